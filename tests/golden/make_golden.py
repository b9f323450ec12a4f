"""Generates tests/golden/*.npz -- frozen outputs of the CPU oracle on fixed seeded scenarios.

The reference holds no golden vectors for the UKF path and cannot be built here, so these
fixtures pin the ORACLE: every scenario is run through the C++ oracle (oracle/ukf_oracle.hpp) AND through the
independent NumPy / SciPy-LAPACK restatement (oracle/numpy_ukf.py); generation fails unless the two agree to
1e-11.  Inputs are regenerated from slam_pose_estimation_b200.synthetic (counter-hash noise), so a fixture holds
only the scenario parameters and the expected final (mu, sigma, status, last_time).

    python tests/golden/make_golden.py        # rewrites the fixtures
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import parity as P  # noqa: E402
from oracle import numpy_ukf as npu  # noqa: E402
from oracle.oracle_lib import OracleBatch  # noqa: E402
from slam_pose_estimation_b200 import synthetic as syn  # noqa: E402


# ---- scenario scripts: drive any object with the UkfBatch method names --------------------------------
def scenario_pose_c3(x, B):
    P.run_pose_c3(x, B, 40)


def scenario_pose_all_kinds(x, B):
    for k in range(1, 28):
        x.predict_dt(0.002 * (1 + k % 5))
        kind = (k - 1) % 9
        z, R = syn.pose_measurement(kind, B, k)
        x.update(kind, z, R)


def scenario_pose_acceleration(x, B):
    acc = 0.02 * syn.noise(np.arange(B), 1, 13, 3)
    cov = np.diag([1e-4, 2e-4, 3e-4])
    mask = (np.arange(B) % 2).astype(np.uint8)
    x.set_acceleration(acc, cov, mask)
    for k in range(1, 11):
        x.predict_dt(0.01)
        z, R = syn.pose_measurement(4, B, k)
        x.update(4, z, R)


def scenario_ori_stream(x, B):
    P.run_ori_c1(x, B, 60, every=10)


def scenario_ori_c1_10k(x, B):
    P.run_ori_c1(x, B, 10_000)


def scenario_pose_c5_events(x, B):
    """per-filter queues of asynchronous IMU / DVL / GPS samples (BASELINE config 5), two launches"""
    ts, kinds, mu3 = syn.pose_c5_events(B, 1, 24, dvl_period=7, gps_period=11)
    tab = syn.sensor_cov_table()
    h = ts.shape[0] // 2
    x.run_events(ts[:h], kinds[:h], mu3[:h], tab)
    x.run_events(ts[h:], kinds[h:], mu3[h:], tab)


def scenario_pose_gate(x, B):
    """Mahalanobis gate at 16: every third filter receives gross outliers, which must leave it untouched"""
    x.set_mahalanobis_gate(16.0)
    out = np.arange(B) % 3 == 0
    for k in range(1, 5):
        for kind in (8, 4, 1):
            z, R = syn.pose_measurement(kind, B, k)
            z = z.copy()
            if kind != 8:
                z[out] += 500.0
            x.step(syn.DT, kind, z, R)


def scenario_ori_events(x, B):
    """OrientationUKF queues: rotation-rate and acceleration samples stored, velocity updates, idle slots"""
    K = 24
    ts = np.zeros((K, B), np.int64)
    kinds = np.full((K, B), syn.EVENT_IDLE, np.int8)
    mu3 = np.zeros((K, B, 3))
    t = np.full(B, syn.T0_US, np.int64)
    for k in range(K):
        tick = k // 3 + 1
        gyro, acc = syn.orientation_imu(B, tick)
        act = (np.arange(B) + k) % 5 != 4
        kind, z, step = ((11, gyro, 400), (12, acc, 0), (9, syn.orientation_velocity(B, tick)[0], 600))[k % 3]
        t = np.where(act, t + step, t)
        ts[k], kinds[k], mu3[k] = t, np.where(act, kind, syn.EVENT_IDLE), z
    x.run_events(ts, kinds, mu3, syn.sensor_cov_table())


SCENARIOS = {
    "pose_c3": (0, 16, scenario_pose_c3),
    "pose_all_kinds": (0, 9, scenario_pose_all_kinds),
    "pose_acceleration": (0, 6, scenario_pose_acceleration),
    "ori_stream": (1, 4, scenario_ori_stream),
    "ori_c1_10k": (1, 1, scenario_ori_c1_10k),
    "pose_c5_events": (0, 12, scenario_pose_c5_events),
    "pose_gate": (0, 9, scenario_pose_gate),
    "ori_events": (1, 7, scenario_ori_events),
}


def make(cls, kind, B, **kw):
    return P.make_pose(cls, B, **kw) if kind == 0 else P.make_ori(cls, B, **kw)


class NumpyBatch:
    """the NumPy restatement behind the same method names (one Python filter object per batch entry)"""

    def __init__(self, kind, B):
        self.kind, self.B, self.f = kind, B, [None] * B

    def initialize(self, mu, sigma):
        for b in range(self.B):
            self.f[b] = (npu.PoseUKF(mu[b], sigma[b]) if self.kind == 0
                         else npu.OrientationUKF(mu[b], sigma[b], np.inf, np.inf, 0.0))

    def set_process_noise(self, Q):
        for f in self.f:
            f.Q = np.array(Q, float)

    def set_orientation_params(self, tg, ta, lat):
        for f in self.f:
            f.tau_g, f.tau_a = tg, ta
            f.earth = np.array([npu.EARTHW * np.cos(lat), 0.0, npu.EARTHW * np.sin(lat)])

    def predict_dt(self, dt):
        dt = np.broadcast_to(np.asarray(dt, float), (self.B,))
        for b, f in enumerate(self.f):
            f.predict_dt(float(dt[b]))

    def predict_time(self, ts):
        ts = np.broadcast_to(np.asarray(ts, np.int64), (self.B,))
        for b, f in enumerate(self.f):
            f.predict_time(int(ts[b]))

    def update(self, kind, z, R, mask=None):
        for b, f in enumerate(self.f):
            if mask is not None and not mask[b]:
                continue
            Rb = R[b] if np.ndim(R) == 3 else R
            f.update_velocity(z[b], Rb) if kind == 9 else f.update(kind, z[b], Rb)

    def step(self, dt, kind, z, R, mask=None):
        self.predict_dt(dt)
        self.update(kind, z, R, mask)

    def set_acceleration(self, mu, cov=None, mask=None):
        for b, f in enumerate(self.f):
            if mask is not None and not mask[b]:
                continue
            if self.kind == 0:
                f.set_acceleration(mu[b], np.eye(3) if cov is None else cov)
            else:
                f.acc = np.array(mu[b], float)

    def set_rotation_rate(self, mu, cov=None, mask=None):
        for b, f in enumerate(self.f):
            f.gyro = np.array(mu[b], float)

    def set_mahalanobis_gate(self, max_d2):
        for f in self.f:
            f.ukf.accept_max_d2 = float(max_d2)

    def run_events(self, ts, kinds, mu3, cov):
        """the caller loop of the reference, filter by filter: predictionStepFromSampleTime, then integrateMeasurement"""
        K = ts.shape[0]
        for b, f in enumerate(self.f):
            for k in range(K):
                kind = int(kinds[k, b])
                if kind == syn.EVENT_IDLE:
                    continue
                f.predict_time(int(ts[k, b]))
                if kind < 0:
                    continue
                R = cov[kind] if cov.shape == (13, 3, 3) else cov[k, b].reshape(3, 3)
                m = {1: 2, 5: 2, 7: 2, 2: 1, 6: 1}.get(kind, 3)
                if kind == 10:
                    f.set_acceleration(mu3[k, b], R)
                elif kind == 11:
                    f.gyro = np.array(mu3[k, b], float)
                elif kind == 12:
                    f.acc = np.array(mu3[k, b], float)
                elif kind == 9:
                    f.update_velocity(mu3[k, b], R)
                else:
                    f.update(kind, mu3[k, b, :m], R[:m, :m])

    def get_state(self):
        return np.stack([f.ukf.mu for f in self.f]), np.stack([f.ukf.sigma for f in self.f])


def main():
    for name, (kind, B, script) in SCENARIOS.items():
        o = make(OracleBatch, kind, B)
        script(o, B)
        mu, sg = o.get_state()
        if name != "ori_c1_10k":  # the 10k-step stream is too slow for the Python restatement; 60 steps cover it
            n = make(NumpyBatch, kind, B)
            script(n, B)
            em, es = P.assert_parity(kind, (mu, sg), n.get_state(), tol=1e-11, what=f"{name}: oracle vs NumPy restatement")
            print(f"{name}: oracle vs numpy  mu {em:.2e}  sigma {es:.2e}")
        np.savez(os.path.join(HERE, name + ".npz"), kind=kind, B=B, mu=mu, sigma=sg, status=o.get_status(),
                 last_time=o.get_last_time())
        print("wrote", name)


if __name__ == "__main__":
    main()
