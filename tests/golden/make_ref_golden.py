"""Generates tests/golden/ref_*.npz from the REFERENCE'S OWN SOURCES.

oracle/_ref/libref.so (oracle/ref_recipe.mk) is /root/reference/src/pose_with_velocity/PoseUKF.cpp,
orientation_estimator/OrientationUKF.cpp and UnscentedKalmanFilter.hpp compiled unmodified against the stand-in
dependency headers of oracle/ref_shim.  In that library the wrapper layers -- time guards and latch, initializeFilter,
the process and measurement models, the process-noise shaping including the shadowed local of PoseUKF.cpp:190 and the
dt^2 of OrientationUKF.cpp:86, checkMeasurment, getRotationRate -- are reference text; ukfom::ukf and the MTK manifold
primitives underneath are the restatement of SURVEY.md App. A (the un-vendored slam/mtk).  The scenarios below aim at
the reference-text layers.  /root/reference exists in the build container only, so the outputs are committed here as
fixtures (scenario parameters are regenerated from slam_pose_estimation_b200.synthetic; a fixture holds the expected
final state) together with the hashes of the reference sources they were produced from.

    python tests/golden/make_ref_golden.py        # rewrites the fixtures (needs /root/reference)
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import parity as P  # noqa: E402
from slam_pose_estimation_b200 import synthetic as syn  # noqa: E402


def dense_spd(n, seed, scale):
    """a dense symmetric positive definite matrix (process noise / covariance with off-diagonal entries)"""
    a = syn.noise(np.arange(n) + 1000 * seed, seed, 12, n)
    return scale * (a @ a.T / n + np.eye(n))


def scenario_ref_pose_quirks(x, B):
    """PoseUKF.cpp:180-196 and UnscentedKalmanFilter.hpp:83-125 line by line"""
    x.set_process_noise(dense_spd(12, 1, 1e-3))          # off-diagonal Q: the two rotated blocks and the rest differ visibly
    x.set_time_bounds(1e-6, 0.5)
    t = syn.T0_US
    x.predict_time(np.full(B, t, np.int64))              # :86-90 first call latches only
    for step_us in (10_000, 0, 1, -3_000, 20_000, 700_000, 15_000):
        # 10 ms; repeated stamp (dt = 0: no-op, no latch); 1 us (<= min_dt: no-op, no latch); backwards (throw, no latch);
        # 20 ms measured from the last latch; 0.7 s (> max_dt: throw AFTER the latch moved, :96-97 then :119); 15 ms
        t_new = t + step_us
        x.predict_time(np.full(B, t_new, np.int64))
        if step_us > 1:
            t = t_new
    for kind in range(9):                                 # the nine integrateMeasurement overloads, :112-173
        z, R = syn.pose_measurement(kind, B, 3 + kind)
        x.update(kind, z, R)
        x.predict_dt(0.004 * (1 + kind % 3))
    acc = 0.05 * syn.noise(np.arange(B), 1, 13, 3)
    x.set_acceleration(acc, dense_spd(3, 2, 2e-3))        # :175-178 stored; :188-193 the shadowing branch from now on
    for k in range(4):
        x.predict_dt(0.01 * (k + 1))                      # Q unrotated and NOT scaled by dt, velocity block 2 acc.cov
        z, R = syn.pose_measurement(4, B, 20 + k)
        x.update(4, z, R)
    nan_acc = acc.copy()
    nan_acc[::2, 1] = np.nan                              # one NaN component: allFinite() false -> the plain branch again
    x.set_acceleration(nan_acc, np.eye(3) * 1e-4)
    x.predict_dt(0.02)
    mu, sg = x.get_state()
    x.initialize(mu, sg)                                  # :40-44 re-initialisation resets the time latch
    x.predict_time(np.full(B, t + 5_000_000, np.int64))   # ... so this only latches (no DT_TOO_LARGE)
    x.predict_time(np.full(B, t + 5_010_000, np.int64))


def scenario_ref_ori_quirks(x, B):
    """OrientationUKF.cpp:12-89: dt^2 noise scaling, stored IMU samples, finite checks, velocity update"""
    Q = dense_spd(13, 3, 1e-6)
    x.set_process_noise(Q)
    t = syn.T0_US
    x.predict_time(np.full(B, t, np.int64))
    for k, step_us in enumerate((1_000, 5_000, 20_000, 1_000, 50_000, 2_000, 2_000, 10_000), start=1):
        gyro, acc = syn.orientation_imu(B, k)
        if k == 3:
            gyro = gyro.copy()
            gyro[::3, 0] = np.inf                         # :55 checkMeasurment throws: the old sample stays
        x.set_rotation_rate(gyro)
        x.set_acceleration(acc)
        t += step_us
        x.predict_time(np.full(B, t, np.int64))           # process noise = dt^2 Q' (:86), not dt Q'
        if k % 2 == 0:
            z, R = syn.orientation_velocity(B, k)
            if k == 6:
                z = z.copy()
                z[1::4, 2] = np.nan                       # :67 rejected
            x.update(9, z, R)


def scenario_ref_pose_c3(x, B):
    P.run_pose_c3(x, B, 120)


def scenario_ref_ori_c1(x, B):
    P.run_ori_c1(x, B, 600, every=50)


def scenario_ref_pose_c5(x, B):
    ts, kinds, mu3 = syn.pose_c5_events(B, 1, 200, dvl_period=7, gps_period=11)
    x.run_events(ts, kinds, mu3, syn.sensor_cov_table())


SCENARIOS = {
    "ref_pose_quirks": (0, 8, scenario_ref_pose_quirks),
    "ref_ori_quirks": (1, 12, scenario_ref_ori_quirks),
    "ref_pose_c3": (0, 8, scenario_ref_pose_c3),
    "ref_ori_c1": (1, 4, scenario_ref_ori_c1),
    "ref_pose_c5": (0, 10, scenario_ref_pose_c5),
}


def make(cls, kind, B, **kw):
    return P.make_pose(cls, B, **kw) if kind == 0 else P.make_ori(cls, B, **kw)


def outputs(x, kind):
    mu, sg = x.get_state()
    out = {"mu": mu, "sigma": sg, "status": x.get_status(), "last_time": x.get_last_time(), "hist": x.get_mean_iter_hist()}
    if kind == 1:
        out["rotation_rate"] = x.get_rotation_rate()
    return out


def main():
    from oracle import oracle_lib as O

    if not O.build_ref() or not os.path.isdir(O.REF_ROOT):
        raise SystemExit("needs /root/reference (the build container)")
    hashes = open(os.path.join(ROOT, "oracle", "_ref", "sources.sha256")).read()
    for name, (kind, B, script) in SCENARIOS.items():
        r = make(O.OracleBatch, kind, B, variant="ref")
        script(r, B)
        out = outputs(r, kind)
        assert np.isfinite(out["sigma"]).all()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind=kind, B=B, source_sha256=np.array(hashes), **out)
        print(f"{name}: kind {kind}, B {B}, status bits {int(np.bitwise_or.reduce(out['status']))}")


if __name__ == "__main__":
    main()
