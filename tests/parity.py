"""Parity helpers shared by the CPU (emulator) and GPU (C ABI) tests.

TEST INFRASTRUCTURE.  The metrics are the ones SURVEY.md section 8(d) fixes:
  mu:    || mu_a [-] mu_b ||_inf / max(1, ||mu_b||_inf)   (boxminus so q == -q)
  sigma: || S_a - S_b ||_F / || S_b ||_F
Tolerance of the north star: 1e-9 on both.
"""
from __future__ import annotations

import numpy as np

TOL = 1e-9


def _qmul(a, b):
    x1, y1, z1, w1 = a.T
    x2, y2, z2, w2 = b.T
    return np.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                     w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2,
                     w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], axis=1)


def _qlog(q):
    nv = np.linalg.norm(q[:, :3], axis=1)
    nv = np.maximum(nv, 1e-300)
    s = 2.0 * np.arctan(nv / q[:, 3]) / nv
    return q[:, :3] * s[:, None]


def mu_error(kind: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """per-filter manifold distance between states a and b (B x MU)"""
    rot = 3 if kind == 0 else 0
    qa, qb = a[:, rot:rot + 4], b[:, rot:rot + 4]
    qbi = qb * np.array([-1.0, -1.0, -1.0, 1.0])
    dq = np.abs(_qlog(_qmul(qa, qbi))).max(axis=1)
    rest = np.delete(a - b, np.s_[rot:rot + 4], axis=1)
    d = np.maximum(np.abs(rest).max(axis=1), dq)
    return d / np.maximum(1.0, np.abs(b).max(axis=1))


def sigma_error(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    num = np.linalg.norm((a - b).reshape(a.shape[0], -1), axis=1)
    den = np.linalg.norm(b.reshape(b.shape[0], -1), axis=1)
    return num / den


def assert_parity(kind, got, ref, tol=TOL, what=""):
    mu_g, sg_g = got
    mu_r, sg_r = ref
    assert np.isfinite(mu_g).all() and np.isfinite(sg_g).all(), f"{what}: non-finite result"
    em = mu_error(kind, mu_g, mu_r).max()
    es = sigma_error(sg_g, sg_r).max()
    assert em <= tol, f"{what}: state error {em:.3e} > {tol}"
    assert es <= tol, f"{what}: covariance error {es:.3e} > {tol}"
    return em, es


def spd_ok(sigma: np.ndarray) -> bool:
    sym = np.abs(sigma - np.transpose(sigma, (0, 2, 1))).max() == 0.0
    w = np.linalg.eigvalsh(sigma)
    return bool(sym and (w > 0).all())


# ---- scenario scripts: the same calls on any object with the UkfBatch method names ------------

def run_pose_c3(x, B, steps, first=0, start=1, fused=True, r_scale=None):
    """C3 schedule on PoseUKF: predict + angular velocity every tick, velocity every 10th, position every 100th."""
    from slam_pose_estimation_b200 import synthetic as syn
    for k in range(start, start + steps):
        for kind in syn.pose_schedule(k):
            z, R = syn.pose_measurement(kind, B, k, first=first, r_scale=r_scale)
            if kind == 8:
                if fused:
                    x.step(syn.DT, kind, z, R)
                else:
                    x.predict_dt(syn.DT)
                    x.update(kind, z, R)
            else:
                x.update(kind, z, R)


def run_ori_c1(x, B, steps, first=0, start=1, every=100):
    """C1/C2 schedule on OrientationUKF: IMU sample + time-stamped predict every tick, velocity update every 100th."""
    from slam_pose_estimation_b200 import synthetic as syn
    for k in range(start, start + steps):
        gyro, acc = syn.orientation_imu(B, k, first=first)
        x.set_rotation_rate(gyro)
        x.set_acceleration(acc)
        x.predict_time(np.array([syn.T0_US + 1000 * k], np.int64))
        if k % every == 0:
            z, R = syn.orientation_velocity(B, k, first=first)
            x.update(9, z, R)


def make_pose(cls, B, perturb=True, first=0, **kw):
    from slam_pose_estimation_b200 import synthetic as syn
    mu, sg = syn.pose_initial(B, perturb=perturb, first=first)
    x = cls(0, B, **kw)
    x.initialize(mu, sg)
    return x


def make_ori(cls, B, **kw):
    from slam_pose_estimation_b200 import synthetic as syn
    mu, sg = syn.orientation_initial(B)
    x = cls(1, B, **kw)
    x.initialize(mu, sg)
    x.set_process_noise(syn.ORI_Q)
    x.set_orientation_params(syn.ORI_TAU, syn.ORI_TAU, syn.LATITUDE_BREMEN)
    return x
