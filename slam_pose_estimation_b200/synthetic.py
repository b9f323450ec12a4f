"""Synthetic IMU / DVL / GPS streams for the batched filters (host side, NumPy only).

The reference ships no data and no UKF test; SURVEY.md section 8(d) fixes the
synthetic workloads (configs C1-C5 of BASELINE.json) used by the parity tests,
smoke() and bench.py.  Noise is a counter hash (splitmix64 of seed, filter, step,
channel) -> sum of four exact 32-bit uniforms, so any slice of any stream can be
generated independently and reproducibly on any host.
"""
from __future__ import annotations

import numpy as np

SEED = 20261018
LATITUDE_BREMEN = 0.92698121  # the only site constant in the reference (test_coordinate_projection.cpp:11)

SIGMA_GYRO = 1e-3
SIGMA_ACC = 1e-2
SIGMA_DVL = 1e-2
SIGMA_GPS = 0.5
DT = 1e-3
T0_US = 1_000_000  # non-zero start: t_last == 0 means "unset" (UnscentedKalmanFilter.hpp:86)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def noise(filters: np.ndarray, step: int, channel: int, ncomp: int, seed: int = SEED) -> np.ndarray:
    """Unit-variance, zero-mean noise, shape (len(filters), ncomp): Irwin-Hall(4) of hashed uniforms."""
    f = np.asarray(filters, np.uint64)[:, None]
    c = (np.uint64(channel) * np.uint64(16) + np.arange(ncomp, dtype=np.uint64))[None, :]
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed) ^ _splitmix64(f * np.uint64(0x10001) + np.uint64(step) * np.uint64(0x1000003)) ^ (c << np.uint64(40)))
        a = _splitmix64(key)
        b = _splitmix64(a)
    s = ((a >> np.uint64(32)).astype(np.float64) + (a & np.uint64(0xFFFFFFFF)).astype(np.float64)
         + (b >> np.uint64(32)).astype(np.float64) + (b & np.uint64(0xFFFFFFFF)).astype(np.float64))
    return (s * 2.0 ** -32 - 2.0) * np.sqrt(3.0)


# ---- PoseUKF workloads (C3 / C4) ---------------------------------------------------
POSE_V_TRUE = np.array([1.0, 0.0, 0.0])
POSE_W_TRUE = np.array([0.0, 0.0, 0.05])


def pose_initial(B: int, perturb: bool = False, first: int = 0):
    """mu0 (B,13), sigma0 (B,12,12).  perturb: mu0 drawn per filter from N(0, sigma0) (C4)."""
    mu = np.zeros((B, 13))
    mu[:, 6] = 1.0
    mu[:, 7:10] = POSE_V_TRUE
    mu[:, 10:13] = POSE_W_TRUE
    d = np.array([1.0] * 3 + [0.01] * 3 + [0.1] * 3 + [0.01] * 3)
    sigma = np.broadcast_to(np.diag(d), (B, 12, 12)).copy()
    if perturb:
        idx = np.arange(first, first + B)
        e = noise(idx, 0, 15, 12) * np.sqrt(d)
        mu[:, 0:3] += e[:, 0:3]
        half = 0.5 * e[:, 3:6]
        ang = np.linalg.norm(half, axis=1, keepdims=True)
        q = np.concatenate([np.sinc(ang / np.pi) * half, np.cos(ang)], axis=1)
        mu[:, 3:7] = q
        mu[:, 7:10] += e[:, 6:9]
        mu[:, 10:13] += e[:, 9:12]
    return mu, sigma


def pose_truth(step: int):
    """Planar circle: body v = (1,0,0), yaw rate 0.05 rad/s, start at the origin, t = step*DT."""
    t = step * DT
    w = POSE_W_TRUE[2]
    yaw = w * t
    pos = np.array([np.sin(yaw) / w, (1 - np.cos(yaw)) / w, 0.0])
    return pos, yaw


def pose_measurement(kind: int, B: int, step: int, first: int = 0, r_scale=None):
    """(mu (B,m), cov (m,m)) for measurement `kind` at tick `step` of filters first..first+B."""
    idx = np.arange(first, first + B)
    pos, yaw = pose_truth(step)
    if kind == 8:
        sig, truth, ch = SIGMA_GYRO, POSE_W_TRUE, 1
    elif kind == 4:
        sig, truth, ch = SIGMA_DVL, POSE_V_TRUE, 2
    elif kind == 0:
        sig, truth, ch = SIGMA_GPS, pos, 3
    elif kind == 1:
        sig, truth, ch = SIGMA_GPS, pos[:2], 4
    elif kind == 2:
        sig, truth, ch = SIGMA_GPS, pos[2:], 5
    elif kind == 3:
        sig, truth, ch = 1e-2, np.array([0.0, 0.0, yaw]), 6
    elif kind == 5:
        sig, truth, ch = SIGMA_DVL, POSE_V_TRUE[:2], 7
    elif kind == 6:
        sig, truth, ch = SIGMA_DVL, POSE_V_TRUE[2:], 8
    elif kind == 7:
        sig, truth, ch = SIGMA_DVL, np.array([POSE_V_TRUE[0], POSE_W_TRUE[2]]), 9
    else:
        raise ValueError(kind)
    m = len(truth)
    z = truth[None, :] + sig * noise(idx, step, ch, m)
    cov = np.eye(m) * sig * sig
    if r_scale is not None:
        cov = cov[None] * np.asarray(r_scale, float).reshape(-1, 1, 1)
    return z, cov


def pose_schedule(step: int):
    """C3 schedule: angular velocity every tick, velocity every 10th, position every 100th."""
    kinds = [8]
    if step % 10 == 0:
        kinds.append(4)
    if step % 100 == 0:
        kinds.append(0)
    return kinds


# ---- OrientationUKF workloads (C1 / C2) -----------------------------------------------
ORI_Q = np.diag([1e-6] * 3 + [1e-4] * 3 + [1e-10] * 3 + [1e-8] * 3 + [1e-12])
ORI_TAU = 3600.0
G0 = 9.81


def orientation_initial(B: int):
    mu = np.zeros((B, 14))
    mu[:, 3] = 1.0
    mu[:, 13] = G0
    d = np.array([0.01] * 3 + [0.01] * 3 + [1e-6] * 3 + [1e-4] * 3 + [1e-4])
    sigma = np.broadcast_to(np.diag(d), (B, 13, 13)).copy()
    return mu, sigma


def orientation_imu(B: int, step: int, first: int = 0):
    """gyro (B,3), acc (B,3): body yaw rate 0.05 rad/s, specific force (0,0,g) + noise."""
    idx = np.arange(first, first + B)
    gyro = POSE_W_TRUE[None, :] + SIGMA_GYRO * noise(idx, step, 10, 3)
    acc = np.array([0.0, 0.0, G0])[None, :] + SIGMA_ACC * noise(idx, step, 11, 3)
    return gyro, acc


def orientation_velocity(B: int, step: int, first: int = 0):
    idx = np.arange(first, first + B)
    z = SIGMA_DVL * noise(idx, step, 12, 3)
    return z, np.eye(3) * SIGMA_DVL**2
