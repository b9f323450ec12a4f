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


# ---- asynchronous sensor queues (C5) ---------------------------------------------------
EVENT_IDLE = -2
EVENT_KIND_COUNT = 13


def sensor_cov_table():
    """cov[kind] (13,3,3): one covariance per sensor kind, leading m x m block used (identity elsewhere)."""
    tab = np.tile(np.eye(3), (EVENT_KIND_COUNT, 1, 1))
    for kind, sig in ((8, SIGMA_GYRO), (4, SIGMA_DVL), (5, SIGMA_DVL), (6, SIGMA_DVL), (7, SIGMA_DVL), (0, SIGMA_GPS),
                      (1, SIGMA_GPS), (2, SIGMA_GPS), (3, 1e-2), (9, SIGMA_DVL), (10, SIGMA_ACC), (11, SIGMA_GYRO),
                      (12, SIGMA_ACC)):
        tab[kind] = np.eye(3) * sig * sig
    return tab


def pack_events(ts_c, kinds_c, mu_c, valid):
    """Per-filter queues from candidate samples: ts_c, kinds_c, valid: (C, B) in per-filter time order; mu_c: (C, B, 3).
    Returns slot-major (K, B) arrays, K = the longest queue, shorter queues padded with EVENT_IDLE."""
    C_, B = valid.shape
    pos = np.cumsum(valid, axis=0) - 1
    K = int(pos[-1].max()) + 1 if C_ else 0
    ts = np.zeros((K, B), np.int64)
    kinds = np.full((K, B), EVENT_IDLE, np.int8)
    mu3 = np.zeros((K, B, 3))
    c, b = np.nonzero(valid)
    k = pos[c, b]
    ts[k, b] = ts_c[c, b]
    kinds[k, b] = kinds_c[c, b]
    mu3[k, b] = mu_c[c, b]
    return ts, kinds, mu3


def pose_c5_events(B: int, tick0: int, n_ticks: int, first: int = 0, dvl_period: int = 100, gps_period: int = 1000):
    """C5 (BASELINE.json config 5): per-filter queues of asynchronous samples over ticks tick0 .. tick0+n_ticks-1.
    IMU 1 kHz -> AngularVelocityMeasurement (kind 8) at t_k; DVL 10 Hz with a per-filter phase -> VelocityMeasurement
    (kind 4) at t_k + 300 us; GPS 1 Hz with a per-filter phase -> XYMeasurement (kind 1, nav-plane coordinates as
    GeographicProjection::worldToNav produces them) at t_k + 600 us.  Returns ts, kinds (K,B), mu3 (K,B,3)."""
    idx = np.arange(first, first + B)
    ph_dvl = (_splitmix64(idx.astype(np.uint64) * np.uint64(3) + np.uint64(SEED)) % np.uint64(dvl_period)).astype(np.int64)
    ph_gps = (_splitmix64(idx.astype(np.uint64) * np.uint64(5) + np.uint64(SEED)) % np.uint64(gps_period)).astype(np.int64)
    ts_c = np.zeros((n_ticks * 3, B), np.int64)
    kinds_c = np.zeros((n_ticks * 3, B), np.int8)
    valid = np.zeros((n_ticks * 3, B), bool)
    mu_c = np.zeros((n_ticks * 3, B, 3))
    for j in range(n_ticks):
        k = tick0 + j
        t = T0_US + 1000 * k
        ts_c[3 * j], kinds_c[3 * j], valid[3 * j] = t, 8, True
        mu_c[3 * j] = pose_measurement(8, B, k, first=first)[0]
        has_dvl = (k + ph_dvl) % dvl_period == 0
        ts_c[3 * j + 1], kinds_c[3 * j + 1], valid[3 * j + 1] = t + 300, 4, has_dvl
        if has_dvl.any():
            mu_c[3 * j + 1] = pose_measurement(4, B, k, first=first)[0]
        has_gps = (k + ph_gps) % gps_period == 0
        ts_c[3 * j + 2], kinds_c[3 * j + 2], valid[3 * j + 2] = t + 600, 1, has_gps
        if has_gps.any():
            mu_c[3 * j + 2, :, :2] = pose_measurement(1, B, k, first=first)[0]
    return pack_events(ts_c, kinds_c, mu_c, valid)
