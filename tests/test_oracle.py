"""Pins the CPU oracle (oracle/ukf_oracle.hpp through oracle/build/liboracle.so).

The reference holds no UKF test or golden vector (PARITY UNPINNED), so
the oracle is pinned by (i) an independent NumPy / SciPy-LAPACK restatement, (ii) analytic
known-answer tests, (iii) committed fixtures (tests/golden) that freeze its outputs.
"""
from __future__ import annotations

import numpy as np
import pytest

import parity as P
from oracle import numpy_ukf as npu
from oracle.oracle_lib import OracleBatch
from slam_pose_estimation_b200 import synthetic as syn

TIGHT = 1e-11  # two independent double-precision restatements of the same algorithm


def _np_state(f):
    return f.ukf.mu[None, :].copy(), f.ukf.sigma[None, :, :].copy()


@pytest.mark.parametrize("variant", ["left", "right"])
def test_pose_against_numpy_restatement(variant):
    left = variant == "left"
    mu, sg = syn.pose_initial(3, perturb=True)
    for b in range(3):
        o = OracleBatch(0, 1, variant=variant)
        o.initialize(mu[b:b + 1], sg[b:b + 1])
        f = npu.PoseUKF(mu[b], sg[b], left=left)
        for k in range(1, 13):
            o.predict_dt(0.01 * k)
            f.predict_dt(0.01 * k)
            kind = (k - 1) % 9
            z, R = syn.pose_measurement(kind, 3, k)
            o.update(kind, z[b:b + 1], R)
            f.update(kind, z[b], R)
            P.assert_parity(0, o.get_state(), _np_state(f), tol=TIGHT, what=f"{variant} filter {b} tick {k} kind {kind}")
        assert int(o.get_mean_iter_hist().sum()) == len(f.ukf.passes)
        assert np.array_equal(np.bincount(f.ukf.passes, minlength=8)[:8], o.get_mean_iter_hist())


def test_pose_acceleration_against_numpy():
    mu, sg = syn.pose_initial(1, perturb=True)
    o = OracleBatch(0, 1)
    o.initialize(mu, sg)
    f = npu.PoseUKF(mu[0], sg[0])
    acc, cov = np.array([[0.02, -0.01, 0.03]]), np.diag([1e-4, 2e-4, 3e-4])
    o.set_acceleration(acc, cov)
    f.set_acceleration(acc[0], cov)
    for _ in range(3):
        o.predict_dt(0.05)
        f.predict_dt(0.05)
    P.assert_parity(0, o.get_state(), _np_state(f), tol=TIGHT, what="acceleration branch")
    # the shadowed noise: velocity block grows by 2*acc.cov per predict, unscaled by dt (PoseUKF.cpp:190-191)
    o2 = OracleBatch(0, 1)
    o2.initialize(mu, sg)
    o2.predict_dt(0.05)
    assert o.get_state()[1][0, 6, 6] - sg[0, 6, 6] > 3 * 2e-4 * 0.99


@pytest.mark.parametrize("variant", ["left", "right"])
def test_orientation_against_numpy_restatement(variant):
    mu, sg = syn.orientation_initial(1)
    o = OracleBatch(1, 1, variant=variant)
    o.initialize(mu, sg)
    o.set_process_noise(syn.ORI_Q)
    o.set_orientation_params(syn.ORI_TAU, 2 * syn.ORI_TAU, syn.LATITUDE_BREMEN)
    f = npu.OrientationUKF(mu[0], sg[0], syn.ORI_TAU, 2 * syn.ORI_TAU, syn.LATITUDE_BREMEN, left=variant == "left")
    f.Q = syn.ORI_Q.copy()
    for k in range(1, 16):
        gyro, acc = syn.orientation_imu(1, k)
        o.set_rotation_rate(gyro)
        o.set_acceleration(acc)
        f.gyro, f.acc = gyro[0], acc[0]
        o.predict_time(np.array([syn.T0_US + 1000 * k], np.int64))
        f.predict_time(syn.T0_US + 1000 * k)
        if k % 5 == 0:
            z, R = syn.orientation_velocity(1, k)
            o.update(9, z, R)
            f.update_velocity(z[0], R)
        P.assert_parity(1, o.get_state(), _np_state(f), tol=TIGHT, what=f"{variant} tick {k}")
    assert np.allclose(o.get_rotation_rate()[0], f.rotation_rate(), rtol=0, atol=1e-15)


# ---- analytic known answers -----------------------------------------------------------------------

def test_kat_linear_measurement_is_the_kalman_update():
    """h = selector of a Euclidean block and sigma without orientation coupling: the unscented update with
    these sigma points (unit spread, weights 1/2) equals the closed-form Kalman update."""
    rng = np.random.default_rng(7)
    n = 12
    A = rng.normal(size=(9, 9)) * 0.3
    S9 = A @ A.T + np.eye(9) * 0.05
    idx = [0, 1, 2, 6, 7, 8, 9, 10, 11]
    sigma = np.zeros((n, n))
    sigma[np.ix_(idx, idx)] = S9
    sigma[3:6, 3:6] = np.eye(3) * 1e-4
    mu, _ = syn.pose_initial(1)
    o = OracleBatch(0, 1)
    o.initialize(mu, sigma[None])
    H = np.zeros((3, n))
    H[:, 6:9] = np.eye(3)  # velocity measurement, PoseUKF.cpp:36-40
    R = np.diag([0.01, 0.02, 0.03])
    z = np.array([[1.1, -0.2, 0.05]])
    o.update(4, z, R)
    x = np.concatenate([mu[0, 0:3], np.zeros(3), mu[0, 7:13]])
    Sm = H @ sigma @ H.T + R
    K = sigma @ H.T @ np.linalg.inv(Sm)
    x_new = x + K @ (z[0] - H @ x)
    s_new = sigma - K @ Sm @ K.T
    mu_o, sg_o = o.get_state()
    got = np.concatenate([mu_o[0, 0:3], np.zeros(3), mu_o[0, 7:13]])
    assert np.abs(got - x_new).max() < 1e-12
    assert np.abs(mu_o[0, 3:7] - mu[0, 3:7]).max() < 1e-12  # orientation untouched
    assert np.abs(sg_o[0] - s_new).max() < 1e-12


def test_kat_pure_rotation_and_translation():
    """near-zero covariance: predict reduces to the process model; yaw advances by w*dt, position by R(q) v dt
    (PoseUKF.cpp:75-83, both rotations with the input orientation)."""
    yaw0, w, v, dt = 0.3, 0.05, 1.0, 0.1
    mu = np.zeros((1, 13))
    mu[0, 3:7] = [0, 0, np.sin(yaw0 / 2), np.cos(yaw0 / 2)]
    mu[0, 7] = v
    mu[0, 12] = w
    o = OracleBatch(0, 1)
    o.initialize(mu, np.eye(12)[None] * 1e-24)
    o.set_process_noise(np.zeros((12, 12)))
    o.predict_dt(dt)
    m, s = o.get_state()
    yaw1 = yaw0 + w * dt
    assert np.abs(m[0, 3:7] - [0, 0, np.sin(yaw1 / 2), np.cos(yaw1 / 2)]).max() < 1e-12
    assert np.abs(m[0, 0:3] - [v * dt * np.cos(yaw0), v * dt * np.sin(yaw0), 0]).max() < 1e-12
    assert np.abs(m[0, 7:13] - mu[0, 7:13]).max() == 0.0


def test_kat_static_predict_adds_rotated_noise():
    """v = w = 0: the state is a fixed point of the model, so sigma' = sigma + dt * blockrot(Q) (PoseUKF.cpp:182-186)
    and for OrientationUKF the scale is dt^2 (OrientationUKF.cpp:86)."""
    mu = np.zeros((1, 13))
    q = np.array([0.1, -0.2, 0.3, 0.9])
    q /= np.linalg.norm(q)
    mu[0, 3:7] = q
    sigma = np.diag(np.linspace(0.5, 1.5, 12))[None] * 1e-3
    sigma[0, 6:, 6:] *= 1e-30  # (nearly) certain zero velocities: no pose/velocity coupling builds up
    Q = np.diag(np.linspace(1, 2, 12)) * 1e-3
    Q[0, 1] = Q[1, 0] = 2e-4
    Q[3, 5] = Q[5, 3] = -1e-4
    o = OracleBatch(0, 1)
    o.initialize(mu, sigma)
    o.set_process_noise(Q)
    dt = 0.25
    o.predict_dt(dt)
    Rm = npu.qmat(q)
    Qr = Q.copy()
    Qr[0:3, 0:3] = Rm @ Q[0:3, 0:3] @ Rm.T
    Qr[3:6, 3:6] = Rm @ Q[3:6, 3:6] @ Rm.T
    want = sigma[0] + dt * Qr
    got = o.get_state()[1][0]
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()  # log(exp()) round trip of a 0.03 rad spread


def test_kat_quaternion_sign_is_irrelevant():
    """boxminus uses atan(nv / w): q and -q are the same orientation."""
    mu, sg = syn.pose_initial(2, perturb=True)
    mu[1] = mu[0]
    mu[1, 3:7] *= -1.0
    sg[1] = sg[0]
    o = OracleBatch(0, 2)
    o.initialize(mu, sg)
    z, R = syn.pose_measurement(3, 1, 4)
    o.predict_dt(0.1)
    o.update(3, np.repeat(z, 2, axis=0), R)
    m, s = o.get_state()
    assert P.mu_error(0, m[0:1], m[1:2]).max() < 1e-13
    assert P.sigma_error(s[0:1], s[1:2]).max() < 1e-12


# ---- shell behaviour (UnscentedKalmanFilter.hpp) ---------------------------------------------------------

def test_time_guards_and_latching():
    o = P.make_pose(OracleBatch, 1)
    f0 = o.get_state()
    o.predict_time(np.array([syn.T0_US], np.int64))  # first call latches only (:86-90)
    assert np.array_equal(o.get_state()[0], f0[0]) and o.get_last_time()[0] == syn.T0_US
    o.predict_time(np.array([syn.T0_US], np.int64))  # dt = 0: no-op, no latch
    assert np.array_equal(o.get_state()[0], f0[0]) and o.get_status()[0] == 0
    o.predict_time(np.array([syn.T0_US - 5], np.int64))  # negative: "throws", time not latched
    assert o.get_status()[0] == 1 and o.get_last_time()[0] == syn.T0_US
    o.clear_status()
    o.set_time_bounds(1e-9, 0.5)
    o.predict_time(np.array([syn.T0_US + 2_000_000], np.int64))  # too large: throws AFTER latching (:96-97,119)
    assert o.get_status()[0] == 2 and o.get_last_time()[0] == syn.T0_US + 2_000_000
    assert np.array_equal(o.get_state()[0], f0[0])
    o.predict_time(np.array([syn.T0_US + 2_001_000], np.int64))
    assert not np.array_equal(o.get_state()[0], f0[0])


def test_orientation_nonfinite_measurements_are_rejected():
    o = P.make_ori(OracleBatch, 2)
    before = o.get_state()
    z = np.array([[0.0, np.nan, 0.0], [0.1, 0.0, -0.1]])
    o.update(9, z, np.eye(3) * 1e-4)
    after = o.get_state()
    assert o.get_status().tolist() == [4, 0]
    assert np.array_equal(after[0][0], before[0][0]) and not np.array_equal(after[0][1], before[0][1])


def test_not_spd_is_flagged_and_skipped():
    mu, sg = syn.pose_initial(2)
    sg[1, 4, 4] = -1.0
    o = OracleBatch(0, 2)
    o.initialize(mu, sg)
    o.predict_dt(0.01)
    assert o.get_status().tolist() == [0, 8]
    assert np.array_equal(o.get_state()[0][1], mu[1])


def test_invariants_after_a_stream():
    B = 16
    o = P.make_pose(OracleBatch, B)
    P.run_pose_c3(o, B, 100)
    mu, sg = o.get_state()
    assert np.abs(np.linalg.norm(mu[:, 3:7], axis=1) - 1).max() < 1e-12
    w = np.linalg.eigvalsh(sg)
    assert (w > 0).all()
    assert not o.get_status().any()
