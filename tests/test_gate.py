"""The `accept` functor slot of ukfom::ukf::update (PoseUKF.cpp:116, OrientationUKF.cpp:69-71).  The reference passes
accept_any_mahalanobis_distance, the engine's default; ukfb_set_mahalanobis_gate(max_d2) gives the thresholded functor:
an outlier leaves state and covariance untouched and sets UKFB_STATUS_MEAS_REJECTED."""
from __future__ import annotations

import numpy as np
import pytest

import parity as P
from oracle.oracle_lib import OracleBatch
from slam_pose_estimation_b200 import synthetic as syn

REJECTED = 64


def check_pose(cls, kw, tol):
    B = 40
    o, e = P.make_pose(OracleBatch, B), P.make_pose(cls, B, **kw)
    out = np.arange(B) % 3 == 0
    for x in (o, e):
        x.set_mahalanobis_gate(16.0)
        for k in range(1, 4):
            for kind in (8, 4, 1, 3):  # angular velocity, velocity, XY position, orientation (the literal path)
                z, R = syn.pose_measurement(kind, B, k)
                z = z.copy()
                if kind != 8:
                    z[out] += 1000.0 if kind != 3 else 2.0  # gross outliers for every third filter
                x.step(syn.DT, kind, z, R)
    P.assert_parity(0, e.get_state(), o.get_state(), tol=tol, what="gated updates")
    st = o.get_status()
    assert (st[out] == REJECTED).all() and not st[~out].any()
    assert np.array_equal(e.get_status(), st)
    # the default accepts everything: the same stream without a gate moves the outlier filters far away
    f = P.make_pose(cls, B, **kw)
    z, R = syn.pose_measurement(4, B, 1)
    f.step(syn.DT, 4, z + 1000.0, R)
    assert not f.get_status().any() and (np.abs(f.get_state()[0][:, 7]) > 100).all()


def check_ori(cls, kw, tol):
    B = 33
    o, e = P.make_ori(OracleBatch, B), P.make_ori(cls, B, **kw)
    out = np.arange(B) % 4 == 1
    for x in (o, e):
        x.set_mahalanobis_gate(25.0)
        for k in range(1, 5):
            gyro, acc = syn.orientation_imu(B, k)
            x.set_rotation_rate(gyro)
            x.set_acceleration(acc)
            x.predict_time(np.array([syn.T0_US + 1000 * k], np.int64))
            z, R = syn.orientation_velocity(B, k)
            z = z.copy()
            z[out] += 50.0
            x.update(9, z, R)
    P.assert_parity(1, e.get_state(), o.get_state(), tol=tol, what="gated orientation updates")
    st = o.get_status()
    assert (st[out] == REJECTED).all() and not st[~out].any()
    assert np.array_equal(e.get_status(), st)


@pytest.mark.parametrize("kernel", ["thread", "fast", "warp"])
def test_emu_pose_gate(kernel):
    from emu_lib import EmuBatch
    check_pose(EmuBatch, dict(kernel=kernel), 1e-12)


@pytest.mark.parametrize("kernel", ["thread", "fast", "warp"])
def test_emu_orientation_gate(kernel):
    from emu_lib import EmuBatch
    check_ori(EmuBatch, dict(kernel=kernel), 1e-12)


@pytest.mark.gpu
def test_gpu_gate():
    from slam_pose_estimation_b200 import UkfBatch
    check_pose(UkfBatch, {}, P.TOL)
    check_ori(UkfBatch, {}, P.TOL)
    g = UkfBatch(0, 4)
    assert g.get_mahalanobis_gate() == np.inf
    g.set_mahalanobis_gate(9.0)
    assert g.get_mahalanobis_gate() == 9.0
