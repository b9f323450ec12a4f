"""BodyStateMeasurement (pose_with_velocity/BodyStateMeasurement.hpp:12-41): RigidBodyState records in and out of a
PoseUKF batch.  The oracle functions restate the two reference functions; the engine's pack / unpack kernels must
reproduce them (copies bit for bit, the rotated velocity to rounding)."""
from __future__ import annotations

import numpy as np
import pytest

import parity as P
from oracle import oracle_lib as O
from slam_pose_estimation_b200 import synthetic as syn


def make_rbs(B, seed=3):
    rng = np.random.default_rng(seed)
    rbs = np.zeros((B, 49))
    rbs[:, 0:3] = rng.normal(size=(B, 3)) * 10
    q = rng.normal(size=(B, 4))
    rbs[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    rbs[:, 7:10] = rng.normal(size=(B, 3))
    rbs[:, 10:13] = rng.normal(size=(B, 3)) * 0.1
    for blk, scale in enumerate((1.0, 0.01, 0.1, 0.01)):
        a = rng.normal(size=(B, 3, 3))
        rbs[:, 13 + blk * 9:22 + blk * 9] = ((a @ a.transpose(0, 2, 1) + 3 * np.eye(3)) * scale).reshape(B, 9)
    return rbs


def test_oracle_body_state_functions():
    rbs = make_rbs(5)
    mu, sg = O.from_body_states(rbs)
    assert np.array_equal(mu, rbs[:, :13])  # velocity taken as is (BodyStateMeasurement.hpp:18)
    for blk in range(4):
        assert np.array_equal(sg[:, 3 * blk:3 * blk + 3, 3 * blk:3 * blk + 3].reshape(5, 9), rbs[:, 13 + 9 * blk:22 + 9 * blk])
    off = sg.copy()
    for blk in range(4):
        off[:, 3 * blk:3 * blk + 3, 3 * blk:3 * blk + 3] = 0
    assert not off.any()  # setZero() outside the four blocks (:21)
    out = O.to_body_states(mu, sg)
    # velocity leaves rotated into the navigation frame (:32): R(q) v with an independent rotation matrix
    x, y, z, w = mu[:, 3:7].T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                  2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                  2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1).reshape(-1, 3, 3)
    assert np.abs(out[:, 7:10] - np.einsum("bij,bj->bi", R, mu[:, 7:10])).max() < 1e-14
    keep = np.r_[0:7, 10:49]
    assert np.array_equal(out[:, keep], rbs[:, keep])  # the covariance blocks are not rotated (:35-38)


@pytest.mark.gpu
def test_gpu_body_states_round_trip_and_filtering():
    from oracle.oracle_lib import OracleBatch
    from slam_pose_estimation_b200 import UkfBatch
    B = 70
    rbs = make_rbs(B)
    g = UkfBatch(0, B)
    g.initialize_from_body_states(rbs)
    mu, sg = O.from_body_states(rbs)
    gm, gs = g.get_state()
    assert np.array_equal(gm, mu) and np.array_equal(gs, sg)
    assert not g.get_last_time().any()
    ref = O.to_body_states(mu, sg)
    got = g.get_body_states()
    keep = np.r_[0:7, 10:49]
    assert np.array_equal(got[:, keep], ref[:, keep])
    assert np.abs(got[:, 7:10] - ref[:, 7:10]).max() < 1e-14
    # a few filter steps from that state, then out again
    o = OracleBatch(0, B)
    o.initialize(mu, sg)
    for k in range(1, 4):
        z, R = syn.pose_measurement(8, B, k)
        g.step(syn.DT, 8, z, R)
        o.step(syn.DT, 8, z, R)
    om, os_ = o.get_state()
    ref = O.to_body_states(om, os_)
    got = g.get_body_states()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-12
    with pytest.raises(Exception):
        UkfBatch(1, 4).initialize_from_body_states(np.zeros((4, 49)))
