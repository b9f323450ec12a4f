"""The product's device code (slam_pose_estimation_b200/csrc/ukf_device.cuh), compiled for the host by
tests/simt_emu (every CUDA thread a host thread, warp shuffles and the FP64 mma tile emulated), against the CPU
oracle.  This runs in the GPU-less container and exercises the kernel's shared-memory maps, lane ownership, the
register Cholesky with shuffled pivot rows and the DMMA fragment layout; the real parity gate is
tests/test_gpu_parity.py on a B200.  Small batches only: the emulation is slow."""
from __future__ import annotations

import numpy as np
import pytest

import parity as P
from emu_lib import EmuBatch
from oracle.oracle_lib import OracleBatch
from slam_pose_estimation_b200 import synthetic as syn

TOL = 1e-12  # same algorithm, libm and the polynomial exp/log kernels differ in the last ulps


@pytest.mark.parametrize("G", [4, 8, 16])
def test_pose_stream(G):
    B = 2 * G + 3  # two full groups and a ragged tail, spread over two emulated warps
    o, e = P.make_pose(OracleBatch, B), P.make_pose(EmuBatch, B, G=G)
    P.run_pose_c3(o, B, 10)
    P.run_pose_c3(e, B, 10)
    P.assert_parity(0, e.get_state(), o.get_state(), tol=TOL, what=f"emulated pose stream G={G}")
    assert np.array_equal(e.get_mean_iter_hist(), o.get_mean_iter_hist())
    assert not e.get_status().any()


def test_pose_every_measurement_kind():
    B = 5
    o, e = P.make_pose(OracleBatch, B), P.make_pose(EmuBatch, B, G=8)
    for kind in range(9):
        z, R = syn.pose_measurement(kind, B, kind + 1)
        for x in (o, e):
            x.predict_dt(0.02)
            x.update(kind, z, R)
    P.assert_parity(0, e.get_state(), o.get_state(), tol=TOL, what="emulated update kinds")


def test_pose_acceleration_mask_and_guards():
    B = 6
    o, e = P.make_pose(OracleBatch, B), P.make_pose(EmuBatch, B, G=4)
    acc = 0.01 * syn.noise(np.arange(B), 1, 13, 3)
    mask = (np.arange(B) % 2).astype(np.uint8)
    dt = np.array([-1.0, 0.0, 0.01, 0.02, 5.0, 0.03])
    z, R = syn.pose_measurement(4, B, 2)
    for x in (o, e):
        x.set_time_bounds(1e-9, 1.0)
        x.set_acceleration(acc, np.eye(3) * 1e-4, mask)
        x.predict_dt(dt)
        x.update(4, z, R, mask)
    assert np.array_equal(e.get_status(), o.get_status())
    P.assert_parity(0, e.get_state(), o.get_state(), tol=TOL, what="emulated acceleration / mask / guards")


def test_not_spd_leaves_the_filter_untouched():
    mu, sg = syn.pose_initial(3)
    sg[1, 7, 7] = -1.0
    o, e = OracleBatch(0, 3), EmuBatch(0, 3, G=4)
    z, R = syn.pose_measurement(8, 3, 1)
    for x in (o, e):
        x.initialize(mu, sg)
        x.step(0.01, 8, z, R)
    assert e.get_status().tolist() == o.get_status().tolist() == [0, 8, 0]
    assert np.array_equal(e.get_state()[0][1], mu[1]) and np.array_equal(np.tril(e.get_state()[1][1]), np.tril(sg[1]))
    P.assert_parity(0, (e.get_state()[0][[0, 2]], e.get_state()[1][[0, 2]]),
                    (o.get_state()[0][[0, 2]], o.get_state()[1][[0, 2]]), tol=TOL, what="neighbours of a non-SPD filter")


def test_orientation_stream():
    B = 3
    o, e = P.make_ori(OracleBatch, B), P.make_ori(EmuBatch, B, G=8)
    P.run_ori_c1(o, B, 12, every=4)
    P.run_ori_c1(e, B, 12, every=4)
    P.assert_parity(1, e.get_state(), o.get_state(), tol=TOL, what="emulated orientation stream")
    assert np.array_equal(e.t_last, o.get_last_time())
