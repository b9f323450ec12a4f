"""Randomised parity: random states, random dense SPD covariances over six orders of magnitude (orientation spreads up
to ~0.7 rad, so manifold means need several passes and some lanes leave the polynomial / guard ranges of the fast
kernels), random time steps, process noise and measurement kinds.  CPU: the kernels' source under tests/simt_emu against
the oracle; GPU: the C ABI against the oracle.  Seeds are fixed; every case is reproducible."""
from __future__ import annotations

import numpy as np
import pytest

import parity as P
from oracle.oracle_lib import OracleBatch

POSE_M = {0: 3, 1: 2, 2: 1, 3: 3, 4: 3, 5: 2, 6: 1, 7: 2, 8: 3}


def random_spd(rng, B, n, lo=-6.0, hi=-0.3):
    A = rng.normal(size=(B, n, n))
    S = A @ np.transpose(A, (0, 2, 1)) / n + 0.05 * np.eye(n)
    d = (10.0 ** rng.uniform(lo, hi, size=(B, n))) ** 0.5  # per-component standard deviations
    return S * d[:, :, None] * d[:, None, :]


def random_quat(rng, B):
    q = rng.normal(size=(B, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def run_pose(cls, kw, seed, B=48, rounds=4):
    rng = np.random.default_rng(seed)
    mu = np.zeros((B, 13))
    mu[:, 0:3] = rng.normal(size=(B, 3)) * 10
    mu[:, 3:7] = random_quat(rng, B)
    mu[:, 7:10] = rng.normal(size=(B, 3)) * 2
    mu[:, 10:13] = rng.normal(size=(B, 3)) * 0.5
    sg = random_spd(rng, B, 12)
    Q = random_spd(rng, B, 12, -8.0, -3.0)
    x = cls(0, B, **kw)
    x.initialize(mu, sg)
    x.set_process_noise(Q)
    acc = rng.normal(size=(B, 3)) * 0.1
    acc[rng.random(B) < 0.5] = np.nan  # half the filters without a stored acceleration
    x.set_acceleration(acc, np.eye(3) * 1e-3)
    for r in range(rounds):
        dt = 10.0 ** rng.uniform(-4, -1, size=B)
        kinds = rng.integers(0, 9, size=B).astype(np.int8)
        kinds[rng.random(B) < 0.1] = -1
        z = np.zeros((B, 3))
        R = np.tile(np.eye(3), (B, 1, 1))
        st = x.get_state()[0]
        for b in range(B):
            k = int(kinds[b])
            if k < 0:
                continue
            m = POSE_M[k]
            truth = {0: st[b, 0:3], 1: st[b, 0:2], 2: st[b, 2:3], 3: np.zeros(3), 4: st[b, 7:10], 5: st[b, 7:9], 6: st[b, 9:10],
                     7: st[b, [7, 12]], 8: st[b, 10:13]}[k]
            z[b, :m] = truth + rng.normal(size=m) * 0.05
            Rb = rng.normal(size=(m, m))
            R[b, :m, :m] = Rb @ Rb.T * 1e-3 + np.eye(m) * 10.0 ** rng.uniform(-5, -1)
        x.predict_dt(dt)
        x.update_mixed(kinds, z, R)
    return x


def run_ori(cls, kw, seed, B=48, rounds=5):
    rng = np.random.default_rng(1000 + seed)
    mu = np.zeros((B, 14))
    mu[:, 0:4] = random_quat(rng, B)
    mu[:, 4:7] = rng.normal(size=(B, 3))
    mu[:, 7:10] = rng.normal(size=(B, 3)) * 1e-2
    mu[:, 10:13] = rng.normal(size=(B, 3)) * 1e-1
    mu[:, 13] = 9.81 + rng.normal(size=B) * 0.01
    sg = random_spd(rng, B, 13)
    Q = random_spd(rng, B, 13, -9.0, -4.0)
    x = cls(1, B, **kw)
    x.set_orientation_params(float(10.0 ** rng.uniform(0, 3)), float(10.0 ** rng.uniform(0, 3)), float(rng.uniform(-1.5, 1.5)))
    x.initialize(mu, sg)
    x.set_process_noise(Q)
    for r in range(rounds):
        x.set_rotation_rate(rng.normal(size=(B, 3)) * 0.3)
        x.set_acceleration(rng.normal(size=(B, 3)) + np.array([0, 0, 9.81]))
        x.predict_dt(10.0 ** rng.uniform(-4, -1.3, size=B))
        if r % 2 == 1:
            Rb = rng.normal(size=(B, 3, 3))
            R = Rb @ np.transpose(Rb, (0, 2, 1)) * 1e-3 + np.eye(3) * 1e-3
            mask = (rng.random(B) < 0.8).astype(np.uint8)
            x.update(9, rng.normal(size=(B, 3)), R, mask)
    return x


def compare(kind, got, ref, tol):
    assert np.array_equal(got.get_status(), ref.get_status())
    ok = ref.get_status() == 0  # a flagged filter (e.g. a covariance that lost definiteness) stops being comparable
    assert ok.sum() > 0.8 * ok.size
    mg, sg = got.get_state()
    mr, sr = ref.get_state()
    P.assert_parity(kind, (mg[ok], sg[ok]), (mr[ok], sr[ok]), tol=tol, what=f"random filters kind {kind}")
    assert np.array_equal(got.get_mean_iter_hist(), ref.get_mean_iter_hist())
    return ref.get_mean_iter_hist()


@pytest.mark.parametrize("kernel", ["thread", "fast"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_emu_random_pose(kernel, seed):
    from emu_lib import EmuBatch
    hist = compare(0, run_pose(EmuBatch, dict(kernel=kernel), seed), run_pose(OracleBatch, {}, seed), 1e-9)
    assert hist[2:].sum() > 0  # some means needed more than one pass


@pytest.mark.parametrize("kernel", ["thread", "fast"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_emu_random_orientation(kernel, seed):
    from emu_lib import EmuBatch
    hist = compare(1, run_ori(EmuBatch, dict(kernel=kernel), seed), run_ori(OracleBatch, {}, seed), 1e-9)
    assert hist[2:].sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_gpu_random_pose(seed):
    from slam_pose_estimation_b200 import UkfBatch
    compare(0, run_pose(UkfBatch, {}, seed, B=200), run_pose(OracleBatch, {}, seed, B=200), P.TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_gpu_random_orientation(seed):
    from slam_pose_estimation_b200 import UkfBatch
    compare(1, run_ori(UkfBatch, {}, seed, B=200), run_ori(OracleBatch, {}, seed, B=200), P.TOL)
