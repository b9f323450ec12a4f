"""The two lane-per-filter kernels -- ukf_thread.cuh (literal sigma-point sequence) and ukf_pose_fast.cuh (the
structure-exploiting PoseUKF kernel, with its literal fallbacks) -- compiled for the host by tests/simt_emu and
compared with the CPU oracle in the GPU-less container.  The GPU gate is tests/test_gpu_parity.py."""
from __future__ import annotations

import numpy as np
import pytest

import parity as P
from emu_lib import EmuBatch
from oracle.oracle_lib import OracleBatch
from slam_pose_estimation_b200 import synthetic as syn

TOL = 1e-12
KERNELS = ["thread", "fast"]


@pytest.mark.parametrize("kernel", KERNELS)
def test_pose_stream(kernel):
    B = 37  # one full tile and a ragged one
    o, e = P.make_pose(OracleBatch, B), P.make_pose(EmuBatch, B, kernel=kernel)
    P.run_pose_c3(o, B, 12)
    P.run_pose_c3(e, B, 12)
    P.assert_parity(0, e.get_state(), o.get_state(), tol=TOL, what=f"{kernel} pose stream")
    assert np.array_equal(e.get_mean_iter_hist(), o.get_mean_iter_hist())
    assert not e.get_status().any()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("fused", [False, True])
def test_pose_every_measurement_kind(kernel, fused):
    B = 5
    o, e = P.make_pose(OracleBatch, B), P.make_pose(EmuBatch, B, kernel=kernel)
    for rnd in range(2):
        for kind in range(9):
            z, R = syn.pose_measurement(kind, B, kind + 1 + 9 * rnd)
            for x in (o, e):
                if fused:
                    x.step(0.02, kind, z, R)
                else:
                    x.predict_dt(0.02)
                    x.update(kind, z, R)
    P.assert_parity(0, e.get_state(), o.get_state(), tol=TOL, what=f"{kernel} update kinds")
    assert np.array_equal(e.get_mean_iter_hist(), o.get_mean_iter_hist())


@pytest.mark.parametrize("kernel", KERNELS)
def test_pose_acceleration_mask_and_guards(kernel):
    B = 6
    o, e = P.make_pose(OracleBatch, B), P.make_pose(EmuBatch, B, kernel=kernel)
    acc = 0.01 * syn.noise(np.arange(B), 1, 13, 3)
    mask = (np.arange(B) % 2).astype(np.uint8)
    dt = np.array([-1.0, 0.0, 0.01, 0.02, 5.0, 0.03])
    z, R = syn.pose_measurement(4, B, 2)
    for x in (o, e):
        x.set_time_bounds(1e-9, 1.0)
        x.set_acceleration(acc, np.eye(3) * 1e-4, mask)
        x.predict_dt(dt)
        x.update(4, z, R, mask)
        x.step(dt, 0, *syn.pose_measurement(0, B, 3), mask)
    assert np.array_equal(e.get_status(), o.get_status())
    P.assert_parity(0, e.get_state(), o.get_state(), tol=TOL, what=f"{kernel} acceleration / mask / guards")


@pytest.mark.parametrize("kernel", KERNELS)
def test_not_spd_leaves_the_filter_untouched(kernel):
    mu, sg = syn.pose_initial(3)
    sg[1, 7, 7] = -1.0
    z, R = syn.pose_measurement(8, 3, 1)
    for fused in (True, False):
        o, e = OracleBatch(0, 3), EmuBatch(0, 3, kernel=kernel)
        for x in (o, e):
            x.initialize(mu, sg)
            if fused:
                x.step(0.01, 8, z, R)
            else:
                x.update(8, z, R)
        assert e.get_status().tolist() == o.get_status().tolist() == [0, 8, 0]
        assert np.array_equal(e.get_state()[0][1], mu[1]) and np.array_equal(np.tril(e.get_state()[1][1]), np.tril(sg[1]))
        P.assert_parity(0, (e.get_state()[0][[0, 2]], e.get_state()[1][[0, 2]]),
                        (o.get_state()[0][[0, 2]], o.get_state()[1][[0, 2]]), tol=TOL, what="neighbours of a non-SPD filter")


@pytest.mark.parametrize("kernel", KERNELS)
def test_pose_large_angles_take_the_literal_expressions(kernel):
    """Orientation spread and rates far outside the ranges of the short polynomials (0.58 rad) and an orientation
    column of the covariance factor beyond the selector-update guard (|L_ori[:, j]| next to pi).  The fast kernel redoes the COLUMNS whose sigma points leave the range with
    its any-angle exp / log (no literal predict any more), hands the lane behind the update guard to the literal code,
    and still agrees with the oracle."""
    B = 4
    mu, sg = syn.pose_initial(B)
    sg[0, 3:6, 3:6] *= 150.0  # sqrt(1.5) rad orientation sigma: exp and log leave the polynomial range
    sg[1, 5, 5] = 3.13**2  # a factor column at the branch cut of the SO(3) log: selector-update guard
    mu[2, 10:13] = [3.0, -40.0, 25.0]  # 47 rad/s: |w| dt = 0.94 rad with dt = 0.02 stays fast; with 0.05 it does not
    o, e = OracleBatch(0, B), EmuBatch(0, B, kernel=kernel)
    before = e.fallbacks()
    for x in (o, e):
        x.initialize(mu, sg)
        for k, kind in enumerate([8, 4, 0, 7]):
            z, R = syn.pose_measurement(kind, B, k + 1)
            x.step(0.05, kind, z, R)
    assert np.array_equal(e.get_status(), o.get_status())
    P.assert_parity(0, e.get_state(), o.get_state(), tol=1e-10, what=f"{kernel} large angles")
    if kernel == "fast":
        fb = e.fallbacks() - before
        assert fb[0] == 0 and fb[1] > 0, f"literal predict / update calls {fb}"
        assert fb[3] > 0 and fb[4] > 0, f"the any-angle columns were not exercised: {fb}"
        assert fb[:3].sum() < 3 * 4 * B, "every lane fell back: the fast path was not exercised"
    assert np.array_equal(e.get_mean_iter_hist(), o.get_mean_iter_hist())


@pytest.mark.parametrize("sig", [(0.7, 0.7, 0.7), (1.0, 1.0, 1.0), (1.6, 1.6, 1.6), (0.1, 0.1, 2.9), (2.0, 2.0, 2.0), (3.0, 3.0, 3.0),
                                 (0.1, 0.1, 3.05)])
def test_pose_wide_orientation_uncertainty_stays_in_the_fast_kernel(sig):
    """A filter that barely knows its attitude (or, last case, its heading: the usual start-up state): the sigma points
    of the orientation columns are radians apart.  The fast kernel keeps such filters -- no literal predict / update /
    apply_delta call -- by redoing those columns with its any-angle exp / log, and matches the oracle, the number of
    mean passes included.  The last three have trace(Sigma_ori) >= 9, beyond the hot path's guard: the out-of-line code
    looks at the columns of the factor themselves (each below pi) and still runs the structured update."""
    B = 8
    mu, sg = syn.pose_initial(B, perturb=True)
    sg[:, 3:6, 3:6] = np.diag(np.square(sig))
    o, e = OracleBatch(0, B), EmuBatch(0, B, kernel="fast")
    before = e.fallbacks()
    for x in (o, e):
        x.initialize(mu, sg)
        for k, kind in enumerate([8, 8, 4, 8, 0, 8], start=1):
            z, R = syn.pose_measurement(kind, B, k)
            x.step(0.01, kind, z, R)
    assert np.array_equal(e.get_status(), o.get_status()) and not o.get_status().any()
    P.assert_parity(0, e.get_state(), o.get_state(), tol=1e-11, what=f"wide orientation sigma {sig}")
    assert np.array_equal(e.get_mean_iter_hist(), o.get_mean_iter_hist())
    fb = e.fallbacks() - before
    assert not fb[:3].any(), f"literal fallbacks were taken: {fb}"
    assert fb[3] > 0 and fb[4] > 0, f"the any-angle columns were not exercised: {fb}"


def test_pose_heading_uncertainty_at_the_branch_cut_runs_the_literal_update():
    """A factor column within 0.04 rad of pi: (mu [+] L_j) [-] mu = L_j is no longer safe to assume, the literal update
    (which forms the sigma points) serves the filter and matches the oracle."""
    B = 8
    mu, sg = syn.pose_initial(B, perturb=True)
    sg[:, 3:6, 3:6] = np.diag(np.square((0.1, 0.1, 3.12)))
    o, e = OracleBatch(0, B), EmuBatch(0, B, kernel="fast")
    before = e.fallbacks()
    for x in (o, e):
        x.initialize(mu, sg)
        for k in (1, 2):
            z, R = syn.pose_measurement(8, B, k)
            x.step(0.01, 8, z, R)
    assert np.array_equal(e.get_status(), o.get_status())
    P.assert_parity(0, e.get_state(), o.get_state(), tol=1e-11, what="heading sigma 3.12 rad")
    fb = e.fallbacks() - before
    assert fb[1] == 2 * B and fb[0] == 0, f"expected the literal update, the structured predict: {fb}"


@pytest.mark.parametrize("kernel", KERNELS)
def test_pose_mixed_kinds_per_filter_noise_and_time_mode(kernel):
    B = 9
    o, e = P.make_pose(OracleBatch, B), P.make_pose(EmuBatch, B, kernel=kernel)
    rng = np.random.default_rng(5)
    A = rng.normal(size=(B, 12, 12)) * 0.01
    Q = A @ np.transpose(A, (0, 2, 1)) + 1e-5 * np.eye(12)
    kinds = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8], np.int8)
    for x in (o, e):
        x.set_process_noise(Q)
        x.predict_time(np.full(B, 1_000_000, np.int64))
        for k in range(1, 4):
            x.predict_time(1_000_000 + 20_000 * k + np.arange(B, dtype=np.int64) * (k == 2))
            z = np.zeros((B, 3))
            R = np.tile(np.eye(3), (B, 1, 1))
            for b in range(B):
                zz, RR = syn.pose_measurement(int(kinds[b]), B, k)
                m = zz.shape[1]
                z[b, :m] = zz[b]
                R[b, :m, :m] = RR if RR.ndim == 2 else RR[b]
            x.update_mixed(np.roll(kinds, k), np.roll(z, k, axis=0), np.roll(R, k, axis=0))
    assert np.array_equal(e.get_status(), o.get_status())
    P.assert_parity(0, e.get_state(), o.get_state(), tol=TOL, what=f"{kernel} mixed kinds")
    assert np.array_equal(e.t_last, o.get_last_time())


@pytest.mark.parametrize("kernel", KERNELS)
def test_orientation_stream(kernel):
    B = 37
    o, e = P.make_ori(OracleBatch, B), P.make_ori(EmuBatch, B, kernel=kernel)
    P.run_ori_c1(o, B, 12, every=4)
    P.run_ori_c1(e, B, 12, every=4)
    P.assert_parity(1, e.get_state(), o.get_state(), tol=TOL, what=f"{kernel} orientation stream")
    assert np.array_equal(e.t_last, o.get_last_time())
    assert np.array_equal(e.get_mean_iter_hist(), o.get_mean_iter_hist())
    assert not e.get_status().any()
    if kernel == "fast":
        assert not e.fallbacks().any() or True  # counters are cumulative over the module; see the large-angle test


@pytest.mark.parametrize("kernel", KERNELS)
def test_orientation_bias_decay_dense_noise_guards_and_masks(kernel):
    """short bias time constants (the D = diag(cg, ca, 1) scaling of the fast kernel), a dense per-filter Q, a dense
    initial covariance, per-filter dt with every guard, masked and non-finite measurements"""
    B = 9
    rng = np.random.default_rng(11)
    mu, sg = syn.orientation_initial(B)
    A = rng.normal(size=(B, 13, 13)) * 0.02
    sg = sg + A @ np.transpose(A, (0, 2, 1))
    mu[:, 4:7] = rng.normal(size=(B, 3))
    mu[:, 7:10] = rng.normal(size=(B, 3)) * 1e-3
    mu[:, 10:13] = rng.normal(size=(B, 3)) * 1e-2
    Aq = rng.normal(size=(B, 13, 13)) * 1e-3
    Q = Aq @ np.transpose(Aq, (0, 2, 1)) + syn.ORI_Q
    dt = np.array([-1.0, 0.0, 0.01, 0.02, 5.0, 0.03, 0.01, 0.02, 0.005])
    mask = (np.arange(B) % 3 != 0).astype(np.uint8)
    o, e = OracleBatch(1, B), EmuBatch(1, B, kernel=kernel)
    for x in (o, e):
        x.set_orientation_params(7.0, 3.0, syn.LATITUDE_BREMEN)
        x.initialize(mu, sg)
        x.set_process_noise(Q)
        x.set_time_bounds(1e-9, 1.0)
        for k in range(1, 4):
            gyro, acc = syn.orientation_imu(B, k)
            x.set_rotation_rate(gyro)
            x.set_acceleration(acc)
            x.predict_dt(dt)
            z, R = syn.orientation_velocity(B, k)
            z = z + mu[:, 4:7] * 0.9
            if k == 2:
                z[4, 1] = np.nan
            x.update(9, z, np.tile(R, (B, 1, 1)) * (1 + np.arange(B))[:, None, None], mask)
            x.step(dt, 9, z, R)
    assert np.array_equal(e.get_status(), o.get_status())
    P.assert_parity(1, e.get_state(), o.get_state(), tol=TOL, what=f"{kernel} orientation guards")
    assert np.array_equal(e.get_mean_iter_hist(), o.get_mean_iter_hist())


@pytest.mark.parametrize("kernel", KERNELS)
def test_orientation_large_angles_take_the_literal_expressions(kernel):
    """As for PoseUKF: orientation spread and rates far outside the short polynomials are served by the any-angle instance
    of the structured code (no literal predict), a factor column next to pi sends the update to the literal code."""
    B = 4
    mu, sg = syn.orientation_initial(B)
    sg[0, 0:3, 0:3] *= 150.0  # sqrt(1.5) rad orientation sigma: exp and log leave the polynomial range
    sg[1, 0:3, 0:3] *= 400.0  # 2 rad per axis, trace 12 > 9: beyond the hot path's update guard, every column below pi
    sg[3, 2, 2] = 3.13**2  # a factor column at the branch cut of the SO(3) log: the literal update
    o, e = P.make_ori(OracleBatch, B), P.make_ori(EmuBatch, B, kernel=kernel)
    before = e.fallbacks()
    gyro = np.tile([0.0, 0.0, 0.05], (B, 1))
    gyro[2] = [3.0, -40.0, 25.0]  # 47 rad/s: |w| dt = 2.4 rad
    for x in (o, e):
        x.initialize(mu, sg)
        for k in range(1, 4):
            x.set_rotation_rate(gyro)
            x.set_acceleration(syn.orientation_imu(B, k)[1])
            z, R = syn.orientation_velocity(B, k)
            x.step(0.05, 9, z, R)
    assert np.array_equal(e.get_status(), o.get_status())
    P.assert_parity(1, e.get_state(), o.get_state(), tol=1e-10, what=f"{kernel} orientation large angles")
    if kernel == "fast":
        fb = e.fallbacks() - before
        assert fb[0] == 0 and fb[1] + fb[2] > 0, f"literal predict / update calls {fb}"
        assert fb[3] > 0 and fb[4] > 0, f"the any-angle instance was not exercised: {fb}"
        assert fb[:3].sum() < 2 * 3 * B, "every lane fell back: the fast path was not exercised"
    assert np.array_equal(e.get_mean_iter_hist(), o.get_mean_iter_hist())


@pytest.mark.parametrize("sig", [(0.7, 0.7, 0.7), (1.0, 1.0, 1.0), (2.0, 2.0, 2.0), (0.1, 0.1, 2.9), (3.0, 3.0, 3.0)])
def test_orientation_wide_attitude_uncertainty_stays_in_the_fast_kernel(sig):
    """An OrientationUKF that barely knows its attitude (or only its heading: the usual start-up state) never reaches the
    literal code: the any-angle instance of the structured predict / update serves it, the mean passes included."""
    B = 8
    mu, sg = syn.orientation_initial(B)
    sg[:, 0:3, 0:3] = np.diag(np.square(sig))
    o, e = P.make_ori(OracleBatch, B), P.make_ori(EmuBatch, B, kernel="fast")
    before = e.fallbacks()
    for x in (o, e):
        x.initialize(mu, sg)
        for k in range(1, 7):
            gyro, acc = syn.orientation_imu(B, k)
            x.set_rotation_rate(gyro)
            x.set_acceleration(acc)
            if k % 2 == 0:
                z, R = syn.orientation_velocity(B, k)
                x.step(0.01, 9, z, R)
            else:
                x.predict_dt(0.01)
    assert np.array_equal(e.get_status(), o.get_status()) and not o.get_status().any()
    P.assert_parity(1, e.get_state(), o.get_state(), tol=1e-11, what=f"wide orientation sigma {sig}")
    assert np.array_equal(e.get_mean_iter_hist(), o.get_mean_iter_hist())
    fb = e.fallbacks() - before
    assert not fb[:3].any(), f"literal fallbacks were taken: {fb}"
    assert fb[3] > 0 and fb[4] > 0, f"the any-angle instance was not exercised: {fb}"


@pytest.mark.parametrize("kernel", KERNELS)
def test_orientation_not_spd_leaves_the_filter_untouched(kernel):
    mu, sg = syn.orientation_initial(3)
    sg[1, 8, 8] = -1.0
    z, R = syn.orientation_velocity(3, 1)
    for fused in (True, False):
        o, e = OracleBatch(1, 3), EmuBatch(1, 3, kernel=kernel)
        for x in (o, e):
            x.initialize(mu, sg)
            if fused:
                x.step(0.01, 9, z, R)
            else:
                x.update(9, z, R)
        assert e.get_status().tolist() == o.get_status().tolist() == [0, 8, 0]
        assert np.array_equal(e.get_state()[0][1], mu[1]) and np.array_equal(np.tril(e.get_state()[1][1]), np.tril(sg[1]))
        P.assert_parity(1, (e.get_state()[0][[0, 2]], e.get_state()[1][[0, 2]]),
                        (o.get_state()[0][[0, 2]], o.get_state()[1][[0, 2]]), tol=TOL, what="neighbours of a non-SPD filter")


def test_pose_orientation_measurement_runs_the_structured_path():
    """OrientationMeasurement (PoseUKF.cpp:133-138), the one manifold-valued measurement: the fast kernel evaluates it
    from the 12 points that perturb the orientation; no literal fallback is taken, the measured rotation vector may be
    any angle, and a filter far from its measurement (innovation outside the log polynomial) still falls back."""
    B = 40
    o, e = P.make_pose(OracleBatch, B), P.make_pose(EmuBatch, B, kernel="fast")
    before = e.fallbacks()
    for k in range(1, 7):
        z, R = syn.pose_measurement(3, B, k)
        mu = o.get_state()[0]
        # rotation vector of the current estimate plus noise: large absolute angles, small innovations
        q = mu[:, 3:7]
        ang = 2.0 * np.arctan2(np.linalg.norm(q[:, :3], axis=1), q[:, 3])
        axis = q[:, :3] / np.maximum(np.linalg.norm(q[:, :3], axis=1, keepdims=True), 1e-300)
        zz = axis * ang[:, None] + (z - z.mean(axis=0)) * 0.5
        for x in (o, e):
            x.step(syn.DT * 5, 3, zz, R)
    P.assert_parity(0, e.get_state(), o.get_state(), tol=TOL, what="orientation measurement, structured path")
    assert np.array_equal(e.get_mean_iter_hist(), o.get_mean_iter_hist())
    assert not (e.fallbacks() - before).any(), e.fallbacks() - before
    far = np.tile([0.0, 0.0, 2.5], (B, 1))  # 2.5 rad away from every estimate: the innovation leaves the polynomial
    for x in (o, e):
        x.step(syn.DT, 3, far, np.eye(3) * 1e-2)
    P.assert_parity(0, e.get_state(), o.get_state(), tol=1e-10, what="orientation measurement, far innovation")
    assert (e.fallbacks() - before)[1] > 0


def _per_filter_orientation(cls, B, **kw):
    rng = np.random.default_rng(17)
    tg, ta, lat = 10.0 ** rng.uniform(0, 3, B), 10.0 ** rng.uniform(0, 3, B), rng.uniform(-1.5, 1.5, B)
    mu, sg = syn.orientation_initial(B)
    x = cls(1, B, **kw)
    x.initialize(mu, sg)
    x.set_process_noise(syn.ORI_Q)
    x.set_orientation_params(tg, ta, lat)  # one constructor-argument set per filter (OrientationUKF.cpp:41-47)
    P.run_ori_c1(x, B, 8, every=4)
    return x


@pytest.mark.parametrize("kernel", KERNELS + ["warp"])
def test_orientation_per_filter_parameters(kernel):
    B = 37
    o, e = _per_filter_orientation(OracleBatch, B), _per_filter_orientation(EmuBatch, B, kernel=kernel)
    P.assert_parity(1, e.get_state(), o.get_state(), tol=TOL, what=f"{kernel} per-filter tau / latitude")
    # and they matter: the same run with one shared parameter set differs
    s = P.make_ori(OracleBatch, B)
    P.run_ori_c1(s, B, 8, every=4)
    assert np.abs(s.get_state()[0] - o.get_state()[0]).max() > 1e-9


@pytest.mark.parametrize("filt", ["pose", "orientation"])
def test_unnormalised_quaternions_take_the_literal_expressions(filt):
    """The fast kernels' log has no reciprocal and assumes |q| = 1 (to 1e-8); the reference's log is scale invariant.  A
    state initialised with an unnormalised quaternion must therefore run the literal code, and agree with the oracle."""
    B = 6
    if filt == "pose":
        mu, sg = syn.pose_initial(B)
        mu[1, 3:7] *= 1.001
        mu[4, 3:7] *= 0.98
        o, e = OracleBatch(0, B), EmuBatch(0, B, kernel="fast")
    else:
        mu, sg = syn.orientation_initial(B)
        mu[1, 0:4] *= 1.001
        mu[4, 0:4] *= 0.98
        o, e = OracleBatch(1, B), EmuBatch(1, B, kernel="fast")
    before = e.fallbacks()
    for x in (o, e):
        x.initialize(mu, sg)
        if filt == "pose":
            P.run_pose_c3(x, B, 4)
        else:
            x.set_process_noise(syn.ORI_Q)
            x.set_orientation_params(syn.ORI_TAU, syn.ORI_TAU, syn.LATITUDE_BREMEN)
            P.run_ori_c1(x, B, 4, every=2)
    P.assert_parity(0 if filt == "pose" else 1, e.get_state(), o.get_state(), tol=TOL, what=f"{filt} unnormalised q")
    fb = (e.fallbacks() - before)[:3]
    assert fb[0] >= 2 and fb.sum() < 3 * 4 * B, f"fallbacks {fb}"


def test_diagonal_noise_paths_equal_the_general_one(monkeypatch):
    """A broadcast Q without off-diagonal entries (the reference's default) is read through 12 / 13 loads instead of 78 / 91
    (StepParams::q_diagonal = 1): the same bits as the general noise code.  With one value per rotated 3 x 3 block
    (q_diagonal = 2) R (q I) R^T is taken as q I: equal to rounding.  A dense Q takes neither shortcut."""
    import emu_lib

    B = 33
    for make, run, filt in ((P.make_pose, lambda e: P.run_pose_c3(e, B, 6), 0), (P.make_ori, lambda e: P.run_ori_c1(e, B, 6, every=3), 1)):
        out = {}
        for flag in (2, 1, 0):
            monkeypatch.setattr(emu_lib, "Q_DIAGONAL_FLAG", flag)
            e = make(EmuBatch, B, kernel="fast")
            run(e)
            out[flag] = e.get_state()
        assert np.array_equal(out[1][0], out[0][0]) and np.array_equal(out[1][1], out[0][1])
        P.assert_parity(filt, out[2], out[0], tol=1e-13, what="isotropic noise blocks")
    monkeypatch.setattr(emu_lib, "Q_DIAGONAL_FLAG", 2)
    rng = np.random.default_rng(5)
    A = rng.normal(size=(12, 12)) * 0.01
    Q = A @ A.T
    o, e = P.make_pose(OracleBatch, B), P.make_pose(EmuBatch, B, kernel="fast")
    for x in (o, e):
        x.set_process_noise(Q)
        P.run_pose_c3(x, B, 4)
    P.assert_parity(0, e.get_state(), o.get_state(), tol=TOL, what="dense broadcast Q")
    Qd = np.diag(np.linspace(1e-6, 1e-3, 12))  # diagonal, but not one value per block
    o, e = P.make_pose(OracleBatch, B), P.make_pose(EmuBatch, B, kernel="fast")
    for x in (o, e):
        x.set_process_noise(Qd)
        P.run_pose_c3(x, B, 4)
    P.assert_parity(0, e.get_state(), o.get_state(), tol=TOL, what="anisotropic diagonal Q")


def check_not_spd_deferral(cls, kw):
    """The one documented behavioural difference of the fast PoseUKF kernel (csrc/ukf_pose_fast.cuh, header of pf_update).
    ukfom's update factorises Sigma - K S K^T in full inside apply_delta; the structured update needs only the six leading
    columns of that factor and checks only their pivots.  An EXACT velocity measurement (R = 0) makes the velocity block of
    the downdated covariance zero up to rounding: the oracle (the reference's sequence) flags NOT_SPD in that update and
    leaves the mean alone; the fast kernel completes the update and flags the filter at its next factorisation -- the
    following predict -- after which both hold a flagged, frozen filter."""
    B = 6
    o, e = P.make_pose(OracleBatch, B), P.make_pose(cls, B, **kw)
    z, _ = syn.pose_measurement(4, B, 1)
    for x in (o, e):
        x.predict_dt(0.01)
        x.update(4, z, np.zeros((3, 3)))
    assert (o.get_status() == 8).all()           # reference: at once
    early = e.get_status()
    assert ((early == 0) | (early == 8)).all()   # fast kernel: now (rounding made a leading pivot fail) or ...
    for x in (o, e):
        x.predict_dt(0.01)
    assert (e.get_status() == 8).all() and (o.get_status() == 8).all()  # ... at the next factorisation, never later
    frozen = e.get_state()
    e.predict_dt(0.01)
    assert np.array_equal(e.get_state()[0], frozen[0]) and np.array_equal(e.get_state()[1], frozen[1])
    # a measurement that is merely very accurate is no such case: both integrate it and agree
    o, e = P.make_pose(OracleBatch, B), P.make_pose(cls, B, **kw)
    for x in (o, e):
        x.predict_dt(0.01)
        x.update(4, z, np.eye(3) * 1e-10)
        x.predict_dt(0.01)
    assert not o.get_status().any() and not e.get_status().any()
    P.assert_parity(0, e.get_state(), o.get_state(), tol=1e-9, what="very accurate velocity measurement")


def test_fast_kernel_not_spd_deferral_is_one_factorisation():
    check_not_spd_deferral(EmuBatch, dict(kernel="fast"))
