"""Event streams (ukfb_run_events): per-filter queues of asynchronous sensor samples, each integrated as the reference's
aggregator callbacks do -- predictionStepFromSampleTime(ts) then integrateMeasurement(sample)
(UnscentedKalmanFilter.hpp:83-100, PoseUKF.cpp:112-178, OrientationUKF.cpp:53-72).  The oracle runs that loop filter by
filter; the kernels run all K slots in one launch.  CPU: the kernels' source under tests/simt_emu; GPU: the C ABI."""
from __future__ import annotations

import numpy as np
import pytest

import parity as P
from oracle.oracle_lib import OracleBatch
from slam_pose_estimation_b200 import synthetic as syn

IDLE = syn.EVENT_IDLE


def c5(B, n_ticks, first=0):
    ts, kinds, mu3 = syn.pose_c5_events(B, 1, n_ticks, first=first, dvl_period=7, gps_period=11)
    return ts, kinds, mu3, syn.sensor_cov_table()


def test_pack_events_keeps_per_filter_time_order():
    ts, kinds, mu3, _ = c5(37, 25)
    assert (kinds == 8).sum() == 37 * 25 and (kinds == 4).sum() > 0 and (kinds == 1).sum() > 0
    assert (kinds == IDLE).any()  # ragged queues
    for b in range(37):
        live = kinds[:, b] != IDLE
        assert (np.diff(ts[live, b]) > 0).all()
        assert not live[np.argmin(live):].any() or live.all()  # padding only at the tail


def pose_edge_events(B):
    """hand-made queues: acceleration samples (finite and the NaN sentinel), a negative time step, a too large one, a
    repeated timestamp (dt <= min_dt: update without predict), an orientation measurement, predict-only slots."""
    K = 12
    ts = np.zeros((K, B), np.int64)
    kinds = np.full((K, B), IDLE, np.int8)
    mu3 = np.zeros((K, B, 3))
    cov = np.zeros((K, B, 3, 3))
    t = np.full(B, syn.T0_US, np.int64)
    plan = [(8, 1000), (10, 500), (4, 1500), (-1, 2000), (0, 0), (3, 1000), (10, 700), (8, -4000), (1, 9000), (7, 3_000_000),
            (2, 1000), (6, 1000)]
    for k, (kind, step_us) in enumerate(plan):
        act = (np.arange(B) + k) % 4 != 3  # every filter sits out some slots
        t = np.where(act, t + step_us, t)
        ts[k], kinds[k] = t, np.where(act, kind, IDLE)
        if kind == 10:
            mu3[k] = 0.05 * syn.noise(np.arange(B), k, 13, 3)
            if k == 6:
                mu3[k, ::2] = np.nan  # back to "no acceleration" for every other filter
            cov[k] = np.eye(3) * 1e-4
        elif kind >= 0:
            m = {1: 2, 7: 2, 2: 1, 6: 1}.get(kind, 3)
            z, R = syn.pose_measurement(kind, B, k + 1)
            mu3[k, :, :m] = z
            cov[k, :, :m, :m] = R
    return ts, kinds, mu3, cov.reshape(K, B, 9)


def check_pose(cls, kw, tol):
    B = 37
    ts, kinds, mu3, tab = c5(B, 25)
    o, e = P.make_pose(OracleBatch, B), P.make_pose(cls, B, **kw)
    h = ts.shape[0] // 2
    for x in (o, e):
        x.run_events(ts[:h], kinds[:h], mu3[:h], tab)
        x.run_events(ts[h:], kinds[h:], mu3[h:], tab)  # queues continue across calls
    P.assert_parity(0, e.get_state(), o.get_state(), tol=tol, what="C5 queues")
    assert np.array_equal(e.get_last_time(), o.get_last_time()) if hasattr(e, "get_last_time") else True
    assert not e.get_status().any() and not o.get_status().any()
    assert np.array_equal(e.get_mean_iter_hist(), o.get_mean_iter_hist())

    ts, kinds, mu3, cov = pose_edge_events(B)
    o, e = P.make_pose(OracleBatch, B), P.make_pose(cls, B, **kw)
    for x in (o, e):
        x.set_time_bounds(1e-9, 2.0)
        x.run_events(ts, kinds, mu3, cov)
    P.assert_parity(0, e.get_state(), o.get_state(), tol=max(tol, 1e-10), what="edge queues")
    st = o.get_status()
    assert (st & 1).any() and (st & 2).any()  # the negative and the too large step were seen
    assert np.array_equal(e.get_status(), st)


def check_ori(cls, kw, tol):
    B = 33
    K = 30
    ts = np.zeros((K, B), np.int64)
    kinds = np.full((K, B), IDLE, np.int8)
    mu3 = np.zeros((K, B, 3))
    t = np.full(B, syn.T0_US, np.int64)
    for k in range(K):
        tick = k // 3 + 1
        gyro, acc = syn.orientation_imu(B, tick)
        act = (np.arange(B) + k) % 5 != 4
        if k % 3 == 0:
            kind, z, step = 11, gyro, 400
        elif k % 3 == 1:
            kind, z, step = 12, acc, 0
        else:
            kind, z, step = (9, syn.orientation_velocity(B, tick)[0], 600) if tick % 2 == 0 else (-1, np.zeros((B, 3)), 600)
        t = np.where(act, t + step, t)
        ts[k], kinds[k], mu3[k] = t, np.where(act, kind, IDLE), z
    mu3[7, 3] = np.inf      # a non-finite acceleration sample: rejected, the old one stays
    mu3[17, 5, 1] = np.nan  # a non-finite velocity measurement: rejected
    kinds[4, 2] = 8         # a PoseUKF kind in an OrientationUKF queue
    tab = syn.sensor_cov_table()
    o, e = P.make_ori(OracleBatch, B), P.make_ori(cls, B, **kw)
    for x in (o, e):
        x.run_events(ts, kinds, mu3, tab)
    P.assert_parity(1, e.get_state(), o.get_state(), tol=tol, what="orientation queues")
    st = o.get_status()
    assert st[3] & 4 and st[5] & 4 and st[2] & 32
    assert np.array_equal(e.get_status(), st)
    # the stored IMU samples survive the launch: one more predict-only slot uses them
    ts2 = (ts.max(axis=0) + 1000)[None]
    for x in (o, e):
        x.run_events(ts2, np.full((1, B), -1, np.int8), np.zeros((1, B, 3)), tab)
    P.assert_parity(1, e.get_state(), o.get_state(), tol=tol, what="orientation queues, next launch")


def check_late_sample_is_dropped(cls, kw):
    """A sample that arrives out of order (timestamp before the filter's last one) or after too long a gap makes
    predictionStep throw (UnscentedKalmanFilter.hpp:110-122): the reference's callback ends there, so the sample is
    neither integrated nor stored.  The queue with such samples must leave the filters exactly where the queue
    without them does -- apart from the status bit and, for the too large step, the time latch (:96-97)."""
    B = 40
    K = 6
    ts = np.zeros((K, B), np.int64)
    kinds = np.full((K, B), IDLE, np.int8)
    mu3 = np.zeros((K, B, 3))
    cov = np.zeros((K, B, 3, 3))
    plan = [(8, 1000), (4, 2000), (8, -1500), (10, -500), (0, 5_000_000), (8, 6000)]  # slots 2, 3: backwards; slot 4: too large
    t = np.full(B, syn.T0_US, np.int64)
    for k, (kind, step_us) in enumerate(plan):
        ts[k] = t + step_us
        if step_us > 0 and step_us < 1_000_000:
            t = t + step_us
        kinds[k] = kind
        if kind == 10:
            mu3[k], cov[k] = 0.3, np.eye(3) * 1e-3
        else:
            z, R = syn.pose_measurement(kind, B, k + 1)
            mu3[k], cov[k] = z, R
    cov = cov.reshape(K, B, 9)
    keep = [0, 1, 5]
    a, b = P.make_pose(cls, B, **kw), P.make_pose(cls, B, **kw)
    for x in (a, b):
        x.set_time_bounds(1e-9, 2.0)
    a.run_events(ts, kinds, mu3, cov)
    b.run_events(ts[keep], kinds[keep], mu3[keep], cov[keep])
    st = a.get_status()
    assert (st == 3).all() and not b.get_status().any()  # NEG_DT | DT_TOO_LARGE, nothing else
    # the too large step moved the latch (:96-97 runs before the throw), so the last sample of `a` sees a negative
    # step and is dropped as well: compare `a` with the queue cut before it, and `b` after its own two samples
    c = P.make_pose(cls, B, **kw)
    c.set_time_bounds(1e-9, 2.0)
    c.run_events(ts[:2], kinds[:2], mu3[:2], cov[:2])
    ma, sa = a.get_state()
    mc, sc = c.get_state()
    assert np.array_equal(ma, mc) and np.array_equal(sa, sc)
    if hasattr(a, "get_last_time"):
        assert np.array_equal(a.get_last_time(), ts[4])
    assert not np.array_equal(b.get_state()[0], mc)  # the in-order sample of slot 5 was integrated there


def test_oracle_late_sample_is_dropped():
    check_late_sample_is_dropped(OracleBatch, {})


@pytest.mark.parametrize("kernel", ["thread", "fast", "warp"])
def test_emu_late_sample_is_dropped(kernel):
    from emu_lib import EmuBatch
    check_late_sample_is_dropped(EmuBatch, dict(kernel=kernel))


@pytest.mark.gpu
def test_gpu_late_sample_is_dropped():
    from slam_pose_estimation_b200 import UkfBatch
    check_late_sample_is_dropped(UkfBatch, {})
    # the single calls of the C++ shim throw at the same points: predict_time flags, the caller then skips the update
    B = 8
    f = P.make_pose(UkfBatch, B)
    f.predict_time(np.full(B, syn.T0_US, np.int64))
    f.predict_time(np.full(B, syn.T0_US + 1000, np.int64))
    f.predict_time(np.full(B, syn.T0_US + 500, np.int64))
    assert (f.get_status() == 1).all()


@pytest.mark.parametrize("kernel", ["thread", "fast", "warp"])
def test_emu_pose_event_queues(kernel):
    from emu_lib import EmuBatch
    check_pose(EmuBatch, dict(kernel=kernel), 1e-12)


@pytest.mark.parametrize("kernel", ["thread", "fast", "warp"])
def test_emu_orientation_event_queues(kernel):
    from emu_lib import EmuBatch
    check_ori(EmuBatch, dict(kernel=kernel), 1e-12)


@pytest.mark.gpu
def test_gpu_pose_event_queues():
    from slam_pose_estimation_b200 import UkfBatch
    check_pose(UkfBatch, {}, P.TOL)


@pytest.mark.gpu
def test_gpu_orientation_event_queues():
    from slam_pose_estimation_b200 import UkfBatch
    check_ori(UkfBatch, {}, P.TOL)


@pytest.mark.gpu
def test_gpu_event_queue_equals_single_calls():
    """one launch over K slots == K x (ukfb_predict_time + ukfb_update_mixed) through the single calls"""
    from slam_pose_estimation_b200 import UkfBatch
    B = 70
    ts, kinds, mu3, tab = c5(B, 12)
    a, b = P.make_pose(UkfBatch, B), P.make_pose(UkfBatch, B)
    a.run_events(ts, kinds, mu3, tab)
    for k in range(ts.shape[0]):
        live = kinds[k] != IDLE
        # idle filters: repeat their last timestamp (dt = 0 is a no-op and leaves the latch alone)
        tk = np.where(live, ts[k], np.maximum(b.get_last_time(), 1))
        if k == 0:
            assert live.all()
        b.predict_time(tk)
        cov = tab[np.maximum(kinds[k], 0)]
        b.update_mixed(np.where(live, kinds[k], -1).astype(np.int8), mu3[k], cov)
    ma, sa = a.get_state()
    mb, sb = b.get_state()
    assert np.array_equal(ma, mb) and np.array_equal(sa, sb)


@pytest.mark.gpu
def test_gpu_event_queue_streaming_calls():
    """ukfb_run_events_async + ukfb_get_state_async over three windows == the blocking calls, bit for bit"""
    import torch
    from slam_pose_estimation_b200 import UkfBatch
    B = 70
    ts, kinds, mu3, tab = c5(B, 18)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    K = ts.shape[0]
    cuts = [0, K // 3, 2 * K // 3, K]
    a, b = P.make_pose(UkfBatch, B), P.make_pose(UkfBatch, B)
    outs = [pin(np.zeros((B, 13))) for _ in range(3)]
    tabp = pin(tab)
    parts = [(pin(ts[lo:hi]), pin(kinds[lo:hi]), pin(mu3[lo:hi])) for lo, hi in zip(cuts[:-1], cuts[1:])]
    for i, (t, k, m) in enumerate(parts):
        a.run_events_async(t, k, m, tabp)
        a.get_state_async(outs[i])
    a.synchronize()
    for i, (t, k, m) in enumerate(parts):
        b.run_events(t, k, m, tab)
        assert np.array_equal(outs[i], b.get_state()[0])
    assert np.array_equal(a.get_state()[1], b.get_state()[1])


def _gps_fixes_through_the_projection(mu3, kinds, tmp_path):
    """XYMeasurement samples as the reference's callers make them: the GPS fix is a (lat, lon) pair that
    GeographicProjection::worldToNav (GeographicProjection.cpp:29-37) turns into nav-plane x, y.  The synthetic stream
    holds noisy nav-plane positions; they become GPS fixes through navToWorld (:39-44) and come back through worldToNav,
    the host-side feeder path (the C++ mirror under include/pose_estimation_b200/)."""
    import ctypes as C
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / "libgeo.so")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "cpp", "geo_capi.cpp"), "-o", so], check=True)
    geo = C.CDLL(so)
    sel = kinds == 1
    x, y = np.ascontiguousarray(mu3[sel][:, 0]), np.ascontiguousarray(mu3[sel][:, 1])
    n = x.size
    lat, lon, x2, y2 = np.empty(n), np.empty(n), np.empty(n), np.empty(n)
    pd = lambda a: a.ctypes.data_as(C.c_void_p)
    lat0, lon0 = C.c_double(syn.LATITUDE_BREMEN), C.c_double(0.154595663)
    assert geo.geo_nav_to_world(lat0, lon0, C.c_long(n), pd(x), pd(y), pd(lat), pd(lon)) == 0
    assert geo.geo_world_to_nav(lat0, lon0, C.c_long(n), pd(lat), pd(lon), pd(x2), pd(y2)) == 0
    assert np.abs(x2 - x).max() < 1e-6 and np.abs(y2 - y).max() < 1e-6 and n > 0  # the projection round trip
    out = mu3.copy()
    xy = out[sel]
    xy[:, 0], xy[:, 1] = x2, y2
    out[sel] = xy
    return out, n


@pytest.mark.gpu
def test_gpu_10k_ticks_of_the_mixed_imu_dvl_gps_stream(tmp_path):
    """north star: 1e-9 after 10 000 steps on the asynchronous IMU / DVL / GPS stream (BASELINE config 5).  64 PoseUKF
    filters, 10 000 IMU ticks at 1 kHz (AngularVelocityMeasurement, PoseUKF.cpp:168-173), DVL at 10 Hz
    (VelocityMeasurement, :140-145) and GPS at 1 Hz (XYMeasurement, :119-124, coordinates through the GeographicProjection
    mirror), per-filter phases, queue windows of 337 slots through ukfb_run_events_async; the oracle runs the reference's
    callback loop over the same queues."""
    import torch
    from slam_pose_estimation_b200 import UkfBatch

    B, ticks, W = 64, 10_000, 337
    ts, kinds, mu3 = syn.pose_c5_events(B, 1, ticks)
    mu3, n_gps = _gps_fixes_through_the_projection(mu3, kinds, tmp_path)
    K = ts.shape[0]
    assert (kinds == 8).sum() == B * ticks and (kinds == 4).sum() >= B * (ticks // 100 - 1) and n_gps >= B * (ticks // 1000 - 1)
    tab = syn.sensor_cov_table()
    g, o = P.make_pose(UkfBatch, B), P.make_pose(OracleBatch, B)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    tabp = pin(tab)
    keep = []
    for lo in range(0, K, W):
        part = (pin(ts[lo:lo + W]), pin(kinds[lo:lo + W]), pin(mu3[lo:lo + W]))
        keep.append(part)
        g.run_events_async(*part, tabp)
    g.synchronize()
    o.run_events(ts, kinds, mu3, tab)
    em, es = P.assert_parity(0, g.get_state(), o.get_state(), tol=1e-9, what="10 000 ticks of the mixed stream")
    assert np.array_equal(g.get_status(), o.get_status()) and not g.get_status().any()
    assert np.array_equal(g.get_last_time(), o.get_last_time())
    assert np.array_equal(g.get_mean_iter_hist(), o.get_mean_iter_hist())
    assert P.spd_ok(g.get_state()[1])
    print(f"10k mixed stream: mu err {em:.2e}, sigma err {es:.2e}, {K} slots, {n_gps} GPS fixes")
