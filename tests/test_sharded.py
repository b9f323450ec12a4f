"""Sharded handles (ukfb_create_sharded): B filters split by filter index over several devices behind ONE handle.

Reference objects share nothing (UnscentedKalmanFilter.hpp:150-154), so a sharded handle must give, filter by filter,
bit for bit what a one-device handle gives -- through every host-pointer entry point, including the slot-major event
arrays a shard sees with a stride.  The device lists used: [0, 0, 0] always (three shards, three host workers, one
GPU: runs on the one-GPU box of the round-end test tier) and one shard per visible GPU when there are at least two.
CPU: argument checks and the loud refusal without a GPU."""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import parity as P
from oracle.oracle_lib import OracleBatch
from slam_pose_estimation_b200 import _build, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_lists():
    import torch

    n = torch.cuda.device_count()
    lists = [[0, 0, 0]]
    if n >= 2:
        lists.append(list(range(n)))
    return lists


def test_create_sharded_argument_checks_and_refusal():
    from slam_pose_estimation_b200 import _capi

    _build.build()
    lib = _capi.load()
    h = C.c_void_p()
    dev = (C.c_int * 2)(0, 0)
    assert lib.ukfb_create_sharded(0, 100, dev, 0, C.byref(h)) == -1
    assert lib.ukfb_create_sharded(0, 1, dev, 2, C.byref(h)) == -1  # fewer filters than shards
    assert lib.ukfb_create_sharded(5, 100, dev, 2, C.byref(h)) == -1
    assert lib.ukfb_create_sharded(0, 100, dev, 2, None) == -1
    assert lib.ukfb_shard_count(None) == 0
    import torch

    if not torch.cuda.is_available():
        assert lib.ukfb_create_sharded(0, 100, dev, 2, C.byref(h)) == -3 and not h.value  # no CPU path
        p = C.c_void_p()
        assert lib.ukfb_host_alloc(C.byref(p), 4096) == -3


def test_shard_ranges_match_the_python_partition():
    """ukfb_create_sharded documents shard_range's partition; the C side is checked on the GPU below, here the rule"""
    from slam_pose_estimation_b200.shard import shard_range

    for total, world in ((100, 3), (7, 7), (1 << 20, 8), (1001, 8)):
        got = [shard_range(total, r, world) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(got[:-1], got[1:]))
        sizes = [b - a for a, b in got]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _same(a, b):
    (ma, sa), (mb, sb) = a.get_state(), b.get_state()
    return np.array_equal(ma, mb) and np.array_equal(sa, sb)


@pytest.mark.gpu
def test_sharded_pose_equals_single_device_bitwise_and_oracle():
    from slam_pose_estimation_b200 import UkfBatch
    from slam_pose_estimation_b200.shard import shard_range

    B = 203  # ragged: shards of 68, 68, 67 -> partial tiles everywhere
    for devs in device_lists():
        s = P.make_pose(UkfBatch, B, devices=devs)
        g = P.make_pose(UkfBatch, B)
        o = P.make_pose(OracleBatch, B)
        assert s.shard_count() == len(devs) and g.shard_count() == 1
        for i in range(len(devs)):
            sh, first, count = s.shard(i)
            assert (first, first + count) == shard_range(B, i, len(devs)) and sh.B == count
        rs = 1.0 + (np.arange(B) % 3)  # per-filter measurement covariances
        for x in (s, g, o):
            P.run_pose_c3(x, B, 12, r_scale=rs)
            # per-filter dt, a mask, per-filter kinds
            dt = syn.DT * (1.0 + (np.arange(B) % 4))
            x.predict_dt(dt)
            z, R = syn.pose_measurement(4, B, 40)
            x.update(4, z, R, mask=(np.arange(B) % 3 != 0).astype(np.uint8))
            kinds = np.array([[-1, 0, 1, 2, 4, 5, 6, 7, 8][b % 9] for b in range(B)], np.int8)
            mu3, cov = np.zeros((B, 3)), np.tile(np.eye(3), (B, 1, 1))
            for k in set(kinds.tolist()) - {-1}:
                zk, Rk = syn.pose_measurement(int(k), B, 41)
                m = zk.shape[1]
                sel = kinds == k
                mu3[sel, :m] = zk[sel]
                cov[sel, :m, :m] = Rk
            x.update_mixed(kinds, mu3, cov)
            x.set_acceleration(0.05 * syn.noise(np.arange(B), 50, 13, 3), np.eye(3) * 1e-4, mask=(np.arange(B) % 2).astype(np.uint8))
            x.predict_dt(syn.DT)
        assert _same(s, g), f"devices {devs}: sharded != single device"
        P.assert_parity(0, s.get_state(), o.get_state(), what=f"sharded PoseUKF on {devs}")
        assert np.array_equal(s.get_status(), g.get_status()) and not s.get_status().any()
        assert np.array_equal(s.get_mean_iter_hist(), g.get_mean_iter_hist())
        assert np.array_equal(s.get_mean_iter_hist(), o.get_mean_iter_hist())
        assert s.launch_count() == len(devs) * g.launch_count()
        # pose-only readback and the body states land shard by shard in one buffer
        assert np.array_equal(s.get_mu_range(0, 7), g.get_state()[0][:, :7])
        assert np.array_equal(s.get_body_states(), g.get_body_states())
        # `_dev` calls belong to the per-device handles
        with pytest.raises(Exception):
            s.step_dev(0, False, -1)
        s.close(), g.close()


@pytest.mark.gpu
def test_sharded_event_queues_blocking_and_streaming():
    """slot-major K x B queues: a shard's slice of every slot is a strided copy (cudaMemcpy2DAsync)"""
    import torch
    from slam_pose_estimation_b200 import UkfBatch

    B = 150
    ts, kinds, mu3 = syn.pose_c5_events(B, 1, 24, dvl_period=7, gps_period=11)
    tab = syn.sensor_cov_table()
    K = ts.shape[0]
    per_event = np.ascontiguousarray(tab[np.maximum(kinds, 0)].reshape(K, B, 9))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    for devs in device_lists():
        g = P.make_pose(UkfBatch, B)
        o = P.make_pose(OracleBatch, B)
        g.run_events(ts, kinds, mu3, tab)
        o.run_events(ts, kinds, mu3, tab)
        s1 = P.make_pose(UkfBatch, B, devices=devs)
        s1.run_events(ts, kinds, mu3, tab)  # per-sensor covariance table (broadcast to every shard)
        assert _same(s1, g)
        s2 = P.make_pose(UkfBatch, B, devices=devs)
        s2.run_events(ts, kinds, mu3, per_event)  # one covariance per event (strided like the samples)
        assert _same(s2, g)
        assert np.array_equal(s1.get_last_time(), g.get_last_time())
        P.assert_parity(0, s1.get_state(), o.get_state(), what=f"sharded event queues on {devs}")
        # streaming calls: three windows, estimates of each window gathered into one pinned buffer
        s3 = P.make_pose(UkfBatch, B, devices=devs)
        b = P.make_pose(UkfBatch, B)
        cuts = [0, K // 3, 2 * K // 3, K]
        outs = [pin(np.zeros((B, 13))) for _ in range(3)]
        pose7 = [pin(np.zeros((B, 7))) for _ in range(3)]
        tabp = pin(tab)
        parts = [(pin(ts[lo:hi]), pin(kinds[lo:hi]), pin(mu3[lo:hi])) for lo, hi in zip(cuts[:-1], cuts[1:])]
        for i, (t, k, m) in enumerate(parts):
            s3.run_events_async(t, k, m, tabp)
            s3.get_state_async(outs[i])
            s3.get_mu_range_async(0, 7, pose7[i])
        s3.synchronize()
        for i, (t, k, m) in enumerate(parts):
            b.run_events(t, k, m, tab)
            assert np.array_equal(outs[i], b.get_state()[0])
            assert np.array_equal(pose7[i], outs[i][:, :7])
        assert _same(s3, g)
        for x in (g, s1, s2, s3, b):
            x.close()


@pytest.mark.gpu
def test_sharded_orientation_per_filter_parameters_status_and_times():
    from slam_pose_estimation_b200 import UkfBatch

    B = 101
    tau_g = 1000.0 + 10.0 * np.arange(B)
    tau_a = 2000.0 + 5.0 * np.arange(B)
    lat = np.linspace(-1.2, 1.2, B)
    for devs in device_lists():
        objs = [P.make_ori(UkfBatch, B, devices=devs), P.make_ori(UkfBatch, B), P.make_ori(OracleBatch, B)]
        for x in objs:
            x.set_orientation_params(tau_g, tau_a, lat)
            Q = np.tile(syn.ORI_Q, (B, 1, 1)) * (1.0 + (np.arange(B) % 5))[:, None, None]
            x.set_process_noise(Q)  # per-filter process noise
            P.run_ori_c1(x, B, 20, every=5)
            bad = syn.orientation_velocity(B, 21)[0]
            bad[::10, 1] = np.nan  # non-finite measurements: flagged per filter, on whichever shard they live
            x.update(9, bad, np.eye(3) * syn.SIGMA_DVL**2)
        s, g, o = objs
        assert _same(s, g)
        P.assert_parity(1, s.get_state(), o.get_state(), what=f"sharded OrientationUKF on {devs}")
        st = s.get_status()
        assert np.array_equal(st, g.get_status()) and np.array_equal(st, o.get_status())
        assert (st[::10] == 4).all() and st.sum() == 4 * len(st[::10])
        assert s.status_summary() == g.status_summary() == (len(st[::10]), 4)
        assert np.array_equal(s.get_rotation_rate(), g.get_rotation_rate())
        assert np.array_equal(s.get_last_time(), g.get_last_time())
        assert np.array_equal(s.get_process_noise(per_filter=True), g.get_process_noise(per_filter=True))
        tl = np.arange(B, dtype=np.int64) + 5_000_000
        s.set_last_time(tl)
        assert np.array_equal(s.get_last_time(), tl)
        s.clear_status()
        assert s.status_summary() == (0, 0)
        s.close(), g.close()


def _wobble(b, k, c):
    return math.sin(0.37 * b + 1.3 * k + 0.71 * c)


def _oracle_for_sharded_demo(B):
    """the calls of tests/cpp/sharded_demo.cpp::drive on the oracle"""
    d0 = np.array([1, 1, 1, 0.01, 0.01, 0.01, 0.1, 0.1, 0.1, 0.01, 0.01, 0.01])
    mu0, sg0 = np.zeros((B, 13)), np.zeros((B, 12, 12))
    for b in range(B):
        yaw = 0.3 * _wobble(b, 20, 0)
        mu0[b, 0], mu0[b, 1] = _wobble(b, 21, 0), _wobble(b, 21, 1)
        mu0[b, 5], mu0[b, 6] = math.sin(0.5 * yaw), math.cos(0.5 * yaw)
        mu0[b, 7], mu0[b, 12] = 1.0 + 0.1 * _wobble(b, 22, 0), 0.05
        sg0[b] = np.diag([d0[i] * (1.0 + 0.2 * _wobble(b, 23, i)) for i in range(12)])
    o = OracleBatch(0, B)
    o.initialize(mu0, sg0)
    o.predict_time(np.full(B, 1000000, np.int64))
    o.predict_time(np.array([1010000 + 10 * (b % 7) for b in range(B)], np.int64))
    w = np.array([[(0.05 if i == 2 else 0.0) + 1e-3 * _wobble(b, 0, i) for i in range(3)] for b in range(B)])
    wc = np.array([np.eye(3) * 1e-6 * (1.0 + b % 3) for b in range(B)])
    o.update(8, w, wc)
    xy = np.array([[0.01 * _wobble(b, 1, i) for i in range(2)] for b in range(B)])
    o.update(1, xy, np.tile(np.eye(2) * 0.25, (B, 1, 1)))
    o.update(4, np.tile([1.01, 0.0, 0.0], (B, 1)), np.eye(3) * 1e-4)
    o.predict_dt(0.01)
    queues = [[] for _ in range(B)]
    for b in range(B):
        for k in range(3 + b % 3):
            mu = [(0.05 if i == 2 else 0.0) + 1e-3 * _wobble(b, 2 + k, i) for i in range(3)]
            queues[b].append((1030000 + 1000 * k + 10 * (b % 5), 8, mu, np.eye(3) * 1e-6))
        if b % 4 == 1:
            c = np.eye(3)
            queues[b].append((1040000, 2, [0.02 * _wobble(b, 9, 0), 0, 0], c))
    K = max(len(q) for q in queues)
    ts, kinds = np.zeros((K, B), np.int64), np.full((K, B), -2, np.int8)
    mu3, cov = np.zeros((K, B, 3)), np.zeros((K, B, 9))
    for b, q in enumerate(queues):
        for k, (t, kind, mu, c) in enumerate(q):
            ts[k, b], kinds[k, b], mu3[k, b], cov[k, b] = t, kind, mu, c.ravel()
    o.run_events(ts, kinds, mu3, cov)
    return o


@pytest.fixture(scope="module")
def sharded_demo(tmp_path_factory):
    _build.build()
    exe = str(tmp_path_factory.mktemp("cpp") / "sharded_demo")
    libdir = os.path.dirname(_build.LIB)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "sharded_demo.cpp"), "-L", libdir, "-lukfb", f"-Wl,-rpath,{libdir}", "-o", exe],
                   check=True)
    return exe


def test_sharded_cpp_host_compiles_and_fails_loudly_without_gpu(sharded_demo):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sharded_demo, "64", "0,0"], capture_output=True, text=True)
    assert r.returncode != 0 and ("no CPU path" in r.stderr or "CUDA" in r.stderr)


@pytest.mark.gpu
def test_cpp_host_drives_all_gpus_through_one_object(sharded_demo):
    """C++ host code, one PoseUKF object over a device list: equals the one-device object bit for bit (checked in the
    program) and the oracle replaying the same calls (checked here)"""
    B = 300
    o = _oracle_for_sharded_demo(B)
    mu_ref, sg_ref = o.get_state()
    tl_ref = o.get_last_time()
    for devs in device_lists():
        r = subprocess.run([sharded_demo, str(B), ",".join(str(d) for d in devs)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        lines = r.stdout.splitlines()
        assert lines[0].split() == ["shards", str(len(devs)), "batch", str(B)]
        assert lines[1] == "sharded_equals_single 1"
        idx, mu, sg = [], [], []
        for line in lines[2:]:
            tok = line.split()
            assert tok[0] == "filter"
            b = int(tok[1])
            assert int(tok[2]) == tl_ref[b]
            v = np.array([float(x) for x in tok[3:]])
            idx.append(b), mu.append(v[:13]), sg.append(v[13:].reshape(12, 12))
        assert len(idx) >= 16
        P.assert_parity(0, (np.array(mu), np.array(sg)), (mu_ref[idx], sg_ref[idx]), what=f"C++ sharded object on {devs}")
