"""The C++ host shim (include/pose_estimation_b200/*.hpp: the reference's class API over the C ABI).
CPU: it compiles against the headers, links libukfb.so and refuses to run without a GPU.
GPU: the demo's PoseUKF / OrientationUKF sequence matches the oracle replaying the same calls, and the
reference's three exceptions are thrown with the reference's messages."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

import parity as P
from oracle.oracle_lib import OracleBatch
from slam_pose_estimation_b200 import _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    _build.build()
    exe = str(tmp_path_factory.mktemp("cpp") / "shim_demo")
    libdir = os.path.dirname(_build.LIB)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "shim_demo.cpp"), "-L", libdir, "-lukfb", f"-Wl,-rpath,{libdir}", "-o", exe],
                   check=True)
    return exe


def test_shim_compiles_and_fails_loudly_without_gpu(demo):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([demo], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU path" in r.stderr or "CUDA" in r.stderr


def _oracle_replay():
    mu0 = np.zeros((1, 13))
    mu0[0, 6], mu0[0, 7], mu0[0, 12] = 1.0, 1.0, 0.05
    sg0 = np.diag([1, 1, 1, 0.01, 0.01, 0.01, 0.1, 0.1, 0.1, 0.01, 0.01, 0.01])[None]
    o = OracleBatch(0, 1)
    o.initialize(mu0, sg0)
    o.predict_time(np.array([1000000], np.int64))
    o.predict_time(np.array([1010000], np.int64))
    o.update(0, [[0.02, -0.01, 0.005]], np.eye(3) * 0.25)
    o.update(1, [[0.01, 0.0]], np.eye(2))
    o.update(2, [[-0.02]], np.eye(1))
    o.update(3, [[0.0, 0.0, 0.001]], np.eye(3) * 1e-4)
    o.update(4, [[1.01, 0.0, 0.0]], np.eye(3) * 1e-4)
    o.update(5, [[0.99, 0.0]], np.eye(2))
    o.update(6, [[0.0]], np.eye(1))
    o.update(7, [[1.0, 0.05]], np.eye(2))
    o.update(8, [[0.0, 0.0, 0.049]], np.eye(3) * 1e-6)
    o.set_acceleration([[0.1, 0.0, 0.0]], np.eye(3) * 1e-4)
    o.predict_dt(0.01)
    pose = o.get_state()
    mu, sg = np.zeros((1, 14)), np.diag([0.01] * 6 + [1e-6] * 3 + [1e-4] * 4)[None]
    mu[0, 3], mu[0, 13] = 1.0, 9.81
    q = OracleBatch(1, 1)
    q.set_orientation_params(3600.0, 3600.0, 0.92698121)
    q.initialize(mu, sg)
    q.set_process_noise(np.diag([1e-6] * 3 + [1e-4] * 3 + [1e-10] * 3 + [1e-8] * 3 + [1e-12]))
    for k in range(5):
        q.set_rotation_rate([[0.0, 0.0, 0.05]])
        q.set_acceleration([[0.0, 0.0, 9.81]])
        q.predict_time(np.array([1000000 + 1000 * k], np.int64))
    q.update(9, [[0.01, 0.0, 0.0]], np.eye(3) * 1e-4)
    return pose, q.get_state(), q.get_rotation_rate()


def _oracle_event_queue():
    """the samples shim_demo.cpp pushes into its EventQueue, as slot-major arrays"""
    B = 3
    queues = [[] for _ in range(B)]
    for b in range(B):
        for k in range(5):
            queues[b].append((1000000 + 1000 * k + 100 * b, 8, [0, 0, 0.05 + 0.001 * b], np.eye(3) * 1e-6))
            if k == 2 and b >= 1:
                queues[b].append((1002500, 4, [1.0 + 0.01 * b, 0, 0], np.eye(3) * 1e-4))
            if k == 3 and b == 2:
                queues[b].append((1003500, 10, [0.2, 0, 0], np.eye(3) * 1e-4))
    K = max(len(q) for q in queues)
    ts, kinds = np.zeros((K, B), np.int64), np.full((K, B), -2, np.int8)
    mu3, cov = np.zeros((K, B, 3)), np.zeros((K, B, 9))
    for b, q in enumerate(queues):
        for k, (t, kind, mu, c) in enumerate(q):
            ts[k, b], kinds[k, b], mu3[k, b], cov[k, b] = t, kind, mu, c.ravel()
    mu0 = np.zeros((B, 13))
    mu0[:, 6], mu0[:, 7], mu0[:, 12] = 1.0, 1.0, 0.05
    sg0 = np.tile(np.diag([1, 1, 1, 0.01, 0.01, 0.01, 0.1, 0.1, 0.1, 0.01, 0.01, 0.01]), (B, 1, 1))
    o = OracleBatch(0, B)
    o.initialize(mu0, sg0)
    o.run_events(ts, kinds, mu3, cov)
    return K, o.get_state()


@pytest.mark.gpu
def test_shim_matches_oracle_and_throws_like_the_reference(demo):
    r = subprocess.run([demo], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    vals, caught, evq_mu, evq_sigma = {}, [], [], []
    for line in r.stdout.splitlines():
        tag, *rest = line.split()
        if tag == "evq_mu":
            evq_mu.append([float(x) for x in rest])
        elif tag == "evq_sigma":
            evq_sigma.append([float(x) for x in rest])
        elif tag == "rbs":
            rbs_out = np.array([float(x) for x in rest])
        elif tag == "evq_depth":
            evq_depth = int(rest[0])
        elif tag == "caught:":
            caught.append(" ".join(rest))
        elif tag in ("pose_mu", "pose_sigma", "ori_mu", "ori_sigma", "ori_rate"):
            vals[tag] = np.array([float(x) for x in rest])
        elif tag == "pose_last_time":
            assert int(rest[0]) == 1010000
    assert caught == ["Delta time is negative!", "Delta time is greater then the allowed maximum!",
                      "Measurement or covariance contains non-finite values!"]
    pose, ori, rate = _oracle_replay()
    P.assert_parity(0, (vals["pose_mu"][None], vals["pose_sigma"].reshape(1, 12, 12)), pose, what="C++ shim PoseUKF")
    P.assert_parity(1, (vals["ori_mu"][None], vals["ori_sigma"].reshape(1, 13, 13)), ori, what="C++ shim OrientationUKF")
    assert np.abs(vals["ori_rate"] - rate[0]).max() < 1e-12
    rbs = np.zeros(49)
    rbs[0:3], rbs[5:7], rbs[7:9], rbs[12] = [1.0, 2.0, -3.0], [np.sin(0.3), np.cos(0.3)], [0.5, 0.1], 0.02
    for blk, v in enumerate((0.5, 0.01, 0.1, 0.01)):
        rbs[13 + 9 * blk:22 + 9 * blk] = (np.eye(3) * v).ravel()
    rbs[14] = rbs[16] = 0.05
    from oracle import oracle_lib as O
    o = OracleBatch(0, 1)
    o.initialize(*O.from_body_states(rbs[None]))
    o.predict_dt(0.05)
    ref_rbs = O.to_body_states(*o.get_state())[0]
    assert np.abs(rbs_out - ref_rbs).max() < 1e-12
    K, ref = _oracle_event_queue()
    assert evq_depth == K == 7
    P.assert_parity(0, (np.array(evq_mu), np.array(evq_sigma).reshape(3, 12, 12)), ref, what="C++ EventQueue")
