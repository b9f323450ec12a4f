#!/usr/bin/env python
"""tools/bench_sharded.py -- ONE process, ONE handle, all GPUs: the C-ABI route a C++ host takes (ukfb_create_sharded).

The driver's scaling bench launches one rank per GPU (torchrun, bench.py); this tool measures the other deployment the
north star names: host code holding a single handle whose filters are split by index over the N GPUs of the box, one
host worker thread per device inside the library, no NCCL anywhere.  Workload = bench.py's C4.  Prints one JSON line per
configuration: weak (1 Mi filters per GPU) and strong (1 Mi filters in total), device-timed (CUDA events per shard, max
over shards) and end to end through the host-pointer streaming calls with pinned buffers, plus the final gather
(ukfb_get_state into one host buffer) and a bitwise check of a strided sample against a one-device handle.

    python tools/bench_sharded.py [--gpus N] [--steps K]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (workload generators only)


def pinned(shape, dtype=np.float64):
    import torch

    t = torch.empty(shape, dtype=torch.float64 if dtype == np.float64 else torch.int64).pin_memory()
    return t.numpy()


def run(n_gpus: int, total: int, steps: int, label: str):
    from slam_pose_estimation_b200 import synthetic as syn
    from slam_pose_estimation_b200.batch import UkfBatch

    pool = 4
    mu0, sg0, zs, R = bench.make_workload(total, 0, pool)
    f = UkfBatch(0, total, devices=list(range(n_gpus)))
    f.initialize(mu0, sg0)
    f.set_measurement_cov(8, R)
    z_pin = []
    for j in range(pool):
        a = pinned((total, 3))
        a[:] = zs[j]
        z_pin.append(a)
    dt_pin = pinned((1,))
    dt_pin[0] = syn.DT
    mu_pin = [pinned((total, 13)) for _ in range(2)]
    pose_pin = [pinned((total, 7)) for _ in range(2)]
    applied = []

    def loop(read, n):
        for k in range(n):
            applied.append(k % pool)
            f.step_async(dt_pin, 8, z_pin[k % pool], None)
            read(k)
        f.synchronize()

    out = {"label": label, "gpus": n_gpus, "filters_total": total, "filters_per_gpu": total // n_gpus, "steps": steps,
           "api": "one ukfb_create_sharded handle, host-pointer streaming calls (ukfb_step_async + ukfb_get_state_async / "
                  "ukfb_get_mu_range_async), pinned host buffers, one host worker per device inside the library"}
    # device-timed: the kernels alone (events on every shard's stream, max over shards); inputs travel, outputs do not
    loop(lambda k: None, 3)
    f.event_record(0)
    loop(lambda k: None, steps)
    t_wall0 = time.perf_counter()
    f.event_record(1)
    f.synchronize()
    out["step_only_ms"] = f.event_elapsed_ms(0, 1) / steps
    out["step_only_value"] = total / (out["step_only_ms"] * 1e-3)
    for name, read in (("e2e_full_mu", lambda k: f.get_state_async(mu_pin[k & 1])),
                       ("e2e_pose_only", lambda k: f.get_mu_range_async(0, 7, pose_pin[k & 1]))):
        loop(read, 3)
        t0 = time.perf_counter()
        loop(read, steps)
        s = time.perf_counter() - t0
        out[name] = {"value": total * steps / s, "ms_per_step": s / steps * 1e3}
    # the final gather: every shard's estimates into ONE host buffer
    mu_all, sg_smp = pinned((total, 13)), None
    t0 = time.perf_counter()
    f.get_state_into(mu_all)
    out["gather_ms"] = (time.perf_counter() - t0) * 1e3
    out["gather_bytes"] = int(total * 13 * 8)
    # a strided sample against a one-device handle that took the same steps (bitwise) -- and the oracle (1e-9)
    S = 64
    idx = (np.arange(S) * (total // S)).astype(np.int64)
    m_s, s_s, z_s, R_s = bench.make_workload_of(idx, pool)
    g = UkfBatch(0, S, device=0)
    g.initialize(m_s, s_s)
    for j in applied:
        g.step(syn.DT, 8, z_s[j], R_s)
    out["sample_bitwise_equal_to_one_device"] = bool(np.array_equal(g.get_state()[0], mu_all[idx]))
    out["launches"] = f.launch_count()
    f.close()
    g.close()
    return out


def main():
    import torch

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    n = args.gpus or torch.cuda.device_count()
    for g in sorted({1, n}):
        print(json.dumps(run(g, g << 20, args.steps, "weak: 1 Mi filters per GPU")), flush=True)
    if n > 1:
        print(json.dumps(run(n, 1 << 20, args.steps * 4, "strong: 1 Mi filters in total")), flush=True)


if __name__ == "__main__":
    main()
