#!/usr/bin/env python
"""tools/bench_mean_tol.py -- what the manifold-mean tolerance costs.

ukfom's sigma_points_mean iterates `while (norm(mean_delta) > tol && ++i < max_it)`; the tolerance (UKFB_MEAN_TOL = 1e-5 in
include/ukfb_constants.h) is a recollection of the un-vendored slam/mtk.  On the benchmark workload 1e-5 means ONE pass per
mean; should the Rock fork use a tighter value, every mean takes two.  This tool runs bench.py's step with the default
build and with a build of engine (lib/variants/tight.so, -DUKFB_MEAN_TOL=1e-9) AND oracle (oracle/build/liboracle_tight.so)
at the tighter value: throughput, passes per mean, and parity of a strided sample against the matching oracle.

    python tools/build_variant.py tight -DUKFB_MEAN_TOL=1e-9        (build container)
    python tools/bench_mean_tol.py                                  (GPU box)
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def run(variant: str):
    import numpy as np
    import torch

    import bench
    import parity as P
    from oracle.oracle_lib import OracleBatch
    from slam_pose_estimation_b200 import synthetic as syn
    from slam_pose_estimation_b200.batch import UkfBatch

    B, steps, warm, pool = 1 << 20, 20, 3, 4
    mu, sg, zs, R = bench.make_workload(B, 0, pool)
    f = UkfBatch(0, B)
    f.initialize(mu, sg)
    dev = torch.device("cuda", 0)
    d_dt = torch.full((1,), syn.DT, dtype=torch.float64, device=dev)
    d_R = torch.from_numpy(R).to(dev)
    d_z = [torch.from_numpy(z).to(dev) for z in zs]
    torch.cuda.synchronize()
    for k in range(warm):
        f.step_dev(d_dt, False, 8, d_z[k % pool], d_R, True)
    f.clear_mean_iter_hist()
    f.synchronize()
    f.event_record(0)
    for k in range(steps):
        f.step_dev(d_dt, False, 8, d_z[(warm + k) % pool], d_R, True)
    f.event_record(1)
    f.synchronize()
    ms = f.event_elapsed_ms(0, 1) / steps
    hist = f.get_mean_iter_hist()
    S = 64
    idx = (np.arange(S) * (B // S)).astype(int)
    o = OracleBatch(0, S, variant="left" if variant == "default" else "tight")
    o.initialize(mu[idx], sg[idx])
    for k in range(warm + steps):
        o.step(syn.DT, 8, zs[k % pool][idx], R[idx])
    mg, sgg = f.get_state()
    mo, so = o.get_state()
    print(json.dumps({"build": variant, "mean_tol": 1e-5 if variant == "default" else 1e-9, "filters": B, "ms_per_step": ms,
                      "value": B / (ms * 1e-3), "unit": "filter-steps/s",
                      "mean_passes_avg": float((hist * np.arange(8)).sum() / max(1, hist.sum())),
                      "max_mu_err": float(P.mu_error(0, mg[idx], mo).max()), "max_sigma_err": float(P.sigma_error(sgg[idx], so).max()),
                      "status_flagged": int(f.status_summary()[0])}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for v in ("default", "tight"):
            env = dict(os.environ)
            if v == "tight":
                env["UKFB_LIB"] = os.path.join(ROOT, "slam_pose_estimation_b200", "lib", "variants", "tight.so")
            subprocess.run([sys.executable, os.path.abspath(__file__), v], env=env, check=True)
