#!/usr/bin/env python
"""Small calls through every kernel family of the C ABI, meant to run under compute-sanitizer on the GPU box:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
Ragged batch sizes (37, 70) exercise the lanes past the end of the last tile."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import parity as P  # noqa: E402
from slam_pose_estimation_b200 import UkfBatch, synthetic as syn  # noqa: E402


def main():
    for kernel in ("fast", "thread", "warp"):
        os.environ["UKFB_KERNEL"] = kernel
        B = 37
        g = P.make_pose(UkfBatch, B)
        P.run_pose_c3(g, B, 3)
        for kind in range(9):
            z, R = syn.pose_measurement(kind, B, 5 + kind)
            g.step(0.02, kind, z, R)
        o = P.make_ori(UkfBatch, B)
        P.run_ori_c1(o, B, 4, every=2)
        if kernel != "warp":
            ts, kinds, mu3 = syn.pose_c5_events(B, 1, 6, dvl_period=3, gps_period=4)
            e = P.make_pose(UkfBatch, B)
            e.set_mahalanobis_gate(25.0)
            e.run_events(ts, kinds, mu3, syn.sensor_cov_table())
            assert np.isfinite(e.get_state()[0]).all()
        rbs = g.get_body_states()
        h = UkfBatch(0, B)
        h.initialize_from_body_states(rbs)
        assert np.isfinite(g.get_state()[1]).all() and np.isfinite(o.get_state()[1]).all() and np.isfinite(h.get_state()[0]).all()
        print(kernel, "ok", g.launch_count() + o.launch_count(), "launches", flush=True)
    os.environ.pop("UKFB_KERNEL")


if __name__ == "__main__":
    main()
