#!/usr/bin/env python
"""tools/bench_c5.py -- BASELINE.json config 5 on one B200: PoseUKF with asynchronous IMU / DVL / GPS sample queues
(IMU 1 kHz -> AngularVelocity, DVL 10 Hz -> Velocity, GPS 1 Hz -> XY in the nav plane), batch size swept 1 .. 1 Mi.

Every launch of ukfb_run_events_dev integrates a window of `ticks` IMU ticks (K queue slots) for all filters; the
device-timed figure has the queues resident in HBM, the e2e figure goes through ukfb_run_events with host arrays
(H2D of timestamps, kinds and samples inside the timed region) plus the D2H of the estimates.  One JSON line per batch
size on stdout.  Not the contract bench (bench.py is); its output is kept under profiles/.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,32,1024,32768,65536,1048576")
    ap.add_argument("--ticks", type=int, default=20)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch

    from slam_pose_estimation_b200 import UkfBatch, synthetic as syn

    dev = torch.device("cuda:0")
    tab = syn.sensor_cov_table()
    d_tab = torch.from_numpy(tab).to(dev)
    for B in [int(x) for x in args.batches.split(",")]:
        ticks = args.ticks
        ts, kinds, mu3 = syn.pose_c5_events(B, 1, ticks)
        K = ts.shape[0]
        live = int((kinds != syn.EVENT_IDLE).sum())
        mu, sg = syn.pose_initial(B, perturb=True)
        f = UkfBatch(0, B)
        f.initialize(mu, sg)
        d_ts, d_kinds, d_mu3 = (torch.from_numpy(a).to(dev) for a in (ts, kinds, mu3))
        period = ticks * 1000
        stream = torch.cuda.ExternalStream(f.stream())

        def launch():
            f.run_events_dev(K, d_ts, d_kinds, d_mu3, d_tab, per_event=False)
            with torch.cuda.stream(stream):
                d_ts.add_(period)  # the next window: same samples, later timestamps

        torch.cuda.synchronize()
        for _ in range(args.warmup):
            launch()
        f.synchronize()
        t_dev = []
        for r in range(args.reps):
            f.event_record(0)
            f.run_events_dev(K, d_ts, d_kinds, d_mu3, d_tab, per_event=False)
            f.event_record(1)
            with torch.cuda.stream(stream):
                d_ts.add_(period)
            f.synchronize()
            t_dev.append(f.event_elapsed_ms(0, 1))
        ms = float(np.median(t_dev))
        # the same as a STREAM of windows, as a host that keeps the queues coming makes them: the timestamp arrays of the
        # next windows are on the device beforehand, so nothing but the launches is in the stream (consecutive launches
        # of a handle of a few waves then overlap at their ends: ukfb_overlapped_launch_count)
        nstream = 5
        with torch.cuda.stream(stream):
            d_ts_next = [d_ts + i * period for i in range(nstream)]
        f.synchronize()
        f.event_record(2)
        for i in range(nstream):
            f.run_events_dev(K, d_ts_next[i], d_kinds, d_mu3, d_tab, per_event=False)
        f.event_record(3)
        f.synchronize()
        ms_stream = f.event_elapsed_ms(2, 3) / nstream
        with torch.cuda.stream(stream):
            d_ts.add_(nstream * period)
        del d_ts_next
        f.synchronize()
        flagged, bits = f.status_summary()
        # end to end: host queues in, estimates out.  Blocking calls, then the streaming ones (the next window's queues are
        # copied while the current window is integrated; estimates leave on the copy-out stream).  Two host copies of the
        # queues alternate so that advancing the timestamps of one never touches a copy in flight.
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
        ts_now = d_ts.cpu().numpy()
        kinds_h, mu3_h = pin(kinds), pin(mu3)
        nrep = 4
        ts_h = [pin(ts_now + i * period) for i in range(nrep)]  # one host copy per window: nothing in flight is modified
        out = [pin(np.empty((B, 13))), pin(np.empty((B, 13)))]
        t_blk = []
        for r in range(nrep):
            t0 = time.perf_counter()
            f.run_events(ts_h[r], kinds_h, mu3_h, tab)
            f.get_state_into(out[0])
            t_blk.append((time.perf_counter() - t0) * 1e3)
        ms_blk = float(np.median(t_blk))
        f.synchronize()
        for a in ts_h:
            a += nrep * period
        for r in range(2):  # warm-up: the pipeline's staging slots are allocated on first use
            f.run_events_async(ts_h[r], kinds_h, mu3_h, tab)
            f.get_state_async(out[r & 1])
        f.synchronize()
        ts_h[0] += nrep * period
        ts_h[1] += nrep * period
        order = [2, 3, 0, 1]
        t0 = time.perf_counter()
        for r in range(nrep):
            f.run_events_async(ts_h[order[r]], kinds_h, mu3_h, tab)
            f.get_state_async(out[r & 1])
        f.synchronize()
        ms_e2e = (time.perf_counter() - t0) * 1e3 / nrep
        print(json.dumps({
            "workload": "C5: PoseUKF, per-filter queues of IMU (kind 8, 1 kHz) / DVL (kind 4, 10 Hz) / GPS-XY (kind 1, 1 Hz) samples",
            "filters": B, "ticks_per_launch": ticks, "slots_per_launch": K, "samples_per_launch": live,
            "launch_ms": ms, "samples_per_s": live / ms * 1e3, "filter_ticks_per_s": B * ticks / ms * 1e3,
            "us_per_sample_per_filter": ms * 1e3 / K,
            "stream_of_launches_ms": ms_stream, "stream_samples_per_s": live / ms_stream * 1e3,
            "launches_overlapped": f.overlapped_launch_count() > 0,
            "e2e_ms": ms_e2e, "e2e_samples_per_s": live / ms_e2e * 1e3, "e2e_api": "ukfb_run_events_async + ukfb_get_state_async, pinned host arrays",
            "e2e_blocking_ms": ms_blk, "e2e_blocking_samples_per_s": live / ms_blk * 1e3,
            "e2e_h2d_bytes": int(ts_h[0].nbytes + kinds_h.nbytes + mu3_h.nbytes + tab.nbytes), "e2e_d2h_bytes": int(out[0].nbytes),
            "status_flagged": int(flagged), "status_bits": int(bits), "gpu_launches": int(f.launch_count()),
        }), flush=True)
        f.close()


if __name__ == "__main__":
    main()
