#!/usr/bin/env python
"""tools/bench_c2.py -- BASELINE.json configs 2 and 3 on one B200 (secondary figures; bench.py is the contract bench).

C2: 65,536 OrientationUKF on synthetic 1 kHz IMU data: every tick stores the IMU sample and predicts, every 100th tick
also integrates a body-velocity measurement.  C3: 65,536 PoseUKF, predict + AngularVelocity update every tick, velocity
every 10th, position every 100th.  K ticks advance per launch (ukfb_run_dev) with the state on chip; the figure is
device-timed filter-ticks/s.  One JSON line per case.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="65536,1048576")
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cases", default="c2,c3")
    ap.add_argument("--update-every-tick", action="store_true", help="C2 with a velocity update on every tick")
    args = ap.parse_args()
    import torch

    from slam_pose_estimation_b200 import UkfBatch, synthetic as syn

    dev = torch.device("cuda:0")
    K = args.K
    for case in args.cases.split(","):
        for B in [int(x) for x in args.batches.split(",")]:
            Kc = K if B <= 65536 else max(10, K // 10)
            d_dt = torch.full((Kc,), syn.DT, dtype=torch.float64, device=dev)
            if case == "c2":
                mu, sg = syn.orientation_initial(B)
                f = UkfBatch(1, B)
                f.set_orientation_params(syn.ORI_TAU, syn.ORI_TAU, syn.LATITUDE_BREMEN)
                f.initialize(mu, sg)
                f.set_process_noise(syn.ORI_Q)
                kinds = np.full(Kc, 9 if args.update_every_tick else -1, np.int8)
                kinds[Kc - 1] = 9
                imu = np.empty((4, B, 6))  # four distinct IMU sample sets, cycled over the ticks
                for j in range(4):
                    g, a = syn.orientation_imu(B, j + 1)
                    imu[j, :, :3], imu[j, :, 3:] = g, a
                d_imu = torch.from_numpy(imu).to(dev)[torch.arange(Kc, device=dev) % 4].contiguous()
                z = syn.orientation_velocity(B, 1)[0]
                d_z = torch.from_numpy(z).to(dev)[None].expand(Kc, B, 3).contiguous()
                R = np.eye(3) * syn.SIGMA_DVL**2
                updates = Kc if args.update_every_tick else 1
            else:
                mu, sg = syn.pose_initial(B, perturb=True)
                f = UkfBatch(0, B)
                f.initialize(mu, sg)
                kinds = np.full(Kc, 8, np.int8)
                d_imu = None
                zs = np.stack([syn.pose_measurement(8, B, j + 1)[0] for j in range(4)])
                d_z = torch.from_numpy(zs).to(dev)[torch.arange(Kc, device=dev) % 4].contiguous()
                R = np.eye(3) * syn.SIGMA_GYRO**2
                updates = Kc
            d_R = torch.from_numpy(np.tile(R, (Kc, 1, 1))).to(dev)
            torch.cuda.synchronize()
            for _ in range(2):
                f.run_dev(Kc, d_dt, False, kinds, d_z, d_R, False, d_imu)
            f.synchronize()
            # a stream of launches, as a caller feeding sensor data makes them (consecutive launches of a handle of a few
            # waves overlap at their ends: ukfb_overlapped_launch_count)
            f.event_record(0)
            for _ in range(args.reps):
                f.run_dev(Kc, d_dt, False, kinds, d_z, d_R, False, d_imu)
            f.event_record(1)
            f.synchronize()
            ms = f.event_elapsed_ms(0, 1) / args.reps
            flagged, bits = f.status_summary()
            print(json.dumps({
                "workload": ("C2: OrientationUKF, IMU store + predict per tick, velocity update on the last tick of the launch"
                             if case == "c2" else "C3-like: PoseUKF, predict + AngularVelocity update per tick"),
                "filters": B, "ticks_per_launch": Kc, "updates_per_launch": updates, "launch_ms": ms,
                "filter_ticks_per_s": B * Kc / ms * 1e3, "us_per_tick": ms * 1e3 / Kc,
                "status_flagged": int(flagged), "status_bits": int(bits),
                "mean_pass_hist": [int(v) for v in f.get_mean_iter_hist()],
                "launches_overlapped": f.overlapped_launch_count() > 0,
            }), flush=True)
            f.close()


if __name__ == "__main__":
    main()
