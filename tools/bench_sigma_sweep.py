#!/usr/bin/env python
"""tools/bench_sigma_sweep.py -- throughput of the PoseUKF step against the orientation uncertainty.

The fast kernel evaluates SO(3) exp / log with short polynomials that are valid for rotations up to 0.58 rad; sigma points
of a filter whose orientation standard deviation is larger leave that range.  This tool measures what that costs: the
bench.py step (predictionStep(1 ms) + AngularVelocityMeasurement update, which carries no orientation information, so the
orientation uncertainty stays where it was set) on B filters whose initial orientation covariance is sigma^2 I,
  (a) for sigma from 0.05 rad to pi (an unknown heading: the usual start-up state of a pose filter), all filters alike;
  (b) for a fraction f of wide filters (sigma = 1 rad) spread evenly among narrow ones (0.1 rad): lanes of one warp
      that take different paths.
One JSON line per point: filter-steps/s, mean passes per state mean, flagged filters, parity of a strided sample against
the CPU oracle (mean and covariance, tests/parity.py metrics).

    python tools/bench_sigma_sweep.py [--filters B] [--steps K]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def workload(B, sigma_ori):
    """bench.py's C4 workload with the orientation block of the initial covariance set per filter"""
    from slam_pose_estimation_b200 import synthetic as syn

    sigma_ori = np.broadcast_to(np.asarray(sigma_ori, float), (B,))
    mu = np.zeros((B, 13))
    mu[:, 6] = 1.0
    mu[:, 7:10] = syn.POSE_V_TRUE
    mu[:, 10:13] = syn.POSE_W_TRUE
    d = np.tile(np.array([1.0] * 3 + [0.01] * 3 + [0.1] * 3 + [0.01] * 3), (B, 1))
    d[:, 3:6] = (sigma_ori**2)[:, None]
    sg = np.zeros((B, 12, 12))
    sg[:, np.arange(12), np.arange(12)] = d
    e = syn.noise(np.arange(B), 0, 15, 12) * np.sqrt(d)
    mu[:, 0:3] += e[:, 0:3]
    half = 0.5 * e[:, 3:6]
    ang = np.linalg.norm(half, axis=1, keepdims=True)
    mu[:, 3:7] = np.concatenate([np.sinc(ang / np.pi) * half, np.cos(ang)], axis=1)
    mu[:, 7:10] += e[:, 6:9]
    mu[:, 10:13] += e[:, 9:12]
    zs = np.stack([syn.pose_measurement(8, B, j + 1)[0] for j in range(4)])
    R = np.eye(3) * syn.SIGMA_GYRO**2
    return mu, sg, zs, R


def run(B, sigma_ori, steps, label):
    import torch

    import parity as P
    from oracle.oracle_lib import OracleBatch
    from slam_pose_estimation_b200 import synthetic as syn
    from slam_pose_estimation_b200.batch import UkfBatch

    mu, sg, zs, R = workload(B, sigma_ori)
    f = UkfBatch(0, B)
    f.initialize(mu, sg)
    dev = torch.device("cuda", 0)
    d_dt = torch.full((1,), syn.DT, dtype=torch.float64, device=dev)
    d_R = torch.from_numpy(R).to(dev)
    d_z = [torch.from_numpy(z).to(dev) for z in zs]
    torch.cuda.synchronize()
    warm = 3
    for k in range(warm):
        f.step_dev(d_dt, False, 8, d_z[k % 4], d_R, False)
    f.clear_mean_iter_hist()
    f.synchronize()
    f.event_record(0)
    for k in range(steps):
        f.step_dev(d_dt, False, 8, d_z[(warm + k) % 4], d_R, False)
    f.event_record(1)
    f.synchronize()
    ms = f.event_elapsed_ms(0, 1) / steps
    hist = f.get_mean_iter_hist()
    S = 32
    idx = (np.arange(S) * (B // S)).astype(int)
    o = OracleBatch(0, S)
    o.initialize(mu[idx], sg[idx])
    for k in range(warm + steps):
        o.step(syn.DT, 8, zs[k % 4][idx], R)
    mg, sgg = f.get_state()
    mo, so = o.get_state()
    out = {"label": label, "filters": B, "steps": steps, "ms_per_step": ms, "value": B / (ms * 1e-3), "unit": "filter-steps/s",
           "mean_passes_avg": float((hist * np.arange(8)).sum() / max(1, hist.sum())), "status_flagged": int(f.status_summary()[0]),
           "max_mu_err": float(P.mu_error(0, mg[idx], mo).max()), "max_sigma_err": float(P.sigma_error(sgg[idx], so).max()),
           "oracle_status_flagged": int((o.get_status() != 0).sum())}
    f.close()
    return out


def run_ori(B, sigma_ori, steps, label, K=10):
    """the same question for OrientationUKF: K IMU ticks (sample stored + predict) per launch, initial attitude covariance
    sigma^2 I (the predicts carry no attitude information either, so it stays wide)"""
    import torch

    import parity as P
    from oracle.oracle_lib import OracleBatch
    from slam_pose_estimation_b200 import synthetic as syn
    from slam_pose_estimation_b200.batch import UkfBatch

    mu, sg = syn.orientation_initial(B)
    sg[:, 0:3, 0:3] = np.eye(3) * sigma_ori**2
    f = P.make_ori(UkfBatch, 1)  # parameters as the tests set them
    f.close()
    f = UkfBatch(1, B)
    f.initialize(mu, sg)
    f.set_process_noise(syn.ORI_Q)
    f.set_orientation_params(syn.ORI_TAU, syn.ORI_TAU, syn.LATITUDE_BREMEN)
    dev = torch.device("cuda", 0)
    imu = np.stack([np.concatenate(syn.orientation_imu(B, k + 1), axis=1) for k in range(K)])
    d_imu = torch.from_numpy(imu).to(dev)
    d_dt = torch.full((K,), syn.DT, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    warm = 2
    for _ in range(warm):
        f.run_dev(K, d_dt, False, d_imu=d_imu)
    f.clear_mean_iter_hist()
    f.synchronize()
    f.event_record(0)
    for _ in range(steps):
        f.run_dev(K, d_dt, False, d_imu=d_imu)
    f.event_record(1)
    f.synchronize()
    ms = f.event_elapsed_ms(0, 1) / steps
    hist = f.get_mean_iter_hist()
    S = 16
    idx = (np.arange(S) * (B // S)).astype(int)
    o = OracleBatch(1, S)
    o.initialize(mu[idx], sg[idx])
    o.set_process_noise(syn.ORI_Q)
    o.set_orientation_params(syn.ORI_TAU, syn.ORI_TAU, syn.LATITUDE_BREMEN)
    for _ in range(warm + steps):
        for k in range(K):
            o.set_rotation_rate(imu[k, idx, 0:3])
            o.set_acceleration(imu[k, idx, 3:6])
            o.predict_dt(syn.DT)
    mg, sgg = f.get_state()
    mo, so = o.get_state()
    out = {"label": label, "filter": "OrientationUKF", "filters": B, "ticks_per_launch": K, "launches": steps, "ms_per_launch": ms,
           "value": B * K / (ms * 1e-3), "unit": "filter-ticks/s",
           "mean_passes_avg": float((hist * np.arange(8)).sum() / max(1, hist.sum())), "status_flagged": int(f.status_summary()[0]),
           "max_mu_err": float(P.mu_error(1, mg[idx], mo).max()), "max_sigma_err": float(P.sigma_error(sgg[idx], so).max()),
           "sigma_ori": sigma_ori}
    f.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filters", type=int, default=1 << 18)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--sigmas", type=float, nargs="*", help="only these points of sweep (a), and no sweep (b)")
    ap.add_argument("--orientation", action="store_true", help="sweep (a) on OrientationUKF instead")
    args = ap.parse_args()
    B = args.filters
    if args.orientation:
        for s in args.sigmas or (0.1, 0.3, 0.6, 1.0, 2.0, 3.0):
            print(json.dumps(run_ori(B, s, args.steps, f"OrientationUKF, all filters sigma_ori = {s:.3f} rad")), flush=True)
        return
    for s in args.sigmas or (0.05, 0.1, 0.2, 0.3, 0.45, 0.6, 0.8, 1.0, 1.5, 2.0, 2.5, 3.0, float(np.pi)):
        r = run(B, s, args.steps, f"all filters sigma_ori = {s:.3f} rad")
        r["sigma_ori"] = s
        print(json.dumps(r), flush=True)
    for frac in () if args.sigmas else (0.0, 1 / 32, 1 / 8, 1 / 4, 1 / 2, 1.0):
        period = int(round(1 / frac)) if frac else 0
        sig = np.full(B, 0.1)
        if period:
            sig[::period] = 1.0
        r = run(B, sig, args.steps, f"fraction {frac:.4f} of the filters (every {period or 'none'}th lane) at 1 rad, the rest at 0.1 rad")
        r["wide_fraction"] = frac
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
