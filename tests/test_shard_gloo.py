"""world_size-2 CPU (gloo) test of the host-side sharding logic: contiguous filter ranges, no collective on the
step path, one gather of the estimates at the end, and per-filter results that do not depend on the sharding
(SURVEY.md section 8e).  The per-shard arithmetic here is the CPU oracle -- this tests the plumbing, not the kernels."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity as P
from oracle.oracle_lib import OracleBatch
from slam_pose_estimation_b200.shard import gather_estimates, shard_range, shard_sizes

TOTAL, STEPS = 37, 12


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 37, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(total, world)
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, last = shard_range(TOTAL, rank, world)
        f = P.make_pose(OracleBatch, last - first, first=first, threads=1)
        P.run_pose_c3(f, last - first, STEPS, first=first)  # no communication on the step path
        mu, sg = f.get_state()
        rows = torch.from_numpy(np.concatenate([mu, sg.reshape(mu.shape[0], -1)], axis=1))
        full = gather_estimates(rows, TOTAL)
        if rank == 0:
            np.save(out_path, full.numpy())
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_run_matches_the_unsharded_batch(tmp_path):
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    whole = P.make_pose(OracleBatch, TOTAL, threads=1)
    P.run_pose_c3(whole, TOTAL, STEPS)
    mu, sg = whole.get_state()
    want = np.concatenate([mu, sg.reshape(TOTAL, -1)], axis=1)
    assert got.shape == want.shape
    assert np.array_equal(got, want)  # bitwise: a filter's result does not depend on its shard
