"""Multi-GPU host logic on CPU (gloo, world_size 2): the batch shards as contiguous blocks with global-index
seeds and NO data-path collective; the only cross-rank traffic is a scalar gather/max of timings and counters.
The shards are checked with the oracle: concatenating per-rank shards reproduces the single-process batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minigrid_b200 import shard_range
from oracle.oracle import OracleVecEnv

ENV_ID, TOTAL, STEPS = "MiniGrid-LavaCrossingS9N1-v0", 203, 40


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_range(TOTAL, rank, world)
    env = OracleVecEnv(ENV_ID, count)
    obs, _ = env.reset(seed=first + 1000)  # what MinigridVecEnv(seed_offset=first).reset(seed=1000) seeds
    actions = np.random.default_rng(7).integers(0, 7, (STEPS, TOTAL)).astype(np.int32)
    trace = [obs.copy()]
    ended = 0
    for t in range(STEPS):
        o, d, r, te, tr = env.step(actions[t, first:first + count])
        trace.append(o.copy())
        ended += int((te | tr).sum())
    np.save(os.path.join(out_dir, f"trace_{rank}.npy"), np.stack(trace))
    # the only collectives of a run: max over ranks of the elapsed time, sum of counters
    t_ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    n_end = torch.tensor([ended], dtype=torch.int64)
    dist.all_reduce(n_end, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), np.array([t_ms.item(), n_end.item()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    full = OracleVecEnv(ENV_ID, TOTAL)
    obs, _ = full.reset(seed=1000)
    actions = np.random.default_rng(7).integers(0, 7, (STEPS, TOTAL)).astype(np.int32)
    trace = [obs.copy()]
    ended = 0
    for t in range(STEPS):
        o, d, r, te, tr = full.step(actions[t])
        trace.append(o.copy())
        ended += int((te | tr).sum())
    trace = np.stack(trace)
    parts = [np.load(tmp_path / f"trace_{r}.npy") for r in range(world)]
    np.testing.assert_array_equal(np.concatenate(parts, axis=1), trace)
    t_ms, n_end = np.load(tmp_path / "reduced.npy")
    assert t_ms == 11.0 and n_end == ended
