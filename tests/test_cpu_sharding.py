"""Multi-GPU host logic on CPU: world_size-2 gloo.  Each rank owns a shard of
the global env batch (no collective on the step path); results must be the
ones a single process computes, and flush_stats sums across ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity


def _worker(rank, world, port, total, steps, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import pyoracle as po
    from masurvival.distributed import all_reduce_stats, shard
    rec = parity.make_config('2v2', auto_reset=True)
    first, count = shard(total, rank, world)
    envs = [po.OracleEnv(rec, seed=9, env_id=first + k) for k in range(count)]
    rng = np.random.default_rng(0)
    acts = parity.random_actions(rng, total * steps, 4).reshape(steps, total, 4, 6)
    rew = np.zeros((steps, count, 4), np.float32)
    for e in envs:
        e.reset()
    for t in range(steps):
        for k, e in enumerate(envs):
            rew[t, k] = e.step(acts[t, first + k])['rewards']
    local = {'steps': 0, 'heals_used': 0, 'boxes_placed': 0, 'episodes': 0, 'reward0': 0.0, 'reward1': 0.0, 'kills0': 0, 'kills1': 0}
    for e in envs:
        s = e.flush_stats()
        local['steps'] += int(s['steps']); local['heals_used'] += int(s['heals_used']); local['boxes_placed'] += int(s['boxes_placed'])
        local['episodes'] += int(s['episodes']); local['reward0'] += float(s['reward'][0]); local['reward1'] += float(s['reward'][1])
        local['kills0'] += int(s['kills'][0]); local['kills1'] += int(s['kills'][1])
    tot = all_reduce_stats(local)
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, count, rew))
    if rank == 0:
        ret['stats'] = tot
        ret['rew'] = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])], axis=1)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_covers_range():
    from masurvival.distributed import shard
    for total in (1, 7, 16, 16384, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1


def test_two_rank_gloo_matches_single_process():
    total, steps, world = 12, 40, 2
    mgr = mp.Manager(); ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, total, steps, ret), nprocs=world, join=True)
    import pyoracle as po
    rec = parity.make_config('2v2', auto_reset=True)
    envs = [po.OracleEnv(rec, seed=9, env_id=k) for k in range(total)]
    rng = np.random.default_rng(0)
    acts = parity.random_actions(rng, total * steps, 4).reshape(steps, total, 4, 6)
    rew = np.zeros((steps, total, 4), np.float32)
    for e in envs:
        e.reset()
    for t in range(steps):
        for k, e in enumerate(envs):
            rew[t, k] = e.step(acts[t, k])['rewards']
    assert np.array_equal(rew, ret['rew'])
    assert ret['stats']['steps'] == total * steps
    assert abs(ret['stats']['reward0'] - sum(float(e.flush_stats()['reward'][0]) for e in envs)) < 1e-6
