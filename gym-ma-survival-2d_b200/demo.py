#!/usr/bin/env python
"""Headless counterpart of the reference's demo.py (demo.py:76-158, 164-293)
for the batched B200 environment: random-policy rollout, optional JSON config
(`-c`, same nested keys as the reference; shapes are given as radii), episode
statistics (`flush_stats`) and the `--benchmark` per-step timing line.
Rendering / interactive play are outside the accelerated path.

    python gym-ma-survival-2d_b200/demo.py -n 4096 --max-steps 1000 --benchmark
"""
import argparse
import json
import os
import pprint
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from masurvival.envs import MaSurvivalVec  # noqa: E402


class RandomPolicy:
    """demo.py:18-22: `action_space.sample()`, here for N envs on the device."""

    def __init__(self, env, seed=0):
        import torch
        self.env, self.g = env, torch.Generator(device=f'cuda:{env.device}')
        self.g.manual_seed(seed)

    def act(self, observations=None):
        import torch
        N, A, dev = self.env.num_envs, self.env.n_agents, f'cuda:{self.env.device}'
        a = torch.empty((N, A, 6), dtype=torch.uint8, device=dev)
        a[..., :3] = torch.randint(0, 3, (N, A, 3), dtype=torch.uint8, device=dev, generator=self.g)
        a[..., 3:] = torch.randint(0, 2, (N, A, 3), dtype=torch.uint8, device=dev, generator=self.g)
        return a


def demo_env(env, max_steps=None, print_benchmark=False, seed=0):
    """Runs until every env finished one episode (auto_reset off) or max_steps."""
    import torch
    policy = RandomPolicy(env, seed)
    times = []
    t, obs = 0, env.reset()
    done_once = torch.zeros(env.num_envs, dtype=torch.bool, device=f'cuda:{env.device}')
    while True:
        action = policy.act(obs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        obs, reward, done, info = env.step(action)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        done_once |= done
        t += 1
        if max_steps is not None and t == max_steps:
            print(f'Maximum number of steps {t} reached, terminating episode.')
            break
        if max_steps is None and bool(done_once.all()):
            break
    print('Episode complete. Stats printed below.')
    pprint.PrettyPrinter().pprint(env.flush_stats())
    if print_benchmark:
        times = np.array(times)
        print(f'Performance test results: {times.mean()}, {times.std()}')
        print(f'  = {env.num_envs * env.n_agents / times.mean():.4g} agent-steps/s over {env.num_envs} envs')
    env.close()


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('-c', '--config', help='JSON file with the (partial) nested config dict')
    ap.add_argument('-n', '--num-envs', type=int, default=1024)
    ap.add_argument('-s', '--max-steps', type=int, default=None)
    ap.add_argument('-b', '--benchmark', action='store_true', help='print mean/std of the per-step time (demo.py:274-279)')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--device', type=int, default=0)
    args = ap.parse_args()
    config = None
    if args.config is not None:
        with open(args.config) as f:
            config = json.load(f)
    env = MaSurvivalVec(config, num_envs=args.num_envs, device=args.device, seed=args.seed, auto_reset=args.max_steps is not None)
    demo_env(env, args.max_steps, args.benchmark, args.seed)


if __name__ == '__main__':
    main()
