#!/usr/bin/env python
"""`ref-python-on-shim` CPU baseline.

Times the reference's OWN, UNMODIFIED `masurvival` package (imported from /root/reference) on the
repo's float32 Box2D/gym shims (oracle/shim), with the loop and the timer of the reference's
demo.py:120-147 -- `action_space.sample()` per step, `time.process_time()` around `env.step`
only -- except that the env is reset on done so that exactly `--steps` steps are timed.  One
process per host core (`os.cpu_count()`), aggregate agent-steps/s = sum over processes.

/root/reference only exists in the build container, so this script runs HERE and its output is
committed as profiles/ref_python_on_shim.json; bench.py copies the matching entry into its JSON
line as `cpu_baseline_ref_python` (never presented as a pybox2d number).

    python tests/golden/time_reference.py [--steps 1000] [--procs N]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def worker(args):
    name, steps, seed = args
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shim'))
    import numpy as np
    import parity
    import Box2D, gym  # noqa: F401,E401  (the shims)
    user = parity.variant({'1v1_default': '1v1'}.get(name, name))
    user.pop('lidars', None)
    for m in [k for k in sys.modules if k == 'masurvival' or k.startswith('masurvival.')]:
        del sys.modules[m]
    sys.path = [p for p in sys.path if 'gym-ma-survival-2d_b200' not in p]
    sys.path.insert(0, '/root/reference')
    from masurvival.envs.masurvival_env import MaSurvival
    import masurvival
    assert masurvival.__file__.startswith('/root/reference')
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):       # "Splitting agents into 2 teams."
        env = MaSurvival(None if name == '1v1_default' else user)
    env.np_random = np.random.default_rng(seed)
    env.action_space.seed(seed) if hasattr(env.action_space, 'seed') else None
    env.reset()
    times, episodes = [], 0
    for _ in range(steps):
        action = env.action_space.sample()                # demo.py:22
        t0 = time.process_time()
        obs, reward, done, info = env.step(action)        # demo.py:135-137
        t1 = time.process_time()
        times.append(t1 - t0)
        if done:
            episodes += 1
            env.reset()
    return sum(times), len(times), episodes, len(env.simulation.groups['agents'].get(
        __import__('masurvival.simulation').simulation.IndexBodies)[0].bodies)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--procs', type=int, default=os.cpu_count() or 1)
    ap.add_argument('--configs', nargs='*', default=['1v1_default', '2v2', '1v1_heal_only', 'ffa'])
    a = ap.parse_args()
    out = {}
    for name in a.configs:
        t0 = time.perf_counter()
        with mp.get_context('spawn').Pool(a.procs) as pool:
            res = pool.map(worker, [(name, a.steps, 100 + p) for p in range(a.procs)])
        wall = time.perf_counter() - t0
        A = res[0][3]
        per_proc = [n / s for s, n, _, _ in res]          # env-steps per CPU-second, each process
        agg = sum(per_proc) * A
        mean_step = sum(s for s, _, _, _ in res) / sum(n for _, n, _, _ in res)
        out[name] = {'value': agg, 'unit': 'agent-steps/s', 'cores': a.procs, 'kind': 'ref-python-on-shim',
                     'sample': f'{a.procs} processes x {a.steps} steps, 1 env each, action_space.sample(), reset on done; '
                               'time.process_time() around env.step only (demo.py:135-137)',
                     'env_steps_per_sec_per_core': sum(per_proc) / len(per_proc), 'ms_per_step_mean': mean_step * 1e3,
                     'agents_per_env': A, 'episodes': sum(e for _, _, e, _ in res), 'wall_seconds': wall,
                     'where': 'build container (%d cores); /root/reference is not available on the GPU box' % (os.cpu_count() or 1),
                     'what': "the reference's unmodified Python (masurvival/*.py from /root/reference) on the float32 Box2D shim of oracle/shim -- NOT pybox2d"}
        print(name, json.dumps(out[name]))
    # bench.py workloads: '2v2' -> 2v2, etc.; config 1 keeps both 1v1 default and 2v2
    with open(os.path.join(ROOT, 'profiles', 'ref_python_on_shim.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
