"""Small dense-interaction run for compute-sanitizer (memcheck): scrambled
2v2 / ffa states, a few steps, every kernel (reset, step, obs, lidar, stats)."""
import numpy as np
import torch
import parity
from parity import make_config, random_actions, scramble_state
import pyoracle as po
from masurvival import _lib

for variant, n, over in (('2v2', 96, {}), ('ffa_lidar', 32, {}), ('2v2', 64, {'observation': {'omniscent': False}})):
    rec = make_config(variant, auto_reset=True, **over)
    A = int(rec['n_agents'])
    h = _lib.Handle(rec, n, 0, 3, 0)
    h.reset()
    rng = np.random.default_rng(0)
    states = []
    for e in range(n):
        o = po.OracleEnv(rec, seed=3, env_id=e); o.reset()
        states.append(scramble_state(o.get_state(), rec, rng)); o.close()
    h.set_state(np.array(states))
    h.observe()
    for t in range(12):
        a = torch.as_tensor(random_actions(rng, n, A, 0.6, 0.5, 0.3)).cuda()
        h.step(a.data_ptr())
    torch.cuda.synchronize()
    print(variant, 'ok', h.flush_stats())
    h.close()
