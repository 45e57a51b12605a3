"""Small target for ncu captures: pre-roll a batch to the stationary episode mix (random policy, 61 action
batches in scrambled order as in bench.py), then a few more steps.
    python tests/gpu_ncu_target.py [variant] [envs] [preroll steps] [extra steps]"""
import sys
import parity
from parity import make_config
import torch
from masurvival import _lib
v = sys.argv[1] if len(sys.argv) > 1 else '2v2'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
extra = int(sys.argv[4]) if len(sys.argv) > 4 else 6
rec = make_config(v, auto_reset=True)
A = int(rec['n_agents'])
h = _lib.Handle(rec, N, 0, 1, 0)
h.reset()
g = torch.Generator(device='cuda'); g.manual_seed(5)
NB = 61
acts = torch.randint(0, 2, (NB, N, A, 6), dtype=torch.uint8, device='cuda', generator=g)
acts[..., 0:3] = torch.randint(0, 3, (NB, N, A, 3), dtype=torch.uint8, device='cuda', generator=g)
for t in range(warm + extra):
    h.step(acts[(t * 7) % NB].data_ptr())
torch.cuda.synchronize()
print('ok', h.overflow_events(), h.flush_stats()['episodes'])
