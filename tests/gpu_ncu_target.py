"""Small target for ncu captures: warm a 2v2 batch into steady state, then a few more steps."""
import sys
import numpy as np
import parity
from parity import make_config
import torch
from masurvival import _lib
v = sys.argv[1] if len(sys.argv) > 1 else '2v2'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 300
rec = make_config(v, auto_reset=True)
A = int(rec['n_agents'])
h = _lib.Handle(rec, N, 0, 1, 0)
h.reset()
g = torch.Generator(device='cuda'); g.manual_seed(5)
acts = torch.randint(0, 2, (8, N, A, 6), dtype=torch.uint8, device='cuda', generator=g)
acts[..., 0:3] = torch.randint(0, 3, (8, N, A, 3), dtype=torch.uint8, device='cuda', generator=g)
for t in range(warm + 4):
    h.step(acts[t % 8].data_ptr())
torch.cuda.synchronize()
print('ok', h.overflow_events())
