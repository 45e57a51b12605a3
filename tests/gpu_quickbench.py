"""Quick device-side timing of the step kernel (development aid; bench.py is
the contract benchmark)."""
import sys, time
import numpy as np
import parity
from parity import make_config
import torch
from masurvival import _lib

def run(variant, N, steps=200, warm=50, prof=True):
    rec = make_config(variant, auto_reset=True)
    A = int(rec['n_agents'])
    h = _lib.Handle(rec, N, 0, 1, 0)
    h.reset()
    torch.manual_seed(1234)   # identical action streams in every run: timing differences come from the code only
    acts = torch.randint(0, 2, (8, N, A, 6), dtype=torch.uint8, device='cuda')
    acts[..., 0:3] = torch.randint(0, 3, (8, N, A, 3), dtype=torch.uint8, device='cuda')
    for t in range(warm): h.step(acts[t % 8].data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps): h.step(acts[t % 8].data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    bps = h.bytes_per_env_step()
    st = h.flush_stats()
    print(f'{variant} N={N}: {ms*1e3:.1f} us/step, {N/ms*1e3:.3e} env-steps/s, {N*A/ms*1e3:.3e} agent-steps/s, '
          f'{bps} B/env-step -> {N*bps/ms/1e6:.1f} GB/s, episodes={int(st["episodes"])}')
    if prof:
        import ctypes
        L = _lib.load()
        L.msv_debug_profile.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        buf = (ctypes.c_ulonglong * 32)()
        L.msv_debug_profile(h.h, 1, buf)
        for t in range(20): h.step(acts[t % 8].data_ptr())
        L.msv_debug_profile(h.h, 0, buf)
        print('   max over threads (any of 20 launches): total=%d; ' % buf[12] + ', '.join(f'{n}={buf[16+i]}' for i, n in enumerate(['load', 'pre_step', 'find_new', 'collide', 'solve', 'toi', 'post_boxes', 'cameras', 'post_rest', 'rewards+reset', 'observe', 'store'])))
        names = ['load', 'pre_step', 'find_new', 'collide', 'solve', 'toi', 'post_boxes', 'cameras', 'post_rest', 'rewards+reset', 'observe', 'store']
        tot = sum(buf[:12]) or 1
        print('   phase cycles/thread/step: ' + ', '.join(f'{n}={buf[i]/N/20:.0f} ({buf[i]/tot:.0%})' for i, n in enumerate(names)))
    h.close()

if __name__ == '__main__':
    for v, N in (('2v2', 16384), ('1v1_heal_only', 4096), ('ffa', 8192), ('ffa_lidar', 32768)):
        run(v, N)
