"""Development aid: is k_step slower back-to-back than isolated? (needs a PROFILE=1 build)"""
import ctypes, time
import numpy as np
import parity
from parity import make_config
import torch
from masurvival import _lib

rec = make_config('2v2', auto_reset=True)
N, A = 16384, 4
L = _lib.load()
L.msv_debug_profile.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
hs = [_lib.Handle(rec, N, 0, 1 + r, r * N) for r in range(4)]
for h in hs: h.reset()
acts = torch.randint(0, 2, (8, N, A, 6), dtype=torch.uint8, device='cuda')
acts[..., 0:3] = torch.randint(0, 3, (8, N, A, 3), dtype=torch.uint8, device='cuda')
for t in range(200): hs[t % 4].step(acts[t % 8].data_ptr())
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 32)()

def run(mode, steps=40, rot=4):
    L.msv_debug_profile(hs[0].h, 1, buf)
    for h in hs: L.msv_debug_profile(h.h, 1, None)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for t in range(steps):
        if mode == 'gap': torch.cuda.synchronize(); time.sleep(0.002)
        ev[t][0].record()
        hs[t % rot].step_kernel_only(acts[t % 8].data_ptr())
        ev[t][1].record()
        hs[t % rot].observe_only()
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    L.msv_debug_profile(hs[0].h, 0, buf)
    tot = sum(buf[:12]) / N / steps
    names = ['load', 'pre_step', 'find_new', 'collide', 'solve', 'toi', 'post_boxes', 'cameras', 'post_rest', 'rewards+reset', 'TOI_SCAN', 'TOI_EVENT']
    print('   slowest thread: env', buf[13], 'toi_calls', buf[14], 'toi_guard_iters', buf[15], {n: buf[16 + i] for i, n in enumerate(names)})
    print('   TOI calls', buf[28], 'max outer iters', buf[29], 'max root iters/call', buf[30], 'max cycles/call', buf[31])
    print(f'{mode:14s} rot={rot}: k_step ms mean={ms.mean():.3f} min={ms.min():.3f} max={ms.max():.3f}; avg thread cycles={tot:.0f}, max thread cycles={buf[12]}, '
          f'implied clock if kernel==slowest thread: {buf[12] / ms.max() / 1e6:.2f} GHz')

run('back-to-back'); run('back-to-back', rot=1); run('back-to-back', rot=1); run('back-to-back', rot=1)
