#!/usr/bin/env python
"""bench.py -- headline benchmark of the masurvival step path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the config the metric is quoted on):
default 2v2 (A=4, teams, melee cd 40, B=4, H=4, shrinking safe zone),
16384 environments PER GPU, random actions, auto-reset.  One "step" = one
MaSurvival.step of every environment of the batch (one kernel launch).
Environments shard trivially: one process per GPU, no collective on the data
path -> weak scaling.

Prints ONE JSON line (rank 0).  `value` = agent-steps/s with actions already
resident in HBM; `e2e` = the same through the host-buffer C-ABI call
(msv_step_host: H2D actions + D2H rewards/dones inside the timed region).
`--impl reference` times the CPU oracle port (the reference's own pybox2d
stack is not installable here) on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'gym-ma-survival-2d_b200'))

WORKLOAD = '2v2'
ENVS_PER_GPU = 16384
METRIC = 'agent_steps_per_sec'
UNIT = 'agent-steps/s'


def workload_config(auto_reset=True):
    from masurvival.config import merge_config, pack_config, variant
    cfg, cm = merge_config(variant(WORKLOAD))
    return pack_config(cfg, cm, auto_reset=auto_reset)


def config_block(n_gpus, envs_per_gpu, l2):
    return {'workload': 'configs[2]: default 2v2 (A=4 teams, melee cd40, B=4, H=4, safe zone), random actions, auto-reset',
            'envs_per_gpu': envs_per_gpu, 'global_envs': envs_per_gpu * n_gpus, 'agents_per_env': 4,
            'parallelism': f'{n_gpus} independent env shards (1 process/GPU, no collective)', 'l2': l2}


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def ncu_traffic():
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(p):
        try:
            return json.load(open(p)).get('dram_bytes_per_launch')
        except Exception:
            return None
    return None


class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu):
        self.gpu, self.p = gpu, None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--id={self.gpu}', f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill(); out = ''
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from masurvival.envs import MaSurvivalVec
    from masurvival.config import variant

    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    n_gpus = world
    N, A = args.envs, 4
    # ROT independent batches of N envs are stepped round-robin: their combined state +
    # outputs (ROT x ~48 MB) exceed the 126 MB L2, so every step finds its data in HBM
    # (timing rule: "inputs larger than L2") and no flush kernel perturbs the timed region.
    ROT = args.rotate
    envs = [MaSurvivalVec(variant(WORKLOAD), num_envs=N, device=local, seed=args.seed + r,
                          env_offset=(rank * ROT + r) * N, auto_reset=True) for r in range(ROT)]
    for e_ in envs:
        e_.reset()
    env = envs[0]
    dev = f'cuda:{local}'
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    NB = 8  # rotating pre-generated action batches

    def gen_actions(device, generator=None):
        a = torch.empty((NB, N, A, 6), dtype=torch.uint8, device=device)
        a[..., 0:3] = torch.randint(0, 3, (NB, N, A, 3), dtype=torch.uint8, device=device, generator=generator)
        a[..., 3:6] = torch.randint(0, 2, (NB, N, A, 3), dtype=torch.uint8, device=device, generator=generator)
        return a
    acts_dev = gen_actions(dev, g)
    acts_host = acts_dev.cpu().pin_memory()
    rew_host = torch.empty((N, A), dtype=torch.float32).pin_memory()
    done_host = torch.empty((N,), dtype=torch.uint8).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ------------------------------------------------
    for t in range(args.warmup):
        envs[t % ROT]._h.step(acts_dev[t % NB].data_ptr(), stream)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local); clk.start()
    barrier()
    l0 = sum(e_.kernel_launches() for e_ in envs)
    t0.record()
    for t in range(args.steps):
        envs[t % ROT]._h.step(acts_dev[t % NB].data_ptr(), stream)
    t1.record()
    barrier()
    launches = sum(e_.kernel_launches() for e_ in envs) - l0
    clocks = clk.stop()
    total_ms = float(t0.elapsed_time(t1))
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = n_gpus * N * A / (ms_per_step * 1e-3)

    # ---- duration of the dominant kernel (k_step), CUDA events on the launch stream ----
    ks = [torch.cuda.Event(enable_timing=True) for _ in range(64)]
    ke = [torch.cuda.Event(enable_timing=True) for _ in range(64)]
    for t in range(64):
        ks[t].record()
        envs[t % ROT]._h.step_kernel_only(acts_dev[t % NB].data_ptr(), stream)
        ke[t].record()
        envs[t % ROT]._h.observe_only(stream)
    torch.cuda.synchronize()
    kernel_ms = [s_.elapsed_time(e_) for s_, e_ in zip(ks, ke)]

    # ---- end-to-end arm: host buffers through msv_step_host ------------------
    for t in range(max(3, args.warmup // 4)):
        envs[t % ROT].step_host(acts_host[t % NB], rew_host, done_host)
    barrier()
    w0 = time.perf_counter()
    t0.record()
    for t in range(args.steps):
        envs[t % ROT].step_host(acts_host[t % NB], rew_host, done_host)   # H2D + kernels + D2H + stream sync
    t1.record()
    barrier()
    e2e_ms = float(t0.elapsed_time(t1))
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_ms = float(tt.item())
    e2e_value = n_gpus * N * A / (e2e_ms / args.steps * 1e-3)
    checksum = float(rew_host.sum())

    bytes_env = env.bytes_per_env_step()            # whole step (k_step + k_obs)
    obs_bytes = env.obs_bytes_per_env()             # written by k_obs
    kstep_bytes = bytes_env - obs_bytes + 8         # k_step: state r/w, actions, rewards, dones, 8 B mask word for k_obs
    peak, peak_src = measured_peak()
    kavg_ms = float(np.mean(kernel_ms[8:]))
    achieved = kstep_bytes * N / (kavg_ms * 1e-3) / 1e9
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_block(n_gpus, N, f'{ROT} batches of {N} envs stepped round-robin (state+outputs {ROT}x~48 MB > 126 MB L2): inputs larger than L2, no flush'),
        'env_steps_per_sec': value / A,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': N * A * 6, 'd2h_bytes_per_step': N * A * 4 + N,
                'api': 'MaSurvivalVec.step_host -> msv_step_host (pinned host buffers)', 'reward_checksum': checksum},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': ncu_traffic(), 'kernel': 'k_step<4,4,4,4>', 'kernel_ms': kavg_ms,
                     'algorithmic_bytes_per_env_step': kstep_bytes, 'whole_step_bytes_per_env_step': bytes_env,
                     'whole_step_gbs': bytes_env * N / (ms_per_step * 1e-3) / 1e9, 'peak_source': peak_src,
                     'note': 'latency/issue bound, not bandwidth bound: see DESIGN.md section 8'},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu:
            line['cpu_baseline'] = cpu_baseline(sample_envs=8192, steps=600)   # ~10-20 s of host work
        print(json.dumps(line))
    for e_ in envs:
        e_.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(sample_envs, steps, threads=None):
    """The oracle port (reference semantics restated in C, oracle/) timed on
    this box's host cores on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pyoracle as po
    threads = threads or os.cpu_count() or 1
    rec = workload_config(auto_reset=True)
    A = int(rec['n_agents'])
    b = po.OracleBatch(rec, 1, sample_envs, threads)
    b.reset()
    rng = np.random.default_rng(0)
    acts = np.zeros((4, sample_envs, A, 6), dtype=np.uint8)
    acts[..., 0:3] = rng.integers(0, 3, size=(4, sample_envs, A, 3))
    acts[..., 3:6] = rng.integers(0, 2, size=(4, sample_envs, A, 3))
    for t in range(10):
        b.step(acts[t % 4])
    t0 = time.perf_counter()
    for t in range(steps):
        b.step(acts[t % 4])
    dt = time.perf_counter() - t0
    b.close()
    return {'value': sample_envs * steps * A / dt, 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'sample': f'{sample_envs} envs x {steps} steps of the 2v2 workload, {threads} host threads, C oracle (oracle/masurv_oracle.c)',
            'seconds': dt}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pyoracle as po
    threads = os.cpu_count() or 1
    n_gpus = int(os.environ.get('WORLD_SIZE', args.gpus))
    # bounded sample of the workload: 4096 envs per "step" keeps K steps + W warm-up within minutes
    sample = min(args.envs, 4096)
    rec = workload_config(auto_reset=True)
    A = int(rec['n_agents'])
    b = po.OracleBatch(rec, args.seed, sample, threads)
    b.reset()
    rng = np.random.default_rng(0)
    acts = np.zeros((4, sample, A, 6), dtype=np.uint8)
    acts[..., 0:3] = rng.integers(0, 3, size=(4, sample, A, 3))
    acts[..., 3:6] = rng.integers(0, 2, size=(4, sample, A, 3))
    for t in range(args.warmup):
        b.step(acts[t % 4])
    t0 = time.perf_counter()
    for t in range(args.steps):
        b.step(acts[t % 4])
    dt = time.perf_counter() - t0
    b.close()
    value = sample * args.steps * A / dt
    cb = {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
          'sample': f'{sample} envs per step (bounded sample of the {args.envs}-env workload), {threads} host threads, C oracle port; '
                    'the reference\'s own pybox2d stack is not installable in this image'}
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_block(n_gpus, args.envs, 'n/a (CPU)'),
        'cpu_baseline': cb,
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=100)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--envs', type=int, default=ENVS_PER_GPU, help='environments per GPU')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--rotate', type=int, default=4, help='independent env batches stepped round-robin (working set > L2)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == '__main__':
    main()
