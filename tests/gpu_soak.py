"""Long lock-step soak (development / evidence): many env-steps of bit-exact
comparison CUDA vs oracle, written to gpurun_out/soak_*.json."""
import json, os, sys, time
import gpu_lockstep

out = {}
t0 = time.time()
for name, n, steps, over in (('2v2', 256, 1500, {}), ('1v1', 256, 1200, {}), ('ffa', 48, 500, {}),
                             ('2v2', 128, 800, {'observation': {'omniscent': False}, 'boxes': {'ownership': True}}),
                             ('ffa', 32, 400, {'spawn_grid': {'grid_size': 8, 'floor_size': 14}, 'safe_zone': {'cooldown': 80, 'radiuses': [7, 4, 2, 1]}, 'inventory': {'slots': 1}}),
                             ('ffa_lidar', 24, 250, {})):
    r = gpu_lockstep.run(name, n, steps, verbose=False, **over)
    r.pop('details', None)
    out[f'{name}_{n}x{steps}_{"+".join(over) or "default"}'] = r
    print(name, n, steps, {k: r[k] for k in ('env_steps', 'state_exact_mismatch', 'obs_exact_mismatch', 'toi_events', 'dones', 'overflow_events')}, f'{time.time() - t0:.0f}s', flush=True)
os.makedirs(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gpurun_out', 'soak_r02.json'), 'w'), indent=1)
bad = sum(v['state_exact_mismatch'] + v['obs_exact_mismatch'] for v in out.values())
sys.exit(1 if bad else 0)
