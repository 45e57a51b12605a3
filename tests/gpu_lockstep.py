"""Lock-step parity driver: CUDA path (through the C ABI) vs the CPU oracle.

Both sides start from the same Philox-seeded reset, receive the same actions
and are compared after EVERY step on the full msv_env_state record plus the
observation dict, rewards and dones.  After a mismatch the oracle's state is
re-injected into the GPU (msv_set_state) so that one divergence does not
mask later ones.  Used by tests/test_gpu_parity.py and runnable by hand:
    python tests/gpu_lockstep.py 2v2 64 300
"""
import json
import sys

import numpy as np

import parity
from parity import compare_obs, compare_states, make_config, random_actions


def run(variant='2v2', n_envs=64, steps=300, seed=7, auto_reset=True, verbose=True, max_report=10,
        p_attack=0.5, p_use=0.5, p_give=0.5, **over):
    import torch
    import pyoracle as po
    from masurvival import _lib

    rec = make_config(variant, auto_reset=auto_reset, **over)
    A = int(rec['n_agents'])
    h = _lib.Handle(rec, n_envs, device=0, seed=seed, env_offset=0)
    orcs = [po.OracleEnv(rec, seed=seed, env_id=e) for e in range(n_envs)]
    keys = list(po.obs_dims(rec).keys())
    extra = [k for k in ('immune', 'br_over', 'br_results') if h.has_tensor(k)]

    def gpu_obs():
        torch.cuda.synchronize()
        out = {}
        for k in keys + ['rewards', 'dones', 'episode_return', 'episode_length'] + extra:
            out[k] = h.tensor(k).cpu().numpy()
        return out

    def expand(k, arr):  # per-env de-duplicated tensors -> reference shape
        if k in ('zone', 'heals', 'boxes', 'box_items'):
            return np.broadcast_to(arr[None], (A,) + arr.shape)
        return arr

    report = {'state_exact_mismatch': 0, 'state_fail': 0, 'obs_exact_mismatch': 0, 'obs_fail': 0,
              'env_steps': 0, 'toi_events': 0, 'dones': 0, 'episode_stats_checked': 0, 'details': []}

    def check(tag, oouts):
        g = gpu_obs()
        sg = h.get_state()
        for e in range(n_envs):
            so = orcs[e].get_state()
            ex, fl = compare_states(sg[e], so)
            og = {k: expand(k, g[k][e]) for k in keys}
            og['rewards'] = g['rewards'][e]
            og['done'] = bool(g['dones'][e])
            oo = dict(oouts[e])
            if tag != 'reset':
                if oo['done']:   # per-env episode statistics: rows valid where done
                    og['episode_return'] = g['episode_return'][e]; og['episode_length'] = int(g['episode_length'][e])
                    report['episode_stats_checked'] += 1
                for k in extra:
                    og[k] = g[k][e] if k == 'br_results' else int(g[k][e])
            elif 'immune' in extra:
                og['immune'] = int(g['immune'][e])
            oex, ofl = compare_obs(og, oo)
            report['state_exact_mismatch'] += bool(ex)
            report['state_fail'] += bool(fl)
            report['obs_exact_mismatch'] += bool(oex)
            report['obs_fail'] += bool(ofl)
            if (ex or oex) and len(report['details']) < max_report:
                report['details'].append({'at': tag, 'env': e, 'state': ex[:6], 'obs': oex[:6],
                                          'state_fail': fl[:6], 'obs_fail': ofl[:6]})
            if ex:
                h.set_state(np.array([so]), first=e)  # teacher forcing

    h.reset()
    oouts = [o.reset() for o in orcs]
    check('reset', oouts)
    rng = np.random.default_rng(seed + 1)
    for t in range(steps):
        act = random_actions(rng, n_envs, A, p_attack, p_use, p_give)
        a_dev = torch.as_tensor(act).cuda()
        h.step(a_dev.data_ptr())
        oouts = [orcs[e].step(act[e]) for e in range(n_envs)]
        report['env_steps'] += n_envs
        report['toi_events'] += sum(o['n_toi_events'] for o in oouts)
        report['dones'] += sum(o['done'] for o in oouts)
        check('step %d' % t, oouts)
    sg = h.flush_stats()
    so = [o.flush_stats() for o in orcs]
    report['stats_gpu'] = {'steps': int(sg['steps']), 'heals_used': int(sg['heals_used']),
                           'boxes_placed': int(sg['boxes_placed']), 'episodes': int(sg['episodes']),
                           'kills': [int(x) for x in sg['kills'][:A]]}
    report['stats_oracle'] = {'steps': int(sum(s['steps'] for s in so)), 'heals_used': int(sum(s['heals_used'] for s in so)),
                              'boxes_placed': int(sum(s['boxes_placed'] for s in so)),
                              'episodes': int(sum(s['episodes'] for s in so)),
                              'kills': [int(sum(s['kills'][i] for s in so)) for i in range(A)]}
    report['overflow_events'] = h.overflow_events()
    ev = {}
    for o in orcs:                                   # which rules / quirks this lock-step run exercised (oracle census)
        for k, v in o.events().items():
            ev[k] = ev.get(k, 0) + v
    report['events'] = {k: v for k, v in ev.items() if v}
    h.close()
    if verbose:
        print(json.dumps(report, default=str, indent=1))
    return report


if __name__ == '__main__':
    v = sys.argv[1] if len(sys.argv) > 1 else '2v2'
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    s = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    r = run(v, n, s)
    sys.exit(1 if (r['state_fail'] or r['obs_fail']) else 0)
