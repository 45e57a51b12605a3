"""Shared helpers for the parity tests: configs, action streams, and
field-by-field comparison of msv_env_state records / observation dicts."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gym-ma-survival-2d_b200'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

from masurvival.config import merge_config, pack_config, variant  # noqa: E402

REL_TOL = 1e-4  # BASELINE.json north_star: continuous outputs within 1e-4 relative


# configs far from the defaults: every config-derived device constant is exercised
EXOTIC_A = {
    'agents': {'n_agents': 3, 'agent_size': 0.8}, 'spawn_grid': {'grid_size': 4, 'floor_size': 16},
    'melee': {'range': 1.5, 'damage': 34, 'cooldown': 5}, 'inventory': {'slots': 2},
    'auto_pickup': {'shape': 0.6}, 'give': {'shape': 3.0}, 'death_drop': {'radius': 0.7},
    'heals': {'reset_spawns': {'n_items': 5, 'item_size': 0.4}, 'heal': {'healing': 30}},
    'boxes': {'reset_spawns': {'n_boxes': 3, 'box_size': 1.4}, 'item': {'item_size': 0.6, 'offset': 1.1}, 'health': 40},
    'safe_zone': {'phases': 4, 'cooldown': 30, 'damage': 2, 'radiuses': [8, 4, 2], 'centers': [[0, 0], [1, 1], [2, -1]]},
    'reward_scheme': {'r_alive': 0.5, 'r_dead': -0.25, 'r_kill': 3, 'r_death': -1}, 'gameover': {'mode': 'lastalive'},
    'health': {'health': 80}, 'motors': {'impulse': (0.3, 0.2, 0.02)}, 'cameras': {'fov': 1.0, 'depth': 6},
}
EXOTIC_B = {
    'heals': {'reset_spawns': {'n_items': 0, 'item_size': 0.5}}, 'boxes': {'reset_spawns': {'n_boxes': 2, 'box_size': 2.0}, 'ownership': True},
    'observation': {'omniscent': False}, 'safe_zone': {'cooldown': 25, 'damage': 3}, 'health': {'health': 50},
}


def apply_overrides(user, over):
    """user-config dict + overrides; a value of None deletes the key (e.g.
    {'melee': {'cooldown': None}} selects ContinuousMelee, env:309-312)."""
    for k, v in over.items():
        sub = user.setdefault(k, {})
        if isinstance(v.get('reset_spawns'), dict):      # shallow |= of the reference (env:53-54): whole sub-dict
            pass
        for kk, vv in v.items():
            if vv is None:
                sub.pop(kk, None)
            else:
                sub[kk] = vv
    return user


def make_config(name, auto_reset=False, **over):
    user = apply_overrides(variant(name), over)
    cfg, cm = merge_config(user)
    return pack_config(cfg, cm, auto_reset=auto_reset)


def random_actions(rng, n_envs, n_agents, p_attack=0.5, p_use=0.5, p_give=0.5):
    a = np.zeros((n_envs, n_agents, 6), dtype=np.uint8)
    a[..., 0:3] = rng.integers(0, 3, size=(n_envs, n_agents, 3))
    a[..., 3] = rng.random((n_envs, n_agents)) < p_attack
    a[..., 4] = rng.random((n_envs, n_agents)) < p_use
    a[..., 5] = rng.random((n_envs, n_agents)) < p_give
    return a


def _flat_fields(dt, prefix=''):
    for name in dt.names:
        sub = dt.fields[name][0]
        base = sub.base if sub.subdtype else sub
        if base.names:
            for f in _flat_fields(base, prefix + name + '.'):
                yield f
        else:
            yield prefix + name


def _get(rec, path):
    for p in path.split('.'):
        rec = rec[p]
    return np.asarray(rec)


def compare_states(a, b, rel_tol=REL_TOL):
    """Compare two msv_env_state records.  Returns (exact_mismatches,
    tolerance_failures): lists of (field, max_abs_diff)."""
    exact, fail = [], []
    for path in _flat_fields(a.dtype):
        x, y = _get(a, path), _get(b, path)
        if np.array_equal(x, y):
            continue
        if x.dtype.kind == 'f':
            d = float(np.max(np.abs(x.astype(np.float64) - y.astype(np.float64))))
            scale = np.maximum(np.abs(x), np.abs(y)).astype(np.float64)
            ok = bool(np.all(np.abs(x.astype(np.float64) - y) <= rel_tol * np.maximum(scale, 1.0)))
            exact.append((path, d))
            if not ok:
                fail.append((path, d))
        else:
            exact.append((path, float(np.max(np.abs(x.astype(np.int64) - y.astype(np.int64))))))
            fail.append(exact[-1])
    return exact, fail


DISCRETE_KEYS = ('others_mask', 'heals_mask', 'boxes_mask', 'box_items_mask', 'heal_slot',
                 'heal_slot_mask', 'box_slot_mask', 'lidar_hit', 'rewards', 'done',
                 'episode_return', 'episode_length', 'immune', 'br_over', 'br_results')


def compare_obs(a, b, rel_tol=REL_TOL):
    """a, b: dict key -> ndarray for ONE env.  Discrete keys (and the id /
    team / health columns of agent rows) must be identical; the rest within
    rel_tol.  Returns (exact_mismatches, failures)."""
    exact, fail = [], []
    for k in a:
        if k not in b or k == 'n_toi_events':
            continue
        x, y = np.asarray(a[k]), np.asarray(b[k])
        if x.shape != y.shape:
            fail.append((k, 'shape %s vs %s' % (x.shape, y.shape)))
            continue
        if np.array_equal(x, y):
            continue
        d = float(np.max(np.abs(x.astype(np.float64) - y.astype(np.float64))))
        exact.append((k, d))
        if k in DISCRETE_KEYS:
            fail.append((k, d))
            continue
        scale = np.maximum(np.maximum(np.abs(x), np.abs(y)), 1.0)
        if not np.all(np.abs(x.astype(np.float64) - y) <= rel_tol * scale):
            fail.append((k, d))
        if k in ('agent', 'others'):
            ncol = x.shape[-1] - 6  # id, (team), health are discrete
            if not np.array_equal(x[..., :ncol], y[..., :ncol]):
                fail.append((k + '[discrete cols]', d))
    return exact, fail


def scramble_state(s, rec, rng):
    """Turn a freshly reset msv_env_state into a dense-interaction scenario:
    agents are teleported next to random entities / each other, get random
    health, inventories and melee cooldowns.  Fat AABBs are rebuilt tight+0.1
    and all contact pairs cleared, so both implementations start by running
    FindNewContacts on the injected world."""
    s = s.copy()
    A, B, H = int(rec['n_agents']), int(s['n_boxes']), int(s['n_heals'])
    r = np.float32(rec['agent_size'] / 2)
    ext = np.float32(0.1)
    for i in range(A):
        kind = rng.integers(0, 5)
        if kind == 0 and B > 0:
            k = rng.integers(0, B); cx, cy = s['box_x'][k], s['box_y'][k]; d = rng.uniform(0.9, 1.6)
        elif kind == 1 and H > 0:
            k = rng.integers(0, H); cx, cy = s['heal_x'][k], s['heal_y'][k]; d = rng.uniform(0.0, 1.0)
        elif kind == 2 and i > 0:
            k = rng.integers(0, i); cx, cy = s['x'][k], s['y'][k]; d = rng.uniform(0.95, 1.8)
        elif kind == 3:
            w = rng.integers(0, 4); t = rng.uniform(-9, 9); off = rng.uniform(9.2, 9.45)
            cx, cy = [(-off, t), (t, off), (off, t), (t, -off)][w]; d = 0.0
        else:
            cx, cy, d = s['x'][i], s['y'][i], 0.0
        ang = rng.uniform(0, 2 * np.pi)
        x = np.float32(np.clip(cx + d * np.cos(ang), -9.45, 9.45))
        y = np.float32(np.clip(cy + d * np.sin(ang), -9.45, 9.45))
        s['x'][i], s['y'][i] = x, y
        s['angle'][i] = np.float32(rng.uniform(-np.pi, np.pi))
        s['vx'][i], s['vy'][i] = np.float32(rng.normal(0, 4)), np.float32(rng.normal(0, 4))
        s['omega'][i] = np.float32(rng.normal(0, 2))
        s['fat'][i] = [(x - r) - ext, (y - r) - ext, (x + r) + ext, (y + r) + ext]
        s['health'][i] = int(rng.choice([1, 2, 19, 20, 21, 40, 100]))
        s['cooldown'][i] = int(rng.choice([0, 0, 1, 2, 39]))
        n = int(rng.integers(0, int(rec['inv_slots']) + 1))
        s['inv_n'][i] = n
        for k in range(n):
            if B > 0 and rng.random() < 0.5:
                s['inv_kind'][i][k] = 2
                s['inv_shape'][i][k]['hx'] = np.float32(rng.uniform(0.2, 0.9))
                s['inv_shape'][i][k]['hy'] = np.float32(rng.uniform(0.2, 0.9))
                s['inv_shape'][i][k]['rehulled'] = 1
                s['inv_owner'][i][k] = int(rng.integers(0, A)) if not rec['teams'] else 100 + int(rng.integers(0, 2))
            else:
                s['inv_kind'][i][k] = 1
                s['inv_owner'][i][k] = -1
    # conservation: live + floor + carried boxes (heals) never exceed what a
    # reset created, so free some floor entities for what the agents carry
    bcap = 4 if A <= 4 and B <= 4 else 8
    hcap = 4 if A <= 4 and H <= 4 else 16
    drop_b = int(rng.integers(0, min(B, 3) + 1)) if B > 0 else 0
    drop_h = int(rng.integers(0, min(H, 3) + 1)) if H > 0 else 0
    s['n_boxes'] = B - drop_b
    s['n_heals'] = H - drop_h
    room_b, room_h = bcap - int(s['n_boxes']), hcap - int(s['n_heals'])
    for i in range(A):
        keep = 0
        for k in range(int(s['inv_n'][i])):
            kind = int(s['inv_kind'][i][k])
            if kind == 2 and room_b > 0:
                room_b -= 1
            elif kind == 1 and room_h > 0:
                room_h -= 1
            else:
                continue
            for f in ('inv_kind', 'inv_owner'):
                s[f][i][keep] = s[f][i][k]
            s['inv_shape'][i][keep] = s['inv_shape'][i][k]
            keep += 1
        s['inv_n'][i] = keep
    for i in range(A):  # canonical form: nothing beyond the list lengths
        for k in range(s['inv_kind'].shape[1]):
            if k >= int(s['inv_n'][i]):
                s['inv_kind'][i][k] = 0; s['inv_owner'][i][k] = 0
            if k >= int(s['inv_n'][i]) or int(s['inv_kind'][i][k]) == 1:
                s['inv_shape'][i][k] = np.zeros((), dtype=s['inv_shape'].dtype)
    for k in range(int(s['n_boxes']), s['box_x'].shape[0]):
        for f in ('box_x', 'box_y', 'box_health', 'box_has_health', 'box_cause', 'box_owner', 'box_seq'):
            s[f][k] = 0
        s['box_shape'][k] = np.zeros((), dtype=s['box_shape'].dtype)
    for k in range(int(s['n_heals']), s['heal_x'].shape[0]):
        s['heal_x'][k] = 0; s['heal_y'][k] = 0; s['heal_seq'][k] = 0
    for name in ('pair_aa', 'pair_ab', 'pair_aw'):
        s[name] = np.zeros_like(s[name])
    return s
