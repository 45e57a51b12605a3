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


def make_config(name, auto_reset=False, **over):
    user = variant(name)
    for k, v in over.items():
        user.setdefault(k, {}).update(v)
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
                 'heal_slot_mask', 'box_slot_mask', 'lidar_hit', 'rewards', 'done')


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
