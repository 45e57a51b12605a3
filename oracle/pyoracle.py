"""ctypes binding of the TEST ORACLE (oracle/liboracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product
package."""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
sys.path.insert(0, os.path.join(_ROOT, 'gym-ma-survival-2d_b200'))
from masurvival._cstruct import parse_header  # noqa: E402

DEFINES, STRUCTS = parse_header(os.path.join(_ROOT, 'include', 'masurv.h'),
                                os.path.join(_HERE, 'masurv_oracle.h'))
CONFIG_DT = STRUCTS['msv_config']
STATE_DT = STRUCTS['msv_env_state']
OUT_DT = STRUCTS['orc_out']
STATS_DT = STRUCTS['msv_stats']


def build(force=False):
    name = os.environ.get('ORACLE_LIB', 'liboracle.so')      # ORACLE_LIB=liboracle_asan.so: the sanitizer build (make asan)
    so = os.path.join(_HERE, name)
    if force or not os.path.exists(so):
        subprocess.check_call(['make', '-C', _HERE, name])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        vp, u64, i64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int32
        L.orc_create.restype = vp
        L.orc_create.argtypes = [vp, u64, i64]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_draws.argtypes = [vp, vp]
        L.orc_reset.argtypes = [vp, vp]
        L.orc_step.argtypes = [vp, vp, vp]
        L.orc_observe.argtypes = [vp, vp]
        L.orc_get_state.argtypes = [vp, vp]
        L.orc_set_state.argtypes = [vp, vp]
        L.orc_flush_stats.argtypes = [vp, vp]
        L.orc_get_events.argtypes = [vp, vp, i32]
        L.orc_rollout.restype = i64
        L.orc_rollout.argtypes = [vp, u64, i32, i32, i32]
        L.orc_batch_create.restype = vp
        L.orc_batch_create.argtypes = [vp, u64, i32, i32]
        L.orc_batch_destroy.argtypes = [vp]
        L.orc_batch_reset.argtypes = [vp]
        L.orc_batch_step.argtypes = [vp, vp, vp, vp]
        L.orc_philox_uniform.restype = ctypes.c_double
        L.orc_philox_uniform.argtypes = [u64] + [ctypes.c_uint32] * 5
        L.orc_philox4x32.argtypes = [vp, vp, vp]
        L.orc_philox_actions.argtypes = [u64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, i32, vp]
        _lib = L
    return _lib


class _Draws(ctypes.Structure):
    _fields_ = [('shuffle_u', ctypes.c_void_p), ('box_z', ctypes.c_void_p),
                ('zone_u', ctypes.c_void_p), ('death_u', ctypes.c_void_p)]


def _event_names():
    import re
    txt = open(os.path.join(_HERE, 'masurv_oracle.h')).read()
    body = re.search(r'enum \{\s*(ORC_EV_DEATH.*?)ORC_EV_COUNT', txt, flags=re.S).group(1)
    return [n.strip()[len('ORC_EV_'):].lower() for n in body.split(',') if n.strip()]


EVENT_NAMES = _event_names()


def obs_dims(cfg):
    A, B, H = int(cfg['n_agents']), int(cfg['n_boxes']), int(cfg['n_heals'])
    S = 8 + (1 if cfg['teams'] else 0)
    L = int(cfg['lidar_n'])
    d = {'agent': (A, S), 'others': (A, A - 1, S), 'others_mask': (A, A - 1), 'zone': (A, 6)}
    if H > 0:
        d.update({'heals': (A, H, 2), 'heals_mask': (A, H), 'heal_slot': (A, 1, 1), 'heal_slot_mask': (A, 1)})
    if B > 0:
        d.update({'boxes': (A, B, 11), 'boxes_mask': (A, B), 'box_items': (A, B, 10),
                  'box_items_mask': (A, B), 'box_slot': (A, 1, 8), 'box_slot_mask': (A, 1)})
    if L > 0:
        d.update({'lidar_frac': (A, L), 'lidar_hit': (A, L)})
    return d


def unpack_out(cfg, out):
    """orc_out record -> dict of arrays with the reference's shapes."""
    res = {}
    for k, shp in obs_dims(cfg).items():
        n = int(np.prod(shp))
        res[k] = np.array(out[k].reshape(-1)[:n]).reshape(shp)
    A = int(cfg['n_agents'])
    res['rewards'] = np.array(out['rewards'][:A])
    res['done'] = bool(out['done'])
    res['n_toi_events'] = int(out['n_toi_events'])
    res['episode_return'] = np.array(out['episode_return'][:A])
    res['episode_length'] = int(out['episode_length'])
    res['immune'] = int(out['immune'])
    res['br_over'] = int(out['br_over'])
    res['br_results'] = np.array(out['br_results'][:A])
    return res


class OracleEnv:
    """One reference-equivalent environment on the CPU oracle."""

    def __init__(self, cfg, seed=0, env_id=0):
        self.cfg = np.array(cfg, dtype=CONFIG_DT).reshape(())
        self._cfgbuf = np.ascontiguousarray(self.cfg.reshape(1))
        self.h = lib().orc_create(self._cfgbuf.ctypes.data, seed, env_id)
        self._out = np.zeros(1, dtype=OUT_DT)
        self._keep = None

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_draws(self, shuffle_u=None, box_z=None, zone_u=None, death_u=None):
        if shuffle_u is None and box_z is None and zone_u is None and death_u is None:
            lib().orc_set_draws(self.h, None)
            self._keep = None
            return
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64)
                for a in (shuffle_u, box_z, zone_u, death_u)]
        d = _Draws(*[None if a is None else a.ctypes.data for a in arrs])
        self._keep = (arrs, d)
        lib().orc_set_draws(self.h, ctypes.addressof(d))

    def reset(self):
        lib().orc_reset(self.h, self._out.ctypes.data)
        return unpack_out(self.cfg, self._out[0])

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        assert a.size == int(self.cfg['n_agents']) * 6
        lib().orc_step(self.h, a.ctypes.data, self._out.ctypes.data)
        return unpack_out(self.cfg, self._out[0])

    def observe(self):
        lib().orc_observe(self.h, self._out.ctypes.data)
        return unpack_out(self.cfg, self._out[0])

    def get_state(self):
        s = np.zeros(1, dtype=STATE_DT)
        lib().orc_get_state(self.h, s.ctypes.data)
        return s[0]

    def set_state(self, s):
        buf = np.ascontiguousarray(np.array(s, dtype=STATE_DT).reshape(1))
        lib().orc_set_state(self.h, buf.ctypes.data)

    def flush_stats(self):
        s = np.zeros(1, dtype=STATS_DT)
        lib().orc_flush_stats(self.h, s.ctypes.data)
        return s[0]

    def events(self):
        """event-coverage census: dict name -> count since creation (ORC_EV_* of masurv_oracle.h)"""
        buf = np.zeros(len(EVENT_NAMES), dtype=np.int64)
        lib().orc_get_events(self.h, buf.ctypes.data, len(buf))
        return dict(zip(EVENT_NAMES, (int(v) for v in buf)))


def rollout(cfg, seed, n_envs, steps, n_threads):
    buf = np.ascontiguousarray(np.array(cfg, dtype=CONFIG_DT).reshape(1))
    return lib().orc_rollout(buf.ctypes.data, seed, n_envs, steps, n_threads)


def philox_uniform(seed, env, episode, step, stream, k):
    return lib().orc_philox_uniform(seed, env, episode, step, stream, k)


class OracleBatch:
    """n reference-equivalent envs stepped together on `n_threads` host
    threads (the CPU arm of bench.py)."""

    def __init__(self, cfg, seed, n_envs, n_threads):
        self._cfgbuf = np.ascontiguousarray(np.array(cfg, dtype=CONFIG_DT).reshape(1))
        self.n, self.A = int(n_envs), int(self._cfgbuf[0]['n_agents'])
        self.h = lib().orc_batch_create(self._cfgbuf.ctypes.data, seed, n_envs, n_threads)
        self.rewards = np.zeros((self.n, self.A), dtype=np.float32)
        self.dones = np.zeros(self.n, dtype=np.uint8)

    def reset(self):
        lib().orc_batch_reset(self.h)

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        lib().orc_batch_step(self.h, a.ctypes.data, self.rewards.ctypes.data, self.dones.ctypes.data)
        return self.rewards, self.dones

    def close(self):
        if self.h:
            lib().orc_batch_destroy(self.h)
            self.h = None
