"""Host-side checks that need no GPU: the C-ABI library loads and exports
every symbol include/masurv.h declares, struct layouts agree, the product
fails loudly without a device, config handling mirrors the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

import parity
from masurvival import _lib
from masurvival.config import CONFIG_DT, STATE_DT, default_config, merge_config, pack_config, variant

ROOT = parity.ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'masurv.h')).read()
    hdr = re.sub(r'/\*.*?\*/', ' ', hdr, flags=re.S)
    declared = set(re.findall(r'\b(msv_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 20
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f'{name} declared in include/masurv.h but not exported'
    assert declared == set(_lib.EXPORTS)


def test_struct_layouts_match_compiler():
    L = _lib.load()
    assert L.msv_sizeof_config() == CONFIG_DT.itemsize
    assert L.msv_sizeof_env_state() == STATE_DT.itemsize


def test_default_config_matches_reference_defaults():
    L = _lib.load()
    c = np.zeros(1, dtype=CONFIG_DT)
    assert L.msv_default_config(c.ctypes.data) == 0
    cfg, cm = merge_config(None)
    rec = pack_config(cfg, cm)
    for name in CONFIG_DT.names:
        assert np.array_equal(c[0][name], rec[name]), name
    # env:166-172: the class default is 1v1 without teams
    assert rec['n_agents'] == 2 and rec['teams'] == 0 and rec['melee_cooldown'] == 40
    assert rec['zone_n_radiuses'] == 4 and rec['zone_phases'] == 5 and rec['health'] == 100


def test_no_cpu_fallback():
    """without a CUDA device msv_create must fail, not silently run on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.MasurvError, match='NO_DEVICE'):
        _lib.Handle(parity.make_config('2v2'), 4)


def test_philox_known_answer():
    """Random123 KAT for Philox4x32-10 (counter = key = 0 and the pi vector)"""
    L = _lib.load()
    import pyoracle as po
    def run(fn, ctr, key):
        c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key); o = (ctypes.c_uint32 * 4)()
        fn(c, k, o)
        return [int(x) for x in o]
    kat = [((0, 0, 0, 0), (0, 0), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert run(L.msv_philox4x32, ctr, key) == want
        assert run(po.lib().orc_philox4x32, ctr, key) == want


def test_config_merge_rules():
    # env:53-54: shallow |= of sub-dicts
    cfg, cm = merge_config({'melee': {'cooldown': 40}, 'boxes': {'reset_spawns': {'n_boxes': 0, 'box_size': 1}}})
    assert cfg['boxes']['reset_spawns'] == {'n_boxes': 0, 'box_size': 1} and cfg['boxes']['health'] == 20 and not cm
    # env:310: a user config must carry 'melee'; without 'cooldown' -> ContinuousMelee
    with pytest.raises(KeyError):
        merge_config({'agents': {'n_agents': 4}})
    cfg, cm = merge_config({'melee': {}})
    assert cm and pack_config(cfg, cm)['melee_cooldown'] == -1
    with pytest.raises(IndexError):   # more spawns than grid cells (sem:76-79 pops an empty list)
        pack_config(*merge_config({'melee': {'cooldown': 1}, 'agents': {'n_agents': 8}, 'heals': {'reset_spawns': {'n_items': 8, 'item_size': 0.5}}}))
    with pytest.raises(AssertionError):
        pack_config(*merge_config({'melee': {'cooldown': 1}, 'gameover': {'mode': 'nope'}}))
    for v in ('1v1', '1v1_heal_only', '2v2', 'ffa', 'ffa_lidar'):
        parity.make_config(v)


def test_wrapper_shapes_without_gpu():
    """observation/action spaces mirror env:391-453"""
    from masurvival.envs.masurvival_env import MaSurvivalVec
    from masurvival.envs import spaces
    class Fake(MaSurvivalVec):
        def __init__(self, user):
            self.config, cm = merge_config(user)
            self._rec = pack_config(self.config, cm)
    f = Fake(variant('2v2'))
    shp = f.obs_shapes()
    assert shp['agent'] == (4, 9) and shp['others'] == (4, 3, 9) and shp['boxes'] == (4, 4, 11)
    assert shp['box_items'] == (4, 4, 10) and shp['heal_slot'] == (4, 1, 1) and shp['box_slot'] == (4, 1, 8)
    assert sum(int(np.prod(s)) for s in shp.values()) == 640
    f1 = Fake(None)
    assert sum(int(np.prod(s)) for s in f1.obs_shapes().values()) == 276
    assert f.entity_keys() == {'others', 'heals', 'heal_slot', 'boxes', 'box_items', 'box_slot'}
    sp = f.compute_action_space()
    assert len(sp) == 4 and sp.contains(tuple([1, 1, 1, 0, 0, 0] for _ in range(4))) and not sp.contains(tuple([3, 0, 0, 0, 0, 0] for _ in range(4)))


def test_create_rejects_bad_arguments_before_touching_the_device():
    """error behaviour of the C ABI: negative codes, no exception, no crash"""
    L = _lib.load()
    good = np.array(parity.make_config('2v2'), dtype=CONFIG_DT).reshape(1)
    h = ctypes.c_void_p()
    def create(rec, n=4):
        return L.msv_create(rec.ctypes.data, n, 0, 0, 0, ctypes.byref(h))
    assert create(good, 0) == -1                                   # MSV_ERR_INVALID: num_envs <= 0
    for field, val in (('n_agents', 0), ('n_agents', 9), ('n_boxes', 9), ('n_heals', 17), ('grid_size', 9),
                       ('inv_slots', 5), ('lidar_n', 33), ('zone_n_radiuses', 8)):
        bad = good.copy(); bad[0][field] = val
        assert create(bad) == -1, field
    bad = good.copy(); bad[0]['grid_size'] = 2                        # 4+4+4 spawns > 4 cells
    assert create(bad) == -1
    assert L.msv_create(None, 4, 0, 0, 0, ctypes.byref(h)) == -1
    for fn in (L.msv_reset, L.msv_observe):
        assert fn(None, None) == -1
    assert L.msv_step(None, None, None) == -1 and L.msv_destroy(None) == -1
    assert L.msv_bytes_per_env_step(None) == 0 and L.msv_kernel_launches(None) == 0


def test_tile_planner_host_logic():
    """msv_plan_tile: the k_step tile planner without a device.  One block of <= 512 threads per SM; a batch that
    fits one wave is spread over every SM, one that needs several waves takes the largest tile."""
    L = _lib.load()
    SMS, SMEM = 148, 232448                       # B200: multiProcessorCount, sharedMemPerBlockOptin

    def plan(variant, n, sms=SMS, smem=SMEM):
        rec = parity.make_config(variant, auto_reset=True)
        out = (ctypes.c_int32 * 4)()
        assert L.msv_plan_tile(rec.ctypes.data, n, sms, smem, out) == 0
        return tuple(int(x) for x in out)

    # the BASELINE.json batch sizes (what bench.py reports as tile_plan on a B200)
    assert plan('2v2', 16384) == (112, 147, 448, 1)
    assert plan('ffa', 8192) == (56, 147, 448, 2)
    assert plan('1v1_heal_only', 4096) == (32, 128, 64, 0)
    assert plan('ffa_lidar', 32768) == (64, 512, 512, 2)      # several waves: the largest tile
    assert plan('2v2', 1) == (8, 1, 32, 1) and plan('ffa', 3) == (4, 1, 32, 2)
    lanes = {0: 2, 1: 4, 2: 8}
    rng = np.random.default_rng(0)
    for variant in ('1v1', '2v2', 'ffa'):
        for n in [int(x) for x in rng.integers(1, 200000, size=40)] + [147 * 112, 148 * 128, 148 * 128 + 1]:
            epb, blocks, threads, cap = plan(variant, n)
            G = lanes[cap]
            assert threads == epb * G and 32 <= threads <= 512 and threads % 32 == 0     # whole warps, whole groups
            assert epb * blocks >= n > epb * (blocks - 1)
            if blocks > SMS:                                                              # several waves -> full tiles
                assert plan(variant, 10 ** 7)[0] == epb
    # a GPU with less shared memory per block gets smaller tiles; invalid arguments are rejected
    assert plan('2v2', 16384, smem=100 * 1024)[0] < 112
    out = (ctypes.c_int32 * 4)()
    rec = parity.make_config('2v2')
    assert L.msv_plan_tile(None, 16, SMS, SMEM, out) != 0 and L.msv_plan_tile(rec.ctypes.data, 0, SMS, SMEM, out) != 0
