"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on identical seeded inputs.  Discrete outputs (health, inventories,
alive/done flags, masks, rewards, lidar hit ids, contact bookkeeping) must be
bit-exact; continuous outputs within 1e-4 relative (BASELINE.json north_star).
In practice every field is bit-exact and the tests also assert that."""
import numpy as np
import pytest

import parity
from parity import compare_obs, compare_states, make_config, random_actions, scramble_state

pytestmark = pytest.mark.gpu


def _assert_clean(r, need_exact=True):
    assert r['state_fail'] == 0 and r['obs_fail'] == 0, r['details']
    if need_exact:
        assert r['state_exact_mismatch'] == 0 and r['obs_exact_mismatch'] == 0, r['details']
    assert r['stats_gpu'] == r['stats_oracle']


@pytest.mark.parametrize('variant,n,steps', [('2v2', 64, 150), ('1v1', 64, 150), ('1v1_heal_only', 64, 100),
                                             ('ffa', 16, 80), ('ffa_lidar', 8, 40)])
def test_lockstep_rollout(variant, n, steps):
    import gpu_lockstep
    r = gpu_lockstep.run(variant, n, steps, verbose=False)
    _assert_clean(r)


def test_lockstep_fast_zone_and_resets():
    """short safe-zone phases: every env runs through shrink, endgame, death,
    death-drop and in-kernel auto-reset several times"""
    import gpu_lockstep
    r = gpu_lockstep.run('2v2', 48, 260, verbose=False, safe_zone={'cooldown': 8}, health={'health': 12})
    _assert_clean(r)
    assert r['dones'] > 40


def test_lockstep_partial_observability():
    """non-omniscient masks (env:706-739): Cameras over every body, Q1 row remap"""
    import gpu_lockstep
    r = gpu_lockstep.run('2v2', 48, 200, verbose=False, observation={'omniscent': False}, safe_zone={'cooldown': 20}, health={'health': 30})
    _assert_clean(r)
    r = gpu_lockstep.run('ffa', 12, 60, verbose=False, observation={'omniscent': False})
    _assert_clean(r)


def test_lockstep_exotic_configs():
    """configs far from the defaults (agent size, room size, fixed zone centres, 3 agents,
    2 inventory slots, lastalive, kill rewards, no heals, ownership without teams)"""
    import gpu_lockstep
    r = gpu_lockstep.run('1v1', 48, 200, verbose=False, **{k: dict(v) for k, v in parity.EXOTIC_A.items()})
    _assert_clean(r)
    assert r['dones'] > 10
    r = gpu_lockstep.run('1v1', 48, 150, verbose=False, **{k: dict(v) for k, v in parity.EXOTIC_B.items()})
    _assert_clean(r)


def test_lockstep_ownership():
    import gpu_lockstep
    r = gpu_lockstep.run('2v2', 32, 120, verbose=False, boxes={'ownership': True}, p_attack=0.9)
    _assert_clean(r)


def _scenario(variant, n, steps, seed, **over):
    """dense-interaction scenario: scrambled states injected on both sides"""
    import torch
    import pyoracle as po
    from masurvival import _lib
    rec = make_config(variant, auto_reset=False, **over)
    A = int(rec['n_agents'])
    h = _lib.Handle(rec, n, 0, seed, 0)
    orcs = [po.OracleEnv(rec, seed=seed, env_id=e) for e in range(n)]
    h.reset()
    rng = np.random.default_rng(seed)
    states = []
    for e in range(n):
        orcs[e].reset()
        s = scramble_state(orcs[e].get_state(), rec, rng)
        orcs[e].set_state(s)
        states.append(s)
    h.set_state(np.array(states))
    keys = list(po.obs_dims(rec).keys())
    bad = []
    events = {'toi': 0, 'deaths': 0, 'pick': 0}
    for t in range(steps):
        act = random_actions(rng, n, A, 0.6, 0.5, 0.3)
        a_dev = torch.as_tensor(act).cuda()
        h.step(a_dev.data_ptr())
        torch.cuda.synchronize()
        sg = h.get_state()
        g = {k: h.tensor(k).cpu().numpy() for k in keys + ['rewards', 'dones']}
        for e in range(n):
            oo = orcs[e].step(act[e])
            events['toi'] += oo['n_toi_events']
            so = orcs[e].get_state()
            ex, fl = compare_states(sg[e], so)
            og = {k: (np.broadcast_to(g[k][e][None], (A,) + g[k][e].shape) if k in ('zone', 'heals', 'boxes', 'box_items') else g[k][e]) for k in keys}
            og['rewards'] = g['rewards'][e]; og['done'] = bool(g['dones'][e])
            oex, ofl = compare_obs(og, oo)
            if ex or oex:
                bad.append((t, e, ex[:4], oex[:4]))
                h.set_state(np.array([so]), first=e)
            events['deaths'] += int((so['alive'][:A] == 0).sum())
            events['pick'] += int(so['inv_n'][:A].sum())
    assert h.overflow_events() == 0
    h.close()
    return bad, events


@pytest.mark.parametrize('variant,n,steps,over', [
    ('2v2', 96, 40, {}), ('1v1', 96, 40, {}), ('ffa', 24, 30, {}),
    ('2v2', 64, 40, {'boxes': {'ownership': True}}),
    ('2v2', 64, 40, {'observation': {'omniscent': False}}),
])
def test_scrambled_scenarios(variant, n, steps, over):
    bad, events = _scenario(variant, n, steps, seed=11, **over)
    assert not bad, bad[:5]
    assert events['toi'] > 0


def test_step_host_matches_step_and_dlpack_views():
    """msv_step_host (host buffers) and msv_step (device buffer) are the same
    computation; DLPack views alias library memory with reference shapes."""
    import torch
    from masurvival.envs import MaSurvivalVec
    from masurvival.config import variant
    N = 256
    e1 = MaSurvivalVec(variant('2v2'), N, seed=3)
    e2 = MaSurvivalVec(variant('2v2'), N, seed=3)
    o1 = e1.reset(); e2.reset()
    assert o1['agent'].shape == (N, 4, 9) and o1['zone'].shape == (N, 4, 6)
    assert o1['boxes'].shape == (N, 4, 4, 11) and o1['boxes'].stride(1) == 0
    assert o1['others'].shape == (N, 4, 3, 9) and o1['heal_slot'].shape == (N, 4, 1, 1)
    rng = np.random.default_rng(0)
    for t in range(30):
        a = random_actions(rng, N, 4)
        obs, rew, done, info = e1.step(torch.as_tensor(a).cuda())
        rh, dh = e2.step_host(a)
        torch.cuda.synchronize()
        assert np.array_equal(rew.cpu().numpy(), rh) and np.array_equal(done.cpu().numpy().astype(np.uint8), dh)
    s1, s2 = e1.get_state(), e2.get_state()
    assert s1.tobytes() == s2.tobytes()
    p = obs['agent'].data_ptr()
    obs2, *_ = e1.step(torch.as_tensor(a).cuda())
    assert obs2['agent'].data_ptr() == p  # zero-copy: same HBM every step
    e1.close(); e2.close()


def test_full_size_properties():
    """BASELINE size (16384 envs): determinism and domain invariants."""
    import torch
    from masurvival.envs import MaSurvivalVec
    from masurvival.config import variant
    N = 16384
    envs = [MaSurvivalVec(variant('2v2'), N, seed=5) for _ in range(2)]
    for e in envs:
        e.reset()
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    tot_done = 0
    for t in range(120):
        a = torch.empty((N, 4, 6), dtype=torch.uint8, device='cuda')
        a[..., :3] = torch.randint(0, 3, (N, 4, 3), dtype=torch.uint8, device='cuda', generator=g)
        a[..., 3:] = torch.randint(0, 2, (N, 4, 3), dtype=torch.uint8, device='cuda', generator=g)
        outs = [e.step(a) for e in envs]
        tot_done += int(outs[0][2].sum())
    o0, r0, d0, _ = outs[0]; o1, r1, d1, _ = outs[1]
    for k in o0:
        assert torch.equal(o0[k], o1[k]), k          # same seed + actions -> identical (no atomics, no races)
    assert torch.equal(r0, r1) and torch.equal(d0, d1)
    ag = o0['agent']
    assert torch.all(ag[..., 3:5].abs() <= 10.0)      # nobody tunnels through the walls (TOI)
    hp = ag[..., 2]
    assert torch.all(hp == hp.round()) and torch.all(hp >= 0)
    assert torch.all((o0['others_mask'] == 0) | (o0['others_mask'] == 1))
    assert torch.all((r0 == 1) | (r0 == -1))
    st = envs[0].flush_stats()
    assert st['steps'] == N * 120 and st['episodes'] == tot_done
    assert envs[0]._h.overflow_events() == 0      # no fixed-capacity list ever overflowed
    for e in envs:
        e.close()


def test_golden_fixtures_on_gpu():
    """the committed golden vectors (recorded from the reference's own Python,
    tests/golden/make_golden.py) replayed through the C ABI on the GPU"""
    import torch
    import test_cpu_golden as tg
    from masurvival import _lib
    import pyoracle as po
    for path in tg.GOLDEN:
        g = np.load(path)
        rec = tg.case_config(path)
        A = int(rec['n_agents'])
        seed, env_id, _ = [int(v) for v in g['meta']]
        h = _lib.Handle(rec, 1, 0, seed, env_id)
        keys = list(po.obs_dims(rec).keys())
        for r in range(len(g['kind'])):
            if g['kind'][r] == 0:
                h.reset()
            else:
                a = torch.as_tensor(g['actions'][r][None]).cuda()
                h.step(a.data_ptr())
            torch.cuda.synchronize()
            for k in keys:
                v = h.tensor(k).cpu().numpy()[0]
                if k in ('zone', 'heals', 'boxes', 'box_items'):
                    v = np.broadcast_to(v[None], (A,) + v.shape)
                assert np.array_equal(v, g[k][r]), (path, r, k)
            if g['kind'][r] == 1:
                assert np.array_equal(h.tensor('rewards').cpu().numpy()[0], g['rewards'][r]), (path, r)
                assert bool(h.tensor('dones').cpu().numpy()[0]) == bool(g['done'][r]), (path, r)
        h.close()


def test_terminal_observation_capture():
    """auto_reset='terminal': same trajectory as the in-kernel reset, and the finished
    episode's last observation equals what an env without auto-reset returns"""
    import torch
    from masurvival.envs import MaSurvivalVec
    from masurvival.config import variant
    N = 512
    cfg = parity.apply_overrides(variant('2v2'), {'safe_zone': {'cooldown': 6}, 'health': {'health': 10}})
    et = MaSurvivalVec(cfg, N, seed=9, auto_reset='terminal')
    e1 = MaSurvivalVec(cfg, N, seed=9, auto_reset=True)
    e0 = MaSurvivalVec(cfg, N, seed=9, auto_reset=False)
    for e in (et, e1, e0):
        e.reset()
    rng = np.random.default_rng(1)
    fresh = torch.ones(N, dtype=torch.bool, device='cuda')     # e0 envs still in their first episode
    checked = 0
    for t in range(90):
        a = torch.as_tensor(random_actions(rng, N, 4)).cuda()
        ot, rt, dt, info = et.step(a)
        o1, r1, d1, _ = e1.step(a)
        o0, r0, d0, _ = e0.step(a)
        for k in o1:
            assert torch.equal(ot[k], o1[k]), k
        assert torch.equal(rt, r1) and torch.equal(dt, d1)
        sel = dt & fresh
        if sel.any():
            for k, v in info['terminal_observation'].items():
                assert torch.equal(v[sel], o0[k][sel]), k
            checked += int(sel.sum())
        fresh &= ~d0
    assert et.get_state().tobytes() == e1.get_state().tobytes()
    assert checked > 100
    for e in (et, e1, e0):
        e.close()


def test_abi_tensor_info_errors_and_lifetime():
    """msv_tensor_info shapes/strides/dtypes, unknown names, byte accounting, state round trip"""
    import ctypes
    import torch
    from masurvival import _lib
    rec = make_config('ffa_lidar', auto_reset=True)
    h = _lib.Handle(rec, 100, 0, 1, 0)                 # 100 is not a multiple of the 64-env block
    L = _lib.load()
    h.reset()
    ptr, nd, dt = ctypes.c_void_p(), ctypes.c_int32(), ctypes.c_int32()
    shape, strides = (ctypes.c_int64 * 4)(), (ctypes.c_int64 * 4)()
    rc = L.msv_tensor_info(h.h, b'others', ctypes.byref(ptr), ctypes.byref(nd), shape, strides, ctypes.byref(dt))
    assert rc == 0 and nd.value == 4 and list(shape) == [100, 8, 7, 8] and list(strides) == [448, 56, 8, 1] and dt.value == 0
    assert L.msv_tensor_info(h.h, b'lidar_hit', ctypes.byref(ptr), ctypes.byref(nd), shape, strides, ctypes.byref(dt)) == 0 and dt.value == 2
    assert L.msv_tensor_info(h.h, b'dones', ctypes.byref(ptr), ctypes.byref(nd), shape, strides, ctypes.byref(dt)) == 0 and dt.value == 1 and list(shape)[:1] == [100]
    assert L.msv_tensor_info(h.h, b'nope', None, None, None, None, None) == -4      # MSV_ERR_NAME
    assert b'nope' in L.msv_last_error(h.h)
    t = h.tensor('agent')
    assert t.shape == (100, 8, 8) and t.is_cuda and t.data_ptr() == h.tensor('agent').data_ptr()
    assert h.tensor('lidar_hit').dtype == torch.int32 and h.tensor('dones').dtype == torch.uint8
    assert h.bytes_per_env_step() > h.obs_bytes_per_env() > 0
    # checkpoint / resume: get_state -> set_state is the identity, and a restored env continues identically
    a = torch.zeros((100, 8, 6), dtype=torch.uint8, device='cuda'); a[..., :3] = 2; a[..., 3] = 1
    for _ in range(25):
        h.step(a.data_ptr())
    snap = h.get_state()
    for _ in range(10):
        h.step(a.data_ptr())
    torch.cuda.synchronize()
    ref_obs = h.tensor('agent').clone(); ref_state = h.get_state()
    h.set_state(snap)
    assert h.get_state().tobytes() == snap.tobytes()
    for _ in range(10):
        h.step(a.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(h.tensor('agent'), ref_obs) and h.get_state().tobytes() == ref_state.tobytes()
    assert h.get_state(first=99, count=1).shape == (1,)
    with pytest.raises(_lib.MasurvError):
        h.get_state(first=100, count=1)
    h.close()
