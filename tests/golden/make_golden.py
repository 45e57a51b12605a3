#!/usr/bin/env python
"""Generate the golden input/output vectors under tests/golden/ by running the
reference's OWN, UNMODIFIED `masurvival` package (imported from
/root/reference) on the Box2D/gym shims of oracle/shim/.

    python tests/golden/make_golden.py        # needs /root/reference (this container only)

For every case it drives the reference env with a scripted + random policy,
feeding its numpy-Generator call sites (SpawnGrid shuffle, RandomizeBoxShapes
normal, SafeZone/DeathDrop random) from the same counter-based Philox streams
the oracle and the CUDA kernel use, and records actions, observations, rewards
and dones.  While generating it also steps the C oracle in lock-step and
asserts bit-equality, so a committed fixture is by construction one the oracle
reproduces.  The fixtures pin layers L1-L3 (reference Python semantics) on the
restated L0 (b2lite); see DESIGN.md "Parity status".
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shim'))

import parity  # noqa: E402  (product config helpers: variant/make_config)
import pyoracle as po  # noqa: E402
import Box2D  # noqa: E402,F401  (the shim)
import gym  # noqa: E402,F401

# switch `masurvival` from the product package to the reference package
for m in [k for k in sys.modules if k == 'masurvival' or k.startswith('masurvival.')]:
    del sys.modules[m]
sys.path = [p for p in sys.path if 'gym-ma-survival-2d_b200' not in p]
sys.path.insert(0, '/root/reference')
from masurvival.envs.masurvival_env import MaSurvival  # noqa: E402
import masurvival  # noqa: E402
assert masurvival.__file__.startswith('/root/reference'), masurvival.__file__

STREAM_SHUFFLE, STREAM_BOX, STREAM_ZONE, STREAM_DEATH = 0, 1, 2, 3


class PhiloxGenerator:
    """Stands in for `np.random.default_rng()` inside the reference env
    (env:50,67,455-465): same three methods, draws from Philox4x32-10 keyed by
    (seed, env id, episode, step) exactly like oracle/masurv_oracle.c."""

    def __init__(self, seed, env_id):
        self.seed, self.env_id, self.episode, self.step = seed, env_id, -1, 0
        self._zone = self._box = 0

    def begin_episode(self):
        self.episode += 1
        self.step = 0
        self._zone = self._box = 0

    def _u(self, stream, k, step=0):
        return po.philox_uniform(self.seed, self.env_id, self.episode, step, stream, k)

    def shuffle(self, x):                      # semantics.py:74
        n = len(x)
        for i in range(n - 1, 0, -1):
            j = int(self._u(STREAM_SHUFFLE, n - 1 - i) * (i + 1))
            x[i], x[j] = x[j], x[i]

    def normal(self, loc=0.0, scale=1.0):      # semantics.py:111-118
        k = self._box
        self._box += 1
        u1, u2 = self._u(STREAM_BOX, 2 * k), self._u(STREAM_BOX, 2 * k + 1)
        z = math.sqrt(-2.0 * math.log(1.0 - u1)) * math.cos(6.283185307179586 * u2)
        return loc + scale * z

    def random(self, n=None):
        if n is None:                          # semantics.py:745-746
            k = self._zone
            self._zone += 1
            return self._u(STREAM_ZONE, k)
        return np.array([self._u(STREAM_DEATH, k, self.step) for k in range(int(n))])  # semantics.py:391


def ref_config(name, over):
    user = parity.apply_overrides(parity.variant(name), over)
    for k in ('lidars', 'modules', 'box2d'):       # extension keys of the product config: not part of the reference's
        user.pop(k, None)
    import masurvival.simulation as rsim          # the reference wants b2CircleShape objects (env:222-227)
    for k in ('auto_pickup', 'give'):
        if k in user and not hasattr(user[k].get('shape'), 'radius'):
            user[k] = dict(user[k]); user[k]['shape'] = rsim.circle_shape(user[k]['shape'])
    return user


def scripted_actions(obs, t, rng, A, p_attack=0.6, p_use_heal=0.5, p_use_box=0.2, p_give=0.08):
    """seek-and-interact policy so that pickups, heals, box placement, melee
    kills, death drops and gives all happen within a few hundred steps"""
    acts = np.zeros((A, 6), dtype=np.uint8)
    off = obs['agent'].shape[1] - 6
    for i in range(A):
        row = obs['agent'][i]
        x, y, ang = row[off], row[off + 1], row[off + 2]
        mode = (t // 40 + i) % 4
        target = None
        if mode == 0 and 'heals' in obs:
            m = obs['heals_mask'][i] == 0
            if m.any():
                p = obs['heals'][i][m]
                target = p[np.argmin(((p - [x, y]) ** 2).sum(1))]
        elif mode == 1 and 'boxes' in obs:
            m = obs['boxes_mask'][i] == 0
            if m.any():
                p = obs['boxes'][i][m][:, 8:10]
                target = p[np.argmin(((p - [x, y]) ** 2).sum(1))]
        elif mode == 2 and 'box_items' in obs:
            m = obs['box_items_mask'][i] == 0
            if m.any():
                p = obs['box_items'][i][m][:, 8:10]
                target = p[np.argmin(((p - [x, y]) ** 2).sum(1))]
        if target is None:
            others = obs['others'][i]
            alive = others[:, off - 1] > 0
            if alive.any():
                p = others[alive][:, off:off + 2]
                target = p[np.argmin(((p - [x, y]) ** 2).sum(1))]
        if target is None or rng.random() < 0.15:
            acts[i, 0:3] = rng.integers(0, 3, 3)
        else:
            bearing = math.atan2(target[1] - y, target[0] - x) - ang
            bearing = (bearing + math.pi) % (2 * math.pi) - math.pi
            acts[i, 0] = 2 if abs(bearing) < 1.0 else 1
            acts[i, 1] = 1
            acts[i, 2] = 2 if bearing > 0.05 else (0 if bearing < -0.05 else 1)
        acts[i, 3] = rng.random() < p_attack
        has_heal = 'heal_slot_mask' in obs and obs['heal_slot_mask'][i][0] == 0
        has_box = 'box_slot_mask' in obs and obs['box_slot_mask'][i][0] == 0
        hp = row[off - 1]
        acts[i, 4] = (has_heal and hp < 70 and rng.random() < p_use_heal) or (has_box and rng.random() < p_use_box) or rng.random() < 0.03
        acts[i, 5] = rng.random() < p_give
    return acts


def wire_unused_modules(env, rec):
    """Append the modules the reference ships but never instantiates to its `agents` group, at the
    places its own source marks (commented lines env:336,338; Lidars has no marked place and goes
    last: it then scans the state the observation describes).  Nothing in /root/reference is
    modified: the instances are added to the live env object."""
    import masurvival.simulation as rsim
    import masurvival.semantics as rsem
    mods = env.simulation.groups['agents'].modules
    extra = {}
    if int(rec['immunity_cooldown']) >= 0:
        i = [k for k, m in enumerate(mods) if isinstance(m, rsem.SafeZone)][0] + 1      # env:336
        extra['immunity'] = rsem.ImmunityPhase(int(rec['immunity_cooldown']))
        mods.insert(i, extra['immunity'])
    if int(rec['battle_royale']):
        i = [k for k, m in enumerate(mods) if isinstance(m, (rsem.Melee, rsem.ContinuousMelee))][0] + 1   # env:338
        extra['br'] = rsem.BattleRoyale()
        mods.insert(i, extra['br'])
    if int(rec['lidar_n']) > 0:
        extra['lidars'] = rsim.Lidars(int(rec['lidar_n']), float(rec['lidar_fov']), float(rec['lidar_depth']))
        mods.append(extra['lidars'])
    return extra


KIND = {'agents': 1, 'boxes': 2, 'box_items': 3, 'heals': 4, 'walls': 5}


def ref_lidar(env, lid, A, L):
    """Lidars.scans (simulation.py:377-383) -> per agent INDEX arrays: fraction (1 where the ray hit
    nothing or the agent is dead) and kind << 8 | position of the hit body in its group's list
    (agents: their stable IndexBodies slot)."""
    import masurvival.simulation as rsim
    groups = env.simulation.groups
    agents = groups['agents']
    index = agents.get(rsim.IndexBodies)[0].bodies
    frac = np.ones((A, L), dtype=np.float32); hit = np.zeros((A, L), dtype=np.int32)
    for row, body in enumerate(agents.bodies):
        i = index.index(body)
        for r, scan in enumerate(lid.scans[row]):
            if scan is None:
                continue
            fixture, f = scan
            hb = fixture.body
            grp = rsim.Group.body_group(hb)
            gname = [k for k, v in groups.items() if v is grp][0]
            idx = index.index(hb) if gname == 'agents' else grp.bodies.index(hb)
            frac[i, r] = np.float32(f); hit[i, r] = (KIND[gname] << 8) | idx
    return frac, hit


def run_case(name, over, seed, env_id, steps, policy='scripted', pol={}):
    rec = parity.make_config(name, auto_reset=False, **{k: dict(v) for k, v in over.items()})
    A = int(rec['n_agents'])
    env = MaSurvival(ref_config(name, over))
    extra = wire_unused_modules(env, rec)
    gen = PhiloxGenerator(seed, env_id)
    env.np_random = gen
    orc = po.OracleEnv(rec, seed=seed, env_id=env_id)
    rng = np.random.default_rng(seed * 1000 + env_id)
    keys = [k for k in po.obs_dims(rec) if not k.startswith('lidar')]
    log = {k: [] for k in keys}
    log.update(kind=[], rewards=[], done=[], actions=[])
    xkeys = (['lidar_frac', 'lidar_hit'] if 'lidars' in extra else []) + (['immune'] if 'immunity' in extra else []) + \
            (['br_over', 'br_results'] if 'br' in extra else [])
    log.update({k: [] for k in xkeys})
    L = int(rec['lidar_n'])

    def extras(oo, tag):
        """state of the extra modules, read off the reference objects and checked against the oracle"""
        x = {}
        if 'lidars' in extra:
            x['lidar_frac'], x['lidar_hit'] = ref_lidar(env, extra['lidars'], A, L)
        if 'immunity' in extra:
            x['immune'] = np.int32(bool(env.simulation.groups['agents'].get(type(extra['immunity']))[0].health.immune))
        if 'br' in extra:
            br = extra['br']
            x['br_over'] = np.int32(bool(br.over))
            x['br_results'] = np.array(br.results if br.over else [0] * A, dtype=np.int32)   # .results only exists once over
        for k, v in x.items():
            if not np.array_equal(np.asarray(v), np.asarray(oo[k])):
                raise AssertionError(f'{name} {tag}: {k} differs: ref {v} oracle {oo[k]}')
        return x

    def record(kind, obs, rew, done, act, x={}):
        for k in keys:
            log[k].append(np.asarray(obs[k], dtype=np.float32))
        for k in xkeys:
            log[k].append(np.asarray(x[k]))
        log['kind'].append(kind); log['rewards'].append(np.asarray(rew, dtype=np.float32))
        log['done'].append(done); log['actions'].append(act)

    def check(tag, obs, oo, rew=None, done=None):
        for k in keys:
            a, b = np.asarray(obs[k]), oo[k]
            assert a.dtype == np.float32 and a.shape == b.shape, (tag, k, a.shape, b.shape)
            if not np.array_equal(a, b):
                idx = np.argwhere(a != b)[:4]
                raise AssertionError(f'{name} {tag}: key {k} differs at {idx.tolist()}: ref {a[tuple(idx[0])]} oracle {b[tuple(idx[0])]}')
        if rew is not None:
            assert np.array_equal(np.asarray(rew, dtype=np.float32), oo['rewards']), (tag, rew, oo['rewards'])
            assert bool(done) == oo['done'], (tag, done, oo['done'])

    def reset():
        gen.begin_episode()
        obs = env.reset()
        oo = orc.reset()
        check('reset', obs, oo)
        record(0, obs, np.zeros(A, np.float32), False, np.zeros((A, 6), np.uint8), extras(oo, 'reset'))
        return obs

    obs = reset()
    counters = dict(dones=0, heals_used=0, boxes_placed=0, kills=0, toi=0)
    for t in range(steps):
        act = scripted_actions(obs, t, rng, A, **pol) if policy == 'scripted' else parity.random_actions(rng, 1, A)[0]
        gen.step = env.steps
        obs, rew, done, _ = env.step(tuple(tuple(int(v) for v in a) for a in act))
        oo = orc.step(act)
        check(f'step {t}', obs, oo, rew, done)
        counters['toi'] += oo['n_toi_events']
        x = extras(oo, f'step {t}')
        if 'lidar_hit' in x:
            for kk, nm in KIND.items():
                counters['lidar_' + kk] = counters.get('lidar_' + kk, 0) + int(((x['lidar_hit'] >> 8) == nm).sum())
            counters['lidar_dead_rows'] = counters.get('lidar_dead_rows', 0) + int((obs['agent'][:, -7] == 0).sum())
        record(1, obs, rew, done, act, x)
        if done:
            counters['dones'] += 1
            st = env.flush_stats()
            so = orc.flush_stats()
            n = 2 if rec['teams'] else A
            assert st['steps'] == so['steps'] and st['heals_used'] == so['heals_used'] and st['boxes_placed'] == so['boxes_placed'], (st, so)
            for i in range(n):
                assert st[f'kills{i}'] == so['kills'][i] and abs(st[f'reward{i}'] - so['reward'][i]) < 1e-3, (st, so)
            counters['heals_used'] += st['heals_used']; counters['boxes_placed'] += st['boxes_placed']
            counters['kills'] += sum(st[f'kills{i}'] for i in range(n))
            obs = reset()
    st = env.flush_stats()
    n = 2 if rec['teams'] else A
    counters['heals_used'] += st['heals_used']; counters['boxes_placed'] += st['boxes_placed']
    counters['kills'] += sum(st[f'kills{i}'] for i in range(n))
    out = {k: np.stack(v) for k, v in log.items()}
    out['kind'] = np.array(log['kind'], dtype=np.uint8)
    out['done'] = np.array(log['done'], dtype=np.uint8)
    out['meta'] = np.array([seed, env_id, steps], dtype=np.int64)
    counters['events'] = {k: v for k, v in orc.events().items() if v}     # rule / quirk census of this trajectory (oracle/masurv_oracle.h ORC_EV_*)
    return out, counters


CASES = [
    # file, variant, overrides, seed, env_id, steps, policy
    ('g_1v1_default', '1v1', {}, 11, 0, 1200, 'scripted'),
    ('g_1v1_random', '1v1', {}, 12, 3, 300, 'random'),
    ('g_1v1_heal_only', '1v1_heal_only', {}, 13, 1, 800, 'scripted'),
    ('g_2v2_teams', '2v2', {}, 14, 2, 1200, 'scripted'),
    ('g_2v2_owned_fastzone', '2v2', {'boxes': {'ownership': True}, 'safe_zone': {'cooldown': 12}, 'health': {'health': 40}}, 15, 5, 1000, 'scripted'),
    ('g_2v2_lastalive_kills', '2v2', {'gameover': {'mode': 'lastalive'}, 'reward_scheme': {'r_kill': 5, 'r_death': -3},
                                      'melee': {'damage': 50}}, 16, 7, 1000, 'scripted'),
    ('g_ffa_randomized', 'ffa', {'health': {'health': 60}}, 17, 4, 500, 'scripted'),
    ('g_1v1_continuous_melee', '1v1', {'melee': {'cooldown': None}}, 18, 6, 600, 'scripted'),
    ('g_exotic_3agents', '1v1', parity.EXOTIC_A, 21, 10, 700, 'scripted'),
    ('g_exotic_noheals', '1v1', parity.EXOTIC_B, 22, 11, 500, 'scripted'),
    ('g_2v2_partial_obs', '2v2', {'observation': {'omniscent': False}, 'safe_zone': {'cooldown': 40}}, 19, 8, 900, 'scripted'),
    ('g_ffa_partial_obs', 'ffa', {'observation': {'omniscent': False}, 'health': {'health': 60}}, 20, 9, 300, 'scripted'),
    # the reference's own Lidars module (simulation.py:357-392) appended to its agents group: 32 and 9 rays
    ('g_ffa_lidar', 'ffa_lidar', {'health': {'health': 50}, 'safe_zone': {'cooldown': 30}}, 23, 12, 400, 'scripted'),
    ('g_2v2_lidar9', '2v2', {'lidars': {'n_lasers': 9, 'fov': 0.8 * math.pi, 'depth': 2.0}, 'safe_zone': {'cooldown': 25}, 'health': {'health': 30}},
     24, 13, 500, 'scripted'),
    # ImmunityPhase (semantics.py:652-674) and BattleRoyale (semantics.py:31-46) wired at env:336,338
    ('g_1v1_modules', '1v1', {'modules': {'immunity_phase': True, 'battle_royale': True}, 'immunity_phase': {'cooldown': 12},
                              'safe_zone': {'cooldown': 10}, 'health': {'health': 25}}, 25, 14, 500, 'scripted'),
    ('g_ffa_modules', 'ffa', {'modules': {'immunity_phase': True, 'battle_royale': True}, 'immunity_phase': {'cooldown': 0},
                              'gameover': {'mode': 'lastalive'}, 'safe_zone': {'cooldown': 10}, 'health': {'health': 25}}, 26, 15, 300, 'scripted'),
    # cases found by searching (oracle only) for the rules the cases above never fire -- see COVERAGE below
    ('g_4v4_crowd', 'ffa', {'teams': {'twoteams': True}, 'inventory': {'slots': 1}, 'give': {'shape': 3.0}, 'spawn_grid': {'grid_size': 8, 'floor_size': 14},
                            'safe_zone': {'cooldown': 80, 'radiuses': [7, 4, 2, 1]}, 'health': {'health': 80}}, 32, 12, 600, 'scripted',
     dict(p_give=0.5, p_use_heal=0.02, p_use_box=0.02, p_attack=0.3)),                                   # give to a full taker, give blocked by a stranger, pickup with a full inventory
    ('g_ffa_brawl', 'ffa', {'safe_zone': {'cooldown': 15, 'damage': 3}, 'health': {'health': 30}, 'melee': {'damage': 30, 'cooldown': 3, 'range': 2.5}},
     31, 11, 800, 'scripted', dict(p_attack=0.95)),                                                       # Q5: kill by an attacker who is dead at reward time
    ('g_ffa_hoard', 'ffa', {'inventory': {'slots': 3}, 'safe_zone': {'cooldown': 40, 'damage': 2}, 'health': {'health': 60}, 'melee': {'damage': 30, 'cooldown': 10}},
     33, 13, 800, 'scripted', dict(p_give=0.02, p_use_heal=0.02, p_use_box=0.02, p_attack=0.9)),          # death drop of a 2+ item inventory
    ('g_2v2_owned_attack', '2v2', {'boxes': {'ownership': True, 'health': 20}, 'safe_zone': {'cooldown': 150}, 'melee': {'cooldown': 5}},
     32, 12, 800, 'scripted', dict(p_attack=0.9, p_use_box=0.6)),                                         # owned box hit by a non-owner
]

# every rule / quirk of SURVEY.md 8a that must fire somewhere in the committed fixtures (Q7, two agents on
# one item in the same step, cannot: the reference destroys the body twice there -- documented deviation)
COVERAGE = ['death_by_zone', 'death_by_melee', 'multi_death_step', 'kill_ffa', 'kill_team', 'q5_dead_killer', 'give_ok', 'give_to_full',
            'give_stranger', 'give_blocked_by_body', 'deathdrop_1', 'deathdrop_2plus', 'pickup_heal', 'pickup_box', 'pickup_full',
            'box_placed', 'box_destroyed', 'q9_fresh_box_hit', 'owned_box_protected', 'melee_hit_agent', 'melee_hit_box',
            'melee_teammate_immune', 'q3_cooldown_burnt', 'q1_stale_seen_row', 'q10_saved_by_heal', 'toi_event', 'episode_end']


def main():
    only = sys.argv[1:]
    for case in CASES:
        fname, name, over, seed, env_id, steps, policy = case[:7]
        if only and fname not in only:
            continue
        out, counters = run_case(name, over, seed, env_id, steps, policy, case[7] if len(case) > 7 else {})
        path = os.path.join(HERE, fname + '.npz')
        np.savez_compressed(path, **out)
        print(f'{fname}: {len(out["kind"])} records, {os.path.getsize(path) / 1024:.0f} KiB, {counters}')


if __name__ == '__main__':
    main()
