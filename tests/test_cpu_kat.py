"""Closed-form known-answer tests for the CPU oracle: the
reference ships no golden vectors for the Box2D layer, so b2lite is pinned by
physics/geometry identities that hold for Box2D v2.3 by construction."""
import math

import numpy as np
import pytest

import parity
import pyoracle as po

f32 = np.float32
DT = f32(1.0 / 60)
DAMP = f32(1.0) / (f32(1.0) + DT * f32(0.8))          # b2Island::Solve, Pade damping
R = f32(0.5)
MASS = f32(1.0) * f32(3.14159265359) * R * R           # b2CircleShape::ComputeMass
INV_M = f32(1.0) / MASS
INV_I = f32(1.0) / (MASS * (f32(0.5) * R * R))
NOOP = [1, 1, 1, 0, 0, 0]


def fresh(variant='1v1', seed=3, **over):
    rec = parity.make_config(variant, **over)
    o = po.OracleEnv(rec, seed=seed, env_id=0)
    o.reset()
    return rec, o


def place(s, i, x, y, angle=0.0, vx=0.0, vy=0.0, w=0.0):
    s['x'][i], s['y'][i], s['angle'][i] = x, y, angle
    s['vx'][i], s['vy'][i], s['omega'][i] = vx, vy, w
    s['fat'][i] = [f32(x) - R - f32(0.1), f32(y) - R - f32(0.1), f32(x) + R + f32(0.1), f32(y) + R + f32(0.1)]
    return s


def clear_world(s):
    """remove boxes/heals so nothing interferes"""
    s['n_boxes'] = 0; s['n_heals'] = 0; s['n_items'] = 0
    for f in ('box_x', 'box_y', 'box_health', 'box_has_health', 'box_cause', 'box_owner', 'box_seq', 'heal_x', 'heal_y', 'heal_seq'):
        s[f][:] = 0
    s['box_shape'][:] = np.zeros((), dtype=s['box_shape'].dtype)
    return s


def test_free_flight_damping_and_integration():
    rec, o = fresh()
    s = clear_world(o.get_state().copy())
    place(s, 0, 0.0, 0.0, 0.3, vx=1.0, vy=-2.0, w=0.5); place(s, 1, 5.0, 5.0)
    o.set_state(s)
    o.step([NOOP, NOOP])
    t = o.get_state()
    v1x, v1y = f32(1.0) * DAMP, f32(-2.0) * DAMP
    v2x, v2y = v1x * DAMP, v1y * DAMP
    x = f32(0.0) + DT * v1x; x = x + DT * v2x
    y = f32(0.0) + DT * v1y; y = y + DT * v2y
    assert t['vx'][0] == v2x and t['vy'][0] == v2y and t['x'][0] == x and t['y'][0] == y
    w1 = f32(0.5) * DAMP; w2 = w1 * DAMP
    a = f32(0.3) + DT * w1; a = a + DT * w2
    assert t['omega'][0] == w2 and t['angle'][0] == a
    assert abs(float(DAMP) - 0.986842) < 1e-6


def test_motor_impulse():
    rec, o = fresh()
    s = clear_world(o.get_state().copy())
    place(s, 0, 0.0, 0.0, 0.0); place(s, 1, 5.0, 5.0)
    o.set_state(s)
    o.step([[2, 1, 2, 0, 0, 0], NOOP])               # forward + turn left
    t = o.get_state()
    dv = INV_M * f32(0.25); dw = INV_I * f32(0.0125)
    assert abs(float(dv) - 0.31831) < 1e-5 and abs(float(dw) - 0.12732) < 1e-5
    assert t['vx'][0] == dv * DAMP * DAMP and t['vy'][0] == 0.0
    assert t['omega'][0] == dw * DAMP * DAMP
    # sideways control acts along the body-frame y axis: rotate the agent by 90 degrees
    s2 = place(s.copy(), 0, 0.0, 0.0, math.pi / 2)
    o.set_state(s2)
    o.step([[2, 1, 1, 0, 0, 0], NOOP])
    t = o.get_state()
    assert abs(t['vx'][0]) < 2e-8 and abs(float(t['vy'][0]) - float(dv * DAMP * DAMP)) < 1e-7


def test_lidar_ray_fractions_circle_wall_box():
    rec, o = fresh('1v1', lidars={'n_lasers': 3, 'fov': math.pi, 'depth': 10})
    s = clear_world(o.get_state().copy())
    place(s, 0, 0.0, 0.0, 0.0); place(s, 1, -6.0, -6.0)
    s['n_heals'] = 1; s['heal_x'][0], s['heal_y'][0], s['heal_seq'][0] = 3.0, 0.0, 4
    s['n_boxes'] = 1; s['box_x'][0], s['box_y'][0] = 0.0, 4.0
    s['box_shape'][0]['hx'], s['box_shape'][0]['hy'] = 0.5, 0.75
    s['box_health'][0], s['box_has_health'][0], s['box_cause'][0], s['box_owner'][0] = 20, 1, -1, -1
    o.set_state(s)
    out = o.observe()
    fr, hit = out['lidar_frac'][0], out['lidar_hit'][0]
    # ray 0 points to -y (angle -pi/2): south wall inner face at y = -9.9
    assert hit[0] == (po.DEFINES['ORC_KIND_WALL'] << 8 | 3) and abs(fr[0] - 0.99) < 1e-6
    # ray 1 points to +x: heal circle r=0.25 centred at x=3
    assert hit[1] == (po.DEFINES['ORC_KIND_HEAL'] << 8 | 0) and abs(fr[1] - 0.275) < 1e-6
    # ray 2 points to +y: box bottom face at y = 4 - 0.75
    assert hit[2] == (po.DEFINES['ORC_KIND_BOX'] << 8 | 0) and abs(fr[2] - 0.325) < 1e-6


def test_zone_schedule():
    """hold 100, shrink 100 (linear), 5 phases, endgame at step 800 (sem:776-811)"""
    rec, o = fresh('1v1', health={'health': 100000})
    radii = [10, 5, 2.5, 1, 0]
    for t in range(1, 901):
        out = o.step([NOOP, NOOP])
        r = float(out['zone'][0][2])
        phase, k = divmod(t, 200)
        if t >= 800:
            assert r == 0.0 and tuple(out['zone'][0][3:]) == (0.0, 0.0, 0.0)
        elif k <= 100:
            assert r == f32(radii[phase])
            assert out['zone'][0][5] == f32(radii[phase + 1])
        else:
            tt = (200 - k) / 100
            assert r == f32(tt * radii[phase] + (1 - tt) * radii[phase + 1])
    st = o.get_state()
    assert st['zone_endgame'] == 1 and st['zone_phase'] == 4


def test_melee_cooldown_cadence_and_kill():
    """hit at t, next possible hit at t+40 (Q3); 5 hits kill; kill credit (Q5)"""
    rec, o = fresh('1v1', reward_scheme={'r_alive': 0, 'r_dead': 0, 'r_kill': 7, 'r_death': -2})
    s = clear_world(o.get_state().copy())
    place(s, 0, 0.0, 0.0, 0.0); place(s, 1, 1.5, 0.0, math.pi)
    for z in range(5):
        s['zone_cx'][z] = 0; s['zone_cy'][z] = 0
    s['zone_cur_x'] = 0; s['zone_cur_y'] = 0
    o.set_state(s)
    hp = []
    for t in range(170):
        out = o.step([[1, 1, 1, 1, 0, 0], NOOP])
        hp.append(float(out['agent'][1][1]))
        if t == 160:
            assert tuple(out['rewards']) == (7.0, -2.0) and not out['done']
    assert hp[0] == 80 and hp[39] == 80 and hp[40] == 60 and hp[80] == 40 and hp[120] == 20 and hp[159] == 20
    assert hp[160] == 0 and o.get_state()['alive'][1] == 0      # dies in the step of the 5th hit


def test_team_immunity_consumes_cooldown():
    rec, o = fresh('2v2')
    s = clear_world(o.get_state().copy())
    place(s, 0, 0.0, 0.0, 0.0); place(s, 1, 1.5, 0.0, 0.0); place(s, 2, -5.0, 5.0); place(s, 3, 5.0, -5.0)
    o.set_state(s)
    out = o.step([[1, 1, 1, 1, 0, 0], NOOP, NOOP, NOOP])
    st = o.get_state()
    assert out['agent'][1][2] == 100 and st['cooldown'][0] == 39          # teammate immune, cooldown burnt (Q3)


def test_resting_contact_against_wall():
    rec, o = fresh('1v1', health={'health': 100000})
    s = clear_world(o.get_state().copy())
    place(s, 0, 8.0, 0.0, 0.0); place(s, 1, -5.0, 5.0)
    o.set_state(s)
    for t in range(240):
        out = o.step([[2, 1, 1, 0, 0, 0], NOOP])
    x = float(out['agent'][0][2])
    # wall face at 9.9, polygon skin 0.01, circle 0.5: touching at 9.39, rest within linearSlop of it
    assert 9.385 < x < 9.41, x
    assert abs(float(out['agent'][0][5])) < 0.35     # velocity re-built by one step of thrust only
    p = o.get_state()['pair_aw'][0][2]
    assert p['seq'] > 0 and (p['flags'] & 1)      # the wall contact exists and is touching


def test_box_item_cycle_and_vertex_order():
    """box dies -> item next step -> pickup -> place re-creates it 0.75 ahead (Q8, Q9)"""
    rec, o = fresh()
    s = o.get_state().copy()
    s['n_heals'] = 0
    place(s, 0, float(s['box_x'][0]) - 1.2, float(s['box_y'][0]), 0.0); place(s, 1, -9.0, -9.0)
    o.set_state(s)
    bx, by = float(s['box_x'][0]), float(s['box_y'][0])
    out = o.step([[1, 1, 1, 1, 0, 0], NOOP])
    assert o.get_state()['n_boxes'] == 3 and o.get_state()['n_pending'] == 1
    out = o.step([NOOP, NOOP])
    st = o.get_state()
    assert st['n_items'] == 1 and st['item_x'][0] == f32(bx) and st['item_y'][0] == f32(by)
    # SetAsBox order (-,-),(+,-),(+,+),(-,+) became the re-hulled order starting at (+,-)
    assert tuple(out['box_items'][0][0][:8]) == (0.5, -0.5, 0.5, 0.5, -0.5, 0.5, -0.5, -0.5)
    assert tuple(out['boxes'][0][0][:8]) == (-0.5, -0.5, 0.5, -0.5, 0.5, 0.5, -0.5, 0.5)
    for t in range(40):
        out = o.step([[2, 1, 1, 0, 0, 0], NOOP])
        if o.get_state()['inv_n'][0] == 1:
            break
    assert o.get_state()['inv_n'][0] == 1 and out['box_slot_mask'][0][0] == 0
    before = o.get_state()
    out = o.step([[1, 1, 1, 0, 1, 0], NOOP])
    st = o.get_state()
    assert st['n_boxes'] == 4 and st['inv_n'][0] == 0 and st['box_shape'][3]['rehulled'] == 1
    ang = float(before['angle'][0])
    assert abs(float(st['box_x'][3]) - (float(before['x'][0]) + 0.75 * math.cos(ang))) < 1e-5
    assert st['box_has_health'][3] == 1 and st['box_health'][3] == 20


# ---- Box2D build variants (masurv.h MSV_B2_*): each switchable detail has its own known answer --------
def test_b2_variant_clamp_damping():
    """bit 0: Box2D <= 2.2 damping v *= clamp(1 - h*d, 0, 1) instead of the Pade form of 2.3.x"""
    out = {}
    for var in (0, 1):
        rec, o = fresh(box2d={'variant': var})
        s = clear_world(o.get_state().copy())
        place(s, 0, 0.0, 0.0, 0.0, vx=1.0, vy=-2.0, w=0.5); place(s, 1, 5.0, 5.0)
        o.set_state(s)
        o.step([NOOP, NOOP])
        out[var] = o.get_state()
    clampd = f32(1.0) - DT * f32(0.8)
    assert out[0]['vx'][0] == f32(1.0) * DAMP * DAMP
    assert out[1]['vx'][0] == f32(1.0) * clampd * clampd and out[1]['omega'][0] == f32(0.5) * clampd * clampd
    assert out[0]['vx'][0] != out[1]['vx'][0]


def test_b2_variant_weld_tolerance():
    """bit 1: b2PolygonShape::Set welds vertices whose SQUARED distance is below 0.5*linearSlop in 2.3.0
    (so below 0.05 apart), below (0.5*linearSlop)^2 in later releases"""
    import ctypes
    L = po.lib()
    L.b2l_set_variant.argtypes = [ctypes.c_int]
    L.b2l_polygon_set.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    L.b2l_polygon_set.restype = ctypes.c_int
    shape = np.zeros(256, dtype=np.uint8)                       # opaque b2l_shape
    tiny = np.array([[-0.02, -0.02], [0.02, -0.02], [0.02, 0.02], [-0.02, 0.02]], dtype=np.float32)   # 0.04 edges
    normal = np.array([[-0.5, -0.5], [0.5, -0.5], [0.5, 0.5], [-0.5, 0.5]], dtype=np.float32)
    try:
        L.b2l_set_variant(0)
        assert L.b2l_polygon_set(shape.ctypes.data, tiny.ctypes.data, 4) == -1      # welded down to < 3 vertices
        assert L.b2l_polygon_set(shape.ctypes.data, normal.ctypes.data, 4) == 0
        L.b2l_set_variant(2)
        assert L.b2l_polygon_set(shape.ctypes.data, tiny.ctypes.data, 4) == 0       # 0.04 > 0.0025: kept
    finally:
        L.b2l_set_variant(0)


def test_b2_variant_toi_substeps():
    """bit 2: a contact leaves SolveTOI at toiCount >= b2_maxSubSteps instead of > : replaying the recorded
    actions of the g_ffa_hoard fixture (an agent gets wedged at record 275), the default reproduces the
    reference-recorded trajectory with 84 TOI events, the variant departs from it at the wedge"""
    import test_cpu_golden as tg
    path = [p for p in tg.GOLDEN if 'g_ffa_hoard' in p][0]
    g = np.load(path)
    variant, over = tg.CASE_CFG['g_ffa_hoard']
    seed, env_id, _ = [int(v) for v in g['meta']]
    res = {}
    for var in (0, 4):
        o2 = {k: dict(v) for k, v in over.items()}
        o2['box2d'] = {'variant': var}
        orc = po.OracleEnv(parity.make_config(variant, auto_reset=False, **o2), seed=seed, env_id=env_id)
        n, first = 0, None
        for r in range(400):
            if g['kind'][r] == 0:
                out = orc.reset()
            else:
                out = orc.step(g['actions'][r]); n += out['n_toi_events']
            if first is None and not np.array_equal(out['agent'], g['agent'][r]):
                first = r
        res[var] = (n, first)
    assert res[0][1] is None and res[4][1] == 275 and res[4][0] < res[0][0]
