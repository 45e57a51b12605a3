"""Box2D shim -- TEST ORACLE infrastructure, not product code.

A pure-Python stand-in for the subset of pybox2d 2.3.10 that the reference's
`masurvival` package imports, so that the reference's
OWN, UNMODIFIED Python (simulation.py / semantics.py / masurvival_env.py) can
be imported from /root/reference and run here: pybox2d itself is not
installable in this image.  All rigid-body arithmetic is delegated to
oracle/b2lite.c (the float32 restatement of the Box2D 2.3 subset) through
ctypes; vector math done on the Python side rounds to float32 after every
operation, like pybox2d's SWIG-wrapped b2Vec2.

Documented deviations from pybox2d:
 - `body.position` & co. return copies (pybox2d returns references into the
   b2Body; the reference only relies on that in Object.pre_despawn, where it
   is a use-after-free -- see DESIGN.md "Q12");
 - QueryAABB / RayCast report fixtures in body creation order (no dynamic
   tree), RayCast reports only the closest hit (the reference's callback keeps
   the closest anyway, simulation.py:471-484).
"""
import ctypes
import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _ORACLE)
import pyoracle as _po  # noqa: E402  (builds/loads liboracle.so)
from masurvival._cstruct import parse_header  # noqa: E402  (header parser only)

_DEFS, _ST = parse_header(os.path.join(_ORACLE, 'b2lite.h'))
_SHAPE_DT, _BODY_DT, _WORLD_DT = _ST['b2l_shape'], _ST['b2l_body'], _ST['b2l_world']
_L = _po.lib()
_vp, _f, _i = ctypes.c_void_p, ctypes.c_float, ctypes.c_int


class _V2(ctypes.Structure):
    _fields_ = [('x', _f), ('y', _f)]


_L.b2l_world_new.restype = _vp
_L.b2l_set_variant.argtypes = [_i]


def set_variant(v):
    """Box2D build variant of the restated subset (include/masurv.h MSV_B2_* bits); process-global"""
    _L.b2l_set_variant(int(v))


_L.b2l_world_free.argtypes = [_vp]
_L.b2l_circle.argtypes = [_vp, _f]
_L.b2l_set_as_box.argtypes = [_vp, _f, _f]
_L.b2l_polygon_set.argtypes = [_vp, _vp, _i]
_L.b2l_test_point.argtypes = [_vp, _V2, _f, _f, _V2]
_L.b2l_test_point.restype = _i
_L.b2l_shape_aabb.argtypes = [_vp, _V2, _f, _f, _vp]
_L.b2l_create_body.argtypes = [_vp, _i, _f, _f, _f, _vp, _f, _i, _f, _f]
_L.b2l_create_body.restype = _i
_L.b2l_destroy_body.argtypes = [_vp, _i]
_L.b2l_step.argtypes = [_vp, _f, _i, _i]
_L.b2l_apply_linear_impulse_center.argtypes = [_vp, _i, _f, _f, _i]
_L.b2l_apply_angular_impulse.argtypes = [_vp, _i, _f, _i]
_L.b2l_raycast.argtypes = [_vp, _V2, _V2, ctypes.POINTER(_f), ctypes.POINTER(_V2)]
_L.b2l_raycast.restype = _i
_L.b2l_query_aabb.argtypes = [_vp, _vp, _vp, _i]
_L.b2l_query_aabb.restype = _i
_L.b2l_rot.argtypes = [_f, ctypes.POINTER(_f), ctypes.POINTER(_f)]


def f32(x):
    return ctypes.c_float(x).value


def _rot(angle):
    s, c = _f(), _f()
    _L.b2l_rot(f32(angle), ctypes.byref(s), ctypes.byref(c))
    return s.value, c.value


b2_staticBody = 0
b2_kinematicBody = 1
b2_dynamicBody = 2
b2_pi = f32(3.14159265359)


class b2Vec2:
    __slots__ = ('x', 'y')

    def __init__(self, *args):
        if len(args) == 0:
            x, y = 0.0, 0.0
        elif len(args) == 1:
            x, y = args[0][0], args[0][1]
        else:
            x, y = args
        self.x, self.y = f32(x), f32(y)

    def __iter__(self):
        yield self.x
        yield self.y

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __add__(self, o):
        return b2Vec2(f32(self.x + o[0]), f32(self.y + o[1]))

    __radd__ = __add__

    def __sub__(self, o):
        return b2Vec2(f32(self.x - o[0]), f32(self.y - o[1]))

    def __rsub__(self, o):
        return b2Vec2(f32(o[0] - self.x), f32(o[1] - self.y))

    def __mul__(self, a):
        a = f32(a)
        return b2Vec2(f32(self.x * a), f32(self.y * a))

    __rmul__ = __mul__

    def __truediv__(self, a):
        a = f32(a)
        return b2Vec2(f32(self.x / a), f32(self.y / a))

    def __neg__(self):
        return b2Vec2(-self.x, -self.y)

    def __eq__(self, o):
        try:
            return self.x == o[0] and self.y == o[1]
        except Exception:
            return False

    def __hash__(self):
        return hash((self.x, self.y))

    @property
    def length(self):
        return f32(math.sqrt(f32(f32(self.x * self.x) + f32(self.y * self.y))))

    def copy(self):
        return b2Vec2(self.x, self.y)

    def __repr__(self):
        return f'b2Vec2({self.x},{self.y})'


class b2Mat22:
    """Only what sim.from_polar needs: `R.angle = a; R * b2Vec2`."""

    def __init__(self, a11=1.0, a12=0.0, a21=0.0, a22=1.0):
        self.ex = b2Vec2(a11, a21)
        self.ey = b2Vec2(a12, a22)

    @property
    def angle(self):
        return f32(math.atan2(self.ex.y, self.ex.x))

    @angle.setter
    def angle(self, a):
        s, c = _rot(a)
        self.ex = b2Vec2(c, s)
        self.ey = b2Vec2(-s, c)

    def __mul__(self, v):  # b2Mul(A, v)
        return b2Vec2(f32(f32(self.ex.x * v[0]) + f32(self.ey.x * v[1])),
                      f32(f32(self.ex.y * v[0]) + f32(self.ey.y * v[1])))


class b2Rot:
    def __init__(self, angle=0.0):
        self.s, self.c = _rot(angle)


class b2Transform:
    def __init__(self):
        self.position = b2Vec2(0, 0)
        self.q = b2Rot(0.0)
        self._angle = 0.0

    def Set(self, position=None, angle=0.0):
        self.position = b2Vec2(position if position is not None else (0, 0))
        self._angle = f32(angle)
        self.q = b2Rot(angle)

    @property
    def angle(self):
        return self._angle

    @property
    def R(self):
        return b2Mat22(self.q.c, -self.q.s, self.q.s, self.q.c)


class b2Shape:
    def __init__(self):
        self._rec = np.zeros(1, dtype=_SHAPE_DT)

    @property
    def _ptr(self):
        return self._rec.ctypes.data

    def TestPoint(self, transform, p):
        p = b2Vec2(p)
        return bool(_L.b2l_test_point(self._ptr, _V2(transform.position.x, transform.position.y),
                                      transform.q.s, transform.q.c, _V2(p.x, p.y)))

    def getAABB(self, transform, child_index=0):
        out = (_f * 4)()
        _L.b2l_shape_aabb(self._ptr, _V2(transform.position.x, transform.position.y),
                          transform.q.s, transform.q.c, out)
        return b2AABB(lowerBound=(out[0], out[1]), upperBound=(out[2], out[3]))


class b2CircleShape(b2Shape):
    def __init__(self, radius=0.0, pos=(0, 0)):
        super().__init__()
        assert tuple(pos) == (0, 0)
        _L.b2l_circle(self._ptr, f32(radius))

    @property
    def radius(self):
        return float(self._rec[0]['radius'])


class b2PolygonShape(b2Shape):
    def __init__(self, box=None, vertices=None):
        super().__init__()
        if box is not None:
            _L.b2l_set_as_box(self._ptr, f32(box[0]), f32(box[1]))
        elif vertices is not None:
            vs = np.array([[f32(v[0]), f32(v[1])] for v in vertices], dtype=np.float32)
            _L.b2l_polygon_set(self._ptr, vs.ctypes.data, len(vs))
        else:
            _L.b2l_set_as_box(self._ptr, 0.0, 0.0)

    @property
    def vertices(self):
        n = int(self._rec[0]['count'])
        return [(float(v['x']), float(v['y'])) for v in self._rec[0]['verts'][:n]]

    @property
    def radius(self):
        return float(self._rec[0]['radius'])


class b2ChainShape(b2Shape):
    pass


class b2EdgeShape(b2Shape):
    pass


class b2AABB:
    def __init__(self, lowerBound=(0, 0), upperBound=(0, 0)):
        self.lowerBound = b2Vec2(lowerBound)
        self.upperBound = b2Vec2(upperBound)


class b2FixtureDef:
    def __init__(self, shape=None, density=0.0, restitution=0.0, isSensor=False, friction=0.2):
        self.shape, self.density, self.restitution, self.isSensor = shape, density, restitution, isSensor


class b2Fixture:
    def __init__(self, body, defn):
        self.body = body
        self.shape = defn.shape
        self.density = f32(defn.density)
        self.restitution = f32(defn.restitution)
        self._sensor = bool(defn.isSensor)

    @property
    def sensor(self):
        return self._sensor

    @sensor.setter
    def sensor(self, flag):
        self._sensor = bool(flag)
        self.body._world._view['bodies'][self.body._slot]['sensor'] = int(bool(flag))


class b2Body:
    def __init__(self, world, slot, type_, fixture_def, linearDamping, userData):
        self._world, self._slot = world, slot
        self.type = type_
        self.userData = userData
        self.linearDamping = f32(linearDamping)
        self.fixtures = [b2Fixture(self, fixture_def)]

    @property
    def _b(self):
        return self._world._view['bodies'][self._slot]

    @property
    def position(self):
        p = self._b['p']
        return b2Vec2(float(p['x']), float(p['y']))

    @property
    def worldCenter(self):
        c = self._b['c']
        return b2Vec2(float(c['x']), float(c['y']))

    @property
    def angle(self):
        return float(self._b['a'])

    @property
    def linearVelocity(self):
        v = self._b['v']
        return b2Vec2(float(v['x']), float(v['y']))

    @property
    def angularVelocity(self):
        return float(self._b['w'])

    @property
    def awake(self):
        return bool(self._b['awake'])

    @property
    def transform(self):
        t = b2Transform()
        b = self._b
        t.position = b2Vec2(float(b['p']['x']), float(b['p']['y']))
        t._angle = float(b['a'])
        t.q = b2Rot.__new__(b2Rot)
        t.q.s, t.q.c = float(b['qs']), float(b['qc'])
        return t

    def ApplyLinearImpulse(self, impulse, point, wake):
        impulse, point = b2Vec2(impulse), b2Vec2(point)
        assert point == self.worldCenter, 'shim: impulses only at the centre of mass'
        _L.b2l_apply_linear_impulse_center(self._world._w, self._slot, impulse.x, impulse.y, int(bool(wake)))

    def ApplyAngularImpulse(self, impulse, wake):
        _L.b2l_apply_angular_impulse(self._world._w, self._slot, f32(impulse), int(bool(wake)))


class b2RayCastCallback:
    def __init__(self):
        pass

    def ReportFixture(self, fixture, point, normal, fraction):
        raise NotImplementedError


class b2QueryCallback:
    def __init__(self):
        pass

    def ReportFixture(self, fixture):
        raise NotImplementedError


class b2ContactListener:
    pass


class b2Joint:
    pass


class b2World:
    def __init__(self, gravity=(0, 0), doSleep=True):
        assert tuple(gravity) == (0, 0) and doSleep
        self._w = _L.b2l_world_new()
        buf = (ctypes.c_char * _WORLD_DT.itemsize).from_address(self._w)
        self._view = np.frombuffer(buf, dtype=_WORLD_DT, count=1)[0]
        self._bodies = {}

    def __del__(self):
        try:
            _L.b2l_world_free(self._w)
        except Exception:
            pass

    def CreateBody(self, type=b2_staticBody, position=(0, 0), angle=0.0, fixtures=None,
                   linearDamping=0.0, angularDamping=0.0, userData=None):
        position = b2Vec2(position)
        slot = _L.b2l_create_body(self._w, int(type), position.x, position.y, f32(angle), fixtures.shape._ptr,
                                  f32(fixtures.density), int(bool(fixtures.isSensor)),
                                  f32(linearDamping), f32(angularDamping))
        body = b2Body(self, slot, type, fixtures, linearDamping, userData)
        self._bodies[slot] = body
        return body

    def DestroyBody(self, body):
        if self._bodies.get(body._slot) is body:
            _L.b2l_destroy_body(self._w, body._slot)
            del self._bodies[body._slot]

    def Step(self, timeStep, velocityIterations, positionIterations):
        _L.b2l_step(self._w, f32(timeStep), int(velocityIterations), int(positionIterations))

    def ClearForces(self):
        pass

    def RayCast(self, callback, point1, point2):
        p1, p2 = b2Vec2(point1), b2Vec2(point2)
        fr, nrm = _f(), _V2()
        slot = _L.b2l_raycast(self._w, _V2(p1.x, p1.y), _V2(p2.x, p2.y), ctypes.byref(fr), ctypes.byref(nrm))
        if slot >= 0:
            f = fr.value
            point = b2Vec2(f32(f32((1.0 - f)) * p1.x) + f32(f * p2.x), f32(f32((1.0 - f)) * p1.y) + f32(f * p2.y))
            callback.ReportFixture(self._bodies[slot].fixtures[0], point, b2Vec2(nrm.x, nrm.y), f)

    def QueryAABB(self, callback, aabb):
        box = (_f * 4)(aabb.lowerBound.x, aabb.lowerBound.y, aabb.upperBound.x, aabb.upperBound.y)
        out = (_i * 256)()
        n = _L.b2l_query_aabb(self._w, box, out, 256)
        for k in range(n):
            if not callback.ReportFixture(self._bodies[out[k]].fixtures[0]):
                break
