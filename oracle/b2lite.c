/*
 * b2lite.c -- TEST ORACLE (see b2lite.h header comment for scope, parity
 * status and documented deviations).  Compile with -O2 -ffp-contract=off.
 *
 * Every routine names the Box2D 2.3.x function it restates.  The reference's
 * call sites into this layer are: simulation.py:199-206 (CreateBody /
 * DestroyBody), :237-240 (Step, ClearForces), :423-424 (impulses), :434
 * (RayCast), :449-455 (getAABB / QueryAABB / TestPoint), semantics.py:764.
 */
#include "b2lite.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>

/* b2Settings.h */
#define b2_pi 3.14159265359f
#define b2_epsilon FLT_EPSILON
#define b2_maxFloat FLT_MAX
#define b2_linearSlop 0.005f
#define b2_angularSlop (2.0f / 180.0f * b2_pi)
#define b2_polygonRadius (2.0f * b2_linearSlop)
#define b2_aabbExtension 0.1f
#define b2_aabbMultiplier 2.0f
#define b2_velocityThreshold 1.0f
#define b2_maxLinearCorrection 0.2f
#define b2_maxTranslation 2.0f
#define b2_maxTranslationSquared (b2_maxTranslation * b2_maxTranslation)
#define b2_maxRotation (0.5f * b2_pi)
#define b2_maxRotationSquared (b2_maxRotation * b2_maxRotation)
#define b2_baumgarte 0.2f
#define b2_toiBaugarte 0.75f
#define b2_timeToSleep 0.5f
#define b2_linearSleepTolerance 0.01f
#define b2_angularSleepTolerance (2.0f / 180.0f * b2_pi)
#define b2_maxSubSteps 8
#define b2_maxTOIContacts 32
#define b2_maxPolygonVertices 8
#define B2L_FRICTION 0.2f /* b2FixtureDef default; reference never sets it */

typedef b2l_vec2 v2;
static inline v2 V(float x, float y) { v2 r = {x, y}; return r; }
static inline v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 vmul(float s, v2 a) { return V(s * a.x, s * a.y); }
static inline v2 vneg(v2 a) { return V(-a.x, -a.y); }
static inline float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float vcross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static inline v2 vcross_vs(v2 a, float s) { return V(s * a.y, -s * a.x); }
static inline v2 vcross_sv(float s, v2 a) { return V(-s * a.y, s * a.x); }
static inline float vlen(v2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
static inline float vlen2(v2 a) { return a.x * a.x + a.y * a.y; }
static inline float vnormalize(v2* a) {
  float length = vlen(*a);
  if (length < b2_epsilon) return 0.0f;
  float inv = 1.0f / length;
  a->x *= inv; a->y *= inv;
  return length;
}
static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }
static inline float fclamp(float a, float lo, float hi) { return fmaxf_(lo, fminf_(a, hi)); }

typedef struct { v2 p; float s, c; } xform;
static inline v2 qmul(float s, float c, v2 v) { return V(c * v.x - s * v.y, s * v.x + c * v.y); }
static inline v2 qmulT(float s, float c, v2 v) { return V(c * v.x + s * v.y, -s * v.x + c * v.y); }
static inline v2 xmul(xform T, v2 v) {
  return V((T.c * v.x - T.s * v.y) + T.p.x, (T.s * v.x + T.c * v.y) + T.p.y);
}
static inline v2 xmulT(xform T, v2 v) {
  float px = v.x - T.p.x, py = v.y - T.p.y;
  return V(T.c * px + T.s * py, -T.s * px + T.c * py);
}

/* b2Rot::Set -- sinf/cosf; evaluated in double and rounded so that every
 * implementation (oracle, shim, CUDA) gets the correctly rounded float. */
void b2l_rot(float angle, float* s, float* c) {
  *s = (float)sin((double)angle);
  *c = (float)cos((double)angle);
}

/* ------------------------------------------------------------------ shapes */
void b2l_circle(b2l_shape* s, float radius) {
  memset(s, 0, sizeof *s);
  s->type = B2L_CIRCLE; s->radius = radius; s->count = 1;
}

/* b2PolygonShape::SetAsBox */
void b2l_set_as_box(b2l_shape* s, float hx, float hy) {
  memset(s, 0, sizeof *s);
  s->type = B2L_POLYGON; s->radius = b2_polygonRadius; s->count = 4;
  s->verts[0] = V(-hx, -hy); s->verts[1] = V(hx, -hy);
  s->verts[2] = V(hx, hy);   s->verts[3] = V(-hx, hy);
  s->normals[0] = V(0.0f, -1.0f); s->normals[1] = V(1.0f, 0.0f);
  s->normals[2] = V(0.0f, 1.0f);  s->normals[3] = V(-1.0f, 0.0f);
}

/* b2PolygonShape::Set (2.3.0): weld, gift-wrap from the right-most (lowest on
 * ties) point, normals = normalize(cross(edge, 1)). */
/* Details that differ between Box2D 2.3.x builds (masurv.h MSV_B2_*): the real pybox2d cannot be run in
 * this image, so the choices it would pin are switchable.  Process-global: the oracle sets it from the
 * env's config at every entry point. */
static int g_b2_variant = 0;
void b2l_set_variant(int v) { g_b2_variant = v; }
int b2l_get_variant(void) { return g_b2_variant; }

int b2l_polygon_set(b2l_shape* s, const v2* vertices, int count) {
  memset(s, 0, sizeof *s);
  s->type = B2L_POLYGON; s->radius = b2_polygonRadius;
  int n = count < b2_maxPolygonVertices ? count : b2_maxPolygonVertices;
  v2 ps[b2_maxPolygonVertices]; int tempCount = 0;
  for (int i = 0; i < n; ++i) {
    v2 v = vertices[i]; int unique = 1;
    for (int j = 0; j < tempCount; ++j)
      /* 2.3.0 compares the SQUARED distance with 0.5*linearSlop; later releases square the tolerance */
      if (vlen2(vsub(v, ps[j])) < ((g_b2_variant & 2) ? (0.5f * b2_linearSlop) * (0.5f * b2_linearSlop) : 0.5f * b2_linearSlop)) { unique = 0; break; }
    if (unique) ps[tempCount++] = v;
  }
  n = tempCount;
  if (n < 3) { b2l_set_as_box(s, 1.0f, 1.0f); return -1; }
  int i0 = 0; float x0 = ps[0].x;
  for (int i = 1; i < n; ++i) {
    float x = ps[i].x;
    if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
  }
  int hull[b2_maxPolygonVertices]; int m = 0; int ih = i0;
  for (;;) {
    hull[m] = ih;
    int ie = 0;
    for (int j = 1; j < n; ++j) {
      if (ie == ih) { ie = j; continue; }
      v2 r = vsub(ps[ie], ps[hull[m]]);
      v2 v = vsub(ps[j], ps[hull[m]]);
      float c = vcross(r, v);
      if (c < 0.0f) ie = j;
      if (c == 0.0f && vlen2(v) > vlen2(r)) ie = j;
    }
    ++m; ih = ie;
    if (ie == i0) break;
  }
  s->count = m;
  for (int i = 0; i < m; ++i) s->verts[i] = ps[hull[i]];
  for (int i = 0; i < m; ++i) {
    int i2 = i + 1 < m ? i + 1 : 0;
    v2 edge = vsub(s->verts[i2], s->verts[i]);
    s->normals[i] = vcross_vs(edge, 1.0f);
    vnormalize(&s->normals[i]);
  }
  return 0;
}

/* b2CircleShape::TestPoint / b2PolygonShape::TestPoint */
int b2l_test_point(const b2l_shape* s, v2 p, float qs, float qc, v2 pt) {
  if (s->type == B2L_CIRCLE) {
    v2 center = vadd(p, qmul(qs, qc, V(0.0f, 0.0f)));
    v2 d = vsub(pt, center);
    return vdot(d, d) <= s->radius * s->radius;
  }
  v2 pLocal = qmulT(qs, qc, vsub(pt, p));
  for (int i = 0; i < s->count; ++i) {
    float dot = vdot(s->normals[i], vsub(pLocal, s->verts[i]));
    if (dot > 0.0f) return 0;
  }
  return 1;
}

/* b2CircleShape::ComputeAABB / b2PolygonShape::ComputeAABB */
void b2l_shape_aabb(const b2l_shape* s, v2 p, float qs, float qc, float out[4]) {
  if (s->type == B2L_CIRCLE) {
    v2 c = vadd(p, qmul(qs, qc, V(0.0f, 0.0f)));
    out[0] = c.x - s->radius; out[1] = c.y - s->radius;
    out[2] = c.x + s->radius; out[3] = c.y + s->radius;
    return;
  }
  xform T = {p, qs, qc};
  v2 lower = xmul(T, s->verts[0]), upper = lower;
  for (int i = 1; i < s->count; ++i) {
    v2 v = xmul(T, s->verts[i]);
    lower = V(fminf_(lower.x, v.x), fminf_(lower.y, v.y));
    upper = V(fmaxf_(upper.x, v.x), fmaxf_(upper.y, v.y));
  }
  out[0] = lower.x - s->radius; out[1] = lower.y - s->radius;
  out[2] = upper.x + s->radius; out[3] = upper.y + s->radius;
}

/* b2CircleShape::RayCast / b2PolygonShape::RayCast */
int b2l_shape_raycast(const b2l_shape* sh, v2 p, float qs, float qc, v2 p1, v2 p2,
                      float maxFraction, float* fraction, v2* normal) {
  if (sh->type == B2L_CIRCLE) {
    v2 position = vadd(p, qmul(qs, qc, V(0.0f, 0.0f)));
    v2 s = vsub(p1, position);
    float b = vdot(s, s) - sh->radius * sh->radius;
    v2 r = vsub(p2, p1);
    float c = vdot(s, r);
    float rr = vdot(r, r);
    float sigma = c * c - rr * b;
    if (sigma < 0.0f || rr < b2_epsilon) return 0;
    float a = -(c + sqrtf(sigma));
    if (0.0f <= a && a <= maxFraction * rr) {
      a /= rr;
      *fraction = a;
      v2 nrm = vadd(s, vmul(a, r));
      vnormalize(&nrm);
      if (normal) *normal = nrm;
      return 1;
    }
    return 0;
  }
  v2 q1 = qmulT(qs, qc, vsub(p1, p));
  v2 q2 = qmulT(qs, qc, vsub(p2, p));
  v2 d = vsub(q2, q1);
  float lower = 0.0f, upper = maxFraction;
  int index = -1;
  for (int i = 0; i < sh->count; ++i) {
    float numerator = vdot(sh->normals[i], vsub(sh->verts[i], q1));
    float denominator = vdot(sh->normals[i], d);
    if (denominator == 0.0f) {
      if (numerator < 0.0f) return 0;
    } else {
      if (denominator < 0.0f && numerator < lower * denominator) {
        lower = numerator / denominator; index = i;
      } else if (denominator > 0.0f && numerator < upper * denominator) {
        upper = numerator / denominator;
      }
    }
    if (upper < lower) return 0;
  }
  if (index >= 0) {
    *fraction = lower;
    if (normal) *normal = qmul(qs, qc, sh->normals[index]);
    return 1;
  }
  return 0;
}

/* ------------------------------------------------------------------- world */
b2l_world* b2l_world_new(void) {
  b2l_world* w = (b2l_world*)calloc(1, sizeof(b2l_world));
  return w;
}
void b2l_world_free(b2l_world* w) { free(w); }
void b2l_world_clear(b2l_world* w) { memset(w, 0, sizeof *w); }

/* b2Body::SynchronizeTransform */
void b2l_sync_transform(b2l_body* b) {
  b2l_rot(b->a, &b->qs, &b->qc);
  /* localCenter == 0 for every body in this subset */
  v2 lc = qmul(b->qs, b->qc, V(0.0f, 0.0f));
  b->p = vsub(b->c, lc);
}

static int fat_overlap(const float* a, const float* b) {
  /* b2TestOverlap(const b2AABB&, const b2AABB&) */
  float d1x = b[0] - a[2], d1y = b[1] - a[3];
  float d2x = a[0] - b[2], d2y = a[1] - b[3];
  if (d1x > 0.0f || d1y > 0.0f) return 0;
  if (d2x > 0.0f || d2y > 0.0f) return 0;
  return 1;
}

/* b2World::CreateBody + b2Body::CreateFixture (single fixture) */
int b2l_create_body(b2l_world* w, int type, float x, float y, float angle,
                    const b2l_shape* shape, float density, int sensor,
                    float linDamp, float angDamp) {
  int slot = -1;
  for (int i = 0; i < B2L_MAX_BODIES; ++i) if (!w->bodies[i].used) { slot = i; break; }
  assert(slot >= 0);
  b2l_body* b = &w->bodies[slot];
  memset(b, 0, sizeof *b);
  b->used = 1; b->id = w->next_id++; b->seq = w->body_seq++;
  b->type = type; b->sensor = sensor; b->shape = *shape;
  b->c0 = b->c = V(x, y); b->a0 = b->a = angle; b->alpha0 = 0.0f;
  b->linDamp = linDamp; b->angDamp = angDamp;
  b->awake = 1; b->sleepTime = 0.0f;
  b2l_rot(angle, &b->qs, &b->qc);
  b->p = V(x, y);
  /* b2Body::ResetMassData; only dynamic circles and static shapes occur */
  if (type == B2L_DYNAMIC) {
    assert(shape->type == B2L_CIRCLE);
    float r = shape->radius;
    b->mass = density * b2_pi * r * r;           /* b2CircleShape::ComputeMass */
    float I = b->mass * (0.5f * r * r + 0.0f);
    b->invMass = 1.0f / b->mass;
    I -= b->mass * 0.0f;                          /* centre at the origin */
    b->I = I;
    b->invI = 1.0f / I;
  }
  /* b2Fixture::CreateProxies -> b2DynamicTree::CreateProxy (fat AABB) */
  float aabb[4];
  b2l_shape_aabb(&b->shape, b->p, b->qs, b->qc, aabb);
  b->fat[0] = aabb[0] - b2_aabbExtension; b->fat[1] = aabb[1] - b2_aabbExtension;
  b->fat[2] = aabb[2] + b2_aabbExtension; b->fat[3] = aabb[3] + b2_aabbExtension;
  b->moved = 1;
  w->newFixture = 1;
  return slot;
}

static void destroy_contact(b2l_world* w, b2l_contact* c) {
  /* b2ContactManager::Destroy: wake nothing here (no listener); if the
   * manifold had points both bodies are woken by b2Contact::Destroy. */
  if (c->pointCount > 0 && !w->bodies[c->a].sensor && !w->bodies[c->b].sensor) {
    b2l_set_awake(&w->bodies[c->a], 1);
    b2l_set_awake(&w->bodies[c->b], 1);
  }
  c->used = 0;
}

void b2l_destroy_body(b2l_world* w, int slot) {
  b2l_body* b = &w->bodies[slot];
  if (!b->used) return;
  for (int i = 0; i < B2L_MAX_CONTACTS; ++i) {
    b2l_contact* c = &w->contacts[i];
    if (c->used && (c->a == slot || c->b == slot)) destroy_contact(w, c);
  }
  b->used = 0;
}

int b2l_slot_of(b2l_world* w, int id) {
  for (int i = 0; i < B2L_MAX_BODIES; ++i)
    if (w->bodies[i].used && w->bodies[i].id == id) return i;
  return -1;
}

/* b2Body::SetAwake */
void b2l_set_awake(b2l_body* b, int flag) {
  if (flag) {
    if (!b->awake) { b->awake = 1; b->sleepTime = 0.0f; }
  } else {
    b->awake = 0; b->sleepTime = 0.0f;
    b->v = V(0.0f, 0.0f); b->w = 0.0f;
  }
}

/* b2Body::ApplyLinearImpulse(impulse, worldCenter, wake) */
void b2l_apply_linear_impulse_center(b2l_world* w, int slot, float ix, float iy, int wake) {
  b2l_body* b = &w->bodies[slot];
  if (b->type != B2L_DYNAMIC) return;
  if (wake && !b->awake) b2l_set_awake(b, 1);
  if (b->awake) {
    b->v = vadd(b->v, vmul(b->invMass, V(ix, iy)));
    v2 point = b->c;
    b->w += b->invI * vcross(vsub(point, b->c), V(ix, iy));
  }
}
/* b2Body::ApplyAngularImpulse */
void b2l_apply_angular_impulse(b2l_world* w, int slot, float imp, int wake) {
  b2l_body* b = &w->bodies[slot];
  if (b->type != B2L_DYNAMIC) return;
  if (wake && !b->awake) b2l_set_awake(b, 1);
  if (b->awake) b->w += b->invI * imp;
}

/* slots sorted by descending seq == Box2D's LIFO body list */
static int body_order(b2l_world* w, int* out) {
  int n = 0;
  for (int i = 0; i < B2L_MAX_BODIES; ++i) if (w->bodies[i].used) out[n++] = i;
  for (int i = 1; i < n; ++i) {
    int k = out[i], j = i - 1;
    while (j >= 0 && w->bodies[out[j]].seq < w->bodies[k].seq) { out[j + 1] = out[j]; --j; }
    out[j + 1] = k;
  }
  return n;
}
/* contact indices sorted by descending seq == LIFO world contact list */
static int contact_order(b2l_world* w, int* out) {
  int n = 0;
  for (int i = 0; i < B2L_MAX_CONTACTS; ++i) if (w->contacts[i].used) out[n++] = i;
  for (int i = 1; i < n; ++i) {
    int k = out[i], j = i - 1;
    while (j >= 0 && w->contacts[out[j]].seq < w->contacts[k].seq) { out[j + 1] = out[j]; --j; }
    out[j + 1] = k;
  }
  return n;
}

b2l_contact* b2l_find_contact(b2l_world* w, int sa, int sb) {
  for (int i = 0; i < B2L_MAX_CONTACTS; ++i) {
    b2l_contact* c = &w->contacts[i];
    if (c->used && ((c->a == sa && c->b == sb) || (c->a == sb && c->b == sa))) return c;
  }
  return 0;
}

static b2l_contact* new_contact(b2l_world* w, int sa, int sb) {
  for (int i = 0; i < B2L_MAX_CONTACTS; ++i) {
    b2l_contact* c = &w->contacts[i];
    if (!c->used) {
      memset(c, 0, sizeof *c);
      c->used = 1;
      /* b2Contact::Create: polygon-vs-circle always has the polygon as A */
      b2l_body* A = &w->bodies[sa]; b2l_body* B = &w->bodies[sb];
      if (A->shape.type == B2L_CIRCLE && B->shape.type == B2L_POLYGON) { int t = sa; sa = sb; sb = t; }
      c->a = sa; c->b = sb;
      c->flags = B2L_ENABLED;
      c->toi = 1.0f;
      return c;
    }
  }
  assert(0 && "contact capacity");
  return 0;
}

b2l_contact* b2l_inject_contact(b2l_world* w, int sa, int sb, int seq, int flags, float ni, float ti) {
  /* sa must be the lower-seq body for circle-circle (AddPair: proxyA = min id) */
  b2l_contact* c = new_contact(w, sa, sb);
  c->seq = seq; c->flags = flags & (B2L_TOUCHING | B2L_ENABLED);
  c->pointCount = (flags & B2L_TOUCHING) ? 1 : 0;
  c->ni = ni; c->ti = ti;
  if (seq > w->contact_seq) w->contact_seq = seq;
  return c;
}

/* b2ContactManager::FindNewContacts + AddPair */
static void find_new_contacts(b2l_world* w) {
  typedef struct { int a, b; } pr; /* slots, ordered by (seqA < seqB) */
  static __thread pr pairs[B2L_MAX_BODIES * 16];
  int np = 0;
  for (int i = 0; i < B2L_MAX_BODIES; ++i) {
    b2l_body* bi = &w->bodies[i];
    if (!bi->used || !bi->moved) continue;
    for (int j = 0; j < B2L_MAX_BODIES; ++j) {
      b2l_body* bj = &w->bodies[j];
      if (!bj->used || j == i) continue;
      if (!fat_overlap(bi->fat, bj->fat)) continue;
      pr p; if (bi->seq < bj->seq) { p.a = i; p.b = j; } else { p.a = j; p.b = i; }
      assert(np < (int)(sizeof pairs / sizeof pairs[0]));
      pairs[np++] = p;
    }
  }
  for (int i = 0; i < B2L_MAX_BODIES; ++i) w->bodies[i].moved = 0;
  /* std::sort by (proxyIdA, proxyIdB); insertion sort is fine and stable */
  for (int i = 1; i < np; ++i) {
    pr k = pairs[i]; int j = i - 1;
    while (j >= 0) {
      int sa = w->bodies[pairs[j].a].seq, sb = w->bodies[pairs[j].b].seq;
      int ka = w->bodies[k.a].seq, kb = w->bodies[k.b].seq;
      if (sa > ka || (sa == ka && sb > kb)) { pairs[j + 1] = pairs[j]; --j; } else break;
    }
    pairs[j + 1] = k;
  }
  for (int i = 0; i < np; ++i) {
    if (i > 0 && pairs[i].a == pairs[i - 1].a && pairs[i].b == pairs[i - 1].b) continue;
    b2l_body* A = &w->bodies[pairs[i].a]; b2l_body* B = &w->bodies[pairs[i].b];
    if (b2l_find_contact(w, pairs[i].a, pairs[i].b)) continue;
    /* b2Body::ShouldCollide: at least one body dynamic (no joints here) */
    if (A->type != B2L_DYNAMIC && B->type != B2L_DYNAMIC) continue;
    /* deviation: sensor contacts are not materialised (see b2lite.h) */
    if (A->sensor || B->sensor) continue;
    b2l_contact* c = new_contact(w, pairs[i].a, pairs[i].b);
    c->seq = ++w->contact_seq;
  }
}

/* b2CollideCircles / b2CollidePolygonAndCircle -> manifold of contact c */
static void evaluate(b2l_world* w, b2l_contact* c) {
  b2l_body* A = &w->bodies[c->a]; b2l_body* B = &w->bodies[c->b];
  c->pointCount = 0;
  if (A->shape.type == B2L_CIRCLE) {
    v2 pA = vadd(qmul(A->qs, A->qc, V(0.0f, 0.0f)), A->p);
    v2 pB = vadd(qmul(B->qs, B->qc, V(0.0f, 0.0f)), B->p);
    v2 d = vsub(pB, pA);
    float distSqr = vdot(d, d);
    float rA = A->shape.radius, rB = B->shape.radius;
    float radius = rA + rB;
    if (distSqr > radius * radius) return;
    c->mtype = 0;
    c->localPoint = V(0.0f, 0.0f); c->localNormal = V(0.0f, 0.0f);
    c->pointCount = 1; c->mpLocal = V(0.0f, 0.0f);
    return;
  }
  /* polygon A, circle B */
  xform xfA = {A->p, A->qs, A->qc}; xform xfB = {B->p, B->qs, B->qc};
  v2 cW = xmul(xfB, V(0.0f, 0.0f));
  v2 cLocal = xmulT(xfA, cW);
  int normalIndex = 0;
  float separation = -b2_maxFloat;
  float radius = A->shape.radius + B->shape.radius;
  int vertexCount = A->shape.count;
  const v2* vertices = A->shape.verts; const v2* normals = A->shape.normals;
  for (int i = 0; i < vertexCount; ++i) {
    float s = vdot(normals[i], vsub(cLocal, vertices[i]));
    if (s > radius) return;
    if (s > separation) { separation = s; normalIndex = i; }
  }
  int vertIndex1 = normalIndex;
  int vertIndex2 = vertIndex1 + 1 < vertexCount ? vertIndex1 + 1 : 0;
  v2 v1 = vertices[vertIndex1], v2_ = vertices[vertIndex2];
  if (separation < b2_epsilon) {
    c->pointCount = 1; c->mtype = 1;
    c->localNormal = normals[normalIndex];
    c->localPoint = vmul(0.5f, vadd(v1, v2_));
    c->mpLocal = V(0.0f, 0.0f);
    return;
  }
  float u1 = vdot(vsub(cLocal, v1), vsub(v2_, v1));
  float u2 = vdot(vsub(cLocal, v2_), vsub(v1, v2_));
  if (u1 <= 0.0f) {
    if (vlen2(vsub(cLocal, v1)) > radius * radius) return;
    c->pointCount = 1; c->mtype = 1;
    c->localNormal = vsub(cLocal, v1); vnormalize(&c->localNormal);
    c->localPoint = v1; c->mpLocal = V(0.0f, 0.0f);
  } else if (u2 <= 0.0f) {
    if (vlen2(vsub(cLocal, v2_)) > radius * radius) return;
    c->pointCount = 1; c->mtype = 1;
    c->localNormal = vsub(cLocal, v2_); vnormalize(&c->localNormal);
    c->localPoint = v2_; c->mpLocal = V(0.0f, 0.0f);
  } else {
    v2 faceCenter = vmul(0.5f, vadd(v1, v2_));
    float sep = vdot(vsub(cLocal, faceCenter), normals[vertIndex1]);
    if (sep > radius) return;
    c->pointCount = 1; c->mtype = 1;
    c->localNormal = normals[vertIndex1];
    c->localPoint = faceCenter; c->mpLocal = V(0.0f, 0.0f);
  }
}

/* b2Contact::Update (non-sensor branch) */
static void contact_update(b2l_world* w, b2l_contact* c) {
  int oldCount = c->pointCount; float oni = c->ni, oti = c->ti;
  c->flags |= B2L_ENABLED;
  int wasTouching = (c->flags & B2L_TOUCHING) != 0;
  evaluate(w, c);
  int touching = c->pointCount > 0;
  /* single point, id.key == 0 on both manifolds: impulses carry over */
  c->ni = 0.0f; c->ti = 0.0f;
  if (touching && oldCount > 0) { c->ni = oni; c->ti = oti; }
  if (touching != wasTouching) {
    b2l_set_awake(&w->bodies[c->a], 1);
    b2l_set_awake(&w->bodies[c->b], 1);
  }
  if (touching) c->flags |= B2L_TOUCHING; else c->flags &= ~B2L_TOUCHING;
}

/* b2ContactManager::Collide */
static void collide(b2l_world* w) {
  int order[B2L_MAX_CONTACTS]; int n = contact_order(w, order);
  for (int k = 0; k < n; ++k) {
    b2l_contact* c = &w->contacts[order[k]];
    b2l_body* A = &w->bodies[c->a]; b2l_body* B = &w->bodies[c->b];
    int activeA = A->awake && A->type != B2L_STATIC;
    int activeB = B->awake && B->type != B2L_STATIC;
    if (!activeA && !activeB) continue;
    if (!fat_overlap(A->fat, B->fat)) { destroy_contact(w, c); continue; }
    contact_update(w, c);
  }
}

/* ------------------------------------------------------- contact solver --- */
typedef struct {
  int ci;          /* contact index */
  int ia, ib;      /* island body indices */
  v2 normal, rA, rB;
  float normalMass, tangentMass, velocityBias;
  float ni, ti;
  float invMassA, invMassB, invIA, invIB;
  float friction;
  /* position constraint */
  int mtype; v2 localNormal, localPoint, mpLocal; float radiusA, radiusB;
} cconstraint;

typedef struct { v2 c; float a; v2 v; float w; } bstate;

typedef struct {
  int nb, nc;
  int bslots[2 * b2_maxTOIContacts + B2L_MAX_BODIES];
  bstate st[2 * b2_maxTOIContacts + B2L_MAX_BODIES];
  int cidx[B2L_MAX_CONTACTS];
  cconstraint cc[B2L_MAX_CONTACTS];
} island_t;

static void solver_setup(b2l_world* w, island_t* is, int warmStarting, float dtRatio) {
  for (int i = 0; i < is->nc; ++i) {
    b2l_contact* c = &w->contacts[is->cidx[i]];
    cconstraint* cc = &is->cc[i];
    b2l_body* A = &w->bodies[c->a]; b2l_body* B = &w->bodies[c->b];
    cc->ci = is->cidx[i];
    cc->ia = A->islandIndex; cc->ib = B->islandIndex;
    cc->friction = sqrtf(B2L_FRICTION * B2L_FRICTION); /* b2MixFriction */
    cc->invMassA = A->invMass; cc->invMassB = B->invMass;
    cc->invIA = A->invI; cc->invIB = B->invI;
    cc->mtype = c->mtype; cc->localNormal = c->localNormal; cc->localPoint = c->localPoint;
    cc->mpLocal = c->mpLocal;
    cc->radiusA = A->shape.radius; cc->radiusB = B->shape.radius;
    if (warmStarting) { cc->ni = dtRatio * c->ni; cc->ti = dtRatio * c->ti; }
    else { cc->ni = 0.0f; cc->ti = 0.0f; }
  }
}

/* b2ContactSolver::InitializeVelocityConstraints (+ b2WorldManifold::Initialize) */
static void init_velocity_constraints(island_t* is) {
  for (int i = 0; i < is->nc; ++i) {
    cconstraint* cc = &is->cc[i];
    bstate* sA = &is->st[cc->ia]; bstate* sB = &is->st[cc->ib];
    float mA = cc->invMassA, mB = cc->invMassB, iA = cc->invIA, iB = cc->invIB;
    v2 cA = sA->c, cB = sB->c; float aA = sA->a, aB = sB->a;
    v2 vA = sA->v, vB = sB->v; float wA = sA->w, wB = sB->w;
    xform xfA, xfB;
    b2l_rot(aA, &xfA.s, &xfA.c); b2l_rot(aB, &xfB.s, &xfB.c);
    xfA.p = vsub(cA, qmul(xfA.s, xfA.c, V(0.0f, 0.0f)));
    xfB.p = vsub(cB, qmul(xfB.s, xfB.c, V(0.0f, 0.0f)));
    v2 normal, point;
    if (cc->mtype == 0) {
      normal = V(1.0f, 0.0f);
      v2 pointA = xmul(xfA, cc->localPoint);
      v2 pointB = xmul(xfB, cc->mpLocal);
      if (vlen2(vsub(pointA, pointB)) > b2_epsilon * b2_epsilon) {
        normal = vsub(pointB, pointA); vnormalize(&normal);
      }
      v2 pA = vadd(pointA, vmul(cc->radiusA, normal));
      v2 pB = vsub(pointB, vmul(cc->radiusB, normal));
      point = vmul(0.5f, vadd(pA, pB));
    } else {
      normal = qmul(xfA.s, xfA.c, cc->localNormal);
      v2 planePoint = xmul(xfA, cc->localPoint);
      v2 clipPoint = xmul(xfB, cc->mpLocal);
      v2 pA = vadd(clipPoint, vmul(cc->radiusA - vdot(vsub(clipPoint, planePoint), normal), normal));
      v2 pB = vsub(clipPoint, vmul(cc->radiusB, normal));
      point = vmul(0.5f, vadd(pA, pB));
    }
    cc->normal = normal;
    cc->rA = vsub(point, cA); cc->rB = vsub(point, cB);
    float rnA = vcross(cc->rA, normal), rnB = vcross(cc->rB, normal);
    float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
    cc->normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
    v2 tangent = vcross_vs(normal, 1.0f);
    float rtA = vcross(cc->rA, tangent), rtB = vcross(cc->rB, tangent);
    float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
    cc->tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
    cc->velocityBias = 0.0f;
    float vRel = vdot(normal, vsub(vsub(vadd(vB, vcross_sv(wB, cc->rB)), vA), vcross_sv(wA, cc->rA)));
    if (vRel < -b2_velocityThreshold) cc->velocityBias = -0.0f * vRel; /* restitution 0 */
  }
}

static void warm_start(island_t* is) {
  for (int i = 0; i < is->nc; ++i) {
    cconstraint* cc = &is->cc[i];
    bstate* sA = &is->st[cc->ia]; bstate* sB = &is->st[cc->ib];
    v2 tangent = vcross_vs(cc->normal, 1.0f);
    v2 P = vadd(vmul(cc->ni, cc->normal), vmul(cc->ti, tangent));
    sA->w -= cc->invIA * vcross(cc->rA, P);
    sA->v = vsub(sA->v, vmul(cc->invMassA, P));
    sB->w += cc->invIB * vcross(cc->rB, P);
    sB->v = vadd(sB->v, vmul(cc->invMassB, P));
  }
}

static void solve_velocity_constraints(island_t* is) {
  for (int i = 0; i < is->nc; ++i) {
    cconstraint* cc = &is->cc[i];
    bstate* sA = &is->st[cc->ia]; bstate* sB = &is->st[cc->ib];
    float mA = cc->invMassA, mB = cc->invMassB, iA = cc->invIA, iB = cc->invIB;
    v2 vA = sA->v, vB = sB->v; float wA = sA->w, wB = sB->w;
    v2 normal = cc->normal; v2 tangent = vcross_vs(normal, 1.0f);
    float friction = cc->friction;
    { /* tangent first */
      v2 dv = vsub(vsub(vadd(vB, vcross_sv(wB, cc->rB)), vA), vcross_sv(wA, cc->rA));
      float vt = vdot(dv, tangent) - 0.0f;
      float lambda = cc->tangentMass * (-vt);
      float maxFriction = friction * cc->ni;
      float newImpulse = fclamp(cc->ti + lambda, -maxFriction, maxFriction);
      lambda = newImpulse - cc->ti;
      cc->ti = newImpulse;
      v2 P = vmul(lambda, tangent);
      vA = vsub(vA, vmul(mA, P)); wA -= iA * vcross(cc->rA, P);
      vB = vadd(vB, vmul(mB, P)); wB += iB * vcross(cc->rB, P);
    }
    { /* normal, one point */
      v2 dv = vsub(vsub(vadd(vB, vcross_sv(wB, cc->rB)), vA), vcross_sv(wA, cc->rA));
      float vn = vdot(dv, normal);
      float lambda = -cc->normalMass * (vn - cc->velocityBias);
      float newImpulse = fmaxf_(cc->ni + lambda, 0.0f);
      lambda = newImpulse - cc->ni;
      cc->ni = newImpulse;
      v2 P = vmul(lambda, normal);
      vA = vsub(vA, vmul(mA, P)); wA -= iA * vcross(cc->rA, P);
      vB = vadd(vB, vmul(mB, P)); wB += iB * vcross(cc->rB, P);
    }
    sA->v = vA; sA->w = wA; sB->v = vB; sB->w = wB;
  }
}

/* b2ContactSolver::SolvePositionConstraints / SolveTOIPositionConstraints */
static int solve_position_constraints(island_t* is, int toi, int toiA, int toiB) {
  float minSeparation = 0.0f;
  for (int i = 0; i < is->nc; ++i) {
    cconstraint* cc = &is->cc[i];
    bstate* sA = &is->st[cc->ia]; bstate* sB = &is->st[cc->ib];
    float mA = cc->invMassA, iA = cc->invIA, mB = cc->invMassB, iB = cc->invIB;
    if (toi) {
      mA = 0.0f; iA = 0.0f; mB = 0.0f; iB = 0.0f;
      if (cc->ia == toiA || cc->ia == toiB) { mA = cc->invMassA; iA = cc->invIA; }
      if (cc->ib == toiA || cc->ib == toiB) { mB = cc->invMassB; iB = cc->invIB; }
    }
    v2 cA = sA->c, cB = sB->c; float aA = sA->a, aB = sB->a;
    xform xfA, xfB;
    b2l_rot(aA, &xfA.s, &xfA.c); b2l_rot(aB, &xfB.s, &xfB.c);
    xfA.p = vsub(cA, qmul(xfA.s, xfA.c, V(0.0f, 0.0f)));
    xfB.p = vsub(cB, qmul(xfB.s, xfB.c, V(0.0f, 0.0f)));
    v2 normal, point; float separation;
    if (cc->mtype == 0) {
      v2 pointA = xmul(xfA, cc->localPoint);
      v2 pointB = xmul(xfB, cc->mpLocal);
      normal = vsub(pointB, pointA); vnormalize(&normal);
      point = vmul(0.5f, vadd(pointA, pointB));
      separation = vdot(vsub(pointB, pointA), normal) - cc->radiusA - cc->radiusB;
    } else {
      normal = qmul(xfA.s, xfA.c, cc->localNormal);
      v2 planePoint = xmul(xfA, cc->localPoint);
      v2 clipPoint = xmul(xfB, cc->mpLocal);
      separation = vdot(vsub(clipPoint, planePoint), normal) - cc->radiusA - cc->radiusB;
      point = clipPoint;
    }
    v2 rA = vsub(point, cA), rB = vsub(point, cB);
    minSeparation = fminf_(minSeparation, separation);
    float C = fclamp((toi ? b2_toiBaugarte : b2_baumgarte) * (separation + b2_linearSlop),
                     -b2_maxLinearCorrection, 0.0f);
    float rnA = vcross(rA, normal), rnB = vcross(rB, normal);
    float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
    float impulse = K > 0.0f ? -C / K : 0.0f;
    v2 P = vmul(impulse, normal);
    cA = vsub(cA, vmul(mA, P)); aA -= iA * vcross(rA, P);
    cB = vadd(cB, vmul(mB, P)); aB += iB * vcross(rB, P);
    sA->c = cA; sA->a = aA; sB->c = cB; sB->a = aB;
  }
  return minSeparation >= (toi ? -1.5f : -3.0f) * b2_linearSlop;
}

static void island_add_body(b2l_world* w, island_t* is, int slot) {
  w->bodies[slot].islandIndex = is->nb;
  is->bslots[is->nb++] = slot;
}

/* b2Island::Solve */
static void island_solve(b2l_world* w, island_t* is, float h, float dtRatio, int velIters, int posIters) {
  for (int i = 0; i < is->nb; ++i) {
    b2l_body* b = &w->bodies[is->bslots[i]];
    v2 c = b->c; float a = b->a; v2 v = b->v; float wv = b->w;
    b->c0 = b->c; b->a0 = b->a;
    if (b->type == B2L_DYNAMIC) {
      /* gravity 0, no forces: v += h * (0 + invMass * 0) */
      v = vadd(v, vmul(h, vadd(vmul(1.0f, V(0.0f, 0.0f)), vmul(b->invMass, V(0.0f, 0.0f)))));
      wv += h * b->invI * 0.0f;
      if (g_b2_variant & 1) {   /* Box2D <= 2.2: v *= b2Clamp(1 - h * damping, 0, 1) */
        v = vmul(fclamp(1.0f - h * b->linDamp, 0.0f, 1.0f), v);
        wv *= fclamp(1.0f - h * b->angDamp, 0.0f, 1.0f);
      } else {                  /* 2.3.x: Pade approximation */
        v = vmul(1.0f / (1.0f + h * b->linDamp), v);
        wv *= 1.0f / (1.0f + h * b->angDamp);
      }
    }
    is->st[i].c = c; is->st[i].a = a; is->st[i].v = v; is->st[i].w = wv;
  }
  solver_setup(w, is, 1, dtRatio);
  init_velocity_constraints(is);
  warm_start(is);
  for (int i = 0; i < velIters; ++i) solve_velocity_constraints(is);
  for (int i = 0; i < is->nc; ++i) { /* StoreImpulses */
    b2l_contact* c = &w->contacts[is->cc[i].ci];
    c->ni = is->cc[i].ni; c->ti = is->cc[i].ti;
  }
  for (int i = 0; i < is->nb; ++i) {
    bstate* s = &is->st[i];
    v2 translation = vmul(h, s->v);
    if (vdot(translation, translation) > b2_maxTranslationSquared) {
      float ratio = b2_maxTranslation / vlen(translation);
      s->v = vmul(ratio, s->v);
    }
    float rotation = h * s->w;
    if (rotation * rotation > b2_maxRotationSquared) {
      float ratio = b2_maxRotation / fabsf(rotation);
      s->w *= ratio;
    }
    s->c = vadd(s->c, vmul(h, s->v));
    s->a += h * s->w;
  }
  int positionSolved = 0;
  for (int i = 0; i < posIters; ++i) {
    if (solve_position_constraints(is, 0, 0, 0)) { positionSolved = 1; break; }
  }
  for (int i = 0; i < is->nb; ++i) {
    b2l_body* b = &w->bodies[is->bslots[i]];
    b->c = is->st[i].c; b->a = is->st[i].a; b->v = is->st[i].v; b->w = is->st[i].w;
    b2l_sync_transform(b);
  }
  /* sleeping (allowSleep: doSleep=True, simulation.py:229) */
  float minSleepTime = b2_maxFloat;
  const float linTolSqr = b2_linearSleepTolerance * b2_linearSleepTolerance;
  const float angTolSqr = b2_angularSleepTolerance * b2_angularSleepTolerance;
  for (int i = 0; i < is->nb; ++i) {
    b2l_body* b = &w->bodies[is->bslots[i]];
    if (b->type == B2L_STATIC) continue;
    if (b->w * b->w > angTolSqr || vdot(b->v, b->v) > linTolSqr) {
      b->sleepTime = 0.0f; minSleepTime = 0.0f;
    } else {
      b->sleepTime += h;
      minSleepTime = fminf_(minSleepTime, b->sleepTime);
    }
  }
  if (minSleepTime >= b2_timeToSleep && positionSolved)
    for (int i = 0; i < is->nb; ++i) b2l_set_awake(&w->bodies[is->bslots[i]], 0);
}

/* b2Body::SynchronizeFixtures + b2DynamicTree::MoveProxy */
static void synchronize_fixtures(b2l_world* w, b2l_body* b) {
  (void)w;
  float s1, c1; b2l_rot(b->a0, &s1, &c1);
  v2 p1 = vsub(b->c0, qmul(s1, c1, V(0.0f, 0.0f)));
  float aabb1[4], aabb2[4], aabb[4];
  b2l_shape_aabb(&b->shape, p1, s1, c1, aabb1);
  b2l_shape_aabb(&b->shape, b->p, b->qs, b->qc, aabb2);
  aabb[0] = fminf_(aabb1[0], aabb2[0]); aabb[1] = fminf_(aabb1[1], aabb2[1]);
  aabb[2] = fmaxf_(aabb1[2], aabb2[2]); aabb[3] = fmaxf_(aabb1[3], aabb2[3]);
  v2 displacement = vsub(b->p, p1);
  if (b->fat[0] <= aabb[0] && b->fat[1] <= aabb[1] && aabb[2] <= b->fat[2] && aabb[3] <= b->fat[3])
    return;
  float f[4] = {aabb[0] - b2_aabbExtension, aabb[1] - b2_aabbExtension,
                aabb[2] + b2_aabbExtension, aabb[3] + b2_aabbExtension};
  v2 d = vmul(b2_aabbMultiplier, displacement);
  if (d.x < 0.0f) f[0] += d.x; else f[2] += d.x;
  if (d.y < 0.0f) f[1] += d.y; else f[3] += d.y;
  memcpy(b->fat, f, sizeof f);
  b->moved = 1;
}

/* b2World::Solve */
static void world_solve(b2l_world* w, float dt, float dtRatio, int velIters, int posIters) {
  static __thread island_t is;
  int border[B2L_MAX_BODIES]; int nbodies = body_order(w, border);
  int inAnyIsland[B2L_MAX_BODIES];
  memset(inAnyIsland, 0, sizeof inAnyIsland);
  for (int i = 0; i < nbodies; ++i) w->bodies[border[i]].islandFlag = 0;
  for (int i = 0; i < B2L_MAX_CONTACTS; ++i) w->contacts[i].flags &= ~B2L_ISLAND;
  int stack[B2L_MAX_BODIES];
  for (int si = 0; si < nbodies; ++si) {
    b2l_body* seed = &w->bodies[border[si]];
    if (seed->islandFlag) continue;
    if (!seed->awake) continue;
    if (seed->type == B2L_STATIC) continue;
    is.nb = 0; is.nc = 0;
    int stackCount = 0;
    stack[stackCount++] = border[si];
    seed->islandFlag = 1;
    while (stackCount > 0) {
      int bs = stack[--stackCount];
      b2l_body* b = &w->bodies[bs];
      island_add_body(w, &is, bs);
      inAnyIsland[bs] = 1;
      b2l_set_awake(b, 1);
      if (b->type == B2L_STATIC) continue;
      /* contact edges of b, newest first */
      int corder[B2L_MAX_CONTACTS]; int nco = contact_order(w, corder);
      for (int k = 0; k < nco; ++k) {
        b2l_contact* c = &w->contacts[corder[k]];
        if (c->a != bs && c->b != bs) continue;
        if (c->flags & B2L_ISLAND) continue;
        if (!(c->flags & B2L_ENABLED) || !(c->flags & B2L_TOUCHING)) continue;
        is.cidx[is.nc++] = corder[k];
        c->flags |= B2L_ISLAND;
        int other = c->a == bs ? c->b : c->a;
        if (w->bodies[other].islandFlag) continue;
        stack[stackCount++] = other;
        w->bodies[other].islandFlag = 1;
      }
    }
    island_solve(w, &is, dt, dtRatio, velIters, posIters);
    for (int i = 0; i < is.nb; ++i) {
      b2l_body* b = &w->bodies[is.bslots[i]];
      if (b->type == B2L_STATIC) b->islandFlag = 0;
    }
  }
  for (int i = 0; i < nbodies; ++i) {
    b2l_body* b = &w->bodies[border[i]];
    if (!b->islandFlag) continue;
    if (b->type == B2L_STATIC) continue;
    synchronize_fixtures(w, b);
  }
  find_new_contacts(w);
}

/* ------------------------------------------------------------ GJK / TOI --- */
typedef struct { const v2* verts; int count; float radius; } dproxy;
typedef struct { float metric; int count; int indexA[3], indexB[3]; } scache;
typedef struct { v2 wA, wB, w; float a; int indexA, indexB; } svertex;
typedef struct { svertex v[3]; int count; } simplex;
typedef struct { v2 localCenter, c0, c; float a0, a, alpha0; } sweep_t;

static const v2 kOrigin = {0.0f, 0.0f};

static int proxy_support(const dproxy* p, v2 d) {
  int best = 0; float bestValue = vdot(p->verts[0], d);
  for (int i = 1; i < p->count; ++i) {
    float value = vdot(p->verts[i], d);
    if (value > bestValue) { best = i; bestValue = value; }
  }
  return best;
}

static float simplex_metric(const simplex* s) {
  switch (s->count) {
    case 1: return 0.0f;
    case 2: return vlen(vsub(s->v[0].w, s->v[1].w));
    case 3: return vcross(vsub(s->v[1].w, s->v[0].w), vsub(s->v[2].w, s->v[0].w));
  }
  return 0.0f;
}

static void simplex_read_cache(simplex* s, const scache* cache, const dproxy* pA, xform xfA,
                               const dproxy* pB, xform xfB) {
  s->count = cache->count;
  for (int i = 0; i < s->count; ++i) {
    svertex* v = &s->v[i];
    v->indexA = cache->indexA[i]; v->indexB = cache->indexB[i];
    v->wA = xmul(xfA, pA->verts[v->indexA]);
    v->wB = xmul(xfB, pB->verts[v->indexB]);
    v->w = vsub(v->wB, v->wA);
    v->a = 0.0f;
  }
  if (s->count > 1) {
    float metric1 = cache->metric;
    float metric2 = simplex_metric(s);
    if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < b2_epsilon) s->count = 0;
  }
  if (s->count == 0) {
    svertex* v = &s->v[0];
    v->indexA = 0; v->indexB = 0;
    v->wA = xmul(xfA, pA->verts[0]);
    v->wB = xmul(xfB, pB->verts[0]);
    v->w = vsub(v->wB, v->wA);
    v->a = 1.0f;
    s->count = 1;
  }
}

static void simplex_solve2(simplex* s) {
  v2 w1 = s->v[0].w, w2 = s->v[1].w;
  v2 e12 = vsub(w2, w1);
  float d12_2 = -vdot(w1, e12);
  if (d12_2 <= 0.0f) { s->v[0].a = 1.0f; s->count = 1; return; }
  float d12_1 = vdot(w2, e12);
  if (d12_1 <= 0.0f) { s->v[1].a = 1.0f; s->count = 1; s->v[0] = s->v[1]; return; }
  float inv = 1.0f / (d12_1 + d12_2);
  s->v[0].a = d12_1 * inv; s->v[1].a = d12_2 * inv; s->count = 2;
}

static void simplex_solve3(simplex* s) {
  v2 w1 = s->v[0].w, w2 = s->v[1].w, w3 = s->v[2].w;
  v2 e12 = vsub(w2, w1);
  float w1e12 = vdot(w1, e12), w2e12 = vdot(w2, e12);
  float d12_1 = w2e12, d12_2 = -w1e12;
  v2 e13 = vsub(w3, w1);
  float w1e13 = vdot(w1, e13), w3e13 = vdot(w3, e13);
  float d13_1 = w3e13, d13_2 = -w1e13;
  v2 e23 = vsub(w3, w2);
  float w2e23 = vdot(w2, e23), w3e23 = vdot(w3, e23);
  float d23_1 = w3e23, d23_2 = -w2e23;
  float n123 = vcross(e12, e13);
  float d123_1 = n123 * vcross(w2, w3);
  float d123_2 = n123 * vcross(w3, w1);
  float d123_3 = n123 * vcross(w1, w2);
  if (d12_2 <= 0.0f && d13_2 <= 0.0f) { s->v[0].a = 1.0f; s->count = 1; return; }
  if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
    float inv = 1.0f / (d12_1 + d12_2);
    s->v[0].a = d12_1 * inv; s->v[1].a = d12_2 * inv; s->count = 2; return;
  }
  if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
    float inv = 1.0f / (d13_1 + d13_2);
    s->v[0].a = d13_1 * inv; s->v[2].a = d13_2 * inv; s->count = 2; s->v[1] = s->v[2]; return;
  }
  if (d12_1 <= 0.0f && d23_2 <= 0.0f) { s->v[1].a = 1.0f; s->count = 1; s->v[0] = s->v[1]; return; }
  if (d13_1 <= 0.0f && d23_1 <= 0.0f) { s->v[2].a = 1.0f; s->count = 1; s->v[0] = s->v[2]; return; }
  if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
    float inv = 1.0f / (d23_1 + d23_2);
    s->v[1].a = d23_1 * inv; s->v[2].a = d23_2 * inv; s->count = 2; s->v[0] = s->v[2]; return;
  }
  float inv = 1.0f / (d123_1 + d123_2 + d123_3);
  s->v[0].a = d123_1 * inv; s->v[1].a = d123_2 * inv; s->v[2].a = d123_3 * inv; s->count = 3;
}

/* b2Distance (useRadii = false) */
static float gjk_distance(scache* cache, const dproxy* pA, xform xfA, const dproxy* pB, xform xfB) {
  simplex s;
  simplex_read_cache(&s, cache, pA, xfA, pB, xfB);
  int saveA[3], saveB[3], saveCount = 0;
  int iter = 0;
  while (iter < 20) {
    saveCount = s.count;
    for (int i = 0; i < saveCount; ++i) { saveA[i] = s.v[i].indexA; saveB[i] = s.v[i].indexB; }
    switch (s.count) {
      case 1: break;
      case 2: simplex_solve2(&s); break;
      case 3: simplex_solve3(&s); break;
    }
    if (s.count == 3) break;
    /* search direction */
    v2 d;
    if (s.count == 1) d = vneg(s.v[0].w);
    else {
      v2 e12 = vsub(s.v[1].w, s.v[0].w);
      float sgn = vcross(e12, vneg(s.v[0].w));
      d = sgn > 0.0f ? vcross_sv(1.0f, e12) : vcross_vs(e12, 1.0f);
    }
    if (vlen2(d) < b2_epsilon * b2_epsilon) break;
    svertex* vx = &s.v[s.count];
    vx->indexA = proxy_support(pA, qmulT(xfA.s, xfA.c, vneg(d)));
    vx->wA = xmul(xfA, pA->verts[vx->indexA]);
    vx->indexB = proxy_support(pB, qmulT(xfB.s, xfB.c, d));
    vx->wB = xmul(xfB, pB->verts[vx->indexB]);
    vx->w = vsub(vx->wB, vx->wA);
    ++iter;
    int duplicate = 0;
    for (int i = 0; i < saveCount; ++i)
      if (vx->indexA == saveA[i] && vx->indexB == saveB[i]) { duplicate = 1; break; }
    if (duplicate) break;
    ++s.count;
  }
  v2 pointA, pointB;
  switch (s.count) {
    case 1: pointA = s.v[0].wA; pointB = s.v[0].wB; break;
    case 2:
      pointA = vadd(vmul(s.v[0].a, s.v[0].wA), vmul(s.v[1].a, s.v[1].wA));
      pointB = vadd(vmul(s.v[0].a, s.v[0].wB), vmul(s.v[1].a, s.v[1].wB));
      break;
    default:
      pointA = vadd(vadd(vmul(s.v[0].a, s.v[0].wA), vmul(s.v[1].a, s.v[1].wA)), vmul(s.v[2].a, s.v[2].wA));
      pointB = pointA; break;
  }
  float distance = vlen(vsub(pointA, pointB));
  cache->metric = simplex_metric(&s);
  cache->count = s.count;
  for (int i = 0; i < s.count; ++i) { cache->indexA[i] = s.v[i].indexA; cache->indexB[i] = s.v[i].indexB; }
  return distance;
}

/* b2Sweep::GetTransform */
static xform sweep_xf(const sweep_t* sw, float beta) {
  xform xf;
  xf.p = vadd(vmul(1.0f - beta, sw->c0), vmul(beta, sw->c));
  float angle = (1.0f - beta) * sw->a0 + beta * sw->a;
  b2l_rot(angle, &xf.s, &xf.c);
  xf.p = vsub(xf.p, qmul(xf.s, xf.c, sw->localCenter));
  return xf;
}
static void sweep_normalize(sweep_t* sw) {
  float twoPi = 2.0f * b2_pi;
  float d = twoPi * floorf(sw->a0 / twoPi);
  sw->a0 -= d; sw->a -= d;
}

typedef struct {
  const dproxy *pA, *pB; sweep_t swA, swB;
  int type; /* 0 points, 1 faceA, 2 faceB */
  v2 localPoint, axis;
} sepfn;

static float sep_init(sepfn* f, const scache* cache, const dproxy* pA, const sweep_t* swA,
                      const dproxy* pB, const sweep_t* swB, float t1) {
  f->pA = pA; f->pB = pB; f->swA = *swA; f->swB = *swB;
  int count = cache->count;
  xform xfA = sweep_xf(swA, t1), xfB = sweep_xf(swB, t1);
  if (count == 1) {
    f->type = 0;
    v2 lA = pA->verts[cache->indexA[0]], lB = pB->verts[cache->indexB[0]];
    v2 pointA = xmul(xfA, lA), pointB = xmul(xfB, lB);
    f->axis = vsub(pointB, pointA);
    return vnormalize(&f->axis);
  } else if (cache->indexA[0] == cache->indexA[1]) {
    f->type = 2;
    v2 lB1 = pB->verts[cache->indexB[0]], lB2 = pB->verts[cache->indexB[1]];
    f->axis = vcross_vs(vsub(lB2, lB1), 1.0f); vnormalize(&f->axis);
    v2 normal = qmul(xfB.s, xfB.c, f->axis);
    f->localPoint = vmul(0.5f, vadd(lB1, lB2));
    v2 pointB = xmul(xfB, f->localPoint);
    v2 lA = pA->verts[cache->indexA[0]];
    v2 pointA = xmul(xfA, lA);
    float s = vdot(vsub(pointA, pointB), normal);
    if (s < 0.0f) { f->axis = vneg(f->axis); s = -s; }
    return s;
  } else {
    f->type = 1;
    v2 lA1 = pA->verts[cache->indexA[0]], lA2 = pA->verts[cache->indexA[1]];
    f->axis = vcross_vs(vsub(lA2, lA1), 1.0f); vnormalize(&f->axis);
    v2 normal = qmul(xfA.s, xfA.c, f->axis);
    f->localPoint = vmul(0.5f, vadd(lA1, lA2));
    v2 pointA = xmul(xfA, f->localPoint);
    v2 lB = pB->verts[cache->indexB[0]];
    v2 pointB = xmul(xfB, lB);
    float s = vdot(vsub(pointB, pointA), normal);
    if (s < 0.0f) { f->axis = vneg(f->axis); s = -s; }
    return s;
  }
}

static float sep_find_min(const sepfn* f, int* indexA, int* indexB, float t) {
  xform xfA = sweep_xf(&f->swA, t), xfB = sweep_xf(&f->swB, t);
  switch (f->type) {
    case 0: {
      v2 axisA = qmulT(xfA.s, xfA.c, f->axis);
      v2 axisB = qmulT(xfB.s, xfB.c, vneg(f->axis));
      *indexA = proxy_support(f->pA, axisA);
      *indexB = proxy_support(f->pB, axisB);
      v2 pointA = xmul(xfA, f->pA->verts[*indexA]);
      v2 pointB = xmul(xfB, f->pB->verts[*indexB]);
      return vdot(vsub(pointB, pointA), f->axis);
    }
    case 1: {
      v2 normal = qmul(xfA.s, xfA.c, f->axis);
      v2 pointA = xmul(xfA, f->localPoint);
      v2 axisB = qmulT(xfB.s, xfB.c, vneg(normal));
      *indexA = -1;
      *indexB = proxy_support(f->pB, axisB);
      v2 pointB = xmul(xfB, f->pB->verts[*indexB]);
      return vdot(vsub(pointB, pointA), normal);
    }
    default: {
      v2 normal = qmul(xfB.s, xfB.c, f->axis);
      v2 pointB = xmul(xfB, f->localPoint);
      v2 axisA = qmulT(xfA.s, xfA.c, vneg(normal));
      *indexB = -1;
      *indexA = proxy_support(f->pA, axisA);
      v2 pointA = xmul(xfA, f->pA->verts[*indexA]);
      return vdot(vsub(pointA, pointB), normal);
    }
  }
}

static float sep_evaluate(const sepfn* f, int indexA, int indexB, float t) {
  xform xfA = sweep_xf(&f->swA, t), xfB = sweep_xf(&f->swB, t);
  switch (f->type) {
    case 0: {
      v2 pointA = xmul(xfA, f->pA->verts[indexA]);
      v2 pointB = xmul(xfB, f->pB->verts[indexB]);
      return vdot(vsub(pointB, pointA), f->axis);
    }
    case 1: {
      v2 normal = qmul(xfA.s, xfA.c, f->axis);
      v2 pointA = xmul(xfA, f->localPoint);
      v2 pointB = xmul(xfB, f->pB->verts[indexB]);
      return vdot(vsub(pointB, pointA), normal);
    }
    default: {
      v2 normal = qmul(xfB.s, xfB.c, f->axis);
      v2 pointB = xmul(xfB, f->localPoint);
      v2 pointA = xmul(xfA, f->pA->verts[indexA]);
      return vdot(vsub(pointA, pointB), normal);
    }
  }
}

enum { TOI_UNKNOWN, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED };

/* b2TimeOfImpact */
static int time_of_impact(const dproxy* pA, sweep_t swA, const dproxy* pB, sweep_t swB,
                          float tMax, float* tOut) {
  int state = TOI_UNKNOWN; *tOut = tMax;
  sweep_normalize(&swA); sweep_normalize(&swB);
  float totalRadius = pA->radius + pB->radius;
  float target = fmaxf_(b2_linearSlop, totalRadius - 3.0f * b2_linearSlop);
  float tolerance = 0.25f * b2_linearSlop;
  float t1 = 0.0f;
  int iter = 0;
  scache cache; cache.count = 0;
  for (;;) {
    xform xfA = sweep_xf(&swA, t1), xfB = sweep_xf(&swB, t1);
    float distance = gjk_distance(&cache, pA, xfA, pB, xfB);
    if (distance <= 0.0f) { state = TOI_OVERLAPPED; *tOut = 0.0f; break; }
    if (distance < target + tolerance) { state = TOI_TOUCHING; *tOut = t1; break; }
    sepfn fcn;
    sep_init(&fcn, &cache, pA, &swA, pB, &swB, t1);
    int done = 0;
    float t2 = tMax;
    int pushBackIter = 0;
    for (;;) {
      int indexA, indexB;
      float s2 = sep_find_min(&fcn, &indexA, &indexB, t2);
      if (s2 > target + tolerance) { state = TOI_SEPARATED; *tOut = tMax; done = 1; break; }
      if (s2 > target - tolerance) { t1 = t2; break; }
      float s1 = sep_evaluate(&fcn, indexA, indexB, t1);
      if (s1 < target - tolerance) { state = TOI_FAILED; *tOut = t1; done = 1; break; }
      if (s1 <= target + tolerance) { state = TOI_TOUCHING; *tOut = t1; done = 1; break; }
      int rootIterCount = 0;
      float a1 = t1, a2 = t2;
      for (;;) {
        float t;
        if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
        else t = 0.5f * (a1 + a2);
        ++rootIterCount;
        float s = sep_evaluate(&fcn, indexA, indexB, t);
        if (fabsf(s - target) < tolerance) { t2 = t; break; }
        if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
        if (rootIterCount == 50) break;
      }
      ++pushBackIter;
      if (pushBackIter == b2_maxPolygonVertices) break;
    }
    ++iter;
    if (done) break;
    if (iter == 20) { state = TOI_FAILED; *tOut = t1; break; }
  }
  return state;
}

/* b2Sweep::Advance / b2Body::Advance */
static void body_advance(b2l_body* b, float alpha) {
  float beta = (alpha - b->alpha0) / (1.0f - b->alpha0);
  b->c0 = vadd(b->c0, vmul(beta, vsub(b->c, b->c0)));
  b->a0 += beta * (b->a - b->a0);
  b->alpha0 = alpha;
  b->c = b->c0; b->a = b->a0;
  b2l_sync_transform(b);
}
static void sweep_advance(b2l_body* b, float alpha) {
  float beta = (alpha - b->alpha0) / (1.0f - b->alpha0);
  b->c0 = vadd(b->c0, vmul(beta, vsub(b->c, b->c0)));
  b->a0 += beta * (b->a - b->a0);
  b->alpha0 = alpha;
}
typedef struct { v2 c0, c; float a0, a, alpha0; } sweep_backup;
static sweep_backup sw_save(const b2l_body* b) { sweep_backup s = {b->c0, b->c, b->a0, b->a, b->alpha0}; return s; }
static void sw_restore(b2l_body* b, sweep_backup s) { b->c0 = s.c0; b->c = s.c; b->a0 = s.a0; b->a = s.a; b->alpha0 = s.alpha0; }

/* b2Island::SolveTOI */
static void island_solve_toi(b2l_world* w, island_t* is, float h, int velIters, int toiA, int toiB) {
  for (int i = 0; i < is->nb; ++i) {
    b2l_body* b = &w->bodies[is->bslots[i]];
    is->st[i].c = b->c; is->st[i].a = b->a; is->st[i].v = b->v; is->st[i].w = b->w;
  }
  solver_setup(w, is, 0, 1.0f);
  for (int i = 0; i < 20; ++i)
    if (solve_position_constraints(is, 1, toiA, toiB)) break;
  w->bodies[is->bslots[toiA]].c0 = is->st[toiA].c; w->bodies[is->bslots[toiA]].a0 = is->st[toiA].a;
  w->bodies[is->bslots[toiB]].c0 = is->st[toiB].c; w->bodies[is->bslots[toiB]].a0 = is->st[toiB].a;
  init_velocity_constraints(is);
  for (int i = 0; i < velIters; ++i) solve_velocity_constraints(is);
  for (int i = 0; i < is->nb; ++i) {
    bstate* s = &is->st[i];
    v2 translation = vmul(h, s->v);
    if (vdot(translation, translation) > b2_maxTranslationSquared) {
      float ratio = b2_maxTranslation / vlen(translation);
      s->v = vmul(ratio, s->v);
    }
    float rotation = h * s->w;
    if (rotation * rotation > b2_maxRotationSquared) {
      float ratio = b2_maxRotation / fabsf(rotation);
      s->w *= ratio;
    }
    s->c = vadd(s->c, vmul(h, s->v));
    s->a += h * s->w;
    b2l_body* b = &w->bodies[is->bslots[i]];
    b->c = s->c; b->a = s->a; b->v = s->v; b->w = s->w;
    b2l_sync_transform(b);
  }
}

/* b2World::SolveTOI (m_stepComplete is always true: no sub-stepping) */
static void world_solve_toi(b2l_world* w, float dt, int velIters) {
  static __thread island_t is;
  for (int i = 0; i < B2L_MAX_BODIES; ++i)
    if (w->bodies[i].used) { w->bodies[i].islandFlag = 0; w->bodies[i].alpha0 = 0.0f; }
  for (int i = 0; i < B2L_MAX_CONTACTS; ++i)
    if (w->contacts[i].used) {
      w->contacts[i].flags &= ~(B2L_TOI | B2L_ISLAND);
      w->contacts[i].toiCount = 0; w->contacts[i].toi = 1.0f;
    }
  for (;;) {
    b2l_contact* minContact = 0; float minAlpha = 1.0f;
    int corder[B2L_MAX_CONTACTS]; int nco = contact_order(w, corder);
    for (int k = 0; k < nco; ++k) {
      b2l_contact* c = &w->contacts[corder[k]];
      if (!(c->flags & B2L_ENABLED)) continue;
      if ((g_b2_variant & 4) ? c->toiCount >= b2_maxSubSteps : c->toiCount > b2_maxSubSteps) continue;
      float alpha = 1.0f;
      if (c->flags & B2L_TOI) alpha = c->toi;
      else {
        b2l_body* bA = &w->bodies[c->a]; b2l_body* bB = &w->bodies[c->b];
        if (bA->sensor || bB->sensor) continue;
        int activeA = bA->awake && bA->type != B2L_STATIC;
        int activeB = bB->awake && bB->type != B2L_STATIC;
        if (!activeA && !activeB) continue;
        int collideA = bA->type != B2L_DYNAMIC; /* no bullets */
        int collideB = bB->type != B2L_DYNAMIC;
        if (!collideA && !collideB) continue;
        float alpha0 = bA->alpha0;
        if (bA->alpha0 < bB->alpha0) { alpha0 = bB->alpha0; sweep_advance(bA, alpha0); }
        else if (bB->alpha0 < bA->alpha0) { alpha0 = bA->alpha0; sweep_advance(bB, alpha0); }
        dproxy pA = {bA->shape.type == B2L_CIRCLE ? &kOrigin : bA->shape.verts, bA->shape.count, bA->shape.radius};
        dproxy pB = {bB->shape.type == B2L_CIRCLE ? &kOrigin : bB->shape.verts, bB->shape.count, bB->shape.radius};
        sweep_t sA = {V(0.0f, 0.0f), bA->c0, bA->c, bA->a0, bA->a, bA->alpha0};
        sweep_t sB = {V(0.0f, 0.0f), bB->c0, bB->c, bB->a0, bB->a, bB->alpha0};
        float beta; int state = time_of_impact(&pA, sA, &pB, sB, 1.0f, &beta);
        if (state == TOI_TOUCHING) alpha = fminf_(alpha0 + (1.0f - alpha0) * beta, 1.0f);
        else alpha = 1.0f;
        c->toi = alpha; c->flags |= B2L_TOI;
      }
      if (alpha < minAlpha) { minContact = c; minAlpha = alpha; }
    }
    if (minContact == 0 || 1.0f - 10.0f * b2_epsilon < minAlpha) break;
    b2l_body* bA = &w->bodies[minContact->a]; b2l_body* bB = &w->bodies[minContact->b];
    sweep_backup backup1 = sw_save(bA), backup2 = sw_save(bB);
    body_advance(bA, minAlpha); body_advance(bB, minAlpha);
    contact_update(w, minContact);
    minContact->flags &= ~B2L_TOI;
    ++minContact->toiCount;
    if (!(minContact->flags & B2L_ENABLED) || !(minContact->flags & B2L_TOUCHING)) {
      minContact->flags &= ~B2L_ENABLED;
      sw_restore(bA, backup1); sw_restore(bB, backup2);
      b2l_sync_transform(bA); b2l_sync_transform(bB);
      continue;
    }
    ++w->n_toi_events;
    b2l_set_awake(bA, 1); b2l_set_awake(bB, 1);
    is.nb = 0; is.nc = 0;
    island_add_body(w, &is, minContact->a);
    island_add_body(w, &is, minContact->b);
    is.cidx[is.nc++] = (int)(minContact - w->contacts);
    bA->islandFlag = 1; bB->islandFlag = 1; minContact->flags |= B2L_ISLAND;
    int pair[2] = {minContact->a, minContact->b};
    for (int i = 0; i < 2; ++i) {
      int bs = pair[i]; b2l_body* body = &w->bodies[bs];
      if (body->type != B2L_DYNAMIC) continue;
      int co[B2L_MAX_CONTACTS]; int n2 = contact_order(w, co);
      for (int k = 0; k < n2; ++k) {
        b2l_contact* contact = &w->contacts[co[k]];
        if (contact->a != bs && contact->b != bs) continue;
        if (is.nb == 2 * b2_maxTOIContacts) break;
        if (is.nc == b2_maxTOIContacts) break;
        if (contact->flags & B2L_ISLAND) continue;
        int os = contact->a == bs ? contact->b : contact->a;
        b2l_body* other = &w->bodies[os];
        if (other->type == B2L_DYNAMIC) continue; /* no bullets */
        if (w->bodies[contact->a].sensor || w->bodies[contact->b].sensor) continue;
        sweep_backup backup = sw_save(other);
        if (!other->islandFlag) body_advance(other, minAlpha);
        contact_update(w, contact);
        if (!(contact->flags & B2L_ENABLED) || !(contact->flags & B2L_TOUCHING)) {
          sw_restore(other, backup); b2l_sync_transform(other); continue;
        }
        contact->flags |= B2L_ISLAND;
        is.cidx[is.nc++] = co[k];
        if (other->islandFlag) continue;
        other->islandFlag = 1;
        if (other->type != B2L_STATIC) b2l_set_awake(other, 1);
        island_add_body(w, &is, os);
      }
    }
    float subdt = (1.0f - minAlpha) * dt;
    island_solve_toi(w, &is, subdt, velIters, bA->islandIndex, bB->islandIndex);
    for (int i = 0; i < is.nb; ++i) {
      b2l_body* body = &w->bodies[is.bslots[i]];
      body->islandFlag = 0;
      if (body->type != B2L_DYNAMIC) continue;
      synchronize_fixtures(w, body);
      for (int k = 0; k < B2L_MAX_CONTACTS; ++k) {
        b2l_contact* c = &w->contacts[k];
        if (c->used && (c->a == is.bslots[i] || c->b == is.bslots[i])) c->flags &= ~(B2L_TOI | B2L_ISLAND);
      }
    }
    find_new_contacts(w);
  }
}

/* b2World::Step */
void b2l_step(b2l_world* w, float dt, int velIters, int posIters) {
  if (w->newFixture) { find_new_contacts(w); w->newFixture = 0; }
  float inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
  float dtRatio = w->inv_dt0 * dt;
  collide(w);
  if (dt > 0.0f) world_solve(w, dt, dtRatio, velIters, posIters);
  if (dt > 0.0f) world_solve_toi(w, dt, velIters);
  if (dt > 0.0f) w->inv_dt0 = inv_dt;
  /* autoClearForces: no forces are ever applied */
}

/* b2World::RayCast with the reference's closest-hit callback
 * (simulation.py:471-484): see deviation note in b2lite.h */
int b2l_raycast(b2l_world* w, v2 p1, v2 p2, float* fraction, v2* normal) {
  int order[B2L_MAX_BODIES]; int n = body_order(w, order);
  int best = -1; float bestF = 0.0f; v2 bestN = V(0.0f, 0.0f);
  for (int k = n - 1; k >= 0; --k) { /* ascending creation order */
    b2l_body* b = &w->bodies[order[k]];
    float f; v2 nn;
    if (b2l_shape_raycast(&b->shape, b->p, b->qs, b->qc, p1, p2, 1.0f, &f, &nn)) {
      if (best < 0 || f < bestF) { best = order[k]; bestF = f; bestN = nn; }
    }
  }
  if (best >= 0) { *fraction = bestF; if (normal) *normal = bestN; }
  return best;
}

int b2l_query_aabb(b2l_world* w, const float aabb[4], int* out, int cap) {
  int order[B2L_MAX_BODIES]; int n = body_order(w, order);
  int cnt = 0;
  for (int k = n - 1; k >= 0; --k) {
    b2l_body* b = &w->bodies[order[k]];
    if (fat_overlap(aabb, b->fat) && cnt < cap) out[cnt++] = order[k];
  }
  return cnt;
}
