/*
 * b2lite.h -- TEST ORACLE, not product code.
 *
 * CPU restatement (plain C, float32, no FMA contraction) of the subset of
 * Box2D v2.3.x that the masurvival step path exercises through pybox2d
 * 2.3.10.  pybox2d / Box2D are third-party
 * dependencies of the reference that are NOT vendored under /root/reference
 * (named only in README.md:16-18,53-57), so the algorithm below is restated
 * from the published Box2D 2.3 sources: b2World::Step/Solve/SolveTOI,
 * b2Island::Solve/SolveTOI, b2ContactSolver, b2CollideCircles,
 * b2CollidePolygonAndCircle, b2Distance (GJK), b2TimeOfImpact,
 * b2ContactManager::Collide/FindNewContacts, b2PolygonShape::Set/RayCast/
 * TestPoint, b2CircleShape::RayCast/TestPoint.
 *
 * PARITY STATUS: "parity unpinned" at this layer -- the reference ships no
 * golden vectors and pybox2d cannot be installed here; this file is pinned
 * only by closed-form known-answer tests (tests/test_b2lite_kat.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use anything under oracle/.
 *
 * Documented deviations from real Box2D (all measure-zero or unobservable):
 *  - broad-phase is brute force over fat AABBs; the dynamic-tree node id that
 *    orders new pairs is replaced by the body creation sequence number;
 *  - contacts involving a sensor fixture are never created (they carry no
 *    constraint and wake nobody);
 *  - ray casts return the minimum fraction over independent per-fixture
 *    tests (ties -> first body in creation order) instead of the tree's
 *    sequential clipping order.
 */
#ifndef B2LITE_H
#define B2LITE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B2L_MAX_BODIES 64
#define B2L_MAX_CONTACTS 192
#define B2L_MAX_VERTS 8

#define B2L_STATIC 0
#define B2L_DYNAMIC 2
#define B2L_CIRCLE 0
#define B2L_POLYGON 1

#define B2L_TOUCHING 1
#define B2L_ENABLED 2
#define B2L_ISLAND 4
#define B2L_TOI 8

typedef struct { float x, y; } b2l_vec2;

typedef struct b2l_shape {
  int type;      /* B2L_CIRCLE / B2L_POLYGON */
  float radius;  /* circle radius, or b2_polygonRadius */
  int count;
  b2l_vec2 verts[B2L_MAX_VERTS], normals[B2L_MAX_VERTS];
} b2l_shape;

typedef struct b2l_body {
  int used;
  int id;        /* unique handle (never reused) */
  int seq;       /* creation sequence: body-list order and proxy-id surrogate */
  int type;
  int sensor;
  b2l_shape shape;
  /* transform */
  b2l_vec2 p; float qs, qc;
  /* sweep */
  b2l_vec2 c0, c; float a0, a, alpha0;
  b2l_vec2 v; float w;
  float mass, invMass, I, invI;
  float linDamp, angDamp;
  float sleepTime; int awake;
  int islandFlag, islandIndex;
  float fat[4];
  int moved;
} b2l_body;

typedef struct b2l_contact {
  int used;
  int seq;
  int a, b;       /* body slots; a = fixtureA's body */
  int flags;
  int toiCount; float toi;
  /* manifold */
  int mtype;      /* 0 circles, 1 faceA */
  int pointCount;
  b2l_vec2 localNormal, localPoint, mpLocal;
  float ni, ti;
} b2l_contact;

typedef struct b2l_world {
  b2l_body bodies[B2L_MAX_BODIES];
  b2l_contact contacts[B2L_MAX_CONTACTS];
  int next_id, body_seq, contact_seq;
  int newFixture;
  float inv_dt0;
  int n_toi_events; /* diagnostics */
} b2l_world;

/* shapes */
void b2l_circle(b2l_shape* s, float radius);
void b2l_set_as_box(b2l_shape* s, float hx, float hy);
int  b2l_polygon_set(b2l_shape* s, const b2l_vec2* verts, int n);
int  b2l_test_point(const b2l_shape* s, b2l_vec2 p, float qs, float qc, b2l_vec2 pt);
void b2l_shape_aabb(const b2l_shape* s, b2l_vec2 p, float qs, float qc, float out[4]);
int  b2l_shape_raycast(const b2l_shape* s, b2l_vec2 p, float qs, float qc,
                       b2l_vec2 p1, b2l_vec2 p2, float maxFraction,
                       float* fraction, b2l_vec2* normal);
void b2l_rot(float angle, float* s, float* c); /* b2Rot::Set */

/* Box2D build variant (masurv.h MSV_B2_* bits), process-global */
void b2l_set_variant(int v);
int  b2l_get_variant(void);

/* world */
b2l_world* b2l_world_new(void);
void b2l_world_free(b2l_world* w);
void b2l_world_clear(b2l_world* w);
int  b2l_create_body(b2l_world* w, int type, float x, float y, float angle,
                     const b2l_shape* shape, float density, int sensor,
                     float linDamp, float angDamp);
void b2l_destroy_body(b2l_world* w, int slot);
int  b2l_slot_of(b2l_world* w, int id);
void b2l_step(b2l_world* w, float dt, int velIters, int posIters);
void b2l_apply_linear_impulse_center(b2l_world* w, int slot, float ix, float iy, int wake);
void b2l_apply_angular_impulse(b2l_world* w, int slot, float imp, int wake);
void b2l_set_awake(b2l_body* b, int flag);
/* closest hit; returns slot or -1 */
int  b2l_raycast(b2l_world* w, b2l_vec2 p1, b2l_vec2 p2, float* fraction,
                 b2l_vec2* normal);
/* bodies whose fat AABB overlaps [lo,hi], in creation order; returns count */
int  b2l_query_aabb(b2l_world* w, const float aabb[4], int* out, int cap);
/* injection helpers for the parity harness */
b2l_contact* b2l_find_contact(b2l_world* w, int slotA, int slotB);
b2l_contact* b2l_inject_contact(b2l_world* w, int slotA, int slotB, int seq,
                                int flags, float ni, float ti);
void b2l_sync_transform(b2l_body* b);

#ifdef __cplusplus
}
#endif
#endif
