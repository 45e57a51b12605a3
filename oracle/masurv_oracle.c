/*
 * masurv_oracle.c -- TEST ORACLE (see masurv_oracle.h).  Compile with
 * -O2 -ffp-contract=off.  Citations are into /root/reference/:
 *   env = masurvival/envs/masurvival_env.py, sim = masurvival/simulation.py,
 *   sem = masurvival/semantics.py.
 */
#include "masurv_oracle.h"
#include "b2lite.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>

typedef b2l_vec2 v2;
static inline v2 V(float x, float y) { v2 r = {x, y}; return r; }

#define MA MSV_MAX_AGENTS
#define MB MSV_MAX_BOXES
#define MH MSV_MAX_HEALS
#define MS MSV_MAX_SLOTS

typedef struct { int kind; msv_box_shape shape; int owner; } inv_item;

struct oracle_env {
  msv_config cfg;
  uint64_t seed; int64_t env_id;
  orc_draws draws; int have_draws;
  int death_cursor;
  b2l_world* w;
  /* derived float32 constants */
  float agent_r, heal_r, item_r, box_h, wall_hx, wall_hy, wall_off;
  b2l_shape cone;
  /* agents */
  int a_slot[MA];  /* b2lite slot, -1 when dead */
  int a_id[MA];    /* b2lite unique id (valid while alive, kept after) */
  int health[MA], cause[MA], cooldown[MA];
  int inv_n[MA]; inv_item inv[MA][MS];
  /* boxes */
  int n_boxes; int box_slot[MB]; msv_box_shape box_shape[MB];
  int box_health[MB], box_has_health[MB], box_cause[MB], box_owner[MB];
  /* box items */
  int n_items; int item_slot[MB]; msv_box_shape item_shape[MB]; int item_owner[MB];
  /* heals */
  int n_heals; int heal_slot[MH];
  int wall_slot[MSV_N_WALLS];
  /* Object.next_spawns */
  int n_pending; float pend_x[MB], pend_y[MB]; msv_box_shape pend_shape[MB]; int pend_owner[MB];
  /* zone */
  int n_zones; float zone_cx[MSV_MAX_ZONES], zone_cy[MSV_MAX_ZONES];
  int zone_phase, zone_t_cooldown, zone_t_shrink, zone_endgame;
  float zone_x, zone_y, zone_r;
  /* cameras (transient, but survives until the observation is built) */
  int n_seen_rows; int seen_n[MA]; int seen[MA][MA + 2 * MB + MH + 8];
  /* lidar (transient) */
  float lidar_frac[MA][MSV_MAX_LASERS]; int lidar_hit[MA][MSV_MAX_LASERS];
  /* per-step trackers */
  int n_deaths; int deaths[MA];
  int n_kills; int kill_cause[MA];
  int use_heal, use_box;
  /* counters */
  int steps, episode;
  float stat_reward[MA]; int stat_kills[MA]; int stat_steps, stat_heals, stat_boxes;
  int64_t stat_episodes;
  float last_rewards[MA]; int last_kills[MA];
  float ep_return[MA];   /* running return of the current episode (per-env view of env:483-508) */
  int64_t ev[ORC_EV_COUNT]; /* event-coverage census (orc_get_events) */
};

/* ------------------------------------------------------------------ Philox */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double orc_philox_uniform(uint64_t seed, uint32_t env, uint32_t episode, uint32_t step,
                          uint32_t stream, uint32_t k) {
  uint32_t ctr[4] = {env, episode, step, (stream << 16) | (k >> 1)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  orc_philox4x32(ctr, key, o);
  uint32_t a = o[(k & 1) * 2], b = o[(k & 1) * 2 + 1];
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

void orc_philox_actions(uint64_t seed, uint32_t env, uint32_t episode, uint32_t step,
                        int32_t n_agents, uint8_t* out) {
  for (int i = 0; i < n_agents; ++i) {
    uint32_t ctr[4] = {env, episode, step, (ORC_STREAM_ACTION << 16) | (uint32_t)i};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    orc_philox4x32(ctr, key, o);
    out[i * 6 + 0] = (uint8_t)((o[0] & 0xFFFF) % 3);
    out[i * 6 + 1] = (uint8_t)((o[0] >> 16) % 3);
    out[i * 6 + 2] = (uint8_t)((o[1] & 0xFFFF) % 3);
    out[i * 6 + 3] = (uint8_t)((o[1] >> 16) & 1);
    out[i * 6 + 4] = (uint8_t)((o[2] >> 16) & 1);
    out[i * 6 + 5] = (uint8_t)((o[3] >> 16) & 1);
  }
}

static double draw_uniform(oracle_env* e, int stream, int k, const double* buf) {
  if (buf) return buf[k];
  return orc_philox_uniform(e->seed, (uint32_t)e->env_id, (uint32_t)e->episode,
                            (uint32_t)e->steps, (uint32_t)stream, (uint32_t)k);
}
static double draw_normal(oracle_env* e, int k) {
  if (e->have_draws && e->draws.box_z) return e->draws.box_z[k];
  double u1 = draw_uniform(e, ORC_STREAM_BOX, 2 * k, 0);
  double u2 = draw_uniform(e, ORC_STREAM_BOX, 2 * k + 1, 0);
  return sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586 * u2);
}

/* --------------------------------------------------------------- helpers -- */
/* sim.from_polar (sim:20-23): R.angle = angle; R * b2Vec2(length, 0) */
static v2 from_polar(double length, double angle) {
  float s, c; b2l_rot((float)angle, &s, &c);
  float L = (float)length;
  return V(c * L + (-s) * 0.0f, s * L + c * 0.0f);
}

static void shape_of(const msv_box_shape* bs, b2l_shape* out) {
  b2l_set_as_box(out, bs->hx, bs->hy);
  if (bs->rehulled) { /* sim.copy_shape -> b2PolygonShape(vertices=...) */
    b2l_shape tmp = *out;
    b2l_polygon_set(out, tmp.verts, tmp.count);
  }
}

static int n_agents_alive(const oracle_env* e) {
  int n = 0; for (int i = 0; i < e->cfg.n_agents; ++i) n += e->a_slot[i] >= 0; return n;
}
static int team_of(const oracle_env* e, int idx) { /* sem:957-966 */
  return idx < e->cfg.n_agents / 2 ? 0 : 1;
}
static int team_alive(const oracle_env* e, int t) { /* sem:978-982 */
  int A = e->cfg.n_agents, split = A / 2;
  for (int i = (t ? split : 0); i < (t ? A : split); ++i) if (e->a_slot[i] >= 0) return 1;
  return 0;
}

/* classify a b2lite slot: kind + index */
static int classify(const oracle_env* e, int slot, int* idx) {
  for (int i = 0; i < e->cfg.n_agents; ++i) if (e->a_slot[i] == slot) { *idx = i; return ORC_KIND_AGENT; }
  for (int i = 0; i < e->n_boxes; ++i) if (e->box_slot[i] == slot) { *idx = i; return ORC_KIND_BOX; }
  for (int i = 0; i < e->n_items; ++i) if (e->item_slot[i] == slot) { *idx = i; return ORC_KIND_ITEM; }
  for (int i = 0; i < e->n_heals; ++i) if (e->heal_slot[i] == slot) { *idx = i; return ORC_KIND_HEAL; }
  for (int i = 0; i < MSV_N_WALLS; ++i) if (e->wall_slot[i] == slot) { *idx = i; return ORC_KIND_WALL; }
  *idx = -1; return 0;
}

/* sim.shape_query (sim:444-457) for a circle of radius r centred on `body`:
 * all bodies (creation order) whose worldCenter passes TestPoint. */
static int circle_query(oracle_env* e, int slot, float r, int* out) {
  b2l_body* b = &e->w->bodies[slot];
  b2l_shape circ; b2l_circle(&circ, r);
  float aabb[4]; b2l_shape_aabb(&circ, b->p, b->qs, b->qc, aabb);
  int cand[B2L_MAX_BODIES];
  int n = b2l_query_aabb(e->w, aabb, cand, B2L_MAX_BODIES), m = 0;
  for (int k = 0; k < n; ++k) {
    b2l_body* o = &e->w->bodies[cand[k]];
    if (b2l_test_point(&circ, b->p, b->qs, b->qc, o->c)) out[m++] = cand[k];
  }
  return m;
}

/* ---------------------------------------------------------- create/reset -- */
oracle_env* orc_create(const msv_config* cfg, uint64_t seed, int64_t env_id) {
  oracle_env* e = (oracle_env*)calloc(1, sizeof *e);
  e->cfg = *cfg; e->seed = seed; e->env_id = env_id;
  e->w = b2l_world_new();
  b2l_set_variant(cfg->b2_variant);
  e->agent_r = (float)(cfg->agent_size / 2);       /* sem:16-18 */
  e->heal_r = (float)(cfg->heal_item_size / 2);    /* sem:24-25 */
  e->item_r = (float)(cfg->box_item_size / 2);
  e->box_h = (float)(cfg->box_size / 2.);          /* sem:20-22, sim:56-57 */
  { /* sem:685-695 */
    double height = cfg->floor_size, width = height / 100;
    e->wall_hx = (float)(width / 2.); e->wall_hy = (float)(height / 2.);
    e->wall_off = (float)(cfg->floor_size / 2);
  }
  { /* sim:321-328 */
    v2 vs[4];
    vs[0] = V(0.0f, 0.0f);
    vs[1] = from_polar(cfg->cam_depth, +cfg->cam_fov / 2);
    vs[2] = V((float)cfg->cam_depth, 0.0f);
    vs[3] = from_polar(cfg->cam_depth, -cfg->cam_fov / 2);
    b2l_polygon_set(&e->cone, vs, 4);
  }
  e->n_zones = cfg->zone_n_radiuses + 1;
  for (int i = 0; i < MA; ++i) e->a_slot[i] = -1;
  e->episode = -1;
  return e;
}
void orc_destroy(oracle_env* e) { b2l_world_free(e->w); free(e); }
void orc_set_draws(oracle_env* e, const orc_draws* d) {
  if (d) { e->draws = *d; e->have_draws = 1; } else { memset(&e->draws, 0, sizeof e->draws); e->have_draws = 0; }
  e->death_cursor = 0;
}

static double zone_radius(const oracle_env* e, int i) {
  return i < e->cfg.zone_n_radiuses ? e->cfg.zone_radiuses[i] : 0.0; /* sem:726-727 */
}

static int spawn_agent(oracle_env* e, float x, float y) {
  b2l_shape s; b2l_circle(&s, e->agent_r);
  return b2l_create_body(e->w, B2L_DYNAMIC, x, y, 0.0f, &s, 1.0f, 0, 0.8f, 0.8f);
}
static int spawn_box_body(oracle_env* e, float x, float y, const msv_box_shape* bs) {
  b2l_shape s; shape_of(bs, &s);
  return b2l_create_body(e->w, B2L_STATIC, x, y, 0.0f, &s, 1.0f, 0, 0.8f, 0.8f);
}
static int spawn_sensor(oracle_env* e, float x, float y, float r) {
  b2l_shape s; b2l_circle(&s, r);
  return b2l_create_body(e->w, B2L_DYNAMIC, x, y, 0.0f, &s, 1.0f, 1, 0.8f, 0.8f);
}

static void cameras_update(oracle_env* e);
static void lidar_update(oracle_env* e);

/* BaseEnv.reset (env:59-74) */
void orc_reset(oracle_env* e, orc_out* out) {
  const msv_config* c = &e->cfg;
  b2l_set_variant(c->b2_variant);
  e->episode += 1; e->steps = 0;
  memset(e->ep_return, 0, sizeof e->ep_return);
  b2l_world_clear(e->w);
  /* SpawnGrid.reset (sem:71-74) + square_grid (sem:987-992) */
  int g = c->grid_size, n = g * g;
  float px[64], py[64]; int perm[64];
  assert(n <= 64);
  for (int k = 0; k < n; ++k) {
    int i = k % g, j = k / g;
    double ci = (double)i / g + 0.5 / g, cj = (double)j / g + 0.5 / g;
    px[k] = (float)(c->floor_size * ci - c->floor_size / 2.);
    py[k] = (float)(c->floor_size * cj - c->floor_size / 2.);
    perm[k] = k;
  }
  for (int i = n - 1; i >= 1; --i) { /* Fisher-Yates with injected uniforms */
    double u = draw_uniform(e, ORC_STREAM_SHUFFLE, n - 1 - i, e->have_draws ? e->draws.shuffle_u : 0);
    int j = (int)(u * (i + 1));
    int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
  }
  int top = n; /* placements pop from the END (sem:76-79) */
  /* groups reset in dict order boxes, box_items, heals, walls, agents (env:382-388) */
  e->n_boxes = 0; e->n_items = 0; e->n_heals = 0; e->n_pending = 0;
  msv_box_shape protos[MB];
  for (int b = 0; b < c->n_boxes; ++b) { /* RandomizeBoxShapes.post_reset sem:107-120 */
    protos[b].rehulled = 0;
    if (c->box_randomized) {
      double w = c->box_avg_w + c->box_std_w * draw_normal(e, 2 * b);
      if (!(w > c->box_min_w)) w = c->box_min_w;
      double h = c->box_avg_h + c->box_std_h * draw_normal(e, 2 * b + 1);
      if (!(h > c->box_min_h)) h = c->box_min_h;
      protos[b].hx = (float)(w / 2.); protos[b].hy = (float)(h / 2.);
    } else { protos[b].hx = e->box_h; protos[b].hy = e->box_h; }
  }
  for (int b = 0; b < c->n_boxes; ++b) {
    int k = perm[--top];
    e->box_slot[b] = spawn_box_body(e, px[k], py[k], &protos[b]);
    e->box_shape[b] = protos[b];
    e->box_health[b] = c->box_health; e->box_has_health[b] = 1; /* sem:421-423 */
    e->box_cause[b] = MSV_CAUSE_NONE; e->box_owner[b] = MSV_CAUSE_NONE;
    e->n_boxes++;
  }
  for (int h = 0; h < c->n_heals; ++h) {
    int k = perm[--top];
    e->heal_slot[h] = spawn_sensor(e, px[k], py[k], e->heal_r);
    e->n_heals++;
  }
  { /* ThickRoomWalls (sem:680-698) */
    b2l_shape s; b2l_set_as_box(&s, e->wall_hx, e->wall_hy);
    float o = e->wall_off; float hp = (float)(M_PI / 2);
    e->wall_slot[0] = b2l_create_body(e->w, B2L_STATIC, -o, 0.0f, 0.0f, &s, 1.0f, 0, 0.8f, 0.8f);
    e->wall_slot[1] = b2l_create_body(e->w, B2L_STATIC, 0.0f, o, hp, &s, 1.0f, 0, 0.8f, 0.8f);
    e->wall_slot[2] = b2l_create_body(e->w, B2L_STATIC, o, 0.0f, 0.0f, &s, 1.0f, 0, 0.8f, 0.8f);
    e->wall_slot[3] = b2l_create_body(e->w, B2L_STATIC, 0.0f, -o, hp, &s, 1.0f, 0, 0.8f, 0.8f);
  }
  for (int i = 0; i < MA; ++i) { e->a_slot[i] = -1; e->inv_n[i] = 0; e->cooldown[i] = 0; }
  for (int i = 0; i < c->n_agents; ++i) {
    int k = perm[--top];
    e->a_slot[i] = spawn_agent(e, px[k], py[k]);
    e->a_id[i] = e->w->bodies[e->a_slot[i]].id;
    e->health[i] = c->health; e->cause[i] = MSV_CAUSE_NONE;
  }
  /* agents module post_reset order (env:320-342): ... Cameras ... SafeZone */
  cameras_update(e);
  lidar_update(e);
  { /* SafeZone.post_reset sem:739-756 */
    if (c->zone_centers_random) {
      int d = 0;
      for (int z = e->n_zones - 1; z >= 0; --z) {
        double r = zone_radius(e, z);
        double L = c->floor_size - 2 * r;
        double ux = draw_uniform(e, ORC_STREAM_ZONE, d, e->have_draws ? e->draws.zone_u : 0); d++;
        double uy = draw_uniform(e, ORC_STREAM_ZONE, d, e->have_draws ? e->draws.zone_u : 0); d++;
        e->zone_cx[z] = (float)((ux * L) - L / 2);
        e->zone_cy[z] = (float)((uy * L) - L / 2);
      }
    } else {
      for (int z = 0; z < e->n_zones; ++z) {
        if (z < c->zone_n_radiuses) { e->zone_cx[z] = (float)c->zone_centers[z][0]; e->zone_cy[z] = (float)c->zone_centers[z][1]; }
        else { e->zone_cx[z] = 0.0f; e->zone_cy[z] = 0.0f; }
      }
    }
    e->zone_t_cooldown = c->zone_cooldown; e->zone_t_shrink = 0;
    e->zone_phase = 0; e->zone_endgame = 0;
    e->zone_r = (float)zone_radius(e, 0);
    e->zone_x = e->zone_cx[0]; e->zone_y = e->zone_cy[0];
  }
  e->n_deaths = 0; e->n_kills = 0; e->use_heal = 0; e->use_box = 0;
  if (out) {
    orc_observe(e, out); memset(out->rewards, 0, sizeof out->rewards); out->done = 0;
    out->immune = c->immunity_cooldown >= 0;   /* ImmunityPhase.post_reset (sem:661-665) */
    out->br_over = 0;                          /* BattleRoyale.post_reset (sem:38-39) */
  }
}

/* -------------------------------------------------------------- sensors --- */
/* Cameras._update_seen (sim:336-354) */
static void cameras_update(oracle_env* e) {
  int row = 0;
  for (int i = 0; i < e->cfg.n_agents; ++i) {
    if (e->a_slot[i] < 0) continue;
    b2l_body* body = &e->w->bodies[e->a_slot[i]];
    float aabb[4]; b2l_shape_aabb(&e->cone, body->p, body->qs, body->qc, aabb);
    int cand[B2L_MAX_BODIES]; int n = b2l_query_aabb(e->w, aabb, cand, B2L_MAX_BODIES);
    e->seen_n[row] = 0;
    for (int k = 0; k < n; ++k) {
      b2l_body* other = &e->w->bodies[cand[k]];
      if (!b2l_test_point(&e->cone, body->p, body->qs, body->qc, other->c)) continue;
      if (cand[k] == e->a_slot[i]) continue;
      v2 d = V(other->p.x - body->p.x, other->p.y - body->p.y);
      float k1 = (float)(1 + 1e-6);
      v2 end = V(body->p.x + k1 * d.x, body->p.y + k1 * d.y);
      float f; int hit = b2l_raycast(e->w, body->p, end, &f, 0);
      if (hit == cand[k]) e->seen[row][e->seen_n[row]++] = other->id;
    }
    row++;
  }
  e->n_seen_rows = row;
}

/* Lidars._update (sim:377-392); extension output, see DESIGN.md */
static void lidar_update(oracle_env* e) {
  const msv_config* c = &e->cfg;
  int L = c->lidar_n;
  for (int i = 0; i < c->n_agents; ++i)
    for (int r = 0; r < L; ++r) { e->lidar_frac[i][r] = 1.0f; e->lidar_hit[i][r] = 0; }
  if (L <= 0) return;
  for (int i = 0; i < c->n_agents; ++i) {
    if (e->a_slot[i] < 0) continue;
    b2l_body* body = &e->w->bodies[e->a_slot[i]];
    for (int r = 0; r < L; ++r) {
      double angle = L > 1 ? r * (c->lidar_fov / (L - 1)) - c->lidar_fov / 2. : 0.0;
      angle += (double)body->a;
      v2 off = from_polar(c->lidar_depth, angle);
      v2 end = V(body->p.x + off.x, body->p.y + off.y);
      float f; int hit = b2l_raycast(e->w, body->p, end, &f, 0);
      if (hit >= 0) {
        int idx; int kind = classify(e, hit, &idx);
        e->lidar_frac[i][r] = f; e->lidar_hit[i][r] = (kind << 8) | idx;
      }
    }
  }
}

/* ------------------------------------------------------------ list ops ---- */
static void remove_box(oracle_env* e, int k) {
  b2l_destroy_body(e->w, e->box_slot[k]);
  for (int j = k; j + 1 < e->n_boxes; ++j) {
    e->box_slot[j] = e->box_slot[j + 1]; e->box_shape[j] = e->box_shape[j + 1];
    e->box_health[j] = e->box_health[j + 1]; e->box_has_health[j] = e->box_has_health[j + 1];
    e->box_cause[j] = e->box_cause[j + 1]; e->box_owner[j] = e->box_owner[j + 1];
  }
  e->n_boxes--;
}
static void remove_item(oracle_env* e, int k) {
  b2l_destroy_body(e->w, e->item_slot[k]);
  for (int j = k; j + 1 < e->n_items; ++j) {
    e->item_slot[j] = e->item_slot[j + 1]; e->item_shape[j] = e->item_shape[j + 1];
    e->item_owner[j] = e->item_owner[j + 1];
  }
  e->n_items--;
}
static void remove_heal(oracle_env* e, int k) {
  b2l_destroy_body(e->w, e->heal_slot[k]);
  for (int j = k; j + 1 < e->n_heals; ++j) e->heal_slot[j] = e->heal_slot[j + 1];
  e->n_heals--;
}
static void drop_box_item(oracle_env* e, float x, float y, const msv_box_shape* sh, int owner) {
  assert(e->n_items < MB);
  int k = e->n_items++;
  e->item_slot[k] = spawn_sensor(e, x, y, e->item_r); /* Item.drop sem:143-148 */
  e->item_shape[k] = *sh; e->item_shape[k].rehulled = 1; e->item_owner[k] = owner;
}
static void drop_heal(oracle_env* e, float x, float y) {
  assert(e->n_heals < MH);
  e->heal_slot[e->n_heals++] = spawn_sensor(e, x, y, e->heal_r);
}

/* Health._change_health (sem:490-500) for agents */
static void agent_change_health(oracle_env* e, int idx, int delta, int cause) {
  if (e->a_slot[idx] < 0) return;
  if (e->cfg.teams && cause == MSV_CAUSE_TEAM0 + team_of(e, idx)) return; /* immunities sem:942-946 */
  e->health[idx] += delta; e->cause[idx] = cause;
}
static void box_change_health(oracle_env* e, int k, int delta, int cause) {
  if (!e->box_has_health[k]) return;                         /* Q9 sem:491-492 */
  if (e->box_owner[k] != MSV_CAUSE_NONE && cause != e->box_owner[k]) return; /* sem:497-498 */
  e->box_health[k] += delta; e->box_cause[k] = cause;
}

/* ----------------------------------------------------------------- step --- */
void orc_step(oracle_env* e, const uint8_t* actions, orc_out* out) {
  const msv_config* c = &e->cfg;
  b2l_set_variant(c->b2_variant);
  const int A = c->n_agents;
  b2l_world* w = e->w;
  static const double dtab[3] = {-1., 0., 1.};
  e->n_deaths = 0; e->n_kills = 0; e->use_heal = 0; e->use_box = 0;
  e->death_cursor = 0;
  w->n_toi_events = 0;
  /* queue_actions (env:741-755): dead agents' actions are dropped */

  /* ---- PRE_STEP, groups boxes, box_items, heals, walls, agents ---- */
  /* boxes/Object.pre_step (sem:853-856, 902-905) */
  for (int k = 0; k < e->n_pending; ++k)
    drop_box_item(e, e->pend_x[k], e->pend_y[k], &e->pend_shape[k], e->pend_owner[k]);
  e->n_pending = 0;
  /* agents/DynamicMotors.pre_step (sim:407-424) */
  for (int i = 0; i < A; ++i) {
    if (e->a_slot[i] < 0) continue;
    const uint8_t* a = actions + 6 * i;
    b2l_body* b = &w->bodies[e->a_slot[i]];
    float par = (float)(dtab[a[0]] * c->motor_impulse[0]);
    float nor = (float)(dtab[a[1]] * c->motor_impulse[1]);
    float ix = b->qc * par + (-b->qs) * nor;   /* R * b2Vec2(par, nor) */
    float iy = b->qs * par + b->qc * nor;
    float ang = (float)(dtab[a[2]] * c->motor_impulse[2]);
    b2l_apply_linear_impulse_center(w, e->a_slot[i], ix, iy, 1);
    b2l_apply_angular_impulse(w, e->a_slot[i], ang, 1);
  }
  /* agents/UseLast.pre_step (sem:300-309) -> Inventory.use (sem:206-213) */
  for (int i = 0; i < A; ++i) {
    if (e->a_slot[i] < 0 || !actions[6 * i + 4]) continue;
    if (e->inv_n[i] == 0) continue;
    inv_item it = e->inv[i][--e->inv_n[i]];
    b2l_body* b = &w->bodies[e->a_slot[i]];
    if (it.kind == MSV_ITEM_HEAL) {
      e->use_heal++;
      if (e->health[i] <= 0) e->ev[ORC_EV_Q10_SAVED_BY_HEAL]++;   /* zone damage brought it to <= 0 after last step's death check */
      agent_change_health(e, i, c->healing, MSV_CAUSE_NONE); /* sem:646-649 */
    } else {
      e->use_box++;
      v2 off = from_polar(c->box_item_offset, (double)b->a); /* sem:830-836 */
      assert(e->n_boxes < MB);
      int k = e->n_boxes++;
      e->box_shape[k] = it.shape;
      e->box_slot[k] = spawn_box_body(e, b->p.x + off.x, b->p.y + off.y, &it.shape);
      e->box_has_health[k] = 0; e->box_health[k] = 0; e->box_cause[k] = MSV_CAUSE_NONE;
      e->box_owner[k] = c->box_ownership ? it.owner : MSV_CAUSE_NONE; /* sem:876-884 */
      e->ev[ORC_EV_BOX_PLACED]++;
    }
  }
  /* agents/GiveLast.pre_step (sem:335-370) */
  {
    int taker[MA];
    for (int i = 0; i < A; ++i) {
      taker[i] = -1;
      if (e->a_slot[i] < 0) continue;
      int nb[B2L_MAX_BODIES]; int n = circle_query(e, e->a_slot[i], (float)c->give_radius, nb);
      float minDist = INFINITY;
      b2l_body* body = &w->bodies[e->a_slot[i]];
      for (int k = 0; k < n; ++k) {
        if (nb[k] == e->a_slot[i]) continue;
        b2l_body* o = &w->bodies[nb[k]];
        float dx = body->p.x - o->p.x, dy = body->p.y - o->p.y;
        float dist = sqrtf(dx * dx + dy * dy);
        if (dist < minDist) { minDist = dist; taker[i] = nb[k]; }
      }
    }
    for (int i = 0; i < A; ++i) {
      if (e->a_slot[i] < 0 || !actions[6 * i + 5] || taker[i] < 0) continue;
      int tidx; int kind = classify(e, taker[i], &tidx);
      if (kind != ORC_KIND_AGENT) { if (e->inv_n[i] > 0) e->ev[ORC_EV_GIVE_BLOCKED_BY_BODY]++; continue; }   /* no Inventory module: sem:196-198 (Q6) */
      if (c->teams && team_of(e, tidx) != team_of(e, i)) { if (e->inv_n[i] > 0) e->ev[ORC_EV_GIVE_STRANGER]++; continue; } /* strangers sem:344-349 */
      if (e->inv_n[i] == 0) continue;          /* IndexError sem:199-202 */
      inv_item it = e->inv[i][--e->inv_n[i]];
      if (e->inv_n[tidx] + 1 <= c->inv_slots) { e->inv[tidx][e->inv_n[tidx]++] = it; e->ev[ORC_EV_GIVE_OK]++; } /* else lost, Q6 */
      else e->ev[ORC_EV_GIVE_TO_FULL]++;
    }
  }
  /* agents/Melee.pre_step (sem:584-617) / ContinuousMelee (sem:531-554) */
  {
    int target[MA];
    for (int i = 0; i < A; ++i) {
      target[i] = -1;
      if (e->a_slot[i] < 0) continue;
      b2l_body* b = &w->bodies[e->a_slot[i]];
      v2 hand = from_polar(c->melee_range, (double)b->a);
      v2 end = V(b->p.x + hand.x, b->p.y + hand.y);
      float f; target[i] = b2l_raycast(w, b->p, end, &f, 0);
    }
    for (int i = 0; i < A; ++i) {
      if (e->a_slot[i] < 0) continue;
      int attack = actions[6 * i + 3];
      int on_cooldown = c->melee_cooldown >= 0 && e->cooldown[i] > 0;
      if (target[i] >= 0 && attack && !on_cooldown) {
        int cause = c->teams ? MSV_CAUSE_TEAM0 + team_of(e, i) : i;
        int tidx; int kind = classify(e, target[i], &tidx);
        if (kind == ORC_KIND_AGENT) {
          e->ev[ORC_EV_MELEE_HIT_AGENT]++;
          if (c->teams && team_of(e, tidx) == team_of(e, i)) e->ev[ORC_EV_MELEE_TEAMMATE_IMMUNE]++;
          agent_change_health(e, tidx, -c->melee_damage, cause);
        } else if (kind == ORC_KIND_BOX) {
          e->ev[ORC_EV_MELEE_HIT_BOX]++;
          if (!e->box_has_health[tidx]) e->ev[ORC_EV_Q9_FRESH_BOX_HIT]++;
          else if (e->box_owner[tidx] != MSV_CAUSE_NONE && cause != e->box_owner[tidx]) e->ev[ORC_EV_OWNED_BOX_PROTECTED]++;
          box_change_health(e, tidx, -c->melee_damage, cause);
        } else e->ev[ORC_EV_Q3_COOLDOWN_BURNT]++;   /* wall / item hit: only the cooldown is consumed */
        if (c->melee_cooldown >= 0) e->cooldown[i] = c->melee_cooldown;
      }
    }
    if (c->melee_cooldown >= 0)
      for (int i = 0; i < A; ++i) if (e->cooldown[i] > 0) e->cooldown[i]--;
  }

  /* ---- world step (sim:236-240) ---- */
  for (int s = 0; s < 2; ++s) b2l_step(w, (float)(1.0 / 60), 10, 10);

  /* ---- POST_STEP ---- */
  /* boxes/Health.post_step (sem:429-435) */
  for (int k = 0; k < e->n_boxes; ++k)
    if (!e->box_has_health[k]) { e->box_has_health[k] = 1; e->box_health[k] = c->box_health; }
  for (int k = 0; k < e->n_boxes;) {
    if (e->box_health[k] <= 0) {
      b2l_body* b = &w->bodies[e->box_slot[k]];
      assert(e->n_pending < MB);
      int q = e->n_pending++;
      e->pend_x[q] = b->p.x; e->pend_y[q] = b->p.y;
      e->pend_shape[q] = e->box_shape[k]; e->pend_shape[q].rehulled = 1; /* sim.prototype -> copy_shape */
      e->pend_owner[q] = c->box_ownership ? e->box_cause[k] : MSV_CAUSE_NONE; /* sem:911-912 */
      e->ev[ORC_EV_BOX_DESTROYED]++;
      remove_box(e, k);
    } else ++k;
  }
  /* agents/Cameras.post_step (sim:333-334) */
  cameras_update(e);
  /* agents/Health.post_step -> despawn(dead) (sem:429-448) */
  {
    int dead[MA], nd = 0;
    for (int i = 0; i < A; ++i) if (e->a_slot[i] >= 0 && e->health[i] <= 0) dead[nd++] = i;
    if (nd > 0) {
      /* TrackDeaths (sim:281-284) */
      for (int k = 0; k < nd; ++k) e->deaths[e->n_deaths++] = dead[k];
      e->ev[ORC_EV_DEATH] += nd;
      if (nd >= 2) e->ev[ORC_EV_MULTI_DEATH_STEP]++;
      for (int k = 0; k < nd; ++k) {
        if (e->inv_n[dead[k]] >= 2) e->ev[ORC_EV_DEATHDROP_2PLUS]++;
        else if (e->inv_n[dead[k]] == 1) e->ev[ORC_EV_DEATHDROP_1]++;
        if (e->cause[dead[k]] == MSV_CAUSE_ZONE) e->ev[ORC_EV_DEATH_BY_ZONE]++;
        else if (e->cause[dead[k]] != MSV_CAUSE_NONE) e->ev[ORC_EV_DEATH_BY_MELEE]++;
        /* Q1: a survivor listed AFTER the dead agent reads a neighbour's stale seen-row this step */
        for (int j = dead[k] + 1; j < A; ++j) if (e->a_slot[j] >= 0 && e->health[j] > 0) { e->ev[ORC_EV_Q1_STALE_SEEN_ROW]++; break; }
      }
      /* DeathDrop.pre_despawn (sem:387-396) */
      int total = 0; for (int k = 0; k < nd; ++k) total += e->inv_n[dead[k]];
      double angles[MA * MS];
      for (int k = 0; k < total; ++k)
        angles[k] = 2 * M_PI * draw_uniform(e, ORC_STREAM_DEATH, e->death_cursor + k,
                                            e->have_draws ? e->draws.death_u : 0);
      e->death_cursor += total;
      int top = total;
      for (int k = 0; k < nd; ++k) {
        int i = dead[k];
        b2l_body* b = &w->bodies[e->a_slot[i]];
        for (int j = 0; j < e->inv_n[i]; ++j) {
          double ang = angles[--top];
          v2 off = from_polar(c->drop_radius, ang);
          float x = b->p.x + off.x, y = b->p.y + off.y;
          if (e->inv[i][j].kind == MSV_ITEM_HEAL) drop_heal(e, x, y);
          else drop_box_item(e, x, y, &e->inv[i][j].shape, e->inv[i][j].owner);
        }
        e->inv_n[i] = 0;
      }
      /* Health.pre_despawn -> TrackKills (sem:437-448, 628-629) */
      for (int k = 0; k < nd; ++k) e->kill_cause[e->n_kills++] = e->cause[dead[k]];
      for (int k = 0; k < nd; ++k) { b2l_destroy_body(w, e->a_slot[dead[k]]); e->a_slot[dead[k]] = -1; }
    }
  }
  /* agents/AutoPickup.post_step (sem:278-283) */
  {
    int found[MA][B2L_MAX_BODIES]; int nf[MA];
    for (int i = 0; i < A; ++i) {
      nf[i] = 0;
      if (e->a_slot[i] < 0) continue;
      int q[B2L_MAX_BODIES]; int n = circle_query(e, e->a_slot[i], (float)c->pickup_radius, q);
      for (int k = 0; k < n; ++k) found[i][nf[i]++] = w->bodies[q[k]].id;
    }
    for (int i = 0; i < A; ++i) {
      if (e->a_slot[i] < 0) continue;
      for (int k = 0; k < nf[i]; ++k) {
        int slot = b2l_slot_of(w, found[i][k]);
        if (slot < 0) { e->ev[ORC_EV_Q7_DOUBLE_PICKUP]++; continue; }  /* Q7: already taken by an earlier agent */
        int idx; int kind = classify(e, slot, &idx);
        if (kind != ORC_KIND_HEAL && kind != ORC_KIND_ITEM) continue;
        if (e->inv_n[i] + 1 > c->inv_slots) { e->ev[ORC_EV_PICKUP_FULL]++; continue; } /* sem:184-185 */
        e->ev[kind == ORC_KIND_HEAL ? ORC_EV_PICKUP_HEAL : ORC_EV_PICKUP_BOX]++;
        inv_item it; memset(&it, 0, sizeof it);
        if (kind == ORC_KIND_HEAL) { it.kind = MSV_ITEM_HEAL; it.owner = MSV_CAUSE_NONE; remove_heal(e, idx); }
        else { it.kind = MSV_ITEM_BOX; it.shape = e->item_shape[idx]; it.owner = e->item_owner[idx]; remove_item(e, idx); }
        e->inv[i][e->inv_n[i]++] = it;
      }
    }
  }
  /* agents/SafeZone.post_step (sem:758-768) + tick (sem:776-811) */
  {
    for (int i = 0; i < A; ++i) {
      if (e->a_slot[i] < 0) continue;
      b2l_body* b = &w->bodies[e->a_slot[i]];
      float dx = b->c.x - e->zone_x, dy = b->c.y - e->zone_y;
      int inside = dx * dx + dy * dy <= e->zone_r * e->zone_r;
      if (e->zone_endgame || !inside) agent_change_health(e, i, -c->zone_damage, MSV_CAUSE_ZONE);
    }
    if (e->zone_t_cooldown == 0) { /* shrinking */
      if (!e->zone_endgame) {
        e->zone_t_shrink -= 1;
        if (e->zone_t_shrink > 0) {
          double t = (double)e->zone_t_shrink / c->zone_cooldown;
          double r1 = zone_radius(e, e->zone_phase), r2 = zone_radius(e, e->zone_phase + 1);
          e->zone_r = (float)(t * r1 + (1 - t) * r2);
          float t1 = (float)t, t2 = (float)(1 - t);
          e->zone_x = t1 * e->zone_cx[e->zone_phase] + t2 * e->zone_cx[e->zone_phase + 1];
          e->zone_y = t1 * e->zone_cy[e->zone_phase] + t2 * e->zone_cy[e->zone_phase + 1];
        } else {
          e->zone_t_cooldown = c->zone_cooldown;
          e->zone_phase += 1;
          e->zone_r = (float)zone_radius(e, e->zone_phase);
          e->zone_x = e->zone_cx[e->zone_phase]; e->zone_y = e->zone_cy[e->zone_phase];
          if (e->zone_phase == c->zone_phases - 1) e->zone_endgame = 1;
        }
      }
    } else {
      e->zone_t_cooldown -= 1;
      if (e->zone_t_cooldown <= 0) e->zone_t_shrink = c->zone_cooldown;
    }
  }

  /* ---- observations, rewards, done, stats (env:84-90) ---- */
  lidar_update(e); /* Lidar extension block: scanned on the state the observation describes */
  if (out) orc_observe(e, out);
  float rewards[MA]; memset(rewards, 0, sizeof rewards);
  memset(e->last_kills, 0, sizeof e->last_kills);
  if (!c->teams) { /* env:768-782 */
    for (int i = 0; i < A; ++i) rewards[i] += e->a_slot[i] >= 0 ? c->r_alive : c->r_dead;
    for (int k = 0; k < e->n_kills; ++k) {
      int killer = e->kill_cause[k];
      if (killer >= 0 && killer < A && e->a_slot[killer] >= 0) { rewards[killer] += c->r_kill; e->last_kills[killer]++; e->ev[ORC_EV_KILL_FFA]++; }
      else if (killer >= 0 && killer < A) e->ev[ORC_EV_Q5_DEAD_KILLER]++;     /* kill credit needs a living killer (env:777-780) */
    }
    for (int k = 0; k < e->n_deaths; ++k) rewards[e->deaths[k]] += c->r_death;
  } else { /* env:783-801 */
    int split = A / 2;
    for (int t = 0; t < 2; ++t) {
      float r = team_alive(e, t) ? c->r_alive : c->r_dead;
      for (int i = (t ? split : 0); i < (t ? A : split); ++i) rewards[i] += r;
    }
    for (int k = 0; k < e->n_kills; ++k) {
      int cz = e->kill_cause[k];
      if (cz != MSV_CAUSE_TEAM0 && cz != MSV_CAUSE_TEAM0 + 1) continue;
      int t = cz - MSV_CAUSE_TEAM0;
      for (int i = (t ? split : 0); i < (t ? A : split); ++i) rewards[i] += c->r_kill;
      e->last_kills[t]++; e->ev[ORC_EV_KILL_TEAM]++;
    }
    for (int k = 0; k < e->n_deaths; ++k) {
      int t = team_of(e, e->deaths[k]);
      for (int i = (t ? split : 0); i < (t ? A : split); ++i) rewards[i] += c->r_death;
    }
  }
  int n_alive = c->teams ? team_alive(e, 0) + team_alive(e, 1) : n_agents_alive(e);
  int done = c->gameover_mode == MSV_GAMEOVER_ALLDEAD ? n_alive == 0 : n_alive <= 1; /* env:810-831 */
  e->steps += 1;
  /* _update_stats (env:483-508) */
  if (!c->teams) for (int i = 0; i < A; ++i) e->stat_reward[i] += rewards[i];
  else { e->stat_reward[0] += rewards[0]; e->stat_reward[1] += rewards[A / 2]; }
  for (int i = 0; i < (c->teams ? 2 : A); ++i) e->stat_kills[i] += e->last_kills[i];
  e->stat_steps += 1; e->stat_heals += e->use_heal; e->stat_boxes += e->use_box;
  memcpy(e->last_rewards, rewards, sizeof rewards);
  for (int i = 0; i < A; ++i) e->ep_return[i] += rewards[i];
  if (out) {
    memcpy(out->rewards, rewards, sizeof rewards); out->done = done; out->n_toi_events = w->n_toi_events;
    if (done) { memcpy(out->episode_return, e->ep_return, sizeof e->ep_return); out->episode_length = e->steps; }
    /* BattleRoyale.post_step (sem:41-46): last module of the agents group, sees the post-death body list */
    out->br_over = c->battle_royale && n_agents_alive(e) <= 1;
    for (int i = 0; i < MA; ++i) out->br_results[i] = out->br_over && i < A && e->a_slot[i] >= 0;   /* .results only exists once over */
    /* ImmunityPhase.post_step (sem:667-674): Health.immune stays True for max(cooldown, 1) steps */
    { int cd = c->immunity_cooldown < 1 ? 1 : c->immunity_cooldown; out->immune = c->immunity_cooldown >= 0 && e->steps < cd; }
  }
  if (done) e->ev[ORC_EV_EPISODE_END]++;
  e->ev[ORC_EV_TOI_EVENT] += w->n_toi_events;
  if (done && c->auto_reset) {
    e->stat_episodes++;
    orc_reset(e, 0);
    if (out) {
      orc_out keep = *out;                       /* orc_observe clears the record */
      orc_observe(e, out);
      memcpy(out->rewards, keep.rewards, sizeof keep.rewards); out->done = 1; out->n_toi_events = keep.n_toi_events;
      memcpy(out->episode_return, keep.episode_return, sizeof keep.episode_return); out->episode_length = keep.episode_length;
      out->br_over = keep.br_over; memcpy(out->br_results, keep.br_results, sizeof keep.br_results);
      out->immune = c->immunity_cooldown >= 0;   /* the new episode's flag; br_* and episode_* describe the finished one */
    }
  }
}

/* --------------------------------------------------------- observations --- */
static void agent_row(const oracle_env* e, int i, float* row) { /* env:659-691 */
  int s = 0;
  row[s++] = (float)i;
  if (e->cfg.teams) row[s++] = (float)team_of(e, i);
  if (e->a_slot[i] < 0) { for (int k = 0; k < 7; ++k) row[s++] = 0.0f; return; }
  const b2l_body* b = &e->w->bodies[e->a_slot[i]];
  row[s++] = (float)e->health[i];
  row[s++] = b->p.x; row[s++] = b->p.y; row[s++] = b->a;
  row[s++] = b->v.x; row[s++] = b->v.y; row[s++] = b->w;
}

static int seen_has(const oracle_env* e, int row, int id) {
  if (row < 0 || row >= e->n_seen_rows) return 0;
  for (int k = 0; k < e->seen_n[row]; ++k) if (e->seen[row][k] == id) return 1;
  return 0;
}

/* fetch_observations (env:510-657) */
void orc_observe(oracle_env* e, orc_out* o) {
  const msv_config* c = &e->cfg;
  const int A = c->n_agents, B = c->n_boxes, H = c->n_heals, S = 8 + (c->teams ? 1 : 0);
  memset(o, 0, sizeof *o);
  float rows[MA][ORC_S_MAX];
  for (int i = 0; i < A; ++i) agent_row(e, i, rows[i]);
  /* rank of alive agent i in the CURRENT agents.bodies list (Q1) */
  int rank[MA]; { int r = 0; for (int i = 0; i < A; ++i) rank[i] = e->a_slot[i] >= 0 ? r++ : -1; }
  for (int i = 0; i < A; ++i) {
    memcpy(o->agent + i * S, rows[i], S * sizeof(float));
    int k = 0;
    for (int j = 0; j < A; ++j) {
      if (j == i) continue;
      memcpy(o->others + (i * (A - 1) + k) * S, rows[j], S * sizeof(float));
      float m = 1.0f;
      if (e->a_slot[i] >= 0 && e->a_slot[j] >= 0 && seen_has(e, rank[i], e->a_id[j])) m = 0.0f;
      o->others_mask[i * (A - 1) + k] = m;
      k++;
    }
  }
  float zone[6] = {e->zone_x, e->zone_y, e->zone_r, 0.0f, 0.0f, 0.0f};
  if (e->zone_phase < c->zone_phases - 1) { /* env:537-544 */
    zone[3] = e->zone_cx[e->zone_phase + 1]; zone[4] = e->zone_cy[e->zone_phase + 1];
    zone[5] = (float)zone_radius(e, e->zone_phase + 1);
  }
  for (int i = 0; i < A; ++i) memcpy(o->zone + i * 6, zone, sizeof zone);
  for (int i = 0; i < A; ++i) {
    for (int h = 0; h < H; ++h) {
      float m;
      if (h < e->n_heals) {
        const b2l_body* b = &e->w->bodies[e->heal_slot[h]];
        o->heals[(i * H + h) * 2 + 0] = b->p.x; o->heals[(i * H + h) * 2 + 1] = b->p.y;
        if (c->omniscient) m = 0.0f;
        else m = (e->a_slot[i] >= 0 && seen_has(e, rank[i], b->id)) ? 0.0f : 1.0f; /* env:706-715 */
      } else m = 1.0f;
      o->heals_mask[i * H + h] = m;
    }
    for (int k = 0; k < B; ++k) {
      float m;
      if (k < e->n_boxes) {
        const b2l_body* b = &e->w->bodies[e->box_slot[k]];
        float* dst = o->boxes + (i * B + k) * 11;
        for (int v = 0; v < 4; ++v) { dst[2 * v] = b->shape.verts[v].x; dst[2 * v + 1] = b->shape.verts[v].y; }
        dst[8] = b->p.x; dst[9] = b->p.y; dst[10] = b->a;
        if (c->omniscient) m = 0.0f;
        else m = (e->a_slot[i] >= 0 && seen_has(e, rank[i], b->id)) ? 0.0f : 1.0f;
      } else m = 1.0f;
      o->boxes_mask[i * B + k] = m;
    }
    for (int k = 0; k < B; ++k) {
      float m;
      if (k < e->n_items) {
        const b2l_body* b = &e->w->bodies[e->item_slot[k]];
        b2l_shape sh; shape_of(&e->item_shape[k], &sh);
        float* dst = o->box_items + (i * B + k) * 10;
        for (int v = 0; v < 4; ++v) { dst[2 * v] = sh.verts[v].x; dst[2 * v + 1] = sh.verts[v].y; }
        dst[8] = b->p.x; dst[9] = b->p.y;
        if (c->omniscient) m = 0.0f;
        else m = (e->a_slot[i] >= 0 && seen_has(e, rank[i], b->id)) ? 0.0f : 1.0f;
      } else m = 1.0f;
      o->box_items_mask[i * B + k] = m;
    }
    o->heal_slot_mask[i] = 1.0f; o->box_slot_mask[i] = 1.0f;
    if (e->a_slot[i] >= 0 && e->inv_n[i] > 0) { /* env:635-654 */
      const inv_item* it = &e->inv[i][e->inv_n[i] - 1];
      if (H > 0 && it->kind == MSV_ITEM_HEAL) { o->heal_slot[i] = (float)c->healing; o->heal_slot_mask[i] = 0.0f; }
      if (B > 0 && it->kind == MSV_ITEM_BOX) {
        b2l_shape sh; shape_of(&it->shape, &sh);
        for (int v = 0; v < 4; ++v) { o->box_slot[i * 8 + 2 * v] = sh.verts[v].x; o->box_slot[i * 8 + 2 * v + 1] = sh.verts[v].y; }
        o->box_slot_mask[i] = 0.0f;
      }
    }
  }
  int L = c->lidar_n;
  for (int i = 0; i < A; ++i)
    for (int r = 0; r < L; ++r) { o->lidar_frac[i * L + r] = e->lidar_frac[i][r]; o->lidar_hit[i * L + r] = e->lidar_hit[i][r]; }
}

/* ------------------------------------------------------- state exchange --- */
static int aa_index(int i, int j) { return j * (j - 1) / 2 + i; } /* i<j */

void orc_get_state(oracle_env* e, msv_env_state* s) {
  const msv_config* c = &e->cfg;
  b2l_world* w = e->w;
  memset(s, 0, sizeof *s);
  for (int i = 0; i < c->n_agents; ++i) {
    s->alive[i] = e->a_slot[i] >= 0;
    s->cooldown[i] = e->cooldown[i];
    s->cause[i] = MSV_CAUSE_NONE;
    if (!s->alive[i]) continue;
    b2l_body* b = &w->bodies[e->a_slot[i]];
    s->health[i] = e->health[i]; s->cause[i] = e->cause[i];
    s->x[i] = b->c.x; s->y[i] = b->c.y; s->angle[i] = b->a;
    s->vx[i] = b->v.x; s->vy[i] = b->v.y; s->omega[i] = b->w;
    s->sleep_time[i] = b->sleepTime; s->awake[i] = b->awake;
    memcpy(s->fat[i], b->fat, sizeof b->fat);
    s->inv_n[i] = e->inv_n[i];
    for (int k = 0; k < e->inv_n[i]; ++k) {
      s->inv_kind[i][k] = e->inv[i][k].kind; s->inv_shape[i][k] = e->inv[i][k].shape;
      s->inv_owner[i][k] = e->inv[i][k].owner;
    }
  }
  s->n_boxes = e->n_boxes;
  for (int k = 0; k < e->n_boxes; ++k) {
    b2l_body* b = &w->bodies[e->box_slot[k]];
    s->box_x[k] = b->p.x; s->box_y[k] = b->p.y; s->box_shape[k] = e->box_shape[k];
    s->box_health[k] = e->box_health[k]; s->box_has_health[k] = e->box_has_health[k];
    s->box_cause[k] = e->box_cause[k]; s->box_owner[k] = e->box_owner[k]; s->box_seq[k] = b->seq;
  }
  s->n_items = e->n_items;
  for (int k = 0; k < e->n_items; ++k) {
    b2l_body* b = &w->bodies[e->item_slot[k]];
    s->item_x[k] = b->p.x; s->item_y[k] = b->p.y; s->item_shape[k] = e->item_shape[k];
    s->item_owner[k] = e->item_owner[k]; s->item_seq[k] = b->seq;
  }
  s->n_heals = e->n_heals;
  for (int k = 0; k < e->n_heals; ++k) {
    b2l_body* b = &w->bodies[e->heal_slot[k]];
    s->heal_x[k] = b->p.x; s->heal_y[k] = b->p.y; s->heal_seq[k] = b->seq;
  }
  s->n_pending = e->n_pending;
  for (int k = 0; k < e->n_pending; ++k) {
    s->pend_x[k] = e->pend_x[k]; s->pend_y[k] = e->pend_y[k];
    s->pend_shape[k] = e->pend_shape[k]; s->pend_owner[k] = e->pend_owner[k];
  }
  for (int z = 0; z < e->n_zones; ++z) { s->zone_cx[z] = e->zone_cx[z]; s->zone_cy[z] = e->zone_cy[z]; }
  s->zone_phase = e->zone_phase; s->zone_t_cooldown = e->zone_t_cooldown;
  s->zone_t_shrink = e->zone_t_shrink; s->zone_endgame = e->zone_endgame;
  s->zone_cur_x = e->zone_x; s->zone_cur_y = e->zone_y; s->zone_cur_r = e->zone_r;
  for (int k = 0; k < B2L_MAX_CONTACTS; ++k) {
    b2l_contact* ct = &w->contacts[k];
    if (!ct->used) continue;
    int ia, ib; int ka = classify(e, ct->a, &ia), kb = classify(e, ct->b, &ib);
    msv_pair* p = 0;
    if (ka == ORC_KIND_AGENT && kb == ORC_KIND_AGENT) p = &s->pair_aa[aa_index(ia < ib ? ia : ib, ia < ib ? ib : ia)];
    else if (ka == ORC_KIND_BOX && kb == ORC_KIND_AGENT) p = &s->pair_ab[ib][ia];
    else if (ka == ORC_KIND_WALL && kb == ORC_KIND_AGENT) p = &s->pair_aw[ib][ia];
    else assert(0);
    p->seq = ct->seq;
    p->flags = (ct->flags & B2L_TOUCHING ? MSV_PAIR_TOUCHING : 0) | (ct->flags & B2L_ENABLED ? MSV_PAIR_ENABLED : 0);
    p->normal_impulse = ct->ni; p->tangent_impulse = ct->ti;
  }
  s->first_step = w->inv_dt0 == 0.0f;
  s->steps = e->steps; s->episode = e->episode;
  s->body_seq = w->body_seq; s->contact_seq = w->contact_seq;
  for (int i = 0; i < MA; ++i) { s->stat_reward[i] = e->stat_reward[i]; s->stat_kills[i] = e->stat_kills[i]; }
  s->stat_steps = e->stat_steps; s->stat_heals_used = e->stat_heals; s->stat_boxes_placed = e->stat_boxes;
  s->stat_episodes = (int32_t)e->stat_episodes;
  for (int i = 0; i < MA; ++i) s->ep_return[i] = i < e->cfg.n_agents ? e->ep_return[i] : 0.0f;
}

static void force_seq(b2l_world* w, int slot, int seq) { w->bodies[slot].seq = seq; }

void orc_set_state(oracle_env* e, const msv_env_state* s) {
  const msv_config* c = &e->cfg;
  b2l_world* w = e->w;
  b2l_world_clear(w);
  int B0 = c->n_boxes, H0 = c->n_heals;
  e->n_boxes = s->n_boxes;
  for (int k = 0; k < s->n_boxes; ++k) {
    e->box_shape[k] = s->box_shape[k];
    e->box_slot[k] = spawn_box_body(e, s->box_x[k], s->box_y[k], &s->box_shape[k]);
    force_seq(w, e->box_slot[k], s->box_seq[k]);
    e->box_health[k] = s->box_health[k]; e->box_has_health[k] = s->box_has_health[k];
    e->box_cause[k] = s->box_cause[k]; e->box_owner[k] = s->box_owner[k];
  }
  e->n_items = s->n_items;
  for (int k = 0; k < s->n_items; ++k) {
    e->item_slot[k] = spawn_sensor(e, s->item_x[k], s->item_y[k], e->item_r);
    force_seq(w, e->item_slot[k], s->item_seq[k]);
    e->item_shape[k] = s->item_shape[k]; e->item_owner[k] = s->item_owner[k];
  }
  e->n_heals = s->n_heals;
  for (int k = 0; k < s->n_heals; ++k) {
    e->heal_slot[k] = spawn_sensor(e, s->heal_x[k], s->heal_y[k], e->heal_r);
    force_seq(w, e->heal_slot[k], s->heal_seq[k]);
  }
  {
    b2l_shape sh; b2l_set_as_box(&sh, e->wall_hx, e->wall_hy);
    float o = e->wall_off; float hp = (float)(M_PI / 2);
    float wx[4] = {-o, 0.0f, o, 0.0f}, wy[4] = {0.0f, o, 0.0f, -o}, wa[4] = {0.0f, hp, 0.0f, hp};
    for (int k = 0; k < 4; ++k) {
      e->wall_slot[k] = b2l_create_body(w, B2L_STATIC, wx[k], wy[k], wa[k], &sh, 1.0f, 0, 0.8f, 0.8f);
      force_seq(w, e->wall_slot[k], B0 + H0 + k);
    }
  }
  for (int i = 0; i < MA; ++i) e->a_slot[i] = -1;
  for (int i = 0; i < c->n_agents; ++i) {
    e->cooldown[i] = s->cooldown[i]; e->inv_n[i] = 0;
    e->health[i] = s->health[i]; e->cause[i] = s->cause[i];
    if (!s->alive[i]) continue;
    int slot = spawn_agent(e, s->x[i], s->y[i]);
    e->a_slot[i] = slot; e->a_id[i] = w->bodies[slot].id;
    force_seq(w, slot, B0 + H0 + 4 + i);
    b2l_body* b = &w->bodies[slot];
    b->a = b->a0 = s->angle[i]; b2l_sync_transform(b);
    b->v.x = s->vx[i]; b->v.y = s->vy[i]; b->w = s->omega[i];
    b->sleepTime = s->sleep_time[i]; b->awake = s->awake[i];
    memcpy(b->fat, s->fat[i], sizeof b->fat);
    e->inv_n[i] = s->inv_n[i];
    for (int k = 0; k < s->inv_n[i]; ++k) {
      e->inv[i][k].kind = s->inv_kind[i][k]; e->inv[i][k].shape = s->inv_shape[i][k];
      e->inv[i][k].owner = s->inv_owner[i][k];
    }
  }
  /* every proxy counts as freshly created: the next Step starts with a full
   * FindNewContacts, exactly what an injected state needs (no-op on any state
   * reachable by stepping, where contact <=> fat-AABB overlap already holds) */
  for (int i = 0; i < B2L_MAX_BODIES; ++i) w->bodies[i].moved = w->bodies[i].used;
  w->newFixture = 1;
  e->n_pending = s->n_pending;
  for (int k = 0; k < s->n_pending; ++k) {
    e->pend_x[k] = s->pend_x[k]; e->pend_y[k] = s->pend_y[k];
    e->pend_shape[k] = s->pend_shape[k]; e->pend_owner[k] = s->pend_owner[k];
  }
  for (int z = 0; z < e->n_zones; ++z) { e->zone_cx[z] = s->zone_cx[z]; e->zone_cy[z] = s->zone_cy[z]; }
  e->zone_phase = s->zone_phase; e->zone_t_cooldown = s->zone_t_cooldown;
  e->zone_t_shrink = s->zone_t_shrink; e->zone_endgame = s->zone_endgame;
  e->zone_x = s->zone_cur_x; e->zone_y = s->zone_cur_y; e->zone_r = s->zone_cur_r;
  /* contacts */
  for (int j = 1; j < c->n_agents; ++j)
    for (int i = 0; i < j; ++i) {
      const msv_pair* p = &s->pair_aa[aa_index(i, j)];
      if (p->seq && e->a_slot[i] >= 0 && e->a_slot[j] >= 0)
        b2l_inject_contact(w, e->a_slot[i], e->a_slot[j], p->seq,
                           (p->flags & MSV_PAIR_TOUCHING ? B2L_TOUCHING : 0) | (p->flags & MSV_PAIR_ENABLED ? B2L_ENABLED : 0),
                           p->normal_impulse, p->tangent_impulse);
    }
  for (int i = 0; i < c->n_agents; ++i) {
    if (e->a_slot[i] < 0) continue;
    for (int k = 0; k < s->n_boxes; ++k) {
      const msv_pair* p = &s->pair_ab[i][k];
      if (p->seq) b2l_inject_contact(w, e->box_slot[k], e->a_slot[i], p->seq,
                           (p->flags & MSV_PAIR_TOUCHING ? B2L_TOUCHING : 0) | (p->flags & MSV_PAIR_ENABLED ? B2L_ENABLED : 0),
                           p->normal_impulse, p->tangent_impulse);
    }
    for (int k = 0; k < 4; ++k) {
      const msv_pair* p = &s->pair_aw[i][k];
      if (p->seq) b2l_inject_contact(w, e->wall_slot[k], e->a_slot[i], p->seq,
                           (p->flags & MSV_PAIR_TOUCHING ? B2L_TOUCHING : 0) | (p->flags & MSV_PAIR_ENABLED ? B2L_ENABLED : 0),
                           p->normal_impulse, p->tangent_impulse);
    }
  }
  w->inv_dt0 = s->first_step ? 0.0f : 1.0f / (float)(1.0 / 60);
  e->steps = s->steps; e->episode = s->episode;
  w->body_seq = s->body_seq; w->contact_seq = s->contact_seq;
  for (int i = 0; i < MA; ++i) { e->stat_reward[i] = s->stat_reward[i]; e->stat_kills[i] = s->stat_kills[i]; }
  e->stat_steps = s->stat_steps; e->stat_heals = s->stat_heals_used; e->stat_boxes = s->stat_boxes_placed;
  e->stat_episodes = s->stat_episodes;
  for (int i = 0; i < MA; ++i) e->ep_return[i] = s->ep_return[i];
  /* sensors are recomputed from the injected world (Cameras.post_reset analogue) */
  cameras_update(e);
  lidar_update(e);
}

void orc_get_events(oracle_env* e, int64_t* out, int32_t n) {
  for (int k = 0; k < n && k < ORC_EV_COUNT; ++k) out[k] = e->ev[k];
}

void orc_flush_stats(oracle_env* e, msv_stats* out) {
  memset(out, 0, sizeof *out);
  for (int i = 0; i < MA; ++i) { out->reward[i] = e->stat_reward[i]; out->kills[i] = e->stat_kills[i]; }
  out->steps = e->stat_steps; out->heals_used = e->stat_heals; out->boxes_placed = e->stat_boxes;
  out->episodes = e->stat_episodes;
  memset(e->stat_reward, 0, sizeof e->stat_reward); memset(e->stat_kills, 0, sizeof e->stat_kills);
  e->stat_steps = e->stat_heals = e->stat_boxes = 0; e->stat_episodes = 0;
}

/* ------------------------------------------------------ CPU baseline run -- */
typedef struct { const msv_config* cfg; uint64_t seed; int first, count, steps; int64_t done_steps; } rollout_job;

static void* rollout_thread(void* arg) {
  rollout_job* j = (rollout_job*)arg;
  msv_config cfg = *j->cfg; cfg.auto_reset = 1;
  orc_out* out = (orc_out*)malloc(sizeof(orc_out));
  for (int k = 0; k < j->count; ++k) {
    oracle_env* e = orc_create(&cfg, j->seed, j->first + k);
    orc_reset(e, out);
    uint8_t act[MA * 6];
    for (int t = 0; t < j->steps; ++t) {
      orc_philox_actions(j->seed, (uint32_t)(j->first + k), 0u, (uint32_t)t, cfg.n_agents, act);
      orc_step(e, act, out);
      j->done_steps++;
    }
    orc_destroy(e);
  }
  free(out);
  return 0;
}

int64_t orc_rollout(const msv_config* cfg, uint64_t seed, int32_t n_envs, int32_t steps, int32_t n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256]; rollout_job jobs[256];
  int per = (n_envs + n_threads - 1) / n_threads, first = 0, used = 0;
  for (int t = 0; t < n_threads && first < n_envs; ++t) {
    int cnt = per < n_envs - first ? per : n_envs - first;
    jobs[t].cfg = cfg; jobs[t].seed = seed; jobs[t].first = first; jobs[t].count = cnt;
    jobs[t].steps = steps; jobs[t].done_steps = 0;
    pthread_create(&th[t], 0, rollout_thread, &jobs[t]);
    first += cnt; used++;
  }
  int64_t total = 0;
  for (int t = 0; t < used; ++t) { pthread_join(th[t], 0); total += jobs[t].done_steps; }
  return total;
}

/* ------------------------------------------------------------ batch API --- */
struct orc_batch { msv_config cfg; int n, n_threads; oracle_env** envs; orc_out* outs; };
typedef struct { orc_batch* b; int first, count; const uint8_t* actions; float* rewards; uint8_t* dones; int op; } batch_job;

static void* batch_thread(void* arg) {
  batch_job* j = (batch_job*)arg;
  int A = j->b->cfg.n_agents;
  for (int k = j->first; k < j->first + j->count; ++k) {
    orc_out* o = &j->b->outs[k - j->first + j->first];
    if (j->op == 0) orc_reset(j->b->envs[k], o);
    else {
      orc_step(j->b->envs[k], j->actions + (size_t)k * A * 6, o);
      if (j->rewards) for (int i = 0; i < A; ++i) j->rewards[(size_t)k * A + i] = o->rewards[i];
      if (j->dones) j->dones[k] = (uint8_t)o->done;
    }
  }
  return 0;
}
static void batch_run(orc_batch* b, int op, const uint8_t* actions, float* rewards, uint8_t* dones) {
  pthread_t th[256]; batch_job jobs[256];
  int nt = b->n_threads, per = (b->n + nt - 1) / nt, first = 0, used = 0;
  for (int t = 0; t < nt && first < b->n; ++t) {
    int cnt = per < b->n - first ? per : b->n - first;
    batch_job jj = {b, first, cnt, actions, rewards, dones, op};
    jobs[t] = jj;
    pthread_create(&th[t], 0, batch_thread, &jobs[t]);
    first += cnt; used++;
  }
  for (int t = 0; t < used; ++t) pthread_join(th[t], 0);
}
orc_batch* orc_batch_create(const msv_config* cfg, uint64_t seed, int32_t n_envs, int32_t n_threads) {
  orc_batch* b = (orc_batch*)calloc(1, sizeof *b);
  b->cfg = *cfg; b->n = n_envs;
  b->n_threads = n_threads < 1 ? 1 : (n_threads > 256 ? 256 : n_threads);
  b->envs = (oracle_env**)calloc((size_t)n_envs, sizeof(oracle_env*));
  b->outs = (orc_out*)calloc((size_t)n_envs, sizeof(orc_out));
  for (int k = 0; k < n_envs; ++k) b->envs[k] = orc_create(cfg, seed, k);
  return b;
}
void orc_batch_destroy(orc_batch* b) {
  for (int k = 0; k < b->n; ++k) orc_destroy(b->envs[k]);
  free(b->envs); free(b->outs); free(b);
}
void orc_batch_reset(orc_batch* b) { batch_run(b, 0, 0, 0, 0); }
void orc_batch_step(orc_batch* b, const uint8_t* actions, float* rewards, uint8_t* dones) {
  batch_run(b, 1, actions, rewards, dones);
}
