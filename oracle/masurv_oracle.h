/*
 * masurv_oracle.h -- TEST ORACLE, not product code.
 *
 * Scalar CPU restatement (plain C) of the reference's step path
 *   MaSurvival.reset / step        masurvival/envs/masurvival_env.py:59-90
 *   Simulation.step                masurvival/simulation.py:233-242
 *   the semantics.py modules in the wiring order of env:320-389
 * on top of b2lite (the restated Box2D subset).  One oracle_env == one
 * reference `MaSurvival` instance.  It is pinned against the reference's OWN
 * Python (imported unmodified from /root/reference on the Box2D/gym shims in
 * oracle/shim/) through the committed golden vectors under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use anything under oracle/.
 */
#ifndef MASURV_ORACLE_H
#define MASURV_ORACLE_H
#include "../include/masurv.h"
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_S_MAX 9 /* agent row: id,(team),health,x,y,angle,vx,vy,w */

/* Observation/reward/done of one env, reference layout, tightly packed with
 * the ACTUAL dims (A, B, H, s, L) of the config. */
typedef struct orc_out {
  float agent[MSV_MAX_AGENTS * ORC_S_MAX];
  float others[MSV_MAX_AGENTS * (MSV_MAX_AGENTS - 1) * ORC_S_MAX];
  float others_mask[MSV_MAX_AGENTS * (MSV_MAX_AGENTS - 1)];
  float zone[MSV_MAX_AGENTS * 6];
  float heals[MSV_MAX_AGENTS * MSV_MAX_HEALS * 2];
  float heals_mask[MSV_MAX_AGENTS * MSV_MAX_HEALS];
  float heal_slot[MSV_MAX_AGENTS];
  float heal_slot_mask[MSV_MAX_AGENTS];
  float boxes[MSV_MAX_AGENTS * MSV_MAX_BOXES * 11];
  float boxes_mask[MSV_MAX_AGENTS * MSV_MAX_BOXES];
  float box_items[MSV_MAX_AGENTS * MSV_MAX_BOXES * 10];
  float box_items_mask[MSV_MAX_AGENTS * MSV_MAX_BOXES];
  float box_slot[MSV_MAX_AGENTS * 8];
  float box_slot_mask[MSV_MAX_AGENTS];
  float lidar_frac[MSV_MAX_AGENTS * MSV_MAX_LASERS];
  int32_t lidar_hit[MSV_MAX_AGENTS * MSV_MAX_LASERS]; /* 0 none, else kind<<8|id */
  float rewards[MSV_MAX_AGENTS];
  int32_t done;
  int32_t n_toi_events;
  float episode_return[MSV_MAX_AGENTS]; /* valid when done: per-agent return of the finished episode */
  int32_t episode_length;               /* valid when done */
  int32_t immune;                       /* Health.immune under ImmunityPhase (sem:652-674) */
  int32_t br_over;                      /* BattleRoyale.over / .results (sem:31-46) */
  int32_t br_results[MSV_MAX_AGENTS];
} orc_out;

#define ORC_KIND_AGENT 1
#define ORC_KIND_BOX 2
#define ORC_KIND_ITEM 3
#define ORC_KIND_HEAL 4
#define ORC_KIND_WALL 5

/* Explicit random draws (parity injection).  When a pointer is NULL the
 * oracle derives the draw from its own Philox4x32-10 stream. */
typedef struct orc_draws {
  const double* shuffle_u; /* grid^2-1 uniforms, Fisher-Yates i=n-1..1 */
  const double* box_z;     /* 2 standard normals per box (w then h) */
  const double* zone_u;    /* 2 uniforms per zone, LAST zone first, x then y */
  const double* death_u;   /* uniforms for DeathDrop, rng.random(n) order */
} orc_draws;

/* Event-coverage census: how often each rule / quirk of SURVEY.md 8a fired in
 * this env since it was created.  Test infrastructure: the golden generator and the lock-step
 * drivers sum these and assert that every one of them was exercised. */
enum {
  ORC_EV_DEATH, ORC_EV_DEATH_BY_ZONE, ORC_EV_DEATH_BY_MELEE, ORC_EV_MULTI_DEATH_STEP,
  ORC_EV_KILL_FFA, ORC_EV_KILL_TEAM, ORC_EV_Q5_DEAD_KILLER,
  ORC_EV_GIVE_OK, ORC_EV_GIVE_TO_FULL, ORC_EV_GIVE_STRANGER, ORC_EV_GIVE_BLOCKED_BY_BODY,
  ORC_EV_DEATHDROP_1, ORC_EV_DEATHDROP_2PLUS,
  ORC_EV_PICKUP_HEAL, ORC_EV_PICKUP_BOX, ORC_EV_PICKUP_FULL, ORC_EV_Q7_DOUBLE_PICKUP,
  ORC_EV_BOX_PLACED, ORC_EV_BOX_DESTROYED, ORC_EV_Q9_FRESH_BOX_HIT, ORC_EV_OWNED_BOX_PROTECTED,
  ORC_EV_MELEE_HIT_AGENT, ORC_EV_MELEE_HIT_BOX, ORC_EV_MELEE_TEAMMATE_IMMUNE, ORC_EV_Q3_COOLDOWN_BURNT,
  ORC_EV_Q1_STALE_SEEN_ROW, ORC_EV_Q10_SAVED_BY_HEAL, ORC_EV_TOI_EVENT, ORC_EV_EPISODE_END,
  ORC_EV_COUNT
};

typedef struct oracle_env oracle_env;

oracle_env* orc_create(const msv_config* cfg, uint64_t seed, int64_t env_id);
void orc_destroy(oracle_env* e);
void orc_set_draws(oracle_env* e, const orc_draws* d); /* NULL = Philox */
void orc_reset(oracle_env* e, orc_out* out);
void orc_step(oracle_env* e, const uint8_t* actions /*[A][6]*/, orc_out* out);
void orc_observe(oracle_env* e, orc_out* out);
void orc_get_state(oracle_env* e, msv_env_state* s);
void orc_set_state(oracle_env* e, const msv_env_state* s);
void orc_flush_stats(oracle_env* e, msv_stats* out);
void orc_get_events(oracle_env* e, int64_t* out, int32_t n); /* ORC_EV_* counters */

/* multi-env, multi-thread driver used as the CPU baseline: n envs, `steps`
 * steps each with Philox actions, auto-reset on done.  Returns env-steps. */
int64_t orc_rollout(const msv_config* cfg, uint64_t seed, int32_t n_envs,
                    int32_t steps, int32_t n_threads);

/* batch of n envs stepped together by n_threads host threads: the CPU arm of
 * bench.py (`--impl reference` / cpu_baseline). */
typedef struct orc_batch orc_batch;
orc_batch* orc_batch_create(const msv_config* cfg, uint64_t seed, int32_t n_envs, int32_t n_threads);
void orc_batch_destroy(orc_batch* b);
void orc_batch_reset(orc_batch* b);
/* actions uint8[n][A][6] -> rewards float[n][A], dones uint8[n] */
void orc_batch_step(orc_batch* b, const uint8_t* actions, float* rewards, uint8_t* dones);

void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* uniform double for (env, episode, step, stream, k) */
double orc_philox_uniform(uint64_t seed, uint32_t env, uint32_t episode,
                          uint32_t step, uint32_t stream, uint32_t k);
/* the synthetic random action stream shared by bench/tests */
void orc_philox_actions(uint64_t seed, uint32_t env, uint32_t episode,
                        uint32_t step, int32_t n_agents, uint8_t* out);

#define ORC_STREAM_SHUFFLE 0
#define ORC_STREAM_BOX 1
#define ORC_STREAM_ZONE 2
#define ORC_STREAM_DEATH 3
#define ORC_STREAM_ACTION 4

#ifdef __cplusplus
}
#endif
#endif
