/*
 * mg_oracle.h — CPU restatement of the reference's MiniGridEnv hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (minigrid_b200/) never imports, links or falls back to this code.
 *
 * Parity status: PINNED. The restatement is checked (tests/test_oracle_vs_reference.py, run in the
 * build container where /root/reference exists) against the unmodified Python reference imported
 * through oracle/ref_shim, and (everywhere) against tests/golden/ fixtures that
 * oracle/gen_golden.py produced from that same Python reference, plus the reference's own
 * known-answer vectors (minigrid/wrappers.py:26-41,226-234,818-830; tests/test_wrappers.py:364-380).
 */
#ifndef MG_ORACLE_H
#define MG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* env kinds (the four generators BASELINE.json's configs name) */
enum { MGO_EMPTY = 0, MGO_DOORKEY = 1, MGO_CROSSING = 2, MGO_FOURROOMS = 3, MGO_LAVAGAP = 4, MGO_DISTSHIFT = 5, MGO_MULTIROOM = 6,
       /* SURVEY 8(f-1) generators restated ahead of their device kernels (oracle + fixtures only so far) */
       MGO_LOCKEDROOM = 7, MGO_PLAYGROUND = 8,
       /* a generator plus a step post-filter (gotodoor.py:130-149) */
       MGO_GOTODOOR = 9, MGO_FETCH = 10 /* fetch.py:118-175 */,
       MGO_REDBLUEDOORS = 11 /* redbluedoors.py:78-126 (width = 2 * height) */,
       MGO_GOTOOBJECT = 12 /* gotoobject.py:92-160 */, MGO_PUTNEAR = 13 /* putnear.py:99-199 */,
       MGO_MEMORY = 14 /* memory.py:90-164 */,
       /* RNG draws inside step */
       MGO_DYNOBSTACLES = 15 /* dynamicobstacles.py:107-167 */,
       /* core/roomgrid.py + unlock.py, unlockpickup.py, blockedunlockpickup.py, keycorridor.py */
       MGO_ROOMGRID = 16 };
/* vector autoreset modes (gymnasium.vector.AutoresetMode) */
enum { MGO_AUTORESET_NEXT_STEP = 0, MGO_AUTORESET_SAME_STEP = 1, MGO_AUTORESET_DISABLED = 2 };

typedef struct mgo_vec mgo_vec;

/* params: EMPTY  {random_start(0/1), start_x, start_y, start_dir}
 *         DOORKEY {}
 *         CROSSING {num_crossings, obstacle_type (9 = lava, 2 = wall)}
 *         FOURROOMS {}
 *         LAVAGAP {obstacle_type}; DISTSHIFT {strip2_row, start_x, start_y, start_dir}
 *         MULTIROOM {minNumRooms, maxNumRooms, maxRoomSize}; LOCKEDROOM {}; PLAYGROUND {}; GOTODOOR {}; FETCH {numObjs}; REDBLUEDOORS {}; GOTOOBJECT {numObjs}; PUTNEAR {numObjs};
 *         MEMORY {random_length}; DYNOBSTACLES {n_obstacles, random_start, start_x, start_y, start_dir};
 *         ROOMGRID {variant (0 Unlock, 1 UnlockPickup, 2 BlockedUnlockPickup, 3 KeyCorridor, 4 ObstructedMaze_1Dlhb,
 *         5 ObstructedMaze_Full, 6 ObstructedMaze_Full_V1), room_size, num_rows, num_cols[, key_in_box, blocked,
 *         agent_room_i | agent_room_j << 4, num_quarters]} */
mgo_vec *mgo_vec_create(int kind, int width, int height, int max_steps, int see_through_walls,
                        const int32_t *params, int n_params, int n_envs);
void mgo_vec_destroy(mgo_vec *v);
/* reward wrappers around every env of the vector (wrappers.py:68-184, 809-882); bonus wrappers are outermost */
int mgo_vec_set_no_death(mgo_vec *v, int type_mask, double death_cost);
int mgo_vec_set_bonus(mgo_vec *v, int mode); /* 0 none, 1 ActionBonus, 2 PositionBonus */

/* np_random = Generator(PCG64(SeedSequence(seed[i]))) for env i (gymnasium Env.reset(seed=...)) */
void mgo_vec_seed(mgo_vec *v, const uint64_t *seeds);
/* MiniGridEnv.reset() for every env (RNG stream continues), obs: [n][7][7][3], dir: [n] */
void mgo_vec_reset(mgo_vec *v, uint8_t *obs, int32_t *dir, int n_threads);
/* partial reset (gymnasium >= 1.1 SyncVectorEnv.reset(options={"reset_mask": mask})): envs with mask[i] != 0 only */
void mgo_vec_seed_masked(mgo_vec *v, const uint8_t *mask, const uint64_t *seeds);
void mgo_vec_reset_masked(mgo_vec *v, const uint8_t *mask, uint8_t *obs, int32_t *dir);
/* one lockstep step with SyncVectorEnv autoreset semantics. returns 0, or -1 on an invalid action
 * (the reference raises ValueError, minigrid_env.py:584-585) */
int mgo_vec_step(mgo_vec *v, const int32_t *actions, uint8_t *obs, int32_t *dir, double *reward,
                 uint8_t *terminated, uint8_t *truncated, int autoreset_mode, int n_threads);
/* FullyObsWrapper.observation: [n][W][H][3] */
void mgo_vec_full_obs(mgo_vec *v, uint8_t *out);
/* ViewSizeWrapper.observation: gen_obs with agent_view_size V (odd, 3..15): [n][V][V][3]; -1 if V is unsupported */
int mgo_vec_gen_obs_view(mgo_vec *v, int V, uint8_t *obs);
/* SymbolicObsWrapper.observation: int64 [n][W][H][3] = (x, y, type or -1), agent cell type 10 */
void mgo_vec_symbolic_obs(mgo_vec *v, int64_t *out);
/* gen_obs() of the current state, no transition */
void mgo_vec_gen_obs(mgo_vec *v, uint8_t *obs, int32_t *dir);

/* state exchange: grid = Grid.encode() [n][W][H][3]; agent = [n][6] {x, y, dir, carry_type (-1 none),
 * carry_color, step_count}; rng = [n][6] {state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger};
 * pending = [n] autoreset flag (NEXT_STEP). Any pointer may be NULL. */
void mgo_vec_get_state(mgo_vec *v, uint8_t *grid, int32_t *agent, uint64_t *rng, uint8_t *pending);
void mgo_vec_set_state(mgo_vec *v, const uint8_t *grid, const int32_t *agent, const uint64_t *rng,
                       const uint8_t *pending);

/* RNG hooks (pin the numpy restatement): draw from env i's generator */
int64_t mgo_rng_integers(mgo_vec *v, int i, int64_t low, int64_t high);
void mgo_rng_shuffle_perm(mgo_vec *v, int i, int32_t *perm, int n); /* shuffles perm in place */
uint32_t mgo_rng_next32(mgo_vec *v, int i);

/* single-env loop used by bench.py's cpu_baseline: runs n_steps lockstep steps of the whole vector with
 * actions[t][n] and returns elapsed seconds (obs etc. written to internal scratch) */
double mgo_vec_rollout(mgo_vec *v, const int32_t *actions, int n_steps, int autoreset_mode,
                       int n_threads, uint64_t *checksum_out);

int mgo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
