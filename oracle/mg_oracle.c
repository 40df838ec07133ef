/*
 * mg_oracle.c — CPU restatement (plain C) of the reference's MiniGridEnv.step/reset/gen_obs.
 *
 * TEST INFRASTRUCTURE ONLY (see mg_oracle.h). It deliberately keeps the reference's own structure —
 * an object grid, Grid.slice, (dir+1) x Grid.rotate_left, Grid.process_vis, Grid.encode — rather than
 * the closed-form gather the CUDA kernels use, so that the two implementations are independent.
 * All file:line citations are relative to /root/reference/minigrid/.
 *
 * Third-party arithmetic on the path (not under /root/reference, restated from the published
 * algorithms and pinned by tests against numpy 2.3 itself and the reference doctest
 * wrappers.py:26-41): numpy.random.SeedSequence, PCG64 (XSL-RR 128/64), Generator.integers
 * (buffered 32-bit Lemire), Generator.shuffle (masked rejection), Generator.choice.
 */
#include "mg_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <unistd.h>
#include <math.h>

typedef unsigned __int128 u128;

/* ---- core/constants.py:25-46 ---- */
enum { T_UNSEEN = 0, T_EMPTY = 1, T_WALL = 2, T_FLOOR = 3, T_DOOR = 4, T_KEY = 5, T_BALL = 6, T_BOX = 7,
       T_GOAL = 8, T_LAVA = 9, T_AGENT = 10 };
enum { C_RED = 0, C_GREEN = 1, C_BLUE = 2, C_PURPLE = 3, C_YELLOW = 4, C_GREY = 5 };
enum { S_OPEN = 0, S_CLOSED = 1, S_LOCKED = 2 };
/* core/actions.py:7-20 */
enum { A_LEFT = 0, A_RIGHT = 1, A_FORWARD = 2, A_PICKUP = 3, A_DROP = 4, A_TOGGLE = 5, A_DONE = 6 };
/* core/constants.py:49-58 DIR_TO_VEC */
static const int DIR_X[4] = {1, 0, -1, 0};
static const int DIR_Y[4] = {0, 1, 0, -1};

#define VIEW 7 /* agent_view_size, minigrid_env.py:42 */

/* A grid slot: the reference stores WorldObj|None (grid.py:35). `None` is type T_EMPTY here, which is
 * also what Grid.encode emits for None (grid.py:258-261); an object is (type, colour, door state).
 * Box.contains (world_object.py:273-293) is None everywhere except ObstructedMaze, whose boxes hide a key:
 * inner = 1 + the colour of the contained Key, 0 = nothing inside. It is not part of encode(). */
typedef struct { uint8_t type, color, state, inner; } cell_t;
static const cell_t CELL_NONE = {T_EMPTY, 0, 0};

static int cell_is_none(cell_t c) { return c.type == T_EMPTY; }
/* world_object.py:45-47,113,128,141,177 */
static int can_overlap(cell_t c) {
  return c.type == T_GOAL || c.type == T_FLOOR || c.type == T_LAVA || (c.type == T_DOOR && c.state == S_OPEN);
}
/* world_object.py:49-51,243,265,277 */
static int can_pickup(cell_t c) { return c.type == T_KEY || c.type == T_BALL || c.type == T_BOX; }
/* world_object.py:57-59,164,181 */
static int see_behind(cell_t c) {
  if (c.type == T_WALL) return 0;
  if (c.type == T_DOOR) return c.state == S_OPEN;
  return 1;
}

/* ---- numpy PCG64 + Generator restatement ---- */
typedef struct { u128 state, inc; int has_uint32; uint32_t uinteger; } pcg64_t;

static const u128 PCG_MULT = ((u128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;

static void pcg_step(pcg64_t *r) { r->state = r->state * PCG_MULT + r->inc; }
static uint64_t rotr64(uint64_t v, unsigned rot) { return (v >> rot) | (v << ((-rot) & 63)); }
static uint64_t pcg_next64(pcg64_t *r) {
  pcg_step(r);
  return rotr64((uint64_t)(r->state >> 64) ^ (uint64_t)r->state, (unsigned)(r->state >> 122));
}
static uint32_t pcg_next32(pcg64_t *r) {
  if (r->has_uint32) { r->has_uint32 = 0; return r->uinteger; }
  uint64_t n = pcg_next64(r);
  r->has_uint32 = 1;
  r->uinteger = (uint32_t)(n >> 32);
  return (uint32_t)n;
}
static void pcg_srandom(pcg64_t *r, u128 initstate, u128 initseq) {
  r->state = 0;
  r->inc = (initseq << 1) | 1;
  pcg_step(r);
  r->state += initstate;
  pcg_step(r);
  r->has_uint32 = 0;
  r->uinteger = 0;
}

/* numpy.random.SeedSequence(entropy=int).generate_state(8, uint32) */
static uint32_t ss_hashmix(uint32_t value, uint32_t *hash_const) {
  value ^= *hash_const;
  *hash_const *= 0x931e8875u;
  value *= *hash_const;
  value ^= value >> 16;
  return value;
}
static uint32_t ss_mix(uint32_t x, uint32_t y) {
  uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
  r ^= r >> 16;
  return r;
}
static void seed_sequence_pcg64(uint64_t seed, pcg64_t *r) {
  uint32_t entropy[2];
  int n_ent = 1;
  entropy[0] = (uint32_t)seed;
  entropy[1] = (uint32_t)(seed >> 32);
  if (entropy[1] != 0) n_ent = 2;
  uint32_t pool[4];
  uint32_t hc = 0x43b0d7e5u;
  for (int i = 0; i < 4; i++) pool[i] = ss_hashmix(i < n_ent ? entropy[i] : 0u, &hc);
  for (int s = 0; s < 4; s++)
    for (int d = 0; d < 4; d++)
      if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], &hc));
  /* generate_state(4, uint64) == 8 uint32 words, little-endian pairs */
  uint32_t out[8];
  uint32_t hb = 0x8b51f9ddu;
  for (int i = 0; i < 8; i++) {
    uint32_t v = pool[i & 3];
    v ^= hb;
    hb *= 0x58f38dedu;
    v *= hb;
    v ^= v >> 16;
    out[i] = v;
  }
  uint64_t w[4];
  for (int i = 0; i < 4; i++) w[i] = (uint64_t)out[2 * i] | ((uint64_t)out[2 * i + 1] << 32);
  /* pcg64_set_seed: state = (w0 high, w1 low), inc = (w2 high, w3 low) */
  pcg_srandom(r, ((u128)w[0] << 64) | w[1], ((u128)w[2] << 64) | w[3]);
}

/* Generator.integers(low, high) for ranges < 2^32: zero draws when the range is one value */
static int64_t rng_integers(pcg64_t *r, int64_t low, int64_t high) {
  uint32_t rng = (uint32_t)(high - 1 - low);
  if (rng == 0) return low;
  uint32_t rng_excl = rng + 1;
  uint64_t m = (uint64_t)pcg_next32(r) * rng_excl;
  uint32_t leftover = (uint32_t)m;
  if (leftover < rng_excl) {
    uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
    while (leftover < threshold) {
      m = (uint64_t)pcg_next32(r) * rng_excl;
      leftover = (uint32_t)m;
    }
  }
  return low + (int64_t)(m >> 32);
}
/* random_interval(max): masked rejection, used by Generator.shuffle on a Python list */
static uint32_t rng_interval(pcg64_t *r, uint32_t max) {
  if (max == 0) return 0;
  uint32_t mask = max;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t v;
  while ((v = (pcg_next32(r) & mask)) > max) {}
  return v;
}

/* ---- Grid (core/grid.py) ---- */
typedef struct { int width, height; cell_t *cells; } grid_t;

static cell_t grid_get(const grid_t *g, int i, int j) { return g->cells[j * g->width + i]; } /* grid.py:74-78 */
static void grid_set(grid_t *g, int i, int j, cell_t v) { g->cells[j * g->width + i] = v; }   /* grid.py:65-72 */
static const cell_t WALL_GREY = {T_WALL, C_GREY, 0};
static void grid_horz_wall(grid_t *g, int x, int y, int length, cell_t obj) { /* grid.py:80-91 */
  if (length < 0) length = g->width - x;
  for (int i = 0; i < length; i++) grid_set(g, x + i, y, obj);
}
static void grid_vert_wall(grid_t *g, int x, int y, int length, cell_t obj) { /* grid.py:93-104 */
  if (length < 0) length = g->height - y;
  for (int j = 0; j < length; j++) grid_set(g, x, y + j, obj);
}
static void grid_wall_rect(grid_t *g, int x, int y, int w, int h) { /* grid.py:106-110 */
  grid_horz_wall(g, x, y, w, WALL_GREY);
  grid_horz_wall(g, x, y + h - 1, w, WALL_GREY);
  grid_vert_wall(g, x, y, h, WALL_GREY);
  grid_vert_wall(g, x + w - 1, y, h, WALL_GREY);
}

typedef struct { cell_t c[VIEW * VIEW]; } view_t; /* a 7x7 Grid, row-major j*7+i */
static cell_t view_get(const view_t *v, int i, int j) { return v->c[j * VIEW + i]; }
static void view_set(view_t *v, int i, int j, cell_t x) { v->c[j * VIEW + i] = x; }

/* grid.py:124-143 */
static void grid_slice(const grid_t *g, int topX, int topY, view_t *out) {
  for (int j = 0; j < VIEW; j++)
    for (int i = 0; i < VIEW; i++) {
      int x = topX + i, y = topY + j;
      cell_t v = (x >= 0 && x < g->width && y >= 0 && y < g->height) ? grid_get(g, x, y) : WALL_GREY;
      view_set(out, i, j, v);
    }
}
/* grid.py:110-122 (square view: width == height == 7) */
static void view_rotate_left(const view_t *in, view_t *out) {
  for (int i = 0; i < VIEW; i++)
    for (int j = 0; j < VIEW; j++) view_set(out, j, VIEW - 1 - i, view_get(in, i, j));
}
/* grid.py:291-328; mask is [i][j] */
static void view_process_vis(view_t *g, int ax, int ay, uint8_t mask[VIEW][VIEW]) {
  memset(mask, 0, VIEW * VIEW);
  mask[ax][ay] = 1;
  for (int j = VIEW - 1; j >= 0; j--) {
    for (int i = 0; i < VIEW - 1; i++) {
      if (!mask[i][j]) continue;
      cell_t c = view_get(g, i, j);
      if (!cell_is_none(c) && !see_behind(c)) continue;
      mask[i + 1][j] = 1;
      if (j > 0) { mask[i + 1][j - 1] = 1; mask[i][j - 1] = 1; }
    }
    for (int i = VIEW - 1; i >= 1; i--) {
      if (!mask[i][j]) continue;
      cell_t c = view_get(g, i, j);
      if (!cell_is_none(c) && !see_behind(c)) continue;
      mask[i - 1][j] = 1;
      if (j > 0) { mask[i - 1][j - 1] = 1; mask[i][j - 1] = 1; }
    }
  }
  for (int j = 0; j < VIEW; j++)
    for (int i = 0; i < VIEW; i++)
      if (!mask[i][j]) view_set(g, i, j, CELL_NONE);
}
/* grid.py:244-268 + world_object.py:65-67,196-212; out is [i][j][3] C-order */
static void encode_cell(cell_t c, uint8_t *o) {
  if (cell_is_none(c)) { o[0] = T_EMPTY; o[1] = 0; o[2] = 0; }
  else { o[0] = c.type; o[1] = c.color; o[2] = (c.type == T_DOOR) ? c.state : 0; }
}

/* ---- MiniGridEnv (minigrid_env.py) ---- */
typedef struct {
  grid_t grid;
  int agent_x, agent_y, agent_dir;
  int carrying;      /* bool */
  cell_t carry;      /* valid iff carrying */
  int step_count;
  pcg64_t rng;
  uint8_t pending_reset; /* SyncVectorEnv._autoreset_envs[i] */
  int target_x, target_y; /* kinds with a step post-filter (GoToDoor: target_pos; Fetch: targetType, targetColor) */
  int aux[4];             /* PutNear: move_type, moveColor; Memory: failure_pos (target = success_pos) */
  int n_obst, obst_x[8], obst_y[8]; /* Dynamic-Obstacles: self.obstacles[i].cur_pos, in list order */
} env_t;

struct mgo_vec {
  int kind, width, height, max_steps, see_through;
  int32_t params[8];
  int n;
  env_t *envs;
  cell_t *arena;
  /* scratch for rollout */
  uint8_t *s_obs; int32_t *s_dir; double *s_rew; uint8_t *s_term, *s_trunc;
  /* reward wrappers around each env, as a SyncVectorEnv of wrapped envs has them (wrappers.py:68-184, 809-882) */
  int no_death_mask; double death_cost; /* NoDeath: bit t set = OBJECT_TO_IDX type t is in no_death_types */
  int bonus_mode;                        /* 0 none, 1 ActionBonus, 2 PositionBonus (outermost wrapper) */
  uint32_t *counts;                      /* [n][W*H*4*7] / [n][W*H]: the wrappers' self.counts dicts, dense */
};

static int64_t rand_int(env_t *e, int64_t lo, int64_t hi) { return rng_integers(&e->rng, lo, hi); } /* :247-252 */

/* minigrid_env.py:313-372 (reject_fn=None, max_tries=inf); obj == NULL places nothing (place_agent) */
static void place_obj(env_t *e, const cell_t *obj, int top_x, int top_y, int size_w, int size_h, int *px, int *py) {
  if (top_x < 0) top_x = 0;
  if (top_y < 0) top_y = 0;
  int hi_x = top_x + size_w < e->grid.width ? top_x + size_w : e->grid.width;
  int hi_y = top_y + size_h < e->grid.height ? top_y + size_h : e->grid.height;
  for (;;) {
    int x = (int)rand_int(e, top_x, hi_x);
    int y = (int)rand_int(e, top_y, hi_y);
    if (!cell_is_none(grid_get(&e->grid, x, y))) continue;
    if (x == e->agent_x && y == e->agent_y) continue;
    *px = x; *py = y;
    break;
  }
  if (obj) grid_set(&e->grid, *px, *py, *obj);
}
/* minigrid_env.py:383-397 */
static void place_agent(env_t *e, int top_x, int top_y, int size_w, int size_h) {
  e->agent_x = -1; e->agent_y = -1;
  int x, y;
  place_obj(e, NULL, top_x, top_y, size_w, size_h, &x, &y);
  e->agent_x = x; e->agent_y = y;
  e->agent_dir = (int)rand_int(e, 0, 4);
}

static void grid_clear(grid_t *g) { for (int k = 0; k < g->width * g->height; k++) g->cells[k] = CELL_NONE; }

/* envs/empty.py:97-114 */
static void gen_empty(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height;
  grid_clear(&e->grid);
  grid_wall_rect(&e->grid, 0, 0, W, H);
  cell_t goal = {T_GOAL, C_GREEN, 0};
  grid_set(&e->grid, W - 2, H - 2, goal);
  if (!v->params[0]) { e->agent_x = v->params[1]; e->agent_y = v->params[2]; e->agent_dir = v->params[3]; }
  else place_agent(e, 0, 0, W, H);
}
/* envs/doorkey.py:74-99 */
static void gen_doorkey(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height;
  grid_clear(&e->grid);
  grid_wall_rect(&e->grid, 0, 0, W, H);
  cell_t goal = {T_GOAL, C_GREEN, 0};
  grid_set(&e->grid, W - 2, H - 2, goal);
  int split = (int)rand_int(e, 2, W - 2);
  grid_vert_wall(&e->grid, split, 0, -1, WALL_GREY);
  place_agent(e, 0, 0, split, H);
  int door_y = (int)rand_int(e, 1, H - 2);
  cell_t door = {T_DOOR, C_YELLOW, S_LOCKED};
  grid_set(&e->grid, split, door_y, door);
  cell_t key = {T_KEY, C_YELLOW, 0};
  int kx, ky;
  place_obj(e, &key, 0, 0, split, H, &kx, &ky);
}
/* envs/crossing.py:131-188 */
static void gen_crossing(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height;
  int num_crossings = v->params[0];
  cell_t obstacle = {(uint8_t)v->params[1], (uint8_t)(v->params[1] == T_LAVA ? C_RED : C_GREY), 0};
  grid_clear(&e->grid);
  grid_wall_rect(&e->grid, 0, 0, W, H);
  e->agent_x = 1; e->agent_y = 1; e->agent_dir = 0;
  cell_t goal = {T_GOAL, C_GREEN, 0};
  grid_set(&e->grid, W - 2, H - 2, goal);
  /* rivers = [(v, i) for i in range(2, height-2, 2)] + [(h, j) for j in range(2, width-2, 2)] */
  int rdir[64], rpos[64], n = 0; /* dir: 0 = v, 1 = h */
  for (int i = 2; i < H - 2; i += 2) { rdir[n] = 0; rpos[n] = i; n++; }
  for (int j = 2; j < W - 2; j += 2) { rdir[n] = 1; rpos[n] = j; n++; }
  for (int i = n - 1; i >= 1; i--) { /* np_random.shuffle(list) */
    int j = (int)rng_interval(&e->rng, (uint32_t)i);
    int td = rdir[i], tp = rpos[i]; rdir[i] = rdir[j]; rpos[i] = rpos[j]; rdir[j] = td; rpos[j] = tp;
  }
  if (num_crossings < n) n = num_crossings;
  int rivers_v[64], nv = 0, rivers_h[64], nh = 0;
  for (int k = 0; k < n; k++) { if (rdir[k] == 0) rivers_v[nv++] = rpos[k]; else rivers_h[nh++] = rpos[k]; }
  /* sorted() */
  for (int a = 1; a < nv; a++) { int t = rivers_v[a], b = a; while (b > 0 && rivers_v[b - 1] > t) { rivers_v[b] = rivers_v[b - 1]; b--; } rivers_v[b] = t; }
  for (int a = 1; a < nh; a++) { int t = rivers_h[a], b = a; while (b > 0 && rivers_h[b - 1] > t) { rivers_h[b] = rivers_h[b - 1]; b--; } rivers_h[b] = t; }
  /* itt.product(range(1, width-1), rivers_h) then itt.product(rivers_v, range(1, height-1)) */
  for (int i = 1; i < W - 1; i++) for (int k = 0; k < nh; k++) grid_set(&e->grid, i, rivers_h[k], obstacle);
  for (int k = 0; k < nv; k++) for (int j = 1; j < H - 1; j++) grid_set(&e->grid, rivers_v[k], j, obstacle);
  /* path = [h]*len(rivers_v) + [v]*len(rivers_h); shuffle */
  int path[128], np_ = 0;
  for (int k = 0; k < nv; k++) path[np_++] = 1;
  for (int k = 0; k < nh; k++) path[np_++] = 0;
  for (int i = np_ - 1; i >= 1; i--) {
    int j = (int)rng_interval(&e->rng, (uint32_t)i);
    int t = path[i]; path[i] = path[j]; path[j] = t;
  }
  int limits_v[66], limits_h[66];
  limits_v[0] = 0; for (int k = 0; k < nv; k++) limits_v[k + 1] = rivers_v[k]; limits_v[nv + 1] = H - 1;
  limits_h[0] = 0; for (int k = 0; k < nh; k++) limits_h[k + 1] = rivers_h[k]; limits_h[nh + 1] = W - 1;
  int room_i = 0, room_j = 0;
  for (int k = 0; k < np_; k++) {
    int i, j;
    if (path[k] == 1) { /* h */
      i = limits_v[room_i + 1];
      int a = limits_h[room_j] + 1, b = limits_h[room_j + 1];
      j = a + (int)rand_int(e, 0, b - a); /* np_random.choice(range(a, b)) */
      room_i++;
    } else {
      int a = limits_v[room_i] + 1, b = limits_v[room_i + 1];
      i = a + (int)rand_int(e, 0, b - a);
      j = limits_h[room_j + 1];
      room_j++;
    }
    grid_set(&e->grid, i, j, CELL_NONE);
  }
}
/* envs/fourrooms.py:78-126 (agent_pos = goal_pos = None) */
static void gen_fourrooms(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height;
  grid_clear(&e->grid);
  grid_horz_wall(&e->grid, 0, 0, -1, WALL_GREY);
  grid_horz_wall(&e->grid, 0, H - 1, -1, WALL_GREY);
  grid_vert_wall(&e->grid, 0, 0, -1, WALL_GREY);
  grid_vert_wall(&e->grid, W - 1, 0, -1, WALL_GREY);
  int room_w = W / 2, room_h = H / 2;
  for (int j = 0; j < 2; j++)
    for (int i = 0; i < 2; i++) {
      int xL = i * room_w, yT = j * room_h, xR = xL + room_w, yB = yT + room_h;
      if (i + 1 < 2) {
        grid_vert_wall(&e->grid, xR, yT, room_h, WALL_GREY);
        int py = (int)rand_int(e, yT + 1, yB);
        grid_set(&e->grid, xR, py, CELL_NONE);
      }
      if (j + 1 < 2) {
        grid_horz_wall(&e->grid, xL, yB, room_w, WALL_GREY);
        int px = (int)rand_int(e, xL + 1, xR);
        grid_set(&e->grid, px, yB, CELL_NONE);
      }
    }
  place_agent(e, 0, 0, W, H);
  cell_t goal = {T_GOAL, C_GREEN, 0};
  int gx, gy;
  place_obj(e, &goal, 0, 0, W, H, &gx, &gy);
}

/* envs/lavagap.py:100-135 */
static void gen_lavagap(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height;
  cell_t obstacle = {(uint8_t)v->params[0], (uint8_t)(v->params[0] == T_LAVA ? C_RED : C_GREY), 0};
  grid_clear(&e->grid);
  grid_wall_rect(&e->grid, 0, 0, W, H);
  e->agent_x = 1; e->agent_y = 1; e->agent_dir = 0;
  cell_t goal = {T_GOAL, C_GREEN, 0};
  grid_set(&e->grid, W - 2, H - 2, goal);
  int gap_x = (int)rand_int(e, 2, W - 2);
  int gap_y = (int)rand_int(e, 1, H - 1);
  grid_vert_wall(&e->grid, gap_x, 1, H - 2, obstacle);
  grid_set(&e->grid, gap_x, gap_y, CELL_NONE);
}
/* envs/distshift.py:98-120 (agent_start_pos given) */
static void gen_distshift(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height;
  int strip2_row = v->params[0];
  grid_clear(&e->grid);
  grid_wall_rect(&e->grid, 0, 0, W, H);
  cell_t goal = {T_GOAL, C_GREEN, 0};
  grid_set(&e->grid, W - 2, 1, goal);
  cell_t lava = {T_LAVA, C_RED, 0};
  for (int i = 0; i < W - 6; i++) {
    grid_set(&e->grid, 3 + i, 1, lava);
    grid_set(&e->grid, 3 + i, strip2_row, lava);
  }
  e->agent_x = v->params[1]; e->agent_y = v->params[2]; e->agent_dir = v->params[3];
}

/* envs/multiroom.py:117-284 */
typedef struct { int top_x, top_y, size_x, size_y, door_x, door_y; } mroom_t;
typedef struct { mroom_t r[16]; int n; } mroom_list_t;

/* MultiRoomEnv._placeRoom (multiroom.py:196-284), recursion kept as in the reference */
static int mr_place_room(const mgo_vec *v, env_t *e, int num_left, mroom_list_t *list, int min_sz, int max_sz,
                         int entry_wall, int ex, int ey) {
  int size_x = (int)rand_int(e, min_sz, max_sz + 1);
  int size_y = (int)rand_int(e, min_sz, max_sz + 1);
  int top_x, top_y;
  if (list->n == 0) { top_x = ex; top_y = ey; }
  else if (entry_wall == 0) { top_x = ex - size_x + 1; top_y = (int)rand_int(e, ey - size_y + 2, ey); }
  else if (entry_wall == 1) { top_x = (int)rand_int(e, ex - size_x + 2, ex); top_y = ey - size_y + 1; }
  else if (entry_wall == 2) { top_x = ex; top_y = (int)rand_int(e, ey - size_y + 2, ey); }
  else { top_x = (int)rand_int(e, ex - size_x + 2, ex); top_y = ey; }
  if (top_x < 0 || top_y < 0) return 0;
  if (top_x + size_x > v->width || top_y + size_y >= v->height) return 0;
  for (int k = 0; k < list->n - 1; k++) { /* roomList[:-1] */
    const mroom_t *r = &list->r[k];
    int non_overlap = top_x + size_x < r->top_x || r->top_x + r->size_x <= top_x || top_y + size_y < r->top_y ||
                      r->top_y + r->size_y <= top_y;
    if (!non_overlap) return 0;
  }
  mroom_t *nr = &list->r[list->n++];
  nr->top_x = top_x; nr->top_y = top_y; nr->size_x = size_x; nr->size_y = size_y; nr->door_x = ex; nr->door_y = ey;
  if (num_left == 1) return 1;
  for (int i = 0; i < 8; i++) {
    /* wallSet = {0,1,2,3} - {entryDoorWall}; exitDoorWall = _rand_elem(sorted(wallSet)) */
    int walls[3], nw = 0;
    for (int w = 0; w < 4; w++) if (w != entry_wall) walls[nw++] = w;
    int exit_wall = walls[rand_int(e, 0, 3)];
    int next_entry = (exit_wall + 2) % 4;
    int px, py;
    if (exit_wall == 0) { px = top_x + size_x - 1; py = top_y + (int)rand_int(e, 1, size_y - 1); }
    else if (exit_wall == 1) { px = top_x + (int)rand_int(e, 1, size_x - 1); py = top_y + size_y - 1; }
    else if (exit_wall == 2) { px = top_x; py = top_y + (int)rand_int(e, 1, size_y - 1); }
    else { px = top_x + (int)rand_int(e, 1, size_x - 1); py = top_y; }
    if (mr_place_room(v, e, num_left - 1, list, min_sz, max_sz, next_entry, px, py)) break;
  }
  return 1;
}
static void gen_multiroom(const mgo_vec *v, env_t *e) {
  int W = v->width;
  int min_rooms = v->params[0], max_rooms = v->params[1], max_size = v->params[2];
  mroom_list_t best; best.n = 0;
  int num_rooms = (int)rand_int(e, min_rooms, max_rooms + 1);
  while (best.n < num_rooms) {
    mroom_list_t cur; cur.n = 0;
    int ex = (int)rand_int(e, 0, W - 2), ey = (int)rand_int(e, 0, W - 2);
    mr_place_room(v, e, num_rooms, &cur, 4, max_size, 2, ex, ey);
    if (cur.n > best.n) best = cur;
  }
  grid_clear(&e->grid);
  /* COLOR_NAMES sorted: blue green grey purple red yellow -> COLOR_TO_IDX */
  static const int SORTED_COLORS[6] = {C_BLUE, C_GREEN, C_GREY, C_PURPLE, C_RED, C_YELLOW};
  int prev_color = -1;
  for (int idx = 0; idx < best.n; idx++) {
    const mroom_t *r = &best.r[idx];
    for (int i = 0; i < r->size_x; i++) { grid_set(&e->grid, r->top_x + i, r->top_y, WALL_GREY); grid_set(&e->grid, r->top_x + i, r->top_y + r->size_y - 1, WALL_GREY); }
    for (int j = 0; j < r->size_y; j++) { grid_set(&e->grid, r->top_x, r->top_y + j, WALL_GREY); grid_set(&e->grid, r->top_x + r->size_x - 1, r->top_y + j, WALL_GREY); }
    if (idx > 0) {
      int colors[6], nc = 0;
      for (int c = 0; c < 6; c++) if (SORTED_COLORS[c] != prev_color) colors[nc++] = SORTED_COLORS[c];
      int color = colors[rand_int(e, 0, nc)];
      cell_t door = {T_DOOR, (uint8_t)color, S_CLOSED};
      grid_set(&e->grid, r->door_x, r->door_y, door);
      prev_color = color;
    }
  }
  place_agent(e, best.r[0].top_x, best.r[0].top_y, best.r[0].size_x, best.r[0].size_y);
  cell_t goal = {T_GOAL, C_GREEN, 0};
  int gx, gy;
  const mroom_t *last = &best.r[best.n - 1];
  place_obj(e, &goal, last->top_x, last->top_y, last->size_x, last->size_y, &gx, &gy);
}

/* COLOR_NAMES = sorted(COLORS) (constants.py:17): blue green grey purple red yellow, as COLOR_TO_IDX values */
static const int COLOR_NAMES_IDX[6] = {C_BLUE, C_GREEN, C_GREY, C_PURPLE, C_RED, C_YELLOW};

/* envs/lockedroom.py:108-173. _rand_elem(lst) = lst[_rand_int(0, len(lst))] (minigrid_env.py:268-275);
 * LockedRoom.rand_pos = _rand_pos(topX + 1, topX + sizeX - 1, topY + 1, topY + sizeY - 1) (lockedroom.py:18-23) */
static void gen_lockedroom(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height;
  grid_clear(&e->grid);
  for (int i = 0; i < W; i++) { grid_set(&e->grid, i, 0, WALL_GREY); grid_set(&e->grid, i, H - 1, WALL_GREY); }
  for (int j = 0; j < H; j++) { grid_set(&e->grid, 0, j, WALL_GREY); grid_set(&e->grid, W - 1, j, WALL_GREY); }
  int lwall = W / 2 - 2, rwall = W / 2 + 2;
  for (int j = 0; j < H; j++) { grid_set(&e->grid, lwall, j, WALL_GREY); grid_set(&e->grid, rwall, j, WALL_GREY); }
  struct { int top_x, top_y, size_x, size_y, door_x, door_y, color, locked; } rooms[6];
  int nr = 0;
  for (int n = 0; n < 3; n++) {
    int j = n * (H / 3);
    for (int i = 0; i < lwall; i++) grid_set(&e->grid, i, j, WALL_GREY);
    for (int i = rwall; i < W; i++) grid_set(&e->grid, i, j, WALL_GREY);
    int room_w = lwall + 1, room_h = H / 3 + 1;
    rooms[nr].top_x = 0; rooms[nr].top_y = j; rooms[nr].size_x = room_w; rooms[nr].size_y = room_h;
    rooms[nr].door_x = lwall; rooms[nr].door_y = j + 3; rooms[nr].color = -1; rooms[nr].locked = 0; nr++;
    rooms[nr].top_x = rwall; rooms[nr].top_y = j; rooms[nr].size_x = room_w; rooms[nr].size_y = room_h;
    rooms[nr].door_x = rwall; rooms[nr].door_y = j + 3; rooms[nr].color = -1; rooms[nr].locked = 0; nr++;
  }
  int locked = (int)rand_int(e, 0, nr);
  rooms[locked].locked = 1;
  {
    int gx = (int)rand_int(e, rooms[locked].top_x + 1, rooms[locked].top_x + rooms[locked].size_x - 1);
    int gy = (int)rand_int(e, rooms[locked].top_y + 1, rooms[locked].top_y + rooms[locked].size_y - 1);
    cell_t goal = {T_GOAL, C_GREEN, 0};
    grid_set(&e->grid, gx, gy, goal);
  }
  int left[6], nleft = 6;  /* sorted(colors): the remaining names keep their sorted order */
  for (int c = 0; c < 6; c++) left[c] = COLOR_NAMES_IDX[c];
  for (int k = 0; k < nr; k++) {
    int pick = (int)rand_int(e, 0, nleft);
    rooms[k].color = left[pick];
    for (int c = pick; c + 1 < nleft; c++) left[c] = left[c + 1];
    nleft--;
    cell_t door = {T_DOOR, (uint8_t)rooms[k].color, (uint8_t)(rooms[k].locked ? S_LOCKED : S_CLOSED)};
    grid_set(&e->grid, rooms[k].door_x, rooms[k].door_y, door);
  }
  int key_room;
  do { key_room = (int)rand_int(e, 0, nr); } while (key_room == locked);
  {
    int kx = (int)rand_int(e, rooms[key_room].top_x + 1, rooms[key_room].top_x + rooms[key_room].size_x - 1);
    int ky = (int)rand_int(e, rooms[key_room].top_y + 1, rooms[key_room].top_y + rooms[key_room].size_y - 1);
    cell_t key = {T_KEY, (uint8_t)rooms[locked].color, 0};
    grid_set(&e->grid, kx, ky, key);
  }
  place_agent(e, lwall, 0, rwall - lwall, H);
}

/* envs/playground.py:33-90 */
static void gen_playground(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height;
  grid_clear(&e->grid);
  grid_horz_wall(&e->grid, 0, 0, -1, WALL_GREY);
  grid_horz_wall(&e->grid, 0, H - 1, -1, WALL_GREY);
  grid_vert_wall(&e->grid, 0, 0, -1, WALL_GREY);
  grid_vert_wall(&e->grid, W - 1, 0, -1, WALL_GREY);
  int room_w = W / 3, room_h = H / 3;
  for (int j = 0; j < 3; j++) {
    for (int i = 0; i < 3; i++) {
      int xl = i * room_w, yt = j * room_h, xr = xl + room_w, yb = yt + room_h;
      if (i + 1 < 3) {
        grid_vert_wall(&e->grid, xr, yt, room_h, WALL_GREY);
        int py = (int)rand_int(e, yt + 1, yb - 1);
        int color = COLOR_NAMES_IDX[rand_int(e, 0, 6)];
        cell_t door = {T_DOOR, (uint8_t)color, S_CLOSED};
        grid_set(&e->grid, xr, py, door);
      }
      if (j + 1 < 3) {
        grid_horz_wall(&e->grid, xl, yb, room_w, WALL_GREY);
        int px = (int)rand_int(e, xl + 1, xr - 1);
        int color = COLOR_NAMES_IDX[rand_int(e, 0, 6)];
        cell_t door = {T_DOOR, (uint8_t)color, S_CLOSED};
        grid_set(&e->grid, px, yb, door);
      }
    }
  }
  place_agent(e, 0, 0, W, H);
  static const int TYPES[3] = {T_KEY, T_BALL, T_BOX};
  for (int k = 0; k < 12; k++) {
    int type = TYPES[rand_int(e, 0, 3)];
    int color = COLOR_NAMES_IDX[rand_int(e, 0, 6)];
    cell_t obj = {(uint8_t)type, (uint8_t)color, 0};
    int ox, oy;
    place_obj(e, &obj, 0, 0, W, H, &ox, &oy);
  }
}

/* envs/gotodoor.py:88-128: a room of random size in the top-left corner, one door per wall, distinct colours */
static void gen_gotodoor(const mgo_vec *v, env_t *e) {
  grid_clear(&e->grid);
  int width = (int)rand_int(e, 5, v->width + 1);
  int height = (int)rand_int(e, 5, v->height + 1);
  grid_wall_rect(&e->grid, 0, 0, width, height);
  int door_x[4], door_y[4], door_color[4], n_colors = 0;
  door_x[0] = (int)rand_int(e, 2, width - 2); door_y[0] = 0;
  door_x[1] = (int)rand_int(e, 2, width - 2); door_y[1] = height - 1;
  door_x[2] = 0; door_y[2] = (int)rand_int(e, 2, height - 2);
  door_x[3] = width - 1; door_y[3] = (int)rand_int(e, 2, height - 2);
  while (n_colors < 4) {
    int color = COLOR_NAMES_IDX[rand_int(e, 0, 6)];
    int dup = 0;
    for (int k = 0; k < n_colors; k++) dup |= door_color[k] == color;
    if (dup) continue;
    door_color[n_colors++] = color;
  }
  for (int k = 0; k < 4; k++) {
    cell_t door = {T_DOOR, (uint8_t)door_color[k], S_CLOSED};
    grid_set(&e->grid, door_x[k], door_y[k], door);
  }
  place_agent(e, 0, 0, width, height);
  int idx = (int)rand_int(e, 0, 4);
  e->target_x = door_x[idx]; e->target_y = door_y[idx];
}

/* envs/fetch.py:118-160: numObjs random keys / balls, then the agent, a target among the objects, and one draw for
 * the wording of the mission (consumed, the string itself is not modelled) */
static void gen_fetch(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height, num_objs = v->params[0];
  grid_clear(&e->grid);
  grid_horz_wall(&e->grid, 0, 0, -1, WALL_GREY);
  grid_horz_wall(&e->grid, 0, H - 1, -1, WALL_GREY);
  grid_vert_wall(&e->grid, 0, 0, -1, WALL_GREY);
  grid_vert_wall(&e->grid, W - 1, 0, -1, WALL_GREY);
  cell_t objs[16];
  for (int k = 0; k < num_objs; k++) {
    int type = rand_int(e, 0, 2) == 0 ? T_KEY : T_BALL;
    int color = COLOR_NAMES_IDX[rand_int(e, 0, 6)];
    cell_t obj = {(uint8_t)type, (uint8_t)color, 0};
    int ox, oy;
    place_obj(e, &obj, 0, 0, W, H, &ox, &oy);
    objs[k] = obj;
  }
  place_agent(e, 0, 0, W, H);
  cell_t target = objs[rand_int(e, 0, num_objs)];
  e->target_x = target.type; e->target_y = target.color;
  (void)rand_int(e, 0, 5);
}

/* place_obj with max_tries (minigrid_env.py:340-343: `if num_tries > max_tries: raise`, i.e. max_tries + 1 attempts).
 * Returns 0 when the reference would raise RecursionError. */
static int place_obj_tries(env_t *e, const cell_t *obj, int top_x, int top_y, int size_w, int size_h, int max_tries, int *px, int *py) {
  if (top_x < 0) top_x = 0;
  if (top_y < 0) top_y = 0;
  int hi_x = top_x + size_w < e->grid.width ? top_x + size_w : e->grid.width;
  int hi_y = top_y + size_h < e->grid.height ? top_y + size_h : e->grid.height;
  for (int tries = 0; tries <= max_tries; tries++) {
    int x = (int)rand_int(e, top_x, hi_x);
    int y = (int)rand_int(e, top_y, hi_y);
    if (!cell_is_none(grid_get(&e->grid, x, y))) continue;
    if (x == e->agent_x && y == e->agent_y) continue;
    grid_set(&e->grid, x, y, *obj);
    *px = x; *py = y;
    return 1;
  }
  return 0;
}
static const cell_t BALL_BLUE = {T_BALL, C_BLUE, 0}; /* Ball() defaults to blue, world_object.py:254 */

/* envs/dynamicobstacles.py:107-133 */
static void gen_dynobstacles(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height;
  grid_clear(&e->grid);
  grid_wall_rect(&e->grid, 0, 0, W, H);
  cell_t goal = {T_GOAL, C_GREEN, 0};
  grid_set(&e->grid, W - 2, H - 2, goal);
  if (!v->params[1]) { e->agent_x = v->params[2]; e->agent_y = v->params[3]; e->agent_dir = v->params[4]; }
  else place_agent(e, 0, 0, W, H);
  e->n_obst = v->params[0];
  for (int i = 0; i < e->n_obst; i++)
    place_obj_tries(e, &BALL_BLUE, 0, 0, W, H, 100, &e->obst_x[i], &e->obst_y[i]);
}

/* ---- core/roomgrid.py: RoomGrid and the envs built on it (second half) ----
 * params: {variant, room_size, num_rows, num_cols}; variant 0 Unlock (envs/unlock.py), 1 UnlockPickup
 * (unlockpickup.py), 2 BlockedUnlockPickup (blockedunlockpickup.py), 3 KeyCorridor (keycorridor.py, obj_type "ball") */
enum { RG_UNLOCK = 0, RG_UNLOCKPICKUP = 1, RG_BLOCKEDUNLOCKPICKUP = 2, RG_KEYCORRIDOR = 3,
       /* envs/obstructedmaze.py, obstructedmaze_v1.py: params {variant, room_size, num_rows, num_cols, key_in_box, blocked,
        * agent_room_i | agent_room_j << 4, num_quarters (Full only)} */
       RG_OBSTRUCTED_1D = 4 /* ObstructedMaze_1Dlhb */, RG_OBSTRUCTED_FULL = 5 /* ObstructedMaze_Full */,
       RG_OBSTRUCTED_FULL_V1 = 6 /* ObstructedMaze_Full_V1 */ };
typedef struct {
  int top_x, top_y;          /* Room.top; Room.size = (room_size, room_size) */
  int door_x[4], door_y[4];  /* Room.door_pos, order right, down, left, up; -1 = None */
  int doors[4];              /* Room.doors: 0 None, 1 a Door, 2 True (wall removed) */
  int locked;                /* Room.locked */
} rg_room_t;
typedef struct { int S, rows, cols; rg_room_t r[9]; } rg_t; /* room (i, j) = r[j * cols + i] */

static int rg_neighbor(const rg_t *g, int i, int j, int k, int *ni, int *nj) { /* Room.neighbors, roomgrid.py:157-168 */
  static const int DI[4] = {1, 0, -1, 0}, DJ[4] = {0, 1, 0, -1};
  *ni = i + DI[k]; *nj = j + DJ[k];
  return *ni >= 0 && *ni < g->cols && *nj >= 0 && *nj < g->rows;
}
/* RoomGrid._gen_grid, roomgrid.py:123-177 */
static void rg_gen_base(const mgo_vec *v, env_t *e, rg_t *g) {
  g->S = v->params[1]; g->rows = v->params[2]; g->cols = v->params[3];
  const int S = g->S;
  grid_clear(&e->grid);
  for (int j = 0; j < g->rows; j++)
    for (int i = 0; i < g->cols; i++) {
      rg_room_t *r = &g->r[j * g->cols + i];
      memset(r, 0, sizeof(*r));
      r->top_x = i * (S - 1); r->top_y = j * (S - 1);
      for (int k = 0; k < 4; k++) { r->door_x[k] = -1; r->door_y[k] = -1; }
      grid_wall_rect(&e->grid, r->top_x, r->top_y, S, S);
    }
  for (int j = 0; j < g->rows; j++)
    for (int i = 0; i < g->cols; i++) {
      rg_room_t *r = &g->r[j * g->cols + i];
      const int x_l = r->top_x + 1, y_l = r->top_y + 1, x_m = r->top_x + S - 1, y_m = r->top_y + S - 1;
      if (i < g->cols - 1) { r->door_x[0] = x_m; r->door_y[0] = (int)rand_int(e, y_l, y_m); }
      if (j < g->rows - 1) { r->door_x[1] = (int)rand_int(e, x_l, x_m); r->door_y[1] = y_m; }
      if (i > 0) { const rg_room_t *n = &g->r[j * g->cols + i - 1]; r->door_x[2] = n->door_x[0]; r->door_y[2] = n->door_y[0]; }
      if (j > 0) { const rg_room_t *n = &g->r[(j - 1) * g->cols + i]; r->door_x[3] = n->door_x[1]; r->door_y[3] = n->door_y[1]; }
    }
  e->agent_x = (g->cols / 2) * (S - 1) + S / 2; /* "the agent starts in the middle, facing right" */
  e->agent_y = (g->rows / 2) * (S - 1) + S / 2;
  e->agent_dir = 0;
}
/* RoomGrid.add_door, roomgrid.py:226-273; door_idx / color (COLOR_TO_IDX) / locked: -1 = draw it. Returns the colour. */
static int rg_add_door(env_t *e, rg_t *g, int i, int j, int door_idx, int color, int locked, int *px, int *py) {
  rg_room_t *r = &g->r[j * g->cols + i];
  int ni, nj;
  if (door_idx < 0)
    for (;;) {
      door_idx = (int)rand_int(e, 0, 4);
      if (rg_neighbor(g, i, j, door_idx, &ni, &nj) && r->doors[door_idx] == 0) break;
    }
  if (color < 0) color = COLOR_NAMES_IDX[rand_int(e, 0, 6)];
  if (locked < 0) locked = rand_int(e, 0, 2) == 0; /* _rand_bool */
  r->locked = locked;
  cell_t door = {T_DOOR, (uint8_t)color, (uint8_t)(locked ? S_LOCKED : S_CLOSED)};
  grid_set(&e->grid, r->door_x[door_idx], r->door_y[door_idx], door);
  rg_neighbor(g, i, j, door_idx, &ni, &nj);
  r->doors[door_idx] = 1;
  g->r[nj * g->cols + ni].doors[(door_idx + 2) % 4] = 1;
  if (px) { *px = r->door_x[door_idx]; *py = r->door_y[door_idx]; }
  return color;
}
/* RoomGrid.place_in_room, roomgrid.py:179-194: place_obj(reject_fn=reject_next_to, max_tries=1000); the RecursionError
 * after 1001 attempts is not modelled (never seen) */
static void rg_place_in_room(env_t *e, const rg_t *g, int i, int j, cell_t obj, int *px, int *py) {
  const rg_room_t *r = &g->r[j * g->cols + i];
  int hi_x = r->top_x + g->S < e->grid.width ? r->top_x + g->S : e->grid.width;
  int hi_y = r->top_y + g->S < e->grid.height ? r->top_y + g->S : e->grid.height;
  for (;;) {
    int x = (int)rand_int(e, r->top_x, hi_x), y = (int)rand_int(e, r->top_y, hi_y);
    if (!cell_is_none(grid_get(&e->grid, x, y))) continue;
    if (x == e->agent_x && y == e->agent_y) continue;
    if (abs(e->agent_x - x) + abs(e->agent_y - y) < 2) continue; /* reject_next_to, roomgrid.py:11-20 */
    grid_set(&e->grid, x, y, obj);
    if (px) { *px = x; *py = y; }
    return;
  }
}
/* RoomGrid.add_object, roomgrid.py:196-224; kind (T_KEY / T_BALL / T_BOX) / color: -1 = draw it */
static cell_t rg_add_object(env_t *e, const rg_t *g, int i, int j, int kind, int color, int *px, int *py) {
  static const int KINDS[3] = {T_KEY, T_BALL, T_BOX};
  if (kind < 0) kind = KINDS[rand_int(e, 0, 3)];
  if (color < 0) color = COLOR_NAMES_IDX[rand_int(e, 0, 6)];
  cell_t obj = {(uint8_t)kind, (uint8_t)color, 0};
  rg_place_in_room(e, g, i, j, obj, px, py);
  return obj;
}
/* RoomGrid.remove_wall, roomgrid.py:275-311 */
static void rg_remove_wall(env_t *e, rg_t *g, int i, int j, int wall_idx) {
  rg_room_t *r = &g->r[j * g->cols + i];
  const int tx = r->top_x, ty = r->top_y, S = g->S;
  for (int k = 1; k < S - 1; k++) {
    if (wall_idx == 0) grid_set(&e->grid, tx + S - 1, ty + k, CELL_NONE);
    else if (wall_idx == 1) grid_set(&e->grid, tx + k, ty + S - 1, CELL_NONE);
    else if (wall_idx == 2) grid_set(&e->grid, tx, ty + k, CELL_NONE);
    else grid_set(&e->grid, tx + k, ty, CELL_NONE);
  }
  int ni, nj;
  rg_neighbor(g, i, j, wall_idx, &ni, &nj);
  r->doors[wall_idx] = 2;
  g->r[nj * g->cols + ni].doors[(wall_idx + 2) % 4] = 2;
}
/* RoomGrid.place_agent, roomgrid.py:313-335 (i, j given, rand_dir=True): retried until the front cell is None or a wall */
static void rg_place_agent(env_t *e, const rg_t *g, int i, int j) {
  const rg_room_t *r = &g->r[j * g->cols + i];
  for (;;) {
    place_agent(e, r->top_x, r->top_y, g->S, g->S);
    cell_t front = grid_get(&e->grid, e->agent_x + DIR_X[e->agent_dir], e->agent_y + DIR_Y[e->agent_dir]);
    if (cell_is_none(front) || front.type == T_WALL) break;
  }
}
/* RoomGrid.connect_all, roomgrid.py:337-393 (door_colors = COLOR_NAMES) */
static void rg_connect_all(env_t *e, rg_t *g) {
  const int S = g->S;
  const int si = e->agent_x / (S - 1), sj = e->agent_y / (S - 1); /* room_from_pos */
  for (;;) {
    unsigned reach = 0, stack = 1u << (sj * g->cols + si);
    while (stack) {
      const int q = __builtin_ctz(stack);
      stack &= stack - 1;
      if ((reach >> q) & 1u) continue;
      reach |= 1u << q;
      for (int k = 0; k < 4; k++) {
        int ni, nj;
        if (g->r[q].doors[k] && rg_neighbor(g, q % g->cols, q / g->cols, k, &ni, &nj)) stack |= 1u << (nj * g->cols + ni);
      }
    }
    if (__builtin_popcount(reach) == g->rows * g->cols) break;
    const int i = (int)rand_int(e, 0, g->cols), j = (int)rand_int(e, 0, g->rows), k = (int)rand_int(e, 0, 4);
    rg_room_t *r = &g->r[j * g->cols + i];
    if (r->door_x[k] < 0 || r->doors[k]) continue;
    int ni, nj;
    rg_neighbor(g, i, j, k, &ni, &nj);
    if (r->locked || g->r[nj * g->cols + ni].locked) continue;
    const int color = COLOR_NAMES_IDX[rand_int(e, 0, 6)];
    rg_add_door(e, g, i, j, k, color, 0, NULL, NULL);
  }
}
/* ObstructedMazeEnv (obstructedmaze.py:112-176): colours fixed by COLOR_NAMES' order, door colours a random permutation */
static const int OM_BALL_TO_FIND = C_BLUE, OM_BLOCKING_BALL = C_GREEN, OM_BOX = C_GREY; /* COLOR_NAMES[0], [1], [2] */
static void om_add_locked_door(env_t *e, rg_t *g, int i, int j, int door_idx, int color, int blocked) { /* v1 :77-85; also the door half of v0's add_door */
  int dx, dy;
  rg_add_door(e, g, i, j, door_idx, color, 1, &dx, &dy);
  if (blocked) { /* grid.set: whatever was there is overwritten */
    cell_t ball = {T_BALL, (uint8_t)OM_BLOCKING_BALL, 0, 0};
    grid_set(&e->grid, dx - DIR_X[door_idx], dy - DIR_Y[door_idx], ball);
  }
}
static void om_add_key(env_t *e, const rg_t *g, int i, int j, int color, int key_in_box) { /* v1 :87-99; the key half of v0's add_door */
  cell_t obj = {T_KEY, (uint8_t)color, 0, 0};
  if (key_in_box) { obj.type = T_BOX; obj.color = (uint8_t)OM_BOX; obj.inner = (uint8_t)(color + 1); }
  rg_place_in_room(e, g, i, j, obj, NULL, NULL);
}
static void gen_obstructedmaze(const mgo_vec *v, env_t *e, rg_t *g) {
  const int variant = v->params[0], key_in_box = v->params[4], blocked = v->params[5];
  /* door_colors = _rand_subset(COLOR_NAMES, 6): _rand_elem on the shrinking list (minigrid_env.py:277-292) */
  int left[6], colors[6], nleft = 6;
  for (int c = 0; c < 6; c++) left[c] = COLOR_NAMES_IDX[c];
  for (int k = 0; k < 6; k++) {
    const int pick = (int)rand_int(e, 0, nleft);
    colors[k] = left[pick];
    for (int c = pick; c < nleft - 1; c++) left[c] = left[c + 1];
    nleft--;
  }
  cell_t obj;
  if (variant == RG_OBSTRUCTED_1D) { /* obstructedmaze.py:188-203 */
    om_add_locked_door(e, g, 0, 0, 0, colors[0], blocked);
    om_add_key(e, g, 0, 0, colors[0], key_in_box);
    obj = rg_add_object(e, g, 1, 0, T_BALL, OM_BALL_TO_FIND, NULL, NULL);
    rg_place_agent(e, g, 0, 0);
  } else { /* obstructedmaze.py:229-262, obstructedmaze_v1.py:37-75 */
    static const int SIDE[4][2] = {{2, 1}, {1, 2}, {0, 1}, {1, 0}}, CORNER[4][2] = {{2, 0}, {2, 2}, {0, 2}, {0, 0}};
    const int nq = v->params[7];
    for (int i = 0; i < nq; i++) {
      rg_add_door(e, g, 1, 1, i, colors[i], 0, NULL, NULL);
      if (variant == RG_OBSTRUCTED_FULL) {
        for (int k = -1; k <= 1; k += 2) {
          om_add_locked_door(e, g, SIDE[i][0], SIDE[i][1], (i + k + 4) % 4, colors[(i + k + 6) % 6], blocked);
          om_add_key(e, g, SIDE[i][0], SIDE[i][1], colors[(i + k + 6) % 6], key_in_box);
        }
      } else {
        for (int k = -1; k <= 1; k += 2) om_add_locked_door(e, g, SIDE[i][0], SIDE[i][1], (i + k + 4) % 4, colors[(i + k + 6) % 6], blocked);
        for (int k = -1; k <= 1; k += 2) om_add_key(e, g, SIDE[i][0], SIDE[i][1], colors[(i + k + 6) % 6], key_in_box);
      }
    }
    const int corner = (int)rand_int(e, 0, nq);
    obj = rg_add_object(e, g, CORNER[corner][0], CORNER[corner][1], T_BALL, OM_BALL_TO_FIND, NULL, NULL);
    rg_place_agent(e, g, v->params[6] & 15, v->params[6] >> 4);
  }
  e->target_x = obj.type; e->target_y = obj.color;
}

static void gen_roomgrid(const mgo_vec *v, env_t *e) {
  rg_t g;
  rg_gen_base(v, e, &g);
  const int variant = v->params[0];
  if (variant >= RG_OBSTRUCTED_1D) { gen_obstructedmaze(v, e, &g); return; }
  if (variant == RG_KEYCORRIDOR) { /* keycorridor.py:104-128 */
    for (int j = 1; j < g.rows; j++) rg_remove_wall(e, &g, 1, j, 3);
    const int room_idx = (int)rand_int(e, 0, g.rows);
    const int door_color = rg_add_door(e, &g, 2, room_idx, 2, -1, 1, NULL, NULL);
    const cell_t obj = rg_add_object(e, &g, 2, room_idx, T_BALL, -1, NULL, NULL);
    rg_add_object(e, &g, 0, (int)rand_int(e, 0, g.rows), T_KEY, door_color, NULL, NULL);
    rg_place_agent(e, &g, 1, g.rows / 2);
    rg_connect_all(e, &g);
    e->target_x = obj.type; e->target_y = obj.color;
  } else {
    cell_t obj = CELL_NONE;
    int dx = -1, dy = -1;
    if (variant != RG_UNLOCK) obj = rg_add_object(e, &g, 1, 0, T_BOX, -1, NULL, NULL); /* unlockpickup.py:85, blockedunlockpickup.py:93 */
    const int door_color = rg_add_door(e, &g, 0, 0, 0, -1, 1, &dx, &dy);
    if (variant == RG_BLOCKEDUNLOCKPICKUP) { /* :97-98: a ball in front of the door */
      cell_t ball = {T_BALL, (uint8_t)COLOR_NAMES_IDX[rand_int(e, 0, 6)], 0};
      grid_set(&e->grid, dx - 1, dy, ball);
    }
    rg_add_object(e, &g, 0, 0, T_KEY, door_color, NULL, NULL);
    rg_place_agent(e, &g, 0, 0);
    if (variant == RG_UNLOCK) { e->target_x = dx; e->target_y = dy; } /* self.door */
    else { e->target_x = obj.type; e->target_y = obj.color; }          /* self.obj */
  }
}

/* envs/putnear.py:99-166: like GoToObject, but no object within one cell of an earlier one (reject_fn), and a second,
 * different object as the target. place_obj order of tests: cell empty, not the agent, then reject_fn (:348-361). */
static void gen_putnear(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height, num_objs = v->params[0];
  grid_clear(&e->grid);
  grid_horz_wall(&e->grid, 0, 0, -1, WALL_GREY);
  grid_horz_wall(&e->grid, 0, H - 1, -1, WALL_GREY);
  grid_vert_wall(&e->grid, 0, 0, -1, WALL_GREY);
  grid_vert_wall(&e->grid, W - 1, 0, -1, WALL_GREY);
  static const int TYPES[3] = {T_KEY, T_BALL, T_BOX};
  int type[16], color[16], px[16], py[16], n = 0;
  while (n < num_objs) {
    int t = TYPES[rand_int(e, 0, 3)];
    int c = COLOR_NAMES_IDX[rand_int(e, 0, 6)];
    int dup = 0;
    for (int k = 0; k < n; k++) dup |= type[k] == t && color[k] == c;
    if (dup) continue;
    for (;;) { /* place_obj(obj, reject_fn=near_obj) over the whole grid */
      int x = (int)rand_int(e, 0, W), y = (int)rand_int(e, 0, H);
      if (!cell_is_none(grid_get(&e->grid, x, y))) continue;
      if (x == e->agent_x && y == e->agent_y) continue;
      int near = 0;
      for (int k = 0; k < n; k++) near |= abs(x - px[k]) <= 1 && abs(y - py[k]) <= 1;
      if (near) continue;
      px[n] = x; py[n] = y;
      break;
    }
    cell_t obj = {(uint8_t)t, (uint8_t)c, 0};
    grid_set(&e->grid, px[n], py[n], obj);
    type[n] = t; color[n] = c; n++;
  }
  place_agent(e, 0, 0, W, H);
  int move = (int)rand_int(e, 0, n), target;
  do { target = (int)rand_int(e, 0, n); } while (target == move);
  e->aux[0] = type[move]; e->aux[1] = color[move];
  e->target_x = px[target]; e->target_y = py[target];
}

/* envs/memory.py:90-150 */
static void gen_memory(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height;
  grid_clear(&e->grid);
  grid_horz_wall(&e->grid, 0, 0, -1, WALL_GREY);
  grid_horz_wall(&e->grid, 0, H - 1, -1, WALL_GREY);
  grid_vert_wall(&e->grid, 0, 0, -1, WALL_GREY);
  grid_vert_wall(&e->grid, W - 1, 0, -1, WALL_GREY);
  int upper = H / 2 - 2, lower = H / 2 + 2;
  int hallway_end = v->params[0] ? (int)rand_int(e, 4, W - 2) : W - 3;
  for (int i = 1; i < 5; i++) { grid_set(&e->grid, i, upper, WALL_GREY); grid_set(&e->grid, i, lower, WALL_GREY); }
  grid_set(&e->grid, 4, upper + 1, WALL_GREY);
  grid_set(&e->grid, 4, lower - 1, WALL_GREY);
  for (int i = 5; i < hallway_end; i++) { grid_set(&e->grid, i, upper + 1, WALL_GREY); grid_set(&e->grid, i, lower - 1, WALL_GREY); }
  for (int j = 0; j < H; j++) {
    if (j != H / 2) grid_set(&e->grid, hallway_end, j, WALL_GREY);
    grid_set(&e->grid, hallway_end + 2, j, WALL_GREY);
  }
  e->agent_x = (int)rand_int(e, 1, hallway_end + 1); e->agent_y = H / 2; e->agent_dir = 0;
  int start_type = rand_int(e, 0, 2) == 0 ? T_KEY : T_BALL;            /* _rand_elem([Key, Ball]) */
  cell_t start_obj = {(uint8_t)start_type, C_GREEN, 0};
  grid_set(&e->grid, 1, H / 2 - 1, start_obj);
  int first_type = rand_int(e, 0, 2) == 0 ? T_BALL : T_KEY;            /* _rand_elem([[Ball, Key], [Key, Ball]]) */
  cell_t o0 = {(uint8_t)first_type, C_GREEN, 0}, o1 = {(uint8_t)(first_type == T_BALL ? T_KEY : T_BALL), C_GREEN, 0};
  int x = hallway_end + 1, y0 = H / 2 - 2, y1 = H / 2 + 2;
  grid_set(&e->grid, x, y0, o0);
  grid_set(&e->grid, x, y1, o1);
  if (start_type == first_type) { e->target_x = x; e->target_y = y0 + 1; e->aux[0] = x; e->aux[1] = y1 - 1; }
  else { e->target_x = x; e->target_y = y1 - 1; e->aux[0] = x; e->aux[1] = y0 + 1; }
}

/* envs/gotoobject.py:92-139: numObjs distinct (type, colour) objects, then the agent, then the target */
static void gen_gotoobject(const mgo_vec *v, env_t *e) {
  int W = v->width, H = v->height, num_objs = v->params[0];
  grid_clear(&e->grid);
  grid_wall_rect(&e->grid, 0, 0, W, H);
  static const int TYPES[3] = {T_KEY, T_BALL, T_BOX};
  int type[16], color[16], px[16], py[16], n = 0;
  while (n < num_objs) {
    int t = TYPES[rand_int(e, 0, 3)];
    int c = COLOR_NAMES_IDX[rand_int(e, 0, 6)];
    int dup = 0;
    for (int k = 0; k < n; k++) dup |= type[k] == t && color[k] == c;
    if (dup) continue;
    cell_t obj = {(uint8_t)t, (uint8_t)c, 0};
    place_obj(e, &obj, 0, 0, W, H, &px[n], &py[n]);
    type[n] = t; color[n] = c; n++;
  }
  place_agent(e, 0, 0, W, H);
  int idx = (int)rand_int(e, 0, n);
  e->target_x = px[idx]; e->target_y = py[idx];
}

/* envs/redbluedoors.py:78-103: size = height, the grid is 2 * size wide; the agent is placed before the doors exist */
static void gen_redbluedoors(const mgo_vec *v, env_t *e) {
  int size = v->height;
  grid_clear(&e->grid);
  grid_wall_rect(&e->grid, 0, 0, 2 * size, size);
  grid_wall_rect(&e->grid, size / 2, 0, size, size);
  place_agent(e, size / 2, 0, size, size);
  int pos = (int)rand_int(e, 1, size - 1);
  cell_t red = {T_DOOR, C_RED, S_CLOSED};
  grid_set(&e->grid, size / 2, pos, red);
  e->target_x = pos;
  pos = (int)rand_int(e, 1, size - 1);
  cell_t blue = {T_DOOR, C_BLUE, S_CLOSED};
  grid_set(&e->grid, size / 2 + size - 1, pos, blue);
  e->target_y = pos;
}

/* minigrid_env.py:119-157 (without the gen_obs at the end) */
static void env_reset(const mgo_vec *v, env_t *e) {
  e->agent_x = -1; e->agent_y = -1; e->agent_dir = -1;
  switch (v->kind) {
    case MGO_EMPTY: gen_empty(v, e); break;
    case MGO_DOORKEY: gen_doorkey(v, e); break;
    case MGO_CROSSING: gen_crossing(v, e); break;
    case MGO_LAVAGAP: gen_lavagap(v, e); break;
    case MGO_DISTSHIFT: gen_distshift(v, e); break;
    case MGO_MULTIROOM: gen_multiroom(v, e); break;
    case MGO_LOCKEDROOM: gen_lockedroom(v, e); break;
    case MGO_PLAYGROUND: gen_playground(v, e); break;
    case MGO_GOTODOOR: gen_gotodoor(v, e); break;
    case MGO_FETCH: gen_fetch(v, e); break;
    case MGO_REDBLUEDOORS: gen_redbluedoors(v, e); break;
    case MGO_GOTOOBJECT: gen_gotoobject(v, e); break;
    case MGO_PUTNEAR: gen_putnear(v, e); break;
    case MGO_MEMORY: gen_memory(v, e); break;
    case MGO_DYNOBSTACLES: gen_dynobstacles(v, e); break;
    case MGO_ROOMGRID: gen_roomgrid(v, e); break;
    default: gen_fourrooms(v, e); break;
  }
  e->carrying = 0;
  e->carry = CELL_NONE;
  e->step_count = 0;
}

/* minigrid_env.py:597-650 */
static void env_gen_obs(const mgo_vec *v, const env_t *e, uint8_t *image, int32_t *direction) {
  int topX, topY; /* get_view_exts :453-484 */
  switch (e->agent_dir) {
    case 0: topX = e->agent_x; topY = e->agent_y - VIEW / 2; break;
    case 1: topX = e->agent_x - VIEW / 2; topY = e->agent_y; break;
    case 2: topX = e->agent_x - VIEW + 1; topY = e->agent_y - VIEW / 2; break;
    default: topX = e->agent_x - VIEW / 2; topY = e->agent_y - VIEW + 1; break;
  }
  view_t a, b;
  grid_slice(&e->grid, topX, topY, &a);
  view_t *cur = &a, *other = &b;
  for (int r = 0; r < e->agent_dir + 1; r++) { view_rotate_left(cur, other); view_t *t = cur; cur = other; other = t; }
  uint8_t mask[VIEW][VIEW];
  if (!v->see_through) view_process_vis(cur, VIEW / 2, VIEW - 1, mask);
  else memset(mask, 1, sizeof(mask));
  view_set(cur, VIEW / 2, VIEW - 1, e->carrying ? e->carry : CELL_NONE); /* :623-630 */
  for (int i = 0; i < VIEW; i++)       /* Grid.encode(vis_mask) grid.py:244-268 */
    for (int j = 0; j < VIEW; j++) {
      uint8_t *o = image + (i * VIEW + j) * 3;
      if (mask[i][j]) encode_cell(view_get(cur, i, j), o);
      else { o[0] = 0; o[1] = 0; o[2] = 0; }
    }
  *direction = e->agent_dir;
}

/* ViewSizeWrapper.observation (wrappers.py:663-673) = gen_obs_grid(agent_view_size) + grid.encode(vis_mask) for an odd
 * view size V >= 3 (minigrid_env.py:453-484, 597-630; grid.py:110-143, 291-328): the same pipeline as env_gen_obs with
 * the 7 replaced by V. image: [V][V][3] */
#define MAX_VIEW 15
static void env_gen_obs_view(const mgo_vec *v, const env_t *e, int V, uint8_t *image) {
  int topX, topY;
  switch (e->agent_dir) {
    case 0: topX = e->agent_x; topY = e->agent_y - V / 2; break;
    case 1: topX = e->agent_x - V / 2; topY = e->agent_y; break;
    case 2: topX = e->agent_x - V + 1; topY = e->agent_y - V / 2; break;
    default: topX = e->agent_x - V / 2; topY = e->agent_y - V + 1; break;
  }
  cell_t a[MAX_VIEW * MAX_VIEW], b[MAX_VIEW * MAX_VIEW]; /* row-major j * V + i */
  for (int j = 0; j < V; j++)
    for (int i = 0; i < V; i++) {
      int x = topX + i, y = topY + j;
      a[j * V + i] = (x >= 0 && x < e->grid.width && y >= 0 && y < e->grid.height) ? grid_get(&e->grid, x, y) : WALL_GREY;
    }
  cell_t *cur = a, *other = b;
  for (int r = 0; r < e->agent_dir + 1; r++) { /* rotate_left: out(j, V - 1 - i) = in(i, j) */
    for (int i = 0; i < V; i++)
      for (int j = 0; j < V; j++) other[(V - 1 - i) * V + j] = cur[j * V + i];
    cell_t *t = cur; cur = other; other = t;
  }
  uint8_t mask[MAX_VIEW][MAX_VIEW];
  if (!v->see_through) {
    memset(mask, 0, sizeof(mask));
    mask[V / 2][V - 1] = 1;
    for (int j = V - 1; j >= 0; j--) {
      for (int i = 0; i < V - 1; i++) {
        if (!mask[i][j]) continue;
        cell_t c = cur[j * V + i];
        if (!cell_is_none(c) && !see_behind(c)) continue;
        mask[i + 1][j] = 1;
        if (j > 0) { mask[i + 1][j - 1] = 1; mask[i][j - 1] = 1; }
      }
      for (int i = V - 1; i >= 1; i--) {
        if (!mask[i][j]) continue;
        cell_t c = cur[j * V + i];
        if (!cell_is_none(c) && !see_behind(c)) continue;
        mask[i - 1][j] = 1;
        if (j > 0) { mask[i - 1][j - 1] = 1; mask[i][j - 1] = 1; }
      }
    }
  } else memset(mask, 1, sizeof(mask));
  cur[(V - 1) * V + V / 2] = e->carrying ? e->carry : CELL_NONE;
  for (int i = 0; i < V; i++)
    for (int j = 0; j < V; j++) {
      uint8_t *o = image + (i * V + j) * 3;
      if (mask[i][j]) encode_cell(cur[j * V + i], o);
      else { o[0] = 0; o[1] = 0; o[2] = 0; }
    }
}

/* minigrid_env.py:240-245; compiled with -ffp-contract=off so no FMA is formed */
static double env_reward(const mgo_vec *v, const env_t *e) {
  volatile double q = (double)e->step_count / (double)v->max_steps;
  volatile double p = 0.9 * q;
  return 1.0 - p;
}

/* minigrid_env.py:525-595 (transition only; the caller generates the obs). returns -1 on bad action */
static int env_step(const mgo_vec *v, env_t *e, int action, double *reward, uint8_t *terminated, uint8_t *truncated) {
  int red_before = 0, blue_before = 0; /* RedBlueDoorEnv.step, redbluedoors.py:105-108 */
  if (v->kind == MGO_REDBLUEDOORS) {
    red_before = grid_get(&e->grid, v->height / 2, e->target_x).state == S_OPEN;
    blue_before = grid_get(&e->grid, v->height / 2 + v->height - 1, e->target_y).state == S_OPEN;
  }
  int not_clear = 0;
  if (v->kind == MGO_DYNOBSTACLES) { /* DynamicObstaclesEnv.step before super().step, dynamicobstacles.py:135-158 */
    if (action >= 3) action = 0; /* action_space = Discrete(forward + 1); the unknown-action error cannot occur */
    cell_t front = grid_get(&e->grid, e->agent_x + DIR_X[e->agent_dir], e->agent_y + DIR_Y[e->agent_dir]);
    not_clear = !cell_is_none(front) && front.type != T_GOAL;
    for (int i = 0; i < e->n_obst; i++) {
      int ox = e->obst_x[i], oy = e->obst_y[i], nx, ny;
      if (place_obj_tries(e, &BALL_BLUE, ox - 1, oy - 1, 3, 3, 100, &nx, &ny)) {
        grid_set(&e->grid, ox, oy, CELL_NONE);
        e->obst_x[i] = nx; e->obst_y[i] = ny;
      }
    }
  }
  const int pre_carrying = e->carrying; /* PutNearEnv.step, putnear.py:168-169 */
  if (v->kind == MGO_MEMORY && action == A_PICKUP) action = A_TOGGLE; /* memory.py:152-154 */
  e->step_count += 1;
  *reward = 0; *terminated = 0; *truncated = 0;
  int fx = e->agent_x + DIR_X[e->agent_dir], fy = e->agent_y + DIR_Y[e->agent_dir];
  cell_t fwd = grid_get(&e->grid, fx, fy);
  int fwd_none = cell_is_none(fwd);
  switch (action) {
    case A_LEFT: e->agent_dir -= 1; if (e->agent_dir < 0) e->agent_dir += 4; break;
    case A_RIGHT: e->agent_dir = (e->agent_dir + 1) % 4; break;
    case A_FORWARD:
      if (fwd_none || can_overlap(fwd)) { e->agent_x = fx; e->agent_y = fy; }
      if (!fwd_none && fwd.type == T_GOAL) { *terminated = 1; *reward = env_reward(v, e); }
      if (!fwd_none && fwd.type == T_LAVA) *terminated = 1;
      break;
    case A_PICKUP:
      if (!fwd_none && can_pickup(fwd) && !e->carrying) {
        e->carrying = 1; e->carry = fwd;
        grid_set(&e->grid, fx, fy, CELL_NONE);
      }
      break;
    case A_DROP:
      if (fwd_none && e->carrying) { grid_set(&e->grid, fx, fy, e->carry); e->carrying = 0; e->carry = CELL_NONE; }
      break;
    case A_TOGGLE:
      if (!fwd_none) {
        if (fwd.type == T_DOOR) { /* Door.toggle world_object.py:184-194 */
          if (fwd.state == S_LOCKED) {
            if (e->carrying && e->carry.type == T_KEY && e->carry.color == fwd.color) { fwd.state = S_OPEN; grid_set(&e->grid, fx, fy, fwd); }
          } else { fwd.state = (fwd.state == S_OPEN) ? S_CLOSED : S_OPEN; grid_set(&e->grid, fx, fy, fwd); }
        } else if (fwd.type == T_BOX) { /* Box.toggle :290-293: the box is replaced by its contents (None, or the hidden key) */
          cell_t inside = CELL_NONE;
          if (fwd.inner) { inside.type = T_KEY; inside.color = (uint8_t)(fwd.inner - 1); inside.state = 0; inside.inner = 0; }
          grid_set(&e->grid, fx, fy, inside);
        }
      }
      break;
    case A_DONE: break;
    default: e->step_count -= 1; return -1;
  }
  if (e->step_count >= v->max_steps) *truncated = 1;
  if (v->kind == MGO_GOTODOOR || v->kind == MGO_GOTOOBJECT) { /* gotodoor.py:130-149, gotoobject.py:141-160 (same filter) */
    if (action == A_TOGGLE) *terminated = 1;
    if (action == A_DONE) {
      int dx = e->agent_x - e->target_x, dy = e->agent_y - e->target_y;
      if ((dx == 0 && (dy == 1 || dy == -1)) || (dy == 0 && (dx == 1 || dx == -1))) *reward = env_reward(v, e);
      *terminated = 1;
    }
  }
  if (v->kind == MGO_DYNOBSTACLES && action == A_FORWARD && not_clear) { *reward = -1.0; *terminated = 1; } /* :162-165 */
  if (v->kind == MGO_ROOMGRID) {
    if (v->params[0] == RG_UNLOCK) { /* unlock.py:88-96: a toggle with self.door open ends the episode */
      if (action == A_TOGGLE) {
        cell_t d = grid_get(&e->grid, e->target_x, e->target_y);
        if (d.type == T_DOOR && d.state == S_OPEN) { *reward = env_reward(v, e); *terminated = 1; }
      }
    } else if (action == A_PICKUP && e->carrying && e->carry.type == e->target_x && e->carry.color == e->target_y) {
      /* "self.carrying == self.obj" (unlockpickup.py:97-105, blockedunlockpickup.py:107-115, keycorridor.py:128-136): an
       * identity test; these generators create exactly one object of self.obj's type, so type + colour decide it */
      *reward = env_reward(v, e); *terminated = 1;
    }
  }
  if (v->kind == MGO_PUTNEAR) { /* putnear.py:171-199 */
    int ox = e->agent_x + DIR_X[e->agent_dir], oy = e->agent_y + DIR_Y[e->agent_dir];
    if (action == A_PICKUP && e->carrying && (e->carry.type != e->aux[0] || e->carry.color != e->aux[1])) *terminated = 1;
    if (action == A_DROP && pre_carrying) {
      /* "self.grid.get(ox, oy) is preCarrying": the drop happened, i.e. nothing is carried any more */
      if (!e->carrying && abs(ox - e->target_x) <= 1 && abs(oy - e->target_y) <= 1) *reward = env_reward(v, e);
      *terminated = 1;
    }
  }
  if (v->kind == MGO_MEMORY) { /* memory.py:156-164 */
    if (e->agent_x == e->target_x && e->agent_y == e->target_y) { *reward = env_reward(v, e); *terminated = 1; }
    if (e->agent_x == e->aux[0] && e->agent_y == e->aux[1]) { *reward = 0.0; *terminated = 1; }
  }
  if (v->kind == MGO_REDBLUEDOORS) { /* redbluedoors.py:110-126 */
    int red_after = grid_get(&e->grid, v->height / 2, e->target_x).state == S_OPEN;
    int blue_after = grid_get(&e->grid, v->height / 2 + v->height - 1, e->target_y).state == S_OPEN;
    if (blue_after) {
      *reward = red_before ? env_reward(v, e) : 0.0;
      *terminated = 1;
    } else if (red_after && blue_before) {
      *reward = 0.0;
      *terminated = 1;
    }
  }
  if (v->kind == MGO_FETCH && e->carrying) { /* FetchEnv.step after super().step, fetch.py:162-175 */
    *reward = (e->carry.type == e->target_x && e->carry.color == e->target_y) ? env_reward(v, e) : 0.0;
    *terminated = 1;
  }
  return 0;
}

/* ---- vector level (gymnasium.vector.SyncVectorEnv semantics, restated) ---- */
mgo_vec *mgo_vec_create(int kind, int width, int height, int max_steps, int see_through_walls,
                        const int32_t *params, int n_params, int n_envs) {
  mgo_vec *v = (mgo_vec *)calloc(1, sizeof(*v));
  v->kind = kind; v->width = width; v->height = height; v->max_steps = max_steps; v->see_through = see_through_walls;
  for (int i = 0; i < 8; i++) v->params[i] = (params && i < n_params) ? params[i] : 0;
  v->n = n_envs;
  v->envs = (env_t *)calloc((size_t)n_envs, sizeof(env_t));
  v->arena = (cell_t *)malloc((size_t)n_envs * width * height * sizeof(cell_t));
  for (int i = 0; i < n_envs; i++) {
    env_t *e = &v->envs[i];
    e->grid.width = width; e->grid.height = height;
    e->grid.cells = v->arena + (size_t)i * width * height;
    grid_clear(&e->grid);
    e->agent_x = e->agent_y = -1; e->agent_dir = -1;
    seed_sequence_pcg64((uint64_t)i, &e->rng);
  }
  return v;
}
/* NoDeath(env, no_death_types, death_cost), wrappers.py:836-850: type_mask bit t = OBJECT_TO_IDX value t; 0 removes it */
int mgo_vec_set_no_death(mgo_vec *v, int type_mask, double death_cost) {
  if (type_mask & (1 << T_GOAL)) return -1; /* assert "goal" not in no_death_types, :845 */
  v->no_death_mask = type_mask; v->death_cost = death_cost;
  return 0;
}
/* ActionBonus(env) (mode 1) / PositionBonus(env) (mode 2), wrappers.py:97-104, 148-157: fresh, empty counts */
int mgo_vec_set_bonus(mgo_vec *v, int mode) {
  free(v->counts); v->counts = NULL; v->bonus_mode = 0;
  if (mode == 0) return 0;
  if (mode != 1 && mode != 2) return -1;
  v->counts = (uint32_t *)calloc((size_t)v->n * v->width * v->height * (mode == 1 ? 28 : 1), sizeof(uint32_t));
  if (!v->counts) return -1;
  v->bonus_mode = mode;
  return 0;
}
void mgo_vec_destroy(mgo_vec *v) {
  if (!v) return;
  free(v->counts);
  free(v->s_obs); free(v->s_dir); free(v->s_rew); free(v->s_term); free(v->s_trunc);
  free(v->arena); free(v->envs); free(v);
}
void mgo_vec_seed(mgo_vec *v, const uint64_t *seeds) {
  for (int i = 0; i < v->n; i++) seed_sequence_pcg64(seeds[i], &v->envs[i].rng);
}
int mgo_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
static int clamp_threads(const mgo_vec *v, int n_threads) {
  if (n_threads <= 0) n_threads = mgo_max_threads();
  if (n_threads > v->n) n_threads = v->n;
  return n_threads < 1 ? 1 : n_threads;
}

/* envs are independent, so a "thread" is just a contiguous slice of the vector (what a farm of
 * SyncVectorEnv worker processes, one per core, amounts to) */
typedef struct {
  mgo_vec *v; int lo, hi; int op; /* 0 reset, 1 step, 2 rollout */
  const int32_t *actions; uint8_t *obs; int32_t *dir; double *reward; uint8_t *term, *trunc;
  int mode, n_steps, bad;
} job_t;

static void reset_range(mgo_vec *v, int lo, int hi, uint8_t *obs, int32_t *dir) {
  for (int i = lo; i < hi; i++) {
    env_t *e = &v->envs[i];
    env_reset(v, e);
    e->pending_reset = 0;
    env_gen_obs(v, e, obs + (size_t)i * 147, dir + i);
  }
}
/* the wrapped env's step(): BonusWrapper(NoDeath(env)).step(action). A wrapper that is not configured is absent. */
static int wrapped_step(mgo_vec *v, int i, int action, double *reward, uint8_t *terminated, uint8_t *truncated) {
  env_t *e = &v->envs[i];
  /* NoDeath.step, wrappers.py:852-882: the cell in front BEFORE the env steps (Dynamic-Obstacles moves its balls inside step) */
  int going_to_death = 0;
  if (v->no_death_mask) {
    int fx = e->agent_x + DIR_X[e->agent_dir], fy = e->agent_y + DIR_Y[e->agent_dir]; /* front_pos; always inside the walls */
    cell_t front = grid_get(&e->grid, fx, fy);
    going_to_death = action == A_FORWARD && !cell_is_none(front) && ((v->no_death_mask >> front.type) & 1);
  }
  int rc = env_step(v, e, action, reward, terminated, truncated);
  if (rc != 0) return rc;
  if (v->no_death_mask) {
    cell_t cur = grid_get(&e->grid, e->agent_x, e->agent_y);
    int in_death = !cell_is_none(cur) && ((v->no_death_mask >> cur.type) & 1);
    if (*terminated && (going_to_death || in_death)) {
      *terminated = 0;
      volatile double r = *reward + v->death_cost; /* reward += self.death_cost */
      *reward = r;
    }
  }
  if (v->bonus_mode) {
    /* ActionBonus.step wrappers.py:106-125: key (agent_pos, agent_dir, action) after the step;
     * PositionBonus.step :163-184: key agent_pos, bonus * self.scale with self.scale = 1 (:157) */
    size_t per = (size_t)v->width * v->height * (v->bonus_mode == 1 ? 28 : 1);
    size_t key = (size_t)e->agent_y * v->width + e->agent_x;
    if (v->bonus_mode == 1) key = (key * 4 + (size_t)e->agent_dir) * 7 + (size_t)action;
    uint32_t c = ++v->counts[(size_t)i * per + key];
    volatile double sq = sqrt((double)c);
    volatile double bonus = 1.0 / sq;
    volatile double r = *reward + bonus;
    *reward = r;
  }
  return 0;
}

static int step_range(mgo_vec *v, int lo, int hi, const int32_t *actions, uint8_t *obs, int32_t *dir,
                      double *reward, uint8_t *terminated, uint8_t *truncated, int mode) {
  int bad = 0;
  for (int i = lo; i < hi; i++) {
    env_t *e = &v->envs[i];
    if (mode == MGO_AUTORESET_NEXT_STEP && e->pending_reset) {
      env_reset(v, e); /* action ignored; unseeded reset: the RNG stream continues (the wrappers' reset() is the env's) */
      reward[i] = 0.0; terminated[i] = 0; truncated[i] = 0;
      e->pending_reset = 0;
    } else {
      if (wrapped_step(v, i, actions[i], &reward[i], &terminated[i], &truncated[i]) != 0) bad = 1;
      int done = terminated[i] | truncated[i];
      if (mode == MGO_AUTORESET_NEXT_STEP) e->pending_reset = (uint8_t)done;
      else if (mode == MGO_AUTORESET_SAME_STEP && done) env_reset(v, e);
    }
    env_gen_obs(v, e, obs + (size_t)i * 147, dir + i);
  }
  return bad;
}
static void *job_main(void *arg) {
  job_t *j = (job_t *)arg;
  if (j->op == 0) reset_range(j->v, j->lo, j->hi, j->obs, j->dir);
  else if (j->op == 1) j->bad = step_range(j->v, j->lo, j->hi, j->actions, j->obs, j->dir, j->reward, j->term, j->trunc, j->mode);
  else
    for (int t = 0; t < j->n_steps; t++)
      j->bad |= step_range(j->v, j->lo, j->hi, j->actions + (size_t)t * j->v->n, j->obs, j->dir, j->reward, j->term, j->trunc, j->mode);
  return NULL;
}
static int run_jobs(job_t proto, int nt) {
  mgo_vec *v = proto.v;
  if (nt == 1) { proto.lo = 0; proto.hi = v->n; job_main(&proto); return proto.bad; }
  job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)nt);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nt);
  for (int k = 0; k < nt; k++) {
    jobs[k] = proto;
    jobs[k].lo = (int)((long long)v->n * k / nt);
    jobs[k].hi = (int)((long long)v->n * (k + 1) / nt);
    pthread_create(&th[k], NULL, job_main, &jobs[k]);
  }
  int bad = 0;
  for (int k = 0; k < nt; k++) { pthread_join(th[k], NULL); bad |= jobs[k].bad; }
  free(jobs); free(th);
  return bad;
}
void mgo_vec_reset(mgo_vec *v, uint8_t *obs, int32_t *dir, int n_threads) {
  job_t j; memset(&j, 0, sizeof(j));
  j.v = v; j.op = 0; j.obs = obs; j.dir = dir;
  run_jobs(j, clamp_threads(v, n_threads));
}
/* gymnasium >= 1.1 SyncVectorEnv.reset(seed=..., options={"reset_mask": mask}) (third-party, not in the reference
 * tree; restated): only the selected sub-environments are seeded and reset, `_autoreset_envs` is cleared for them,
 * the other envs and their slots of the observation buffer are left as they are. */
void mgo_vec_seed_masked(mgo_vec *v, const uint8_t *mask, const uint64_t *seeds) {
  for (int i = 0; i < v->n; i++)
    if (mask[i]) seed_sequence_pcg64(seeds[i], &v->envs[i].rng);
}
void mgo_vec_reset_masked(mgo_vec *v, const uint8_t *mask, uint8_t *obs, int32_t *dir) {
  for (int i = 0; i < v->n; i++)
    if (mask[i]) reset_range(v, i, i + 1, obs, dir);
}
void mgo_vec_gen_obs(mgo_vec *v, uint8_t *obs, int32_t *dir) {
  for (int i = 0; i < v->n; i++) env_gen_obs(v, &v->envs[i], obs + (size_t)i * 147, dir + i);
}
int mgo_vec_step(mgo_vec *v, const int32_t *actions, uint8_t *obs, int32_t *dir, double *reward,
                 uint8_t *terminated, uint8_t *truncated, int mode, int n_threads) {
  job_t j; memset(&j, 0, sizeof(j));
  j.v = v; j.op = 1; j.actions = actions; j.obs = obs; j.dir = dir; j.reward = reward; j.term = terminated; j.trunc = truncated; j.mode = mode;
  return run_jobs(j, clamp_threads(v, n_threads)) ? -1 : 0;
}
/* ViewSizeWrapper: image [n][V][V][3]; returns -1 for an unsupported size */
int mgo_vec_gen_obs_view(mgo_vec *v, int V, uint8_t *obs) {
  if (V < 3 || V > MAX_VIEW || V % 2 == 0) return -1;
  for (int i = 0; i < v->n; i++) env_gen_obs_view(v, &v->envs[i], V, obs + (size_t)i * V * V * 3);
  return 0;
}
/* SymbolicObsWrapper.observation (wrappers.py:762-782): [n][W][H][3] int64 = (x, y, OBJECT_TO_IDX[type] or -1), the
 * agent's cell gets OBJECT_TO_IDX["agent"] */
void mgo_vec_symbolic_obs(mgo_vec *v, int64_t *out) {
  int W = v->width, H = v->height;
  for (int n = 0; n < v->n; n++) {
    env_t *e = &v->envs[n];
    int64_t *o = out + (size_t)n * W * H * 3;
    for (int i = 0; i < W; i++)
      for (int j = 0; j < H; j++) {
        cell_t c = grid_get(&e->grid, i, j);
        int64_t *t = o + (i * H + j) * 3;
        t[0] = i; t[1] = j; t[2] = cell_is_none(c) ? -1 : c.type;
      }
    o[(e->agent_x * H + e->agent_y) * 3 + 2] = T_AGENT;
  }
}
/* wrappers.py:419-426 */
void mgo_vec_full_obs(mgo_vec *v, uint8_t *out) {
  int W = v->width, H = v->height;
  for (int n = 0; n < v->n; n++) {
    env_t *e = &v->envs[n];
    uint8_t *o = out + (size_t)n * W * H * 3;
    for (int i = 0; i < W; i++)
      for (int j = 0; j < H; j++) encode_cell(grid_get(&e->grid, i, j), o + (i * H + j) * 3);
    uint8_t *a = o + (e->agent_x * H + e->agent_y) * 3;
    a[0] = T_AGENT; a[1] = C_RED; a[2] = (uint8_t)e->agent_dir;
  }
}
void mgo_vec_get_state(mgo_vec *v, uint8_t *grid, int32_t *agent, uint64_t *rng, uint8_t *pending) {
  int W = v->width, H = v->height;
  for (int n = 0; n < v->n; n++) {
    env_t *e = &v->envs[n];
    if (grid) {
      uint8_t *o = grid + (size_t)n * W * H * 3;
      for (int i = 0; i < W; i++)
        for (int j = 0; j < H; j++) encode_cell(grid_get(&e->grid, i, j), o + (i * H + j) * 3);
    }
    if (agent) {
      int32_t *a = agent + (size_t)n * 6;
      a[0] = e->agent_x; a[1] = e->agent_y; a[2] = e->agent_dir;
      a[3] = e->carrying ? e->carry.type : -1; a[4] = e->carrying ? e->carry.color : 0; a[5] = e->step_count;
    }
    if (rng) {
      uint64_t *r = rng + (size_t)n * 6;
      r[0] = (uint64_t)(e->rng.state >> 64); r[1] = (uint64_t)e->rng.state;
      r[2] = (uint64_t)(e->rng.inc >> 64); r[3] = (uint64_t)e->rng.inc;
      r[4] = (uint64_t)e->rng.has_uint32; r[5] = e->rng.uinteger;
    }
    if (pending) pending[n] = e->pending_reset;
  }
}
void mgo_vec_set_state(mgo_vec *v, const uint8_t *grid, const int32_t *agent, const uint64_t *rng,
                       const uint8_t *pending) {
  int W = v->width, H = v->height;
  for (int n = 0; n < v->n; n++) {
    env_t *e = &v->envs[n];
    if (grid) {
      const uint8_t *o = grid + (size_t)n * W * H * 3;
      for (int i = 0; i < W; i++)
        for (int j = 0; j < H; j++) {
          const uint8_t *c = o + (i * H + j) * 3;
          cell_t x = {c[0], c[1], c[2]};
          if (x.type == T_EMPTY || x.type == T_UNSEEN || x.type == T_AGENT) x = CELL_NONE; /* WorldObj.decode :77-78 */
          grid_set(&e->grid, i, j, x);
        }
    }
    if (agent) {
      const int32_t *a = agent + (size_t)n * 6;
      e->agent_x = a[0]; e->agent_y = a[1]; e->agent_dir = a[2];
      e->carrying = a[3] >= 0;
      e->carry = CELL_NONE;
      if (e->carrying) { e->carry.type = (uint8_t)a[3]; e->carry.color = (uint8_t)a[4]; e->carry.state = 0; }
      e->step_count = a[5];
    }
    if (rng) {
      const uint64_t *r = rng + (size_t)n * 6;
      e->rng.state = ((u128)r[0] << 64) | r[1];
      e->rng.inc = ((u128)r[2] << 64) | r[3];
      e->rng.has_uint32 = (int)r[4]; e->rng.uinteger = (uint32_t)r[5];
    }
    if (pending) e->pending_reset = pending[n];
  }
}

int64_t mgo_rng_integers(mgo_vec *v, int i, int64_t low, int64_t high) { return rng_integers(&v->envs[i].rng, low, high); }
uint32_t mgo_rng_next32(mgo_vec *v, int i) { return pcg_next32(&v->envs[i].rng); }
void mgo_rng_shuffle_perm(mgo_vec *v, int i, int32_t *perm, int n) {
  for (int k = n - 1; k >= 1; k--) {
    int j = (int)rng_interval(&v->envs[i].rng, (uint32_t)k);
    int32_t t = perm[k]; perm[k] = perm[j]; perm[j] = t;
  }
}

double mgo_vec_rollout(mgo_vec *v, const int32_t *actions, int n_steps, int mode, int n_threads,
                       uint64_t *checksum_out) {
  size_t n = (size_t)v->n;
  if (!v->s_obs) {
    v->s_obs = (uint8_t *)malloc(n * 147); v->s_dir = (int32_t *)malloc(n * 4); v->s_rew = (double *)malloc(n * 8);
    v->s_term = (uint8_t *)malloc(n); v->s_trunc = (uint8_t *)malloc(n);
  }
  job_t j; memset(&j, 0, sizeof(j));
  j.v = v; j.op = 2; j.actions = actions; j.obs = v->s_obs; j.dir = v->s_dir; j.reward = v->s_rew;
  j.term = v->s_term; j.trunc = v->s_trunc; j.mode = mode; j.n_steps = n_steps;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  run_jobs(j, clamp_threads(v, n_threads));
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (checksum_out) {
    uint64_t h = 1469598103934665603ULL;
    for (size_t k = 0; k < n * 147; k++) { h ^= v->s_obs[k]; h *= 1099511628211ULL; }
    *checksum_out = h;
  }
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
