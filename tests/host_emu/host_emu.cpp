// host_emu.cpp — TEST BUILD ONLY. Compiles the engine's per-lane device headers (mg_obs.cuh, mg_transition.cuh,
// mg_levels.cuh, mg_pcg64.cuh) with g++ through mg_host_shim.h and replays the kernels' orchestration
// (mg_step.cu / mg_reset.cu / mg_state.cu) tile by tile on the CPU, so that the byte-code layout, the gather,
// process_vis bit boards, the stream assembly, the staged-emit lane shift and the generators can be checked
// against the oracle WITHOUT a GPU (tests/test_host_emu.py, `-m "not gpu"`). It is not part of the product:
// minigrid_b200/ never builds or loads it, and the product has no CPU path.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../minigrid_b200/csrc/mg_common.cuh"
#include "../../minigrid_b200/csrc/mg_levels.cuh"
#include "../../minigrid_b200/csrc/mg_obs.cuh"
#include "../../minigrid_b200/csrc/mg_pcg64.cuh"
#include "../../minigrid_b200/csrc/mg_transition.cuh"

using namespace mg;

struct Emu {
  Params p;
  std::vector<uint32_t> grid;
  std::vector<uint4> agent;
  std::vector<RngRec> rng;
  std::vector<double> reward_lut;
  std::vector<uint32_t> cell_lut;
  std::vector<int> list[2];
  int count[2];
  int cur;
  int err;
};

template <int KIND>
static void reset_env(Emu *e, int env, uint8_t *obs, int32_t *dir_out, int set_fresh) {  // k_reset body
  Params &p = e->p;
  Pcg r = load_rng(&e->rng[env]);
  Level L;
  draw_level<KIND>(p, r, L);
  store_rng(&e->rng[env], r);
  uint32_t *col = p.grid + (size_t)(env >> 5) * p.g.wpe * 32 + (env & 31);
  fill_level<KIND>(p, L, col);
  uint4 rec;
  rec.x = (uint32_t)L.ax | ((uint32_t)L.ay << 8);
  rec.y = (uint32_t)L.adir | ((set_fresh ? FLAG_FRESH : 0u) << 8);
  rec.z = 0; rec.w = 0;
  p.agent[env] = rec;
  if (dir_out) dir_out[env] = L.adir;
  if (obs) {
    uint32_t S[OBS_WORDS];
    if (p.see_through) gen_obs_words<true, false>(p.g, col, p.cell_lut, L.ax, L.ay, L.adir, 0u, S);
    else gen_obs_words<false, false>(p.g, col, p.cell_lut, L.ax, L.ay, L.adir, 0u, S);
    emit_obs_bytes(obs + (size_t)env * OBS_BYTES, S);
  }
}
static void reset_one(Emu *e, int env, uint8_t *obs, int32_t *dir_out, int set_fresh) {
  switch (e->p.kind) {
    case KIND_EMPTY: reset_env<KIND_EMPTY>(e, env, obs, dir_out, set_fresh); break;
    case KIND_DOORKEY: reset_env<KIND_DOORKEY>(e, env, obs, dir_out, set_fresh); break;
    case KIND_CROSSING: reset_env<KIND_CROSSING>(e, env, obs, dir_out, set_fresh); break;
    default: reset_env<KIND_FOURROOMS>(e, env, obs, dir_out, set_fresh); break;
  }
}
static void reset_list(Emu *e, int which, uint8_t *obs, int32_t *dir, int set_fresh) {
  for (int i = 0; i < e->count[which]; ++i) reset_one(e, e->list[which][i], obs, dir, set_fresh);
}

template <bool ST>
static void step_tiles(Emu *e, const int32_t *actions, uint8_t *obs, int32_t *dir_out, double *reward_out,
                       uint8_t *term_out, uint8_t *trunc_out, int cur) {  // k_step body
  Params &p = e->p;
  const Geom g = p.g;
  const bool stepping = actions != nullptr;
  if (stepping) e->count[cur ^ 1] = 0;
  std::vector<uint32_t> gtile((size_t)g.wpe * 32), stage(1184);
  for (int tile = 0; tile < p.n_tiles; ++tile) {
    uint32_t *gsrc = p.grid + (size_t)tile * g.wpe * 32;
    memcpy(gtile.data(), gsrc, (size_t)g.wpe * 128);  // the TMA bulk load
    uint32_t S[32][OBS_WORDS];
    const bool full = (tile + 1) * TILE <= p.n_envs;
    for (int lane = 0; lane < 32; ++lane) {
      const int env = tile * TILE + lane;
      const bool active = env < p.n_envs;
      uint4 rec = p.agent[env];
      int ax = rec.x & 0xFF, ay = (rec.x >> 8) & 0xFF, dir = rec.y & 3;
      uint32_t flags = rec.y >> 8, carry = rec.z;
      int steps = (int)rec.w;
      const int action = (actions && active) ? actions[env] : A_DONE;
      const uint32_t *base = gtile.data() + lane;
      double reward = 0.0;
      uint32_t terminated = 0, truncated = 0;
      const bool fresh = (flags & FLAG_FRESH) != 0;
      if (stepping && !fresh) {
        steps += 1;
        int fx, fy;
        front_pos(g, ax, ay, dir, fx, fy);
        const int rw = r_word(g, fx, fy), cw = c_word(g, fx, fy);
        const uint32_t fc = (tile_word<true>(base, rw) >> (8 * (fx & 3))) & 0xFFu;
        const StepOut so = transition(action, fc, fx, fy, ax, ay, dir, carry);
        terminated = so.terminated;
        if (so.goal) {
          if (steps <= p.max_steps) reward = p.reward_lut[steps];
          else { volatile double q = (double)steps / (double)p.max_steps; volatile double m = 0.9 * q; reward = 1.0 - m; }
        }
        if (so.bad_action) e->err |= 1;
        if (so.newc != fc && active) {
          uint8_t *sb = reinterpret_cast<uint8_t *>(gtile.data());
          uint8_t *gb = reinterpret_cast<uint8_t *>(gsrc);
          const size_t ro = ((size_t)rw * 32 + lane) * 4 + (fx & 3), co = ((size_t)cw * 32 + lane) * 4 + (fy & 3);
          sb[ro] = (uint8_t)so.newc; sb[co] = (uint8_t)so.newc;
          gb[ro] = (uint8_t)so.newc; gb[co] = (uint8_t)so.newc;
        }
        truncated = steps >= p.max_steps;
      }
      const bool done = (terminated | truncated) != 0;
      if (stepping) {
        flags &= ~FLAG_FRESH;
        if (p.mode == AUTORESET_NEXT_STEP) flags = done ? (flags | FLAG_PENDING) : (flags & ~FLAG_PENDING);
      }
      if (stepping && p.mode != AUTORESET_DISABLED && done && active) e->list[cur][e->count[cur]++] = env;
      if (obs) {
        gen_obs_words<ST, true>(g, base, p.cell_lut, ax, ay, dir, carry, S[lane]);
        if (!full && active) emit_obs_bytes(obs + (size_t)env * OBS_BYTES, S[lane]);
      }
      if (active) {
        if (stepping) {
          rec.x = (uint32_t)ax | ((uint32_t)ay << 8);
          rec.y = (uint32_t)dir | (flags << 8);
          rec.z = carry; rec.w = (uint32_t)steps;
          p.agent[env] = rec;
        }
        if (dir_out) dir_out[env] = dir;
        if (reward_out) reward_out[env] = reward;
        if (term_out) term_out[env] = (uint8_t)terminated;
        if (trunc_out) trunc_out[env] = (uint8_t)truncated;
      }
    }
    if (obs && full) {
      for (int lane = 0; lane < 32; ++lane)
        emit_obs_staged(stage.data(), lane, S[lane], lane < 31 ? S[lane + 1][0] : S[lane][0]);
      memcpy(obs + (size_t)tile * OBS_TILE_BYTES, stage.data(), OBS_TILE_BYTES);  // the TMA bulk store
    }
  }
}

extern "C" {

void *emu_create(int kind, int W, int H, int max_steps, int see_through, const int32_t *params, int n_params,
                 int n_envs, int mode) {
  Emu *e = new Emu();
  Params &p = e->p;
  memset(&p, 0, sizeof(p));
  p.g = make_geom(W, H);
  p.n_envs = n_envs; p.n_tiles = (n_envs + 31) / 32;
  p.max_steps = max_steps; p.see_through = see_through; p.mode = mode; p.kind = kind;
  for (int i = 0; i < 8; ++i) p.kp[i] = (params && i < n_params) ? params[i] : 0;
  const size_t n_pad = (size_t)p.n_tiles * 32;
  e->grid.assign((size_t)p.n_tiles * p.g.wpe * 32, CODE_WALL4);
  e->agent.assign(n_pad, make_uint4(1u | (1u << 8), 0, 0, 0));
  e->rng.resize(n_pad);
  memset(e->rng.data(), 0, n_pad * sizeof(RngRec));
  e->reward_lut.resize(max_steps + 1);
  for (int k = 0; k <= max_steps; ++k) { volatile double q = (double)k / (double)max_steps; volatile double m = 0.9 * q; e->reward_lut[k] = 1.0 - m; }
  e->cell_lut.resize(256);
  for (uint32_t c = 0; c < 256; ++c) e->cell_lut[c] = decode_cell(c);
  e->list[0].resize(n_pad); e->list[1].resize(n_pad);
  e->count[0] = e->count[1] = 0; e->cur = 0; e->err = 0;
  p.grid = e->grid.data(); p.agent = e->agent.data(); p.rng = e->rng.data();
  p.reward_lut = e->reward_lut.data(); p.cell_lut = e->cell_lut.data();
  return e;
}
void emu_destroy(void *h) { delete (Emu *)h; }
void emu_seed(void *h, const uint64_t *seeds) {
  Emu *e = (Emu *)h;
  for (int i = 0; i < e->p.n_envs; ++i) { Pcg r = seed_pcg64(seeds[i]); store_rng(&e->rng[i], r); }
}
void emu_reset(void *h, uint8_t *obs, int32_t *dir) {
  Emu *e = (Emu *)h;
  e->count[0] = e->count[1] = 0;
  for (int i = 0; i < e->p.n_envs; ++i) reset_one(e, i, obs, dir, 0);
}
int emu_step(void *h, const int32_t *actions, uint8_t *obs, int32_t *dir, double *reward, uint8_t *term, uint8_t *trunc) {
  Emu *e = (Emu *)h;
  if (!actions) {  // mg_gen_obs
    if (e->p.see_through) step_tiles<true>(e, nullptr, obs, dir, nullptr, nullptr, nullptr, e->cur);
    else step_tiles<false>(e, nullptr, obs, dir, nullptr, nullptr, nullptr, e->cur);
    return 0;
  }
  const int append = e->cur ^ 1;
  if (e->p.mode == AUTORESET_NEXT_STEP) reset_list(e, e->cur, nullptr, nullptr, 1);
  if (e->p.see_through) step_tiles<true>(e, actions, obs, dir, reward, term, trunc, append);
  else step_tiles<false>(e, actions, obs, dir, reward, term, trunc, append);
  if (e->p.mode == AUTORESET_SAME_STEP) reset_list(e, append, obs, dir, 0);
  e->cur = append;
  const int bad = e->err; e->err = 0;
  return bad ? -1 : 0;
}
void emu_full_obs(void *h, uint8_t *out, int with_agent) {  // k_full_obs body
  Emu *e = (Emu *)h;
  const Params &p = e->p;
  for (int env = 0; env < p.n_envs; ++env)
    for (int x = 0; x < p.g.W; ++x)
      for (int y = 0; y < p.g.H; ++y) {
        const uint32_t *col = p.grid + (size_t)(env >> 5) * p.g.wpe * 32 + (env & 31);
        const uint32_t code = (col[c_word(p.g, x, y) * 32] >> (8 * (y & 3))) & 0xFFu;
        uint32_t t = p.cell_lut[code];
        const uint4 rec = p.agent[env];
        if (with_agent && (int)(rec.x & 0xFF) == x && (int)((rec.x >> 8) & 0xFF) == y) t = T_AGENT | (C_RED << 8) | ((rec.y & 3u) << 16);
        uint8_t *o = out + (((size_t)env * p.g.W + x) * p.g.H + y) * 3;
        o[0] = (uint8_t)t; o[1] = (uint8_t)(t >> 8); o[2] = (uint8_t)(t >> 16);
      }
}
void emu_get_state(void *h, int32_t *agent, uint64_t *rng, uint8_t *pending) {  // k_get_agent body
  Emu *e = (Emu *)h;
  for (int env = 0; env < e->p.n_envs; ++env) {
    const uint4 rec = e->agent[env];
    int32_t *a = agent + (size_t)env * 6;
    a[0] = rec.x & 0xFF; a[1] = (rec.x >> 8) & 0xFF; a[2] = rec.y & 3;
    a[3] = rec.z ? (int32_t)(rec.z & 15u) : -1; a[4] = rec.z ? (int32_t)((rec.z >> 4) & 7u) : 0; a[5] = (int32_t)rec.w;
    const RngRec r = e->rng[env];
    uint64_t *o = rng + (size_t)env * 6;
    o[0] = r.state_hi; o[1] = r.state_lo; o[2] = r.inc_hi; o[3] = r.inc_lo; o[4] = r.has_uint32; o[5] = r.uinteger;
    pending[env] = ((rec.y >> 8) & FLAG_PENDING) ? 1 : 0;
  }
}
void emu_set_state(void *h, const uint8_t *grid, const int32_t *agent) {  // k_set_grid / k_set_agent bodies
  Emu *e = (Emu *)h;
  const Params &p = e->p;
  for (int env = 0; env < p.n_envs; ++env) {
    if (grid)
      for (int x = 0; x < p.g.W; ++x)
        for (int y = 0; y < p.g.H; ++y) {
          const uint8_t *in = grid + (((size_t)env * p.g.W + x) * p.g.H + y) * 3;
          const uint8_t code = (uint8_t)encode_cell(in[0], in[1], in[2]);
          uint8_t *col = reinterpret_cast<uint8_t *>(p.grid + (size_t)(env >> 5) * p.g.wpe * 32 + (env & 31));
          col[(size_t)r_word(p.g, x, y) * 128 + (x & 3)] = code;
          col[(size_t)c_word(p.g, x, y) * 128 + (y & 3)] = code;
        }
    if (agent) {
      const int32_t *a = agent + (size_t)env * 6;
      uint4 rec = p.agent[env];
      rec.x = (uint32_t)(a[0] & 0xFF) | ((uint32_t)(a[1] & 0xFF) << 8);
      rec.y = (uint32_t)(a[2] & 3);
      rec.z = a[3] >= 0 ? ((uint32_t)(a[3] & 15) | ((uint32_t)(a[4] & 7) << 4)) : 0u;
      rec.w = (uint32_t)a[5];
      p.agent[env] = rec;
    }
  }
}
}
