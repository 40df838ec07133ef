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
#include "../../minigrid_b200/csrc/mg_postfilter.cuh"

using namespace mg;

static size_t step_words(const Geom &g) {
  size_t w = g.layout == LAYOUT_TILED ? (size_t)g.wpe * 32 : (size_t)32 * WIN_LANE_BYTES / 4;
  return w < 1184 ? 1184 : w;
}

struct Emu {
  Params p;
  std::vector<uint32_t> grid;
  std::vector<uint4> agent;
  std::vector<RngRec> rng;
  std::vector<uint4> extra;
  std::vector<double> reward_lut;
  std::vector<uint32_t> cell_lut;
  std::vector<uint16_t> vis_tbl;
  std::vector<uint32_t> tmpl;
  int use_tbl;
  int err;
  uint32_t *packed_out = nullptr;  // MG_HOST_PACKED: 13 words per env instead of the image and the result arrays
};

template <int KIND>
static void reset_env(Emu *e, int env, uint8_t *obs, int32_t *dir_out) {  // k_reset body
  Params &p = e->p;
  Pcg r = load_rng(&e->rng[env]);
  Level L;
  draw_level<KIND>(p, r, L);
  store_rng(&e->rng[env], r);
  fill_level<KIND>(p, L, env);
  uint4 rec;
  rec.x = (uint32_t)L.ax | ((uint32_t)L.ay << 8);
  rec.y = (uint32_t)L.adir;
  if (has_post_filter<KIND>()) {  // post-filter targets
    rec.x |= ((uint32_t)level_tx(L) << 16) | ((uint32_t)level_ty(L) << 24);
    rec.y |= level_aux(L) << 16;
  }
  rec.z = 0; rec.w = 0;
  if (KIND == KIND_DYNOBS) {
    uint32_t ex[4];
    dynobs_pack(L, ex);
    p.extra[env] = make_uint4(ex[0], ex[1], ex[2], ex[3]);
  }
  p.agent[env] = rec;
  if (dir_out) dir_out[env] = L.adir;
  if (obs) {
    uint32_t S[OBS_WORDS];
    const uint32_t *base = p.grid + grid_word(p.g, env, 0);  // gen_obs_global body
    if (p.g.layout == LAYOUT_TILED) {
      const AccTiled acc = {base, false};
      if (p.see_through) gen_obs_words<VIS_NONE>(p.g, acc, p.cell_lut, p.vis_tbl, L.ax, L.ay, L.adir, 0u, S);
      else gen_obs_words<VIS_ALU>(p.g, acc, p.cell_lut, p.vis_tbl, L.ax, L.ay, L.adir, 0u, S);
    } else {
      const AccFlat acc = {base};
      if (p.see_through) gen_obs_words<VIS_NONE>(p.g, acc, p.cell_lut, p.vis_tbl, L.ax, L.ay, L.adir, 0u, S);
      else gen_obs_words<VIS_ALU>(p.g, acc, p.cell_lut, p.vis_tbl, L.ax, L.ay, L.adir, 0u, S);
    }
    emit_obs_bytes(obs + (size_t)env * OBS_BYTES, S);
  }
}
static void reset_one(Emu *e, int env, uint8_t *obs, int32_t *dir_out) {
  switch (e->p.kind) {
    case KIND_EMPTY: reset_env<KIND_EMPTY>(e, env, obs, dir_out); break;
    case KIND_DOORKEY: reset_env<KIND_DOORKEY>(e, env, obs, dir_out); break;
    case KIND_CROSSING: reset_env<KIND_CROSSING>(e, env, obs, dir_out); break;
    case KIND_LAVAGAP: reset_env<KIND_LAVAGAP>(e, env, obs, dir_out); break;
    case KIND_DISTSHIFT: reset_env<KIND_DISTSHIFT>(e, env, obs, dir_out); break;
    case KIND_MULTIROOM: reset_env<KIND_MULTIROOM>(e, env, obs, dir_out); break;
    case KIND_LOCKEDROOM: reset_env<KIND_LOCKEDROOM>(e, env, obs, dir_out); break;
    case KIND_PLAYGROUND: reset_env<KIND_PLAYGROUND>(e, env, obs, dir_out); break;
    case KIND_GOTODOOR: reset_env<KIND_GOTODOOR>(e, env, obs, dir_out); break;
    case KIND_FETCH: reset_env<KIND_FETCH>(e, env, obs, dir_out); break;
    case KIND_REDBLUEDOORS: reset_env<KIND_REDBLUEDOORS>(e, env, obs, dir_out); break;
    case KIND_GOTOOBJECT: reset_env<KIND_GOTOOBJECT>(e, env, obs, dir_out); break;
    case KIND_PUTNEAR: reset_env<KIND_PUTNEAR>(e, env, obs, dir_out); break;
    case KIND_MEMORY: reset_env<KIND_MEMORY>(e, env, obs, dir_out); break;
    case KIND_DYNOBS: reset_env<KIND_DYNOBS>(e, env, obs, dir_out); break;
    case KIND_ROOMGRID: reset_env<KIND_ROOMGRID>(e, env, obs, dir_out); break;
    default: reset_env<KIND_FOURROOMS>(e, env, obs, dir_out); break;
  }
}
// warp_reset body: phase 1 lane-per-env draws, phase 2 template copy + byte patches per environment
struct ResetOut { int ax, ay, dir, tx, ty; uint32_t aux; };
template <int KIND>
static void warp_reset_k(Emu *e, unsigned pend, int tile, uint32_t *gtile, ResetOut out[32]) {
  Params &p = e->p;
  const Geom &g = p.g;
  std::vector<Level> Ls(32, blank_level());
  for (int lane = 0; lane < 32; ++lane)
    if ((pend >> lane) & 1u) {
      const int env = tile * TILE + lane;
      Pcg r = load_rng(&e->rng[env]);
      draw_level<KIND>(p, r, Ls[lane]);
      store_rng(&e->rng[env], r);
      if (KIND == KIND_DYNOBS) {
        uint32_t ex[4];
        dynobs_pack(Ls[lane], ex);
        p.extra[env] = make_uint4(ex[0], ex[1], ex[2], ex[3]);
      }
    }
  for (int lane = 0; lane < 32; ++lane) {
    out[lane].ax = Ls[lane].ax; out[lane].ay = Ls[lane].ay; out[lane].dir = Ls[lane].adir;
    const bool pf = has_post_filter<KIND>();  // kinds with a post-filter: their targets ride in Level::ov
    out[lane].tx = pf ? level_tx(Ls[lane]) : 0; out[lane].ty = pf ? level_ty(Ls[lane]) : 0; out[lane].aux = pf ? level_aux(Ls[lane]) : 0u;
  }
  if (__builtin_popcount(pend) >= 8) {  // DENSE_RESET_MIN: every pending lane fills its own env
    uint8_t *sb = reinterpret_cast<uint8_t *>(gtile), *gb = reinterpret_cast<uint8_t *>(p.grid);
    for (int lane = 0; lane < 32; ++lane) {
      if (!((pend >> lane) & 1u)) continue;
      const Level &L = Ls[lane];
      const int env = tile * TILE + lane;
      for (int w = 0; w < g.wpe; ++w) {
        if (gtile) gtile[w * 32 + lane] = e->tmpl[w];
        p.grid[grid_word(g, env, w)] = e->tmpl[w];
      }
      for (int share = 0; share < 32; ++share)
        patch_level<KIND>(p, L, share, [&](int x, int y) {
          const uint8_t code = (uint8_t)cell_of<KIND>(p, L, x, y);
          const int rw = r_word(g, x, y), cw = c_word(g, x, y);
          if (gtile) { sb[((size_t)rw * 32 + lane) * 4 + (x & 3)] = code; sb[((size_t)cw * 32 + lane) * 4 + (y & 3)] = code; }
          gb[grid_word(g, env, rw) * 4 + (x & 3)] = code;
          gb[grid_word(g, env, cw) * 4 + (y & 3)] = code;
        });
    }
    return;
  }
  while (pend) {
    const int src = __ffs(pend) - 1;
    pend &= pend - 1;
    const Level &B = Ls[src];
    const int env = tile * TILE + src;
    for (int w = 0; w < g.wpe; ++w) {
      if (gtile) gtile[w * 32 + src] = e->tmpl[w];
      p.grid[grid_word(g, env, w)] = e->tmpl[w];
    }
    uint8_t *sb = reinterpret_cast<uint8_t *>(gtile), *gb = reinterpret_cast<uint8_t *>(p.grid);
    for (int lane = 0; lane < 32; ++lane)
      patch_level<KIND>(p, B, lane, [&](int x, int y) {
        const uint8_t code = (uint8_t)cell_of<KIND>(p, B, x, y);
        const int rw = r_word(g, x, y), cw = c_word(g, x, y);
        if (gtile) { sb[((size_t)rw * 32 + src) * 4 + (x & 3)] = code; sb[((size_t)cw * 32 + src) * 4 + (y & 3)] = code; }
        gb[grid_word(g, env, rw) * 4 + (x & 3)] = code;
        gb[grid_word(g, env, cw) * 4 + (y & 3)] = code;
      });
  }
}
static void warp_reset(Emu *e, unsigned pend, int tile, uint32_t *gtile, ResetOut out[32]) {
  switch (e->p.kind) {
    case KIND_EMPTY: warp_reset_k<KIND_EMPTY>(e, pend, tile, gtile, out); break;
    case KIND_DOORKEY: warp_reset_k<KIND_DOORKEY>(e, pend, tile, gtile, out); break;
    case KIND_CROSSING: warp_reset_k<KIND_CROSSING>(e, pend, tile, gtile, out); break;
    case KIND_LAVAGAP: warp_reset_k<KIND_LAVAGAP>(e, pend, tile, gtile, out); break;
    case KIND_DISTSHIFT: warp_reset_k<KIND_DISTSHIFT>(e, pend, tile, gtile, out); break;
    case KIND_MULTIROOM: warp_reset_k<KIND_MULTIROOM>(e, pend, tile, gtile, out); break;
    case KIND_LOCKEDROOM: warp_reset_k<KIND_LOCKEDROOM>(e, pend, tile, gtile, out); break;
    case KIND_PLAYGROUND: warp_reset_k<KIND_PLAYGROUND>(e, pend, tile, gtile, out); break;
    case KIND_GOTODOOR: warp_reset_k<KIND_GOTODOOR>(e, pend, tile, gtile, out); break;
    case KIND_FETCH: warp_reset_k<KIND_FETCH>(e, pend, tile, gtile, out); break;
    case KIND_REDBLUEDOORS: warp_reset_k<KIND_REDBLUEDOORS>(e, pend, tile, gtile, out); break;
    case KIND_GOTOOBJECT: warp_reset_k<KIND_GOTOOBJECT>(e, pend, tile, gtile, out); break;
    case KIND_PUTNEAR: warp_reset_k<KIND_PUTNEAR>(e, pend, tile, gtile, out); break;
    case KIND_MEMORY: warp_reset_k<KIND_MEMORY>(e, pend, tile, gtile, out); break;
    case KIND_DYNOBS: warp_reset_k<KIND_DYNOBS>(e, pend, tile, gtile, out); break;
    case KIND_ROOMGRID: warp_reset_k<KIND_ROOMGRID>(e, pend, tile, gtile, out); break;
    default: warp_reset_k<KIND_FOURROOMS>(e, pend, tile, gtile, out); break;
  }
}

// the step post-filters (mg_postfilter.cuh) of the kinds that have one
static int emu_pre_filter(int kind, int action) { return kind == KIND_MEMORY ? pre_filter<KIND_MEMORY>(action) : action; }
static PostOut emu_post_filter(int kind, const PostIn &in, uint32_t terminated) {
  switch (kind) {
    case KIND_GOTODOOR: return post_filter<KIND_GOTODOOR>(in, terminated);
    case KIND_GOTOOBJECT: return post_filter<KIND_GOTOOBJECT>(in, terminated);
    case KIND_FETCH: return post_filter<KIND_FETCH>(in, terminated);
    case KIND_PUTNEAR: return post_filter<KIND_PUTNEAR>(in, terminated);
    case KIND_MEMORY: return post_filter<KIND_MEMORY>(in, terminated);
    case KIND_REDBLUEDOORS: return post_filter<KIND_REDBLUEDOORS>(in, terminated);
    case KIND_ROOMGRID: return post_filter<KIND_ROOMGRID>(in, terminated);
    default: { PostOut o = {terminated, POST_KEEP}; return o; }
  }
}

template <int VIS>
static void step_tiles(Emu *e, const int32_t *actions, uint8_t *obs, int32_t *dir_out, double *reward_out,
                       uint8_t *term_out, uint8_t *trunc_out) {  // k_step body, phase by phase over the 32 lanes
  Params &p = e->p;
  const Geom g = p.g;
  const bool stepping = actions != nullptr;
  const bool WIN = g.layout == LAYOUT_WINDOW;
  std::vector<uint32_t> gtile((size_t)step_words(g)), S_all(32 * OBS_WORDS);
  uint8_t *gb = reinterpret_cast<uint8_t *>(p.grid);
  for (int tile = 0; tile < p.n_tiles; ++tile) {
    if (!WIN) memcpy(gtile.data(), p.grid + (size_t)tile * g.wpe * 32, (size_t)g.wpe * 128);  // the TMA bulk load
    const bool full = (tile + 1) * TILE <= p.n_envs;
    int ax[32], ay[32], dir[32], steps[32], tx[32], ty[32], rsteps[32] = {0};
    uint32_t flags[32], carry[32], terminated[32] = {0}, truncated[32] = {0};
    double reward[32] = {0};
    bool active[32], fresh[32] = {false};
    uint4 rec[32];
    for (int lane = 0; lane < 32; ++lane) {
      const int env = tile * TILE + lane;
      active[lane] = env < p.n_envs;
      rec[lane] = p.agent[env];
      ax[lane] = rec[lane].x & 0xFF; ay[lane] = (rec[lane].x >> 8) & 0xFF; dir[lane] = rec[lane].y & 3;
      tx[lane] = (rec[lane].x >> 16) & 0xFF; ty[lane] = rec[lane].x >> 24;  // post-filter targets; aux rides in flags >> 8
      flags[lane] = rec[lane].y >> 8; carry[lane] = rec[lane].z; steps[lane] = (int)rec[lane].w;
    }
    if (stepping && p.mode == AUTORESET_NEXT_STEP) {
      unsigned pend = 0;
      for (int lane = 0; lane < 32; ++lane) { fresh[lane] = active[lane] && (flags[lane] & FLAG_PENDING); if (fresh[lane]) pend |= 1u << lane; }
      if (pend) {
        ResetOut ro[32];
        warp_reset(e, pend, tile, WIN ? nullptr : gtile.data(), ro);
        for (int lane = 0; lane < 32; ++lane)
          if (fresh[lane]) {
            ax[lane] = ro[lane].ax; ay[lane] = ro[lane].ay; dir[lane] = ro[lane].dir; carry[lane] = 0; steps[lane] = 0; flags[lane] &= ~FLAG_PENDING;
            tx[lane] = ro[lane].tx; ty[lane] = ro[lane].ty; flags[lane] = (flags[lane] & 0xFFu) | (ro[lane].aux << 8);
          }
      }
    }
    std::vector<ViewWords> vws(32);
    if (WIN)  // K1's window branch: the view words of the post-action direction, loaded before the transition
      for (int lane = 0; lane < 32; ++lane) {
        const int env = tile * TILE + lane;
        int dirn = dir[lane];
        if (stepping && !fresh[lane]) {
          const int action = emu_pre_filter(p.kind, active[lane] ? actions[env] : A_DONE);
          dirn = (dir[lane] + (action == A_LEFT ? 3 : 0) + (action == A_RIGHT ? 1 : 0)) & 3;
        }
        const uint32_t *envw = p.grid + (size_t)env * g.wpe;
        load_view_words(g, ax[lane], ay[lane], dirn, vws[lane], [&](int w) { return envw[w]; });
      }
    for (int lane = 0; lane < 32; ++lane) {
      if (!(stepping && !fresh[lane])) continue;
      const int env = tile * TILE + lane;
      int action = emu_pre_filter(p.kind, active[lane] ? actions[env] : A_DONE);
      const uint32_t carry_before = carry[lane];
      const uint32_t *base = gtile.data() + lane;
      steps[lane] += 1;
      int fx, fy;
      front_pos(g, ax[lane], ay[lane], dir[lane], fx, fy);
      const int rw = r_word(g, fx, fy), cw = c_word(g, fx, fy);
      bool not_clear = false;
      if (p.kind == KIND_DYNOBS && !WIN) {  // K1's Dynamic-Obstacles block
        if (action >= 3) action = A_LEFT;
        const uint32_t fc0 = (tile_word<true>(base, rw) >> (8 * (fx & 3))) & 0xFFu;
        not_clear = fc0 != CODE_EMPTY && (fc0 & 15u) != T_GOAL;
        if (active[lane]) {
          Pcg r = load_rng(&e->rng[env]);
          const uint4 e4 = p.extra[env];
          uint32_t ex[4] = {e4.x, e4.y, e4.z, e4.w};
          uint8_t *sb = reinterpret_cast<uint8_t *>(gtile.data());
          uint8_t *tb = reinterpret_cast<uint8_t *>(p.grid + (size_t)tile * g.wpe * 32);
          dynobs_move(g, r, p.kp[0], ex, ax[lane], ay[lane],
                      [&](int x, int y) { return (tile_word<true>(base, r_word(g, x, y)) >> (8 * (x & 3))) & 0xFFu; },
                      [&](int x, int y, uint32_t code) {
                        const int o_r = (r_word(g, x, y) * 32 + lane) * 4 + (x & 3), o_c = (c_word(g, x, y) * 32 + lane) * 4 + (y & 3);
                        sb[o_r] = (uint8_t)code; sb[o_c] = (uint8_t)code;
                        tb[o_r] = (uint8_t)code; tb[o_c] = (uint8_t)code;
                      });
          store_rng(&e->rng[env], r);
          p.extra[env] = make_uint4(ex[0], ex[1], ex[2], ex[3]);
        }
      }
      uint32_t fc;
      const int fpos = ((dir[lane] & 1) ? ay[lane] : ax[lane]) + ((dir[lane] < 2) ? 1 : -1);
      if (WIN) fc = view_words_byte(vws[lane], fpos);
      else fc = (tile_word<true>(base, rw) >> (8 * (fx & 3))) & 0xFFu;
      const StepOut so = transition(action, fc, fx, fy, ax[lane], ay[lane], dir[lane], carry[lane]);
      terminated[lane] = so.terminated;
      if (so.goal) {
        if (steps[lane] <= p.max_steps) reward[lane] = p.reward_lut[steps[lane]];
        else { volatile double q = (double)steps[lane] / (double)p.max_steps; volatile double m = 0.9 * q; reward[lane] = 1.0 - m; }
      }
      if (so.bad_action) e->err |= 1;
      if (so.newc != fc && active[lane]) {
        if (!WIN) {
          uint8_t *sb = reinterpret_cast<uint8_t *>(gtile.data());
          sb[((size_t)rw * 32 + lane) * 4 + (fx & 3)] = (uint8_t)so.newc;
          sb[((size_t)cw * 32 + lane) * 4 + (fy & 3)] = (uint8_t)so.newc;
        } else {
          view_words_set_byte(vws[lane], fpos, so.newc);
        }
        gb[grid_word(g, env, rw) * 4 + (fx & 3)] = (uint8_t)so.newc;
        gb[grid_word(g, env, cw) * 4 + (fy & 3)] = (uint8_t)so.newc;
      }
      if (p.kind == KIND_DYNOBS && action == A_FORWARD && not_clear) { reward[lane] = -1.0; terminated[lane] = 1u; }
      if ((p.kind >= KIND_GOTODOOR && p.kind <= KIND_MEMORY) || p.kind == KIND_ROOMGRID) {  // step post-filter
        PostIn in;
        in.action = action; in.ax = ax[lane]; in.ay = ay[lane]; in.dir = dir[lane];
        in.carry_before = carry_before; in.carry = carry[lane];
        in.tx = tx[lane]; in.ty = ty[lane]; in.aux = flags[lane] >> 8;
        in.red_before = in.blue_before = in.red_after = in.blue_after = false;
        in.variant = p.kp[0]; in.door_open = false;
        if (p.kind == KIND_ROOMGRID && p.kp[0] == RG_UNLOCK) in.door_open = (gb[cell_byte_R(g, env, tx[lane], ty[lane])] & 15u) == T_DOOR;
        if (p.kind == KIND_REDBLUEDOORS) {  // a door changes only as the front cell of a toggle
          const int xl = g.H / 2, xr = g.H / 2 + g.H - 1;
          in.red_after = (gb[cell_byte_R(g, env, xl, tx[lane])] & 15u) == T_DOOR;
          in.blue_after = (gb[cell_byte_R(g, env, xr, ty[lane])] & 15u) == T_DOOR;
          in.red_before = (fx == xl && fy == tx[lane]) ? (fc & 15u) == T_DOOR : in.red_after;
          in.blue_before = (fx == xr && fy == ty[lane]) ? (fc & 15u) == T_DOOR : in.blue_after;
        }
        const PostOut po = emu_post_filter(p.kind, in, terminated[lane]);
        terminated[lane] = po.terminated;
        if (po.reward == POST_ZERO) reward[lane] = 0.0;
        if (po.reward == POST_REWARD) {
          if (steps[lane] <= p.max_steps) reward[lane] = p.reward_lut[steps[lane]];
          else { volatile double q = (double)steps[lane] / (double)p.max_steps; volatile double m = 0.9 * q; reward[lane] = 1.0 - m; }
        }
      }
      truncated[lane] = steps[lane] >= p.max_steps;
      rsteps[lane] = steps[lane];
      const bool done = (terminated[lane] | truncated[lane]) != 0;
      if (p.mode == AUTORESET_NEXT_STEP) flags[lane] = done ? (flags[lane] | FLAG_PENDING) : (flags[lane] & ~FLAG_PENDING);
    }
    if (stepping && p.mode == AUTORESET_SAME_STEP) {
      unsigned pend = 0;
      bool again[32];
      for (int lane = 0; lane < 32; ++lane) { again[lane] = active[lane] && ((terminated[lane] | truncated[lane]) != 0); if (again[lane]) pend |= 1u << lane; }
      if (pend) {
        ResetOut ro[32];
        warp_reset(e, pend, tile, WIN ? nullptr : gtile.data(), ro);
        for (int lane = 0; lane < 32; ++lane)
          if (again[lane]) {
            ax[lane] = ro[lane].ax; ay[lane] = ro[lane].ay; dir[lane] = ro[lane].dir; carry[lane] = 0; steps[lane] = 0;
            tx[lane] = ro[lane].tx; ty[lane] = ro[lane].ty; flags[lane] = (flags[lane] & 0xFFu) | (ro[lane].aux << 8);
            if (WIN) {
              const uint32_t *envw = p.grid + (size_t)(tile * TILE + lane) * g.wpe;
              load_view_words(g, ax[lane], ay[lane], dir[lane], vws[lane], [&](int w) { return envw[w]; });
            }
          }
      }
    }
    if (e->packed_out) {  // K1's packed branch: gather_view + pack_codes, 13 words per lane
      for (int lane = 0; lane < 32; ++lane) {
        const int env = tile * TILE + lane;
        uint32_t clo[VIEW], chi[VIEW], P[PACKED_WORDS];
        if (WIN) {
          gather_from_words<VIS>(g, vws[lane], p.vis_tbl, ax[lane], ay[lane], dir[lane], carry[lane], clo, chi);
        } else {
          const AccTiled acc = {gtile.data() + lane, true};
          gather_view<VIS>(g, acc, p.vis_tbl, ax[lane], ay[lane], dir[lane], carry[lane], clo, chi);
        }
        pack_codes(clo, chi, packed_tail(dir[lane], terminated[lane], truncated[lane], reward[lane] != 0.0 ? 1u : 0u,
                                         reward[lane] < 0.0 ? PACKED_MAX_STEPS : (uint32_t)rsteps[lane]), P);
        if (active[lane]) memcpy(e->packed_out + (size_t)env * PACKED_WORDS, P, sizeof(P));
      }
    } else if (obs) {
      for (int lane = 0; lane < 32; ++lane) {
        uint32_t(&S)[OBS_WORDS] = *reinterpret_cast<uint32_t(*)[OBS_WORDS]>(&S_all[lane * OBS_WORDS]);
        if (WIN) {  // the lane's preloaded view words
          uint32_t clo[VIEW], chi[VIEW];
          gather_from_words<VIS>(g, vws[lane], p.vis_tbl, ax[lane], ay[lane], dir[lane], carry[lane], clo, chi);
          encode_stream(p.cell_lut, clo, chi, S);
        } else {
          const AccTiled acc = {gtile.data() + lane, true};
          gen_obs_words<VIS>(g, acc, p.cell_lut, p.vis_tbl, ax[lane], ay[lane], dir[lane], carry[lane], S);
        }
      }
      const int nvalid = p.n_envs - tile * TILE < TILE ? p.n_envs - tile * TILE : TILE;
      (void)full;
      {  // the consumed tile buffer becomes the stage; full tiles leave by bulk store, the ragged tail byte by byte
        for (int lane = 0; lane < 32; ++lane) {
          uint32_t(&S)[OBS_WORDS] = *reinterpret_cast<uint32_t(*)[OBS_WORDS]>(&S_all[lane * OBS_WORDS]);
          emit_obs_staged(gtile.data(), lane, S, S_all[(lane < 31 ? lane + 1 : lane) * OBS_WORDS]);
        }
        memcpy(obs + (size_t)tile * OBS_TILE_BYTES, gtile.data(), (size_t)nvalid * OBS_BYTES);
      }
    }
    for (int lane = 0; lane < 32; ++lane) {
      if (!active[lane]) continue;
      const int env = tile * TILE + lane;
      if (stepping) {
        uint4 r = rec[lane];
        r.x = (uint32_t)ax[lane] | ((uint32_t)ay[lane] << 8) | ((uint32_t)tx[lane] << 16) | ((uint32_t)ty[lane] << 24);
        r.y = (uint32_t)dir[lane] | (flags[lane] << 8);
        r.z = carry[lane]; r.w = (uint32_t)steps[lane];
        p.agent[env] = r;
      }
      if (dir_out) dir_out[env] = dir[lane];
      if (reward_out) reward_out[env] = reward[lane];
      if (term_out) term_out[env] = (uint8_t)terminated[lane];
      if (trunc_out) trunc_out[env] = (uint8_t)truncated[lane];
    }
  }
}

extern "C" {

void *emu_create(int kind, int W, int H, int max_steps, int see_through, const int32_t *params, int n_params,
                 int n_envs, int mode, int layout) {
  Emu *e = new Emu();
  Params &p = e->p;
  memset(&p, 0, sizeof(p));
  if (layout < 0) layout = make_geom(W, H, LAYOUT_TILED).wpe * 4 > 512 ? LAYOUT_WINDOW : LAYOUT_TILED;  // mg_create rule
  if (kind == KIND_DYNOBS) layout = LAYOUT_TILED;
  p.g = make_geom(W, H, layout);
  p.n_envs = n_envs; p.n_tiles = (n_envs + 31) / 32;
  p.max_steps = max_steps; p.see_through = see_through; p.mode = mode; p.kind = kind;
  for (int i = 0; i < 8; ++i) p.kp[i] = (params && i < n_params) ? params[i] : 0;
  const size_t n_pad = (size_t)p.n_tiles * 32;
  e->grid.assign((size_t)p.n_tiles * p.g.wpe * 32 + 64, CODE_WALL4);
  e->agent.assign(n_pad, make_uint4(1u | (1u << 8), 0, 0, 0));
  e->extra.assign(n_pad, make_uint4(0, 0, 0, 0));
  p.extra = e->extra.data();
  e->rng.resize(n_pad);
  memset(e->rng.data(), 0, n_pad * sizeof(RngRec));
  e->reward_lut.resize(max_steps + 1);
  for (int k = 0; k <= max_steps; ++k) { volatile double q = (double)k / (double)max_steps; volatile double m = 0.9 * q; e->reward_lut[k] = 1.0 - m; }
  e->cell_lut.resize(256);
  for (uint32_t c = 0; c < 256; ++c) e->cell_lut[c] = decode_cell(c);
  e->vis_tbl.resize(128 * 128);
  build_vis_table(e->vis_tbl.data());
  p.vis_tbl = e->vis_tbl.data();
  e->use_tbl = (n_envs % 2) == 0;
  e->tmpl.resize(p.g.wpe);
  {  // k_template body
    const Level L = blank_level();
    for (int w = 0; w < p.g.wpe; ++w)
      switch (kind) {
        case KIND_EMPTY: e->tmpl[w] = level_word<KIND_EMPTY>(p, L, w); break;
        case KIND_DOORKEY: e->tmpl[w] = level_word<KIND_DOORKEY>(p, L, w); break;
        case KIND_CROSSING: e->tmpl[w] = level_word<KIND_CROSSING>(p, L, w); break;
        case KIND_LAVAGAP: e->tmpl[w] = level_word<KIND_LAVAGAP>(p, L, w); break;
        case KIND_DISTSHIFT: e->tmpl[w] = level_word<KIND_DISTSHIFT>(p, L, w); break;
        case KIND_MULTIROOM: e->tmpl[w] = level_word<KIND_MULTIROOM>(p, L, w); break;
        case KIND_LOCKEDROOM: e->tmpl[w] = level_word<KIND_LOCKEDROOM>(p, L, w); break;
        case KIND_PLAYGROUND: e->tmpl[w] = level_word<KIND_PLAYGROUND>(p, L, w); break;
        case KIND_GOTODOOR: e->tmpl[w] = level_word<KIND_GOTODOOR>(p, L, w); break;
        case KIND_FETCH: e->tmpl[w] = level_word<KIND_FETCH>(p, L, w); break;
        case KIND_REDBLUEDOORS: e->tmpl[w] = level_word<KIND_REDBLUEDOORS>(p, L, w); break;
        case KIND_GOTOOBJECT: e->tmpl[w] = level_word<KIND_GOTOOBJECT>(p, L, w); break;
        case KIND_PUTNEAR: e->tmpl[w] = level_word<KIND_PUTNEAR>(p, L, w); break;
        case KIND_MEMORY: e->tmpl[w] = level_word<KIND_MEMORY>(p, L, w); break;
        case KIND_DYNOBS: e->tmpl[w] = level_word<KIND_DYNOBS>(p, L, w); break;
        case KIND_ROOMGRID: e->tmpl[w] = level_word<KIND_ROOMGRID>(p, L, w); break;
        default: e->tmpl[w] = level_word<KIND_FOURROOMS>(p, L, w); break;
      }
    p.tmpl = e->tmpl.data();
  }  // exercise both process_vis forms across the test matrix
  e->err = 0;
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
  for (int i = 0; i < e->p.n_envs; ++i) reset_one(e, i, obs, dir);
}
int emu_step(void *h, const int32_t *actions, uint8_t *obs, int32_t *dir, double *reward, uint8_t *term, uint8_t *trunc) {
  Emu *e = (Emu *)h;  // mg_step / mg_gen_obs (actions == NULL): one K1 launch
  if (e->p.see_through) step_tiles<VIS_NONE>(e, actions, obs, dir, reward, term, trunc);
  else if (e->use_tbl) step_tiles<VIS_TBL>(e, actions, obs, dir, reward, term, trunc);
  else step_tiles<VIS_ALU>(e, actions, obs, dir, reward, term, trunc);
  const int bad = e->err; e->err = 0;
  return bad ? -1 : 0;
}
int emu_step_packed(void *h, const int32_t *actions, uint32_t *packed) {  // mg_step_host with MG_HOST_PACKED: K1 with packed_out
  Emu *e = (Emu *)h;
  e->packed_out = packed;
  const int rc = emu_step(h, actions, nullptr, nullptr, nullptr, nullptr, nullptr);
  e->packed_out = nullptr;
  return rc;
}
void emu_full_obs(void *h, uint8_t *out, int with_agent) {  // k_full_obs body
  Emu *e = (Emu *)h;
  const Params &p = e->p;
  for (int env = 0; env < p.n_envs; ++env)
    for (int x = 0; x < p.g.W; ++x)
      for (int y = 0; y < p.g.H; ++y) {
        const uint32_t code = reinterpret_cast<const uint8_t *>(p.grid)[cell_byte_C(p.g, env, x, y)];
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
    const bool boxed = (rec.z & 15u) == T4_BOX_WITH_KEY;  // k_get_agent: a grey box, whatever is inside
    a[3] = rec.z ? (boxed ? (int32_t)T_BOX : (int32_t)(rec.z & 15u)) : -1;
    a[4] = rec.z ? (boxed ? (int32_t)C_GREY : (int32_t)((rec.z >> 4) & 7u)) : 0; a[5] = (int32_t)rec.w;
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
          uint8_t *gbp = reinterpret_cast<uint8_t *>(p.grid);
          gbp[cell_byte_R(p.g, env, x, y)] = code;
          gbp[cell_byte_C(p.g, env, x, y)] = code;
        }
    if (agent) {
      const int32_t *a = agent + (size_t)env * 6;
      uint4 rec = p.agent[env];
      rec.x = (rec.x & 0xFFFF0000u) | (uint32_t)(a[0] & 0xFF) | ((uint32_t)(a[1] & 0xFF) << 8);  // k_set_agent: the post-filter targets stay
      rec.y = (rec.y & ~3u) | (uint32_t)(a[2] & 3);
      rec.z = a[3] >= 0 ? ((uint32_t)(a[3] & 15) | ((uint32_t)(a[4] & 7) << 4)) : 0u;
      rec.w = (uint32_t)a[5];
      p.agent[env] = rec;
    }
  }
}
}
