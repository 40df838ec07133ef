// mg_step.cu — K1: MiniGridEnv.step (minigrid_env.py:525-595) fused with gen_obs (:597-650) and with the
// vector-level autoreset (MiniGridEnv.reset, :119-157) for a batch: ONE launch per vector step.
//
// One warp = one tile of 32 environments, one lane per environment.
//   1. lane 0 issues a TMA bulk copy (cp.async.bulk -> SASS UBLKCP) of the tile's interleaved grid words
//      into shared memory and arms an mbarrier with the byte count; meanwhile every lane loads its
//      action and 16-byte agent record with coalesced loads. Warps are persistent; with NBUF == 2 they are
//      double-buffered (the copy for the next tile is issued before the current one is processed).
//   2. autoreset (NEXT_STEP: envs flagged last step, before the transition; SAME_STEP: envs that just ended,
//      after it): rare, so the whole warp regenerates one environment at a time — every lane replays the
//      numpy-exact RNG draws (uniform control flow) and fills a 1/32 share of the level's words.
//   3. transition: the 7-action rule on (agent, carrying, the one cell in front), predicated, with the
//      rare cell mutation written to the staged tile and straight back to HBM (2 byte stores).
//   4. observation in registers (mg_obs.cuh), staged into the consumed tile buffer in output layout, then one
//      TMA bulk store of the warp's 32 x 147 = 4704 contiguous bytes.
//   5. coalesced stores of direction / reward / terminated / truncated and the agent record.
#include <cstdlib>

#include "mg_common.cuh"
#include "mg_levels.cuh"
#include "mg_obs.cuh"
#include "mg_pcg64.cuh"
#include "mg_transition.cuh"

namespace mg {

constexpr int STEP_WARPS = 4;
constexpr int STEP_THREADS = STEP_WARPS * 32;

// per-warp buffer: holds the staged tile, then (once the gather has consumed it) the warp's 4704-byte
// observation block in output layout.
__host__ __device__ inline uint32_t step_buf_bytes(const Geom &g) {
  uint32_t b = (uint32_t)g.wpe * 128u;
  if (b < (uint32_t)OBS_TILE_BYTES) b = OBS_TILE_BYTES;
  return (b + 127u) & ~127u;
}
__host__ __device__ inline size_t step_smem_bytes(const Geom &g, int nbuf) {
  return 1024 /*cell table*/ + (size_t)STEP_WARPS * nbuf * step_buf_bytes(g) + 128 /*mbarriers + tile counter*/;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_1d(void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ int load_action(const void *actions, int dtype, int env) {
  if (dtype == 1) return (int)reinterpret_cast<const long long *>(actions)[env];
  if (dtype == 2) return (int)reinterpret_cast<const uint8_t *>(actions)[env];
  return reinterpret_cast<const int *>(actions)[env];
}

// MiniGridEnv.reset() for the lanes in `pend`, one environment at a time with the whole warp: all lanes
// replay the draws (same RNG state, uniform control flow), lane L fills words L, L+32, ... of the level into
// the staged tile and HBM, and the owning lane takes the new agent state. Out of line: it is the rare path
// and must not cost the hot loop registers.
struct ResetOut { int ax, ay, dir; };

template <int KIND>
__device__ __noinline__ ResetOut warp_reset(const Params &p, unsigned pend, int tile, uint32_t *gtile, int lane) {
  ResetOut out = {0, 0, 0};
  const Geom &g = p.g;
  uint32_t *gsrc = p.grid + (size_t)tile * g.wpe * 32;
  while (pend) {
    const int src = __ffs(pend) - 1;
    pend &= pend - 1;
    const int env = tile * TILE + src;
    Pcg r = load_rng(p.rng + env);
    Level L;
    draw_level<KIND>(p, r, L);
    if (lane == 0) store_rng(p.rng + env, r);
    for (int w = lane; w < g.wpe; w += 32) {
      const uint32_t word = level_word<KIND>(p, L, w);
      gtile[w * 32 + src] = word;
      gsrc[w * 32 + src] = word;
    }
    if (lane == src) { out.ax = L.ax; out.ay = L.ay; out.dir = L.adir; }
  }
  __syncwarp();
  return out;
}

template <int KIND, bool SEE_THROUGH, int NBUF>
__global__ void __launch_bounds__(STEP_THREADS)
k_step(Params p, const void *__restrict__ actions, int act_dtype, uint8_t *__restrict__ obs,
       int32_t *__restrict__ dir_out, double *__restrict__ reward_out, uint8_t *__restrict__ term_out,
       uint8_t *__restrict__ trunc_out, int obs_tma_ok) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const Geom g = p.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tile_bytes = (uint32_t)g.wpe * 128u;
  const uint32_t buf_bytes = step_buf_bytes(g);

  uint32_t *lut = reinterpret_cast<uint32_t *>(smem_raw);
  uint8_t *bufs = smem_raw + 1024 + (size_t)warp * NBUF * buf_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + 1024 + (size_t)STEP_WARPS * NBUF * buf_bytes);
  const uint32_t bar0 = smem_u32(bars + 2 * warp);
  int *s_next = reinterpret_cast<int *>(bars + 2 * STEP_WARPS);

  // Programmatic dependent launch: let the next kernel in the stream start its prologue while this grid drains,
  // and do our own prologue (no global memory touched) before waiting for the previous grid to complete.
  asm volatile("griddepcontrol.launch_dependents;");
  const bool stepping = actions != nullptr;  // nullptr: observation-only pass (MiniGridEnv.gen_obs), state untouched
  // one wave of persistent CTAs; CTA c owns tiles [c T/G, (c+1) T/G), its warps pull from a shared counter
  const int t_lo = (int)(((long long)p.n_tiles * blockIdx.x) / gridDim.x);
  const int t_hi = (int)(((long long)p.n_tiles * (blockIdx.x + 1)) / gridDim.x);
  if (threadIdx.x == 0) *s_next = t_lo + NBUF * STEP_WARPS;
  int tile = t_lo + warp;
  int next = (NBUF == 2) ? t_lo + STEP_WARPS + warp : p.n_tiles;
  if (tile >= t_hi) tile = p.n_tiles;
  if (next >= t_hi) next = p.n_tiles;
  uint4 rec = make_uint4(0, 0, 0, 0);
  int action = A_DONE;
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the 256-entry (type, colour, state) table is pure arithmetic: no global load anywhere near the critical path
  for (int i = threadIdx.x; i < 256; i += STEP_THREADS) lut[i] = decode_cell((uint32_t)i);
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");  // everything below reads state the previous step wrote
  if (NBUF == 2 && tile < p.n_tiles) {
    if (lane == 0) {
      mbar_expect_tx(bar0, tile_bytes);
      tma_load_1d(smem_u32(bufs), p.grid + (size_t)tile * g.wpe * 32, tile_bytes, bar0);
    }
    const int env = tile * TILE + lane;
    rec = p.agent[env];
    if (stepping && env < p.n_envs) action = load_action(actions, act_dtype, env);
  }

  uint32_t phase = 0;  // bit b = parity to wait for on buffer b
  int b = 0;
  while (tile < p.n_tiles) {
    uint4 rec_n = make_uint4(0, 0, 0, 0);
    int action_n = A_DONE, nn = p.n_tiles;
    if (NBUF == 2) {
      // ---- prefetch tile `next` into the other buffer, and the index of the tile after it ----
      if (next < p.n_tiles) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the obs block staged there two tiles ago
          const uint32_t nb = bar0 + 8u * (uint32_t)(b ^ 1);
          mbar_expect_tx(nb, tile_bytes);
          tma_load_1d(smem_u32(bufs + (size_t)(b ^ 1) * buf_bytes), p.grid + (size_t)next * g.wpe * 32, tile_bytes, nb);
          nn = atomicAdd(s_next, 1);  // shared-memory atomic, consumed one tile later
          if (nn >= t_hi) nn = p.n_tiles;
        }
        const int env_n = next * TILE + lane;
        rec_n = p.agent[env_n];
        if (stepping && env_n < p.n_envs) action_n = load_action(actions, act_dtype, env_n);
      }
    } else {
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous obs block has left the buffer
        mbar_expect_tx(bar0, tile_bytes);
        tma_load_1d(smem_u32(bufs), p.grid + (size_t)tile * g.wpe * 32, tile_bytes, bar0);
        nn = atomicAdd(s_next, 1);
        if (nn >= t_hi) nn = p.n_tiles;
      }
      const int env0 = tile * TILE + lane;
      rec = p.agent[env0];
      action = (stepping && env0 < p.n_envs) ? load_action(actions, act_dtype, env0) : A_DONE;
    }

    uint32_t *gtile = reinterpret_cast<uint32_t *>(bufs + (size_t)b * buf_bytes);
    uint32_t *gsrc = p.grid + (size_t)tile * g.wpe * 32;
    const int env = tile * TILE + lane;
    const bool active = env < p.n_envs;
    int ax = rec.x & 0xFF, ay = (rec.x >> 8) & 0xFF;
    int dir = rec.y & 3;
    uint32_t flags = rec.y >> 8;
    uint32_t carry = rec.z;
    int steps = (int)rec.w;

    mbar_wait(bar0 + 8u * (uint32_t)b, (phase >> b) & 1u);
    phase ^= 1u << b;

    const uint32_t *base = gtile + lane;
    double reward = 0.0;
    uint32_t terminated = 0, truncated = 0;
    // NEXT_STEP autoreset (gymnasium >= 1.0 SyncVectorEnv): an env that ended last step ignores its action,
    // is reset now, and returns the reset obs with reward 0 / False / False
    bool fresh = false;
    if (stepping && p.mode == AUTORESET_NEXT_STEP) {
      fresh = active && (flags & FLAG_PENDING);
      const unsigned pend = __ballot_sync(0xFFFFFFFFu, fresh);
      if (pend) {
        const ResetOut ro = warp_reset<KIND>(p, pend, tile, gtile, lane);
        if (fresh) { ax = ro.ax; ay = ro.ay; dir = ro.dir; carry = 0; steps = 0; flags &= ~FLAG_PENDING; }
      }
    }
    if (stepping && !fresh) {
      // ---- MiniGridEnv.step, minigrid_env.py:525-588 ----
      steps += 1;
      int fx, fy;
      front_pos(g, ax, ay, dir, fx, fy);
      const int rw = r_word(g, fx, fy), cw = c_word(g, fx, fy);
      const uint32_t fc = (tile_word<true>(base, rw) >> (8 * (fx & 3))) & 0xFFu;
      const StepOut so = transition(action, fc, fx, fy, ax, ay, dir, carry);
      const uint32_t newc = so.newc;
      terminated = so.terminated;
      if (so.goal)  // _reward(), minigrid_env.py:240-245: host-computed table, never an FMA
        reward = steps <= p.max_steps ? p.reward_lut[steps]
                                      : __dsub_rn(1.0, __dmul_rn(0.9, __ddiv_rn((double)steps, (double)p.max_steps)));
      if (so.bad_action) atomicOr(p.err, 1);  // ValueError("Unknown action"), minigrid_env.py:584-585
      if (newc != fc && active) {
        uint8_t *sb = reinterpret_cast<uint8_t *>(gtile);
        uint8_t *gb = reinterpret_cast<uint8_t *>(gsrc);
        const size_t ro = ((size_t)rw * 32 + lane) * 4 + (fx & 3), co = ((size_t)cw * 32 + lane) * 4 + (fy & 3);
        sb[ro] = (uint8_t)newc; sb[co] = (uint8_t)newc;
        gb[ro] = (uint8_t)newc; gb[co] = (uint8_t)newc;
      }
      truncated = steps >= p.max_steps;
      const bool done = (terminated | truncated) != 0;
      if (p.mode == AUTORESET_NEXT_STEP) flags = done ? (flags | FLAG_PENDING) : (flags & ~FLAG_PENDING);
    }
    // SAME_STEP autoreset: the env is reset inside the step that ended it and the reset obs is returned
    if (stepping && p.mode == AUTORESET_SAME_STEP) {
      const bool again = active && ((terminated | truncated) != 0);
      const unsigned pend = __ballot_sync(0xFFFFFFFFu, again);
      if (pend) {
        const ResetOut ro = warp_reset<KIND>(p, pend, tile, gtile, lane);
        if (again) { ax = ro.ax; ay = ro.ay; dir = ro.dir; carry = 0; steps = 0; }
      }
    }

    // ---- gen_obs ----
    if (obs != nullptr) {
      uint32_t S[OBS_WORDS];
      gen_obs_words<SEE_THROUGH, true>(g, base, lut, ax, ay, dir, carry, S);
      const bool full = (tile + 1) * TILE <= p.n_envs;
      if (full && obs_tma_ok) {
        const uint32_t n0 = __shfl_down_sync(0xFFFFFFFFu, S[0], 1);  // also: every lane is past its tile reads
        emit_obs_staged(gtile, lane, S, n0);                         // the consumed tile buffer becomes the stage
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_1d(obs + (size_t)tile * OBS_TILE_BYTES, smem_u32(gtile), OBS_TILE_BYTES);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (active) {
        emit_obs_bytes(obs + (size_t)env * OBS_BYTES, S);
      }
    }
    if (active) {
      if (stepping) {
        rec.x = (uint32_t)ax | ((uint32_t)ay << 8);
        rec.y = (uint32_t)dir | (flags << 8);
        rec.z = carry;
        rec.w = (uint32_t)steps;
        p.agent[env] = rec;
      }
      if (dir_out) dir_out[env] = dir;
      if (reward_out) reward_out[env] = reward;
      if (term_out) term_out[env] = (uint8_t)terminated;
      if (trunc_out) trunc_out[env] = (uint8_t)truncated;
    }
    __syncwarp();  // lanes may still be reading this buffer (partial-tile path) before it is refilled
    if (NBUF == 2) {
      tile = next;
      next = __shfl_sync(0xFFFFFFFFu, nn, 0);
      rec = rec_n;
      action = action_n;
      b ^= 1;
    } else {
      tile = __shfl_sync(0xFFFFFFFFu, nn, 0);
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

typedef void (*StepKernel)(Params, const void *, int, uint8_t *, int32_t *, double *, uint8_t *, uint8_t *, int);

template <int NBUF>
static StepKernel pick_kernel(int kind, int see_through) {
#define MG_K(K) (see_through ? (StepKernel)k_step<K, true, NBUF> : (StepKernel)k_step<K, false, NBUF>)
  switch (kind) {
    case KIND_EMPTY: return MG_K(KIND_EMPTY);
    case KIND_DOORKEY: return MG_K(KIND_DOORKEY);
    case KIND_CROSSING: return MG_K(KIND_CROSSING);
    default: return MG_K(KIND_FOURROOMS);
  }
#undef MG_K
}
static StepKernel step_kernel(const Params &p, int nbuf) {
  return nbuf == 2 ? pick_kernel<2>(p.kind, p.see_through) : pick_kernel<1>(p.kind, p.see_through);
}

// opt in to the tile-dependent dynamic shared memory once per handle, choose single or double buffering
// (double buffering needs twice the shared memory per warp) and size the persistent grid: one wave of CTAs,
// never more CTAs than there are groups of STEP_WARPS tiles
cudaError_t configure_step(const Params &p, int nbuf_request, int *nbuf_out, int *grid_out) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int best_nbuf = 0, best_ctas = 0;
  for (int nbuf = 2; nbuf >= 1; --nbuf) {
    if (nbuf_request && nbuf != nbuf_request) continue;
    const size_t smem = step_smem_bytes(p.g, nbuf);
    if (smem > 227 * 1024) continue;
    StepKernel k = step_kernel(p, nbuf);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int ctas = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k, STEP_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (ctas < 1) continue;
    // prefer double buffering unless it leaves fewer than 4 CTAs (16 warps) per SM and single buffering has more
    if (best_nbuf == 0 || (best_ctas < 4 && ctas > best_ctas)) { best_nbuf = nbuf; best_ctas = ctas; }
  }
  if (best_nbuf == 0) return cudaErrorInvalidValue;
  const long long want = ((long long)p.n_tiles + STEP_WARPS - 1) / STEP_WARPS;
  long long grid = (long long)sms * best_ctas;
  if (grid > want) grid = want;
  *grid_out = (int)(grid < 1 ? 1 : grid);
  *nbuf_out = best_nbuf;
  return cudaSuccess;
}

cudaError_t launch_step(const Params &p, int nbuf, int grid, const void *actions, int action_dtype, uint8_t *obs,
                        int32_t *dir, double *reward, uint8_t *term, uint8_t *trunc, cudaStream_t stream) {
  const size_t smem = step_smem_bytes(p.g, nbuf);
  const int tma_ok = ((reinterpret_cast<uintptr_t>(obs) & 15u) == 0) ? 1 : 0;
  StepKernel k = step_kernel(p, nbuf);
  static const bool use_pdl = []() { const char *e = getenv("MINIGRID_B200_PDL"); return !e || atoi(e) != 0; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(STEP_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k, p, actions, action_dtype, obs, dir, reward, term, trunc, tma_ok);
}

}  // namespace mg
