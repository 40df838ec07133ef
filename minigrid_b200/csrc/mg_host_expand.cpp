// mg_host_expand.cpp — host side of MG_HOST_PACKED (include/minigrid_b200.h): the end-to-end path is PCIe-bound
// (161 B per env-step at ~55 GB/s), so K1 ships 52 B per env-step — the view's 49 one-byte cell codes, i.e. the image
// BEFORE Grid.encode's (type, colour, state) table (grid.py:244-268, world_object.py:65-67,196-212), plus direction,
// flags and the step count the reward is a function of — and the table lookup, which is host-computed data in both
// formats, runs here on the host cores while the next chunk is still crossing the bus. Bit-identical by construction:
// the same 256-entry table (decode_cell) and the same IEEE-double reward expression (minigrid_env.py:240-245).
// Host code only; no CUDA in this file.
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <immintrin.h>
#include <sched.h>

#include "../../include/minigrid_b200.h"
#include "mg_host_expand.h"

namespace mg {

namespace {

constexpr int kObsBytes = 147, kPacked = 52, kCells = 49;
constexpr uint32_t kNegOne = (1u << 19) - 1u;  // PACKED_MAX_STEPS: reserved step count meaning reward -1

// WorldObj.encode / Door.encode of a cell code: bits 0-3 t4 (type with the door state folded in: 4 open, 11 closed,
// 12 locked), bits 4-6 colour. Same function as decode_cell() in mg_common.cuh, restated for a host-only file.
inline uint32_t decode_code(uint32_t code) {
  const uint32_t t4 = code & 15u, color = (code >> 4) & 7u;
  if (t4 == 0) return 0;
  if (t4 == 1) return 1;
  if (t4 == 11) return 4u | (color << 8) | (1u << 16);
  if (t4 == 12) return 4u | (color << 8) | (2u << 16);
  if (t4 == 13) return 7u | (5u << 8);  // a grey box with a key inside (ObstructedMaze): the contents do not show
  return t4 | (color << 8);
}

struct Tables {
  uint32_t lut[256];
  Tables() {
    for (uint32_t c = 0; c < 256; ++c) lut[c] = decode_code(c);
  }
};
const Tables &tables() {
  static const Tables t;
  return t;
}

inline double reward_of(uint32_t steps, int max_steps) {
  // _reward(): 1 - 0.9 * (step_count / max_steps), evaluated like Python does: three separately rounded operations
  volatile double q = (double)steps / (double)max_steps;
  volatile double m = 0.9 * q;
  return 1.0 - m;
}

// destination of `count` consecutive records (any pointer may be NULL)
struct Dst { uint8_t *obs; int32_t *dir; double *reward; uint8_t *term, *trunc; };

inline void expand_tail(const uint8_t *rec, int k, const ExpandJob &j, const Dst &d) {
  const uint32_t f = rec[49];
  if (d.dir) d.dir[k] = (int32_t)(f & 3u);
  if (d.term) d.term[k] = (uint8_t)((f >> 2) & 1u);
  if (d.trunc) d.trunc[k] = (uint8_t)((f >> 3) & 1u);
  if (d.reward) {
    double r = 0.0;
    if (f & 16u) {
      const uint32_t steps = (uint32_t)rec[50] | ((uint32_t)rec[51] << 8) | ((f >> 5) << 16);
      if (steps == kNegOne) r = -1.0;  // Dynamic-Obstacles' collision reward (dynamicobstacles.py:162-165)
      else r = (j.reward_lut && (int)steps <= j.max_steps) ? j.reward_lut[steps] : reward_of(steps, j.max_steps);
    }
    d.reward[k] = r;
  }
}

void expand_scalar(const ExpandJob &j, const uint8_t *rec, int count, const Dst &d) {
  const uint32_t *lut = tables().lut;
  for (int k = 0; k < count; ++k, rec += kPacked) {
    if (d.obs) {
      uint8_t *o = d.obs + (size_t)k * kObsBytes;
      for (int q = 0; q < kCells - 1; ++q) {  // 4-byte stores, the 4th byte is overwritten by the next triple
        const uint32_t t = lut[rec[q]];
        memcpy(o + 3 * q, &t, 4);
      }
      const uint32_t t = lut[rec[kCells - 1]];
      o[144] = (uint8_t)t; o[145] = (uint8_t)(t >> 8); o[146] = (uint8_t)(t >> 16);
    }
    expand_tail(rec, k, j, d);
  }
}

// 16 codes -> 48 image bytes with three 16-entry byte tables on t4 and byte shuffles for the 3-way interleave
__attribute__((target("ssse3"))) void expand_ssse3(const ExpandJob &j, const uint8_t *rec, int count, const Dst &d) {
  const __m128i type_lut = _mm_setr_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 4, 4, 7, 14, 15);
  const __m128i cfix_lut = _mm_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 5, 0, 0);  // t4 = 13: the colour is the box's (grey)
  const __m128i state_lut = _mm_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 0, 0, 0);
  const __m128i cmask_lut = _mm_setr_epi8(0, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, -1, -1);
  const __m128i low4 = _mm_set1_epi8(15), low3 = _mm_set1_epi8(7);
  // output bytes 0..15 / 16..31 / 32..47 of (t0 c0 s0 t1 c1 s1 ...): shuffle masks per source plane (0x80 = zero)
  const __m128i t_a = _mm_setr_epi8(0, -128, -128, 1, -128, -128, 2, -128, -128, 3, -128, -128, 4, -128, -128, 5);
  const __m128i c_a = _mm_setr_epi8(-128, 0, -128, -128, 1, -128, -128, 2, -128, -128, 3, -128, -128, 4, -128, -128);
  const __m128i s_a = _mm_setr_epi8(-128, -128, 0, -128, -128, 1, -128, -128, 2, -128, -128, 3, -128, -128, 4, -128);
  const __m128i t_b = _mm_setr_epi8(-128, -128, 6, -128, -128, 7, -128, -128, 8, -128, -128, 9, -128, -128, 10, -128);
  const __m128i c_b = _mm_setr_epi8(5, -128, -128, 6, -128, -128, 7, -128, -128, 8, -128, -128, 9, -128, -128, 10);
  const __m128i s_b = _mm_setr_epi8(-128, 5, -128, -128, 6, -128, -128, 7, -128, -128, 8, -128, -128, 9, -128, -128);
  const __m128i t_c = _mm_setr_epi8(-128, 11, -128, -128, 12, -128, -128, 13, -128, -128, 14, -128, -128, 15, -128, -128);
  const __m128i c_c = _mm_setr_epi8(-128, -128, 11, -128, -128, 12, -128, -128, 13, -128, -128, 14, -128, -128, 15, -128);
  const __m128i s_c = _mm_setr_epi8(10, -128, -128, 11, -128, -128, 12, -128, -128, 13, -128, -128, 14, -128, -128, 15);
  const uint32_t *lut = tables().lut;
  for (int k = 0; k < count; ++k, rec += kPacked) {
    if (d.obs) {
      uint8_t *o = d.obs + (size_t)k * kObsBytes;
#pragma GCC unroll 3
      for (int blk = 0; blk < 3; ++blk) {
        const __m128i code = _mm_loadu_si128(reinterpret_cast<const __m128i *>(rec + 16 * blk));
        const __m128i t4 = _mm_and_si128(code, low4);
        const __m128i ty = _mm_shuffle_epi8(type_lut, t4);
        const __m128i st = _mm_shuffle_epi8(state_lut, t4);
        const __m128i co = _mm_or_si128(_mm_and_si128(_mm_and_si128(_mm_srli_epi16(code, 4), low3), _mm_shuffle_epi8(cmask_lut, t4)),
                                        _mm_shuffle_epi8(cfix_lut, t4));
        const __m128i a = _mm_or_si128(_mm_or_si128(_mm_shuffle_epi8(ty, t_a), _mm_shuffle_epi8(co, c_a)), _mm_shuffle_epi8(st, s_a));
        const __m128i b = _mm_or_si128(_mm_or_si128(_mm_shuffle_epi8(ty, t_b), _mm_shuffle_epi8(co, c_b)), _mm_shuffle_epi8(st, s_b));
        const __m128i c = _mm_or_si128(_mm_or_si128(_mm_shuffle_epi8(ty, t_c), _mm_shuffle_epi8(co, c_c)), _mm_shuffle_epi8(st, s_c));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(o + 48 * blk), a);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(o + 48 * blk + 16), b);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(o + 48 * blk + 32), c);
      }
      const uint32_t t = lut[rec[kCells - 1]];
      o[144] = (uint8_t)t; o[145] = (uint8_t)(t >> 8); o[146] = (uint8_t)(t >> 16);
    }
    expand_tail(rec, k, j, d);
  }
}

inline void expand_records(const ExpandJob &j, const uint8_t *rec, int count, const Dst &d) {
  static const bool have_ssse3 = __builtin_cpu_supports("ssse3");
  static const bool force_scalar = getenv("MINIGRID_B200_EXPAND_SCALAR") != nullptr;
  if (have_ssse3 && !force_scalar) expand_ssse3(j, rec, count, d);
  else expand_scalar(j, rec, count, d);
}

// A block of 64 consecutive records whose index is a multiple of 64 fills whole cache lines of every output array
// (64 x 147 B = 147 lines, 64 x 8 B = 8, 64 x 4 B = 4, 64 x 1 B = 1), so it is expanded into a staging area that
// stays in L1 / L2 and streamed out with non-temporal stores: the destination lines are never read for ownership,
// which halves the DRAM traffic of a step whose output (42 MB at 262144 envs) is far larger than the caches.
inline void stream_lines(void *dst, const void *src, size_t bytes) {
  __m128i *d = reinterpret_cast<__m128i *>(dst);
  const __m128i *s = reinterpret_cast<const __m128i *>(src);
  for (size_t i = 0; i < bytes / 16; ++i) _mm_stream_si128(d + i, _mm_load_si128(s + i));
}
void expand_block64(const ExpandJob &j, int64_t i0) {
  alignas(64) uint8_t st_obs[64 * kObsBytes];
  alignas(64) int32_t st_dir[64];
  alignas(64) double st_rew[64];
  alignas(64) uint8_t st_term[64], st_trunc[64];
  const Dst d = {j.obs ? st_obs : nullptr, j.dir ? st_dir : nullptr, j.reward ? st_rew : nullptr, j.term ? st_term : nullptr,
                 j.trunc ? st_trunc : nullptr};
  expand_records(j, j.packed + i0 * kPacked, 64, d);
  if (j.obs) stream_lines(j.obs + i0 * kObsBytes, st_obs, sizeof(st_obs));
  if (j.dir) stream_lines(j.dir + i0, st_dir, sizeof(st_dir));
  if (j.reward) stream_lines(j.reward + i0, st_rew, sizeof(st_rew));
  if (j.term) stream_lines(j.term + i0, st_term, sizeof(st_term));
  if (j.trunc) stream_lines(j.trunc + i0, st_trunc, sizeof(st_trunc));
}

}  // namespace

void expand_range(const ExpandJob &j, int64_t lo, int64_t hi) {
  // Which store form is faster depends on where the caller's arrays live (profiles/r02n_gpu_call.log, 262144 envs, 16
  // threads): one set of output arrays (42 MB) stays in the host's last-level cache and plain stores win (7.4e8 against
  // 5.7e8 env-steps/s end to end); four sets cycled (bench.py's rotating batches) live in DRAM, where plain stores pay a
  // read-for-ownership per line and streaming wins (7.1e8 against 4.5e8). The caller says which (mg_abi.cu calibrates per
  // handle over its first steps); -1 = the environment variable, else plain.
  static const bool env_stream = [] { const char *e = getenv("MINIGRID_B200_EXPAND_STREAM"); return e && atoi(e) != 0; }();
  const bool no_stream = j.stream < 0 ? !env_stream : j.stream == 0;
  auto aligned = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 63u) == 0; };
  const bool stream = !no_stream && aligned(j.obs) && aligned(j.dir) && aligned(j.reward) && aligned(j.term) && aligned(j.trunc);
  auto plain = [&](int64_t a, int64_t b) {
    if (b <= a) return;
    const Dst d = {j.obs ? j.obs + a * kObsBytes : nullptr, j.dir ? j.dir + a : nullptr, j.reward ? j.reward + a : nullptr,
                   j.term ? j.term + a : nullptr, j.trunc ? j.trunc + a : nullptr};
    expand_records(j, j.packed + a * kPacked, (int)(b - a), d);
  };
  if (!stream) {
    for (int64_t a = lo; a < hi; a += 1 << 20) plain(a, a + (1 << 20) < hi ? a + (1 << 20) : hi);
    return;
  }
  int64_t i = lo;
  const int64_t head_end = ((lo + 63) & ~(int64_t)63) < hi ? ((lo + 63) & ~(int64_t)63) : hi;
  plain(i, head_end);
  for (i = head_end; i + 64 <= hi; i += 64) expand_block64(j, i);
  plain(i, hi);
  _mm_sfence();  // the streamed lines are globally visible before the caller is told the range is done
}

int usable_host_threads() {
  int n = (int)std::thread::hardware_concurrency();
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) {
    const int a = CPU_COUNT(&set);
    if (a > 0 && (n <= 0 || a < n)) n = a;
  }
  if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {  // cgroup v2 CPU quota
    char quota[32];
    long period = 0;
    if (fscanf(f, "%31s %ld", quota, &period) == 2 && strcmp(quota, "max") != 0 && period > 0) {
      const long q = (atol(quota) + period - 1) / period;
      if (q > 0 && q < n) n = (int)q;
    }
    fclose(f);
  }
  return n < 1 ? 1 : n;
}

// ---- a small persistent pool: workers park on a condition variable between steps, and inside a step wait (spinning
// briefly) for the chunk the copy engine is still delivering. Work is handed out in SLICES of 1024 records through one
// counter per chunk: a worker the OS has descheduled (the pool may be as wide as the process's CPU quota, and the
// calling thread is polling CUDA events next to it) then holds up one slice, not a sixteenth of every chunk, and the
// calling thread expands slices itself once the last chunk has landed. Counters carry the job's generation, so a worker
// that is late leaving job g can never take (or skip) a slice of job g + 1.
struct HostPool::Impl {
  static constexpr int64_t SLICE = 1024;  // records; a multiple of 64 (whole cache lines of every output array)
  static constexpr int MAX_CHUNKS = 16;
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_start;
  uint64_t generation = 0;
  bool stop = false;
  ExpandJob job{};
  int64_t bounds[MAX_CHUNKS + 1] = {0};  // chunk c = envs [bounds[c], bounds[c + 1])
  int n_chunks = 0;
  int64_t total_slices = 0;
  std::atomic<int> ready{0};                    // chunks whose bytes have arrived on the host
  std::atomic<uint64_t> next[MAX_CHUNKS];       // generation << 32 | next slice of the chunk
  std::atomic<int64_t> finished{0};             // slices of the current job that are done
  std::atomic<uint64_t> gen_fast{0};            // mirrors `generation` for the spinning phase

  static int64_t slices_of(int64_t len) { return (len + SLICE - 1) / SLICE; }

  // take slices of chunk c of job `gen` until there are none left (or the job is no longer `gen`)
  void drain_chunk(const ExpandJob &j, const int64_t *bnd, int c, uint64_t gen) {
    const int64_t lo = bnd[c], hi = bnd[c + 1], n = slices_of(hi - lo);
    for (;;) {
      uint64_t cur = next[c].load(std::memory_order_acquire);
      if ((cur >> 32) != (gen & 0xFFFFFFFFu) || (int64_t)(cur & 0xFFFFFFFFu) >= n) return;
      if (!next[c].compare_exchange_weak(cur, cur + 1, std::memory_order_acq_rel)) continue;
      const int64_t s = (int64_t)(cur & 0xFFFFFFFFu);
      const int64_t a = lo + s * SLICE, b = a + SLICE < hi ? a + SLICE : hi;
      expand_range(j, a, b);
      finished.fetch_add(1, std::memory_order_release);
    }
  }

  void worker() {
    uint64_t seen = 0;
    for (;;) {
      // A training loop calls step after step: the next job arrives tens of microseconds after the last one ended, less
      // than a condition-variable wake-up costs. Spin for it briefly, park only when the caller has gone quiet.
      bool got = false;
      for (int spins = 0; spins < 20000; ++spins) {  // ~0.2-0.4 ms
        if (gen_fast.load(std::memory_order_acquire) != seen) { got = true; break; }
        _mm_pause();
      }
      ExpandJob j;
      int64_t bnd[MAX_CHUNKS + 1];
      int nc;
      {
        std::unique_lock<std::mutex> lk(mu);
        if (!got) cv_start.wait(lk, [&] { return stop || generation != seen; });
        if (stop) return;
        seen = generation;
        j = job; nc = n_chunks;
        for (int c = 0; c <= nc; ++c) bnd[c] = bounds[c];
      }
      for (int c = 0; c < nc; ++c) {
        int spins = 0;
        bool stale = false;
        while (ready.load(std::memory_order_acquire) <= c) {
          if (gen_fast.load(std::memory_order_acquire) != seen) { stale = true; break; }  // the job ended without us
          if (++spins < 2000) _mm_pause();
          else { std::this_thread::yield(); spins = 0; }
        }
        if (stale) break;
        drain_chunk(j, bnd, c, seen);
      }
    }
  }
};

HostPool::HostPool(int n_threads) : impl_(new Impl), n_threads_(n_threads < 1 ? 1 : n_threads) {
  for (int c = 0; c < Impl::MAX_CHUNKS; ++c) impl_->next[c].store(0);
  for (int w = 0; w < n_threads_; ++w) impl_->threads.emplace_back([this] { impl_->worker(); });
}
HostPool::~HostPool() {
  {
    std::lock_guard<std::mutex> lk(impl_->mu);
    impl_->stop = true;
  }
  impl_->cv_start.notify_all();
  for (auto &t : impl_->threads) t.join();
  delete impl_;
}
void HostPool::begin(const ExpandJob &job, const int64_t *bounds, int n_chunks) {
  std::lock_guard<std::mutex> lk(impl_->mu);
  if (n_chunks > Impl::MAX_CHUNKS) n_chunks = Impl::MAX_CHUNKS;  // (callers never ask for more)
  impl_->job = job;
  impl_->n_chunks = n_chunks;
  impl_->total_slices = 0;
  impl_->generation += 1;
  for (int c = 0; c <= n_chunks; ++c) impl_->bounds[c] = bounds[c];
  for (int c = 0; c < n_chunks; ++c) {
    impl_->total_slices += Impl::slices_of(bounds[c + 1] - bounds[c]);
    impl_->next[c].store((impl_->generation & 0xFFFFFFFFu) << 32, std::memory_order_release);
  }
  impl_->ready.store(0, std::memory_order_release);
  impl_->finished.store(0, std::memory_order_release);
  impl_->gen_fast.store(impl_->generation, std::memory_order_release);
  impl_->cv_start.notify_all();
}
void HostPool::chunk_ready() { impl_->ready.fetch_add(1, std::memory_order_release); }
void HostPool::abort_chunks(int n_chunks) { impl_->ready.store(n_chunks, std::memory_order_release); }
void HostPool::wait() {
  // every chunk has been released: the calling thread takes slices too, then waits for the slices still in other hands
  const uint64_t gen = impl_->generation;  // (only this thread starts jobs)
  for (int c = 0; c < impl_->n_chunks; ++c) impl_->drain_chunk(impl_->job, impl_->bounds, c, gen);
  int spins = 0;
  while (impl_->finished.load(std::memory_order_acquire) < impl_->total_slices) {
    if (++spins < 4000) _mm_pause();
    else { std::this_thread::yield(); spins = 0; }
  }
}

}  // namespace mg

extern "C" int mg_expand_packed(const uint8_t *packed, int64_t n_envs, int32_t max_steps, uint8_t *obs, int32_t *dir,
                                double *reward, uint8_t *terminated, uint8_t *truncated) {
  if (!packed || n_envs < 0 || max_steps < 1) return MG_ERR_INVALID_ARG;
  mg::ExpandJob j{};
  j.packed = packed; j.max_steps = max_steps; j.reward_lut = nullptr; j.stream = -1;
  j.obs = obs; j.dir = dir; j.reward = reward; j.term = terminated; j.trunc = truncated;
  mg::expand_range(j, 0, n_envs);
  return MG_OK;
}

extern "C" int mg_expand_packed_mt(const uint8_t *packed, int64_t n_envs, int32_t max_steps, uint8_t *obs, int32_t *dir,
                                   double *reward, uint8_t *terminated, uint8_t *truncated, int n_threads) {
  if (!packed || n_envs < 0 || max_steps < 1) return MG_ERR_INVALID_ARG;
  static mg::HostPool *pool = nullptr;  // one shared pool for this convenience entry point (not thread-safe, like a handle)
  const int want = n_threads > 0 ? n_threads : mg::usable_host_threads();
  if (!pool || pool->threads() != want) { delete pool; pool = new mg::HostPool(want); }
  mg::ExpandJob j{};
  j.packed = packed; j.max_steps = max_steps; j.reward_lut = nullptr; j.stream = -1;
  j.obs = obs; j.dir = dir; j.reward = reward; j.term = terminated; j.trunc = truncated;
  const int64_t bounds[2] = {0, n_envs};
  pool->begin(j, bounds, 1);
  pool->chunk_ready();
  pool->wait();
  return MG_OK;
}
