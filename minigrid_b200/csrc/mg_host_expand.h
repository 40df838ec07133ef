// mg_host_expand.h — host-side expansion of the packed step records (MG_HOST_PACKED), see mg_host_expand.cpp.
#pragma once
#include <stdint.h>

namespace mg {

struct ExpandJob {
  const uint8_t *packed;     // [n][52]
  int max_steps;
  const double *reward_lut;  // [max_steps + 1] host copy of the reward table, or NULL (evaluate the expression)
  uint8_t *obs;              // [n][147] or NULL
  int32_t *dir;              // [n] or NULL
  double *reward;            // [n] or NULL
  uint8_t *term, *trunc;     // [n] or NULL
  int stream;                // 1: non-temporal stores through a 64-record staging block (outputs that live in DRAM),
                             // 0: plain stores (outputs that stay in the last-level cache); -1: MINIGRID_B200_EXPAND_STREAM or plain
};

// expands records [lo, hi)
void expand_range(const ExpandJob &job, int64_t lo, int64_t hi);
// CPUs this process may use at once: min(affinity mask, cgroup v2 quota)
int usable_host_threads();

class HostPool {
 public:
  explicit HostPool(int n_threads);
  ~HostPool();
  int threads() const { return n_threads_; }
  // start a step: the workers expand chunk c = records [bounds[c], bounds[c + 1]) as soon as chunk_ready() has been
  // called c + 1 times; wait() returns when every worker has finished the last chunk
  void begin(const ExpandJob &job, const int64_t *bounds, int n_chunks);
  void chunk_ready();
  void abort_chunks(int n_chunks);  // error path: release the workers
  void wait();

 private:
  struct Impl;
  Impl *impl_;
  int n_threads_;
};

}  // namespace mg
