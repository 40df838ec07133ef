// mg_host_shim.h — lets g++ compile the engine's per-lane headers for tests/host_emu (TEST BUILD ONLY; nvcc
// never sees this file and the product library contains no host implementation of the hot path).
#pragma once
#include <stdint.h>

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }

static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {  // PTX prmt.b32, default mode
  const uint64_t ab = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t n = (sel >> (4 * i)) & 15u;
    uint32_t byte = (uint32_t)((ab >> (8 * (n & 7u))) & 0xFFu);
    if (n & 8u) byte = (byte & 0x80u) ? 0xFFu : 0x00u;  // replicate the sign bit
    r |= byte << (8 * i);
  }
  return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {  // shf.r.wrap
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> (shift & 31u));
}
static inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t shift) {  // shf.r.clamp
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> (shift > 32u ? 32u : shift));
}
static inline uint32_t __brev(uint32_t v) {
  uint32_t r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
static inline int __ffs(uint32_t v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
