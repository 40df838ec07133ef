// mg_pcg64.cuh — numpy's random stack, restated for the device, bit-exact:
//   SeedSequence(entropy=int)  -> 4 x uint64 words        (numpy/random/bit_generator.pyx, published algorithm)
//   PCG64 (XSL-RR 128/64, "setseq" stream)                 (numpy/random/src/pcg64/pcg64.h)
//   next_uint32: low half first, high half buffered        (pcg64_next32)
//   Generator.integers(lo, hi) for ranges < 2^32: Lemire   (buffered_bounded_lemire_uint32; no draw if range 1)
//   Generator.shuffle(list): masked rejection per swap     (random_interval)
// numpy is a dependency of the reference (pyproject.toml:28, numpy>=1.18; 2.3.5 in this image) and is not
// under /root/reference; the call sites are MiniGridEnv._rand_int (minigrid_env.py:247-252) and
// crossing.py:154,167,176,181. Pinned by tests/test_oracle_rng.py (oracle vs numpy) and the GPU parity tests.
#pragma once
#include "mg_common.cuh"

namespace mg {

typedef unsigned __int128 u128;

struct Pcg {
  u128 state, inc;
  uint32_t has_uint32, uinteger;
};

MG_D u128 pcg_mult() {
  return ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL;
}
MG_D void pcg_step(Pcg &r) { r.state = r.state * pcg_mult() + r.inc; }
MG_D uint64_t pcg_next64(Pcg &r) {
  pcg_step(r);
  const uint64_t hi = (uint64_t)(r.state >> 64), lo = (uint64_t)r.state;
  const uint64_t x = hi ^ lo;
  const unsigned rot = (unsigned)(hi >> 58);
  return (x >> rot) | (x << ((64u - rot) & 63u));
}
MG_D uint32_t pcg_next32(Pcg &r) {
  if (r.has_uint32) { r.has_uint32 = 0; return r.uinteger; }
  const uint64_t n = pcg_next64(r);
  r.has_uint32 = 1;
  r.uinteger = (uint32_t)(n >> 32);
  return (uint32_t)n;
}
MG_D int rng_integers(Pcg &r, int low, int high) {
  const uint32_t rng = (uint32_t)(high - 1 - low);
  if (rng == 0) return low;
  const uint32_t rng_excl = rng + 1u;
  uint64_t m = (uint64_t)pcg_next32(r) * rng_excl;
  uint32_t leftover = (uint32_t)m;
  if (leftover < rng_excl) {
    const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
    while (leftover < threshold) {
      m = (uint64_t)pcg_next32(r) * rng_excl;
      leftover = (uint32_t)m;
    }
  }
  return low + (int)(m >> 32);
}
MG_D uint32_t rng_interval(Pcg &r, uint32_t max) {
  if (max == 0) return 0;
  uint32_t mask = max;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t v;
  while ((v = (pcg_next32(r) & mask)) > max) {}
  return v;
}

MG_D Pcg load_rng(const RngRec *rec) {
  Pcg r;
  r.state = ((u128)rec->state_hi << 64) | rec->state_lo;
  r.inc = ((u128)rec->inc_hi << 64) | rec->inc_lo;
  r.has_uint32 = rec->has_uint32;
  r.uinteger = rec->uinteger;
  return r;
}
MG_D void store_rng(RngRec *rec, const Pcg &r) {
  rec->state_hi = (uint64_t)(r.state >> 64); rec->state_lo = (uint64_t)r.state;
  rec->inc_hi = (uint64_t)(r.inc >> 64); rec->inc_lo = (uint64_t)r.inc;
  rec->has_uint32 = r.has_uint32; rec->uinteger = r.uinteger;
}

// SeedSequence(seed).generate_state(4, uint64) -> pcg64_set_seed
MG_D uint32_t ss_hashmix(uint32_t value, uint32_t &hash_const) {
  value ^= hash_const;
  hash_const *= 0x931e8875u;
  value *= hash_const;
  value ^= value >> 16;
  return value;
}
MG_D uint32_t ss_mix(uint32_t x, uint32_t y) {
  uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
  r ^= r >> 16;
  return r;
}
MG_D Pcg seed_pcg64(uint64_t seed) {
  const uint32_t e0 = (uint32_t)seed, e1 = (uint32_t)(seed >> 32);
  uint32_t pool[4];
  uint32_t hc = 0x43b0d7e5u;
  pool[0] = ss_hashmix(e0, hc);
  pool[1] = ss_hashmix(e1, hc);  // a missing second entropy word hashes as 0, which is what e1 == 0 is
  pool[2] = ss_hashmix(0u, hc);
  pool[3] = ss_hashmix(0u, hc);
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int d = 0; d < 4; ++d)
      if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], hc));
  uint32_t out[8];
  uint32_t hb = 0x8b51f9ddu;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t v = pool[i & 3];
    v ^= hb;
    hb *= 0x58f38dedu;
    v *= hb;
    v ^= v >> 16;
    out[i] = v;
  }
  const uint64_t w0 = (uint64_t)out[0] | ((uint64_t)out[1] << 32), w1 = (uint64_t)out[2] | ((uint64_t)out[3] << 32);
  const uint64_t w2 = (uint64_t)out[4] | ((uint64_t)out[5] << 32), w3 = (uint64_t)out[6] | ((uint64_t)out[7] << 32);
  Pcg r;
  r.state = 0;
  r.inc = ((((u128)w2 << 64) | w3) << 1) | 1;
  pcg_step(r);
  r.state += ((u128)w0 << 64) | w1;
  pcg_step(r);
  r.has_uint32 = 0;
  r.uinteger = 0;
  return r;
}

}  // namespace mg
