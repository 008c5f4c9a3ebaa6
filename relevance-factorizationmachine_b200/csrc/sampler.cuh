// sampler.cuh -- the perf-mode batch sampler: a keyed Feistel bijection of [0, N) with cycle
// walking. Not part of the reference (its sampler is sklearn.utils.resample, src/fm.py:72-79,
// reproduced on the host in sampler.cu); specification: oracle/sampler_oracle.py.
#pragma once
#include <cstdint>

namespace rfm {

constexpr int FEISTEL_ROUNDS = 6;

struct FeistelKey {
  uint32_t k[FEISTEL_ROUNDS];
  uint32_t half_bits;
  uint32_t mask;
  uint64_t n_rows;
};

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

inline FeistelKey make_feistel_key(uint64_t n_rows, uint32_t seed, uint32_t epoch) {
  FeistelKey key;
  const uint32_t base = seed * 0x9E3779B9u + epoch * 0x7F4A7C15u;
  for (int r = 0; r < FEISTEL_ROUNDS; ++r) key.k[r] = mix32(base + (uint32_t)(r + 1) * 0x632BE5ABu);
  uint32_t bits = 0;
  for (uint64_t v = n_rows > 0 ? n_rows - 1 : 0; v; v >>= 1) ++bits;
  if (bits < 2) bits = 2;
  key.half_bits = (bits + 1) / 2;
  key.mask = (1u << key.half_bits) - 1u;
  key.n_rows = n_rows;
  return key;
}

__host__ __device__ __forceinline__ uint64_t feistel_permute(uint64_t q, const FeistelKey &key) {
  uint64_t x = q;
  do {
    uint32_t L = (uint32_t)(x >> key.half_bits), R = (uint32_t)x & key.mask;
#pragma unroll
    for (int r = 0; r < FEISTEL_ROUNDS; ++r) {
      const uint32_t f = mix32(R ^ key.k[r]) & key.mask;
      const uint32_t t = L ^ f;
      L = R;
      R = t;
    }
    x = ((uint64_t)L << key.half_bits) | R;
  } while (x >= key.n_rows);
  return x;
}

}  // namespace rfm
