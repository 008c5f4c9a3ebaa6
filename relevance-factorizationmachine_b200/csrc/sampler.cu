// sampler.cu -- host-side batch samplers.
//
// rfm_legacy_batch reproduces, bit for bit, what the reference's call
//   sklearn.utils.resample(X, y, ps, replace=False, n_samples=B, random_state=epoch)
// (src/fm.py:72-79, src/mf.py:88-95) selects:  idx = arange(N);
// RandomState(epoch).shuffle(idx); idx[:B].  That is MT19937 seeded with init_genrand(epoch),
// and NumPy's legacy shuffle: for i = N-1 .. 1: j = rk_interval(i); swap(idx[i], idx[j]),
// where rk_interval draws 32-bit outputs masked to the smallest 2^b-1 >= i and rejects
// values > i. The algorithm is NumPy's published one (numpy/random/mtrand.pyx `shuffle`,
// `_legacy_seeding`; src/distributions `random_interval`); tests check it against NumPy.
#include "common.cuh"
#include "sampler.cuh"

namespace {

struct MT19937 {
  uint32_t mt[624];
  int pos;
  explicit MT19937(uint32_t seed) {
    for (int i = 0; i < 624; ++i) {
      mt[i] = seed;
      seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)(i + 1);
    }
    pos = 624;
  }
  void refill() {
    constexpr uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAGIC = 0x9908b0dfu;
    int i;
    for (i = 0; i < 624 - 397; ++i) {
      uint32_t y = (mt[i] & UPPER) | (mt[i + 1] & LOWER);
      mt[i] = mt[i + 397] ^ (y >> 1) ^ ((y & 1u) ? MAGIC : 0u);
    }
    for (; i < 623; ++i) {
      uint32_t y = (mt[i] & UPPER) | (mt[i + 1] & LOWER);
      mt[i] = mt[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MAGIC : 0u);
    }
    uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MAGIC : 0u);
    pos = 0;
  }
  inline uint32_t next() {
    if (pos == 624) refill();
    uint32_t y = mt[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
};

}  // namespace

using namespace rfm;

extern "C" {

int rfm_legacy_batch(int64_t n_rows, int64_t batch, uint32_t epoch, int64_t *out_rows, int32_t *scratch) {
  RFM_REQUIRE(n_rows >= 0 && batch >= 0 && out_rows != nullptr, "rfm_legacy_batch: bad argument");
  RFM_REQUIRE(batch <= n_rows,
              "Cannot sample %lld out of arrays with dim %lld when replace is False",
              (long long)batch, (long long)n_rows);
  RFM_REQUIRE(n_rows <= 0x7fffffffLL, "rfm_legacy_batch: at most 2^31-1 rows");
  std::vector<int32_t> own;
  int32_t *idx = scratch;
  if (!idx) {
    own.resize((size_t)n_rows);
    idx = own.data();
  }
  for (int64_t i = 0; i < n_rows; ++i) idx[i] = (int32_t)i;
  MT19937 rng(epoch);
  uint32_t mask = 0;
  for (int64_t i = n_rows - 1; i >= 1; --i) {
    // smallest bit mask >= i; i only decreases, so recompute when i drops below a power of two
    const uint32_t ui = (uint32_t)i;
    if (mask == 0 || ui <= (mask >> 1)) {
      mask = ui;
      mask |= mask >> 1;
      mask |= mask >> 2;
      mask |= mask >> 4;
      mask |= mask >> 8;
      mask |= mask >> 16;
    }
    uint32_t j;
    do {
      j = rng.next() & mask;
    } while (j > ui);
    const int32_t t = idx[i];
    idx[i] = idx[j];
    idx[j] = t;
  }
  for (int64_t q = 0; q < batch; ++q) out_rows[q] = idx[q];
  return RFM_OK;
}

int rfm_feistel_batch(int64_t n_rows, int64_t batch, uint32_t seed, uint32_t epoch, int64_t *out_rows) {
  RFM_REQUIRE(n_rows >= 0 && batch >= 0 && out_rows != nullptr, "rfm_feistel_batch: bad argument");
  RFM_REQUIRE(batch <= n_rows,
              "Cannot sample %lld out of arrays with dim %lld when replace is False",
              (long long)batch, (long long)n_rows);
  RFM_REQUIRE(n_rows <= (1LL << 32), "rfm_feistel_batch: at most 2^32 rows");
  const FeistelKey key = make_feistel_key((uint64_t)n_rows, seed, epoch);
  for (int64_t q = 0; q < batch; ++q) out_rows[q] = (int64_t)feistel_permute((uint64_t)q, key);
  return RFM_OK;
}

}  // extern "C"
