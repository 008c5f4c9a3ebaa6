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
#include <algorithm>
#include <cstdlib>
#include <thread>

#include "common.cuh"
#include "sampler.cuh"

namespace {

// MT19937 with the tempered outputs of a whole refill produced in one (vectorisable) loop.
struct MT19937 {
  uint32_t mt[624], out[624];
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
      const uint32_t y = (mt[i] & UPPER) | (mt[i + 1] & LOWER);
      mt[i] = mt[i + 397] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAGIC);
    }
    for (; i < 623; ++i) {
      const uint32_t y = (mt[i] & UPPER) | (mt[i + 1] & LOWER);
      mt[i] = mt[i + (397 - 624)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAGIC);
    }
    const uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
    mt[623] = mt[396] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAGIC);
    for (i = 0; i < 624; ++i) {
      uint32_t t = mt[i];
      t ^= (t >> 11);
      t ^= (t << 7) & 0x9d2c5680u;
      t ^= (t << 15) & 0xefc60000u;
      t ^= (t >> 18);
      out[i] = t;
    }
    pos = 0;
  }
  inline uint32_t next() {
    if (pos == 624) refill();
    return out[pos++];
  }
};

// Swap partners of the shuffle steps i, i-1, ..., i-nb+1 (js[t] belongs to step i-t): NumPy's masked rejection
// sampling, written without a data-dependent branch -- a rejected draw is simply overwritten by the next one.
// (The branchy do/while mispredicts on a quarter of the draws: 170 ms instead of 60 ms for 12 M steps.)
// `mask` is the smallest 2^b - 1 >= the current bound; it only changes when the bound halves.
inline void draw_partners(MT19937 &rng, uint32_t &mask, int64_t i, int nb, uint32_t *js) {
  int cnt = 0;
  uint32_t ui = (uint32_t)i;
  while (cnt < nb) {
    if (mask == 0 || ui <= (mask >> 1)) {
      mask = ui;
      mask |= mask >> 1;
      mask |= mask >> 2;
      mask |= mask >> 4;
      mask |= mask >> 8;
      mask |= mask >> 16;
    }
    const uint32_t half = mask >> 1;
    while (cnt < nb && ui > half) {
      const uint32_t c = rng.next() & mask;
      const uint32_t ok = c <= ui ? 1u : 0u;
      js[cnt] = c;
      cnt += ok;
      ui -= ok;
    }
  }
}

// position -> batch slot: open addressing, linear probing, backward-shift deletion (no tombstones). The number of
// live entries never changes (every update moves one entry), so the load factor stays below one half.
struct PosMap {
  static constexpr uint32_t EMPTY = 0xFFFFFFFFu;
  std::vector<uint32_t> key, val;
  uint32_t cmask;
  explicit PosMap(uint64_t live) {
    uint64_t cap = 16;
    while (cap < 2 * live + 2) cap <<= 1;
    key.assign((size_t)cap, EMPTY);
    val.resize((size_t)cap);
    cmask = (uint32_t)(cap - 1);
  }
  static inline uint32_t hash(uint32_t k) {
    k *= 0x9E3779B1u;
    return k ^ (k >> 15);
  }
  inline void insert(uint32_t k, uint32_t v) {
    uint32_t s = hash(k) & cmask;
    while (key[s] != EMPTY) s = (s + 1) & cmask;
    key[s] = k;
    val[s] = v;
  }
  inline uint32_t *find(uint32_t k) {   // k must be present
    uint32_t s = hash(k) & cmask;
    while (key[s] != k) s = (s + 1) & cmask;
    return &val[s];
  }
  inline uint32_t erase(uint32_t k) {   // k must be present
    uint32_t s = hash(k) & cmask;
    while (key[s] != k) s = (s + 1) & cmask;
    const uint32_t v = val[s];
    uint32_t hole = s, t = (s + 1) & cmask;
    while (key[t] != EMPTY) {
      const uint32_t home = hash(key[t]) & cmask;
      // the entry at t may fill the hole unless its home slot lies cyclically in (hole, t]
      if (((t - home) & cmask) >= ((t - hole) & cmask)) {
        key[hole] = key[t];
        val[hole] = val[t];
        hole = t;
      }
      t = (t + 1) & cmask;
    }
    key[hole] = EMPTY;
    return v;
  }
};

// The whole shuffle on an index array (batch is a large share of the rows). Partners are drawn a block ahead and
// their cache lines requested before the swaps run in NumPy's order.
void shuffle_array(int64_t n_rows, uint32_t epoch, int32_t *idx) {
  for (int64_t i = 0; i < n_rows; ++i) idx[i] = (int32_t)i;
  MT19937 rng(epoch);
  uint32_t mask = 0;
  constexpr int BLOCK = 64;
  uint32_t js[BLOCK];
  int64_t i = n_rows - 1;
  while (i >= 1) {
    const int nb = (int)(i < BLOCK ? i : BLOCK);
    draw_partners(rng, mask, i, nb, js);
    for (int t = 0; t < nb; ++t) __builtin_prefetch(idx + js[t], 1, 0);
    for (int t = 0; t < nb; ++t) {
      const int32_t v = idx[i - t];
      idx[i - t] = idx[js[t]];
      idx[js[t]] = v;
    }
    i -= nb;
  }
}

// Only the first `batch` entries of the shuffled array are wanted. Instead of permuting a 4 n_rows-byte array
// (a cache miss per step once it outgrows the caches), all partners are drawn (sequential writes), then the swaps
// are undone last-to-first while following just the `batch` positions of interest: position p of the final array
// holds the element that sat at the swap-undone position before the step, and the initial array is arange.
// Going backwards i grows from 1 to n_rows-1 and a followed position is never above the current i, so for
// i >= batch only "is js followed?" must be asked -- a bitmap of n_rows bits (cache resident) answers it, and
// the expected number of hits is batch * ln(n_rows / batch), each a hash-map move. Same draws, same result, bit
// for bit. Measured (build container): 12 M rows, batch 65,536: 118 ms against 147 ms for the prefetched array
// path; 4 M rows, same batch: 48 against 37 ms -- hence the size rule in rfm_legacy_batch.
void shuffle_prefix(int64_t n_rows, int64_t batch, uint32_t epoch, int64_t *out_rows, uint32_t *js) {
  MT19937 rng(epoch);
  uint32_t mask = 0;
  {
    constexpr int BLOCK = 256;
    int64_t i = n_rows - 1, w = 0;
    while (i >= 1) {
      const int nb = (int)(i < BLOCK ? i : BLOCK);
      draw_partners(rng, mask, i, nb, js + w);        // js[s] is the partner of step i = n_rows-1-s
      w += nb;
      i -= nb;
    }
  }
  std::vector<uint64_t> followed((size_t)((n_rows + 63) / 64), 0);
  auto test = [&](uint32_t p) { return (followed[p >> 6] >> (p & 63)) & 1u; };
  auto set = [&](uint32_t p) { followed[p >> 6] |= 1ull << (p & 63); };
  auto clear = [&](uint32_t p) { followed[p >> 6] &= ~(1ull << (p & 63)); };
  PosMap slot_of((uint64_t)batch);
  for (int64_t p = 0; p < batch; ++p) {
    set((uint32_t)p);
    slot_of.insert((uint32_t)p, (uint32_t)p);
  }
  int64_t i = 1;
  for (; i < n_rows && i < batch; ++i) {              // both ends of the swap may be followed
    const uint32_t j = js[n_rows - 1 - i], ui = (uint32_t)i;
    if (j == ui) continue;
    const bool fi = test(ui), fj = test(j);
    if (fi && fj) {
      uint32_t *a = slot_of.find(ui), *b = slot_of.find(j);
      const uint32_t t = *a;
      *a = *b;
      *b = t;
    } else if (fi) {
      slot_of.insert(j, slot_of.erase(ui));
      clear(ui);
      set(j);
    } else if (fj) {
      slot_of.insert(ui, slot_of.erase(j));
      clear(j);
      set(ui);
    }
  }
  for (; i < n_rows; ++i) {                           // position i itself cannot be followed yet
    const uint32_t j = js[n_rows - 1 - i];
    if (test(j)) {
      const uint32_t ui = (uint32_t)i;
      if (j == ui) continue;
      slot_of.insert(ui, slot_of.erase(j));
      clear(j);
      set(ui);
    }
  }
  for (size_t s = 0; s < slot_of.key.size(); ++s)
    if (slot_of.key[s] != PosMap::EMPTY) out_rows[slot_of.val[s]] = (int64_t)slot_of.key[s];
}

// fn(begin, end) over [0, n) on up to n_threads host threads (contiguous ranges; the caller's thread takes one)
template <typename F>
void parallel_ranges(int64_t n, int n_threads, F fn) {
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n / 4096));
  if (n_threads <= 1) {
    fn((int64_t)0, n);
    return;
  }
  std::vector<std::thread> pool;
  const int64_t per = (n + n_threads - 1) / n_threads;
  for (int t = 1; t < n_threads; ++t) {
    const int64_t b = std::min(n, t * per), e = std::min(n, (t + 1) * per);
    if (b < e) pool.emplace_back([=] { fn(b, e); });
  }
  fn((int64_t)0, std::min(n, per));
  for (auto &th : pool) th.join();
}

}  // namespace

using namespace rfm;

extern "C" {

int rfm_legacy_batch(int64_t n_rows, int64_t batch, uint32_t epoch, int64_t *out_rows, int32_t *scratch) {
  RFM_REQUIRE(n_rows >= 0 && batch >= 0 && out_rows != nullptr, "rfm_legacy_batch: bad argument");
  RFM_REQUIRE(batch <= n_rows,
              "Cannot sample %lld out of arrays with dim %lld when replace is False",
              (long long)batch, (long long)n_rows);
  RFM_REQUIRE(n_rows <= 0x7fffffffLL, "rfm_legacy_batch: at most 2^31-1 rows");
  if (batch == 0) return RFM_OK;
  std::vector<int32_t> own;
  int32_t *buf = scratch;
  if (!buf) {
    own.resize((size_t)n_rows);
    buf = own.data();
  }
  const char *force = getenv("RFM_LEGACY_SAMPLER_PATH");     // "array" / "prefix": tests and measurements
  const bool prefix = force ? force[0] == 'p' : (n_rows >= (1 << 23) && batch * 64 <= n_rows);
  if (prefix) {
    shuffle_prefix(n_rows, batch, epoch, out_rows, reinterpret_cast<uint32_t *>(buf));
  } else {
    shuffle_array(n_rows, epoch, buf);
    for (int64_t q = 0; q < batch; ++q) out_rows[q] = buf[q];
  }
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


int rfm_feistel_batches(int64_t n_rows, int64_t batch, uint32_t seed, uint32_t epoch0, int32_t n_epochs,
                        int64_t begin, int64_t count, int64_t *out_rows, int32_t n_threads) {
  RFM_REQUIRE(n_rows >= 0 && batch >= 0 && n_epochs >= 0 && out_rows != nullptr, "rfm_feistel_batches: bad argument");
  RFM_REQUIRE(batch <= n_rows,
              "Cannot sample %lld out of arrays with dim %lld when replace is False",
              (long long)batch, (long long)n_rows);
  RFM_REQUIRE(begin >= 0 && count >= 0 && begin + count <= batch,
              "rfm_feistel_batches: slice [%lld, %lld) outside the batch of %lld", (long long)begin,
              (long long)(begin + count), (long long)batch);
  RFM_REQUIRE(n_rows <= (1LL << 32), "rfm_feistel_batches: at most 2^32 rows");
  const int64_t total = (int64_t)n_epochs * count;
  parallel_ranges(total, n_threads, [=](int64_t b, int64_t e) {
    int64_t cur_epoch = -1;
    FeistelKey key = make_feistel_key((uint64_t)n_rows, seed, epoch0);
    for (int64_t q = b; q < e; ++q) {
      const int64_t ep = q / count, pos = q - ep * count;
      if (ep != cur_epoch) {
        key = make_feistel_key((uint64_t)n_rows, seed, epoch0 + (uint32_t)ep);
        cur_epoch = ep;
      }
      out_rows[q] = (int64_t)feistel_permute((uint64_t)(begin + pos), key);
    }
  });
  return RFM_OK;
}

int rfm_csr_gather_rows(int64_t n_rows, const void *indptr, int indptr_is_int64, const int32_t *indices,
                        const double *data, const int64_t *labels, const double *pscores, const int64_t *rows,
                        int64_t n_sel, int64_t *out_indptr, int32_t *out_indices, double *out_data,
                        int64_t *out_labels, double *out_pscores, int32_t n_threads) {
  RFM_REQUIRE(n_rows >= 0 && n_sel >= 0 && indptr && out_indptr && (rows || n_sel == 0),
              "rfm_csr_gather_rows: bad argument");
  RFM_REQUIRE((labels == nullptr) == (pscores == nullptr), "rfm_csr_gather_rows: labels and pscores go together");
  const int64_t *p64 = static_cast<const int64_t *>(indptr);
  const int32_t *p32 = static_cast<const int32_t *>(indptr);
  auto ptr_at = [=](int64_t i) -> int64_t { return indptr_is_int64 ? p64[i] : (int64_t)p32[i]; };
  // pass 1 (always): row lengths -> out_indptr (exclusive prefix sum, serial: n_sel adds)
  out_indptr[0] = 0;
  for (int64_t q = 0; q < n_sel; ++q) {
    const int64_t r = rows[q];
    RFM_REQUIRE(r >= 0 && r < n_rows, "rfm_csr_gather_rows: row id %lld outside [0, %lld)", (long long)r,
                (long long)n_rows);
    const int64_t len = ptr_at(r + 1) - ptr_at(r);
    RFM_REQUIRE(len >= 0, "rfm_csr_gather_rows: indptr decreases at row %lld", (long long)r);
    out_indptr[q + 1] = out_indptr[q] + len;
  }
  if (!out_indices) return RFM_OK;   // sizing call: the caller allocates out_indptr[n_sel] entries and calls again
  RFM_REQUIRE(out_data && (out_indptr[n_sel] == 0 || (indices && data)), "rfm_csr_gather_rows: NULL data arrays");
  RFM_REQUIRE(!labels || (out_labels && out_pscores), "rfm_csr_gather_rows: NULL label outputs");
  parallel_ranges(n_sel, n_threads, [=](int64_t b, int64_t e) {
    constexpr int AHEAD = 8;        // the rows are random: request the source lines a few rows early
    for (int64_t q = b; q < e; ++q) {
      if (q + AHEAD < e) {
        const int64_t z = ptr_at(rows[q + AHEAD]);
        __builtin_prefetch(indices + z, 0, 0);
        __builtin_prefetch(data + z, 0, 0);
        __builtin_prefetch(data + z + 8, 0, 0);
      }
      const int64_t r = rows[q], z0 = ptr_at(r), len = ptr_at(r + 1) - z0, o = out_indptr[q];
      memcpy(out_indices + o, indices + z0, (size_t)len * sizeof(int32_t));
      memcpy(out_data + o, data + z0, (size_t)len * sizeof(double));
      if (labels) {
        out_labels[q] = labels[r];
        out_pscores[q] = pscores[r];
      }
    }
  });
  return RFM_OK;
}

}  // extern "C"
