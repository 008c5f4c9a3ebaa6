// rank.cu -- per-user ranking metrics (reference utils/evaluate.py, utils/metrics.py).
//
//   group-by     utils/evaluate.py:129-156, 209-239   interaction_df.groupby("user"): users ascending,
//                                                     rows inside a user in original order
//   ranking      utils/evaluate.py:93, 197            scores.argsort()[::-1]; canonical tie rule here:
//                                                     score descending, later row first among exact ties
//                                                     (== argsort(kind="stable")[::-1])
//   skip rule    utils/evaluate.py:98-99, 201-202     users whose labels sum to 0 contribute nothing
//   DCG@k        utils/metrics.py:83-107    IPS-DCG@k  utils/metrics.py:53-80
//   ME@k         utils/metrics.py:110-127 (nan when the list is shorter than k; nanmean'd by the caller)
//   Recall@k     utils/metrics.py:32-50     MAP@k      utils/metrics.py:9-29
//   coverage     utils/metrics.py:152-166   |union of kept users' top-k items| (integer, exact)
//
// One CTA per user sorts the user's list in shared memory (bitonic; order-preserving 64-bit score keys with NaN
// first, the in-list position breaks ties, so the order is total), then one thread per K walks
// the ranked prefix and emits the per-user metric terms. A second kernel sums the per-user terms
// in a fixed order, so the aggregates are bit-reproducible. Item hit counts use integer atomics.
#include <algorithm>
#include <numeric>

#include "common.cuh"

using namespace rfm;

struct rfm_ranker {
  rfm_ctx *ctx = nullptr;
  int64_t n_rows = 0, n_users = 0, n_items = 0;
  DevBuf<int32_t> order;       // grouped position -> original row id
  DevBuf<int64_t> user_ptr;    // [n_users + 1]
  DevBuf<int32_t> item;        // grouped
  DevBuf<double> label, pscore;
  DevBuf<double> scores;       // original row order
  DevBuf<int32_t> K_dev;
  DevBuf<double> per_user;     // [n_users][n_k][NTERMS]
  DevBuf<double> metrics;      // [n_k][RFM_RANK_NCOLS]
  DevBuf<int32_t> hits;        // [n_k][n_items]
  DevBuf<int64_t> top_rows;    // [n_users][k_max]
  DevBuf<double> totals;       // optional per-user label totals supplied by the caller
  bool has_totals = false;
  DevBuf<double> history;      // [slots][RK_HISTORY_K][RFM_RANK_NCOLS]: metrics kept on the device (rfm_ranker_evaluate_dev)
  int64_t history_slots = 0;
};

// Full-catalog evaluation state (rfm_catalog_eval_*): the held-out labels as a CSR by user (items ascending,
// duplicates summed, like scipy's csr_matrix((label, (user, item)))), label totals per user, item exposures.
struct rfm_catalog_eval {
  rfm_ctx *ctx = nullptr;
  int64_t n_users = 0, n_items = 0;
  DevBuf<int64_t> lab_ptr;
  DevBuf<int32_t> lab_item;
  DevBuf<double> lab_val, totals, item_ps;
  DevBuf<int32_t> K_dev;
  DevBuf<double> per_user, metrics;
  DevBuf<int32_t> hits;
};

namespace {

constexpr int RK_THREADS = 128;
constexpr int RK_MAX_K = 128;     // largest supported ranking position
constexpr int RK_MAX_NK = 16;     // how many K values per call
constexpr int NTERMS = 7;         // dcg, ipsdcg, me, me_valid, recall, map, kept
constexpr int RK_SORT_CAP = 2048; // longest candidate list that is sorted whole in shared memory

// Total order on scores, as an order-preserving 64-bit key (larger key ranks first). The reference ranks with
// scores.argsort()[::-1] (utils/evaluate.py:93,197): NumPy sorts NaN behind every number, so the reversal puts NaN
// FIRST; -0.0 and +0.0 compare equal. Key 0 is never produced (it would be the bit pattern of a NaN), so it
// serves as the padding value that sorts last.
__device__ __forceinline__ uint64_t score_key(double s) {
  if (s != s) return ~0ull;
  if (s == 0.0) s = 0.0;                                  // -0.0 -> +0.0
  const uint64_t b = static_cast<uint64_t>(__double_as_longlong(s));
  return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}

// (key, pos) ranks before (key2, pos2): larger key first, later row first among equal keys
__device__ __forceinline__ bool before(uint64_t ka, int pa, uint64_t kb, int pb) {
  return ka > kb || (ka == kb && pa > pb);
}

// descending bitonic sort of n (a power of two) (key, pos) pairs in shared memory by the whole CTA
__device__ __forceinline__ void block_bitonic_sort(uint64_t *key, int *pos, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += RK_THREADS) {
        const int lo = ((t / stride) * stride << 1) + (t % stride), hi = lo + stride;
        const bool descending = (lo & size) == 0;
        const uint64_t ka = key[lo], kb = key[hi];
        const int pa = pos[lo], pb = pos[hi];
        if (before(kb, pb, ka, pa) == descending) {
          key[lo] = kb; key[hi] = ka;
          pos[lo] = pb; pos[hi] = pa;
        }
      }
    }
  }
  __syncthreads();
}

// CTA-wide sum of a per-thread count (every thread gets the total)
__device__ __forceinline__ int block_count(int c, int *wcnt) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) wcnt[threadIdx.x >> 5] = c;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < RK_THREADS / 32; ++w) t += wcnt[w];
  return t;
}

// One CTA per user. The user's top max(K) candidates in the canonical order are found by a bitonic sort of the
// whole list in shared memory (lists up to RK_SORT_CAP rows: every list of the reference's own workloads), or,
// for longer lists, by a two-level bisection (on the score key, then on the row position among exact ties) that
// selects exactly max(K) rows, which are then sorted. O(L log^2 L / threads) resp. O(96 L / threads) per user
// instead of max(K) passes over the list.
__global__ void __launch_bounds__(RK_THREADS)
rank_users_kernel(const int64_t *__restrict__ user_ptr, const int32_t *__restrict__ order,
                  const int32_t *__restrict__ item, const double *__restrict__ label,
                  const double *__restrict__ pscore, const double *__restrict__ scores,
                  const double *__restrict__ user_totals, int64_t n_users,
                  const int32_t *__restrict__ K, int n_k, int k_max, int64_t n_items,
                  double *__restrict__ per_user, int32_t *__restrict__ hits, int64_t *__restrict__ top_rows) {
  __shared__ uint64_t skey[RK_SORT_CAP];
  __shared__ int spos[RK_SORT_CAP];
  __shared__ double wsum[RK_THREADS / 32];
  __shared__ int wcnt[RK_THREADS / 32];
  __shared__ int n_sel;
  __shared__ double total_y_s;
  for (int64_t u = blockIdx.x; u < n_users; u += gridDim.x) {
    const int64_t beg = user_ptr[u];
    const int L = static_cast<int>(user_ptr[u + 1] - beg);
    // sum of labels decides whether the user counts at all
    double ysum = 0.0;
    for (int j = threadIdx.x; j < L; j += RK_THREADS) ysum += label[beg + j];
    ysum = warp_sum(ysum);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = ysum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < RK_THREADS / 32; ++w) t += wsum[w];
      total_y_s = t;
      n_sel = 0;
    }
    __syncthreads();
    // label total that decides the skip rule and Recall's denominator: over the rows given, or the
    // caller's total when the rows are only a prefix of the user's candidates (full-catalog evaluation)
    const double total_y = user_totals ? user_totals[u] : total_y_s;
    const int n_top = L < k_max ? L : k_max;
    int n_sort = 1;
    if (L <= RK_SORT_CAP) {
      while (n_sort < L) n_sort <<= 1;
      for (int j = threadIdx.x; j < n_sort; j += RK_THREADS) {
        skey[j] = j < L ? score_key(scores[order[beg + j]]) : 0ull;
        spos[j] = j < L ? j : -1;
      }
    } else {
      // n_top-th largest key: bisection on the key bits
      uint64_t thr = 0ull;
      for (int bit = 63; bit >= 0; --bit) {
        const uint64_t cand = thr | (1ull << bit);
        int c = 0;
        for (int j = threadIdx.x; j < L; j += RK_THREADS) c += score_key(scores[order[beg + j]]) >= cand;
        if (block_count(c, wcnt) >= n_top) thr = cand;
      }
      int above = 0;
      for (int j = threadIdx.x; j < L; j += RK_THREADS) above += score_key(scores[order[beg + j]]) > thr;
      above = block_count(above, wcnt);
      // among the rows that tie with it, the n_top - above latest ones: bisection on the position
      const int need = n_top - above;
      int pthr = 0;
      for (int bit = 30; bit >= 0; --bit) {
        const int cand = pthr | (1 << bit);
        int c = 0;
        for (int j = threadIdx.x; j < L; j += RK_THREADS)
          c += (j >= cand) && score_key(scores[order[beg + j]]) == thr;
        if (block_count(c, wcnt) >= need) pthr = cand;
      }
      while (n_sort < n_top) n_sort <<= 1;
      for (int j = threadIdx.x; j < n_sort; j += RK_THREADS) {
        skey[j] = 0ull;
        spos[j] = -1;
      }
      __syncthreads();
      for (int j = threadIdx.x; j < L; j += RK_THREADS) {
        const uint64_t kj = score_key(scores[order[beg + j]]);
        if (kj > thr || (kj == thr && j >= pthr)) {
          const int at = atomicAdd(&n_sel, 1);
          if (at < n_sort) {
            skey[at] = kj;
            spos[at] = j;
          }
        }
      }
    }
    block_bitonic_sort(skey, spos, n_sort);
    const int *top_pos = spos;             // positions inside the user's list, best first; -1 = padding
    if (top_rows) {
      for (int r = threadIdx.x; r < k_max; r += RK_THREADS)
        top_rows[u * k_max + r] = (r < n_top && top_pos[r] >= 0) ? static_cast<int64_t>(order[beg + top_pos[r]]) : -1;
    }
    if (threadIdx.x < n_k) {
      const int kk = threadIdx.x;
      const int k = K[kk];
      double *out = per_user + ((size_t)u * n_k + kk) * NTERMS;
      if (total_y == 0.0) {
        for (int c = 0; c < NTERMS; ++c) out[c] = 0.0;
      } else {
        const int n = L < k ? L : k;
        double dcg = 0.0, ips = 0.0, hit = 0.0, ap = 0.0;
        for (int j = 0; j < n; ++j) {
          const int tp = top_pos[j];
          if (tp < 0) break;
          const double y = label[beg + tp], ps = pscore[beg + tp];
          if (j == 0) {
            dcg += y;
            ips += y / ps;
          } else {
            const double d = log2(static_cast<double>(j + 1));
            dcg += y / d;
            ips += y / (ps * d);
          }
          hit += y;
          if (y >= 1.0) ap += hit / static_cast<double>(j + 1);
          if (hits) atomicAdd(hits + (size_t)kk * n_items + item[beg + tp], 1);
        }
        out[0] = dcg;
        out[1] = ips;
        out[2] = (L >= k && top_pos[k - 1] >= 0) ? pscore[beg + top_pos[k - 1]] : 0.0;
        out[3] = L >= k ? 1.0 : 0.0;
        out[4] = hit / total_y;
        out[5] = ap;
        out[6] = 1.0;
      }
    }
    __syncthreads();
  }
}

// fixed-order sum over users of every (k, term); one CTA per (k, term)
__global__ void __launch_bounds__(256)
rank_reduce_kernel(const double *__restrict__ per_user, int64_t n_users, int n_k, double *__restrict__ metrics) {
  __shared__ double wsum[8];
  const int kk = blockIdx.x / NTERMS, term = blockIdx.x % NTERMS;
  double s = 0.0;
  for (int64_t u = threadIdx.x; u < n_users; u += 256) s += per_user[((size_t)u * n_k + kk) * NTERMS + term];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += wsum[w];
    metrics[kk * RFM_RANK_NCOLS + term] = t;
  }
}

__global__ void __launch_bounds__(256)
rank_covered_kernel(const int32_t *__restrict__ hits, int64_t n_items, double *__restrict__ metrics) {
  __shared__ int wcnt[8];
  const int kk = blockIdx.x;
  int c = 0;
  for (int64_t i = threadIdx.x; i < n_items; i += 256) c += hits[(size_t)kk * n_items + i] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
  if ((threadIdx.x & 31) == 0) wcnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += wcnt[w];
    metrics[kk * RFM_RANK_NCOLS + RFM_RANK_COVERED] = static_cast<double>(t);
  }
}


// ---- full-catalog lists -> metrics, on the device (utils/evaluate.py:80-127 on the Cartesian frame) ------------
// Input: every user's top-k_list items, best first, as score.cu leaves them on the device (canonical order, -1
// padded). One warp per user: lanes look the items' held-out labels up in the label CSR (binary search in the
// user's row; 0 where the pair was not held out), then one lane per K walks the ranked prefix exactly like
// rank_users_kernel does, so both evaluators produce the same bits for the same ranked lists.
constexpr int CE_WARPS = 4;
__global__ void __launch_bounds__(CE_WARPS * 32)
catalog_metrics_kernel(const int32_t *__restrict__ lists, int k_list, int64_t user_begin, int64_t n_rows,
                       const int64_t *__restrict__ lab_ptr, const int32_t *__restrict__ lab_item,
                       const double *__restrict__ lab_val, const double *__restrict__ totals,
                       const double *__restrict__ item_ps, const int32_t *__restrict__ K, int n_k, int64_t n_items,
                       double *__restrict__ per_user, int32_t *__restrict__ hits) {
  __shared__ double sy[CE_WARPS][RK_MAX_K];
  __shared__ int si[CE_WARPS][RK_MAX_K];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = gw; row < n_rows; row += nw) {
    const int64_t u = user_begin + row;
    const int64_t lb = lab_ptr[u], le = lab_ptr[u + 1];
    int n_valid = 0;
    for (int j = lane; j < k_list; j += 32) {
      const int it = lists[row * k_list + j];
      double y = 0.0;
      if (it >= 0) {
        int64_t lo = lb, hi = le;
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (lab_item[mid] < it) lo = mid + 1; else hi = mid;
        }
        if (lo < le && lab_item[lo] == it) y = lab_val[lo];
      }
      si[wid][j] = it;
      sy[wid][j] = y;
      n_valid += it >= 0;
    }
    n_valid = __reduce_add_sync(FULL, n_valid);      // valid entries are a prefix of the list
    __syncwarp();
    const double total_y = totals[u];
    for (int kk = lane; kk < n_k; kk += 32) {
      const int k = K[kk];
      double *out = per_user + ((size_t)row * n_k + kk) * NTERMS;
      if (total_y == 0.0) {
        for (int c = 0; c < NTERMS; ++c) out[c] = 0.0;
      } else {
        const int n = n_valid < k ? n_valid : k;
        double dcg = 0.0, ips = 0.0, hit = 0.0, ap = 0.0;
        for (int j = 0; j < n; ++j) {
          const int it = si[wid][j];
          const double y = sy[wid][j], ps = item_ps[it];
          if (j == 0) {
            dcg += y;
            ips += y / ps;
          } else {
            const double d = log2(static_cast<double>(j + 1));
            dcg += y / d;
            ips += y / (ps * d);
          }
          hit += y;
          if (y >= 1.0) ap += hit / static_cast<double>(j + 1);
          if (hits) atomicAdd(hits + (size_t)kk * n_items + it, 1);
        }
        out[0] = dcg;
        out[1] = ips;
        out[2] = n_valid >= k ? item_ps[si[wid][k - 1]] : 0.0;
        out[3] = n_valid >= k ? 1.0 : 0.0;
        out[4] = hit / total_y;
        out[5] = ap;
        out[6] = 1.0;
      }
    }
    __syncwarp();
  }
}

// one ranking position's metric row -> a history slot on the device (epoch search: read back once per fit)
__global__ void rank_store_slot_kernel(const double *__restrict__ metrics, int n_k, double *__restrict__ slot) {
  for (int i = threadIdx.x; i < n_k * RFM_RANK_NCOLS; i += blockDim.x) slot[i] = metrics[i];
}

}  // namespace

extern "C" {

int rfm_ranker_create(rfm_ctx *ctx, int64_t n_rows, const int64_t *users, const int64_t *items,
                      const double *labels, const double *pscores, int64_t n_items, rfm_ranker **out) {
  RFM_REQUIRE(ctx && out, "rfm_ranker_create: NULL ctx/out");
  *out = nullptr;
  RFM_REQUIRE(n_rows >= 0 && n_rows < 0x7fffffffLL && n_items >= 1, "rfm_ranker_create: bad sizes");
  RFM_REQUIRE(n_rows == 0 || (users && items && labels && pscores), "rfm_ranker_create: NULL column");
  RFM_CUDA(cudaSetDevice(ctx->device));
  std::vector<int32_t> order((size_t)n_rows);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return users[a] < users[b]; });
  std::vector<int64_t> ptr;
  std::vector<int32_t> g_item((size_t)n_rows);
  std::vector<double> g_label((size_t)n_rows), g_ps((size_t)n_rows);
  for (int64_t j = 0; j < n_rows; ++j) {
    const int32_t r = order[(size_t)j];
    if (j == 0 || users[r] != users[order[(size_t)j - 1]]) ptr.push_back(j);
    RFM_REQUIRE(items[r] >= 0 && items[r] < n_items, "rfm_ranker_create: item id %lld outside [0, %lld)",
                (long long)items[r], (long long)n_items);
    g_item[(size_t)j] = (int32_t)items[r];
    g_label[(size_t)j] = labels[r];
    g_ps[(size_t)j] = pscores[r];
  }
  ptr.push_back(n_rows);
  rfm_ranker *r = new (std::nothrow) rfm_ranker();
  if (!r) return fail(RFM_ERR_NOMEM, "rfm_ranker_create: out of host memory");
  r->ctx = ctx;
  r->n_rows = n_rows;
  r->n_users = (int64_t)ptr.size() - 1;
  r->n_items = n_items;
  auto body = [&]() -> int {
    RFM_TRY(r->order.alloc(n_rows));
    RFM_TRY(r->user_ptr.alloc(ptr.size()));
    RFM_TRY(r->item.alloc(n_rows));
    RFM_TRY(r->label.alloc(n_rows));
    RFM_TRY(r->pscore.alloc(n_rows));
    RFM_TRY(r->scores.alloc(n_rows));
    RFM_TRY(r->K_dev.alloc(RK_MAX_NK));
    RFM_TRY(r->metrics.alloc((size_t)RK_MAX_NK * RFM_RANK_NCOLS));
    RFM_CUDA(cudaMemcpyAsync(r->order.p, order.data(), (size_t)n_rows * 4, cudaMemcpyHostToDevice, ctx->stream));
    RFM_CUDA(cudaMemcpyAsync(r->user_ptr.p, ptr.data(), ptr.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    RFM_CUDA(cudaMemcpyAsync(r->item.p, g_item.data(), (size_t)n_rows * 4, cudaMemcpyHostToDevice, ctx->stream));
    RFM_CUDA(cudaMemcpyAsync(r->label.p, g_label.data(), (size_t)n_rows * 8, cudaMemcpyHostToDevice, ctx->stream));
    RFM_CUDA(cudaMemcpyAsync(r->pscore.p, g_ps.data(), (size_t)n_rows * 8, cudaMemcpyHostToDevice, ctx->stream));
    RFM_CUDA(cudaStreamSynchronize(ctx->stream));
    return RFM_OK;
  };
  const int rc = body();
  if (rc != RFM_OK) {
    delete r;
    return rc;
  }
  *out = r;
  return RFM_OK;
}

int rfm_ranker_destroy(rfm_ranker *r) {
  if (r) {
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
    delete r;
  }
  return RFM_OK;
}

int rfm_ranker_num_users(rfm_ranker *r, int64_t *out) {
  RFM_REQUIRE(r && out, "rfm_ranker_num_users: NULL argument");
  *out = r->n_users;
  return RFM_OK;
}

int rfm_ranker_set_user_totals(rfm_ranker *r, const double *totals) {
  RFM_REQUIRE(r, "rfm_ranker_set_user_totals: ranker is NULL");
  rfm_ctx *ctx = r->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  if (!totals) {
    r->has_totals = false;
    return RFM_OK;
  }
  RFM_TRY(r->totals.ensure((size_t)(r->n_users ? r->n_users : 1)));
  RFM_CUDA(cudaMemcpyAsync(r->totals.p, totals, (size_t)r->n_users * 8, cudaMemcpyHostToDevice, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  r->has_totals = true;
  return RFM_OK;
}

}  // extern "C" (pause)
namespace {
constexpr int RK_HISTORY_K = 4;      // ranking positions a history slot holds

// ranks with the scores already in r->scores (device) and leaves the metric rows in r->metrics (device)
int ranker_run(rfm_ranker *r, const int32_t *K, int32_t n_k, bool want_hits, bool want_top, int *k_max_out) {
  RFM_REQUIRE(n_k >= 1 && n_k <= RK_MAX_NK, "rfm_ranker_evaluate: between 1 and %d ranking positions", RK_MAX_NK);
  int k_max = 0;
  for (int j = 0; j < n_k; ++j) {
    RFM_REQUIRE(K[j] >= 1 && K[j] <= RK_MAX_K, "rfm_ranker_evaluate: K[%d]=%d outside [1, %d]", j, K[j], RK_MAX_K);
    k_max = std::max(k_max, (int)K[j]);
  }
  rfm_ctx *ctx = r->ctx;
  RFM_TRY(r->per_user.ensure((size_t)(r->n_users ? r->n_users : 1) * n_k * NTERMS));
  RFM_TRY(r->hits.ensure((size_t)n_k * r->n_items));
  if (want_top) RFM_TRY(r->top_rows.ensure((size_t)(r->n_users ? r->n_users : 1) * k_max));
  RFM_CUDA(cudaMemcpyAsync(r->K_dev.p, K, (size_t)n_k * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (want_hits) RFM_CUDA(cudaMemsetAsync(r->hits.p, 0, (size_t)n_k * r->n_items * 4, ctx->stream));
  RFM_CUDA(cudaMemsetAsync(r->metrics.p, 0, (size_t)RK_MAX_NK * RFM_RANK_NCOLS * 8, ctx->stream));
  if (r->n_users > 0) {
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    const int grid = (int)(r->n_users < cap ? r->n_users : cap);
    RFM_LAUNCH(ctx, rank_users_kernel, grid, RK_THREADS, 0, r->user_ptr.p, r->order.p, r->item.p, r->label.p,
               r->pscore.p, r->scores.p, r->has_totals ? r->totals.p : (const double *)nullptr, r->n_users, r->K_dev.p,
               (int)n_k, k_max, r->n_items, r->per_user.p, want_hits ? r->hits.p : (int32_t *)nullptr,
               want_top ? r->top_rows.p : (int64_t *)nullptr);
    RFM_LAUNCH(ctx, rank_reduce_kernel, n_k * NTERMS, 256, 0, r->per_user.p, r->n_users, (int)n_k, r->metrics.p);
  }
  if (want_hits) RFM_LAUNCH(ctx, rank_covered_kernel, n_k, 256, 0, r->hits.p, r->n_items, r->metrics.p);
  if (k_max_out) *k_max_out = k_max;
  return RFM_OK;
}
}  // namespace
extern "C" {

int rfm_ranker_evaluate(rfm_ranker *r, const double *scores, const int32_t *K, int32_t n_k, double *out_metrics,
                        int32_t *out_item_hits, int64_t *out_top_rows) {
  RFM_REQUIRE(r && K && out_metrics, "rfm_ranker_evaluate: NULL argument");
  RFM_REQUIRE(r->n_rows == 0 || scores, "rfm_ranker_evaluate: scores is NULL");
  rfm_ctx *ctx = r->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  RFM_CUDA(cudaMemcpyAsync(r->scores.p, scores, (size_t)r->n_rows * 8, cudaMemcpyHostToDevice, ctx->stream));
  int k_max = 0;
  // the covered-items column is part of every metric row, so the hit counters always run here
  RFM_TRY(ranker_run(r, K, n_k, true, out_top_rows != nullptr, &k_max));
  RFM_CUDA(cudaMemcpyAsync(out_metrics, r->metrics.p, (size_t)n_k * RFM_RANK_NCOLS * 8, cudaMemcpyDeviceToHost,
                           ctx->stream));
  if (out_item_hits)
    RFM_CUDA(cudaMemcpyAsync(out_item_hits, r->hits.p, (size_t)n_k * r->n_items * 4, cudaMemcpyDeviceToHost,
                             ctx->stream));
  if (out_top_rows && r->n_users > 0)
    RFM_CUDA(cudaMemcpyAsync(out_top_rows, r->top_rows.p, (size_t)r->n_users * k_max * 8, cudaMemcpyDeviceToHost,
                             ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

int rfm_ranker_scores_ptr_dev(rfm_ranker *r, void **scores_dev) {
  RFM_REQUIRE(r && scores_dev, "rfm_ranker_scores_ptr_dev: NULL argument");
  *scores_dev = r->scores.p;
  return RFM_OK;
}

int rfm_ranker_evaluate_dev(rfm_ranker *r, const double *scores_dev, const int32_t *K, int32_t n_k, int64_t slot,
                            int64_t max_slots) {
  RFM_REQUIRE(r && K, "rfm_ranker_evaluate_dev: NULL argument");
  RFM_REQUIRE(n_k >= 1 && n_k <= RK_HISTORY_K, "rfm_ranker_evaluate_dev: between 1 and %d ranking positions", RK_HISTORY_K);
  RFM_REQUIRE(max_slots >= 1 && slot >= 0 && slot < max_slots, "rfm_ranker_evaluate_dev: slot %lld outside [0, %lld)",
              (long long)slot, (long long)max_slots);
  rfm_ctx *ctx = r->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  if (r->history_slots < max_slots) {
    RFM_REQUIRE(r->history_slots == 0 || slot == 0, "rfm_ranker_evaluate_dev: max_slots may only grow at slot 0");
    RFM_TRY(r->history.alloc((size_t)max_slots * RK_HISTORY_K * RFM_RANK_NCOLS));
    r->history_slots = max_slots;
  }
  if (scores_dev && scores_dev != r->scores.p)
    RFM_CUDA(cudaMemcpyAsync(r->scores.p, scores_dev, (size_t)r->n_rows * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  RFM_TRY(ranker_run(r, K, n_k, false, false, nullptr));
  RFM_LAUNCH(ctx, rank_store_slot_kernel, 1, 64, 0, r->metrics.p, (int)n_k,
             r->history.p + (size_t)slot * RK_HISTORY_K * RFM_RANK_NCOLS);
  return RFM_OK;
}

int rfm_ranker_read_slots(rfm_ranker *r, int64_t first_slot, int64_t n_slots, int32_t n_k, double *out_metrics) {
  RFM_REQUIRE(r && out_metrics, "rfm_ranker_read_slots: NULL argument");
  RFM_REQUIRE(n_k >= 1 && n_k <= RK_HISTORY_K, "rfm_ranker_read_slots: bad n_k");
  RFM_REQUIRE(first_slot >= 0 && n_slots >= 0 && first_slot + n_slots <= r->history_slots,
              "rfm_ranker_read_slots: slot range out of bounds");
  rfm_ctx *ctx = r->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  std::vector<double> tmp((size_t)n_slots * RK_HISTORY_K * RFM_RANK_NCOLS);
  if (n_slots > 0)
    RFM_CUDA(cudaMemcpyAsync(tmp.data(), r->history.p + (size_t)first_slot * RK_HISTORY_K * RFM_RANK_NCOLS,
                             tmp.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int64_t sl = 0; sl < n_slots; ++sl)
    for (int j = 0; j < n_k; ++j)
      memcpy(out_metrics + ((size_t)sl * n_k + j) * RFM_RANK_NCOLS,
             tmp.data() + ((size_t)sl * RK_HISTORY_K + j) * RFM_RANK_NCOLS, RFM_RANK_NCOLS * 8);
  return RFM_OK;
}

// ---- full-catalog evaluation on the device ------------------------------------------------------------------
int rfm_catalog_eval_create(rfm_ctx *ctx, int64_t n_users, int64_t n_items, const int64_t *label_indptr,
                            const int32_t *label_items, const double *label_values, const double *item_pscores,
                            rfm_catalog_eval **out) {
  RFM_REQUIRE(ctx && out, "rfm_catalog_eval_create: NULL ctx/out");
  *out = nullptr;
  RFM_REQUIRE(n_users >= 1 && n_items >= 1 && label_indptr && item_pscores, "rfm_catalog_eval_create: bad arguments");
  const int64_t nnz = label_indptr[n_users];
  RFM_REQUIRE(label_indptr[0] == 0 && nnz >= 0 && (nnz == 0 || (label_items && label_values)),
              "rfm_catalog_eval_create: bad label CSR");
  std::vector<double> totals((size_t)n_users, 0.0);
  for (int64_t u = 0; u < n_users; ++u) {
    RFM_REQUIRE(label_indptr[u + 1] >= label_indptr[u], "rfm_catalog_eval_create: indptr decreases at user %lld", (long long)u);
    double t = 0.0;
    for (int64_t z = label_indptr[u]; z < label_indptr[u + 1]; ++z) {
      RFM_REQUIRE(label_items[z] >= 0 && label_items[z] < n_items && (z == label_indptr[u] || label_items[z] > label_items[z - 1]),
                  "rfm_catalog_eval_create: items of user %lld are not strictly ascending ids in [0, %lld)", (long long)u,
                  (long long)n_items);
      t += label_values[z];
    }
    totals[(size_t)u] = t;
  }
  RFM_CUDA(cudaSetDevice(ctx->device));
  rfm_catalog_eval *e = new (std::nothrow) rfm_catalog_eval();
  if (!e) return fail(RFM_ERR_NOMEM, "rfm_catalog_eval_create: out of host memory");
  e->ctx = ctx;
  e->n_users = n_users;
  e->n_items = n_items;
  auto body = [&]() -> int {
    RFM_TRY(e->lab_ptr.alloc(n_users + 1));
    RFM_TRY(e->lab_item.alloc(nnz));
    RFM_TRY(e->lab_val.alloc(nnz));
    RFM_TRY(e->totals.alloc(n_users));
    RFM_TRY(e->item_ps.alloc(n_items));
    RFM_TRY(e->K_dev.alloc(RK_MAX_NK));
    RFM_TRY(e->metrics.alloc((size_t)RK_MAX_NK * RFM_RANK_NCOLS));
    RFM_CUDA(cudaMemcpyAsync(e->lab_ptr.p, label_indptr, (size_t)(n_users + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (nnz > 0) {
      RFM_CUDA(cudaMemcpyAsync(e->lab_item.p, label_items, (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
      RFM_CUDA(cudaMemcpyAsync(e->lab_val.p, label_values, (size_t)nnz * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    RFM_CUDA(cudaMemcpyAsync(e->totals.p, totals.data(), (size_t)n_users * 8, cudaMemcpyHostToDevice, ctx->stream));
    RFM_CUDA(cudaMemcpyAsync(e->item_ps.p, item_pscores, (size_t)n_items * 8, cudaMemcpyHostToDevice, ctx->stream));
    RFM_CUDA(cudaStreamSynchronize(ctx->stream));
    return RFM_OK;
  };
  const int rc = body();
  if (rc != RFM_OK) {
    delete e;
    return rc;
  }
  *out = e;
  return RFM_OK;
}

int rfm_catalog_eval_destroy(rfm_catalog_eval *e) {
  if (e) {
    cudaSetDevice(e->ctx->device);
    cudaStreamSynchronize(e->ctx->stream);
    delete e;
  }
  return RFM_OK;
}

int rfm_catalog_eval_run(rfm_catalog_eval *e, const int32_t *lists_dev, int32_t k_list, int64_t user_begin,
                         int64_t n_rows, const int32_t *K, int32_t n_k, double *out_metrics, int32_t *out_item_hits) {
  RFM_REQUIRE(e && K && out_metrics, "rfm_catalog_eval_run: NULL argument");
  RFM_REQUIRE(n_rows == 0 || lists_dev, "rfm_catalog_eval_run: lists_dev is NULL");
  RFM_REQUIRE(n_k >= 1 && n_k <= RK_MAX_NK, "rfm_catalog_eval_run: between 1 and %d ranking positions", RK_MAX_NK);
  RFM_REQUIRE(k_list >= 1 && k_list <= RK_MAX_K, "rfm_catalog_eval_run: k_list=%d outside [1, %d]", k_list, RK_MAX_K);
  RFM_REQUIRE(user_begin >= 0 && n_rows >= 0 && user_begin + n_rows <= e->n_users, "rfm_catalog_eval_run: user range out of bounds");
  for (int j = 0; j < n_k; ++j)
    RFM_REQUIRE(K[j] >= 1 && K[j] <= RK_MAX_K, "rfm_catalog_eval_run: K[%d]=%d outside [1, %d]", j, K[j], RK_MAX_K);
  rfm_ctx *ctx = e->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  RFM_TRY(e->per_user.ensure((size_t)(n_rows ? n_rows : 1) * n_k * NTERMS));
  RFM_TRY(e->hits.ensure((size_t)n_k * e->n_items));
  RFM_CUDA(cudaMemcpyAsync(e->K_dev.p, K, (size_t)n_k * 4, cudaMemcpyHostToDevice, ctx->stream));
  RFM_CUDA(cudaMemsetAsync(e->hits.p, 0, (size_t)n_k * e->n_items * 4, ctx->stream));
  RFM_CUDA(cudaMemsetAsync(e->metrics.p, 0, (size_t)RK_MAX_NK * RFM_RANK_NCOLS * 8, ctx->stream));
  if (n_rows > 0) {
    const int grid = (int)std::min<int64_t>((n_rows + CE_WARPS - 1) / CE_WARPS, (int64_t)ctx->sm_count * 8);
    RFM_LAUNCH(ctx, catalog_metrics_kernel, grid, CE_WARPS * 32, 0, lists_dev, (int)k_list, user_begin, n_rows,
               e->lab_ptr.p, e->lab_item.p, e->lab_val.p, e->totals.p, e->item_ps.p, e->K_dev.p, (int)n_k, e->n_items,
               e->per_user.p, e->hits.p);
    RFM_LAUNCH(ctx, rank_reduce_kernel, n_k * NTERMS, 256, 0, e->per_user.p, n_rows, (int)n_k, e->metrics.p);
  }
  RFM_LAUNCH(ctx, rank_covered_kernel, n_k, 256, 0, e->hits.p, e->n_items, e->metrics.p);
  RFM_CUDA(cudaMemcpyAsync(out_metrics, e->metrics.p, (size_t)n_k * RFM_RANK_NCOLS * 8, cudaMemcpyDeviceToHost,
                           ctx->stream));
  if (out_item_hits)
    RFM_CUDA(cudaMemcpyAsync(out_item_hits, e->hits.p, (size_t)n_k * e->n_items * 4, cudaMemcpyDeviceToHost,
                             ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

}  // extern "C"
