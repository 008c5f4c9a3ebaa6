// mf.cu -- logistic Matrix Factorization hot path (reference src/mf.py).
//
//   predict   src/mf.py:136-170   sigmoid(P_u . Q_i + b_u[u] + b_i[i] + b)
//   epoch     src/mf.py:97-108    strictly sequential per-sample SGD inside the batch:
//               err  = y/ps - predict(u, i)                         (before any update)
//               P_u -= lr(-err Q_i + reg P_u)                       mf.py:181-182
//               Q_i -= lr(-err P_u(new) + reg Q_i)                  mf.py:193-194
//               b_u[u] -= lr(-err + reg b_u[u]); b_i[i] likewise    mf.py:204-216
//   losses    src/mf.py:110-124   post-update batch loss, full val loss (src/base.py:37-61)
//
// Two samples of a batch commute unless they share a user or an item. The host assigns every
// sample the wavefront level 1 + max(level of the previous sample with the same user, same item);
// samples of one level touch disjoint rows, so a cooperative kernel runs level after level with a
// grid-wide barrier in between and reproduces the sequential result exactly (each sample's
// arithmetic is done by one warp with a fixed reduction order). This is latency-bound by design
// (SURVEY.md H5): the number of levels is the longest chain of repeats in the batch.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;
using namespace rfm;

struct rfm_pairs {
  rfm_ctx *ctx = nullptr;
  int dtype = RFM_F64;
  int64_t n_rows = 0;
  int64_t max_user = -1, max_item = -1;
  bool has_targets = false;
  DevBuf<int32_t> user, item;
  DevBuf<unsigned char> yp;
  std::vector<int32_t> h_user, h_item;   // host copy for the wavefront scheduler
};

struct rfm_mf {
  rfm_ctx *ctx = nullptr;
  int dtype = RFM_F64;
  int64_t n_users = 0, n_items = 0;
  int k = 0, kp = 0, nch = 0;
  double b = 0.0;
  DevBuf<unsigned char> P, Q, bu, bi;
  // per-epoch workspaces
  DevBuf<int64_t> idx;
  DevBuf<uint32_t> order, level_ptr;
  DevBuf<double> partials, result;
  std::vector<int32_t> last_u, last_i;      // wavefront level of the last sample touching a row
  std::vector<uint32_t> h_level, h_order, h_level_ptr;
  int coop_blocks_per_sm = 0;
};

namespace {

constexpr int MF_THREADS = 256;
constexpr int MF_WARPS = MF_THREADS / 32;

size_t dsize(int dtype) { return dtype == RFM_F64 ? 8 : 4; }

__device__ __forceinline__ double sigmoid_ref(double z) {
  z = fmin(fmax(z, -700.0), 700.0);
  return 1.0 / (1.0 + exp(-z));
}

template <typename T>
struct MfArgs {
  const int32_t *user, *item;
  const T *yp;
  const int64_t *idx;        // nullptr: rows [0, n)
  int64_t n;
  T *P, *Q, *bu, *bi;
  int kp;
  double b;
  double *out;               // predict
  double *partials;          // loss: one per warp
  // train
  const uint32_t *order, *level_ptr;
  int n_levels;
  T lr, reg;
};

template <typename T, int NCH>
__device__ __forceinline__ T pair_logit(const MfArgs<T> &a, int32_t u, int32_t i, int lane,
                                        typename Vec2<T>::type (&p)[NCH], typename Vec2<T>::type (&q)[NCH]) {
  using V2 = typename Vec2<T>::type;
  const V2 *prow = reinterpret_cast<const V2 *>(a.P + (size_t)u * a.kp) + lane;
  const V2 *qrow = reinterpret_cast<const V2 *>(a.Q + (size_t)i * a.kp) + lane;
  T dot = T(0);
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    p[ch] = prow[ch * 32];
    q[ch] = qrow[ch * 32];
    dot += p[ch].x * q[ch].x + p[ch].y * q[ch].y;
  }
  dot = warp_sum(dot);
  return ((dot + a.bu[u]) + a.bi[i]) + static_cast<T>(a.b);
}

// mode 0: predict -> out[]; mode 1: loss partials
template <typename T, int NCH, int MODE>
__global__ void __launch_bounds__(MF_THREADS)
mf_rows_kernel(const MfArgs<T> a) {
  using V2 = typename Vec2<T>::type;
  const int lane = lane_id();
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double partial = 0.0;
  for (int64_t s = gw; s < a.n; s += nw) {
    const int64_t t = a.idx ? a.idx[s] : s;
    V2 p[NCH], q[NCH];
    const double pr = sigmoid_ref(static_cast<double>(pair_logit<T, NCH>(a, a.user[t], a.item[t], lane, p, q)));
    if (MODE == 0) {
      if (lane == 0) a.out[s] = pr;
    } else {
      const double r = static_cast<double>(a.yp[t]);
      partial -= r * log(pr + 1e-8) + (1.0 - r) * log(1.0 - pr + 1e-8);
    }
  }
  if (MODE == 1 && lane == 0) a.partials[gw] = partial;
}

__global__ void __launch_bounds__(1024)
mf_reduce_kernel(const double *__restrict__ partials, int n, double scale, double *dst) {
  __shared__ double wsum[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) s += partials[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = warp_sum(wsum[threadIdx.x]);
    if (threadIdx.x == 0) *dst = v * scale;
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(MF_THREADS)
mf_wavefront_kernel(const MfArgs<T> a) {
  using V2 = typename Vec2<T>::type;
  cg::grid_group grid = cg::this_grid();
  const int lane = lane_id();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
  for (int level = 0; level < a.n_levels; ++level) {
    const uint32_t beg = a.level_ptr[level], end = a.level_ptr[level + 1];
    for (uint32_t o = beg + gw; o < end; o += nw) {
      const int64_t t = a.idx[a.order[o]];
      const int32_t u = a.user[t], i = a.item[t];
      V2 p[NCH], q[NCH];
      const T z = pair_logit<T, NCH>(a, u, i, lane, p, q);
      const T err = static_cast<T>(static_cast<double>(a.yp[t]) - sigmoid_ref(static_cast<double>(z)));
      V2 *prow = reinterpret_cast<V2 *>(a.P + (size_t)u * a.kp) + lane;
      V2 *qrow = reinterpret_cast<V2 *>(a.Q + (size_t)i * a.kp) + lane;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        V2 pn, qn;
        pn.x = p[ch].x - a.lr * (-err * q[ch].x + a.reg * p[ch].x);
        pn.y = p[ch].y - a.lr * (-err * q[ch].y + a.reg * p[ch].y);
        qn.x = q[ch].x - a.lr * (-err * pn.x + a.reg * q[ch].x);    // new P_u, mf.py:193
        qn.y = q[ch].y - a.lr * (-err * pn.y + a.reg * q[ch].y);
        prow[ch * 32] = pn;
        qrow[ch * 32] = qn;
      }
      if (lane == 0) {
        const T bu = a.bu[u], bi = a.bi[i];
        a.bu[u] = bu - a.lr * (-err + a.reg * bu);
        a.bi[i] = bi - a.lr * (-err + a.reg * bi);
      }
    }
    grid.sync();
  }
}

template <typename T>
__global__ void pad_kernel(const double *__restrict__ in, T *__restrict__ out, int64_t n, int k, int kp) {
  const int64_t total = n * kp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kp;
    const int f = static_cast<int>(i - r * kp);
    out[i] = f < k ? static_cast<T>(in[r * k + f]) : T(0);
  }
}

template <typename T>
__global__ void unpad_kernel(const T *__restrict__ in, double *__restrict__ out, int64_t n, int k, int kp) {
  const int64_t total = n * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k;
    const int f = static_cast<int>(i - r * k);
    out[i] = static_cast<double>(in[r * kp + f]);
  }
}

template <typename T>
__global__ void mf_targets_kernel(const int64_t *__restrict__ y, const double *__restrict__ ps, T *__restrict__ yp,
                                  int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    yp[i] = static_cast<T>(static_cast<double>(y[i]) / ps[i]);
}

template <typename T>
__global__ void mf_convert_targets_kernel(const double *__restrict__ in, T *__restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = static_cast<T>(in[i]);
}

#define MF_DISPATCH_NCH(nch, ...)                                  \
  switch (nch) {                                                   \
    case 1: { constexpr int NCH = 1; __VA_ARGS__; } break;         \
    case 2: { constexpr int NCH = 2; __VA_ARGS__; } break;         \
    case 3: { constexpr int NCH = 3; __VA_ARGS__; } break;         \
    case 4: { constexpr int NCH = 4; __VA_ARGS__; } break;         \
    case 5: { constexpr int NCH = 5; __VA_ARGS__; } break;         \
    case 6: { constexpr int NCH = 6; __VA_ARGS__; } break;         \
    case 7: { constexpr int NCH = 7; __VA_ARGS__; } break;         \
    case 8: { constexpr int NCH = 8; __VA_ARGS__; } break;         \
    default: return fail(RFM_ERR_INVALID, "n_factors too large (kp/64 = %d > 8)", nch); \
  }

int grid_for(rfm_ctx *ctx, int64_t blocks_wanted, int blocks_per_sm) {
  const int64_t cap = (int64_t)ctx->sm_count * blocks_per_sm;
  const int64_t g = blocks_wanted < cap ? blocks_wanted : cap;
  return g < 1 ? 1 : (int)g;
}

template <typename T>
MfArgs<T> mf_args(const rfm_mf *m, const rfm_pairs *rows) {
  MfArgs<T> a;
  memset(&a, 0, sizeof(a));
  a.user = rows->user.p;
  a.item = rows->item.p;
  a.yp = reinterpret_cast<const T *>(rows->yp.p);
  a.P = reinterpret_cast<T *>(m->P.p);
  a.Q = reinterpret_cast<T *>(m->Q.p);
  a.bu = reinterpret_cast<T *>(m->bu.p);
  a.bi = reinterpret_cast<T *>(m->bi.p);
  a.kp = m->kp;
  a.b = m->b;
  return a;
}

int check_pairs(const rfm_mf *m, const rfm_pairs *rows, const char *who) {
  RFM_REQUIRE(m && rows, "%s: NULL argument", who);
  RFM_REQUIRE(rows->ctx == m->ctx, "%s: rows and model live in different contexts", who);
  RFM_REQUIRE(rows->dtype == m->dtype, "%s: rows dtype %d != model dtype %d", who, rows->dtype, m->dtype);
  RFM_REQUIRE(rows->max_user < m->n_users && rows->max_item < m->n_items,
              "%s: user/item id out of range (max user %lld of %lld, max item %lld of %lld)", who,
              (long long)rows->max_user, (long long)m->n_users, (long long)rows->max_item, (long long)m->n_items);
  return RFM_OK;
}

// mean loss over rows (idx == nullptr) or over the batch idx[0..n) -> *dst_dev
template <typename T>
int mf_loss(rfm_mf *m, const rfm_pairs *rows, const int64_t *idx_dev, int64_t n, double *dst_dev) {
  rfm_ctx *ctx = m->ctx;
  MfArgs<T> a = mf_args<T>(m, rows);
  a.idx = idx_dev;
  a.n = n;
  const int grid = grid_for(ctx, ceil_div(n, MF_WARPS), 6);
  RFM_TRY(m->partials.ensure((size_t)grid * MF_WARPS));
  a.partials = m->partials.p;
  MF_DISPATCH_NCH(m->nch, {
    auto mf_rows_loss = mf_rows_kernel<T, NCH, 1>;
    RFM_LAUNCH(ctx, mf_rows_loss, grid, MF_THREADS, 0, a);
  });
  RFM_LAUNCH(ctx, mf_reduce_kernel, 1, 1024, 0, m->partials.p, grid * MF_WARPS, 1.0 / (double)n, dst_dev);
  return RFM_OK;
}

template <typename T>
int mf_epoch(rfm_mf *m, const rfm_pairs *train, const rfm_pairs *val, const int64_t *batch_rows, int64_t batch,
             double lr, double reg, double *train_loss, double *val_loss) {
  rfm_ctx *ctx = m->ctx;
  // ---- wavefront schedule on the host (O(batch)) ----
  m->h_level.resize((size_t)batch);
  uint32_t n_levels = 0;
  for (int64_t s = 0; s < batch; ++s) {
    const int64_t t = batch_rows[s];
    RFM_REQUIRE(t >= 0 && t < train->n_rows, "rfm_mf_train_epoch: row id %lld out of range", (long long)t);
    const int32_t u = train->h_user[(size_t)t], i = train->h_item[(size_t)t];
    const int32_t lvl = std::max(m->last_u[(size_t)u], m->last_i[(size_t)i]) + 1;   // levels start at 1
    m->last_u[(size_t)u] = lvl;
    m->last_i[(size_t)i] = lvl;
    m->h_level[(size_t)s] = (uint32_t)lvl;
    if ((uint32_t)lvl > n_levels) n_levels = (uint32_t)lvl;
  }
  for (int64_t s = 0; s < batch; ++s) {   // reset only what was touched
    const int64_t t = batch_rows[s];
    m->last_u[(size_t)train->h_user[(size_t)t]] = 0;
    m->last_i[(size_t)train->h_item[(size_t)t]] = 0;
  }
  m->h_level_ptr.assign((size_t)n_levels + 1, 0u);
  for (int64_t s = 0; s < batch; ++s) m->h_level_ptr[m->h_level[(size_t)s]]++;
  uint32_t max_width = 0;
  for (uint32_t l = 1; l <= n_levels; ++l) {
    max_width = std::max(max_width, m->h_level_ptr[l]);
    m->h_level_ptr[l] += m->h_level_ptr[l - 1];
  }
  m->h_order.resize((size_t)batch);
  {
    std::vector<uint32_t> cursor(m->h_level_ptr.begin(), m->h_level_ptr.end() - 1);
    for (int64_t s = 0; s < batch; ++s) m->h_order[cursor[m->h_level[(size_t)s] - 1]++] = (uint32_t)s;
  }
  RFM_TRY(m->idx.ensure((size_t)batch));
  RFM_TRY(m->order.ensure((size_t)batch));
  RFM_TRY(m->level_ptr.ensure((size_t)n_levels + 1));
  RFM_TRY(m->result.ensure(2));
  RFM_CUDA(cudaMemcpyAsync(m->idx.p, batch_rows, (size_t)batch * 8, cudaMemcpyHostToDevice, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(m->order.p, m->h_order.data(), (size_t)batch * 4, cudaMemcpyHostToDevice, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(m->level_ptr.p, m->h_level_ptr.data(), ((size_t)n_levels + 1) * 4,
                           cudaMemcpyHostToDevice, ctx->stream));

  MfArgs<T> a = mf_args<T>(m, train);
  a.idx = m->idx.p;
  a.n = batch;
  a.order = m->order.p;
  a.level_ptr = m->level_ptr.p;
  a.n_levels = (int)n_levels;
  a.lr = static_cast<T>(lr);
  a.reg = static_cast<T>(reg);
  MF_DISPATCH_NCH(m->nch, {
    auto mf_wavefront = mf_wavefront_kernel<T, NCH>;
    if (m->coop_blocks_per_sm == 0) {
      int per_sm = 0;
      RFM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mf_wavefront, MF_THREADS, 0));
      m->coop_blocks_per_sm = per_sm < 1 ? 1 : per_sm;
    }
    const int grid = grid_for(ctx, ceil_div(max_width, MF_WARPS), m->coop_blocks_per_sm);
    void *params[] = {&a};
    RFM_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(mf_wavefront), dim3(grid), dim3(MF_THREADS),
                                         params, 0, ctx->stream));
    ctx->launches++;
  });
  RFM_TRY(mf_loss<T>(m, train, m->idx.p, batch, m->result.p));
  if (val && val->n_rows > 0) RFM_TRY(mf_loss<T>(m, val, nullptr, val->n_rows, m->result.p + 1));
  double host[2] = {0.0, 0.0};
  RFM_CUDA(cudaMemcpyAsync(host, m->result.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (train_loss) *train_loss = host[0];
  if (val_loss) *val_loss = (val && val->n_rows > 0) ? host[1] : 0.0;
  return RFM_OK;
}

}  // namespace

extern "C" {

int rfm_pairs_create(rfm_ctx *ctx, int64_t n_rows, const int64_t *user_item, const int64_t *labels,
                     const double *pscores, int dtype, rfm_pairs **out) {
  RFM_REQUIRE(ctx && out, "rfm_pairs_create: NULL ctx/out");
  *out = nullptr;
  RFM_REQUIRE(n_rows >= 0 && (n_rows == 0 || user_item), "rfm_pairs_create: bad rows");
  RFM_REQUIRE(dtype == RFM_F32 || dtype == RFM_F64, "rfm_pairs_create: bad dtype %d", dtype);
  RFM_REQUIRE((labels == nullptr) == (pscores == nullptr), "rfm_pairs_create: labels and pscores go together");
  RFM_CUDA(cudaSetDevice(ctx->device));
  rfm_pairs *r = new (std::nothrow) rfm_pairs();
  if (!r) return fail(RFM_ERR_NOMEM, "rfm_pairs_create: out of host memory");
  r->ctx = ctx;
  r->dtype = dtype;
  r->n_rows = n_rows;
  r->has_targets = labels != nullptr;
  auto body = [&]() -> int {
    r->h_user.resize((size_t)n_rows);
    r->h_item.resize((size_t)n_rows);
    for (int64_t s = 0; s < n_rows; ++s) {
      const int64_t u = user_item[2 * s], i = user_item[2 * s + 1];
      RFM_REQUIRE(u >= 0 && i >= 0 && u < 0x7fffffffLL && i < 0x7fffffffLL,
                  "rfm_pairs_create: negative or oversized id at row %lld", (long long)s);
      r->h_user[(size_t)s] = (int32_t)u;
      r->h_item[(size_t)s] = (int32_t)i;
      if (u > r->max_user) r->max_user = u;
      if (i > r->max_item) r->max_item = i;
    }
    const size_t es = dsize(dtype);
    RFM_TRY(r->user.alloc(n_rows));
    RFM_TRY(r->item.alloc(n_rows));
    RFM_TRY(r->yp.alloc((size_t)(n_rows ? n_rows : 1) * es));
    RFM_CUDA(cudaMemcpyAsync(r->user.p, r->h_user.data(), (size_t)n_rows * 4, cudaMemcpyHostToDevice, ctx->stream));
    RFM_CUDA(cudaMemcpyAsync(r->item.p, r->h_item.data(), (size_t)n_rows * 4, cudaMemcpyHostToDevice, ctx->stream));
    DevBuf<int64_t> ytmp;
    DevBuf<double> pstmp;
    if (labels && n_rows > 0) {
      RFM_TRY(ytmp.alloc(n_rows));
      RFM_TRY(pstmp.alloc(n_rows));
      RFM_CUDA(cudaMemcpyAsync(ytmp.p, labels, (size_t)n_rows * 8, cudaMemcpyHostToDevice, ctx->stream));
      RFM_CUDA(cudaMemcpyAsync(pstmp.p, pscores, (size_t)n_rows * 8, cudaMemcpyHostToDevice, ctx->stream));
      const int g = grid_for(ctx, ceil_div(n_rows, 256), 8);
      if (dtype == RFM_F64) {
        RFM_LAUNCH(ctx, mf_targets_kernel<double>, g, 256, 0, ytmp.p, pstmp.p, reinterpret_cast<double *>(r->yp.p),
                   n_rows);
      } else {
        RFM_LAUNCH(ctx, mf_targets_kernel<float>, g, 256, 0, ytmp.p, pstmp.p, reinterpret_cast<float *>(r->yp.p),
                   n_rows);
      }
    } else {
      RFM_CUDA(cudaMemsetAsync(r->yp.p, 0, (size_t)(n_rows ? n_rows : 1) * es, ctx->stream));
    }
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

// Overwrite the per-row targets with y/pscore computed by the caller in float64 (fractional labels, src/mf.py:99).
int rfm_pairs_set_targets(rfm_pairs *rows, const double *targets) {
  RFM_REQUIRE(rows && targets, "rfm_pairs_set_targets: NULL argument");
  rfm_ctx *ctx = rows->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  if (rows->n_rows == 0) return RFM_OK;
  DevBuf<double> tmp;
  RFM_TRY(tmp.alloc(rows->n_rows));
  RFM_CUDA(cudaMemcpyAsync(tmp.p, targets, (size_t)rows->n_rows * 8, cudaMemcpyHostToDevice, ctx->stream));
  const int g = grid_for(ctx, ceil_div(rows->n_rows, 256), 8);
  if (rows->dtype == RFM_F64) {
    RFM_LAUNCH(ctx, mf_convert_targets_kernel<double>, g, 256, 0, tmp.p, reinterpret_cast<double *>(rows->yp.p),
               rows->n_rows);
  } else {
    RFM_LAUNCH(ctx, mf_convert_targets_kernel<float>, g, 256, 0, tmp.p, reinterpret_cast<float *>(rows->yp.p),
               rows->n_rows);
  }
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  rows->has_targets = true;
  return RFM_OK;
}

int rfm_pairs_destroy(rfm_pairs *rows) {
  if (rows) {
    cudaSetDevice(rows->ctx->device);
    cudaStreamSynchronize(rows->ctx->stream);
    delete rows;
  }
  return RFM_OK;
}

int rfm_mf_create(rfm_ctx *ctx, int64_t n_users, int64_t n_items, int32_t n_factors, int dtype, rfm_mf **out) {
  RFM_REQUIRE(ctx && out, "rfm_mf_create: NULL ctx/out");
  *out = nullptr;
  RFM_REQUIRE(n_users >= 1 && n_items >= 1 && n_factors >= 1, "rfm_mf_create: bad shape");
  RFM_REQUIRE(n_factors <= 512, "rfm_mf_create: n_factors %d > 512 is not supported", n_factors);
  RFM_REQUIRE(dtype == RFM_F32 || dtype == RFM_F64, "rfm_mf_create: bad dtype %d", dtype);
  RFM_CUDA(cudaSetDevice(ctx->device));
  rfm_mf *m = new (std::nothrow) rfm_mf();
  if (!m) return fail(RFM_ERR_NOMEM, "rfm_mf_create: out of host memory");
  m->ctx = ctx;
  m->dtype = dtype;
  m->n_users = n_users;
  m->n_items = n_items;
  m->k = n_factors;
  m->nch = (n_factors + 63) / 64;
  m->kp = m->nch * 64;
  const size_t es = dsize(dtype);
  int rc = m->P.alloc((size_t)n_users * m->kp * es);
  if (rc == RFM_OK) rc = m->Q.alloc((size_t)n_items * m->kp * es);
  if (rc == RFM_OK) rc = m->bu.alloc((size_t)n_users * es);
  if (rc == RFM_OK) rc = m->bi.alloc((size_t)n_items * es);
  if (rc != RFM_OK) {
    delete m;
    return rc;
  }
  cudaMemsetAsync(m->P.p, 0, (size_t)n_users * m->kp * es, ctx->stream);
  cudaMemsetAsync(m->Q.p, 0, (size_t)n_items * m->kp * es, ctx->stream);
  cudaMemsetAsync(m->bu.p, 0, (size_t)n_users * es, ctx->stream);
  cudaMemsetAsync(m->bi.p, 0, (size_t)n_items * es, ctx->stream);
  m->last_u.assign((size_t)n_users, 0);
  m->last_i.assign((size_t)n_items, 0);
  *out = m;
  return RFM_OK;
}

int rfm_mf_destroy(rfm_mf *m) {
  if (m) {
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    delete m;
  }
  return RFM_OK;
}

int rfm_mf_set_params(rfm_mf *m, const double *P, const double *Q, const double *b_u, const double *b_i, double b) {
  RFM_REQUIRE(m && P && Q && b_u && b_i, "rfm_mf_set_params: NULL argument");
  rfm_ctx *ctx = m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  m->b = b;
  const int64_t nP = m->n_users * m->k, nQ = m->n_items * m->k;
  DevBuf<double> tmp;
  RFM_TRY(tmp.alloc((size_t)(nP + nQ + m->n_users + m->n_items)));
  double *tP = tmp.p, *tQ = tP + nP, *tbu = tQ + nQ, *tbi = tbu + m->n_users;
  RFM_CUDA(cudaMemcpyAsync(tP, P, (size_t)nP * 8, cudaMemcpyHostToDevice, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(tQ, Q, (size_t)nQ * 8, cudaMemcpyHostToDevice, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(tbu, b_u, (size_t)m->n_users * 8, cudaMemcpyHostToDevice, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(tbi, b_i, (size_t)m->n_items * 8, cudaMemcpyHostToDevice, ctx->stream));
  const int g = grid_for(ctx, ceil_div(std::max(m->n_users, m->n_items) * m->kp, 256), 8);
  if (m->dtype == RFM_F64) {
    RFM_LAUNCH(ctx, pad_kernel<double>, g, 256, 0, tP, reinterpret_cast<double *>(m->P.p), m->n_users, m->k, m->kp);
    RFM_LAUNCH(ctx, pad_kernel<double>, g, 256, 0, tQ, reinterpret_cast<double *>(m->Q.p), m->n_items, m->k, m->kp);
    RFM_LAUNCH(ctx, pad_kernel<double>, g, 256, 0, tbu, reinterpret_cast<double *>(m->bu.p), m->n_users, 1, 1);
    RFM_LAUNCH(ctx, pad_kernel<double>, g, 256, 0, tbi, reinterpret_cast<double *>(m->bi.p), m->n_items, 1, 1);
  } else {
    RFM_LAUNCH(ctx, pad_kernel<float>, g, 256, 0, tP, reinterpret_cast<float *>(m->P.p), m->n_users, m->k, m->kp);
    RFM_LAUNCH(ctx, pad_kernel<float>, g, 256, 0, tQ, reinterpret_cast<float *>(m->Q.p), m->n_items, m->k, m->kp);
    RFM_LAUNCH(ctx, pad_kernel<float>, g, 256, 0, tbu, reinterpret_cast<float *>(m->bu.p), m->n_users, 1, 1);
    RFM_LAUNCH(ctx, pad_kernel<float>, g, 256, 0, tbi, reinterpret_cast<float *>(m->bi.p), m->n_items, 1, 1);
  }
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

int rfm_mf_get_params(rfm_mf *m, double *P, double *Q, double *b_u, double *b_i) {
  RFM_REQUIRE(m && P && Q && b_u && b_i, "rfm_mf_get_params: NULL argument");
  rfm_ctx *ctx = m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  const int64_t nP = m->n_users * m->k, nQ = m->n_items * m->k;
  DevBuf<double> tmp;
  RFM_TRY(tmp.alloc((size_t)(nP + nQ + m->n_users + m->n_items)));
  double *tP = tmp.p, *tQ = tP + nP, *tbu = tQ + nQ, *tbi = tbu + m->n_users;
  const int g = grid_for(ctx, ceil_div(std::max(m->n_users, m->n_items) * m->k, 256), 8);
  if (m->dtype == RFM_F64) {
    RFM_LAUNCH(ctx, unpad_kernel<double>, g, 256, 0, reinterpret_cast<const double *>(m->P.p), tP, m->n_users, m->k, m->kp);
    RFM_LAUNCH(ctx, unpad_kernel<double>, g, 256, 0, reinterpret_cast<const double *>(m->Q.p), tQ, m->n_items, m->k, m->kp);
    RFM_LAUNCH(ctx, unpad_kernel<double>, g, 256, 0, reinterpret_cast<const double *>(m->bu.p), tbu, m->n_users, 1, 1);
    RFM_LAUNCH(ctx, unpad_kernel<double>, g, 256, 0, reinterpret_cast<const double *>(m->bi.p), tbi, m->n_items, 1, 1);
  } else {
    RFM_LAUNCH(ctx, unpad_kernel<float>, g, 256, 0, reinterpret_cast<const float *>(m->P.p), tP, m->n_users, m->k, m->kp);
    RFM_LAUNCH(ctx, unpad_kernel<float>, g, 256, 0, reinterpret_cast<const float *>(m->Q.p), tQ, m->n_items, m->k, m->kp);
    RFM_LAUNCH(ctx, unpad_kernel<float>, g, 256, 0, reinterpret_cast<const float *>(m->bu.p), tbu, m->n_users, 1, 1);
    RFM_LAUNCH(ctx, unpad_kernel<float>, g, 256, 0, reinterpret_cast<const float *>(m->bi.p), tbi, m->n_items, 1, 1);
  }
  RFM_CUDA(cudaMemcpyAsync(P, tP, (size_t)nP * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(Q, tQ, (size_t)nQ * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(b_u, tbu, (size_t)m->n_users * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(b_i, tbi, (size_t)m->n_items * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

static int mf_predict_to(rfm_mf *m, const rfm_pairs *rows, double *out_dev) {
  rfm_ctx *ctx = m->ctx;
  const int grid = grid_for(ctx, ceil_div(rows->n_rows, MF_WARPS), 6);
  if (m->dtype == RFM_F64) {
    MfArgs<double> a = mf_args<double>(m, rows);
    a.n = rows->n_rows;
    a.out = out_dev;
    MF_DISPATCH_NCH(m->nch, {
      auto mf_rows_predict = mf_rows_kernel<double, NCH, 0>;
      RFM_LAUNCH(ctx, mf_rows_predict, grid, MF_THREADS, 0, a);
    });
  } else {
    MfArgs<float> a = mf_args<float>(m, rows);
    a.n = rows->n_rows;
    a.out = out_dev;
    MF_DISPATCH_NCH(m->nch, {
      auto mf_rows_predict = mf_rows_kernel<float, NCH, 0>;
      RFM_LAUNCH(ctx, mf_rows_predict, grid, MF_THREADS, 0, a);
    });
  }
  return RFM_OK;
}

int rfm_mf_predict(rfm_mf *m, const rfm_pairs *rows, double *out_scores) {
  RFM_TRY(check_pairs(m, rows, "rfm_mf_predict"));
  if (rows->n_rows == 0) return RFM_OK;
  RFM_REQUIRE(out_scores, "rfm_mf_predict: out_scores is NULL");
  rfm_ctx *ctx = m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double> out;
  RFM_TRY(out.alloc(rows->n_rows));
  RFM_TRY(mf_predict_to(m, rows, out.p));
  RFM_CUDA(cudaMemcpyAsync(out_scores, out.p, (size_t)rows->n_rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

// the same scores left in device memory, asynchronously (the evaluation chain of fit(evaluator=...))
int rfm_mf_predict_dev(rfm_mf *m, const rfm_pairs *rows, double *out_scores_dev) {
  RFM_TRY(check_pairs(m, rows, "rfm_mf_predict_dev"));
  if (rows->n_rows == 0) return RFM_OK;
  RFM_REQUIRE(out_scores_dev, "rfm_mf_predict_dev: out_scores_dev is NULL");
  RFM_CUDA(cudaSetDevice(m->ctx->device));
  return mf_predict_to(m, rows, out_scores_dev);
}

int rfm_mf_logloss(rfm_mf *m, const rfm_pairs *rows, double *out_loss) {
  RFM_TRY(check_pairs(m, rows, "rfm_mf_logloss"));
  RFM_REQUIRE(out_loss, "rfm_mf_logloss: out_loss is NULL");
  RFM_REQUIRE(rows->has_targets && rows->n_rows > 0, "rfm_mf_logloss: rows need labels/pscores");
  rfm_ctx *ctx = m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  RFM_TRY(m->result.ensure(2));
  if (m->dtype == RFM_F64) RFM_TRY(mf_loss<double>(m, rows, nullptr, rows->n_rows, m->result.p));
  else RFM_TRY(mf_loss<float>(m, rows, nullptr, rows->n_rows, m->result.p));
  RFM_CUDA(cudaMemcpyAsync(out_loss, m->result.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

int rfm_mf_train_epoch(rfm_mf *m, const rfm_pairs *train, const rfm_pairs *val, const int64_t *batch_rows,
                       int64_t batch, double lr, double reg, double *train_loss, double *val_loss) {
  RFM_TRY(check_pairs(m, train, "rfm_mf_train_epoch(train)"));
  if (val) RFM_TRY(check_pairs(m, val, "rfm_mf_train_epoch(val)"));
  RFM_REQUIRE(train->has_targets && (!val || val->has_targets), "rfm_mf_train_epoch: rows need labels/pscores");
  RFM_REQUIRE(batch_rows && batch >= 1, "rfm_mf_train_epoch: empty batch");
  RFM_REQUIRE(batch <= train->n_rows, "Cannot sample %lld out of arrays with dim %lld when replace is False",
              (long long)batch, (long long)train->n_rows);
  RFM_CUDA(cudaSetDevice(m->ctx->device));
  return m->dtype == RFM_F64
             ? mf_epoch<double>(m, train, val, batch_rows, batch, lr, reg, train_loss, val_loss)
             : mf_epoch<float>(m, train, val, batch_rows, batch, lr, reg, train_loss, val_loss);
}

}  // extern "C"
