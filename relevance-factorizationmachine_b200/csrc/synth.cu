// synth.cu -- device-side generator of semi-synthetic interaction logs.
//
// Mirrors the reference's simulation (utils/dataloader/kuairec/_click.py):
//   relevance  :148-171   gamma = clip(watch_ratio / relevance_clip, 0, 1)
//   exposure   :173-205   theta_i = max(sigmoid(3 z_i - 1) ** exposure_bias, eps), per item, from popularity counts
//   clicks     :207-235   O ~ Be(theta), R ~ Be(gamma), Y = O * R
//   pscore     kuairec/loader.py:167   theta ** pow_used
// The reference draws from NumPy's legacy global stream, which is sequential; here every random number is a pure
// function of (seed, stream, index): Philox4x32-10, so row g of the log is the same bits whichever GPU or launch
// produces it, and a 10^9-row log is generated shard by shard where it is consumed. Specification, bit for bit for
// the integer outputs: oracle/clicks_oracle.py.
#include "rows.cuh"

namespace {

struct Philox {
  uint32_t k0, k1;
};

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  c[0] = hi1 ^ c[1] ^ k0;
  c[1] = lo1;
  c[2] = hi0 ^ c[3] ^ k1;
  c[3] = lo0;
}

// Philox4x32-10: counter (index low, index high, stream, lane) under key (seed low, seed high)
__host__ __device__ __forceinline__ void philox4x32(uint64_t seed, uint64_t index, uint32_t stream, uint32_t lane,
                                                    uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), stream, lane};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0];
  out[1] = c[1];
  out[2] = c[2];
  out[3] = c[3];
}

// 53-bit uniform in (0, 1): ((a >> 5) * 2^26 + (b >> 6) + 0.5) / 2^53 -- NumPy's double construction, shifted off zero
__host__ __device__ __forceinline__ double u01(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

// Box-Muller, first branch only
__device__ __forceinline__ double normal_from(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  const double u1 = u01(a, b), u2 = u01(c, d);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// smallest index whose inclusive cumulative probability exceeds u (the last index absorbs rounding)
__device__ __forceinline__ int cdf_search(const double *__restrict__ cdf, int64_t n, double u) {
  int64_t lo = 0, hi = n - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(cdf + mid) > u) hi = mid; else lo = mid + 1;
  }
  return (int)lo;
}

enum Stream : uint32_t { ST_PAIR = 0, ST_CTX = 1, ST_NOISE = 2, ST_CLICK = 3, ST_USER_FACTOR = 16, ST_ITEM_FACTOR = 17 };

struct GenArgs {
  uint64_t seed;
  int64_t row0, n_rows, n_users, n_items;
  const double *user_cdf, *item_cdf, *theta, *pscore;
  double pow_used, hidden_scale, noise_scale, watch_shift, relevance_clip;
  int n_hidden, n_ctx;
  int32_t *user, *item;
  void *ctx, *targets;
  signed char *labels, *relevance;
};

template <typename T>
__global__ void __launch_bounds__(256)
synth_rows_kernel(const GenArgs a) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < a.n_rows; t += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t g = (uint64_t)(a.row0 + t);
    uint32_t r[4];
    philox4x32(a.seed, g, ST_PAIR, 0u, r);
    const int u = cdf_search(a.user_cdf, a.n_users, u01(r[0], r[1]));
    const int i = cdf_search(a.item_cdf, a.n_items, u01(r[2], r[3]));
    a.user[t] = u;
    a.item[t] = i;
    for (int j = 0; j < a.n_ctx; ++j) {                      // context columns ~ N(0, 1) (a standardised timestamp)
      philox4x32(a.seed, g, ST_CTX, (uint32_t)j, r);
      static_cast<T *>(a.ctx)[t * a.n_ctx + j] = static_cast<T>(normal_from(r[0], r[1], r[2], r[3]));
    }
    // hidden relevance: <p_u, q_i> over rank-n_hidden factors that are functions of the ids alone
    double dot = 0.0;
    for (int f = 0; f < a.n_hidden; ++f) {
      uint32_t ru[4], ri[4];
      philox4x32(a.seed, (uint64_t)u, ST_USER_FACTOR, (uint32_t)f, ru);
      philox4x32(a.seed, (uint64_t)i, ST_ITEM_FACTOR, (uint32_t)f, ri);
      dot += normal_from(ru[0], ru[1], ru[2], ru[3]) * normal_from(ri[0], ri[1], ri[2], ri[3]);
    }
    philox4x32(a.seed, g, ST_NOISE, 0u, r);
    const double eps = normal_from(r[0], r[1], r[2], r[3]);
    const double watch_ratio = exp(a.hidden_scale * dot + a.noise_scale * eps + a.watch_shift);
    const double gamma = fmin(fmax(watch_ratio / a.relevance_clip, 0.0), 1.0);        // _click.py:168-171
    const double theta = __ldg(a.theta + i);                                          // _click.py:193-205
    philox4x32(a.seed, g, ST_CLICK, 0u, r);
    const int O = u01(r[0], r[1]) < theta ? 1 : 0;                                    // exposure label ~ Be(theta)
    const int R = u01(r[2], r[3]) < gamma ? 1 : 0;                                    // relevance label ~ Be(gamma)
    const int Y = O * R;                                                              // _click.py:231
    const double ps = __ldg(a.pscore + i);                                            // theta ** pow_used, kuairec/loader.py:167 (table: pow() is not correctly rounded)
    static_cast<T *>(a.targets)[t] = static_cast<T>((double)Y / ps);
    if (a.labels) {
      a.labels[t] = (signed char)Y;
      a.relevance[t] = (signed char)R;
    }
  }
}

}  // namespace

int rfm_synth_fill_rows(rfm_ctx *ctx, const rfm_click_model *m, int64_t n_rows, int32_t *user_dev, int32_t *item_dev,
                        void *ctx_dev, int n_ctx, void *targets_dev, signed char *labels_dev, signed char *relevance_dev,
                        int dtype) {
  RFM_REQUIRE(m->n_users >= 1 && m->n_items >= 1 && m->n_users < 0x7fffffffLL && m->n_items < 0x7fffffffLL,
              "rfm_factored_generate: bad id ranges");
  RFM_REQUIRE(m->user_cdf && m->item_cdf && m->item_exposure && m->item_pscore, "rfm_factored_generate: NULL table in the click model");
  RFM_REQUIRE(m->n_hidden >= 0 && m->n_hidden <= 64 && m->relevance_clip > 0.0 && m->row0 >= 0,
              "rfm_factored_generate: bad click model parameters");
  DevBuf<double> ucdf, icdf, theta, pscore;
  RFM_TRY(ucdf.alloc(m->n_users));
  RFM_TRY(icdf.alloc(m->n_items));
  RFM_TRY(theta.alloc(m->n_items));
  RFM_TRY(pscore.alloc(m->n_items));
  RFM_CUDA(cudaMemcpyAsync(ucdf.p, m->user_cdf, (size_t)m->n_users * 8, cudaMemcpyHostToDevice, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(icdf.p, m->item_cdf, (size_t)m->n_items * 8, cudaMemcpyHostToDevice, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(theta.p, m->item_exposure, (size_t)m->n_items * 8, cudaMemcpyHostToDevice, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(pscore.p, m->item_pscore, (size_t)m->n_items * 8, cudaMemcpyHostToDevice, ctx->stream));
  GenArgs a;
  a.seed = m->seed;
  a.row0 = m->row0;
  a.n_rows = n_rows;
  a.n_users = m->n_users;
  a.n_items = m->n_items;
  a.user_cdf = ucdf.p;
  a.item_cdf = icdf.p;
  a.theta = theta.p;
  a.pscore = pscore.p;
  a.pow_used = m->pow_used;
  a.hidden_scale = m->hidden_scale;
  a.noise_scale = m->noise_scale;
  a.watch_shift = m->watch_shift;
  a.relevance_clip = m->relevance_clip;
  a.n_hidden = m->n_hidden;
  a.n_ctx = n_ctx;
  a.user = user_dev;
  a.item = item_dev;
  a.ctx = ctx_dev;
  a.targets = targets_dev;
  a.labels = labels_dev;
  a.relevance = relevance_dev;
  const int grid = (int)std::min<int64_t>((n_rows + 255) / 256, (int64_t)ctx->sm_count * 16);
  if (dtype == RFM_F64) {
    auto synth_rows = synth_rows_kernel<double>;
    RFM_LAUNCH(ctx, synth_rows, grid, 256, 0, a);
  } else {
    auto synth_rows = synth_rows_kernel<float>;
    RFM_LAUNCH(ctx, synth_rows, grid, 256, 0, a);
  }
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));     // the tables above are freed on return
  return RFM_OK;
}
