// runtime.cu -- context, error reporting, pinned memory, device-wide exclusive scan.
#include <cstdarg>
#include <cstdlib>

#include "common.cuh"

namespace rfm {

std::string &last_error() {
  static thread_local std::string err;
  return err;
}

int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

void prof_begin(rfm_ctx *ctx, const char *name) {
  rfm_prof_rec rec;
  rec.name = name;
  rec.a = rec.b = nullptr;
  if (cudaEventCreate(&rec.a) != cudaSuccess || cudaEventCreate(&rec.b) != cudaSuccess) return;
  cudaEventRecord(rec.a, ctx->stream);
  ctx->prof.push_back(rec);
}

void prof_end(rfm_ctx *ctx) {
  if (!ctx->prof.empty() && ctx->prof.back().b) cudaEventRecord(ctx->prof.back().b, ctx->stream);
}

// ---- exclusive scan of uint32 ---------------------------------------------------------------
// Three launches: per-tile totals, scan of the totals by one CTA, per-tile rescan + offset.
// Integer adds, so the result does not depend on scheduling.
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
  // v: this thread's value; returns the exclusive prefix over the CTA in thread order.
  __shared__ uint32_t warp_tot[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = (lane < (blockDim.x >> 5)) ? warp_tot[lane] : 0u;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(FULL, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc - w;  // exclusive prefix of warp totals
    if (lane == 31 && total) *total = winc;
  }
  __syncthreads();
  uint32_t res = warp_tot[wid] + inc - v;
  __syncthreads();  // warp_tot is reused by the next call
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_totals_kernel(const uint32_t *__restrict__ in, int64_t n, uint32_t *__restrict__ tile_tot) {
  __shared__ uint32_t tot;
  const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += in[base + i];
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) tile_tot[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_totals_kernel(uint32_t *__restrict__ tile_tot, int n_tiles, uint32_t *__restrict__ grand_total) {
  __shared__ uint32_t chunk_total;
  uint32_t carry = 0;
  for (int base = 0; base < n_tiles; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const uint32_t v = (i < n_tiles) ? tile_tot[i] : 0u;
    const uint32_t ex = block_exclusive_scan(v, &chunk_total);
    if (i < n_tiles) tile_tot[i] = carry + ex;
    carry += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && grand_total) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int64_t n,
                  const uint32_t *__restrict__ tile_off) {
  const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0u;
    s += v[i];
  }
  uint32_t run = block_exclusive_scan(s, nullptr) + tile_off[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

int exclusive_scan_u32(rfm_ctx *ctx, const uint32_t *in_dev, uint32_t *out_dev, int64_t n,
                       uint32_t *block_sums_dev, uint32_t *total_dev) {
  if (n <= 0) return RFM_OK;
  const int n_tiles = ceil_div(n, SCAN_TILE);
  RFM_LAUNCH(ctx, scan_tile_totals_kernel, n_tiles, SCAN_THREADS, 0, in_dev, n, block_sums_dev);
  RFM_LAUNCH(ctx, scan_totals_kernel, 1, SCAN_THREADS, 0, block_sums_dev, n_tiles, total_dev);
  RFM_LAUNCH(ctx, scan_apply_kernel, n_tiles, SCAN_THREADS, 0, in_dev, out_dev, n, block_sums_dev);
  return RFM_OK;
}

}  // namespace rfm

using namespace rfm;

extern "C" {

int rfm_abi_version(void) { return RFM_ABI_VERSION; }

const char *rfm_last_error(void) { return last_error().c_str(); }

int rfm_device_count(int *out) {
  RFM_REQUIRE(out != nullptr, "rfm_device_count: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *out = n;
  return RFM_OK;
}

int rfm_ctx_create(int device, void *cuda_stream, rfm_ctx **out) {
  RFM_REQUIRE(out != nullptr, "rfm_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(RFM_ERR_NO_DEVICE,
                "rfm_ctx_create: no CUDA device visible (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  RFM_REQUIRE(device >= 0 && device < n, "rfm_ctx_create: device %d out of range [0,%d)", device, n);
  RFM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  RFM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(RFM_ERR_NO_DEVICE,
                "rfm_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                device, prop.major, prop.minor);
  rfm_ctx *ctx = new (std::nothrow) rfm_ctx();
  if (!ctx) return fail(RFM_ERR_NOMEM, "rfm_ctx_create: out of host memory");
  ctx->device = device;
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->sm_count = prop.multiProcessorCount;
  {  // keep freed device memory in the default pool instead of returning it to the driver
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  const char *dc = getenv("RFM_DP_CACHE");
  ctx->dp_cache = !(dc && dc[0] == '0');
  const char *pd = getenv("RFM_PDL");
  ctx->pdl = !(pd && pd[0] == '0');     // on unless RFM_PDL=0
  const char *sl = getenv("RFM_SYNC_LAUNCHES");
  ctx->sync_launches = sl && sl[0] == '1';
  if (cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
    delete ctx;
    return fail(RFM_ERR_CUDA, "rfm_ctx_create: cudaEventCreate failed");
  }
  *out = ctx;
  return RFM_OK;
}

int rfm_ctx_destroy(rfm_ctx *ctx) {
  if (!ctx) return RFM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int q = 0; q < rfm_ctx::DpRegion::MAX_WORLD; ++q)
    if (q != ctx->dp.rank && ctx->dp.peer[q]) cudaIpcCloseMemHandle(ctx->dp.peer[q]);
  if (ctx->dp.base) cudaFree(ctx->dp.base);
  for (unsigned char *p : ctx->dp.retired) cudaFree(p);
  cudaGetLastError();
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  delete ctx;
  return RFM_OK;
}

int rfm_ctx_synchronize(rfm_ctx *ctx) {
  RFM_REQUIRE(ctx != nullptr, "rfm_ctx_synchronize: ctx is NULL");
  RFM_CUDA(cudaSetDevice(ctx->device));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

int rfm_ctx_launch_count(rfm_ctx *ctx, int64_t *out) {
  RFM_REQUIRE(ctx != nullptr && out != nullptr, "rfm_ctx_launch_count: NULL argument");
  *out = ctx->launches;
  return RFM_OK;
}

int rfm_ctx_timer_start(rfm_ctx *ctx) {
  RFM_REQUIRE(ctx != nullptr, "rfm_ctx_timer_start: ctx is NULL");
  RFM_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  return RFM_OK;
}

int rfm_ctx_timer_stop_ms(rfm_ctx *ctx, double *ms) {
  RFM_REQUIRE(ctx != nullptr && ms != nullptr, "rfm_ctx_timer_stop_ms: NULL argument");
  RFM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  RFM_CUDA(cudaEventSynchronize(ctx->ev1));
  float f = 0.f;
  RFM_CUDA(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
  *ms = f;
  return RFM_OK;
}

int rfm_ctx_profile_begin(rfm_ctx *ctx) {
  RFM_REQUIRE(ctx != nullptr, "rfm_ctx_profile_begin: ctx is NULL");
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->prof.clear();
  ctx->profiling = true;
  return RFM_OK;
}

int rfm_ctx_profile_end(rfm_ctx *ctx, char *out_text, size_t capacity) {
  RFM_REQUIRE(ctx != nullptr && out_text != nullptr && capacity > 0, "rfm_ctx_profile_end: bad argument");
  ctx->profiling = false;
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  struct Agg { const char *name; long count; double ms; };
  std::vector<Agg> agg;
  for (auto &rec : ctx->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, rec.a, rec.b) != cudaSuccess) ms = 0.f;
    cudaEventDestroy(rec.a);
    cudaEventDestroy(rec.b);
    bool found = false;
    for (auto &g : agg)
      if (strcmp(g.name, rec.name) == 0) { g.count++; g.ms += ms; found = true; break; }
    if (!found) agg.push_back({rec.name, 1, (double)ms});
  }
  ctx->prof.clear();
  std::string text;
  for (auto &g : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s\t%ld\t%.6f\n", g.name, g.count, g.ms);
    text += line;
  }
  RFM_REQUIRE(text.size() + 1 <= capacity, "rfm_ctx_profile_end: buffer too small (%zu needed)", text.size() + 1);
  memcpy(out_text, text.c_str(), text.size() + 1);
  return RFM_OK;
}

int rfm_host_register(void *p, size_t bytes) {
  RFM_REQUIRE(p != nullptr, "rfm_host_register: NULL pointer");
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(RFM_ERR_CUDA, "rfm_host_register: cudaHostRegister(%zu) failed: %s", bytes, cudaGetErrorString(e));
  }
  return RFM_OK;
}

int rfm_host_unregister(void *p) {
  if (p && cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
  return RFM_OK;
}

int rfm_host_alloc(size_t bytes, void **out) {
  RFM_REQUIRE(out != nullptr, "rfm_host_alloc: out is NULL");
  *out = nullptr;
  cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(RFM_ERR_NOMEM, "rfm_host_alloc: cudaMallocHost(%zu) failed: %s", bytes,
                cudaGetErrorString(e));
  }
  return RFM_OK;
}

int rfm_host_free(void *p) {
  if (p) cudaFreeHost(p);
  return RFM_OK;
}

}  // extern "C"
