// common.cuh -- shared host/device helpers for librfm_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/rfm_b200.h"

namespace rfm {

// ---- error plumbing -------------------------------------------------------------------------
std::string &last_error();
int fail(int code, const char *fmt, ...);

#define RFM_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t err__ = (call);                                                            \
    if (err__ != cudaSuccess)                                                              \
      return ::rfm::fail(RFM_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                     \
                         cudaGetErrorString(err__), __FILE__, __LINE__);                   \
  } while (0)

#define RFM_TRY(expr)                      \
  do {                                     \
    int rc__ = (expr);                     \
    if (rc__ != RFM_OK) return rc__;       \
  } while (0)

#define RFM_REQUIRE(cond, ...)                                    \
  do {                                                            \
    if (!(cond)) return ::rfm::fail(RFM_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// ---- device memory --------------------------------------------------------------------------
// Device buffers come from CUDA's stream-ordered pool (cudaMallocAsync / cudaFreeAsync on the legacy default
// stream, which orders with every blocking stream). rfm_ctx_create raises the pool's release threshold so freed
// blocks stay cached: creating and destroying a trainer per fit() costs microseconds instead of the ~80 ms of
// cudaMalloc / cudaFree (each cudaFree is a device-wide synchronisation) measured on B200.
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p && cudaFreeAsync(p, nullptr) != cudaSuccess) cudaGetLastError();
    p = nullptr;
    n = 0;
  }
  int alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&p), count * sizeof(T), nullptr);
    if (e != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return fail(RFM_ERR_NOMEM, "cudaMallocAsync(%zu bytes) failed: %s", count * sizeof(T),
                  cudaGetErrorString(e));
    }
    n = count;
    return RFM_OK;
  }
  int ensure(size_t count) { return count <= n ? RFM_OK : alloc(count); }
};

template <typename T>
struct PinnedBuf {
  T *p = nullptr;
  size_t n = 0;
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf &) = delete;
  PinnedBuf &operator=(const PinnedBuf &) = delete;
  ~PinnedBuf() {
    if (p) cudaFreeHost(p);
  }
  int ensure(size_t count) {
    if (count <= n) return RFM_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&p), count * sizeof(T));
    if (e != cudaSuccess) {
      p = nullptr;
      n = 0;
      return fail(RFM_ERR_NOMEM, "cudaMallocHost(%zu bytes) failed: %s", count * sizeof(T),
                  cudaGetErrorString(e));
    }
    n = count;
    return RFM_OK;
  }
};

}  // namespace rfm

// ---- the context ----------------------------------------------------------------------------
struct rfm_prof_rec {
  const char *name;
  cudaEvent_t a, b;
};

struct rfm_ctx {
  int device = 0;
  bool profiling = false;
  bool sync_launches = false;   // RFM_SYNC_LAUNCHES=1: synchronise after every kernel and name the one that faulted
  std::vector<rfm_prof_rec> prof;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // staging ring for small host->device / device->host transfers
  rfm::PinnedBuf<unsigned char> stage;
  // Data-parallel exchange region and the CUDA-IPC mappings of the peers' regions (rfm_fm_dp_*). They belong
  // to the context, not to a trainer: cudaMalloc / cudaIpcOpenMemHandle / cudaIpcCloseMemHandle / cudaFree cost
  // tens of milliseconds per fit, more than the epochs of a short fit. A trainer borrows the region; a later
  // trainer of the same (or a smaller) gradient size re-zeroes and reuses it, and peers whose handle did not
  // change stay mapped. RFM_DP_CACHE=0 restores one region per trainer.
  struct DpRegion {
    static constexpr int MAX_WORLD = 8;
    unsigned char *base = nullptr;
    size_t bytes = 0;
    unsigned char own_handle[64] = {0};
    int world = 0, rank = -1;
    int borrowers = 0;                      // a second live trainer gets a region of its own
    unsigned char *peer[MAX_WORLD] = {nullptr};
    unsigned char peer_handle[MAX_WORLD][64] = {{0}};
    std::vector<unsigned char *> retired;   // outgrown regions: peers may still map them, freed with the context
  } dp;
  bool dp_cache = true;
  bool pdl = true;              // programmatic dependent launch along the FM step's kernel chain (RFM_PDL=0: off)
};

namespace rfm {

// per-kernel CUDA-event timing (bench.py's roofline numbers); off unless rfm_ctx_profile_begin
void prof_begin(rfm_ctx *ctx, const char *name);
void prof_end(rfm_ctx *ctx);

// every kernel launch goes through this so gpu_launches is an honest count
#define RFM_LAUNCH(ctx, kernel, grid, block, smem, ...)                                     \
  do {                                                                                      \
    if ((ctx)->profiling) ::rfm::prof_begin((ctx), #kernel);                                \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                        \
    (ctx)->launches++;                                                                      \
    if ((ctx)->profiling) ::rfm::prof_end((ctx));                                           \
    if ((ctx)->sync_launches) {                                                             \
      cudaError_t serr__ = cudaStreamSynchronize((ctx)->stream);                            \
      if (serr__ != cudaSuccess)                                                            \
        return ::rfm::fail(RFM_ERR_CUDA, "kernel %s faulted: %s (%s:%d)", #kernel,          \
                           cudaGetErrorString(serr__), __FILE__, __LINE__);                 \
    }                                                                                       \
    cudaError_t err__ = cudaGetLastError();                                                 \
    if (err__ != cudaSuccess)                                                               \
      return ::rfm::fail(RFM_ERR_CUDA, "launch of %s failed: %s (%s:%d)", #kernel,          \
                         cudaGetErrorString(err__), __FILE__, __LINE__);                    \
  } while (0)

// Dependent launch is used (and kernels may release their dependents early) only on the plain path: the per-kernel
// profile and RFM_SYNC_LAUNCHES put events / synchronisations between the kernels, i.e. operations that do not wait.
inline bool pdl_on(const rfm_ctx *ctx) { return ctx->pdl && !ctx->profiling && !ctx->sync_launches; }

// The same launch with programmatic stream serialization allowed (pdl_on(ctx)): the kernel may be scheduled while its
// predecessor in the stream drains. Only for kernels whose first statement is pdl_wait_and_release(): their bodies
// still run strictly after the predecessor has completed and flushed; what overlaps is the launch latency and the
// block scheduling of the dependent kernel with the tail of the primary.
#define RFM_LAUNCH_PDL(ctx, kernel, grid, block, smem, ...)                                 \
  do {                                                                                      \
    if ((ctx)->profiling) ::rfm::prof_begin((ctx), #kernel);                                \
    cudaLaunchConfig_t cfg__ = {};                                                          \
    cfg__.gridDim = dim3((unsigned)(grid));                                                 \
    cfg__.blockDim = dim3((unsigned)(block));                                               \
    cfg__.dynamicSmemBytes = (smem);                                                        \
    cfg__.stream = (ctx)->stream;                                                           \
    cudaLaunchAttribute attr__[1];                                                          \
    attr__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                      \
    attr__[0].val.programmaticStreamSerializationAllowed = ::rfm::pdl_on(ctx) ? 1 : 0;      \
    cfg__.attrs = attr__;                                                                   \
    cfg__.numAttrs = 1;                                                                     \
    cudaError_t lerr__ = cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__);                   \
    (ctx)->launches++;                                                                      \
    if ((ctx)->profiling) ::rfm::prof_end((ctx));                                           \
    if (lerr__ == cudaSuccess && (ctx)->sync_launches) {                                    \
      cudaError_t serr__ = cudaStreamSynchronize((ctx)->stream);                            \
      if (serr__ != cudaSuccess)                                                            \
        return ::rfm::fail(RFM_ERR_CUDA, "kernel %s faulted: %s (%s:%d)", #kernel,          \
                           cudaGetErrorString(serr__), __FILE__, __LINE__);                 \
    }                                                                                       \
    if (lerr__ != cudaSuccess) {                                                            \
      cudaGetLastError();                                                                   \
      return ::rfm::fail(RFM_ERR_CUDA, "launch of %s failed: %s (%s:%d)", #kernel,          \
                         cudaGetErrorString(lerr__), __FILE__, __LINE__);                   \
    }                                                                                       \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

template <typename T>
struct Vec2;
template <>
struct Vec2<float> {
  using type = float2;
};
template <>
struct Vec2<double> {
  using type = double2;
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
  // xor butterfly: every lane ends with the same, order-fixed total
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// First statement of a kernel that may be launched with RFM_LAUNCH_PDL: wait until the preceding kernel of the stream
// has completed and its writes are visible (a no-op under a plain launch), then let the NEXT kernel's blocks be
// scheduled as slots free up (they park at their own wait).
// The fence in between matters: a block launched programmatically becomes resident while the predecessor still
// runs, i.e. AFTER the L1 invalidation that a kernel boundary implies, so its SM's L1 can hold lines the predecessor's
// own blocks fetched and other SMs then overwrote; loads through the non-coherent path (__ldg, const __restrict__)
// would hit them (observed: a stale V row in the first loss pass after a fix-up on a 4-block problem). A gpu-scope
// fence makes ptxas emit CCTL.IVALL, which drops the SM's L1 lines, as the launch boundary would have.
// release: let the dependents go early. ONLY when the next operation of the stream is a kernel that starts with this
// call -- measured: after a kernel that released early, a following operation that does not wait (the next epoch's
// batch upload) did overtake it (a first-epoch loss computed on a half-overwritten batch, 15 of 25 runs of a 4-block
// problem). Every launch therefore says explicitly whether its successor waits (pdl_release in the argument
// structs, 0 by default).
__device__ __forceinline__ void pdl_wait_and_release(bool release = true) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __threadfence();
  if (release) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// host-callable device-wide utilities (scan.cu)
int exclusive_scan_u32(rfm_ctx *ctx, const uint32_t *in_dev, uint32_t *out_dev, int64_t n,
                       uint32_t *block_sums_dev /* >= ceil(n/4096)+1 */, uint32_t *total_dev);

}  // namespace rfm
