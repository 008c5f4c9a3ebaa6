// Micro-benchmark: 1-CTA (M128 N256 K16) vs 2-CTA (cta_group::2, M256 N256 K16) tcgen05.mma with and without a
// bulk-copy producer writing shared memory, to see how much of the MMA slowdown in the scoring kernel is
// shared-memory operand bandwidth (a CTA pair reads half of B per SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pair_mma_bench pair_mma_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// PAIR = 0: every CTA alone; PAIR = 1: clusters of two, the leader issues cta_group::2 MMAs for both
template <int PAIR, int COPY>
__global__ void __launch_bounds__(160, 1) mma_kernel(int n_mma, unsigned long long *out, const unsigned char *src) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar, cbar[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_rank() : 0;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_ptr;
  const unsigned long long t0 = clock64();
  if (warp == 1 && lane == 0 && rank == 0) {
    // A: this CTA's 128 rows x 64 (16 KB at smem + 0); B: 256 rows (1 CTA) or this CTA's 128 of them (pair) at + 16 KB
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256 >> 3) << 17) | (((PAIR ? 256 : 128) >> 4) << 24);
    const uint64_t da = desc_sw128(smem_u32(smem)), db = desc_sw128(smem_u32(smem + 16384));
    for (int i = 0; i < n_mma; ++i) {
      if (PAIR)
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tb + (i & 1) * 256), "l"(da + (uint64_t)((i & 3) * 2)), "l"(db + (uint64_t)((i & 3) * 2)), "r"(idesc), "r"(1u) : "memory");
      else
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tb + (i & 1) * 256), "l"(da + (uint64_t)((i & 3) * 2)), "l"(db + (uint64_t)((i & 3) * 2)), "r"(idesc), "r"(1u) : "memory");
    }
    if (PAIR)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  if (warp == 2 && lane == 0 && COPY) {
    // copy engine traffic into this CTA's shared memory while the MMAs run: 40 KB (1 CTA) or 24 KB (pair: half of B) per 640 cycles
    const uint32_t bytes = PAIR ? 24576u : 40960u;
    const int n_copy = n_mma / 5 / 4;
    for (int it = 0; it < n_copy; ++it) {
      const int s = it & 1;
      if (it >= 2) {
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&cbar[s])), "r"((uint32_t)(((it >> 1) - 1) & 1)) : "memory");
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[s])), "r"(bytes) : "memory");
      const unsigned char *g = src + ((size_t)(blockIdx.x * 977 + it) % 4096) * 40960;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(smem + 49152 + s * 40960)), "l"(g), "r"(bytes), "r"(smem_u32(&cbar[s])) : "memory");
    }
  }
  // everyone waits for the MMAs (the commit is multicast to both CTAs of a pair)
  if (threadIdx.x == 0) {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
  }
}

template <int PAIR, int COPY>
void run(const char *name, const unsigned char *src) {
  unsigned long long *out, h[148];
  cudaMalloc(&out, 148 * 8);
  auto k = mma_kernel<PAIR, COPY>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
  const int n_mma = 8000;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(160);
  cfg.dynamicSmemBytes = 49152 + 2 * 40960 + 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, n_mma, out, src);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-52s %s  %7.1f cycles per MMA (%s)\n", name, cudaGetErrorString(e), (double)h[0] / n_mma,
         PAIR ? "M256 N256 K16 on two SMs" : "M128 N256 K16");
  cudaFree(out);
}

int main() {
  unsigned char *src;
  cudaMalloc(&src, (size_t)4096 * 40960);
  cudaMemset(src, 0, (size_t)4096 * 40960);
  run<0, 0>("1 CTA, MMA only", src);
  run<0, 1>("1 CTA, MMA + copy engine (40 KB per 20 MMAs)", src);
  run<1, 0>("CTA pair, MMA only", src);
  run<1, 1>("CTA pair, MMA + copy engine (24 KB per 20 MMAs)", src);
  return 0;
}
