// Micro-benchmark: sustained tcgen05.ld (TMEM -> registers) bandwidth per SM for several shapes and warp counts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bw ldtm_bw.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>   // 0: 32x32b.x32, 1: 32x32b.x64, 2: 32x32b.x128, 3: 16x256b.x8 (two per 32 lanes), 4: 32x32b.x16
__device__ __forceinline__ uint32_t ld_once(uint32_t taddr) {
  uint32_t acc = 0;
  if (SHAPE == 0) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= v[i];
  } else if (SHAPE == 4) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= v[i];
  } else if (SHAPE == 3) {
    uint32_t v[8];    // 16 lanes x 256 bit per instruction, x2 repeats -> 16 lanes x 16 columns
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= v[i];
  }
  return acc;
}
template <int SHAPE> struct Bytes { static constexpr int v = SHAPE == 0 ? 4096 : SHAPE == 4 ? 2048 : 1024; };
template <int SHAPE> struct Cols { static constexpr int v = SHAPE == 0 ? 32 : SHAPE == 4 ? 16 : 16; };

template <int SHAPE, bool TWO_IN_FLIGHT>
__global__ void bw_kernel(int iters, unsigned long long *cycles, uint32_t *sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const unsigned long long t0 = clock64();
  const int part = warp >> 2, nparts = blockDim.x >> 7;
  for (int it = 0; it < iters; ++it) {
    for (int c = part * Cols<SHAPE>::v; c < 512; c += nparts * Cols<SHAPE>::v) {
      acc ^= ld_once<SHAPE>(base + c);
      if (SHAPE == 3) acc ^= ld_once<SHAPE>(base + c + (16u << 16));   // second half of the 32 lanes
    }
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512));
}

template <int SHAPE>
void run(const char *name, int threads) {
  unsigned long long *cyc;
  uint32_t *sink;
  cudaMalloc(&cyc, 148 * 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  bw_kernel<SHAPE, false><<<148, threads>>>(iters, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * 128 * 512 * 4;    // the whole TMEM per iteration
  printf("%-14s warps=%2d  %s  cycles=%llu  %.1f B/clk/SM\n", name, threads / 32, cudaGetErrorString(e), h[0], bytes / (double)h[0]);
  cudaFree(cyc);
  cudaFree(sink);
}


// ---- does tcgen05.ld overlap with tcgen05.mma? ------------------------------------------------------------------
// warp 16 issues MMAs (M=128, N=256, K=16, bf16, operands = whatever is in shared memory) into columns [0,256);
// warps 0..15 read columns [256,512). Timed: MMA alone, loads alone, both.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__global__ void __launch_bounds__(544, 1) overlap_kernel(int mode, int n_mma, int ld_iters, unsigned long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  __shared__ unsigned long long t_mma, t_ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    t_mma = 0; t_ld = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_ptr;
  const unsigned long long t0 = clock64();
  if (warp == 16) {
    if (lane == 0 && (mode & 1)) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256 >> 3) << 17) | ((128 >> 4) << 24);
      const uint64_t da = desc_sw128(smem_u32(smem)), db = desc_sw128(smem_u32(smem + 16384));
      for (int i = 0; i < n_mma; ++i) {
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tb), "l"(da + (uint64_t)((i & 3) * 2)), "l"(db + (uint64_t)((i & 3) * 2)), "r"(idesc), "r"(1u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      t_mma = clock64() - t0;
    }
  } else if (mode & 2) {
    const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    uint32_t acc = 0;
    const int part = warp >> 2;
    for (int it = 0; it < ld_iters; ++it)
      for (int c = part * 32; c < 256; c += 4 * 32) acc ^= ld_once<0>(base + c);
    if (acc == 0x12345678u) out[7] = acc;
    if (threadIdx.x == 0) t_ld = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t_mma; out[1] = t_ld; }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
  }
}
void overlap() {
  unsigned long long *out, h[8];
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  const int n_mma = 4000, ld_iters = 2000;
  for (int mode = 1; mode <= 3; ++mode) {
    cudaMemset(out, 0, 64);
    overlap_kernel<<<148, 544, 16384 + 32768 + 1024>>>(mode, n_mma, ld_iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
    printf("mode %d (%s): %s  mma %.1f cyc/MMA   ld %.1f B/clk/SM (16 warps, half of TMEM)\n", mode,
           mode == 1 ? "MMA only" : mode == 2 ? "loads only" : "both", cudaGetErrorString(e),
           h[0] ? (double)h[0] / n_mma : 0.0, h[1] ? (double)ld_iters * 128 * 256 * 4 / (double)h[1] : 0.0);
  }
}

int main() {
  overlap();
  for (int threads : {128, 256, 512}) {
    run<4>("32x32b.x16", threads);
    run<0>("32x32b.x32", threads);
    run<3>("16x256b.x2", threads);
  }
  return 0;
}
