// Micro-benchmark of the scoring kernel's MMA <-> epilogue pipeline without TMA: operands sit in shared memory,
// one thread issues 5 x tcgen05.mma (M128 N256 K16) per step into one of two 256-column accumulator stages, the
// epilogue warps drain the stage and hand it back. Prints cycles per step for several epilogue variants, to find
// where the ~1,300 cycles per step of the real kernel come from (the MMAs alone need 640).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o score_pipe_bench score_pipe_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
template <int SLEEP>
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (done) break;
    if (SLEEP > 0) __nanosleep(SLEEP);
  }
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldwait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

// EPI: 0 empty epilogue, 1 loads (serial wait each), 2 both loads in flight, 3 = 2 + max tree
// EWARPS: epilogue warps (8 or 16); ESLEEP: nanosleep in the epilogue's tfull wait; MSLEEP: in the issuer's waits
// N_MMA: MMAs per step (5 = k 64 + beta)
template <int EPI, int EWARPS, int ESLEEP, int MSLEEP, int N_MMA, int BETA = 0, int STORE = 0, int TMA = 0>
__global__ void __launch_bounds__(128 + 32 * EWARPS, 1) pipe_kernel(int steps, unsigned long long *out, float *sink, float *gmax, const unsigned char *src) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t tfull[2], tempty[2], full[3], empty[3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, EWARPS); }
    for (int s = 0; s < 3; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_ptr;
  const unsigned long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0 && TMA) {     // producer: 40 KB per tile of 4 steps into a 3-stage ring after the 48 KB of operands
      const int tiles = steps / 4;
      for (int it = 0; it < tiles; ++it) {
        const int s = it % 3;
        mbar_wait<32>(empty + s, ((it / 3) & 1) ^ 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full + s)), "r"(40960u) : "memory");
        const unsigned char *g = src + ((size_t)(blockIdx.x * 977 + it) % 4096) * 40960;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + 49152 + s * 40960)), "l"(g), "r"(40960u), "r"(smem_u32(full + s)) : "memory");
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256 >> 3) << 17) | ((128 >> 4) << 24);
      const uint64_t da = desc_sw128(smem_u32(smem)), db = desc_sw128(smem_u32(smem + 16384));
      // no-swizzle K-major descriptors (LBO 128 B, SBO 256 B), as the beta operand of the real kernel
      auto desc_il = [](uint32_t saddr) {
        uint64_t d = 0;
        d |= (uint64_t)((saddr >> 4) & 0x3FFF);
        d |= (uint64_t)(128 >> 4) << 16;
        d |= (uint64_t)(256 >> 4) << 32;
        d |= (uint64_t)1 << 46;
        return d;
      };
      const uint64_t ia = desc_il(smem_u32(smem)), ib = desc_il(smem_u32(smem + 16384));
      for (int st = 0; st < steps; ++st) {
        const int acc = st & 1;
        if (TMA && (st & 3) == 0 && st / 4 < steps / 4) mbar_wait<32>(full + (st / 4) % 3, ((st / 4) / 3) & 1);
        mbar_wait<MSLEEP>(tempty + acc, ((st >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int i = 0; i < N_MMA; ++i)
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                       ::"r"(tb + acc * 256), "l"((BETA && i == 0) ? ia : da + (uint64_t)((i & 3) * 2)),
                         "l"((BETA && i == 0) ? ib : db + (uint64_t)((i & 3) * 2)), "r"(idesc), "r"(i ? 1u : 0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(tfull + acc)) : "memory");
        if (TMA && (st & 3) == 3)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(empty + (st / 4) % 3)) : "memory");
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, part = (warp - 4) >> 2;
    constexpr int COLS = 256 / (EWARPS / 4);      // columns per warp per step
    float acc_max = -1e30f;
    for (int st = 0; st < steps; ++st) {
      const int acc = st & 1;
      mbar_wait<ESLEEP>(tfull + acc, (st >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = tb + ((uint32_t)(q * 32) << 16) + acc * 256 + part * COLS;
      if (EPI >= 1) {
        for (int c = 0; c < COLS; c += 64) {
          uint32_t v0[32], v1[32];
          if (EPI == 1) {
            ld32(base + c, v0); ldwait();
            ld32(base + c + 32, v1); ldwait();
          } else {
            ld32(base + c, v0); ld32(base + c + 32, v1); ldwait();
          }
          if (EPI == 3) {
#pragma unroll
            for (int i = 0; i < 30; i += 3) {
              acc_max = max3(acc_max, __uint_as_float(v0[i]), __uint_as_float(v0[i + 1]));
              acc_max = fmaxf(acc_max, __uint_as_float(v0[i + 2]));
              acc_max = max3(acc_max, __uint_as_float(v1[i]), __uint_as_float(v1[i + 1]));
              acc_max = fmaxf(acc_max, __uint_as_float(v1[i + 2]));
            }
          } else {
            acc_max = fmaxf(acc_max, __uint_as_float(v0[0] ^ v1[31]));
          }
        }
      }
      if (STORE == 1) gmax[((size_t)blockIdx.x * 128 + q * 32 + lane) * 4096 + (st & 1023) * 4 + part] = acc_max;   // one 4-byte store per row
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
      if (STORE == 2) gmax[((size_t)blockIdx.x * 128 + q * 32 + lane) * 4096 + (st & 1023) * 4 + part] = acc_max;   // after the hand-back
    }
    if (acc_max == 12345.f) sink[0] = acc_max;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
  }
}

template <int EPI, int EWARPS, int ESLEEP, int MSLEEP, int N_MMA, int BETA = 0, int STORE = 0, int TMA = 0>
void run(const char *name) {
  unsigned long long *out, h[148];
  float *sink;
  cudaMalloc(&out, 148 * 8);
  cudaMalloc(&sink, 4);
  auto k = pipe_kernel<EPI, EWARPS, ESLEEP, MSLEEP, N_MMA, BETA, STORE, TMA>;
  unsigned char *src;
  cudaMalloc(&src, (size_t)4096 * 40960);
  cudaMemset(src, 0, (size_t)4096 * 40960);
  float *gmax;
  cudaMalloc(&gmax, (size_t)148 * 128 * 4096 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  const int steps = 2000;
  k<<<148, 128 + 32 * EWARPS, 49152 + 3 * 40960 + 1024>>>(steps, out, sink, gmax, src);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-58s %s  %7.1f cycles/step\n", name, cudaGetErrorString(e), (double)h[0] / steps);
  cudaFree(out);
  cudaFree(sink);
  cudaFree(gmax);
  cudaFree(src);
}

int main() {
  run<3, 16, 0, 32, 5, 1, 2>("loads + max tree, stores AFTER the hand-back");
  run<3, 16, 0, 32, 5, 1, 0, 1>("loads + max tree, bulk-copy producer (40 KB per 4 steps)");
  run<3, 16, 0, 32, 5, 1, 2, 1>("loads + max tree, stores after hand-back, producer");
  run<3, 16, 0, 32, 5, 1, 0>("loads + max tree, first MMA on no-swizzle operands");
  run<3, 16, 0, 32, 5, 0, 1>("loads + max tree + group-max stores");
  run<3, 16, 0, 32, 5, 1, 1>("loads + max tree, no-swizzle first MMA, stores");
  run<0, 16, 0, 0, 5>("empty epilogue, 16 warps, spin");
  run<0, 16, 0, 32, 5>("empty epilogue, 16 warps, issuer sleeps 32 ns");
  run<1, 16, 0, 32, 5>("loads (wait each), 16 warps");
  run<2, 16, 0, 32, 5>("loads (both in flight), 16 warps");
  run<3, 16, 0, 32, 5>("loads + max tree, 16 warps");
  run<3, 8, 0, 32, 5>("loads + max tree, 8 warps");
  run<3, 16, 64, 32, 5>("loads + max tree, 16 warps, epilogue sleeps 64 ns");
  run<3, 16, 0, 0, 5>("loads + max tree, 16 warps, nobody sleeps");
  run<3, 16, 0, 32, 9>("loads + max tree, 16 warps, 9 MMAs per step (k = 128)");
  run<0, 16, 0, 32, 9>("empty epilogue, 9 MMAs per step");
  return 0;
}
