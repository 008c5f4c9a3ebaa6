// Issue rates of the epilogue's candidate instructions on sm_100a: warp-instructions per clock per SM for
// FMNMX (2-input max), FMNMX3, FADD2 (add.f32x2), LOP3, FSETP-with-predicate-OR, with 16 resident warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rate_bench alu_rate_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512, 1) rate_kernel(int iters, float seed, unsigned long long *out, float *sink) {
  float x[16];
  uint32_t u[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = seed * (threadIdx.x + i); u[i] = threadIdx.x * 2654435761u + i; }
  unsigned anyp = 0;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int j = (i + 1) & 15, k = (i + 5) & 15;
      if (OP == 0) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(x[j]));
      if (OP == 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(x[j]), "f"(x[k]));
      if (OP == 2) {   // packed add on pairs (i even)
        if ((i & 1) == 0) {
          uint64_t a, b;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x[i]), "f"(x[i + 1]));
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(x[j]), "f"(x[k]));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(a));
        }
      }
      if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x80;" : "+r"(u[i]) : "r"(u[j]), "r"(u[k]));
      if (OP == 4) asm volatile("{\n .reg .pred p;\n setp.ge.f32 p, %1, %2;\n @p or.b32 %0, %0, 1;\n}" : "+r"(anyp) : "f"(x[i]), "f"(x[j]));
      if (OP == 5) asm volatile("add.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(x[j]));
      if (OP == 6) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[j]));
      if (OP == 7) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(x[j]));
      if (OP == 8) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[j]));
      if (OP == 9) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[j]), "r"(u[k]));
    }
  }
  const unsigned long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i] + (float)u[i];
  if (s == 12345.f || anyp == 77u) sink[0] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int per_iter) {
  unsigned long long *out, h[148];
  float *sink;
  cudaMalloc(&out, 148 * 8);
  cudaMalloc(&sink, 4);
  const int iters = 4000;
  rate_kernel<OP><<<148, 512>>>(iters, 1.0001f, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const double warp_instr = (double)iters * per_iter * 16;     // 16 warps per SM
  printf("%-28s %s  %.2f warp-instructions / clk / SM  (%.2f per scheduler)\n", name, cudaGetErrorString(e),
         warp_instr / (double)h[0], warp_instr / (double)h[0] / 4);
  cudaFree(out);
  cudaFree(sink);
}

int main() {
  run<5>("FADD", 16);
  run<0>("FMNMX (max.f32 2-input)", 16);
  run<1>("FMNMX3 (max.f32 3-input)", 16);
  run<2>("FADD2 (add.f32x2)", 8);
  run<3>("LOP3", 16);
  run<4>("FSETP + predicated OR", 32);
  run<6>("HMNMX2 (max.bf16x2)", 16);
  run<8>("HMNMX2 (max.f16x2)", 16);
  run<7>("F2FP (cvt.rn.bf16x2.f32)", 16);
  run<9>("HFMA2 (fma.rn.bf16x2)", 16);
  return 0;
}
