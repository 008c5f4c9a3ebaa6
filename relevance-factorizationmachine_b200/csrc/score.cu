// score.cu -- full-catalog scoring + exact per-user top-K (SURVEY.md Appendix A.4, section 8 f-rows).
//
// FM and MF scores decompose as  score(u, i) = bias + alpha[u] + beta[i] + <A_u, C_i>  when a row is
// [user-side features | item-side features] (for MF: A = P, C = Q, alpha = b_u, beta = b_i, bias = b;
// reference src/mf.py:165-170, src/fm.py:125-132). The reference never scores the full user x item
// grid; the parity statement for this path is "equal to the reference's predict on the
// Cartesian-product rows followed by its per-user argsort", checked in tests/test_score_gpu.py.
//
// Pipeline (all on the device):
//   1. A, C -> bf16 tiles, K padded to 64, rows padded to the tile; row norms in float64.
//   2. score_filter_kernel: a persistent warp-specialised tcgen05 GEMM. One CTA owns a block of 128
//      users and a range of 256-item tiles. TMA (cp.async.bulk.tensor, 128B swizzle) stages the
//      operands, one elected thread issues tcgen05.mma (M=128, N=256, K=16, bf16 -> fp32 in TMEM,
//      two accumulator stages), and four epilogue warps read the accumulator with tcgen05.ld: thread r
//      streams the scores of user r, adds beta, and keeps the Kc best (score, item) candidates above a
//      running threshold. Scores are never written to memory.
//   3. score_rescore_kernel: exact float64 scores of the candidates, exact top-K with the canonical tie
//      rule (score descending, larger item id first), and a proof that no item outside the candidate
//      list can belong to the top-K:  exact_K(u) > threshold(u) + eps(u), with
//      eps(u) = 2^-7 ||A_u|| max_i ||C_i|| bounding the bf16 rounding of both operands.
//   4. score_exact_kernel: users that fail the proof (or every user, in exact mode / for k > 128) are
//      scored in float64 against the whole catalog.
// The result is therefore always the exact float64 top-K; the tensor-core pass only prunes.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"

using namespace rfm;

namespace {

constexpr int BM = 128;          // users per CTA tile (= TMEM lanes)
constexpr int BN = 256;          // items per MMA tile (= TMEM columns per accumulator stage)
constexpr int BK = 64;           // bf16 elements per K block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
#ifndef RFM_EPI_WARPS
#define RFM_EPI_WARPS 8
#endif
constexpr int EPI_WARPS = RFM_EPI_WARPS;             // EPI_WARPS/4 per TMEM lane quarter; each owns a slice of a tile's columns
constexpr int SCORE_THREADS = 128 + 32 * EPI_WARPS;  // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4.. epilogue
constexpr int EPI_HALVES = EPI_WARPS / 4;
constexpr int MAX_KB = 2;        // k <= 128 on the tensor-core path
constexpr int MAX_K = 120;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier / TMA / tcgen05 wrappers (inline PTX) ---------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// same, for the producer / MMA threads: they run far ahead of the epilogue, and a bare spin would steal issue
// slots from the epilogue warps that share their schedulers (measured: 30 % of all issued instructions)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(200);
  }
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups of
// 1024 B. Descriptor fields: start address >> 4 (bits 0-13), leading byte offset (16-29, unused for
// swizzled K-major), stride byte offset = 1024 >> 4 (32-45), version 1 (46-47), layout SWIZZLE_128B = 2 (61-63).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// kind::f16 instruction descriptor: D = fp32 (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1), both
// K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
constexpr uint32_t UMMA_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);

struct FilterArgs {
  const float *beta;       // [n_items_pad], -inf beyond n_items
  int tile_begin;          // first tile of the catalog range being ranked (item-sharded calls)
  int n_item_tiles;        // tiles of BN items in that range
  int item_end;            // items >= item_end are not candidates (the last tile may run past the range)
  int tiles_per_split;
  int kc;                  // candidates kept per user per split
  int n_users_pad;
  float *cand_score;       // [n_splits][n_users_pad][kc]
  int32_t *cand_item;
  float *cand_tau;         // [n_splits][n_users_pad]  smallest kept score when the list is full, else -inf
  float *glist_score;      // global candidate lists when kc > MAX_KC_SMEM: [grid][kc][BM]
  int32_t *glist_item;
};

// Candidate lists are two-level: kc entries in groups of 8, plus the minimum of every group (gm). An
// insert finds the group holding the smallest kept score from the group minima, replaces that entry,
// rescans only that group's 8 entries (independent loads) and refreshes the threshold tau = min(gm).
// Everything is stored [index][row] so the 32 rows of a warp hit different banks / coalesce.
// It has exactly one call site, inside a rolled loop (see the epilogue), so it costs no code in the hot path.
__device__ __forceinline__ void insert_candidate(float s, int item, float *ls, int32_t *li, float *gm, int r, int n_groups,
                                              float &tau) {
  if (!(s > tau)) return;   // an earlier insert of the same 32-column slice may have raised the threshold
  float m1 = gm[r], m2 = INFINITY;
  int g1 = 0;
  for (int g = 1; g < n_groups; ++g) {
    const float x = gm[g * BM + r];
    if (x < m1) {
      m2 = m1;
      m1 = x;
      g1 = g;
    } else if (x < m2) {
      m2 = x;
    }
  }
  float e[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) e[j] = ls[(g1 * 8 + j) * BM + r];
  // smallest (to be replaced) and second smallest entry of the group, positions tracked in registers
  float e1 = e[0], e2 = INFINITY;
  int j1 = 0;
#pragma unroll
  for (int j = 1; j < 8; ++j) {
    if (e[j] < e1) {
      e2 = e1;
      e1 = e[j];
      j1 = j;
    } else if (e[j] < e2) {
      e2 = e[j];
    }
  }
  const float gmin = fminf(s, e2);
  ls[(g1 * 8 + j1) * BM + r] = s;
  li[(g1 * 8 + j1) * BM + r] = item;
  gm[g1 * BM + r] = gmin;
  tau = fminf(gmin, m2);
}

template <int KB>
struct FilterSmem {
  static constexpr int STAGES = (KB == 1 && EPI_WARPS <= 8) ? 3 : 2;
  static constexpr int A_BYTES = KB * BM * 128;
  static constexpr int B_STAGE_BYTES = KB * BN * 128;
  static constexpr int BAR_OFF = A_BYTES + STAGES * B_STAGE_BYTES;
  static constexpr int GM_OFF = BAR_OFF + 256;                 // group minima [EPI_HALVES][n_groups][BM] floats
  __host__ __device__ static size_t list_off(int n_groups) { return GM_OFF + (size_t)EPI_HALVES * n_groups * BM * 4; }
  // lists [EPI_HALVES][kc][BM] of (float score, int32 item) when they live in shared memory
  static size_t bytes(int n_groups, int kc_smem) {
    return 1024 + list_off(n_groups) + (size_t)EPI_HALVES * kc_smem * BM * 8;
  }
};

template <int KB>
__global__ void __launch_bounds__(SCORE_THREADS, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_c,
                    const FilterArgs a) {
  using L = FilterSmem<KB>;
  constexpr int STAGES = L::STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char *sA = smem;
  unsigned char *sB = smem + L::A_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::BAR_OFF);
  uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES, *tempty = bars + 2 * STAGES + 2;
  uint64_t *afull = bars + 2 * STAGES + 4;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int user_block = blockIdx.x, split = blockIdx.y;
  const int tile0 = a.tile_begin + split * a.tiles_per_split;
  int n_tiles = a.n_item_tiles - split * a.tiles_per_split;
  if (n_tiles > a.tiles_per_split) n_tiles = a.tiles_per_split;
  if (n_tiles < 0) n_tiles = 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull + s, 1);
      mbar_init(tempty + s, EPI_WARPS);     // one arrival per epilogue warp
    }
    mbar_init(afull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {   // ===== TMA producer =====
      mbar_expect_tx(afull, L::A_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sA + kb * BM * 128, &tmap_a, afull, kb * BK, user_block * BM);
      for (int it = 0; it < n_tiles; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait_relaxed(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, L::B_STAGE_BYTES);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(sB + s * L::B_STAGE_BYTES + kb * BN * 128, &tmap_c, full + s, kb * BK, (tile0 + it) * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ===== MMA issuer =====
      mbar_wait(afull, 0);
      for (int it = 0; it < n_tiles; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const int acc = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait_relaxed(tempty + acc, aph ^ 1);
        mbar_wait_relaxed(full + s, ph);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t a_base = smem_u32(sA + kb * BM * 128);
          const uint32_t b_base = smem_u32(sB + s * L::B_STAGE_BYTES + kb * BN * 128);
#pragma unroll
          for (int k4 = 0; k4 < BK / UMMA_K; ++k4) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle atom
            tc_mma_bf16(tmem_base + acc * BN, umma_desc_sw128(a_base + k4 * 32), umma_desc_sw128(b_base + k4 * 32),
                        UMMA_IDESC, (kb | k4) != 0 ? 1u : 0u);
          }
        }
        tc_commit(empty + s);      // the B stage is free once these MMAs have read it
        tc_commit(tfull + acc);    // the accumulator is complete
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread r keeps the kc best (score, item) pairs of user row r =====
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;       // which half of every tile's columns this warp filters
    const int r = q * 32 + lane;
    const int kc = a.kc;
    const int n_groups = kc >> 3;
    float *gm = reinterpret_cast<float *>(smem + L::GM_OFF) + (size_t)half * n_groups * BM;
    float *ls;
    int32_t *li;
    if (a.glist_score == nullptr) {
      unsigned char *lists = smem + L::list_off(n_groups);
      ls = reinterpret_cast<float *>(lists) + (size_t)half * kc * BM;
      li = reinterpret_cast<int32_t *>(lists + (size_t)EPI_HALVES * kc * BM * 4) + (size_t)half * kc * BM;
    } else {
      const size_t cta = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * EPI_HALVES + half;
      ls = a.glist_score + cta * kc * BM;
      li = a.glist_item + cta * kc * BM;
    }
    for (int j = 0; j < kc; ++j) {
      ls[j * BM + r] = -INFINITY;
      li[j * BM + r] = -1;
    }
    for (int g = 0; g < n_groups; ++g) gm[g * BM + r] = -INFINITY;
    float tau = -INFINITY;                  // smallest kept score
    for (int it = 0; it < n_tiles; ++it) {
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(tfull + acc, aph);
      tc_fence_after();
      const int item0 = (tile0 + it) * BN;
#pragma unroll 1
      for (int c = half * (BN / EPI_HALVES); c < (half + 1) * (BN / EPI_HALVES); c += 32) {
        uint32_t v[32];
        tc_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c, v);
        const float4 *b4 = reinterpret_cast<const float4 *>(a.beta + item0 + c);
        // four independent partial masks: a single "pass |= ..." chain is 32 dependent instructions long
        uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 b = __ldg(b4 + j4);
          p0 |= (__uint_as_float(v[j4 * 4 + 0]) + b.x > tau ? 1u : 0u) << (j4 * 4 + 0);
          p1 |= (__uint_as_float(v[j4 * 4 + 1]) + b.y > tau ? 1u : 0u) << (j4 * 4 + 1);
          p2 |= (__uint_as_float(v[j4 * 4 + 2]) + b.z > tau ? 1u : 0u) << (j4 * 4 + 2);
          p3 |= (__uint_as_float(v[j4 * 4 + 3]) + b.w > tau ? 1u : 0u) << (j4 * 4 + 3);
        }
        uint32_t pass = (p0 | p1) | (p2 | p3);
        // Rare after warm-up. Every lane walks ITS OWN passing columns (lanes pass at different columns, so
        // this takes max-over-lanes iterations, usually one, instead of one iteration per column). The
        // lane's score is pulled out of the register tile with a select chain (no run-time indexed array).
        while (__any_sync(FULL, pass != 0)) {
          const int j = pass ? __ffs(pass) - 1 : 0;
          // v[j] for a run-time j without a run-time indexed array: 5-level select tree on the bits of j
          uint32_t t16[16], t8[8], t4[4], t2[2];
#pragma unroll
          for (int i = 0; i < 16; ++i) t16[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
          for (int i = 0; i < 8; ++i) t8[i] = (j & 2) ? t16[2 * i + 1] : t16[2 * i];
#pragma unroll
          for (int i = 0; i < 4; ++i) t4[i] = (j & 4) ? t8[2 * i + 1] : t8[2 * i];
#pragma unroll
          for (int i = 0; i < 2; ++i) t2[i] = (j & 8) ? t4[2 * i + 1] : t4[2 * i];
          const float vj = __uint_as_float((j & 16) ? t2[1] : t2[0]);
          if (pass) {
            pass &= pass - 1;
            if (item0 + c + j < a.item_end)
              insert_candidate(vj + __ldg(a.beta + item0 + c + j), item0 + c + j, ls, li, gm, r, n_groups, tau);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
    }
    // hand the list to the exact re-scoring pass
    const size_t row = ((size_t)split * EPI_HALVES + half) * a.n_users_pad + (size_t)user_block * BM + r;
    for (int j = 0; j < kc; ++j) {
      a.cand_score[row * kc + j] = ls[j * BM + r];
      a.cand_item[row * kc + j] = li[j * BM + r];
    }
    a.cand_tau[row] = tau;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---- operand preparation ---------------------------------------------------------------------------------
// float64 [rows][k] -> bf16 [rows_pad][kpad] (zero padded) and float64 row norms
__global__ void to_bf16_kernel(const double *__restrict__ in, int64_t rows, int k, int64_t rows_pad, int kpad,
                               __nv_bfloat16 *__restrict__ out, double *__restrict__ norms) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t rrow = gw; rrow < rows_pad; rrow += nw) {
    double s = 0.0;
    for (int f = lane; f < kpad; f += 32) {
      const double v = (rrow < rows && f < k) ? in[rrow * k + f] : 0.0;
      out[rrow * kpad + f] = __double2bfloat16(v);
      s += v * v;
    }
    s = warp_sum(s);
    if (lane == 0 && rrow < rows) norms[rrow] = sqrt(s);
  }
}

__global__ void beta_f32_kernel(const double *__restrict__ beta, int64_t n_items, int64_t n_pad,
                                float *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = i < n_items ? static_cast<float>(beta ? beta[i] : 0.0) : -INFINITY;
}

__global__ void max_reduce_kernel(const double *__restrict__ v, int64_t n, double *__restrict__ out) {
  __shared__ double w[32];
  double m = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(v[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmax(m, w[i]);
    *out = m;
  }
}

// ---- exact float64 scoring ---------------------------------------------------------------------------------
struct ExactArgs {
  const double *A, *C, *alpha, *beta;
  double bias;
  int64_t n_users, n_items;
  int k;
  int item_begin, item_end;   // catalog range this call ranks (item-sharded runs)
};

// The exact float64 score, computed by a whole warp: lane l accumulates factors l, l+32, ... in order
// (coalesced reads of both rows), the 32 partial sums are combined by an xor butterfly, and every lane
// returns the same value. Both the re-scoring pass and the exact fallback use this one association, so
// a (user, item) pair gets the same bits whichever path scores it.
__device__ __forceinline__ double warp_exact_dot(const ExactArgs &e, int64_t u, int64_t i, int lane) {
  const double *au = e.A + u * e.k, *ci = e.C + i * e.k;
  double dot = 0.0;
  for (int f = lane; f < e.k; f += 32) dot += au[f] * ci[f];
  return warp_sum(dot);
}
__device__ __forceinline__ double warp_exact_score(const ExactArgs &e, int64_t u, int64_t i, int lane) {
  // outer association of the reference: (((dot + alpha_u) + beta_i) + bias), src/mf.py:165-170
  return ((warp_exact_dot(e, u, i, lane) + (e.alpha ? e.alpha[u] : 0.0)) + (e.beta ? e.beta[i] : 0.0)) + e.bias;
}

struct Best {
  double s;
  int item;   // -1 = none
};
__device__ __forceinline__ bool ranks_before(const Best &x, const Best &y) {
  if (x.item < 0) return false;
  if (y.item < 0) return true;
  return x.s > y.s || (x.s == y.s && x.item > y.item);   // canonical tie rule: later row (larger item id) first
}
__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best t;
    t.s = __shfl_xor_sync(FULL, b.s, o);
    t.item = __shfl_xor_sync(FULL, b.item, o);
    if (ranks_before(t, b)) b = t;
  }
  return b;
}

// one warp per user: exact scores of the candidates, exact top-K, and the pruning proof. The warp's
// candidates (item, exact score) are staged in shared memory so the K selection rounds never leave the SM.
constexpr int RESCORE_WARPS = 4;

__global__ void __launch_bounds__(RESCORE_WARPS * 32)
score_rescore_kernel(const ExactArgs e, const float *__restrict__ cand_score, const int32_t *__restrict__ cand_item,
                     const float *__restrict__ cand_tau,
                     int n_lists, int n_users_pad, int kc, int K, const double *__restrict__ a_norm,
                     const double *__restrict__ c_norm_max, const double *__restrict__ beta_abs_max,
                     int32_t *__restrict__ out_items, double *__restrict__ out_scores, int32_t *__restrict__ fail_list,
                     uint32_t *__restrict__ n_fail) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_cand = n_lists * kc;
  double *ex = reinterpret_cast<double *>(rs_smem) + (size_t)wid * n_cand;
  int32_t *it = reinterpret_cast<int32_t *>(rs_smem + (size_t)RESCORE_WARPS * n_cand * 8) + (size_t)wid * n_cand;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t u = gw; u < e.n_users; u += nw) {
    // 1. stage (item, approximate score as an order-preserving uint key) of every list entry
    uint32_t *key = reinterpret_cast<uint32_t *>(ex);     // the float64 slots are reused: keys first, scores later
    int n_valid = 0;
    for (int c = lane; c < n_cand; c += 32) {
      const int sp = c / kc, j = c - sp * kc;
      const size_t at = ((size_t)sp * n_users_pad + u) * kc + j;
      const int item = cand_item[at];
      const bool ok = item >= e.item_begin && item < e.item_end;
      const uint32_t bits = __float_as_uint(cand_score[at]);
      it[c] = ok ? item : -1;
      key[2 * c] = ok ? (bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u)) : 0u;
      n_valid += ok;
    }
    n_valid = __reduce_add_sync(FULL, n_valid);
    __syncwarp();
    // 2. only the kc best approximate scores of the merged lists can matter; their kc-th value is a valid
    //    pruning threshold (it is >= every list's own threshold). Find it by bisection on the key bits.
    uint32_t thr = 0u;
    float tau_max = -INFINITY;
    if (n_valid > kc) {
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand_thr = thr | (1u << bit);
        int cnt = 0;
        for (int c = lane; c < n_cand; c += 32) cnt += (it[c] >= 0 && key[2 * c] >= cand_thr);
        cnt = __reduce_add_sync(FULL, cnt);
        if (cnt >= kc) thr = cand_thr;       // at least kc keys are >= cand_thr: the kc-th largest is too
      }
      const uint32_t bits = (thr & 0x80000000u) ? (thr ^ 0x80000000u) : ~thr;
      tau_max = __uint_as_float(bits);
    } else {
      // every list entry is kept; items outside the lists are bounded by the lists' own thresholds
      for (int sp = lane; sp < n_lists; sp += 32) tau_max = fmaxf(tau_max, cand_tau[(size_t)sp * n_users_pad + u]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tau_max = fmaxf(tau_max, __shfl_xor_sync(FULL, tau_max, o));
    }
    for (int c = lane; c < n_cand; c += 32)
      if (it[c] >= 0 && key[2 * c] < thr) it[c] = -1;
    __syncwarp();
    // 3. exact float64 scores of the survivors, one candidate at a time, all lanes on its dot product
    for (int c = 0; c < n_cand; ++c) {
      const int item = it[c];
      if (item < 0) continue;
      const double sc = warp_exact_score(e, u, item, lane);
      __syncwarp();
      if (lane == 0) ex[c] = sc;
    }
    __syncwarp();
    Best last;
    last.s = 0.0;
    last.item = -2;
    for (int r = 0; r < K; ++r) {
      Best mine;
      mine.s = 0.0;
      mine.item = -1;
      for (int c = lane; c < n_cand; c += 32) {
        Best b;
        b.item = it[c];
        if (b.item < 0) continue;
        b.s = ex[c];
        const bool below = last.item == -2 || ranks_before(last, b);
        if (below && ranks_before(b, mine)) mine = b;
      }
      last = warp_best(mine);
      if (lane == 0) {
        out_items[u * K + r] = last.item;
        out_scores[u * K + r] = last.item >= 0 ? last.s : -INFINITY;
      }
      if (last.item < 0) {   // fewer than K candidates: pad the rest
        for (int rr = r + 1 + lane; rr < K; rr += 32) {
          out_items[u * K + rr] = -1;
          out_scores[u * K + rr] = -INFINITY;
        }
        break;
      }
    }
    // Every item outside the lists has approximate ranking score (<A_u, C_i> + beta_i, what the tensor-core
    // pass sees) <= tau_max, hence exact ranking score <= tau_max + eps. Compare with the K-th item's exact
    // ranking score (alpha_u and the bias are constant per user and do not affect the order).
    double kth_rank = -INFINITY;
    if (last.item >= 0) kth_rank = warp_exact_dot(e, u, last.item, lane) + (e.beta ? e.beta[last.item] : 0.0);
    const double eps = ldexp(a_norm[u] * *c_norm_max, -7) + ldexp(fabs((double)tau_max) + *beta_abs_max, -20);
    const bool complete = tau_max == -INFINITY;          // lists never filled: every item is a candidate
    const bool proven = complete || (last.item >= 0 && kth_rank > (double)tau_max + eps);
    if (!proven && lane == 0) fail_list[atomicAdd(n_fail, 1u)] = (int32_t)u;
    __syncwarp();
  }
}

// one CTA per user: exact top-K over the whole catalog range (fallback and exact mode)
__global__ void __launch_bounds__(256)
score_exact_kernel(const ExactArgs e, const int32_t *__restrict__ users, const uint32_t *__restrict__ n_users_dev,
                   int64_t n_users_host, int K, int32_t *__restrict__ out_items, double *__restrict__ out_scores) {
  __shared__ Best wbest[8];
  const int64_t n_u = n_users_dev ? (int64_t)*n_users_dev : n_users_host;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t ui = blockIdx.x; ui < n_u; ui += gridDim.x) {
    const int64_t u = users ? users[ui] : ui;
    Best last;
    last.s = 0.0;
    last.item = -2;
    for (int r = 0; r < K; ++r) {
      Best mine;
      mine.s = 0.0;
      mine.item = -1;
      for (int64_t i = e.item_begin + wid; i < e.item_end; i += 8) {   // a warp per item
        Best b;
        b.s = warp_exact_score(e, u, i, lane);
        b.item = (int)i;
        const bool below = last.item == -2 || ranks_before(last, b);
        if (below && ranks_before(b, mine)) mine = b;
      }
      if (lane == 0) wbest[wid] = mine;
      __syncthreads();
      Best best = wbest[0];
      for (int w = 1; w < 8; ++w)
        if (ranks_before(wbest[w], best)) best = wbest[w];
      __syncthreads();
      last = best;
      if (threadIdx.x == 0) {
        out_items[u * K + r] = last.item;
        out_scores[u * K + r] = last.item >= 0 ? last.s : -INFINITY;
      }
      if (last.item < 0) break;
    }
  }
}

// merge per-shard top-K lists: one warp per user picks the K best of n_lists * K (item, score) pairs with
// the canonical order. lists are [n_lists][n_users][K].
__global__ void __launch_bounds__(256)
topk_merge_kernel(const int32_t *__restrict__ items, const double *__restrict__ scores, int64_t n_users, int K,
                  int n_lists, int32_t *__restrict__ out_items, double *__restrict__ out_scores) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int n_cand = n_lists * K;
  for (int64_t u = gw; u < n_users; u += nw) {
    Best last;
    last.s = 0.0;
    last.item = -2;
    for (int r = 0; r < K; ++r) {
      Best mine;
      mine.s = 0.0;
      mine.item = -1;
      if (last.item != -1) {
        for (int c = lane; c < n_cand; c += 32) {
          const int l = c / K, j = c - l * K;
          const size_t at = ((size_t)l * n_users + u) * K + j;
          Best b;
          b.item = items[at];
          if (b.item < 0) continue;
          b.s = scores[at];
          const bool below = last.item == -2 || ranks_before(last, b);
          if (below && ranks_before(b, mine)) mine = b;
        }
      }
      last = warp_best(mine);
      if (lane == 0) {
        out_items[u * K + r] = last.item;
        out_scores[u * K + r] = last.item >= 0 ? last.s : -INFINITY;
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap(CUtensorMap *tm, void *base, int64_t rows_pad, int kpad, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    RFM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p) return fail(RFM_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows_pad};
  const cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RFM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return RFM_OK;
}

}  // namespace

struct rfm_topk {
  rfm_ctx *ctx = nullptr;
  int64_t n_users = 0, n_items = 0, n_users_pad = 0, n_items_pad = 0;
  int k = 0, kpad = 0, kb = 0;
  double bias = 0.0;
  bool has_alpha = false, has_beta = false, ready = false;
  DevBuf<double> A, C, alpha, beta, a_norm, c_norm, c_norm_max, beta_abs_max;
  DevBuf<__nv_bfloat16> A16, C16;
  DevBuf<float> beta32, cand_score, cand_tau, glist_score;
  DevBuf<int32_t> cand_item, glist_item, out_items, fail_list;
  DevBuf<double> out_scores, cand_exact;
  DevBuf<uint32_t> n_fail;
  CUtensorMap tmap_a, tmap_c;
};

extern "C" {

int rfm_topk_create(rfm_ctx *ctx, int64_t n_users, int64_t n_items, int32_t n_factors, rfm_topk **out) {
  RFM_REQUIRE(ctx && out, "rfm_topk_create: NULL ctx/out");
  *out = nullptr;
  RFM_REQUIRE(n_users >= 1 && n_items >= 1 && n_factors >= 1, "rfm_topk_create: bad shape");
  RFM_REQUIRE(n_items < 0x7fffff00LL && n_users < 0x7fffff00LL, "rfm_topk_create: too many users/items");
  RFM_CUDA(cudaSetDevice(ctx->device));
  rfm_topk *t = new (std::nothrow) rfm_topk();
  if (!t) return fail(RFM_ERR_NOMEM, "rfm_topk_create: out of host memory");
  t->ctx = ctx;
  t->n_users = n_users;
  t->n_items = n_items;
  t->k = n_factors;
  t->kb = (n_factors + BK - 1) / BK;
  t->kpad = t->kb * BK;
  t->n_users_pad = (n_users + BM - 1) / BM * BM;
  t->n_items_pad = (n_items + BN - 1) / BN * BN;
  auto body = [&]() -> int {
    RFM_TRY(t->A.alloc((size_t)n_users * n_factors));
    RFM_TRY(t->C.alloc((size_t)n_items * n_factors));
    RFM_TRY(t->alpha.alloc(n_users));
    RFM_TRY(t->beta.alloc(n_items));
    RFM_TRY(t->a_norm.alloc(n_users));
    RFM_TRY(t->c_norm.alloc(n_items));
    RFM_TRY(t->c_norm_max.alloc(1));
    RFM_TRY(t->beta_abs_max.alloc(1));
    RFM_TRY(t->n_fail.alloc(1));
    RFM_TRY(t->fail_list.alloc(n_users));
    if (t->kb <= MAX_KB) {
      RFM_TRY(t->A16.alloc((size_t)t->n_users_pad * t->kpad));
      RFM_TRY(t->C16.alloc((size_t)t->n_items_pad * t->kpad));
      RFM_TRY(t->beta32.alloc(t->n_items_pad));
    }
    return RFM_OK;
  };
  const int rc = body();
  if (rc != RFM_OK) {
    delete t;
    return rc;
  }
  *out = t;
  return RFM_OK;
}

int rfm_topk_destroy(rfm_topk *t) {
  if (t) {
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    delete t;
  }
  return RFM_OK;
}

int rfm_topk_set_factors(rfm_topk *t, const double *A, const double *C, const double *alpha, const double *beta,
                         double bias) {
  RFM_REQUIRE(t && A && C, "rfm_topk_set_factors: NULL argument");
  rfm_ctx *ctx = t->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  t->bias = bias;
  t->has_alpha = alpha != nullptr;
  t->has_beta = beta != nullptr;
  RFM_CUDA(cudaMemcpyAsync(t->A.p, A, (size_t)t->n_users * t->k * 8, cudaMemcpyHostToDevice, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(t->C.p, C, (size_t)t->n_items * t->k * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (alpha) RFM_CUDA(cudaMemcpyAsync(t->alpha.p, alpha, (size_t)t->n_users * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (beta) RFM_CUDA(cudaMemcpyAsync(t->beta.p, beta, (size_t)t->n_items * 8, cudaMemcpyHostToDevice, ctx->stream));
  else RFM_CUDA(cudaMemsetAsync(t->beta.p, 0, (size_t)t->n_items * 8, ctx->stream));
  if (t->kb <= MAX_KB) {
    const int g = ctx->sm_count * 8;
    RFM_LAUNCH(ctx, to_bf16_kernel, g, 256, 0, t->A.p, t->n_users, t->k, t->n_users_pad, t->kpad, t->A16.p,
               t->a_norm.p);
    RFM_LAUNCH(ctx, to_bf16_kernel, g, 256, 0, t->C.p, t->n_items, t->k, t->n_items_pad, t->kpad, t->C16.p,
               t->c_norm.p);
    RFM_LAUNCH(ctx, beta_f32_kernel, g, 256, 0, beta ? t->beta.p : (const double *)nullptr, t->n_items,
               t->n_items_pad, t->beta32.p);
    RFM_LAUNCH(ctx, max_reduce_kernel, 1, 1024, 0, t->c_norm.p, t->n_items, t->c_norm_max.p);
    RFM_LAUNCH(ctx, max_reduce_kernel, 1, 1024, 0, t->beta.p, t->n_items, t->beta_abs_max.p);
    RFM_TRY(make_tmap(&t->tmap_a, t->A16.p, t->n_users_pad, t->kpad, BM));
    RFM_TRY(make_tmap(&t->tmap_c, t->C16.p, t->n_items_pad, t->kpad, BN));
  }
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  t->ready = true;
  return RFM_OK;
}

int rfm_topk_run(rfm_topk *t, int32_t K, int32_t mode, int64_t item_begin, int64_t item_end, int32_t *out_items,
                 double *out_scores, int64_t *stats) {
  RFM_REQUIRE(t, "rfm_topk_run: NULL argument");
  RFM_REQUIRE((out_items == nullptr) == (out_scores == nullptr), "rfm_topk_run: out_items and out_scores go together");
  RFM_REQUIRE(t->ready, "rfm_topk_run: call rfm_topk_set_factors first");
  RFM_REQUIRE(K >= 1 && K <= MAX_K, "rfm_topk_run: K=%d outside [1, %d]", K, MAX_K);
  RFM_REQUIRE(mode == 0 || mode == 1, "rfm_topk_run: mode must be 0 (tensor-core prune + exact) or 1 (exact only)");
  if (item_end <= 0) item_end = t->n_items;
  RFM_REQUIRE(item_begin >= 0 && item_begin < item_end && item_end <= t->n_items, "rfm_topk_run: bad item range");
  rfm_ctx *ctx = t->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  RFM_TRY(t->out_items.ensure((size_t)t->n_users * K));
  RFM_TRY(t->out_scores.ensure((size_t)t->n_users * K));
  ExactArgs e;
  e.A = t->A.p;
  e.C = t->C.p;
  e.alpha = t->has_alpha ? t->alpha.p : nullptr;
  e.beta = t->has_beta ? t->beta.p : nullptr;
  e.bias = t->bias;
  e.n_users = t->n_users;
  e.n_items = t->n_items;
  e.k = t->k;
  e.item_begin = (int)item_begin;
  e.item_end = (int)item_end;
  int64_t n_failed = 0;
  // the tensor-core path needs a tile-aligned start (item shards are cut at multiples of 256) and k <= 128
  const bool tensor_path = mode == 0 && t->kb <= MAX_KB && item_begin % BN == 0;
  if (tensor_path) {
    int kc = std::max(2 * K, K + 16);
    kc = (kc + 7) / 8 * 8;
    const int tile_begin = (int)(item_begin / BN);
    const int n_item_tiles = (int)((item_end + BN - 1) / BN) - tile_begin;
    const int n_user_blocks = (int)(t->n_users_pad / BM);
    // split the catalog so that the grid covers the SMs about twice when there are few user blocks
    int n_splits = std::max(1, std::min(n_item_tiles, (2 * ctx->sm_count + n_user_blocks - 1) / n_user_blocks));
    const int tiles_per_split = (n_item_tiles + n_splits - 1) / n_splits;
    n_splits = (n_item_tiles + tiles_per_split - 1) / tiles_per_split;
    const int n_lists = n_splits * EPI_HALVES;          // candidate lists per user
    const size_t rows = (size_t)n_lists * t->n_users_pad;
    RFM_TRY(t->cand_score.ensure(rows * kc));
    RFM_TRY(t->cand_item.ensure(rows * kc));
    RFM_TRY(t->cand_tau.ensure(rows));
    FilterArgs fa;
    fa.beta = t->beta32.p;
    fa.tile_begin = tile_begin;
    fa.n_item_tiles = n_item_tiles;
    fa.item_end = (int)item_end;
    fa.tiles_per_split = tiles_per_split;
    fa.kc = kc;
    fa.n_users_pad = (int)t->n_users_pad;
    fa.cand_score = t->cand_score.p;
    fa.cand_item = t->cand_item.p;
    fa.cand_tau = t->cand_tau.p;
    fa.glist_score = nullptr;
    fa.glist_item = nullptr;
    const int n_groups = kc / 8;
    const size_t smem_limit = 227 * 1024;
    const size_t smem_with_lists = t->kb == 1 ? FilterSmem<1>::bytes(n_groups, kc) : FilterSmem<2>::bytes(n_groups, kc);
    const int kc_smem = smem_with_lists <= smem_limit ? kc : 0;     // otherwise the lists go to global memory
    if (!kc_smem) {
      RFM_TRY(t->glist_score.ensure((size_t)n_user_blocks * n_splits * EPI_HALVES * kc * BM));
      RFM_TRY(t->glist_item.ensure((size_t)n_user_blocks * n_splits * EPI_HALVES * kc * BM));
      fa.glist_score = t->glist_score.p;
      fa.glist_item = t->glist_item.p;
    }
    const dim3 grid(n_user_blocks, n_splits);
    if (t->kb == 1) {
      const size_t smem = FilterSmem<1>::bytes(n_groups, kc_smem);
      RFM_REQUIRE(smem <= smem_limit, "rfm_topk_run: K=%d needs %zu bytes of shared memory", K, smem);
      RFM_CUDA(cudaFuncSetAttribute(score_filter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      auto score_filter = score_filter_kernel<1>;
      RFM_LAUNCH(ctx, score_filter, grid, SCORE_THREADS, smem, t->tmap_a, t->tmap_c, fa);
    } else {
      const size_t smem = FilterSmem<2>::bytes(n_groups, kc_smem);
      RFM_REQUIRE(smem <= smem_limit, "rfm_topk_run: K=%d needs %zu bytes of shared memory", K, smem);
      RFM_CUDA(cudaFuncSetAttribute(score_filter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      auto score_filter = score_filter_kernel<2>;
      RFM_LAUNCH(ctx, score_filter, grid, SCORE_THREADS, smem, t->tmap_a, t->tmap_c, fa);
    }
    RFM_CUDA(cudaMemsetAsync(t->n_fail.p, 0, sizeof(uint32_t), ctx->stream));
    const size_t rs_smem = (size_t)RESCORE_WARPS * n_lists * kc * 12;
    RFM_REQUIRE(rs_smem <= 200 * 1024, "rfm_topk_run: %d candidates per user do not fit the re-scoring kernel",
                n_lists * kc);
    RFM_CUDA(cudaFuncSetAttribute(score_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem));
    const int rgrid = (int)std::min<int64_t>((t->n_users + RESCORE_WARPS - 1) / RESCORE_WARPS,
                                             (int64_t)ctx->sm_count * 8);
    RFM_LAUNCH(ctx, score_rescore_kernel, rgrid, RESCORE_WARPS * 32, rs_smem, e, t->cand_score.p, t->cand_item.p,
               t->cand_tau.p,
               n_lists, (int)t->n_users_pad, kc, (int)K, t->a_norm.p, t->c_norm_max.p, t->beta_abs_max.p,
               t->out_items.p, t->out_scores.p, t->fail_list.p, t->n_fail.p);
    // users whose pruning could not be proven are ranked exactly against the whole catalog
    const int fgrid = (int)std::min<int64_t>(t->n_users, (int64_t)ctx->sm_count * 4);
    RFM_LAUNCH(ctx, score_exact_kernel, fgrid, 256, 0, e, t->fail_list.p, t->n_fail.p, (int64_t)0, (int)K,
               t->out_items.p, t->out_scores.p);
    uint32_t nf = 0;
    RFM_CUDA(cudaMemcpyAsync(&nf, t->n_fail.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    RFM_CUDA(cudaStreamSynchronize(ctx->stream));
    n_failed = nf;
  } else {
    const int fgrid = (int)std::min<int64_t>(t->n_users, (int64_t)ctx->sm_count * 4);
    RFM_LAUNCH(ctx, score_exact_kernel, fgrid, 256, 0, e, (const int32_t *)nullptr, (const uint32_t *)nullptr,
               t->n_users, (int)K, t->out_items.p, t->out_scores.p);
  }
  if (out_items) {   // NULL: the caller reads the result on the device (rfm_topk_result_ptr_dev)
    RFM_CUDA(cudaMemcpyAsync(out_items, t->out_items.p, (size_t)t->n_users * K * 4, cudaMemcpyDeviceToHost,
                             ctx->stream));
    RFM_CUDA(cudaMemcpyAsync(out_scores, t->out_scores.p, (size_t)t->n_users * K * 8, cudaMemcpyDeviceToHost,
                             ctx->stream));
  }
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (stats) {
    stats[0] = tensor_path ? 1 : 0;
    stats[1] = n_failed;
  }
  return RFM_OK;
}

int rfm_topk_result_ptr_dev(rfm_topk *t, void **items_dev, void **scores_dev) {
  RFM_REQUIRE(t && items_dev && scores_dev, "rfm_topk_result_ptr_dev: NULL argument");
  RFM_REQUIRE(t->out_items.p && t->out_scores.p, "rfm_topk_result_ptr_dev: call rfm_topk_run first");
  *items_dev = t->out_items.p;
  *scores_dev = t->out_scores.p;
  return RFM_OK;
}

int rfm_topk_merge_dev(rfm_ctx *ctx, int64_t n_users, int32_t K, int32_t n_lists, const int32_t *items_dev,
                       const double *scores_dev, int32_t *out_items, double *out_scores) {
  RFM_REQUIRE(ctx && items_dev && scores_dev && out_items && out_scores, "rfm_topk_merge_dev: NULL argument");
  RFM_REQUIRE(n_users >= 1 && K >= 1 && n_lists >= 1, "rfm_topk_merge_dev: bad sizes");
  RFM_CUDA(cudaSetDevice(ctx->device));
  DevBuf<int32_t> oi;
  DevBuf<double> os;
  RFM_TRY(oi.alloc((size_t)n_users * K));
  RFM_TRY(os.alloc((size_t)n_users * K));
  const int grid = (int)std::min<int64_t>((n_users + 7) / 8, (int64_t)ctx->sm_count * 8);
  RFM_LAUNCH(ctx, topk_merge_kernel, grid, 256, 0, items_dev, scores_dev, n_users, (int)K, (int)n_lists, oi.p, os.p);
  RFM_CUDA(cudaMemcpyAsync(out_items, oi.p, (size_t)n_users * K * 4, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(out_scores, os.p, (size_t)n_users * K * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

}  // extern "C"
