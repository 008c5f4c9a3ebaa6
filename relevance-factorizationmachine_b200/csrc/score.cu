// score.cu -- full-catalog scoring + exact per-user top-K (SURVEY.md Appendix A.4, section 8 f-rows).
//
// FM and MF scores decompose as  score(u, i) = bias + alpha[u] + beta[i] + <A_u, C_i>  when a row is
// [user-side features | item-side features] (for MF: A = P, C = Q, alpha = b_u, beta = b_i, bias = b;
// reference src/mf.py:165-170, src/fm.py:125-132). The reference never scores the full user x item
// grid; the parity statement for this path is "equal to the reference's predict on the
// Cartesian-product rows followed by its per-user argsort", checked in tests/test_score_gpu.py.
//
// Pipeline (all on the device):
//   1. A, C -> bf16 tiles, K padded to 64, rows padded to the tile; row norms and rounding-residual norms in
//      float64; the item terms beta as a 16-wide bf16 operand (three-way split, exact to 2^-24).
//   2. score_pass_kernel<PASS 0> ("score_sample"): a warp-specialised tcgen05 GEMM over every stride-th 256-item
//      tile. One CTA keeps up to four blocks of 128 users resident; TMA (cp.async.bulk.tensor, 128B swizzle)
//      stages the item tiles, one elected thread issues tcgen05.mma (M=128, N=256, K=16, bf16 -> fp32 in TMEM,
//      two accumulator stages; the first MMA of a step adds beta), and 16 epilogue warps read the accumulator
//      with tcgen05.ld: thread r streams the scores of user row r through FMNMX3 and writes one maximum per
//      group of 64 items. Scores are never written to memory.
//   3. score_threshold_kernel: per user, the K-th largest group maximum. K different items reach it, so it is a
//      lower bound on the K-th best approximate score; tau = bound - 2 eps(u), where eps(u) bounds
//      |approximate - exact| rigorously from the data (score_error_bound).
//   4. score_pass_kernel<PASS 1> ("score_collect"): the same GEMM over every tile; the epilogue compares each
//      32-column slice's maximum with tau and appends the few items that reach it (about K * stride per user)
//      to the user's candidate buffer. Every item whose exact score can reach the exact K-th best is collected.
//   5. score_rescore_kernel: candidates within 2 eps of the K-th best approximate score are re-scored in
//      float64 (a candidate per lane) and ranked with the canonical tie rule (score descending, larger item id
//      first).
//   6. score_exact_kernel: users whose candidate buffer overflowed (or every user, in exact mode / for k > 128)
//      are scored in float64 against the whole catalog.
// The result is therefore always the exact float64 top-K; the tensor-core passes only prune. At k = 64 the passes
// run at ~1,300-1,450 cycles per 128 x 256 step against a 640-cycle MMA step (tensor pipe ~60 % busy): hand-shake
// gaps, an MMA slowed by shared-memory contention with the copy engine, and per-CTA prologue -- not TMEM bandwidth
// (372 B/clk/SM measured) and not the epilogue math; DESIGN.md 4.4, tools/*.cu.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>

#include "common.cuh"

using namespace rfm;

namespace {

constexpr int BM = 128;          // users per CTA tile (= TMEM lanes)
constexpr int BN = 256;          // items per MMA tile (= TMEM columns per accumulator stage)
constexpr int BK = 64;           // bf16 elements per K block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
#ifndef RFM_EPI_WARPS
#define RFM_EPI_WARPS 16
#endif
constexpr int EPI_WARPS = RFM_EPI_WARPS;             // EPI_WARPS/4 per TMEM lane quarter; each owns a slice of a tile's columns
constexpr int SCORE_THREADS = 128 + 32 * EPI_WARPS;  // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4.. epilogue
// every epilogue warp drains every accumulator stage: 4 TMEM lane quarters x EPI_PARTS column ranges, so the MMA of
// step i + 1 overlaps the whole epilogue of step i (two warp groups ping-ponging the stages were slower)
constexpr int EPI_PARTS = EPI_WARPS / 4;
static_assert(EPI_WARPS == 16 || EPI_WARPS == 8, "the drain handles 64 or 128 columns per warp");
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
// the same with the 32-bit shared address computed once by the caller: inside the per-step loops the generic ->
// shared conversion (and the 1 KB alignment of the dynamic shared memory behind it) was re-derived every step
__device__ __forceinline__ void mbar_wait_at(uint32_t bar_addr, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar_addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive_at(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
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
    __nanosleep(32);
  }
}
// one lane of a converged warp; the code around it stays warp-uniform, so the operands of the tcgen05 / TMA
// instructions it guards are known to be uniform and go to uniform registers directly. Inside an `if (lane == 0)`
// region the compiler must assume per-thread values and wraps every such operand in a re-execution loop
// (R2UR + BRA.U.ANY, ~10 instructions each).
__device__ __forceinline__ bool elect_one() {
  uint32_t leader;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(leader));
  return leader != 0;
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

// K-major operand tile without swizzle (canonical "interleave" layout): element (row, k) of a 16-bit type lives at
// (row % 8) * 16 + (row / 8) * SBO + (k / 8) * LBO + (k % 8) * 2 bytes.
__device__ __forceinline__ uint64_t umma_desc_interleaved(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// kind::f16 instruction descriptor: D = fp32 (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1), both
// K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
constexpr uint32_t UMMA_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);

struct PassArgs {
  const __nv_bfloat16 *beta16;   // [n_items_pad / 8][2][8][8]: item terms as three bf16 parts, canonical operand order
  int tile_begin;          // first 256-item tile of the catalog range being ranked (item-sharded calls)
  int tile_stride;         // pass 1 visits every tile_stride-th tile (the sample), pass 2 every tile
  int n_visit;             // tiles this pass visits
  int tiles_per_split;     // visited tiles per CTA (blockIdx.y)
  int item_end;            // end of the catalog range (padding columns of the last tile are never candidates)
  int user_block0;         // first 128-user block of this chunk of users
  int n_user_blocks;       // 128-user blocks in this chunk
  int ub;                  // user blocks per CTA (<= PassSmem::UB)
  int64_t n_groups;        // pass 1: group maxima per user row
  float *gmax;             // pass 1 out: [chunk users (padded)][n_groups]
  const float *tau;        // pass 2 in: [chunk users (padded)] collect threshold (+inf for padding rows)
  int kc;                  // pass 2: capacity of a user's candidate buffer
  uint32_t *cand_cnt;      // pass 2 out: [chunk users (padded)] candidates seen (may exceed kc: overflow)
  int2 *cand;              // pass 2 out: [chunk users (padded)][kc] (item, approximate score bits)
};

constexpr int WARP_COLS = BN / EPI_PARTS;      // columns of an accumulator one epilogue warp filters
static_assert(WARP_COLS % 64 == 0, "an epilogue warp owns whole groups of 64 columns");
constexpr int MAX_GCOLS = 64;   // widest pass-1 group (the threshold kernel merges adjacent groups on load)

// The item term beta_i is added by the tensor core as well: every step starts with one extra K = 16 MMA of a
// constant operand [1 1 1 0 ... 0] against [b0 b1 b2 0 ... 0], beta split into three bf16 terms (exact to
// 2^-24 |beta|), so the accumulator the epilogue reads already holds <A_u, C_i> + beta_i. Both small operands
// use the no-swizzle K-major canonical layout (8-row x 16-byte core matrices; LBO = 128 B between the two K
// halves, SBO = 256 B between 8-row groups); the item operand is stored in global memory in exactly that order,
// so a tile's 8 KB arrive with one 1-D bulk copy.
// A CTA keeps UB blocks of 128 users resident and runs every B tile it stages against all of them (the
// accumulator stages alternate between consecutive (tile, user block) steps), so the catalog is streamed from L2
// once per UB * 128 users: with one block per CTA the B-tile traffic, not the tensor pipe, bounds the pass.
template <int KB>
struct PassSmem {
  static constexpr int UB = KB == 1 ? 4 : 2;
  static constexpr int STAGES = KB == 1 ? 3 : 2;
  static constexpr int A_BLOCK_BYTES = KB * BM * 128;
  static constexpr int A_BYTES = UB * A_BLOCK_BYTES;
  static constexpr int ONES_OFF = A_BYTES;                           // the [1 1 1 0 ...] operand, BM x 16 bf16
  static constexpr int ONES_BYTES = BM * 32;
  static constexpr int B_OFF = ONES_OFF + ONES_BYTES;
  static constexpr int B_MAIN_BYTES = KB * BN * 128;
  static constexpr int B_BETA_BYTES = BN * 32;                       // the item-term operand, BN x 16 bf16
  static constexpr int B_STAGE_BYTES = B_MAIN_BYTES + B_BETA_BYTES;
  static constexpr int BAR_OFF = B_OFF + STAGES * B_STAGE_BYTES;
  static constexpr size_t BYTES = 1024 + BAR_OFF + 256;
};

// 1-D bulk copy global -> shared, completing on an mbarrier (the item terms of a tile ride with its B tile)
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tc_ld_issue_x32(uint32_t taddr, uint32_t (&v)[32]) {
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
}
// wait for the outstanding tcgen05.ld; the registers are listed so that no use of them is scheduled above it
__device__ __forceinline__ void tc_ld_wait_x32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]),
                 "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),
                 "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
// three-input max (FMNMX3): one instruction per three elements
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// m8[i] = max of columns 8i .. 8i+7 of a 32-column slice of the accumulator (which already includes beta)
__device__ __forceinline__ void slice_max(const uint32_t (&v)[32], float (&m8)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    m8[i] = max3(max3(__uint_as_float(v[8 * i]), __uint_as_float(v[8 * i + 1]), __uint_as_float(v[8 * i + 2])),
                 max3(__uint_as_float(v[8 * i + 3]), __uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5])),
                 fmaxf(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7])));
}

// pass 1: write the group maxima of one 32-column slice (groups of GCOLS <= 32 columns)
template <int GCOLS>
__device__ __forceinline__ void store_group_max(float *__restrict__ dst, const float (&m8)[4]) {
  if (GCOLS == 8) {
    *reinterpret_cast<float4 *>(dst) = make_float4(m8[0], m8[1], m8[2], m8[3]);
  } else if (GCOLS == 16) {
    *reinterpret_cast<float2 *>(dst) = make_float2(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
  } else {
    *dst = fmaxf(max3(m8[0], m8[1], m8[2]), m8[3]);
  }
}

// pass 2: append every item of a slice whose approximate score reaches the user's threshold (cold path). Only
// the 8-column sub-groups whose maximum passes are walked: typically one lane and one sub-group per call.
// The slot comes from a global atomic whose round trip (~1 us) must not stall the warp: an append only issues
// the atomic and parks (slot, row, item); the store happens at the thread's next append (or at the end).
struct PendingAppend {
  uint32_t pos;
  int row, item;      // item < 0: nothing pending
  float score;
};
__device__ __forceinline__ void flush_append(const PendingAppend &p, int kc, int2 *__restrict__ cand) {
  if (p.item >= 0 && p.pos < (uint32_t)kc) cand[(size_t)p.row * kc + p.pos] = make_int2(p.item, __float_as_int(p.score));
}
__device__ __forceinline__ void collect_slice(const uint32_t (&v)[32], const float (&m8)[4], float tau, int item_first,
                                              int item_end, int kc, int row, uint32_t *__restrict__ cand_cnt,
                                              int2 *__restrict__ cand, PendingAppend &pend) {
  // warp-uniform control flow: a sub-group of 8 columns is examined only if some lane's maximum over it passes,
  // and then every lane builds its 8-bit pass mask without branching (all but one or two lanes get zero)
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (__any_sync(FULL, m8[i] >= tau)) {
      uint32_t mask = 0u;
#pragma unroll
      for (int j = 0; j < 8; ++j) mask |= (__uint_as_float(v[8 * i + j]) >= tau ? 1u : 0u) << j;
      while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        uint32_t bits = v[8 * i];
#pragma unroll
        for (int jj = 1; jj < 8; ++jj)
          if (j == jj) bits = v[8 * i + jj];
        const int item = item_first + 8 * i + j;
        if (item < item_end) {
          flush_append(pend, kc, cand);
          pend.pos = atomicAdd(cand_cnt + row, 1u);
          pend.row = row;
          pend.item = item;
          pend.score = __uint_as_float(bits);
        }
      }
    }
}

// The tensor-core pass. PASS 0 (the sample): per user row, the maximum approximate score of every group of GCOLS
// items of the visited tiles goes to gmax. PASS 1 (collect): every item whose approximate score reaches the
// user's threshold is appended to the user's candidate buffer. Neither pass keeps a sorted list: thread r
// streams row r of the accumulator through FADD2 + FMNMX3 (one instruction per element) and the append
// happens for a handful of items per user over the whole catalog.
template <int KB, int PASS, int GCOLS>
__global__ void __launch_bounds__(SCORE_THREADS, 1)
score_pass_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_c,
                  const PassArgs a) {
  using L = PassSmem<KB>;
  constexpr int STAGES = L::STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char *sA = smem;
  unsigned char *sOnes = smem + L::ONES_OFF;
  unsigned char *sB = smem + L::B_OFF;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::BAR_OFF);
  uint64_t *full = bars, *empty = bars + STAGES, *tfull = bars + 2 * STAGES, *tempty = bars + 2 * STAGES + 2;
  uint64_t *afull = bars + 2 * STAGES + 4;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int UB = L::UB;
  const int split = blockIdx.y;
  const int block0 = blockIdx.x * a.ub;                 // first user block of this CTA, inside the chunk
  const int nh = min(a.ub, a.n_user_blocks - block0);   // user blocks this CTA holds
  const int visit0 = split * a.tiles_per_split;       // index of this CTA's first visited tile
  int n_tiles = a.n_visit - visit0;
  if (n_tiles > a.tiles_per_split) n_tiles = a.tiles_per_split;
  if (n_tiles < 0) n_tiles = 0;
  auto tile_of = [&](int it) { return a.tile_begin + (visit0 + it) * a.tile_stride; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull + s, 1);
      mbar_init(tempty + s, EPI_WARPS);       // one arrival per epilogue warp
    }
    mbar_init(afull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // the constant operand: row r = [1 1 1 0 ... 0] (16 bf16), canonical no-swizzle K-major layout
  for (int i = threadIdx.x; i < BM * 2; i += SCORE_THREADS) {
    const int row = i >> 1, khalf = i & 1;
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (khalf == 0) {
      w.x = 0x3F803F80u;   // bf16 1.0, 1.0
      w.y = 0x00003F80u;   // bf16 1.0, 0.0
    }
    *reinterpret_cast<uint4 *>(sOnes + (row >> 3) * 256 + khalf * 128 + (row & 7) * 16) = w;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes, read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer: uniform control flow, one elected lane issues the copies =====
    if (elect_one()) {
      mbar_expect_tx(afull, nh * L::A_BLOCK_BYTES);
      for (int h = 0; h < nh; ++h)
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(sA + h * L::A_BLOCK_BYTES + kb * BM * 128, &tmap_a, afull, kb * BK,
                      (a.user_block0 + block0 + h) * BM);
    }
    __syncwarp();
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait_relaxed(empty + s, ph ^ 1);
      const int tile = tile_of(it);
      if (elect_one()) {
        mbar_expect_tx(full + s, L::B_STAGE_BYTES);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_2d(sB + s * L::B_STAGE_BYTES + kb * BN * 128, &tmap_c, full + s, kb * BK, tile * BN);
        bulk_load_1d(sB + s * L::B_STAGE_BYTES + L::B_MAIN_BYTES, a.beta16 + (size_t)tile * BN * 16,
                     L::B_BETA_BYTES, full + s);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the loop (uniform control flow), one elected lane issues =====
    mbar_wait(afull, 0);
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait_relaxed(full + s, ph);
      for (int h = 0; h < nh; ++h) {
        const int step = it * nh + h;
        const int acc = step & 1;
        const uint32_t aph = (step >> 1) & 1;
        mbar_wait_relaxed(tempty + acc, aph ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        const uint32_t ones = smem_u32(sOnes), bop = smem_u32(sB + s * L::B_STAGE_BYTES + L::B_MAIN_BYTES);
        const uint32_t a_blk = smem_u32(sA + h * L::A_BLOCK_BYTES), b_blk = smem_u32(sB + s * L::B_STAGE_BYTES);
        if (elect_one()) {
          // accumulator = 1 * beta (overwrite), then += <A_u, C_i> over the K blocks
          tc_mma_bf16(d, umma_desc_interleaved(ones, 128, 256), umma_desc_interleaved(bop, 128, 256), UMMA_IDESC, 0u);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
            for (int k4 = 0; k4 < BK / UMMA_K; ++k4) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle atom
              tc_mma_bf16(d, umma_desc_sw128(a_blk + kb * BM * 128 + k4 * 32),
                          umma_desc_sw128(b_blk + kb * BN * 128 + k4 * 32), UMMA_IDESC, 1u);
            }
          }
          tc_commit(tfull + acc);                    // the accumulator is complete
          if (h == nh - 1) tc_commit(empty + s);     // the B stage is free once the MMAs of every user block have read it
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread r streams the scores of row r of every user block =====
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int part = (warp - 4) >> 2;       // its range of the accumulator's columns
    const int r = q * 32 + lane;
    PendingAppend pend;
    pend.pos = 0u;
    pend.row = 0;
    pend.item = -1;
    pend.score = 0.f;
    float tau_h[UB];
#pragma unroll
    for (int h = 0; h < UB; ++h)
      tau_h[h] = (PASS == 1 && h < nh) ? a.tau[(size_t)(block0 + h) * BM + r] : INFINITY;
    // Loop-invariant pieces are hoisted and (tile, user block) advance by counters: the epilogue's instruction
    // count per step, not its math, is what competes with the MMA issue for the four schedulers.
    const int c = part * WARP_COLS;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c;
    const int row0 = block0 * BM + r;                      // row of user block h = row0 + h * BM (inside the chunk)
    const int kc = a.kc, item_end = a.item_end;
    uint32_t *const cand_cnt = a.cand_cnt;
    int2 *const cand = a.cand;
    float *const gbase = PASS == 0 ? a.gmax + (size_t)row0 * a.n_groups + (int64_t)visit0 * (BN / GCOLS) + c / GCOLS
                                   : nullptr;
    const size_t gstride_h = (size_t)BM * a.n_groups;      // gmax: from one user block to the next
    const int n_steps = n_tiles * nh;
    const uint32_t tfull_at = smem_u32(tfull), tempty_at = smem_u32(tempty);      // + 8 * stage
    int h = 0, item0 = tile_of(0) * BN + c;
    const int item_step = a.tile_stride * BN;
    float *gtile = gbase;                                  // gmax slot of (current tile, user block 0)
    for (int step = 0; step < n_steps; ++step) {
      const int acc = step & 1;
      const uint32_t aph = (step >> 1) & 1;
      float tau = tau_h[0];
#pragma unroll
      for (int hh = 1; hh < UB; ++hh)
        if (h == hh) tau = tau_h[hh];
      const int row = row0 + h * BM;
      mbar_wait_at(tfull_at + acc * 8, aph);
      tc_fence_after();
      const uint32_t tbase = tlane + acc * BN;
      // Drain first, work later: both of the warp's loads are issued back to back, and the accumulator stage is
      // handed back to the MMA issuer as soon as they have landed in registers. The max tree, the group-maximum
      // store (32 scattered sectors per instruction) and above all the cold append path (a global atomic) then
      // overlap the next MMA instead of sitting between this step's MMA and the next-but-one: every step has
      // some warp on the cold path, and all 16 must arrive before the stage can be reused.
#pragma unroll
      for (int cc = 0; cc < WARP_COLS; cc += 64) {       // one chunk with 16 warps, two with 8
        uint32_t v0[32], v1[32];
        tc_ld_issue_x32(tbase + cc, v0);
        tc_ld_issue_x32(tbase + cc + 32, v1);
        tc_ld_wait_x32(v0);
        tc_ld_wait_x32(v1);
        if (cc + 64 == WARP_COLS) {                      // everything of this stage is in registers: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_at(tempty_at + acc * 8);
        }
        float m0[4], m1[4];
        slice_max(v0, m0);
        slice_max(v1, m1);
        if (PASS == 1) {
          const float mx0 = fmaxf(max3(m0[0], m0[1], m0[2]), m0[3]), mx1 = fmaxf(max3(m1[0], m1[1], m1[2]), m1[3]);
          if (__any_sync(FULL, fmaxf(mx0, mx1) >= tau)) {
            if (__any_sync(FULL, mx0 >= tau))
              collect_slice(v0, m0, tau, item0 + cc, item_end, kc, row, cand_cnt, cand, pend);
            if (__any_sync(FULL, mx1 >= tau))
              collect_slice(v1, m1, tau, item0 + cc + 32, item_end, kc, row, cand_cnt, cand, pend);
          }
        } else {
          float *dst = gtile + h * gstride_h + cc / GCOLS;
          if (GCOLS == 64) {
            *dst = max3(max3(m0[0], m0[1], m0[2]), max3(m0[3], m1[0], m1[1]), fmaxf(m1[2], m1[3]));
          } else {
            store_group_max<GCOLS>(dst, m0);
            store_group_max<GCOLS>(dst + 32 / GCOLS, m1);
          }
        }
      }
      if (++h == nh) {        // next tile
        h = 0;
        item0 += item_step;
        gtile += BN / GCOLS;
      }
    }
    if (PASS == 1) flush_append(pend, a.kc, a.cand);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// eps(u) >= |approximate score - exact ranking score| for every item, where the approximate score is what the
// tensor core accumulates, sum_f bf16(a_f) bf16(c_f) + b0 + b1 + b2, and the exact one is <a, c> + beta:
//   operands:      |sum a^c^ - sum ac| <= ||a^|| ||c^ - c|| + ||a^ - a|| ||c||   (Cauchy-Schwarz; the residual
//                  norms are computed from the data, not from the worst-case 2^-8 relative rounding error)
//   beta split:    |b0 + b1 + b2 - beta| <= 2^-24 |beta|
//   accumulation:  n = kpad + 16 products (exact in float32) summed with at most one ulp of error each
//                  (the tensor core may truncate): <= n 2^-23 (||a^|| ||c^|| + |beta|); doubled for margin.
__device__ __forceinline__ double score_error_bound(double a_norm, double a_res, double c_norm_max, double c_res_max,
                                                    double beta_abs_max, int kpad) {
  const double a_hat = a_norm + a_res, c_hat = c_norm_max + c_res_max;
  const double operands = a_hat * c_res_max + a_res * c_norm_max;
  const double accumulation = ldexp((double)(2 * (kpad + 16)), -23) * (a_hat * c_hat + beta_abs_max);
  return (operands + accumulation + ldexp(beta_abs_max, -23)) * (1.0 + 0x1p-20) + 1e-300;
}

__device__ __forceinline__ uint32_t thr_keep(float x) {    // order-preserving uint key of a float
  const uint32_t bits = __float_as_uint(x);
  return bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t key) {
  return __uint_as_float((key & 0x80000000u) ? (key ^ 0x80000000u) : ~key);
}

// K-th largest of the keys a warp holds in registers (NJ per lane, 0 = none). For small K, K rounds of "take the
// maximum out" (one warp reduction each) beat the 32 rounds of the bit-wise bisection.
template <int NJ>
__device__ __forceinline__ uint32_t warp_kth_largest(uint32_t (&key)[NJ], int K, int lane) {
  if (K <= 24) {
    uint32_t kth = 0u;
    for (int r = 0; r < K; ++r) {
      uint32_t m = 0u;
#pragma unroll
      for (int j = 0; j < NJ; ++j) m = max(m, key[j]);
      kth = __reduce_max_sync(FULL, m);
      if (kth == 0u) break;                                   // fewer than K keys
      const unsigned holders = __ballot_sync(FULL, m == kth);
      if (lane == __ffs(holders) - 1) {                       // one holder drops one instance of the maximum
        bool done = false;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (!done && key[j] == kth) {
            key[j] = 0u;
            done = true;
          }
      }
    }
    return kth;
  }
  uint32_t thr = 0u;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand_thr = thr | (1u << bit);
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) cnt += key[j] >= cand_thr;
    cnt = __reduce_add_sync(FULL, cnt);
    if (cnt >= K) thr = cand_thr;
  }
  return thr;
}

// One warp per user: the K-th largest group maximum of the sample is a lower bound on the user's K-th best
// approximate score (K different items reach it). tau = that bound - 2 eps: every item whose EXACT score can
// reach the exact K-th best has an approximate score >= tau (eps bounds |approximate - exact|).
constexpr int SELECT_WARPS = 4;
// NJ > 0: every lane keeps its NJ keys (groups lane, lane + 32, ...) in registers; NJ = 0: keys staged in shared memory
template <int NJ>
__global__ void __launch_bounds__(SELECT_WARPS * 32)
score_threshold_kernel(const float *__restrict__ gmax, int64_t n_groups, int fold, int64_t n_rows,
                       int64_t n_users_chunk, int K,
                       const double *__restrict__ a_norm, const double *__restrict__ a_res,
                       const double *__restrict__ c_norm_max, const double *__restrict__ c_res_max,
                       const double *__restrict__ beta_abs_max, int kpad, float *__restrict__ tau,
                       float *__restrict__ eps_out, uint32_t *__restrict__ cand_cnt, uint32_t *__restrict__ gtop) {
  extern __shared__ __align__(16) unsigned char th_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t *skey = reinterpret_cast<uint32_t *>(th_smem) + (size_t)wid * (NJ > 0 ? 0 : n_groups);
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  auto to_key = [](float x) {              // order-preserving uint key
    const uint32_t bits = __float_as_uint(x);
    return bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
  };
  for (int64_t u = gw; u < n_rows; u += nw) {
    if (lane == 0) cand_cnt[u] = 0u;
    if (u >= n_users_chunk) {           // padding rows never collect
      if (lane == 0) tau[u] = INFINITY;
      continue;
    }
    // fold = 2 or 4: adjacent stored groups are merged on load (the maximum of group maxima is the maximum of the
    // wider group), shrinking the key set when the sample holds more groups than the threshold needs
    const float *g = gmax + u * n_groups * fold;
    auto group_max = [&](int64_t i) {
      if (fold == 1) return g[i];
      if (fold == 2) {
        const float2 p = *reinterpret_cast<const float2 *>(g + 2 * i);
        return fmaxf(p.x, p.y);
      }
      const float4 p = *reinterpret_cast<const float4 *>(g + 4 * i);
      return fmaxf(fmaxf(p.x, p.y), fmaxf(p.z, p.w));
    };
    uint32_t key[NJ > 0 ? NJ : 1];
    if (NJ > 0) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) key[j] = lane + 32 * j < n_groups ? to_key(group_max(lane + 32 * j)) : 0u;
    } else {
      for (int64_t i = lane; i < n_groups; i += 32) skey[i] = to_key(group_max(i));
      __syncwarp();
    }
    float t0 = -INFINITY;
    bool exported = false;
    if (NJ > 0 && K <= 24 && n_groups >= K) {
      // small K: K rounds of "take the maximum out"; the extracted values ARE the K largest, exported as they come
      uint32_t kth = 0u;
      for (int r = 0; r < K; ++r) {
        uint32_t m = 0u;
#pragma unroll
        for (int j = 0; j < (NJ > 0 ? NJ : 1); ++j) m = max(m, key[j]);
        kth = __reduce_max_sync(FULL, m);
        const unsigned holders = __ballot_sync(FULL, m == kth);
        if (lane == __ffs(holders) - 1) {
          bool done = false;
#pragma unroll
          for (int j = 0; j < (NJ > 0 ? NJ : 1); ++j)
            if (!done && key[j] == kth) {
              key[j] = 0u;
              done = true;
            }
          if (gtop) gtop[u * (int64_t)K + r] = kth;
        }
      }
      t0 = key_to_float(kth);
      exported = true;
    } else if (n_groups >= K) {
      uint32_t thr = 0u;                // bisection on the key bits for the K-th largest key
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand_thr = thr | (1u << bit);
        int cnt = 0;
        if (NJ > 0) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) cnt += key[j] >= cand_thr;
        } else {
          for (int64_t i = lane; i < n_groups; i += 32) cnt += skey[i] >= cand_thr;
        }
        cnt = __reduce_add_sync(FULL, cnt);
        if (cnt >= K) thr = cand_thr;
      }
      const uint32_t bits = (thr & 0x80000000u) ? (thr ^ 0x80000000u) : ~thr;
      t0 = __uint_as_float(bits);
    }
    __syncwarp();
    if (gtop && !exported) {
      // item-sharded runs: this rank's K largest group maxima (as keys, any order) go to the exchange region; the
      // K-th largest of the union over all ranks is the global bound (score_global_threshold_kernel)
      uint32_t *out = gtop + u * (int64_t)K;
      const uint32_t none = 0x007FFFFFu;                       // key of -inf
      if (n_groups >= K) {
        int base = 0;
        const uint32_t lt = (1u << lane) - 1u;
        if (NJ > 0) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const bool up = key[j] > thr_keep(t0);
            const uint32_t m = __ballot_sync(FULL, up);
            if (up) out[base + __popc(m & lt)] = key[j];
            base += __popc(m);
          }
        } else {
          for (int64_t i0 = 0; i0 < n_groups; i0 += 32) {
            const int64_t i = i0 + lane;
            const bool up = i < n_groups && skey[i] > thr_keep(t0);
            const uint32_t m = __ballot_sync(FULL, up);
            if (up) out[base + __popc(m & lt)] = skey[i];
            base += __popc(m);
          }
        }
        for (int r = base + lane; r < K; r += 32) out[r] = thr_keep(t0);     // the ties at the K-th value
      } else {
        for (int r = lane; r < K; r += 32) {
          uint32_t v = none;
          if (r < n_groups) v = to_key(group_max(r));
          out[r] = v;
        }
      }
    }
    if (lane == 0) {
      const double eps = score_error_bound(a_norm[u], a_res[u], *c_norm_max, *c_res_max, *beta_abs_max, kpad);
      tau[u] = t0 > -INFINITY ? __double2float_rd((double)t0 - 2.0 * eps) : -INFINITY;
      eps_out[u] = __double2float_ru(eps);
    }
  }
}

// ---- operand preparation ---------------------------------------------------------------------------------
// float64 [rows][k] -> bf16 [rows_pad][kpad] (zero padded), float64 row norms and the norms of the rounding
// residuals row - bf16(row) (what the error bound of the tensor-core scores is made of)
__global__ void to_bf16_kernel(const double *__restrict__ in, int64_t rows, int k, int64_t rows_pad, int kpad,
                               __nv_bfloat16 *__restrict__ out, double *__restrict__ norms,
                               double *__restrict__ res_norms) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t rrow = gw; rrow < rows_pad; rrow += nw) {
    double s = 0.0, d = 0.0;
    for (int f = lane; f < kpad; f += 32) {
      const double v = (rrow < rows && f < k) ? in[rrow * k + f] : 0.0;
      const __nv_bfloat16 h = __double2bfloat16(v);
      out[rrow * kpad + f] = h;
      const double e = v - (double)__bfloat162float(h);
      s += v * v;
      d += e * e;
    }
    s = warp_sum(s);
    d = warp_sum(d);
    if (lane == 0 && rrow < rows) {
      norms[rrow] = sqrt(s) * (1.0 + 0x1p-40);        // rounded up: these feed an upper bound
      res_norms[rrow] = sqrt(d) * (1.0 + 0x1p-40);
    }
  }
}

// item terms as a bf16 MMA operand: beta = b0 + b1 + b2 (bf16 each, residual <= 2^-24 |beta|) in columns 0..2 of a
// 16-wide row, rows stored in the canonical no-swizzle K-major order (see PassSmem). Items beyond the catalog
// get -inf so that they never reach a threshold.
__global__ void beta_operand_kernel(const double *__restrict__ beta, int64_t n_items, int64_t n_pad,
                                    __nv_bfloat16 *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * blockDim.x) {
    __nv_bfloat16 *row = out + (i >> 3) * 128 + (i & 7) * 8;      // first K half; the second is 64 elements on
    const __nv_bfloat16 zero = __float2bfloat16(0.f);
    __nv_bfloat16 b0 = zero, b1 = zero, b2 = zero;
    if (i >= n_items) {
      b0 = __float2bfloat16(-INFINITY);
    } else if (beta) {
      const double v = beta[i];
      b0 = __double2bfloat16(v);
      const double r1 = v - (double)__bfloat162float(b0);
      b1 = __double2bfloat16(r1);
      b2 = __double2bfloat16(r1 - (double)__bfloat162float(b1));
    }
    row[0] = b0;
    row[1] = b1;
    row[2] = b2;
    for (int k = 3; k < 8; ++k) row[k] = zero;
    for (int k = 0; k < 8; ++k) row[64 + k] = zero;
  }
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

// The same value computed by ONE lane (32 candidates of a user are re-scored in parallel, one per lane): 32
// strided partial sums, each the same fma chain a lane of warp_exact_dot runs, combined in the butterfly's order
// (lane 0's view of it: x[l] + x[l + o] for o = 16, 8, 4, 2, 1). au is the user's row staged in shared memory.
__device__ __forceinline__ double lane_exact_dot(const double *__restrict__ au, const double *__restrict__ ci, int k) {
  double x[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) x[l] = 0.0;
  for (int f0 = 0; f0 < k; f0 += 32) {
    if (f0 + 32 <= k) {
#pragma unroll
      // the row is read once: keep it out of L1 (ld.global.cg), 16 bytes at a time when rows are 16-byte aligned
      if ((k & 1) == 0) {
#pragma unroll
        for (int l = 0; l < 32; l += 2) {
          const double2 c2 = __ldcg(reinterpret_cast<const double2 *>(ci + f0 + l));
          x[l] += au[f0 + l] * c2.x;
          x[l + 1] += au[f0 + l + 1] * c2.y;
        }
      } else {
#pragma unroll
        for (int l = 0; l < 32; ++l) x[l] += au[f0 + l] * __ldcg(ci + f0 + l);
      }
    } else {
#pragma unroll
      for (int l = 0; l < 32; ++l)
        if (f0 + l < k) x[l] += au[f0 + l] * ci[f0 + l];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int l = 0; l < o; ++l) x[l] += x[l + o];
  }
  return x[0];
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

// one warp per user: exact float64 scores of the collected candidates and the exact top-K among them with the
// canonical tie rule. A user whose buffer overflowed goes to the exact fallback instead. The warp's
// candidates (item, exact score) are staged in shared memory so the K selection rounds never leave the SM.
constexpr int RESCORE_WARPS = 4;

// SHARDED: item-sharded runs leave a handful of candidates per user and rank, so the pass is a chain of memory
// round trips per user, not arithmetic: twice the resident warps (registers capped at 64, the 32 partial sums of
// lane_exact_dot spill to L1) beat the single-GPU register budget there.
template <bool SHARDED>
__global__ void __launch_bounds__(RESCORE_WARPS * 32, SHARDED ? 8 : 4)
score_rescore_kernel(const ExactArgs e, int64_t user0, int64_t n_users_chunk, const uint32_t *__restrict__ cand_cnt,
                     const int2 *__restrict__ cand, const float *__restrict__ eps, int kc, int K,
                     int32_t *__restrict__ out_items, double *__restrict__ out_scores, int32_t *__restrict__ fail_list,
                     uint32_t *__restrict__ n_fail, unsigned long long *__restrict__ n_cand_total) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double *ex = reinterpret_cast<double *>(rs_smem) + (size_t)wid * kc;
  double *au = reinterpret_cast<double *>(rs_smem) + (size_t)RESCORE_WARPS * kc + (size_t)wid * e.k;
  int32_t *it = reinterpret_cast<int32_t *>(rs_smem + (size_t)RESCORE_WARPS * (kc + e.k) * 8) + (size_t)wid * 2 * kc;
  uint32_t *key = reinterpret_cast<uint32_t *>(it + kc);
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long seen = 0ull;        // candidates of this warp's users: ONE atomic per warp at the end (a per-user
                                         // atomic on a single address serialised all 32,768 users of a pass at L2)
  // A user's count, first 32 candidates and factor row are requested one loop iteration ahead (all independent
  // loads), so that the chain per user is candidates' item rows -> rank -> store instead of four round trips.
  struct Pre {
    uint32_t cnt;
    int2 first;
    double a[4];        // factors lane, lane + 32, ... (k <= 128 on this path)
  };
  auto fetch = [&](int64_t row) {
    Pre p;
    p.cnt = 0u;
    p.first = make_int2(-1, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) p.a[j] = 0.0;
    if (row < n_users_chunk) {
      p.cnt = cand_cnt[row];
      if (lane < kc) p.first = cand[row * kc + lane];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (lane + 32 * j < e.k) p.a[j] = e.A[(user0 + row) * e.k + lane + 32 * j];
    }
    return p;
  };
  Pre nxt = fetch(gw);
  for (int64_t row = gw; row < n_users_chunk; row += nw) {
    const int64_t u = user0 + row;
    const Pre cur = nxt;
    nxt = fetch(row + nw);
    const uint32_t cnt = cur.cnt;
    const int2 first = cur.first;
    seen += cnt;
    if (cnt > (uint32_t)kc) {            // more candidates than the buffer holds: rank this user exactly
      if (lane == 0) fail_list[atomicAdd(n_fail, 1u)] = (int32_t)u;
      continue;
    }
    int n_cand = (int)cnt;
    if (n_cand == 0) {                   // nothing of this user's top-K can live in this range (sharded runs)
      for (int r = lane; r < K; r += 32) {
        out_items[u * K + r] = -1;
        out_scores[u * K + r] = -INFINITY;
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (lane + 32 * j < e.k) au[lane + 32 * j] = cur.a[j];
    for (int c = lane; c < n_cand; c += 32) {
      const int2 v = c == lane ? first : cand[row * kc + c];
      const uint32_t bits = (uint32_t)v.y;
      it[c] = v.x;
      key[c] = bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);      // order-preserving uint key
    }
    __syncwarp();
    if (n_cand > K) {
      // The K-th largest approximate score over ALL collected items is a tighter bound than the sampled
      // threshold: only candidates within 2 eps of it can reach the exact top-K. Bisection on the key bits.
      uint32_t thr = 0u;
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand_thr = thr | (1u << bit);
        int above = 0;
        for (int c = lane; c < n_cand; c += 32) above += key[c] >= cand_thr;
        above = __reduce_add_sync(FULL, above);
        if (above >= K) thr = cand_thr;
      }
      const uint32_t bits = (thr & 0x80000000u) ? (thr ^ 0x80000000u) : ~thr;
      const float cut = __double2float_rd((double)__uint_as_float(bits) - 2.0 * (double)eps[row]);
      const uint32_t cbits = __float_as_uint(cut);
      const uint32_t cut_key = cbits ^ ((cbits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
      int kept = 0;                      // compact the survivors to the front of it[]
      for (int c0 = 0; c0 < n_cand; c0 += 32) {
        const int c = c0 + lane;
        const bool keep = c < n_cand && key[c] >= cut_key;
        const int item = c < n_cand ? it[c] : -1;
        const uint32_t m = __ballot_sync(FULL, keep);
        __syncwarp();
        if (keep) it[kept + __popc(m & ((1u << lane) - 1u))] = item;
        kept += __popc(m);
        __syncwarp();
      }
      n_cand = kept;
    }
    const double alpha_u = e.alpha ? e.alpha[u] : 0.0;
    for (int c = lane; c < n_cand; c += 32) {        // a candidate per lane
      const int item = it[c];
      // outer association of the reference: (((dot + alpha_u) + beta_i) + bias), src/mf.py:165-170
      ex[c] = ((lane_exact_dot(au, e.C + (size_t)item * e.k, e.k) + alpha_u) + (e.beta ? e.beta[item] : 0.0)) + e.bias;
    }
    __syncwarp();
    // rank by counting: the canonical order is a strict total order (items are distinct), so the number of
    // candidates ranking before c is c's position; positions below K are the answer
    for (int c = lane; c < n_cand; c += 32) {
      Best mine;
      mine.s = ex[c];
      mine.item = it[c];
      int rank = 0;
      for (int d = 0; d < n_cand; ++d) {
        Best other;
        other.s = ex[d];
        other.item = it[d];
        rank += ranks_before(other, mine) ? 1 : 0;
      }
      if (rank < K) {
        out_items[u * K + rank] = mine.item;
        out_scores[u * K + rank] = mine.s;
      }
    }
    for (int r = n_cand + lane; r < K; r += 32) {    // fewer than K items in the range: pad the rest
      out_items[u * K + r] = -1;
      out_scores[u * K + r] = -INFINITY;
    }
    __syncwarp();
  }
  if (lane == 0 && seen) atomicAdd(n_cand_total, seen);
}

// one CTA per user: exact top-K over the whole catalog range (fallback and exact mode). The float64 scores of
// the range are computed once (a warp per item) into the CTA's scratch row, then K selection rounds read them back.
constexpr int EXACT_THREADS = 256;
__global__ void __launch_bounds__(EXACT_THREADS)
score_exact_kernel(const ExactArgs e, const int32_t *__restrict__ users, const uint32_t *__restrict__ n_users_dev,
                   int64_t n_users_host, int K, double *__restrict__ scratch, int32_t *__restrict__ out_items,
                   double *__restrict__ out_scores) {
  __shared__ Best wbest[EXACT_THREADS / 32];
  const int64_t n_u = n_users_dev ? (int64_t)*n_users_dev : n_users_host;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t n_range = e.item_end - e.item_begin;
  double *sc = scratch + (size_t)blockIdx.x * n_range;
  for (int64_t ui = blockIdx.x; ui < n_u; ui += gridDim.x) {
    const int64_t u = users ? users[ui] : ui;
    for (int64_t i = wid; i < n_range; i += EXACT_THREADS / 32) {
      const double v = warp_exact_score(e, u, e.item_begin + i, lane);
      if (lane == 0) sc[i] = v;
    }
    __syncthreads();
    Best last;
    last.s = 0.0;
    last.item = -2;
    for (int r = 0; r < K; ++r) {
      Best mine;
      mine.s = 0.0;
      mine.item = -1;
      for (int64_t i = threadIdx.x; i < n_range; i += EXACT_THREADS) {
        Best b;
        b.s = sc[i];
        b.item = (int)(e.item_begin + i);
        const bool below = last.item == -2 || ranks_before(last, b);
        if (below && ranks_before(b, mine)) mine = b;
      }
      mine = warp_best(mine);
      if (lane == 0) wbest[wid] = mine;
      __syncthreads();
      Best best = wbest[0];
      for (int w = 1; w < EXACT_THREADS / 32; ++w)
        if (ranks_before(wbest[w], best)) best = wbest[w];
      __syncthreads();
      last = best;
      if (threadIdx.x == 0) {
        out_items[u * K + r] = last.item;
        out_scores[u * K + r] = last.item >= 0 ? last.s : -INFINITY;
      }
      if (last.item < 0) {   // fewer than K items in the range: pad the rest
        for (int rr = r + 1 + threadIdx.x; rr < K; rr += EXACT_THREADS) {
          out_items[u * K + rr] = -1;
          out_scores[u * K + rr] = -INFINITY;
        }
        break;
      }
    }
    __syncthreads();
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


// ---- item-sharded runs: exchange over NVLink peer memory ---------------------------------
// Every rank ranks a slice of the catalog for ALL users. Two things cross the GPUs, both through a cudaMalloc'ed
// region that every rank maps with CUDA IPC (no NCCL, no staging):
//   (1) after the sampled pass, each rank's K largest group maxima per user. The K-th largest of their union is a
//       bound on the user's K-th best score over the WHOLE catalog, so every rank collects against the global
//       threshold: the candidates of a user add up to ~K x stride over all ranks instead of per rank, and the
//       exact re-scoring shrinks with the number of ranks like the tensor passes do;
//   (2) the per-rank exact top-K lists, which the owner of a user range (rank r owns users [r U/G, (r+1) U/G))
//       merges with a K-way merge by rank counting.
// A barrier is its own tiny kernel in stream order: one thread per peer stores this rank's step number into the
// peer's flag word (st.release.sys) and polls its own copy of the peer's (ld.acquire.sys). Step numbers only grow,
// waits are bounded by a wall-clock limit and raise a sticky status word instead of hanging.
constexpr int XCH_MAX_WORLD = 8;
constexpr unsigned long long XCH_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

struct XchPeers {
  const unsigned char *base[XCH_MAX_WORLD];
};
struct ListPtrs {           // the sorted (item, score) lists to merge, [n_users][K] each
  const int32_t *items[XCH_MAX_WORLD];
  const double *scores[XCH_MAX_WORLD];
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void topk_xch_barrier_kernel(const XchPeers peers, size_t flags_off, int rank, int world, uint32_t seq,
                                        uint32_t *__restrict__ status) {
  const int q = threadIdx.x;
  if (q >= world) return;
  __threadfence_system();
  uint32_t *theirs = reinterpret_cast<uint32_t *>(const_cast<unsigned char *>(peers.base[q]) + flags_off) + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(seq) : "memory");
  const uint32_t *mine = reinterpret_cast<const uint32_t *>(peers.base[rank] + flags_off) + q;
  const unsigned long long t0 = global_ns();
  while (true) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (v >= seq) break;
    if (global_ns() - t0 > XCH_TIMEOUT_NS) {
      atomicCAS(status, 0u, 1u + (uint32_t)q);
      break;
    }
    __nanosleep(64);
  }
}

// One warp per user: K-th largest of the world x K group-maximum keys the ranks exported -> collect threshold.
// A lane keeps XCH_MAX_WORLD x KJ keys in registers (KJ = ceil(K / 32) per peer; the peer loop is uniform, so the
// peers' base pointers stay in uniform registers) and issues every peer load before the first use.
template <int KJ>
__global__ void __launch_bounds__(SELECT_WARPS * 32)
score_global_threshold_kernel(const XchPeers peers, size_t gtop_off, int world, int K, int64_t user0,
                              int64_t n_users_chunk, const float *__restrict__ eps, float *__restrict__ tau) {
  constexpr int NJ = XCH_MAX_WORLD * KJ;
  const int lane = threadIdx.x & 31;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = gw; row < n_users_chunk; row += nw) {
    const int64_t u = user0 + row;
    uint32_t key[NJ];
#pragma unroll
    for (int q = 0; q < XCH_MAX_WORLD; ++q) {
      const uint32_t *src = q < world ? reinterpret_cast<const uint32_t *>(peers.base[q] + gtop_off) + u * K : nullptr;
#pragma unroll
      for (int r = 0; r < KJ; ++r) {
        const int jj = lane + 32 * r;
        key[q * KJ + r] = (src && jj < K) ? __ldcg(src + jj) : 0u;
      }
    }
    const uint32_t thr = warp_kth_largest<NJ>(key, K, lane);
    const float t0 = key_to_float(thr);
    if (lane == 0) tau[row] = t0 > -INFINITY ? __double2float_rd((double)t0 - 2.0 * (double)eps[row]) : -INFINITY;
  }
}

// K-way merge of the ranks' sorted top-K lists for the users this rank owns. One warp per user: the world x K
// (item, score) pairs are staged in shared memory (coalesced peer loads), then every pair finds its global
// position = its position in its own list + the number of pairs of every other list that rank before it (binary
// search; the canonical order is strict because items are distinct). Positions below K are the answer.
constexpr int MERGE_WARPS = 8;
__global__ void __launch_bounds__(MERGE_WARPS * 32)
topk_xch_merge_kernel(const ListPtrs lists, int world, int K, int64_t user_begin,
                      int64_t user_end, int32_t *__restrict__ out_items, double *__restrict__ out_scores) {
  extern __shared__ __align__(16) unsigned char mg_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n = world * K;
  double *ss = reinterpret_cast<double *>(mg_smem) + (size_t)wid * n;
  int32_t *si = reinterpret_cast<int32_t *>(mg_smem + (size_t)MERGE_WARPS * n * 8) + (size_t)wid * n;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t u = user_begin + gw; u < user_end; u += nw) {
    int valid = 0;
#pragma unroll
    for (int q = 0; q < XCH_MAX_WORLD; ++q) {           // uniform loop: the list pointers stay in uniform registers
      if (q >= world) break;
      const int32_t *pi = lists.items[q] + u * K;
      const double *ps = lists.scores[q] + u * K;
      int32_t it[4];
      double sv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {                     // K <= 120: at most four entries per lane and list, all in flight
        const int j = lane + 32 * r;
        it[r] = j < K ? __ldcg(pi + j) : -1;
        sv[r] = j < K ? __ldcg(ps + j) : 0.0;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = lane + 32 * r;
        if (j < K) {
          si[q * K + j] = it[r];
          ss[q * K + j] = sv[r];
          valid += it[r] >= 0;
        }
      }
    }
    valid = __reduce_add_sync(FULL, valid);
    __syncwarp();
    int32_t *oi = out_items + (u - user_begin) * K;
    double *os = out_scores + (u - user_begin) * K;
    for (int i = lane; i < n; i += 32) {
      Best mine;
      mine.item = si[i];
      mine.s = ss[i];
      if (mine.item < 0) continue;
      const int q = i / K;
      int rank = i - q * K;
      for (int l = 0; l < world; ++l) {
        if (l == q) continue;
        int lo = 0, hi = K;                   // first position of list l that does NOT rank before mine
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          Best other;
          other.item = si[l * K + mid];
          other.s = ss[l * K + mid];
          if (ranks_before(other, mine)) lo = mid + 1; else hi = mid;
        }
        rank += lo;
      }
      if (rank < K) {
        oi[rank] = mine.item;
        os[rank] = mine.s;
      }
    }
    for (int r = valid + lane; r < K; r += 32) {
      oi[r] = -1;
      os[r] = -INFINITY;
    }
    __syncwarp();
  }
}

__global__ void fill_lists_kernel(int32_t *__restrict__ items, double *__restrict__ scores, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    items[i] = -1;
    scores[i] = -INFINITY;
  }
}
__global__ void fill_u32_kernel(uint32_t *__restrict__ p, uint32_t v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
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

// CTAs of the exact kernel: each keeps the float64 scores of the whole range, within a 512 MiB scratch budget
int exact_grid(const rfm_ctx *ctx, int64_t n_users, int64_t n_range) {
  const int64_t by_scratch = std::max<int64_t>(1, ((int64_t)512 << 20) / (8 * n_range));
  return (int)std::min<int64_t>(std::min<int64_t>(n_users, (int64_t)ctx->sm_count * 4), by_scratch);
}

constexpr int MAX_GROUPS = 8192;     // group maxima per user the threshold kernel stages in shared memory

// Split the visited tiles of a pass over blockIdx.y so that the grid fills whole waves of the SMs. Every CTA
// pays about one and a half tiles of prologue (TMEM allocation, the A tile, pipeline fill).
int pick_splits(int n_user_blocks, int n_visit, int sm_count) {
  int best = 1;
  double best_score = -1.0;
  for (int s = 1; s <= std::min(n_visit, 256); ++s) {
    const int per = (n_visit + s - 1) / s;
    const int s_eff = (n_visit + per - 1) / per;
    const int64_t ctas = (int64_t)n_user_blocks * s_eff;
    const int64_t waves = (ctas + sm_count - 1) / sm_count;
    const double score = (double)ctas / (double)(waves * sm_count) * (per / (per + 1.5));
    if (score > best_score + 1e-9) {
      best_score = score;
      best = s_eff;
    }
  }
  return best;
}

template <int KB, int PASS, int GCOLS>
int launch_pass_as(rfm_ctx *ctx, dim3 grid, const CUtensorMap &tmap_a, const CUtensorMap &tmap_c, const PassArgs &pa) {
  auto kernel = score_pass_kernel<KB, PASS, GCOLS>;
  RFM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PassSmem<KB>::BYTES));
  if (PASS == 0) {
    auto score_sample = kernel;
    RFM_LAUNCH(ctx, score_sample, grid, SCORE_THREADS, PassSmem<KB>::BYTES, tmap_a, tmap_c, pa);
  } else {
    auto score_collect = kernel;
    RFM_LAUNCH(ctx, score_collect, grid, SCORE_THREADS, PassSmem<KB>::BYTES, tmap_a, tmap_c, pa);
  }
  return RFM_OK;
}
template <int KB>
int launch_pass_kb(rfm_ctx *ctx, int pass, int gcols, dim3 grid, const CUtensorMap &tmap_a, const CUtensorMap &tmap_c,
                   const PassArgs &pa) {
  if (pass == 1) return launch_pass_as<KB, 1, 64>(ctx, grid, tmap_a, tmap_c, pa);
  switch (gcols) {
    case 64: return launch_pass_as<KB, 0, 64>(ctx, grid, tmap_a, tmap_c, pa);
    case 32: return launch_pass_as<KB, 0, 32>(ctx, grid, tmap_a, tmap_c, pa);
    case 16: return launch_pass_as<KB, 0, 16>(ctx, grid, tmap_a, tmap_c, pa);
    default: return launch_pass_as<KB, 0, 8>(ctx, grid, tmap_a, tmap_c, pa);
  }
}
int launch_pass(rfm_ctx *ctx, int kb, int pass, int gcols, dim3 grid, const CUtensorMap &tmap_a,
                const CUtensorMap &tmap_c, const PassArgs &pa) {
  return kb == 1 ? launch_pass_kb<1>(ctx, pass, gcols, grid, tmap_a, tmap_c, pa)
                 : launch_pass_kb<2>(ctx, pass, gcols, grid, tmap_a, tmap_c, pa);
}

template <int NJ>
int launch_threshold_as(rfm_ctx *ctx, int grid, const float *gmax, int64_t n_groups, int fold, int64_t n_rows,
                        int64_t n_users_chunk, int K, const double *a_norm, const double *a_res, const double *c_norm_max,
                        const double *c_res_max, const double *beta_abs_max, int kpad, float *tau, float *eps_out,
                        uint32_t *cand_cnt, uint32_t *gtop) {
  auto score_threshold = score_threshold_kernel<NJ>;
  const size_t smem = NJ > 0 ? 0 : (size_t)SELECT_WARPS * n_groups * 4;
  if (smem) RFM_CUDA(cudaFuncSetAttribute(score_threshold, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RFM_LAUNCH(ctx, score_threshold, grid, SELECT_WARPS * 32, smem, gmax, n_groups, fold, n_rows, n_users_chunk, K, a_norm,
             a_res, c_norm_max, c_res_max, beta_abs_max, kpad, tau, eps_out, cand_cnt, gtop);
  return RFM_OK;
}
template <typename... Args>
int launch_threshold(rfm_ctx *ctx, int grid, const float *gmax, int64_t n_groups, Args... args) {
  const int64_t per_lane = (n_groups + 31) / 32;
  if (per_lane <= 8) return launch_threshold_as<8>(ctx, grid, gmax, n_groups, args...);
  if (per_lane <= 16) return launch_threshold_as<16>(ctx, grid, gmax, n_groups, args...);
  if (per_lane <= 32) return launch_threshold_as<32>(ctx, grid, gmax, n_groups, args...);
  if (per_lane <= 64) return launch_threshold_as<64>(ctx, grid, gmax, n_groups, args...);
  return launch_threshold_as<0>(ctx, grid, gmax, n_groups, args...);
}

}  // namespace

struct rfm_topk {
  rfm_ctx *ctx = nullptr;
  int64_t n_users = 0, n_items = 0, n_users_pad = 0, n_items_pad = 0;
  int k = 0, kpad = 0, kb = 0;
  double bias = 0.0;
  bool has_alpha = false, has_beta = false, ready = false;
  DevBuf<double> A, C, alpha, beta, a_norm, a_res, c_norm, c_res, c_norm_max, c_res_max, beta_abs_max;
  DevBuf<__nv_bfloat16> A16, C16, beta16;
  DevBuf<float> gmax, tau, eps;
  DevBuf<int32_t> out_items, fail_list;
  DevBuf<int2> cand;
  PinnedBuf<int32_t> stage_items;
  PinnedBuf<double> stage_scores;
  DevBuf<double> out_scores, exact_scratch;
  DevBuf<uint32_t> n_fail, cand_cnt;
  DevBuf<unsigned long long> n_cand;
  CUtensorMap tmap_a, tmap_c;
  // item-sharded exchange (rfm_topk_dp_*): [items int32 [U][k_cap] | scores f64 [U][k_cap] | gtop u32 [U][k_cap] | flags]
  unsigned char *xchg = nullptr;
  size_t x_items_off = 0, x_scores_off = 0, x_gtop_off = 0, x_flags_off = 0, x_bytes = 0;
  int x_kcap = 0, x_rank = -1, x_world = 0;
  uint32_t x_seq = 0;
  unsigned char *x_peer[XCH_MAX_WORLD] = {nullptr};
  DevBuf<uint32_t> x_status;
  int64_t own_begin = 0, own_end = 0;       // users whose merged lists the last sharded run left in out_items/out_scores
  int64_t result_rows = 0;                  // rows of out_items / out_scores that hold the last result
  PinnedBuf<unsigned long long> counters;   // [users ranked exactly (low 32 bits), candidates collected] of the last run
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
    RFM_TRY(t->a_res.alloc(n_users));
    RFM_TRY(t->c_res.alloc(n_items));
    RFM_TRY(t->c_res_max.alloc(1));
    RFM_TRY(t->c_norm_max.alloc(1));
    RFM_TRY(t->beta_abs_max.alloc(1));
    RFM_TRY(t->n_fail.alloc(1));
    RFM_TRY(t->fail_list.alloc(n_users));
    if (t->kb <= MAX_KB) {
      RFM_TRY(t->A16.alloc((size_t)t->n_users_pad * t->kpad));
      RFM_TRY(t->C16.alloc((size_t)t->n_items_pad * t->kpad));
      RFM_TRY(t->beta16.alloc((size_t)t->n_items_pad * 16));
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
    for (int q = 0; q < t->x_world; ++q)
      if (q != t->x_rank && t->x_peer[q]) cudaIpcCloseMemHandle(t->x_peer[q]);
    if (t->xchg) cudaFree(t->xchg);
    cudaGetLastError();
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
               t->a_norm.p, t->a_res.p);
    RFM_LAUNCH(ctx, to_bf16_kernel, g, 256, 0, t->C.p, t->n_items, t->k, t->n_items_pad, t->kpad, t->C16.p,
               t->c_norm.p, t->c_res.p);
    RFM_LAUNCH(ctx, beta_operand_kernel, g, 256, 0, beta ? t->beta.p : (const double *)nullptr, t->n_items,
               t->n_items_pad, t->beta16.p);
    RFM_LAUNCH(ctx, max_reduce_kernel, 1, 1024, 0, t->c_norm.p, t->n_items, t->c_norm_max.p);
    RFM_LAUNCH(ctx, max_reduce_kernel, 1, 1024, 0, t->c_res.p, t->n_items, t->c_res_max.p);
    RFM_LAUNCH(ctx, max_reduce_kernel, 1, 1024, 0, t->beta.p, t->n_items, t->beta_abs_max.p);
    RFM_TRY(make_tmap(&t->tmap_a, t->A16.p, t->n_users_pad, t->kpad, BM));
    RFM_TRY(make_tmap(&t->tmap_c, t->C16.p, t->n_items_pad, t->kpad, BN));
  }
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  t->ready = true;
  return RFM_OK;
}

}  // extern "C" (pause)
namespace {

struct ShardPlan {          // how the item-sharded run differs from the single-GPU one
  bool on = false;
  int tiles_max = 0;        // item tiles of the largest shard: every rank cuts the users into the same chunks
};

XchPeers xch_peers(const rfm_topk *t) {
  XchPeers p;
  for (int q = 0; q < XCH_MAX_WORLD; ++q) p.base[q] = q < t->x_world ? t->x_peer[q] : nullptr;
  return p;
}

int xch_barrier(rfm_topk *t) {
  rfm_ctx *ctx = t->ctx;
  ++t->x_seq;
  RFM_LAUNCH(ctx, topk_xch_barrier_kernel, 1, 32, 0, xch_peers(t), t->x_flags_off, t->x_rank, t->x_world, t->x_seq,
             t->x_status.p);
  return RFM_OK;
}

// tile range [begin, end) of a rank's slice of the catalog (the rule of rfm_b200.dist.item_shard)
void shard_tiles(int64_t n_items, int world, int rank, int64_t *begin, int64_t *end) {
  const int64_t n_tiles = (n_items + BN - 1) / BN, base = n_tiles / world, extra = n_tiles % world;
  *begin = rank * base + std::min<int64_t>(rank, extra);
  *end = *begin + base + (rank < extra ? 1 : 0);
}

// The pipeline of one call over the catalog range [item_begin, item_end); the per-user lists go to
// dst_items / dst_scores ([n_users][K], device). plan.on: thresholds are global (exchange (1) above) and the range
// may be empty (a rank beyond the last tile still takes part in every barrier).
int topk_run_core(rfm_topk *t, int K, int mode, int64_t item_begin, int64_t item_end, const ShardPlan &plan,
                  int32_t *dst_items, double *dst_scores, bool *tensor_path_out, int64_t *n_failed_out,
                  int64_t *n_candidates_out, int64_t *stride_out) {
  rfm_ctx *ctx = t->ctx;
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
  int64_t n_failed = 0, n_candidates = 0, sample_stride = 0;
  const bool empty = item_end <= item_begin;
  // the tensor-core path needs a tile-aligned range (item shards are cut at multiples of 256; the catalog's own
  // end is padded with beta = -inf) and k <= 128
  const bool tensor_path = mode == 0 && t->kb <= MAX_KB && (empty || (item_begin % BN == 0 &&
                           (item_end % BN == 0 || item_end == t->n_items)));
  if (tensor_path) {
    const int tile_begin = (int)(item_begin / BN);
    const int n_item_tiles = empty ? 0 : (int)((item_end + BN - 1) / BN) - tile_begin;
    const int plan_tiles = plan.on ? plan.tiles_max : n_item_tiles;       // identical on every rank
    const int n_user_blocks = (int)(t->n_users_pad / BM);
    // Pass 1 scores a sample of the catalog, every stride-th tile. A sparser sample is cheaper but gives a
    // lower threshold: about K * stride items per user reach it (negative binomial, sd stride * sqrt(K (1 - 1/stride))).
    int stride = K <= 16 ? 4 : (K <= 64 ? 2 : 1);
    if (const char *env = getenv("RFM_SCORE_STRIDE")) stride = std::max(1, atoi(env));
    // The threshold is the K-th largest group maximum, so the sample must hold many more groups than K: small
    // catalogs are sampled densely and in narrower groups (128 -> 8 items). Sharded runs size the sample from the
    // largest shard so that every rank samples alike (the union over the ranks then holds world x as many groups).
    int gcols = MAX_GCOLS;
    auto groups_of = [&](int tiles, int st, int g) { return (int64_t)((tiles + st - 1) / st) * (BN / g); };
    // (a sharded run's threshold comes from the union of the ranks' samples: each needs 1 / world of the groups)
    const int64_t want_groups = plan.on ? std::max<int64_t>(64, (8 * (int64_t)K + 64 + t->x_world - 1) / t->x_world)
                                        : 8 * (int64_t)K + 64;
    while (stride > 1 && groups_of(plan_tiles, stride, gcols) < want_groups) stride /= 2;
    while (gcols > 8 && groups_of(plan_tiles, stride, gcols) < want_groups) gcols /= 2;
    while (groups_of(plan_tiles, stride, gcols) > MAX_GROUPS) ++stride;
    const int n_sample = (n_item_tiles + stride - 1) / stride;
    const int64_t n_groups = groups_of(n_item_tiles, stride, gcols);
    const int64_t plan_groups = std::max<int64_t>(1, groups_of(plan_tiles, stride, gcols));
    // About c K stride items per user reach the threshold (negative binomial in the sampling, sd
    // stride sqrt(K (1 - 1/stride)); c = -ln(1 - f) / f corrects for top items sharing a group, f = K / groups).
    // the threshold kernel merges `fold` adjacent stored groups on load when the sample holds more groups than the
    // threshold needs: fewer keys per user to select from
    int fold = 1;
    for (int cand = 4; cand >= 2 && fold == 1; cand /= 2)
      if (n_groups % cand == 0 && n_groups / cand >= want_groups) fold = cand;
    const double f = std::min(0.5, (double)K * fold / (double)plan_groups);
    const double cf = -std::log1p(-f) / f;
    const double sd = stride * std::sqrt((double)K * (1.0 - 1.0 / stride));
    // The 2 eps safety margin lowers the threshold further (how much depends on the score distribution), so the
    // buffer is four times the expected count: an overflow costs an exact re-rank of that user.
    int kc = (int)std::ceil(cf * (4.0 * K * stride + 6.0 * sd)) + 32;
    kc = std::max<int64_t>(kc, std::min<int64_t>((int64_t)plan_tiles * BN, 256));
    kc = (kc + 7) / 8 * 8;
    // users are processed in chunks so that the group maxima of a chunk stay within a fixed scratch budget
    int64_t scratch_bytes = (int64_t)1 << 30;
    if (const char *env = getenv("RFM_SCORE_SCRATCH_MB")) scratch_bytes = std::max<int64_t>(1, atoll(env)) << 20;
    const int64_t scratch_rows = std::max<int64_t>(BM, scratch_bytes / (plan_groups * 4) / BM * BM);
    const int chunk_blocks = (int)std::min<int64_t>(n_user_blocks, scratch_rows / BM);
    const size_t chunk_rows = (size_t)chunk_blocks * BM;
    RFM_TRY(t->gmax.ensure(chunk_rows * std::max<int64_t>(1, n_groups)));
    RFM_TRY(t->tau.ensure(chunk_rows));
    RFM_TRY(t->cand_cnt.ensure(chunk_rows));
    RFM_TRY(t->cand.ensure(chunk_rows * kc));
    RFM_TRY(t->eps.ensure(chunk_rows));
    RFM_TRY(t->n_cand.ensure(1));
    RFM_CUDA(cudaMemsetAsync(t->n_fail.p, 0, sizeof(uint32_t), ctx->stream));
    RFM_CUDA(cudaMemsetAsync(t->n_cand.p, 0, sizeof(unsigned long long), ctx->stream));
    const size_t rs_smem = (size_t)RESCORE_WARPS * ((size_t)kc * 16 + (size_t)t->k * 8);
    RFM_REQUIRE(rs_smem <= 200 * 1024, "rfm_topk_run: %d candidates per user do not fit the re-scoring kernel", kc);
    RFM_CUDA(cudaFuncSetAttribute(score_rescore_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem));
    RFM_CUDA(cudaFuncSetAttribute(score_rescore_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem));
    uint32_t *gtop = plan.on ? reinterpret_cast<uint32_t *>(t->xchg + t->x_gtop_off) : nullptr;
    if (plan.on && empty) {     // nothing to rank here: no group maxima, empty lists
      RFM_LAUNCH(ctx, fill_u32_kernel, ctx->sm_count * 4, 256, 0, gtop, 0x007FFFFFu, (int64_t)t->n_users * K);
      RFM_LAUNCH(ctx, fill_lists_kernel, ctx->sm_count * 4, 256, 0, dst_items, dst_scores, (int64_t)t->n_users * K);
    }
    for (int ub0 = 0; ub0 < n_user_blocks; ub0 += chunk_blocks) {
      const int nb = std::min(chunk_blocks, n_user_blocks - ub0);
      const int64_t user0 = (int64_t)ub0 * BM;
      const int64_t users_here = std::min<int64_t>((int64_t)nb * BM, t->n_users - user0);
      if (empty) {
        RFM_TRY(xch_barrier(t));      // the peers' thresholds wait for this rank's (empty) export
        continue;
      }
      PassArgs pa;
      pa.beta16 = t->beta16.p;
      pa.tile_begin = tile_begin;
      pa.item_end = (int)item_end;
      pa.user_block0 = ub0;
      pa.n_user_blocks = nb;
      // every CTA holds up to ub blocks of 128 users, as many as still leave two waves of CTAs
      int ub = t->kb == 1 ? PassSmem<1>::UB : PassSmem<2>::UB;
      while (ub > 1 && (int64_t)((nb + ub - 1) / ub) * n_sample < 2 * (int64_t)ctx->sm_count) ub /= 2;
      pa.ub = ub;
      const int n_ctas_x = (nb + ub - 1) / ub;
      pa.n_groups = n_groups;
      pa.gmax = t->gmax.p;
      pa.tau = t->tau.p;
      pa.kc = kc;
      pa.cand_cnt = t->cand_cnt.p;
      pa.cand = t->cand.p;
      // pass 1: group maxima of the sample
      pa.tile_stride = stride;
      pa.n_visit = n_sample;
      int n_splits = pick_splits(n_ctas_x, n_sample, ctx->sm_count);
      pa.tiles_per_split = (n_sample + n_splits - 1) / n_splits;
      n_splits = (n_sample + pa.tiles_per_split - 1) / pa.tiles_per_split;
      RFM_TRY(launch_pass(ctx, t->kb, 0, gcols, dim3(n_ctas_x, n_splits), t->tmap_a, t->tmap_c, pa));
      const int tgrid = (int)std::min<int64_t>(((int64_t)nb * BM + SELECT_WARPS - 1) / SELECT_WARPS,
                                               (int64_t)ctx->sm_count * 16);
      RFM_TRY(launch_threshold(ctx, tgrid, t->gmax.p, n_groups / fold, fold, (int64_t)nb * BM, users_here, (int)K,
                               t->a_norm.p + user0, t->a_res.p + user0, t->c_norm_max.p, t->c_res_max.p,
                               t->beta_abs_max.p, t->kpad, t->tau.p, t->eps.p, t->cand_cnt.p,
                               gtop ? gtop + user0 * K : (uint32_t *)nullptr));
      if (plan.on) {
        // every rank's K largest group maxima are in place -> the global K-th largest replaces the local bound
        RFM_TRY(xch_barrier(t));
        const int kj = (K + 31) / 32;
        const int ggrid = (int)std::min<int64_t>((users_here + SELECT_WARPS - 1) / SELECT_WARPS,
                                                 (int64_t)ctx->sm_count * 16);
#define RFM_GLOBAL_THRESHOLD(KJ)                                                                             \
  do {                                                                                                       \
    auto score_global_threshold = score_global_threshold_kernel<KJ>;                                         \
    RFM_LAUNCH(ctx, score_global_threshold, ggrid, SELECT_WARPS * 32, 0, xch_peers(t), t->x_gtop_off,        \
               t->x_world, (int)K, user0, users_here, t->eps.p, t->tau.p);                                   \
  } while (0)
        if (kj <= 1) RFM_GLOBAL_THRESHOLD(1);
        else if (kj == 2) RFM_GLOBAL_THRESHOLD(2);
        else if (kj == 3) RFM_GLOBAL_THRESHOLD(3);
        else RFM_GLOBAL_THRESHOLD(4);            // K <= 120
#undef RFM_GLOBAL_THRESHOLD
      }
      // pass 2: collect every item that reaches the threshold
      pa.tile_stride = 1;
      pa.n_visit = n_item_tiles;
      n_splits = pick_splits(n_ctas_x, n_item_tiles, ctx->sm_count);
      pa.tiles_per_split = (n_item_tiles + n_splits - 1) / n_splits;
      n_splits = (n_item_tiles + pa.tiles_per_split - 1) / pa.tiles_per_split;
      RFM_TRY(launch_pass(ctx, t->kb, 1, 64, dim3(n_ctas_x, n_splits), t->tmap_a, t->tmap_c, pa));
      const int rs_per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (200 * 1024) / std::max<size_t>(rs_smem, 1)));
      const int rgrid = (int)std::min<int64_t>((users_here + RESCORE_WARPS - 1) / RESCORE_WARPS,
                                               (int64_t)ctx->sm_count * rs_per_sm);
      // (measured at 8 GPUs: the register-capped variant is SLOWER -- 0.146 vs 0.115 ms at top-9, 0.27 vs 0.12 ms at
      // top-100: the spilled partial sums cost more than the extra warps hide; kept behind an environment switch)
      if (plan.on && getenv("RFM_RESCORE_LOWREG") != nullptr) {
        auto score_rescore = score_rescore_kernel<true>;
        RFM_LAUNCH(ctx, score_rescore, rgrid, RESCORE_WARPS * 32, rs_smem, e, user0, users_here, t->cand_cnt.p,
                   t->cand.p, t->eps.p, kc, (int)K, dst_items, dst_scores, t->fail_list.p, t->n_fail.p, t->n_cand.p);
      } else {
        auto score_rescore_kernel_ = score_rescore_kernel<false>;
        RFM_LAUNCH(ctx, score_rescore_kernel_, rgrid, RESCORE_WARPS * 32, rs_smem, e, user0, users_here, t->cand_cnt.p,
                   t->cand.p, t->eps.p, kc, (int)K, dst_items, dst_scores, t->fail_list.p, t->n_fail.p, t->n_cand.p);
      }
    }
    if (!empty) {
      // users whose candidate buffer overflowed are ranked exactly against the whole catalog range
      const int fgrid = exact_grid(ctx, t->n_users, item_end - item_begin);
      RFM_TRY(t->exact_scratch.ensure((size_t)fgrid * (item_end - item_begin)));
      RFM_LAUNCH(ctx, score_exact_kernel, fgrid, EXACT_THREADS, 0, e, t->fail_list.p, t->n_fail.p, (int64_t)0, (int)K,
                 t->exact_scratch.p, dst_items, dst_scores);
    }
    // the two counters come back through page-locked memory (a copy into pageable memory would block the host
    // until the stream drains); sharded runs read them after the merge, single-GPU runs right here
    RFM_TRY(t->counters.ensure(2));
    RFM_CUDA(cudaMemcpyAsync(t->counters.p, t->n_fail.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    RFM_CUDA(cudaMemcpyAsync(t->counters.p + 1, t->n_cand.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             ctx->stream));
    if (!plan.on) {
      RFM_CUDA(cudaStreamSynchronize(ctx->stream));
      n_failed = (int64_t)(uint32_t)t->counters.p[0];
      n_candidates = (int64_t)t->counters.p[1];
    }
    sample_stride = stride;
  } else if (empty) {
    RFM_LAUNCH(ctx, fill_lists_kernel, ctx->sm_count * 4, 256, 0, dst_items, dst_scores, (int64_t)t->n_users * K);
  } else {
    const int fgrid = exact_grid(ctx, t->n_users, item_end - item_begin);
    RFM_TRY(t->exact_scratch.ensure((size_t)fgrid * (item_end - item_begin)));
    RFM_LAUNCH(ctx, score_exact_kernel, fgrid, EXACT_THREADS, 0, e, (const int32_t *)nullptr, (const uint32_t *)nullptr,
               t->n_users, (int)K, t->exact_scratch.p, dst_items, dst_scores);
  }
  *tensor_path_out = tensor_path;
  *n_failed_out = n_failed;
  *n_candidates_out = n_candidates;
  *stride_out = sample_stride;
  return RFM_OK;
}

// result rows [0, n_rows) of t->out_items / out_scores -> caller memory through the page-locked staging buffers
int topk_result_to_host(rfm_topk *t, int64_t n_rows, int K, int32_t *out_items, double *out_scores) {
  rfm_ctx *ctx = t->ctx;
  const size_t n_out = (size_t)n_rows * K;
  RFM_TRY(t->stage_items.ensure(n_out ? n_out : 1));
  RFM_TRY(t->stage_scores.ensure(n_out ? n_out : 1));
  RFM_CUDA(cudaMemcpyAsync(t->stage_items.p, t->out_items.p, n_out * 4, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(t->stage_scores.p, t->out_scores.p, n_out * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (out_items) {
    memcpy(out_items, t->stage_items.p, n_out * 4);
    memcpy(out_scores, t->stage_scores.p, n_out * 8);
  }
  return RFM_OK;
}

}  // namespace
extern "C" {

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
  bool tensor_path = false;
  int64_t n_failed = 0, n_candidates = 0, sample_stride = 0;
  RFM_TRY(topk_run_core(t, K, mode, item_begin, item_end, ShardPlan(), t->out_items.p, t->out_scores.p, &tensor_path,
                        &n_failed, &n_candidates, &sample_stride));
  t->result_rows = t->n_users;
  t->own_begin = 0;
  t->own_end = t->n_users;
  if (out_items) {   // NULL: the caller reads the result on the device (rfm_topk_result_ptr_dev)
    RFM_TRY(topk_result_to_host(t, t->n_users, K, out_items, out_scores));
  } else {
    RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (stats) {
    stats[0] = tensor_path ? 1 : 0;
    stats[1] = n_failed;
    stats[2] = n_candidates;
    stats[3] = sample_stride;
  }
  return RFM_OK;
}

int rfm_topk_dp_export(rfm_topk *t, int32_t k_cap, void *handle_out) {
  RFM_REQUIRE(t && handle_out, "rfm_topk_dp_export: NULL argument");
  RFM_REQUIRE(k_cap >= 1 && k_cap <= MAX_K, "rfm_topk_dp_export: k_cap=%d outside [1, %d]", k_cap, MAX_K);
  RFM_REQUIRE(!t->xchg, "rfm_topk_dp_export: already exported");
  static_assert(sizeof(cudaIpcMemHandle_t) == RFM_DP_HANDLE_BYTES, "IPC handle size");
  rfm_ctx *ctx = t->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t n = (size_t)t->n_users * k_cap;
  t->x_items_off = 0;
  t->x_scores_off = up(n * 4);
  t->x_gtop_off = t->x_scores_off + up(n * 8);
  t->x_flags_off = t->x_gtop_off + up(n * 4);
  t->x_bytes = t->x_flags_off + 256;
  t->x_kcap = k_cap;
  // cudaMalloc, not the stream-ordered pool: pool memory cannot be exported through legacy CUDA IPC
  RFM_CUDA(cudaMalloc(reinterpret_cast<void **>(&t->xchg), t->x_bytes));
  RFM_CUDA(cudaMemsetAsync(t->xchg, 0, t->x_bytes, ctx->stream));
  RFM_TRY(t->x_status.alloc(1));
  RFM_CUDA(cudaMemsetAsync(t->x_status.p, 0, sizeof(uint32_t), ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaIpcMemHandle_t h;
  RFM_CUDA(cudaIpcGetMemHandle(&h, t->xchg));
  memcpy(handle_out, &h, sizeof(h));
  return RFM_OK;
}

int rfm_topk_dp_connect(rfm_topk *t, int32_t rank, int32_t world, const void *all_handles) {
  RFM_REQUIRE(t && all_handles, "rfm_topk_dp_connect: NULL argument");
  RFM_REQUIRE(t->xchg, "rfm_topk_dp_connect: call rfm_topk_dp_export first");
  RFM_REQUIRE(world >= 1 && world <= XCH_MAX_WORLD && rank >= 0 && rank < world,
              "rfm_topk_dp_connect: rank %d / world %d out of range (at most %d ranks)", rank, world, XCH_MAX_WORLD);
  RFM_REQUIRE(t->x_world == 0, "rfm_topk_dp_connect: already connected");
  RFM_CUDA(cudaSetDevice(t->ctx->device));
  for (int q = 0; q < world; ++q) {
    if (q == rank) {
      t->x_peer[q] = t->xchg;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const unsigned char *>(all_handles) + (size_t)q * sizeof(h), sizeof(h));
    void *p = nullptr;
    const cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) {
      cudaGetLastError();
      return fail(RFM_ERR_CUDA, "rfm_topk_dp_connect: cudaIpcOpenMemHandle(rank %d) failed: %s", q, cudaGetErrorString(err));
    }
    t->x_peer[q] = static_cast<unsigned char *>(p);
  }
  t->x_rank = rank;
  t->x_world = world;
  return RFM_OK;
}

int rfm_topk_run_sharded(rfm_topk *t, int32_t K, int32_t mode, int32_t *out_items, double *out_scores,
                         int64_t *user_range, int64_t *stats) {
  RFM_REQUIRE(t, "rfm_topk_run_sharded: NULL argument");
  RFM_REQUIRE((out_items == nullptr) == (out_scores == nullptr), "rfm_topk_run_sharded: out_items and out_scores go together");
  RFM_REQUIRE(t->ready, "rfm_topk_run_sharded: call rfm_topk_set_factors first");
  RFM_REQUIRE(t->x_world >= 1, "rfm_topk_run_sharded: call rfm_topk_dp_export / rfm_topk_dp_connect first");
  RFM_REQUIRE(K >= 1 && K <= t->x_kcap, "rfm_topk_run_sharded: K=%d outside [1, %d] (the exchange region's capacity)", K,
              t->x_kcap);
  RFM_REQUIRE(mode == 0 || mode == 1, "rfm_topk_run_sharded: mode must be 0 or 1");
  rfm_ctx *ctx = t->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  const int world = t->x_world, rank = t->x_rank;
  int64_t tb = 0, te = 0, tiles_max = 0;
  for (int q = 0; q < world; ++q) {
    int64_t b, e2;
    shard_tiles(t->n_items, world, q, &b, &e2);
    tiles_max = std::max(tiles_max, e2 - b);
    if (q == rank) {
      tb = b;
      te = e2;
    }
  }
  const int64_t item_begin = std::min<int64_t>(tb * BN, t->n_items), item_end = std::min<int64_t>(te * BN, t->n_items);
  // users this rank owns after the merge (the rule of rfm_b200.dist.slice_bounds)
  const int64_t ubase = t->n_users / world, uextra = t->n_users % world;
  const int64_t own_begin = rank * ubase + std::min<int64_t>(rank, uextra);
  const int64_t own_end = own_begin + ubase + (rank < uextra ? 1 : 0);
  const int64_t n_own = own_end - own_begin;
  RFM_TRY(t->out_items.ensure((size_t)std::max<int64_t>(1, n_own) * K));
  RFM_TRY(t->out_scores.ensure((size_t)std::max<int64_t>(1, n_own) * K));
  ShardPlan plan;
  plan.on = true;
  plan.tiles_max = (int)tiles_max;
  bool tensor_path = false;
  int64_t n_failed = 0, n_candidates = 0, sample_stride = 0;
  int32_t *lists_i = reinterpret_cast<int32_t *>(t->xchg + t->x_items_off);
  double *lists_s = reinterpret_cast<double *>(t->xchg + t->x_scores_off);
  RFM_TRY(topk_run_core(t, K, mode, item_begin, item_end, plan, lists_i, lists_s, &tensor_path, &n_failed, &n_candidates,
                        &sample_stride));
  RFM_TRY(xch_barrier(t));      // every rank's lists are complete
  if (n_own > 0) {
    const size_t msmem = (size_t)MERGE_WARPS * world * K * 12;
    RFM_CUDA(cudaFuncSetAttribute(topk_xch_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    const int mgrid = (int)std::min<int64_t>((n_own + MERGE_WARPS - 1) / MERGE_WARPS,
                                             (int64_t)ctx->sm_count * std::max<int64_t>(1, std::min<int64_t>(8, (200 * 1024) / std::max<size_t>(msmem, 1))));
    ListPtrs lp;
    for (int q = 0; q < XCH_MAX_WORLD; ++q) {
      lp.items[q] = q < world ? reinterpret_cast<const int32_t *>(t->x_peer[q] + t->x_items_off) : nullptr;
      lp.scores[q] = q < world ? reinterpret_cast<const double *>(t->x_peer[q] + t->x_scores_off) : nullptr;
    }
    RFM_LAUNCH(ctx, topk_xch_merge_kernel, mgrid, MERGE_WARPS * 32, msmem, lp, world, (int)K, own_begin, own_end,
               t->out_items.p, t->out_scores.p);
  }
  uint32_t status = 0;
  RFM_CUDA(cudaMemcpyAsync(&status, t->x_status.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  t->result_rows = n_own;
  t->own_begin = own_begin;
  t->own_end = own_end;
  RFM_TRY(topk_result_to_host(t, out_items ? n_own : 0, K, out_items, out_scores));   // synchronises the stream
  if (status != 0)
    return fail(RFM_ERR_CUDA, "rfm_topk_run_sharded: rank %u did not reach a cross-GPU barrier within %.0f s",
                status - 1u, (double)XCH_TIMEOUT_NS * 1e-9);
  if (user_range) {
    user_range[0] = own_begin;
    user_range[1] = own_end;
  }
  if (stats) {
    stats[0] = tensor_path ? 1 : 0;
    stats[1] = tensor_path ? (int64_t)(uint32_t)t->counters.p[0] : 0;
    stats[2] = tensor_path ? (int64_t)t->counters.p[1] : 0;
    stats[3] = sample_stride;
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

int rfm_topk_result_host(rfm_topk *t, int32_t K, const int32_t **items_host, const double **scores_host) {
  RFM_REQUIRE(t && items_host && scores_host, "rfm_topk_result_host: NULL argument");
  RFM_REQUIRE(t->out_items.p && t->out_scores.p, "rfm_topk_result_host: call rfm_topk_run first");
  RFM_REQUIRE(K >= 1 && (size_t)t->result_rows * K <= t->out_items.n, "rfm_topk_result_host: K does not match the last run");
  rfm_ctx *ctx = t->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  const size_t n_out = (size_t)t->result_rows * K;
  RFM_TRY(t->stage_items.ensure(n_out));
  RFM_TRY(t->stage_scores.ensure(n_out));
  RFM_CUDA(cudaMemcpyAsync(t->stage_items.p, t->out_items.p, n_out * 4, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(t->stage_scores.p, t->out_scores.p, n_out * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  *items_host = t->stage_items.p;
  *scores_host = t->stage_scores.p;
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
  const size_t msmem = (size_t)MERGE_WARPS * n_lists * K * 12;
  if (n_lists <= XCH_MAX_WORLD && msmem <= 200 * 1024) {      // K-way merge by rank counting (lists are sorted)
    ListPtrs lp;
    for (int q = 0; q < XCH_MAX_WORLD; ++q) {
      lp.items[q] = q < n_lists ? items_dev + (size_t)q * n_users * K : nullptr;
      lp.scores[q] = q < n_lists ? scores_dev + (size_t)q * n_users * K : nullptr;
    }
    RFM_CUDA(cudaFuncSetAttribute(topk_xch_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    const int mgrid = (int)std::min<int64_t>((n_users + MERGE_WARPS - 1) / MERGE_WARPS, (int64_t)ctx->sm_count * 2);
    RFM_LAUNCH(ctx, topk_xch_merge_kernel, mgrid, MERGE_WARPS * 32, msmem, lp, (int)n_lists, (int)K, (int64_t)0, n_users,
               oi.p, os.p);
  } else {
    const int grid = (int)std::min<int64_t>((n_users + 7) / 8, (int64_t)ctx->sm_count * 8);
    RFM_LAUNCH(ctx, topk_merge_kernel, grid, 256, 0, items_dev, scores_dev, n_users, (int)K, (int)n_lists, oi.p, os.p);
  }
  RFM_CUDA(cudaMemcpyAsync(out_items, oi.p, (size_t)n_users * K * 4, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(out_scores, os.p, (size_t)n_users * K * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

}  // extern "C"
