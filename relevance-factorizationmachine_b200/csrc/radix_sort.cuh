// radix_sort.cuh -- stable LSD radix sort of (key u32, pos u32, val T) triples on the device.
//
// Used once per minibatch to turn the batch's row-major non-zeros into a column-major list (all
// entries of feature column j contiguous, in batch order), which is what makes the gradient
// reduction per column deterministic without floating-point atomics. 8-bit digits; the number of
// passes is ceil(key_bits/8) (2 for n_features < 65,536).
//
// "Onesweep" structure: one kernel builds the global digit histograms of every pass; then each pass
// is ONE kernel in which a CTA takes the next tile (dynamic, in order), ranks its elements, publishes
// its per-digit counts, and obtains its exclusive per-digit prefix over earlier tiles by decoupled
// look-back over their published {aggregate | inclusive-prefix} status words. A tile only ever waits
// on tiles with smaller ids, which were handed out earlier and never wait on later ones, so the
// scheme cannot deadlock; spins are bounded anyway and raise a flag the host turns into an error.
//
// Stability: inside a tile each warp owns a contiguous run of elements and ranks them round by
// round with __match_any_sync, so equal digits keep their input order; tiles are ordered by id.
// The element count lives on the device (*count_dev) because the ragged layout computes it there.
#pragma once
#include "common.cuh"

namespace rfm {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef RFM_RS_ITEMS
#define RFM_RS_ITEMS 8
#endif
constexpr int RS_ITEMS = RFM_RS_ITEMS;            // elements per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // 4096 elements per tile
constexpr int RS_RADIX = 256;
constexpr int RS_MAX_PASSES = 4;
constexpr uint32_t RS_FLAG_AGG = 1u << 30, RS_FLAG_PREFIX = 2u << 30, RS_VALUE_MASK = (1u << 30) - 1u;
constexpr uint32_t RS_SPIN_LIMIT = 1u << 22;

__global__ void __launch_bounds__(RS_THREADS)
rs_global_hist_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ count_dev, int passes,
                      uint32_t *__restrict__ ghist /* [passes][256] */) {
  __shared__ uint32_t h[RS_MAX_PASSES][RS_RADIX];
  for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_RADIX; i += RS_THREADS) (&h[0][0])[i] = 0;
  __syncthreads();
  const uint32_t count = *count_dev;
  for (uint32_t e = blockIdx.x * RS_THREADS + threadIdx.x; e < count; e += gridDim.x * RS_THREADS) {
    const uint32_t k = keys[e];
    for (int p = 0; p < passes; ++p) atomicAdd(&h[p][(k >> (8 * p)) & 0xFF], 1u);
  }
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    const uint32_t c = h[p][threadIdx.x];
    if (c) atomicAdd(ghist + p * RS_RADIX + threadIdx.x, c);
  }
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
  return *reinterpret_cast<const volatile uint32_t *>(p);
}

__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) {
  *reinterpret_cast<volatile uint32_t *>(p) = v;
}

template <typename T>
__global__ void __launch_bounds__(RS_THREADS, 3)
rs_onesweep_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ pos_in,
                   const T *__restrict__ val_in, uint32_t *__restrict__ keys_out,
                   uint32_t *__restrict__ pos_out, T *__restrict__ val_out,
                   const uint32_t *__restrict__ count_dev, int shift,
                   const uint32_t *__restrict__ ghist /* [256] of this pass */,
                   uint32_t *status /* [n_tiles][256], zeroed */, uint32_t *tile_counter /* zeroed */,
                   uint32_t *error_flag, int pdl_release) {
  pdl_wait_and_release(pdl_release != 0);
  __shared__ uint32_t wcnt[RS_WARPS][RS_RADIX];
  __shared__ uint32_t digit_base[RS_RADIX];
  __shared__ uint32_t gbase[RS_RADIX];        // global position of tile-local slot 0 of each digit's run
  __shared__ uint32_t warp_tot[RS_WARPS];
  __shared__ uint32_t s_tile;
  // the tile in tile-sorted order, so the write-out below is coalesced inside every digit run
  __shared__ uint32_t sk[RS_TILE], sp[RS_TILE];
  __shared__ T sv[RS_TILE];
  const uint32_t count = *count_dev;
  const uint32_t n_tiles = (count + RS_TILE - 1) / RS_TILE;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  {  // exclusive scan of the global digit histogram: where each digit's output range starts
    const uint32_t v = ghist[threadIdx.x];
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < wid; ++w) before += warp_tot[w];
    digit_base[threadIdx.x] = before + inc - v;
  }
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= n_tiles) break;
    const uint32_t warp_base = tile * RS_TILE + wid * (32 * RS_ITEMS);
    // all of this thread's elements are loaded up front (independent loads in flight together); the scatter
    // phase below only stores
    uint32_t key[RS_ITEMS], rank[RS_ITEMS], ppos[RS_ITEMS];
    T pval[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const uint32_t e = warp_base + r * 32 + lane;
      const bool valid = e < count;
      key[r] = valid ? keys_in[e] : 0u;
      ppos[r] = valid ? pos_in[e] : 0u;
      pval[r] = valid ? val_in[e] : T(0);
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const uint32_t e = warp_base + r * 32 + lane;
      const bool valid = e < count;
      const uint32_t d = valid ? ((key[r] >> shift) & 0xFF) : 0xFFFFFFFFu;
      const uint32_t peers = __match_any_sync(FULL, d);
      const uint32_t before = valid ? wcnt[wid][d] : 0u;
      __syncwarp();
      rank[r] = before + __popc(peers & lt_mask);
      if (valid && (peers & lt_mask) == 0) wcnt[wid][d] = before + __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    {  // digit threadIdx.x: tile count, publish, look back, then per-warp start offsets
      const int d = threadIdx.x;
      uint32_t tile_count = 0;
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) tile_count += wcnt[w][d];
      uint32_t excl = 0;
      uint32_t *mine = status + (size_t)tile * RS_RADIX + d;
      if (tile == 0) {
        st_volatile_u32(mine, RS_FLAG_PREFIX | tile_count);
      } else {
        st_volatile_u32(mine, RS_FLAG_AGG | tile_count);
        // walk back over earlier tiles, LB predecessors per round trip; stop at the first inclusive
        // prefix; re-poll from the first status word that is not published yet
#ifndef RFM_RS_LOOKBACK
#define RFM_RS_LOOKBACK 8
#endif
        constexpr int LB = RFM_RS_LOOKBACK;
        int64_t t = (int64_t)tile - 1;
        uint32_t spins = 0;
        bool done = false;
        while (!done && t >= 0) {
          uint32_t s[LB];
#pragma unroll
          for (int u = 0; u < LB; ++u)
            s[u] = t - u >= 0 ? ld_volatile_u32(status + (size_t)(t - u) * RS_RADIX + d) : RS_FLAG_PREFIX;
          int used = 0;
#pragma unroll
          for (int u = 0; u < LB; ++u) {
            if (!done && used == u) {
              const uint32_t flag = s[u] >> 30;
              if (flag != 0u) {
                excl += s[u] & RS_VALUE_MASK;
                used = u + 1;
                if (flag != 1u) done = true;   // inclusive prefix
              }
            }
          }
          t -= used;
          if (used == 0) {
            if (++spins > RS_SPIN_LIMIT) {
              atomicExch(error_flag, 1u);
              done = true;
            }
            __nanosleep(20);
          }
        }
        st_volatile_u32(mine, RS_FLAG_PREFIX | ((excl + tile_count) & RS_VALUE_MASK));
      }
      // tile-local start of this digit's run (exclusive scan of the tile's digit counts over the CTA)
      uint32_t inc = tile_count;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t2 = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t2;
      }
      if (lane == 31) warp_tot[wid] = inc;
      __syncthreads();
      uint32_t before = 0;
      for (int w = 0; w < wid; ++w) before += warp_tot[w];
      const uint32_t tstart = before + inc - tile_count;
      gbase[d] = digit_base[d] + excl - tstart;
      uint32_t run = tstart;
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t c = wcnt[w][d];
        wcnt[w][d] = run;          // tile-local slot where warp w's first element of this digit goes
        run += c;
      }
    }
    __syncthreads();
    const uint32_t tile_n = count - tile * RS_TILE < (uint32_t)RS_TILE ? count - tile * RS_TILE : (uint32_t)RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const uint32_t e = warp_base + r * 32 + lane;
      if (e < count) {
        const uint32_t slot = wcnt[wid][(key[r] >> shift) & 0xFF] + rank[r];
        sk[slot] = key[r];
        sp[slot] = ppos[r];
        sv[slot] = pval[r];
      }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tile_n; i += RS_THREADS) {
      const uint32_t k = sk[i];
      const uint32_t dst = gbase[(k >> shift) & 0xFF] + i;
      if (dst >= count) {   // cannot happen unless a prefix is wrong: record it instead of faulting
        if (atomicCAS(error_flag, 0u, 2u) == 0u) {
          error_flag[5] = tile;
          error_flag[6] = ((k >> shift) & 0xFF) | (uint32_t)shift << 16;
          error_flag[7] = dst;
        }
        continue;
      }
      keys_out[dst] = k;
      pos_out[dst] = sp[i];
      val_out[dst] = sv[i];
    }
  }
}

template <typename T>
struct RadixSorter {
  DevBuf<uint32_t> keys[2], pos[2];
  DevBuf<T> val[2];
  DevBuf<uint32_t> scratch;   // [ghist: passes*256 | error flag, tile counters: 1 + passes | status: passes*n_tiles*256]
  int64_t capacity = 0;
  int n_tiles_cap = 0;
  int passes = 0;
  size_t scratch_words = 0;

  uint32_t *ghist() { return scratch.p; }
  uint32_t *error_flag() { return scratch.p + RS_MAX_PASSES * RS_RADIX; }
  uint32_t *tile_counter(int p) { return scratch.p + RS_MAX_PASSES * RS_RADIX + 1 + p; }
  uint32_t *status(int p) { return scratch.p + RS_MAX_PASSES * RS_RADIX + 8 + (size_t)p * n_tiles_cap * RS_RADIX; }

  int init(int64_t cap, int64_t n_keys) {
    if (cap >= (int64_t)RS_VALUE_MASK) return fail(RFM_ERR_INVALID, "radix sort: %lld elements is too many", (long long)cap);
    capacity = cap;
    n_tiles_cap = ceil_div(cap > 0 ? cap : 1, RS_TILE);
    int bits = 0;
    for (int64_t v = n_keys > 0 ? n_keys - 1 : 0; v; v >>= 1) ++bits;
    passes = bits == 0 ? 1 : (bits + 7) / 8;
    if (passes > RS_MAX_PASSES) return fail(RFM_ERR_INVALID, "radix sort: keys wider than 32 bits");
    for (int b = 0; b < 2; ++b) {
      RFM_TRY(keys[b].alloc(cap));
      RFM_TRY(pos[b].alloc(cap));
      RFM_TRY(val[b].alloc(cap));
    }
    scratch_words = (size_t)RS_MAX_PASSES * RS_RADIX + 8 + (size_t)passes * n_tiles_cap * RS_RADIX;
    RFM_TRY(scratch.alloc(scratch_words));
    RFM_CUDA(cudaMemsetAsync(scratch.p, 0, scratch_words * sizeof(uint32_t), nullptr));
    return RFM_OK;
  }
  int clear_histograms(rfm_ctx *ctx) {
    RFM_CUDA(cudaMemsetAsync(ghist(), 0, (size_t)RS_MAX_PASSES * RS_RADIX * sizeof(uint32_t), ctx->stream));
    return RFM_OK;
  }
  // input is in buffer 0; returns the index of the buffer holding the sorted output. histograms_ready: the
  // producer of the keys already accumulated ghist() (after clear_histograms), so the histogram pass is skipped.
  // zero the tile counters and status words (the error flag in between is sticky); sort() does it itself unless the
  // caller did it earlier in the stream (prepared = true: no memset then sits between the producer of the keys
  // and the sort kernels, which lets them launch programmatically)
  int prepare(rfm_ctx *ctx) {
    RFM_CUDA(cudaMemsetAsync(tile_counter(0), 0, RS_MAX_PASSES * sizeof(uint32_t), ctx->stream));
    RFM_CUDA(cudaMemsetAsync(status(0), 0, (size_t)passes * n_tiles_cap * RS_RADIX * sizeof(uint32_t), ctx->stream));
    return RFM_OK;
  }
  int sort(rfm_ctx *ctx, const uint32_t *count_dev, int *out_buf, bool histograms_ready = false, bool prepared = false) {
    int cur = 0;
    if (!prepared) RFM_TRY(prepare(ctx));
    if (!histograms_ready) {
      RFM_TRY(clear_histograms(ctx));
      const int hgrid = n_tiles_cap < ctx->sm_count * 4 ? n_tiles_cap : ctx->sm_count * 4;
      RFM_LAUNCH(ctx, rs_global_hist_kernel, hgrid, RS_THREADS, 0, keys[0].p, count_dev, passes, ghist());
    }
    const int grid = n_tiles_cap < ctx->sm_count * 4 ? n_tiles_cap : ctx->sm_count * 4;
    for (int p = 0; p < passes; ++p) {
      if (prepared) {
        RFM_LAUNCH_PDL(ctx, rs_onesweep_kernel<T>, grid, RS_THREADS, 0, (const uint32_t *)keys[cur].p,
                       (const uint32_t *)pos[cur].p, (const T *)val[cur].p, keys[cur ^ 1].p, pos[cur ^ 1].p,
                       val[cur ^ 1].p, count_dev, 8 * p, (const uint32_t *)(ghist() + p * RS_RADIX), status(p),
                       tile_counter(p), error_flag(), pdl_on(ctx) ? 1 : 0);     // prepared: the caller's next kernel waits
      } else {
        RFM_LAUNCH(ctx, rs_onesweep_kernel<T>, grid, RS_THREADS, 0, keys[cur].p, pos[cur].p, val[cur].p,
                   keys[cur ^ 1].p, pos[cur ^ 1].p, val[cur ^ 1].p, count_dev, 8 * p, ghist() + p * RS_RADIX,
                   status(p), tile_counter(p), error_flag(), 0);
      }
      cur ^= 1;
    }
    *out_buf = cur;
    return RFM_OK;
  }
  // host check of the look-back watchdog (call at a point that synchronises anyway)
  int check(rfm_ctx *ctx) {
    uint32_t flag = 0;
    RFM_CUDA(cudaMemcpyAsync(&flag, error_flag(), sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    RFM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag == 2u) {
      uint32_t diag[3] = {0, 0, 0};
      cudaMemcpy(diag, error_flag() + 5, sizeof(diag), cudaMemcpyDeviceToHost);
      return fail(RFM_ERR_CUDA, "radix sort: scatter destination %u out of range (tile %u, digit %u, shift %u)", diag[2],
                  diag[0], diag[1] & 0xFFFF, diag[1] >> 16);
    }
    if (flag) return fail(RFM_ERR_CUDA, "radix sort: decoupled look-back timed out");
    return RFM_OK;
  }
};

}  // namespace rfm
