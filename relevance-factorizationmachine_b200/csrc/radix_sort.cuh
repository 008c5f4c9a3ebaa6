// radix_sort.cuh -- stable LSD radix sort of (key u32, pos u32, val T) triples on the device.
//
// Used once per minibatch to turn the batch's row-major non-zeros into a column-major list
// (all entries of feature column j contiguous, in batch order), which is what makes the
// gradient reduction per column deterministic without atomics. 8-bit digits; the number of
// passes is ceil(key_bits/8) (2 for n_features < 65,536). The element count lives on the
// device (*count_dev) because the perf-mode sampler draws the batch there; grids are sized
// from the host-known capacity and every kernel loops over tiles.
//
// Stability: inside a tile each warp owns a contiguous run of elements and ranks them round by
// round with __match_any_sync, so equal digits keep their input order; tiles are ordered by the
// digit-major exclusive scan of the per-tile histograms.
#pragma once
#include "common.cuh"

namespace rfm {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;                       // elements per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // 2048 elements per tile
constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS)
rs_histogram_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ count_dev,
                    int shift, int n_tiles_cap, uint32_t *__restrict__ hist /* [256][n_tiles_cap] */) {
  __shared__ uint32_t h[RS_RADIX];
  const uint32_t count = *count_dev;
  for (int tile = blockIdx.x; tile < n_tiles_cap; tile += gridDim.x) {
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = (uint32_t)tile * RS_TILE;
    if (base < count) {
#pragma unroll
      for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t e = base + i * RS_THREADS + threadIdx.x;
        if (e < count) atomicAdd(&h[(keys[e] >> shift) & 0xFF], 1u);
      }
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_tiles_cap + tile] = h[threadIdx.x];
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ pos_in,
                  const T *__restrict__ val_in, uint32_t *__restrict__ keys_out,
                  uint32_t *__restrict__ pos_out, T *__restrict__ val_out,
                  const uint32_t *__restrict__ count_dev, int shift, int n_tiles_cap,
                  const uint32_t *__restrict__ offsets /* scanned hist */) {
  __shared__ uint32_t wcnt[RS_WARPS][RS_RADIX];
  const uint32_t count = *count_dev;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int tile = blockIdx.x; tile < n_tiles_cap; tile += gridDim.x) {
    const uint32_t tile_base = (uint32_t)tile * RS_TILE;
    if (tile_base >= count) break;  // tiles are visited in increasing order per block
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t warp_base = tile_base + wid * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const uint32_t e = warp_base + r * 32 + lane;
      const bool valid = e < count;
      key[r] = valid ? keys_in[e] : 0u;
      const uint32_t d = valid ? ((key[r] >> shift) & 0xFF) : 0xFFFFFFFFu;
      const uint32_t peers = __match_any_sync(FULL, d);
      const uint32_t before = valid ? wcnt[wid][d] : 0u;
      __syncwarp();
      rank[r] = before + __popc(peers & lt_mask);
      if (valid && (peers & lt_mask) == 0) wcnt[wid][d] = before + __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    {  // digit threadIdx.x: turn per-warp counts into global start offsets, warp by warp
      uint32_t run = offsets[(size_t)threadIdx.x * n_tiles_cap + tile];
#pragma unroll
      for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t c = wcnt[w][threadIdx.x];
        wcnt[w][threadIdx.x] = run;
        run += c;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const uint32_t e = warp_base + r * 32 + lane;
      if (e < count) {
        const uint32_t dst = wcnt[wid][(key[r] >> shift) & 0xFF] + rank[r];
        keys_out[dst] = key[r];
        pos_out[dst] = pos_in[e];
        val_out[dst] = val_in[e];
      }
    }
    __syncthreads();
  }
}

template <typename T>
struct RadixSorter {
  DevBuf<uint32_t> keys[2], pos[2];
  DevBuf<T> val[2];
  DevBuf<uint32_t> hist, tile_sums;
  int64_t capacity = 0;
  int n_tiles_cap = 0;
  int passes = 0;

  int init(int64_t cap, int64_t n_keys) {
    capacity = cap;
    n_tiles_cap = ceil_div(cap > 0 ? cap : 1, RS_TILE);
    int bits = 0;
    for (int64_t v = n_keys > 0 ? n_keys - 1 : 0; v; v >>= 1) ++bits;
    passes = bits == 0 ? 1 : (bits + 7) / 8;
    for (int b = 0; b < 2; ++b) {
      RFM_TRY(keys[b].alloc(cap));
      RFM_TRY(pos[b].alloc(cap));
      RFM_TRY(val[b].alloc(cap));
    }
    RFM_TRY(hist.alloc((size_t)RS_RADIX * n_tiles_cap));
    RFM_TRY(tile_sums.alloc(ceil_div((int64_t)RS_RADIX * n_tiles_cap, 4096) + 2));
    return RFM_OK;
  }
  // input is in buffer 0; returns the index of the buffer holding the sorted output
  int sort(rfm_ctx *ctx, const uint32_t *count_dev, int *out_buf) {
    int cur = 0;
    const int grid = n_tiles_cap < ctx->sm_count * 8 ? n_tiles_cap : ctx->sm_count * 8;
    for (int p = 0; p < passes; ++p) {
      const int shift = 8 * p;
      RFM_LAUNCH(ctx, rs_histogram_kernel, grid, RS_THREADS, 0, keys[cur].p, count_dev, shift, n_tiles_cap,
                 hist.p);
      RFM_TRY(exclusive_scan_u32(ctx, hist.p, hist.p, (int64_t)RS_RADIX * n_tiles_cap, tile_sums.p, nullptr));
      RFM_LAUNCH(ctx, rs_scatter_kernel<T>, grid, RS_THREADS, 0, keys[cur].p, pos[cur].p, val[cur].p,
                 keys[cur ^ 1].p, pos[cur ^ 1].p, val[cur ^ 1].p, count_dev, shift, n_tiles_cap, hist.p);
      cur ^= 1;
    }
    *out_buf = cur;
    return RFM_OK;
  }
};

}  // namespace rfm
