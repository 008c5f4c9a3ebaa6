// fm.cu -- Factorization Machine hot path: predict, IPS log-loss, and the fused training epoch.
//
// Reference semantics (file:line under the reference root):
//   predict      src/fm.py:114-133   y^ = w0 + X.w + 1/2 sum_f[(X.V)_f^2 - (X^2.V^2)_f], sigmoid
//   residual     src/fm.py:80        e = y/ps - predict(batch)              (pre-update params)
//   _update_w0   src/fm.py:135-143   w0 += lr * sum_t e_t
//   _update_w    src/fm.py:145-154   w_j += lr * sum_t e_t x_tj
//   _update_V    src/fm.py:156-187   v_j += lr * sum_t e_t (x_tj s_t - x_tj^2 v_j),  s_t = (X V)_t
//   losses       src/base.py:37-61   on the batch after the update, and on the val rows
//
// Device layout (T = float or double, chosen at model creation):
//   V    [n_features][kp]   row-major, kp = n_factors rounded up to a multiple of 64, zero padded;
//        a warp reads one row as NCH = kp/64 coalesced 16- or 8-byte-per-lane loads
//   w    [n_features],  w0 [1]
//   rows CSR: row_ptr int64 [n_rows+1], col int32 [nnz], val T [nnz], yp T [n_rows] = y/ps
//   S    [batch][kp]        s_t of the current batch (written by the row pass, read by the column pass)
//   E    [batch]            e_t
//
// One epoch = row pass (forward, residual, emits (column, position, x) triples) -> stable radix
// sort by column -> column pass (segmented reduction over 32-entry chunks of the sorted list,
// in-place SGD for columns that live in one chunk, carry records otherwise) -> carry fix-up ->
// loss pass on the batch -> loss pass on val. No floating-point atomics anywhere: every sum has a
// fixed association, so results are bit-reproducible run to run.
#include "common.cuh"
#include "radix_sort.cuh"
#include "rows.cuh"
#include "sampler.cuh"

using namespace rfm;

// ---- handles ----------------------------------------------------------------------------------
struct rfm_fm {
  rfm_ctx *ctx = nullptr;
  int dtype = RFM_F64;
  int64_t n = 0;
  int k = 0, kp = 0, nch = 0;
  DevBuf<unsigned char> w0, w, V, vn;   // vn[j] = ||v_j||^2
  uint64_t version = 0;                 // bumped by everything that changes the parameters (two-level trainers cache
                                        // per-entity aggregates of them)
};

namespace {

// Tuning knobs, measured on B200 at the KuaiRec-big shape (float64, k = 64, B = 65,536):
// rows kernels capped at 85 registers (3 CTAs/SM), column kernel unrolled 4 deep at 2 CTAs/SM.
#ifndef RFM_ROWS_MIN_BLOCKS
#define RFM_ROWS_MIN_BLOCKS 3
#endif
#ifndef RFM_ROWS_UNROLL
#define RFM_ROWS_UNROLL 4
#endif
#ifndef RFM_COLS_UNROLL
#define RFM_COLS_UNROLL 4
#endif
#ifndef RFM_COLS_MIN_BLOCKS
#define RFM_COLS_MIN_BLOCKS 2
#endif
#define RFM_PRAGMA(x) _Pragma(#x)
#define RFM_UNROLL(n) RFM_PRAGMA(unroll n)
constexpr int ROWS_THREADS = 256;
constexpr int ROWS_WARPS = ROWS_THREADS / 32;
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;

enum RowsMode { MODE_TRAIN = 0, MODE_LOSS = 1, MODE_PREDICT = 2 };

size_t dsize(int dtype) { return dtype == RFM_F64 ? 8 : 4; }

// dense gradient buffer (data-parallel mode): [sum_e, pad x3 | dw (n, padded to x4) | dV (n x kp)].
// Offsets are multiples of 4 elements so dV is 16-byte aligned for float and double vector stores.
constexpr int64_t GRAD_W_OFF = 4;
int64_t grad_v_off(int64_t n) { return GRAD_W_OFF + ((n + 3) / 4) * 4; }
__device__ __forceinline__ int64_t grad_v_off_dev(int64_t n) { return GRAD_W_OFF + ((n + 3) / 4) * 4; }
int64_t grad_total(int64_t n, int kp) { return grad_v_off(n) + n * kp; }

// ---- small conversion kernels -------------------------------------------------------------------
template <typename T>
__global__ void convert_f64_kernel(const double *__restrict__ in, T *__restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = static_cast<T>(in[i]);
}

template <typename T>
__global__ void widen_to_f64_kernel(const T *__restrict__ in, double *__restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = static_cast<double>(in[i]);
}

__global__ void widen_i32_kernel(const int32_t *__restrict__ in, int64_t *__restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}

template <typename T>
__global__ void targets_kernel(const int64_t *__restrict__ y, const double *__restrict__ ps,
                               T *__restrict__ yp, int64_t n) {
  // y / ps exactly as NumPy does it: int64 -> float64, one IEEE division (src/fm.py:80)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    yp[i] = static_cast<T>(static_cast<double>(y[i]) / ps[i]);
}

template <typename T>
__global__ void pad_rows_kernel(const double *__restrict__ in, T *__restrict__ out, int64_t n, int k, int kp) {
  const int64_t total = n * kp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kp;
    const int f = static_cast<int>(i - r * kp);
    out[i] = f < k ? static_cast<T>(in[r * k + f]) : T(0);
  }
}

template <typename T>
__global__ void unpad_rows_kernel(const T *__restrict__ in, double *__restrict__ out, int64_t n, int k, int kp) {
  const int64_t total = n * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k;
    const int f = static_cast<int>(i - r * k);
    out[i] = static_cast<double>(in[r * kp + f]);
  }
}

__global__ void check_columns_kernel(const int32_t *__restrict__ col, int64_t nnz, int64_t n_cols,
                                     int *__restrict__ bad) {
  bool mine = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = col[i];
    mine |= (c < 0 || c >= n_cols);
  }
  if (mine) atomicOr(bad, 1);
}

// ---- batch preparation --------------------------------------------------------------------------
__global__ void feistel_sample_kernel(FeistelKey key, int64_t q0, int64_t batch, int64_t *__restrict__ idx) {
  // positions [q0, q0 + batch) of the epoch's permutation (q0 > 0: a rank's slice of a global batch)
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < batch; q += (int64_t)gridDim.x * blockDim.x)
    idx[q] = static_cast<int64_t>(feistel_permute(static_cast<uint64_t>(q0 + q), key));
}

__global__ void row_len_kernel(const int64_t *__restrict__ row_ptr, const int64_t *__restrict__ idx,
                               int64_t batch, uint32_t *__restrict__ len) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < batch; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx[q];
    len[q] = static_cast<uint32_t>(row_ptr[t + 1] - row_ptr[t]);
  }
}

// ---- the row pass ---------------------------------------------------------------------------------
// How a kernel's per-warp partial sums (sum of residuals, or of loss terms) become one scalar
// without a second launch: every CTA writes its partial, takes a ticket, and the CTA that draws
// the last ticket adds all CTA partials in a fixed order and finishes the scalar.
struct Finish {
  int op;                  // 0: *dst_T += scale * sum (w0 += lr * sum_e), also *dst_d = sum if set
                           // 1: *dst_d = sum * scale  (loss mean)
  double scale;
  void *dst_T;
  double *dst_d;
  double *block_partials;  // [gridDim.x]
  unsigned int *ticket;    // zero before the launch; reset by the finishing CTA
};

template <typename T>
struct RowsArgs {
  const int64_t *row_ptr;
  const int32_t *col;
  const T *val;
  const T *yp;
  const int64_t *idx;     // batch row ids, or nullptr: rows [row0, row0 + n)
  int64_t row0, n;
  const T *w0, *w, *V, *vn;   // vn[j] = ||v_j||^2, kept in step with V
  int kp;
  // MODE_TRAIN outputs
  T *S, *E;
  const uint32_t *bptr;   // ragged layout: exclusive scan of the batch's row lengths
  uint32_t stride;        // fixed-stride layout when > 0: row q owns triples [q*stride, (q+1)*stride)
  uint32_t sentinel;      // key written into unused slots of the fixed-stride layout (== n_features)
  uint32_t *keys, *pos;
  T *xs;
  uint32_t *ghist;        // [n_passes][256] digit histograms of the keys written here (for the radix sort)
  int n_passes;
  // SAMPLED: the batch is drawn here: row id of position q is feistel(q0 + q); also stored to idx_out
  FeistelKey fkey;
  int64_t q0;
  int64_t *idx_out;
  // MODE_PREDICT output
  double *out;
  Finish fin;
  // MODE_LOSS only: an optional second row set scored in the same launch (the val rows next to the batch);
  // positions [n, n + n2) are rows [0, n2) of this set and feed fin2
  const int64_t *row_ptr2;
  const int32_t *col2;
  const T *val2;
  const T *yp2;
  int64_t n2;
  Finish fin2;
  // FAC kernels: the rows (and the optional second row set) are factored; row_ptr / col / val are unused
  FacDev fac, fac2;
  int pdl_release;        // the next operation of the stream is a kernel that waits (common.cuh: pdl_wait_and_release)
};

// ---- factored rows: per-row block layout --------------------------------------------------------------------
// cum[s] = entries of blocks 0..s of the row; base[s] + (position inside the row) = index into block s's own
// arrays (table entry, context column) -- for an id block base[s] is the id itself.
struct FacRow {
  int cum[FAC_MAX_SEG];
  int base[FAC_MAX_SEG];
  int len;
};
// Bit j: context column j of row t holds a non-zero. scipy's csr_matrix(dense) stores no zeros, so the stacked
// matrix the reference builds has no entry there; rows are assembled without them to keep the entry positions
// (and with them the association of every per-row sum) identical to the CSR path.
__device__ __forceinline__ unsigned fac_ctx_mask(const FacDev &f, int64_t t) {
  unsigned m = 0u;
  if (f.es == 8) {
    const double *c = static_cast<const double *>(f.ctx) + t * f.n_ctx;
    for (int j = 0; j < f.n_ctx; ++j) m |= (c[j] != 0.0 ? 1u : 0u) << j;
  } else {
    const float *c = static_cast<const float *>(f.ctx) + t * f.n_ctx;
    for (int j = 0; j < f.n_ctx; ++j) m |= (c[j] != 0.f ? 1u : 0u) << j;
  }
  return m;
}
__device__ __forceinline__ unsigned seg_ctx_mask(const FacSegDev &g, unsigned mask) {
  return (mask >> g.ctx0) & (g.width >= 32 ? 0xFFFFFFFFu : ((1u << g.width) - 1u));
}
__device__ __forceinline__ FacRow fac_row(const FacDev &f, int u, int i, unsigned mask) {
  FacRow r;
  int at = 0;
#pragma unroll
  for (int s = 0; s < FAC_MAX_SEG; ++s) {
    int n = 0, b = 0;
    if (s < f.n_seg) {
      const FacSegDev &g = f.seg[s];
      const int key = g.key == 0 ? u : i;
      if (g.kind == SEG_ID) {
        n = 1;
        b = key;
      } else if (g.kind == SEG_TABLE) {
        const int e0 = __ldg(g.ptr + key);
        n = __ldg(g.ptr + key + 1) - e0;
        b = e0 - at;
      } else {
        const unsigned m = seg_ctx_mask(g, mask);
        n = __popc(m);
        b = static_cast<int>(m);        // the block's non-zero columns; its first entry sits at cum[s - 1]
      }
    }
    at += n;
    r.cum[s] = at;
    r.base[s] = b;
  }
  r.len = at;
  return r;
}
__device__ __forceinline__ int fac_row_len(const FacDev &f, int u, int i, unsigned mask) {
  int at = 0;
  for (int s = 0; s < f.n_seg; ++s) {
    const FacSegDev &g = f.seg[s];
    const int key = g.key == 0 ? u : i;
    at += g.kind == SEG_ID ? 1 : g.kind == SEG_TABLE ? __ldg(g.ptr + key + 1) - __ldg(g.ptr + key)
                                                     : __popc(seg_ctx_mask(g, mask));
  }
  return at;
}
// The block descriptors the entry fetch indexes with a per-lane block number live in shared memory (structure of
// arrays, two row sets): kernel parameters sit in the constant bank, where a warp's loads of different addresses
// are serialised.
struct FacSmem {
  int kind[2 * FAC_MAX_SEG];
  uint32_t col0[2 * FAC_MAX_SEG];
  int ctx0[2 * FAC_MAX_SEG];
  const int32_t *col[2 * FAC_MAX_SEG];
  const void *val[2 * FAC_MAX_SEG];
};
__device__ __forceinline__ void fac_smem_fill(FacSmem &sm, const FacDev &f, const FacDev &f2, bool two) {
  for (int i = threadIdx.x; i < 2 * FAC_MAX_SEG; i += blockDim.x) {
    const bool second = i >= FAC_MAX_SEG;
    const int s = second ? i - FAC_MAX_SEG : i;
    const FacDev &src = (second && two) ? f2 : f;
    const bool on = s < src.n_seg;
    sm.kind[i] = on ? src.seg[s].kind : SEG_ID;
    sm.col0[i] = on ? src.seg[s].col0 : 0u;
    sm.ctx0[i] = on ? src.seg[s].ctx0 : 0;
    sm.col[i] = on ? src.seg[s].col : nullptr;
    sm.val[i] = on ? src.seg[s].val : nullptr;
  }
}
// entry `off` (< r.len) of row t: global column and value. Branch-free: every lane issues exactly two loads (a
// table block's column and value; an id / context block loads a dummy column and its value from `one` / the row's
// context record), so the lanes of a row group, which sit in different blocks, do not diverge.
template <typename T>
__device__ __forceinline__ void fac_entry(const FacSmem &sm, int set, const FacDev &f, const FacRow &r, int64_t t,
                                          int off, int &c, T &x) {
  int s = 0, b = r.base[0], start = 0;
#pragma unroll
  for (int j = 1; j < FAC_MAX_SEG; ++j)
    if (off >= r.cum[j - 1]) {        // cum is non-decreasing and off < len: the last block that starts at or before off
      s = j;
      b = r.base[j];
      start = r.cum[j - 1];
    }
  const int at = set * FAC_MAX_SEG + s;
  const int kind = sm.kind[at];
  // context block: the (off - start)-th non-zero column of the block (b holds the block's non-zero mask)
  const int j = f.n_ctx <= 1 ? 0 : (kind == SEG_CTX ? static_cast<int>(__fns(static_cast<unsigned>(b), 0u, off - start + 1)) : 0);
  const int e = kind == SEG_TABLE ? b + off : 0;
  const int32_t *cp = kind == SEG_TABLE ? sm.col[at] + e : f.user;
  const T *xp = kind == SEG_TABLE ? static_cast<const T *>(sm.val[at]) + e
              : kind == SEG_CTX   ? static_cast<const T *>(f.ctx) + t * f.n_ctx + sm.ctx0[at] + j
                                  : static_cast<const T *>(f.one);
  const int cv = __ldg(cp);
  x = __ldg(xp);
  c = static_cast<int>(sm.col0[at]) + (kind == SEG_TABLE ? cv : kind == SEG_ID ? b : j);
}

// ---- factored rows: upload helpers ---------------------------------------------------------------------------
template <typename S>
__global__ void narrow_ids_kernel(const S *in, int32_t *out, int64_t n, int64_t limit,   // in may alias out
                                  int *__restrict__ bad) {
  bool mine = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = static_cast<int64_t>(in[i]);
    mine |= (v < 0 || v >= limit);
    out[i] = static_cast<int32_t>(v);
  }
  if (mine) atomicOr(bad, 1);
}

// labels of 1, 4 or 8 bytes (signed integers) -> y / pscore as NumPy computes it (src/fm.py:80)
template <typename T, typename Y>
__global__ void targets_any_kernel(const Y *__restrict__ y, const double *__restrict__ ps, T *__restrict__ yp, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    yp[i] = static_cast<T>(static_cast<double>(y[i]) / ps[i]);
}

// the same with the propensity looked up per ITEM (pscore = theta_item ** pow_used is an item-level quantity in both of
// the reference's loaders: coat/_preparer.py:56-62, kuairec/loader.py:160-168): 8 bytes per interaction less to upload
template <typename T, typename Y>
__global__ void targets_by_item_kernel(const Y *__restrict__ y, const int32_t *__restrict__ item,
                                       const double *__restrict__ item_ps, T *__restrict__ yp, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    yp[i] = static_cast<T>(static_cast<double>(y[i]) / __ldg(item_ps + item[i]));
}

// one context block's columns into the per-row context record
template <typename T>
__global__ void ctx_pack_kernel(const double *__restrict__ in, int64_t n_rows, int width, T *__restrict__ out,
                                int n_ctx, int ctx0) {
  const int64_t total = n_rows * width;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / width;
    const int c = static_cast<int>(i - r * width);
    out[r * n_ctx + ctx0 + c] = static_cast<T>(in[i]);
  }
}

// non-zeros of the whole row set and the longest row (what the CSR path reads off the row pointers)
__global__ void fac_stats_kernel(const FacDev f, int64_t n_rows, unsigned long long *__restrict__ nnz,
                                 int *__restrict__ max_len) {
  unsigned long long s = 0;
  int m = 0;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_rows; t += (int64_t)gridDim.x * blockDim.x) {
    const int len = fac_row_len(f, f.user[t], f.item[t], fac_ctx_mask(f, t));
    s += (unsigned long long)len;
    m = max(m, len);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(FULL, s, o);
    m = max(m, __shfl_xor_sync(FULL, m, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(nnz, s);
    atomicMax(max_len, m);
  }
}

__global__ void fac_row_len_kernel(const FacDev f, const int64_t *__restrict__ idx, int64_t batch,
                                   uint32_t *__restrict__ len) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < batch; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = idx[q];
    len[q] = static_cast<uint32_t>(fac_row_len(f, f.user[t], f.item[t], fac_ctx_mask(f, t)));
  }
}

// ---- factored rows -> stacked CSR on the device (rfm_rows_materialize) ------------------------------------------
__global__ void fac_all_row_len_kernel(const FacDev f, int64_t n_rows, uint32_t *__restrict__ len) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_rows; t += (int64_t)gridDim.x * blockDim.x)
    len[t] = static_cast<uint32_t>(fac_row_len(f, f.user[t], f.item[t], fac_ctx_mask(f, t)));
}

// row_ptr (int64) from the exclusive scan of the lengths, and the entries of every row in the stacked column order
template <typename T>
__global__ void __launch_bounds__(256)
fac_fill_csr_kernel(const FacDev f, int64_t n_rows, const uint32_t *__restrict__ start, uint32_t total,
                    int64_t *__restrict__ row_ptr, int32_t *__restrict__ col, T *__restrict__ val) {
  __shared__ FacSmem fsm;
  fac_smem_fill(fsm, f, f, false);
  __syncthreads();
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t <= n_rows; t += (int64_t)gridDim.x * blockDim.x) {
    if (t == n_rows) {
      row_ptr[t] = total;
      continue;
    }
    const uint32_t at = start[t];
    row_ptr[t] = at;
    const FacRow fr = fac_row(f, f.user[t], f.item[t], fac_ctx_mask(f, t));
    for (int off = 0; off < fr.len; ++off) {
      int c;
      T x;
      fac_entry<T>(fsm, 0, f, fr, t, off, c, x);
      col[at + off] = c;
      val[at + off] = x;
    }
  }
}

__device__ __forceinline__ double sigmoid_ref(double z) {
  // src/base.py:63-66
  z = fmin(fmax(z, -700.0), 700.0);
  return 1.0 / (1.0 + exp(-z));
}

template <typename T>
__device__ __forceinline__ void block_finish(double warp_partial, const Finish &f) {
  __shared__ double wsum[32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  if (lane == 0) wsum[wid] = warp_partial;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < n_warps; ++w) s += wsum[w];
    f.block_partials[blockIdx.x] = s;
    __threadfence();
    is_last = atomicAdd(f.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double s = 0.0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += __ldcg(f.block_partials + i);
  s = warp_sum(s);
  if (lane == 0) wsum[wid] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < n_warps; ++w) tot += wsum[w];
    if (f.op == 0) {
      if (f.dst_T) {
        T *d = static_cast<T *>(f.dst_T);
        *d = static_cast<T>(static_cast<double>(*d) + f.scale * tot);
      }
      if (f.dst_d) *f.dst_d = tot;
    } else {
      *f.dst_d = tot * f.scale;
    }
    *f.ticket = 0u;
  }
}

// sum over the TPR lanes of a row group (xor butterfly; every lane of the group gets the total)
template <int TPR, typename T>
__device__ __forceinline__ T group_sum(T v, unsigned mask) {
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// TPR lanes cooperate on one row (32/TPR rows per warp); lane g of the group holds the NCV
// 2-element chunks {(ch*TPR + g)*2, +1}, so each load instruction of a group reads TPR*2 contiguous
// elements of a V row. The per-non-zero bookkeeping (broadcast of (column, x), address arithmetic,
// loop control) and the scalar epilogue are shared by all rows of the warp.
template <typename T, int TPR, int NCV, int MODE, bool SAMPLED, bool FAC>
__global__ void __launch_bounds__(ROWS_THREADS, NCV <= 4 ? RFM_ROWS_MIN_BLOCKS : 1)
fm_rows_kernel(const RowsArgs<T> a) {
  pdl_wait_and_release(a.pdl_release != 0);
  using V2 = typename Vec2<T>::type;
  constexpr int GPW = 32 / TPR;
  const int lane = lane_id(), g = lane % TPR, grp = lane / TPR;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_groups = nw * GPW;
  const T w0 = __ldg(a.w0);
  // MODE_TRAIN: this kernel writes every sort key, so it also counts their digits (integer shared-memory
  // atomics, flushed once per CTA) and the sort needs no histogram pass of its own
  __shared__ uint32_t hist[MODE == MODE_TRAIN ? RS_MAX_PASSES * RS_RADIX : 1];
  if (MODE == MODE_TRAIN) {
    for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_RADIX; i += ROWS_THREADS) hist[i] = 0;
    __syncthreads();
  }
  __shared__ FacSmem fsm;
  if (FAC) {
    fac_smem_fill(fsm, a.fac, a.fac2, MODE == MODE_LOSS && a.n2 > 0);
    __syncthreads();
  }
  double partial = 0.0, partial2 = 0.0;
  const int64_t n_total = a.n + (MODE == MODE_LOSS ? a.n2 : 0);
  // Row metadata (row id, CSR extent, target) of the NEXT loop iteration is fetched while the current row is
  // being processed: these are dependent random DRAM reads (id -> row_ptr pair, yp) and would otherwise sit
  // on the critical path of every row. (A second stage that also prefetched the next row's first column
  // entries was measured and did not pay: 59 -> 61-65 us.)
  struct RowMeta {
    int64_t t, beg, end;      // FAC: beg = the row's user id, end = its item id
    T ypv;
    unsigned mask;            // FAC: non-zero context columns of the row
    bool active, second;
  };
  auto fetch = [&](int64_t qq) {
    RowMeta r;
    r.t = r.beg = r.end = 0;
    r.ypv = T(0);
    r.mask = 0u;
    r.active = qq < n_total;
    r.second = MODE == MODE_LOSS && qq >= a.n;
    if (r.active) {
      if (r.second) {
        r.t = qq - a.n;
        if (FAC) {
          r.beg = a.fac2.user[r.t];
          r.end = a.fac2.item[r.t];
          r.mask = fac_ctx_mask(a.fac2, r.t);
        } else {
          r.beg = a.row_ptr2[r.t];
          r.end = a.row_ptr2[r.t + 1];
        }
        r.ypv = a.yp2[r.t];
      } else {
        if (SAMPLED) {
          r.t = static_cast<int64_t>(feistel_permute(static_cast<uint64_t>(a.q0 + qq), a.fkey));
          if (g == 0) a.idx_out[qq] = r.t;
        } else {
          r.t = a.idx ? a.idx[qq] : a.row0 + qq;
        }
        if (FAC) {
          r.beg = a.fac.user[r.t];
          r.end = a.fac.item[r.t];
          r.mask = fac_ctx_mask(a.fac, r.t);
        } else {
          r.beg = a.row_ptr[r.t];
          r.end = a.row_ptr[r.t + 1];
        }
        if (MODE != MODE_PREDICT) r.ypv = a.yp[r.t];
      }
    }
    return r;
  };
  RowMeta nxt = fetch(gw * GPW + grp);
  for (int64_t qb = gw * GPW; qb < n_total; qb += n_groups) {   // warp-uniform trip count
    const int64_t q = qb + grp;
    const RowMeta cur = nxt;
    nxt = fetch(q + n_groups);                                   // loads for the next row start now
    const bool active = cur.active, second = cur.second;
    const int32_t *colp = second ? a.col2 : a.col;
    const T *valp = second ? a.val2 : a.val;
    const int64_t t = cur.t, beg = cur.beg, end = cur.end;
    FacRow fr;
    if (FAC) {
      if (MODE == MODE_LOSS && second) fr = fac_row(a.fac2, static_cast<int>(beg), static_cast<int>(end), cur.mask);
      else fr = fac_row(a.fac, static_cast<int>(beg), static_cast<int>(end), cur.mask);
    }
    const int len = FAC ? (active ? fr.len : 0) : static_cast<int>(end - beg);
    const int maxlen = __reduce_max_sync(FULL, len);
    V2 acc[NCV];
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) acc[ch].x = acc[ch].y = T(0);
    // scalar part of the logit owned by this lane: sum over its own entries of x w_j - x^2 ||v_j||^2 / 2
    T sl = T(0);
    uint32_t out_base = 0;
    if (MODE == MODE_TRAIN && active) out_base = a.stride ? static_cast<uint32_t>(q) * a.stride : a.bptr[q];
    for (int off0 = 0; off0 < maxlen; off0 += TPR) {
      const int off = off0 + g;
      int c = 0;
      T x = T(0);
      if (off < len) {
        if (FAC) {
          if (MODE == MODE_LOSS && second) fac_entry<T>(fsm, 1, a.fac2, fr, t, off, c, x);
          else fac_entry<T>(fsm, 0, a.fac, fr, t, off, c, x);
        } else {
          c = colp[beg + off];
          x = valp[beg + off];
        }
        sl += x * __ldg(a.w + c) - T(0.5) * (x * x) * __ldg(a.vn + c);
        if (MODE == MODE_TRAIN) {
          const uint32_t o = out_base + static_cast<uint32_t>(off);
          a.keys[o] = static_cast<uint32_t>(c);
          a.pos[o] = static_cast<uint32_t>(q);
          a.xs[o] = x;
          for (int ps = 0; ps < a.n_passes; ++ps)
            atomicAdd(&hist[ps * RS_RADIX + ((static_cast<uint32_t>(c) >> (8 * ps)) & 0xFF)], 1u);
        }
      }
      const int cnt = maxlen - off0 < TPR ? maxlen - off0 : TPR;
      RFM_UNROLL(RFM_ROWS_UNROLL)
      for (int i = 0; i < cnt; ++i) {
        const int cj = __shfl_sync(FULL, c, i, TPR);     // entries past a shorter row's end carry x = 0
        const T xj = __shfl_sync(FULL, x, i, TPR);
        const V2 *vrow = reinterpret_cast<const V2 *>(a.V + (size_t)cj * a.kp) + g;
#pragma unroll
        for (int ch = 0; ch < NCV; ++ch) {
          const V2 v = __ldg(vrow + ch * TPR);
          acc[ch].x += xj * v.x;
          acc[ch].y += xj * v.y;
        }
      }
    }
    if (MODE == MODE_TRAIN && active && a.stride) {   // unused slots of this row sort behind every real column
      for (uint32_t si = static_cast<uint32_t>(len) + g; si < a.stride; si += TPR) {
        a.keys[out_base + si] = a.sentinel;
        for (int ps = 0; ps < a.n_passes; ++ps)
          atomicAdd(&hist[ps * RS_RADIX + ((a.sentinel >> (8 * ps)) & 0xFF)], 1u);
      }
    }
    T ss = T(0);
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) ss += acc[ch].x * acc[ch].x + acc[ch].y * acc[ch].y;
    const T z = w0 + group_sum<TPR>(sl + T(0.5) * ss, FULL);
    const double p = sigmoid_ref(static_cast<double>(z));
    if (active) {
      if (MODE == MODE_TRAIN) {
        const double e = static_cast<double>(cur.ypv) - p;
        V2 *srow = reinterpret_cast<V2 *>(a.S + (size_t)q * a.kp) + g;
#pragma unroll
        for (int ch = 0; ch < NCV; ++ch) srow[ch * TPR] = acc[ch];
        if (g == 0) {
          a.E[q] = static_cast<T>(e);
          partial += e;
        }
      } else if (MODE == MODE_LOSS) {
        // src/base.py:56-59 term by term (1 - p formed by subtraction, eps inside both logs)
        if (g == 0) {
          const double r = static_cast<double>(cur.ypv);
          const double term = r * log(p + 1e-8) + (1.0 - r) * log(1.0 - p + 1e-8);
          if (second) partial2 -= term; else partial -= term;
        }
      } else {
        if (g == 0) a.out[q] = p;
      }
    }
  }
  if (MODE == MODE_TRAIN) {
    __syncthreads();
    for (int i = threadIdx.x; i < a.n_passes * RS_RADIX; i += ROWS_THREADS)
      if (hist[i]) atomicAdd(a.ghist + i, hist[i]);
  }
  if (MODE != MODE_PREDICT) block_finish<T>(warp_sum(partial), a.fin);
  if (MODE == MODE_LOSS && a.n2 > 0) block_finish<T>(warp_sum(partial2), a.fin2);
}

// ||v_j||^2 for every row of V (after set_params and after a dense data-parallel apply)
template <typename T>
__global__ void __launch_bounds__(ROWS_THREADS)
row_norms_kernel(const T *__restrict__ V, T *__restrict__ vn, int64_t n, int kp) {
  const int lane = lane_id();
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = gw; j < n; j += nw) {
    T s = T(0);
    for (int f = lane * 2; f < kp; f += 64) {
      const T v0 = V[j * kp + f], v1 = V[j * kp + f + 1];
      s += v0 * v0 + v1 * v1;
    }
    s = warp_sum(s);
    if (lane == 0) vn[j] = s;
  }
}

// ---- the column pass --------------------------------------------------------------------------------
template <typename T>
struct ColsArgs {
  const uint32_t *keys, *pos;
  const T *xs;
  const uint32_t *count;
  uint32_t sentinel;  // keys equal to this are padding of the fixed-stride layout (sorted last)
  const T *E, *S;
  T *V, *w, *vn;
  int kp;
  T lr;
  uint32_t unit;     // sorted entries covered by one carry record = one CTA iteration of the column pass
  T *carry_vec;      // [n_units][2][kp]    slot 0 = HEAD (run began in an earlier unit), 1 = TAIL
  T *carry_ac;       // [n_units][2][2]     (a, c)
  // OUT_GRAD (data-parallel mode): write the gradient instead of applying it.
  // OUT_RAW (level 1 of the two-level step): the raw sums go out as they are: grad_V[j] = sum e x s, grad_w[j] = sum e x,
  // raw_c[j] = sum e x^2
  T *grad_w, *grad_V, *raw_c;
  // L2 = true (level 2 of the two-level step): the list is the static (real column, virtual column, x) list and
  // "S" holds level 1's vector sums: acc += x S[p], a += x Ea[p], c += x^2 Ec[p]
  const T *Ea, *Ec;
  // Level 1 of the two-level step with one dense context column: the first ctx_blocks CTAs of the launch do not walk
  // the sorted list but sum that column over all ctx_rows batch positions (c_q = Cq[q]; every row holds the column, so
  // it needs no sort): per-CTA partials in cpart, combined by the last of them into the OUT_RAW outputs of
  // virtual column ctx_col
  uint32_t ctx_blocks, ctx_col;
  int64_t ctx_rows;
  const T *Cq;
  T *cpart;
  uint32_t *ctx_ticket;
  // fix-up work lists (order of the lists is irrelevant: every entry is handled independently)
  uint32_t *tails;       // chunks that own a column continuing into later chunks
  uint32_t *n_tails;
  int pdl_release;       // the next operation of the stream is a kernel that waits (set per launch)
};

enum ColsOut { OUT_SGD = 0, OUT_GRAD = 1, OUT_RAW = 2 };

template <typename T, int TPR, int NCV, int OUT>
__device__ __forceinline__ void finish_column(const ColsArgs<T> &a, uint32_t colj,
                                              const typename Vec2<T>::type (&acc)[NCV], T sa, T sc, int g,
                                              unsigned gmask) {
  using V2 = typename Vec2<T>::type;
  if (OUT == OUT_RAW) {
    V2 *grow = reinterpret_cast<V2 *>(a.grad_V + (size_t)colj * a.kp) + g;
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) grow[ch * TPR] = acc[ch];
    if (g == 0) {
      a.grad_w[colj] = sa;
      a.raw_c[colj] = sc;
    }
    return;
  }
  V2 *vrow = reinterpret_cast<V2 *>(a.V + (size_t)colj * a.kp) + g;
  if (OUT == OUT_GRAD) {
    V2 *grow = reinterpret_cast<V2 *>(a.grad_V + (size_t)colj * a.kp) + g;
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) {
      const V2 v = vrow[ch * TPR];
      V2 gr;
      gr.x = acc[ch].x - sc * v.x;
      gr.y = acc[ch].y - sc * v.y;
      grow[ch * TPR] = gr;
    }
    if (g == 0) a.grad_w[colj] = sa;
  } else {
    T nrm = T(0);
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) {
      V2 v = vrow[ch * TPR];
      v.x += a.lr * (acc[ch].x - sc * v.x);
      v.y += a.lr * (acc[ch].y - sc * v.y);
      vrow[ch * TPR] = v;
      nrm += v.x * v.x + v.y * v.y;
    }
    nrm = group_sum<TPR>(nrm, gmask);
    if (g == 0) {
      a.w[colj] += a.lr * sa;
      a.vn[colj] = nrm;
    }
  }
}

template <typename T, int TPR, int NCV>
__device__ __forceinline__ void store_carry(const ColsArgs<T> &a, uint32_t chunk, int slot,
                                            const typename Vec2<T>::type (&acc)[NCV], T sa, T sc, int g) {
  using V2 = typename Vec2<T>::type;
  V2 *cv = reinterpret_cast<V2 *>(a.carry_vec + ((size_t)chunk * 2 + slot) * a.kp) + g;
#pragma unroll
  for (int ch = 0; ch < NCV; ++ch) cv[ch * TPR] = acc[ch];
  if (g == 0) {
    a.carry_ac[((size_t)chunk * 2 + slot) * 2 + 0] = sa;
    a.carry_ac[((size_t)chunk * 2 + slot) * 2 + 1] = sc;
  }
}

// The dense context column of level 1 (see ColsArgs::ctx_blocks): sum over q of (c_q e_q) s_q, c_q e_q and c_q^2 e_q.
// Rows are dealt to the CTAs in contiguous ranges and inside a CTA to the row groups round-robin; every group adds
// its rows in order, groups of a warp combine by xor butterfly, warps in warp order, CTAs in three contiguous
// blocks of CTA order by the last CTA to finish: one fixed association per (batch, launch shape).
template <typename T, int TPR, int NCV>
__device__ __forceinline__ void ctx_column_sums(const ColsArgs<T> &a, unsigned char *smem) {
  using V2 = typename Vec2<T>::type;
  constexpr int GPW = 32 / TPR;
  const int lane = lane_id(), g = lane % TPR, grp = lane / TPR, wid = threadIdx.x >> 5;
  const int pstride = a.kp + 2;
  T *wpart = reinterpret_cast<T *>(smem);            // [ROWS_WARPS][pstride], then [3][pstride]
  __shared__ bool ctx_last;
  const int64_t nb = a.ctx_blocks, per = (a.ctx_rows + nb - 1) / nb;
  const int64_t lo = blockIdx.x * per, hi = min(a.ctx_rows, lo + per);
  V2 acc[NCV];
#pragma unroll
  for (int ch = 0; ch < NCV; ++ch) acc[ch].x = acc[ch].y = T(0);
  T sa = T(0), sc = T(0);
  RFM_UNROLL(4)
  for (int64_t r = lo + wid * GPW + grp; r < hi; r += ROWS_WARPS * GPW) {
    const T c = __ldg(a.Cq + r), ev = __ldg(a.E + r);
    const T xe = c * ev, xxe = c * c * ev;
    const V2 *srow = reinterpret_cast<const V2 *>(a.S + (size_t)r * a.kp) + g;
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) {
      const V2 sv = __ldg(srow + ch * TPR);
      acc[ch].x += xe * sv.x;
      acc[ch].y += xe * sv.y;
    }
    sa += xe;
    sc += xxe;
  }
#pragma unroll
  for (int o = TPR; o < 32; o <<= 1) {
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) {
      acc[ch].x += __shfl_xor_sync(FULL, acc[ch].x, o);
      acc[ch].y += __shfl_xor_sync(FULL, acc[ch].y, o);
    }
    sa += __shfl_xor_sync(FULL, sa, o);
    sc += __shfl_xor_sync(FULL, sc, o);
  }
  if (lane < TPR) {
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) {
      wpart[wid * pstride + (ch * TPR + g) * 2] = acc[ch].x;
      wpart[wid * pstride + (ch * TPR + g) * 2 + 1] = acc[ch].y;
    }
    if (lane == 0) {
      wpart[wid * pstride + a.kp] = sa;
      wpart[wid * pstride + a.kp + 1] = sc;
    }
  }
  __syncthreads();
  for (int f = threadIdx.x; f < pstride; f += ROWS_THREADS) {
    T s = T(0);
    for (int w = 0; w < ROWS_WARPS; ++w) s += wpart[w * pstride + f];
    a.cpart[(size_t)blockIdx.x * pstride + f] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ctx_last = atomicAdd(a.ctx_ticket, 1u) == a.ctx_blocks - 1u;
  __syncthreads();
  if (!ctx_last) return;
  __threadfence();
  const unsigned seg_len = (a.ctx_blocks + 2u) / 3u;
  for (int i = threadIdx.x; i < 3 * pstride; i += ROWS_THREADS) {
    const unsigned sg = i / pstride, f = i - sg * pstride;
    const unsigned b0 = min(a.ctx_blocks, sg * seg_len), b1 = min(a.ctx_blocks, b0 + seg_len);
    T s = T(0);
    unsigned b = b0;
    for (; b + 8 <= b1; b += 8) {
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(a.cpart + (size_t)(b + u) * pstride + f);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; b < b1; ++b) s += __ldcg(a.cpart + (size_t)b * pstride + f);
    wpart[i] = s;
  }
  __syncthreads();
  for (int f = threadIdx.x; f < pstride; f += ROWS_THREADS) {
    const T s = (wpart[f] + wpart[pstride + f]) + wpart[2 * pstride + f];
    if (f < a.kp) a.grad_V[(size_t)a.ctx_col * a.kp + f] = s;
    else if (f == a.kp) a.grad_w[a.ctx_col] = s;
    else a.raw_c[a.ctx_col] = s;
  }
  if (threadIdx.x == 0) *a.ctx_ticket = 0u;
}

// Column pass. A group of TPR lanes walks one chunk of 32 consecutive entries of the column-sorted list
// (entry i lives in lane i % TPR, register i / TPR); a CTA covers GPC = 8 * 32/TPR consecutive chunks per
// iteration ("unit"). Columns that live inside one chunk are finished by their group. Partial sums of a
// column that crosses a chunk's start (HEAD) or end (TAIL) are parked in shared memory; after a CTA barrier
// the group where a run starts walks the following chunks' HEAD partials in chunk order and either finishes
// the column (the run ends inside the unit) or emits ONE carry record for the unit. Only runs that cross
// unit boundaries reach the global fix-up kernels.
// SMALL: chunks of TPR entries instead of 32 (a unit is then 256 entries for every TPR): four times the groups for
// lists too short to fill the machine with 32-entry chunks (the two-level step's lists at the KuaiRec shape).
template <typename T, int TPR, int NCV, int OUT, bool L2, bool SMALL = false>
__global__ void __launch_bounds__(ROWS_THREADS, NCV <= 4 ? RFM_COLS_MIN_BLOCKS : 1)
fm_cols_kernel(const ColsArgs<T> a) {
  pdl_wait_and_release(a.pdl_release != 0);
  using V2 = typename Vec2<T>::type;
  constexpr int GPW = 32 / TPR, NJ = SMALL ? 1 : 32 / TPR, GPC = ROWS_WARPS * GPW;
  constexpr uint32_t CH = TPR * NJ;                              // entries per chunk
  extern __shared__ __align__(16) unsigned char cols_smem[];
  const int pstride = a.kp + 2;                                  // partial = kp values + (a, c)
  T *part = reinterpret_cast<T *>(cols_smem);                    // [GPC][2][pstride]
  uint32_t *hkey = reinterpret_cast<uint32_t *>(part + (size_t)GPC * 2 * pstride);   // [GPC] key of the HEAD partial
  uint32_t *tkey = hkey + GPC;                                   // [GPC] key of the TAIL partial
  uint32_t *flag = tkey + GPC;                                   // [GPC] bit0 HEAD, bit1 TAIL, bit2 HEAD runs on past the chunk
  const int lane = lane_id(), g = lane % TPR, grp = lane / TPR;
  const int wid = threadIdx.x >> 5;
  const int gi = wid * GPW + grp;                                // this group's chunk inside the unit
  const unsigned gmask = TPR == 32 ? FULL : (((1u << TPR) - 1u) << (grp * TPR));
  if (OUT == OUT_RAW && blockIdx.x < a.ctx_blocks) {             // block-uniform: these CTAs sum the dense context column
    ctx_column_sums<T, TPR, NCV>(a, cols_smem);
    return;
  }
  const uint32_t M = *a.count;
  const uint32_t n_chunks = (M + CH - 1u) / CH;
  const uint32_t n_units = (n_chunks + GPC - 1) / GPC;
  const uint32_t first_cta = OUT == OUT_RAW ? a.ctx_blocks : 0u, n_ctas = gridDim.x - first_cta;
  for (uint32_t unit = blockIdx.x - first_cta; unit < n_units; unit += n_ctas) {   // CTA-uniform trip count
    const uint32_t chunk = unit * GPC + gi;
    const uint32_t base = chunk * CH;
    uint32_t key[NJ], p[NJ];
    T xe[NJ], xxe[NJ], xa[L2 ? NJ : 1];
    int n_valid = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const uint32_t e = base + j * TPR + g;
      uint32_t k = (chunk < n_chunks && e < M) ? a.keys[e] : KEY_NONE;
      if (k == a.sentinel) k = KEY_NONE;
      const bool valid = k != KEY_NONE;
      key[j] = k;
      p[j] = valid ? a.pos[e] : 0u;
      const T x = valid ? a.xs[e] : T(0);
      if (L2) {
        const T ea = valid ? __ldg(a.Ea + p[j]) : T(0), ec = valid ? __ldg(a.Ec + p[j]) : T(0);
        xe[j] = x;
        xa[L2 ? j : 0] = x * ea;
        xxe[j] = x * x * ec;
      } else {
        const T ev = valid ? __ldg(a.E + p[j]) : T(0);
        xe[j] = x * ev;
        xxe[j] = x * x * ev;
      }
      n_valid += __popc(__ballot_sync(FULL, valid) & gmask);   // padding sorts last: valid entries are a prefix
    }
    uint32_t prev_key = KEY_NONE, next_key = KEY_NONE;
    if (n_valid > 0) {
      if (base > 0) prev_key = a.keys[base - 1];
      if (base + CH < M) next_key = a.keys[base + CH];
      if (next_key == a.sentinel) next_key = KEY_NONE;
    }
    V2 acc[NCV];
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) acc[ch].x = acc[ch].y = T(0);
    T sa = T(0), sc = T(0);
    uint32_t cur = __shfl_sync(FULL, key[0], 0, TPR);
    bool head = (n_valid > 0) && (cur == prev_key);
    uint32_t my_flag = 0;
    T *my_head = part + ((size_t)gi * 2 + 0) * pstride, *my_tail = part + ((size_t)gi * 2 + 1) * pstride;
    auto park = [&](T *dst) {
#pragma unroll
      for (int ch = 0; ch < NCV; ++ch) {
        dst[(ch * TPR + g) * 2] = acc[ch].x;
        dst[(ch * TPR + g) * 2 + 1] = acc[ch].y;
      }
      if (g == 0) {
        dst[a.kp] = sa;
        dst[a.kp + 1] = sc;
      }
    };
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      RFM_UNROLL(RFM_COLS_UNROLL)
      for (int ii = 0; ii < TPR; ++ii) {
        const int i = j * TPR + ii;
        const uint32_t ki = __shfl_sync(FULL, key[j], ii, TPR);
        const uint32_t pi = __shfl_sync(FULL, p[j], ii, TPR);
        const T xei = __shfl_sync(FULL, xe[j], ii, TPR);
        const T xxei = __shfl_sync(FULL, xxe[j], ii, TPR);
        const T xai = L2 ? __shfl_sync(FULL, xa[L2 ? j : 0], ii, TPR) : xei;
        if (i < n_valid) {                                      // group-uniform
          const V2 *srow = reinterpret_cast<const V2 *>(a.S + (size_t)pi * a.kp) + g;
          V2 sv[NCV];
#pragma unroll
          for (int ch = 0; ch < NCV; ++ch) sv[ch] = __ldg(srow + ch * TPR);
          if (ki != cur) {  // the previous column's segment ended inside this chunk
            if (head) {
              park(my_head);
              if (g == 0) hkey[gi] = cur;
              my_flag |= 1u;
            } else {
              finish_column<T, TPR, NCV, OUT>(a, cur, acc, sa, sc, g, gmask);
            }
            head = false;
            cur = ki;
#pragma unroll
            for (int ch = 0; ch < NCV; ++ch) acc[ch].x = acc[ch].y = T(0);
            sa = sc = T(0);
          }
#pragma unroll
          for (int ch = 0; ch < NCV; ++ch) {
            acc[ch].x += xei * sv[ch].x;
            acc[ch].y += xei * sv[ch].y;
          }
          sa += xai;
          sc += xxei;
        }
      }
    }
    if (n_valid > 0) {
      const bool tail = (cur == next_key);
      if (head) {                 // the chunk's only segment began earlier; bit2: it also runs on
        park(my_head);
        if (g == 0) hkey[gi] = cur;
        my_flag |= 1u | (tail ? 4u : 0u);
      } else if (tail) {
        park(my_tail);
        if (g == 0) tkey[gi] = cur;
        my_flag |= 2u;
      } else {
        finish_column<T, TPR, NCV, OUT>(a, cur, acc, sa, sc, g, gmask);
      }
    }
    if (g == 0) flag[gi] = my_flag;
    __syncthreads();
    // ---- combine inside the unit: the group where a run starts owns it ----
    // which = 0: the run entering the unit from an earlier one (only group 0 can own it); which = 1: a run
    // that starts in this chunk (TAIL).
    for (int which = 0; which < 2; ++which) {
      const bool own = which == 0 ? (gi == 0 && (my_flag & 1u)) : (my_flag & 2u) != 0;
      if (!own) continue;
      const T *src = which == 0 ? my_head : my_tail;
      const uint32_t rkey = which == 0 ? hkey[gi] : tkey[gi];
#pragma unroll
      for (int ch = 0; ch < NCV; ++ch) {
        acc[ch].x = src[(ch * TPR + g) * 2];
        acc[ch].y = src[(ch * TPR + g) * 2 + 1];
      }
      sa = src[a.kp];
      sc = src[a.kp + 1];
      bool open = which == 0 ? (my_flag & 4u) != 0 : true;      // does the run continue past this chunk?
      int nx = gi + 1;
      while (open && nx < GPC) {
        const uint32_t f = flag[nx];
        if (!(f & 1u) || hkey[nx] != rkey) break;               // cannot happen for a consistent list
        const T *hp = part + ((size_t)nx * 2 + 0) * pstride;
#pragma unroll
        for (int ch = 0; ch < NCV; ++ch) {
          acc[ch].x += hp[(ch * TPR + g) * 2];
          acc[ch].y += hp[(ch * TPR + g) * 2 + 1];
        }
        sa += hp[a.kp];
        sc += hp[a.kp + 1];
        open = (f & 4u) != 0;
        ++nx;
      }
      if (which == 1 && !open) {
        finish_column<T, TPR, NCV, OUT>(a, rkey, acc, sa, sc, g, gmask);       // began and ended inside the unit
      } else {
        // crosses a unit boundary: one carry record for the whole unit (slot 0 = entering run, 1 = leaving run)
        V2 *cv = reinterpret_cast<V2 *>(a.carry_vec + ((size_t)unit * 2 + which) * a.kp) + g;
#pragma unroll
        for (int ch = 0; ch < NCV; ++ch) cv[ch * TPR] = acc[ch];
        if (g == 0) {
          a.carry_ac[((size_t)unit * 2 + which) * 2 + 0] = sa;
          a.carry_ac[((size_t)unit * 2 + which) * 2 + 1] = sc;
          if (which == 1) a.tails[atomicAdd(a.n_tails, 1u)] = unit;            // this unit owns the column's fix-up
        }
      }
    }
    __syncthreads();
  }
}

// Fix-up: one CTA per run that leaves a unit of the column pass (its TAIL record). Thread 0 finds the last
// unit the column reaches (one probe settles the common case, a binary search the rest); warp w then sums a
// contiguous block of the following units' HEAD records in unit order (loads issued FIX_UNROLL at a time,
// additions in order) and warp 0 adds TAIL + the warp blocks in warp order. The association depends only on
// the run length and the CTA shape, never on timing.
constexpr int FIX_THREADS = 256;
constexpr int FIX_WARPS = FIX_THREADS / 32;

template <typename T, int NCH, int OUT>
__global__ void __launch_bounds__(FIX_THREADS)
fm_fixup_kernel(const ColsArgs<T> a) {
  pdl_wait_and_release(a.pdl_release != 0);
  using V2 = typename Vec2<T>::type;
  constexpr int FIX_UNROLL = NCH <= 1 ? 8 : NCH <= 2 ? 4 : NCH <= 4 ? 2 : 1;
  extern __shared__ __align__(16) unsigned char fix_smem[];
  T *sm = reinterpret_cast<T *>(fix_smem);              // [FIX_WARPS][kp + 2]
  __shared__ uint32_t s_last, s_key;
  const int lane = lane_id(), wid = threadIdx.x >> 5;
  const uint32_t M = *a.count;
  const uint32_t n_tails = *a.n_tails;
  const int stride = a.kp + 2;
  for (uint32_t ti = blockIdx.x; ti < n_tails; ti += gridDim.x) {
    const uint32_t first = a.tails[ti];
    if (threadIdx.x == 0) {
      const uint32_t base = first * a.unit;
      const uint32_t klast = a.keys[base + a.unit - 1u];   // a unit with a leaving run is full; its last key is the column
      uint32_t last_unit = first + 1u;
      const uint32_t probe = base + 2u * a.unit - 1u;
      if (probe < M && a.keys[probe] == klast) {            // the next unit is entirely this column: search on
        // A unit belongs to the run iff its FIRST key is the column (the list is sorted), so the search runs over
        // units, galloping from the run's start: 2 log2(run length in units) dependent probes instead of log2(M)
        const uint32_t n_units = (M + a.unit - 1u) / a.unit;
        uint32_t known = first + 1u, step = 1u, hi = n_units;        // unit `known` is in the run
        while (known + step < n_units) {
          if (a.keys[(known + step) * a.unit] == klast) {
            known += step;
            step <<= 1;
          } else {
            hi = known + step;
            break;
          }
        }
        uint32_t lo = known + 1u;                                    // first unit not known to be in the run
        while (lo < hi) {
          const uint32_t mid = lo + ((hi - lo) >> 1);
          if (a.keys[mid * a.unit] == klast) lo = mid + 1; else hi = mid;
        }
        last_unit = lo - 1u;
      }
      s_last = last_unit;
      s_key = klast;
    }
    __syncthreads();
    const uint32_t last = s_last, klast = s_key;
    const uint32_t n_heads = last - first;
    const uint32_t per = (n_heads + FIX_WARPS - 1) / FIX_WARPS;
    {
      V2 acc[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) acc[ch].x = acc[ch].y = T(0);
      T sa = T(0), sc = T(0);
      const uint32_t c0 = first + 1 + wid * per;
      uint32_t c1 = c0 + per;
      if (c1 > last + 1) c1 = last + 1;
      for (uint32_t c = c0; c < c1; c += FIX_UNROLL) {
        V2 h[FIX_UNROLL][NCH];
        T ha[FIX_UNROLL], hc[FIX_UNROLL];
#pragma unroll
        for (int u = 0; u < FIX_UNROLL; ++u) {
          const bool on = c + u < c1;
          const size_t rec = (size_t)(on ? c + u : c) * 2;
          const V2 *hv = reinterpret_cast<const V2 *>(a.carry_vec + rec * a.kp) + lane;
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            h[u][ch] = __ldcg(hv + ch * 32);
            if (!on) h[u][ch].x = h[u][ch].y = T(0);
          }
          ha[u] = on ? __ldcg(a.carry_ac + rec * 2 + 0) : T(0);
          hc[u] = on ? __ldcg(a.carry_ac + rec * 2 + 1) : T(0);
        }
#pragma unroll
        for (int u = 0; u < FIX_UNROLL; ++u) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            acc[ch].x += h[u][ch].x;
            acc[ch].y += h[u][ch].y;
          }
          sa += ha[u];
          sc += hc[u];
        }
      }
      T *mine = sm + (size_t)wid * stride;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        mine[ch * 64 + lane * 2] = acc[ch].x;
        mine[ch * 64 + lane * 2 + 1] = acc[ch].y;
      }
      if (lane == 0) {
        mine[a.kp] = sa;
        mine[a.kp + 1] = sc;
      }
    }
    __syncthreads();
    if (wid == 0) {
      V2 acc[NCH];
      const V2 *tv = reinterpret_cast<const V2 *>(a.carry_vec + ((size_t)first * 2 + 1) * a.kp) + lane;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) acc[ch] = __ldcg(tv + ch * 32);
      T sa = __ldcg(a.carry_ac + ((size_t)first * 2 + 1) * 2 + 0), sc = __ldcg(a.carry_ac + ((size_t)first * 2 + 1) * 2 + 1);
      for (int w = 0; w < FIX_WARPS; ++w) {
        const T *part = sm + (size_t)w * stride;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          acc[ch].x += part[ch * 64 + lane * 2];
          acc[ch].y += part[ch * 64 + lane * 2 + 1];
        }
        sa += part[a.kp];
        sc += part[a.kp + 1];
      }
      finish_column<T, 32, NCH, OUT>(a, klast, acc, sa, sc, lane, FULL);
    }
    __syncthreads();
  }
}

// dense apply of an all-reduced gradient (data-parallel mode), identical on every rank: one warp per
// feature row updates v_j, w_j and ||v_j||^2 in a single pass; the first thread also updates w0
template <typename T>
__global__ void __launch_bounds__(ROWS_THREADS)
apply_grad_rows_kernel(T *__restrict__ w0, T *__restrict__ w, T *__restrict__ V, T *__restrict__ vn,
                       const T *__restrict__ g0, const T *__restrict__ gw, const T *__restrict__ gV, int64_t n,
                       int kp, T lr) {
  const int lane = lane_id();
  const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (gwarp == 0 && lane == 0) *w0 += lr * *g0;
  for (int64_t j = gwarp; j < n; j += nw) {
    T s = T(0);
    for (int f = lane * 2; f < kp; f += 64) {      // same association as row_norms_kernel
      const T v0 = V[j * kp + f] + lr * gV[j * kp + f];
      const T v1 = V[j * kp + f + 1] + lr * gV[j * kp + f + 1];
      V[j * kp + f] = v0;
      V[j * kp + f + 1] = v1;
      s += v0 * v0 + v1 * v1;
    }
    s = warp_sum(s);
    if (lane == 0) {
      vn[j] = s;
      w[j] += lr * gw[j];
    }
  }
}

// Dense optimizer step on a gradient buffer (rfm_fm_train_epoch_opt). The buffer holds the ASCENT direction d of
// the reference's closed form (theta += lr d is its SGD step), so the loss gradient is g = -d + l2 theta.
// ADAM (Kingma & Ba, bias-corrected; c1 = 1 / (1 - beta1^t), c2 = 1 / (1 - beta2^t) computed on the host):
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  theta -= lr (m c1) / (sqrt(v c2) + eps)
// otherwise SGD with L2: theta -= lr g. Specification: oracle/optimizer_oracle.py.
template <typename T>
struct OptArgs {
  T lr, l2, b1, b2, eps, c1, c2;
};
template <typename T, bool ADAM>
__device__ __forceinline__ T opt_update(T theta, T d, T *__restrict__ m, T *__restrict__ v, const OptArgs<T> &o) {
  const T g = o.l2 * theta - d;
  if (!ADAM) return theta - o.lr * g;
  const T mm = o.b1 * *m + (T(1) - o.b1) * g;
  const T vv = o.b2 * *v + (T(1) - o.b2) * (g * g);
  *m = mm;
  *v = vv;
  return theta - o.lr * (mm * o.c1) / (sqrt(vv * o.c2) + o.eps);
}
template <typename T, bool ADAM>
__global__ void __launch_bounds__(ROWS_THREADS)
optimizer_rows_kernel(T *__restrict__ w0, T *__restrict__ w, T *__restrict__ V, T *__restrict__ vn,
                      const T *__restrict__ grad, T *__restrict__ am, T *__restrict__ av, int64_t n, int k, int kp,
                      int64_t w_off, int64_t v_off, const OptArgs<T> o) {
  const int lane = lane_id();
  const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (gwarp == 0 && lane == 0) *w0 = opt_update<T, ADAM>(*w0, grad[0], am, av, o);
  for (int64_t j = gwarp; j < n; j += nw) {
    T s = T(0);
    for (int f = lane * 2; f < kp; f += 64) {      // same association as row_norms_kernel
      const int64_t at = v_off + j * kp + f;
      // padding columns (f >= k) stay exactly zero: no gradient, and Adam's 0 / (0 + eps) is skipped outright
      const T v0 = f < k ? opt_update<T, ADAM>(V[j * kp + f], grad[at], am + at, av + at, o) : T(0);
      const T v1 = f + 1 < k ? opt_update<T, ADAM>(V[j * kp + f + 1], grad[at + 1], am + at + 1, av + at + 1, o) : T(0);
      V[j * kp + f] = v0;
      V[j * kp + f + 1] = v1;
      s += v0 * v0 + v1 * v1;
    }
    s = warp_sum(s);
    if (lane == 0) {
      vn[j] = s;
      w[j] = opt_update<T, ADAM>(w[j], grad[w_off + j], am + w_off + j, av + w_off + j, o);
    }
  }
}

// row / column kernels: (threads per row, chunks per lane) for kp = 64 * nch
#define RFM_DISPATCH_TPR(nch, ...)                                                       \
  switch (nch) {                                                                         \
    case 1: { constexpr int TPR = 8, NCV = 4; __VA_ARGS__; } break;                      \
    case 2: { constexpr int TPR = 16, NCV = 4; __VA_ARGS__; } break;                     \
    case 3: { constexpr int TPR = 32, NCV = 3; __VA_ARGS__; } break;                     \
    case 4: { constexpr int TPR = 32, NCV = 4; __VA_ARGS__; } break;                     \
    case 5: { constexpr int TPR = 32, NCV = 5; __VA_ARGS__; } break;                     \
    case 6: { constexpr int TPR = 32, NCV = 6; __VA_ARGS__; } break;                     \
    case 7: { constexpr int TPR = 32, NCV = 7; __VA_ARGS__; } break;                     \
    case 8: { constexpr int TPR = 32, NCV = 8; __VA_ARGS__; } break;                     \
    default: return fail(RFM_ERR_INVALID, "n_factors too large (kp/64 = %d > 8)", nch);  \
  }

#define RFM_DISPATCH_NCH(nch, ...)                                 \
  switch (nch) {                                                   \
    case 1: { constexpr int NCH = 1; __VA_ARGS__; } break;         \
    case 2: { constexpr int NCH = 2; __VA_ARGS__; } break;         \
    case 3: { constexpr int NCH = 3; __VA_ARGS__; } break;         \
    case 4: { constexpr int NCH = 4; __VA_ARGS__; } break;         \
    case 5: { constexpr int NCH = 5; __VA_ARGS__; } break;         \
    case 6: { constexpr int NCH = 6; __VA_ARGS__; } break;         \
    case 7: { constexpr int NCH = 7; __VA_ARGS__; } break;         \
    case 8: { constexpr int NCH = 8; __VA_ARGS__; } break;         \
    default: return fail(RFM_ERR_INVALID, "n_factors too large (kp/64 = %d > 8)", nch); \
  }

int grid_for(rfm_ctx *ctx, int64_t work_items_per_block_unit, int blocks_per_sm) {
  const int64_t cap = (int64_t)ctx->sm_count * blocks_per_sm;
  int64_t g = work_items_per_block_unit < cap ? work_items_per_block_unit : cap;
  return g < 1 ? 1 : (int)g;
}

// copies a host array to the device through cudaMemcpyAsync on the context stream
int upload(rfm_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
  if (bytes == 0) return RFM_OK;
  RFM_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return RFM_OK;
}

template <typename T, bool FAC>
int launch_rows_as(rfm_ctx *ctx, int nch, int mode, bool sampled, const RowsArgs<T> &args, int grid) {
  RFM_DISPATCH_TPR(nch, {
    if (mode == MODE_TRAIN) {
      if (sampled) {
        auto fm_rows_train = fm_rows_kernel<T, TPR, NCV, MODE_TRAIN, true, FAC>;
        RFM_LAUNCH_PDL(ctx, fm_rows_train, grid, ROWS_THREADS, 0, args);
      } else {
        auto fm_rows_train = fm_rows_kernel<T, TPR, NCV, MODE_TRAIN, false, FAC>;
        RFM_LAUNCH_PDL(ctx, fm_rows_train, grid, ROWS_THREADS, 0, args);
      }
    } else if (mode == MODE_LOSS) {
      auto fm_rows_loss = fm_rows_kernel<T, TPR, NCV, MODE_LOSS, false, FAC>;
      RFM_LAUNCH_PDL(ctx, fm_rows_loss, grid, ROWS_THREADS, 0, args);
    } else {
      auto fm_rows_predict = fm_rows_kernel<T, TPR, NCV, MODE_PREDICT, false, FAC>;
      RFM_LAUNCH_PDL(ctx, fm_rows_predict, grid, ROWS_THREADS, 0, args);
    }
  });
  return RFM_OK;
}
// fac: the rows (both sets of a loss launch) are factored
template <typename T>
int launch_rows(rfm_ctx *ctx, int nch, int mode, bool sampled, const RowsArgs<T> &args, int grid, bool fac = false) {
  return fac ? launch_rows_as<T, true>(ctx, nch, mode, sampled, args, grid)
             : launch_rows_as<T, false>(ctx, nch, mode, sampled, args, grid);
}

// rows (or 32-entry chunks) a CTA of ROWS_THREADS handles per pass of its loop
int units_per_block(int nch) { return ROWS_WARPS * (nch == 1 ? 4 : nch == 2 ? 2 : 1); }
// sorted entries one CTA iteration of the column pass covers (= entries per carry record)
int64_t cols_unit(int nch) { return 32 * (int64_t)units_per_block(nch); }
size_t cols_smem_bytes(int nch, int kp, size_t es) {
  const size_t gpc = (size_t)units_per_block(nch);
  return gpc * 2 * (size_t)(kp + 2) * es + 3 * gpc * sizeof(uint32_t);
}

#include "two_level.cuh"

}  // namespace

// ---- trainer handle -----------------------------------------------------------------------------
struct rfm_fm_trainer {
  rfm_fm *m = nullptr;
  const rfm_csr *train = nullptr, *val = nullptr;
  int64_t max_batch = 0, max_slots = 0, nnz_cap = 0;
  int rows_grid = 0;
  uint32_t stride = 0;        // > 0: fixed-stride triple layout (rows of near-uniform length)
  uint32_t count_host = 0;    // value currently stored in count (fixed-stride layout)
  DevBuf<int64_t> idx;
  DevBuf<uint32_t> row_len, bptr, scan_tmp, count, tails, n_tails, ticket;
  DevBuf<unsigned char> S, E, carry_vec, carry_ac, grad;
  DevBuf<double> block_partials, losses, loss_sums;
  // NVLink peer exchange of the data-parallel step (rfm_fm_dp_*): a cudaMalloc'ed region every rank maps
  // through CUDA IPC: [gradient buffer, parity 0 | gradient buffer, parity 1 | flags]
  static constexpr int DP_MAX_WORLD = 8;
  unsigned char *xchg = nullptr;
  bool xchg_borrowed = false;   // xchg and peer_base[] belong to the context's DpRegion cache
  size_t xchg_grad_bytes = 0;
  unsigned char *peer_base[DP_MAX_WORLD] = {nullptr};
  int dp_rank = -1, dp_world = 0, dp_parity = 0;
  uint32_t dp_seq = 0;
  bool dp_pending_loss = false;
  DevBuf<unsigned char> adam_m, adam_v;   // Adam moments, laid out like the gradient buffer
  DevBuf<double> dp_prev_loss;
  DevBuf<uint32_t> dp_local;   // [arrive counter, go flag, error flag]
  unsigned char *grad_ptr() const { return xchg ? xchg + (size_t)dp_parity * xchg_grad_bytes : grad.p; }
  RadixSorter<float> sort32;
  RadixSorter<double> sort64;
  TwoLevel *tl = nullptr;      // two-level step (factored rows; rfm_fm_trainer_set_two_level)
  bool broken = false;         // a failed rfm_fm_trainer_set_two_level left the batch buffers half re-sized
  // host staging ring for batch row ids
  static constexpr int RING = 4;
  PinnedBuf<int64_t> stage[RING];
  cudaEvent_t stage_ev[RING] = {nullptr, nullptr, nullptr, nullptr};
  bool stage_used[RING] = {false, false, false, false};
  int ring_pos = 0;
};

namespace {

template <typename T>
RadixSorter<T> &sorter_of(rfm_fm_trainer *t);
template <>
RadixSorter<float> &sorter_of<float>(rfm_fm_trainer *t) { return t->sort32; }
template <>
RadixSorter<double> &sorter_of<double>(rfm_fm_trainer *t) { return t->sort64; }

template <typename T>
RowsArgs<T> rows_args(const rfm_fm *m, const rfm_csr *rows) {
  RowsArgs<T> a;
  memset(&a, 0, sizeof(a));
  a.row_ptr = rows->row_ptr.p;
  a.col = rows->col.p;
  a.val = reinterpret_cast<const T *>(rows->val.p);
  a.yp = reinterpret_cast<const T *>(rows->yp.p);
  a.w0 = reinterpret_cast<const T *>(m->w0.p);
  a.w = reinterpret_cast<const T *>(m->w.p);
  a.V = reinterpret_cast<const T *>(m->V.p);
  a.vn = reinterpret_cast<const T *>(m->vn.p);
  a.kp = m->kp;
  if (rows->factored) a.fac = rows->fac_dev();
  return a;
}

Finish make_finish(int op, double scale, void *dst_T, double *dst_d, double *block_partials, uint32_t *ticket) {
  Finish f;
  f.op = op;
  f.scale = scale;
  f.dst_T = dst_T;
  f.dst_d = dst_d;
  f.block_partials = block_partials;
  f.ticket = ticket;
  return f;
}

// loss pass over a batch (idx != null) or a row range, mean (scale = 1/count) or raw sum into *dst
template <typename T>
int loss_pass(rfm_fm_trainer *t, const rfm_csr *rows, const int64_t *idx_dev, int64_t row0, int64_t n,
              double scale, double *dst_dev) {
  rfm_fm *m = t->m;
  rfm_ctx *ctx = m->ctx;
  if (n <= 0) {
    RFM_CUDA(cudaMemsetAsync(dst_dev, 0, sizeof(double), ctx->stream));
    return RFM_OK;
  }
  RowsArgs<T> a = rows_args<T>(m, rows);
  a.idx = idx_dev;
  a.row0 = row0;
  a.n = n;
  a.fin = make_finish(1, scale, nullptr, dst_dev, t->block_partials.p, t->ticket.p);
  const int grid = grid_for(ctx, ceil_div(n, units_per_block(m->nch)), t->rows_grid / ctx->sm_count);
  return launch_rows<T>(ctx, m->nch, MODE_LOSS, false, a, grid, rows->factored);
}

// forward + residual + triples + sort + column pass (+ fix-up). DP=false applies SGD in place.
// sampled: the batch is positions [q0, q0+batch) of the epoch's Feistel permutation, drawn on the device.
template <typename T, bool DP>
int step_two_level(rfm_fm_trainer *t, int64_t batch, double lr, bool sampled, const FeistelKey &fkey, int64_t q0);

template <typename T, bool DP>
int step_core(rfm_fm_trainer *t, int64_t batch, double lr, bool sampled, const FeistelKey &fkey, int64_t q0) {
  if (t->tl) return step_two_level<T, DP>(t, batch, lr, sampled, fkey, q0);
  rfm_fm *m = t->m;
  rfm_ctx *ctx = m->ctx;
  const rfm_csr *tr = t->train;
  RadixSorter<T> &sorter = sorter_of<T>(t);
  const bool fused_draw = sampled && t->stride > 0;
  if (sampled && !fused_draw)   // ragged layout needs the row ids before the row pass (row lengths)
    RFM_LAUNCH(ctx, feistel_sample_kernel, grid_for(ctx, ceil_div(batch, 256), 4), 256, 0, fkey, q0, batch,
               t->idx.p);
  if (t->stride == 0) {
    const int small_grid = grid_for(ctx, ceil_div(batch, 256), 4);
    if (tr->factored) {
      RFM_LAUNCH(ctx, fac_row_len_kernel, small_grid, 256, 0, tr->fac_dev(), t->idx.p, batch, t->row_len.p);
    } else {
      RFM_LAUNCH(ctx, row_len_kernel, small_grid, 256, 0, tr->row_ptr.p, t->idx.p, batch, t->row_len.p);
    }
    RFM_TRY(exclusive_scan_u32(ctx, t->row_len.p, t->bptr.p, batch, t->scan_tmp.p, t->count.p));
  } else {
    const uint32_t cnt = (uint32_t)batch * t->stride;
    if (cnt != t->count_host) {
      t->count_host = cnt;
      RFM_CUDA(cudaMemcpyAsync(t->count.p, &t->count_host, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
  }

  RowsArgs<T> a = rows_args<T>(m, tr);
  a.idx = t->idx.p;
  a.n = batch;
  a.S = reinterpret_cast<T *>(t->S.p);
  a.E = reinterpret_cast<T *>(t->E.p);
  a.bptr = t->bptr.p;
  a.stride = t->stride;
  a.sentinel = (uint32_t)m->n;
  a.keys = sorter.keys[0].p;
  a.pos = sorter.pos[0].p;
  a.xs = sorter.val[0].p;
  a.fkey = fkey;
  a.q0 = q0;
  a.idx_out = t->idx.p;
  RFM_TRY(sorter.clear_histograms(ctx));
  // every memset of the step happens before the row pass, so that its kernels follow each other in the stream
  // (RFM_LAUNCH_PDL)
  RFM_TRY(sorter.prepare(ctx));
  RFM_CUDA(cudaMemsetAsync(t->n_tails.p, 0, sizeof(uint32_t), ctx->stream));
  a.ghist = sorter.ghist();
  a.n_passes = sorter.passes;
  if (DP) {
    T *g = reinterpret_cast<T *>(t->grad_ptr());
    RFM_CUDA(cudaMemsetAsync(g, 0, (size_t)grad_total(m->n, m->kp) * sizeof(T), ctx->stream));
    // sum_e lands in grad[0]; w0 itself is updated by the dense apply after the all-reduce
    a.fin = make_finish(0, 1.0, g, nullptr, t->block_partials.p, t->ticket.p);
  } else {
    a.fin = make_finish(0, lr, m->w0.p, nullptr, t->block_partials.p, t->ticket.p);
  }
  const int grid = grid_for(ctx, ceil_div(batch, units_per_block(m->nch)), t->rows_grid / ctx->sm_count);
  a.pdl_release = pdl_on(ctx) ? 1 : 0;       // the sort passes, the column pass and the fix-up follow back to back, each waiting
  RFM_TRY(launch_rows<T>(ctx, m->nch, MODE_TRAIN, fused_draw, a, grid, tr->factored));
  int sorted = 0;
  RFM_TRY(sorter.sort(ctx, t->count.p, &sorted, /*histograms_ready=*/true, /*prepared=*/true));

  ColsArgs<T> c;
  memset(&c, 0, sizeof(c));
  c.keys = sorter.keys[sorted].p;
  c.pos = sorter.pos[sorted].p;
  c.xs = sorter.val[sorted].p;
  c.count = t->count.p;
  c.sentinel = t->stride ? (uint32_t)m->n : KEY_NONE;
  c.E = reinterpret_cast<const T *>(t->E.p);
  c.S = reinterpret_cast<const T *>(t->S.p);
  c.V = reinterpret_cast<T *>(m->V.p);
  c.w = reinterpret_cast<T *>(m->w.p);
  c.vn = reinterpret_cast<T *>(m->vn.p);
  c.kp = m->kp;
  c.lr = static_cast<T>(lr);
  c.unit = (uint32_t)cols_unit(m->nch);
  c.carry_vec = reinterpret_cast<T *>(t->carry_vec.p);
  c.carry_ac = reinterpret_cast<T *>(t->carry_ac.p);
  c.tails = t->tails.p;
  c.n_tails = t->n_tails.p;
  if (DP) {
    T *g = reinterpret_cast<T *>(t->grad_ptr());
    c.grad_w = g + GRAD_W_OFF;
    c.grad_V = g + grad_v_off(m->n);
  }
  const int64_t unit_cap = ceil_div(t->nnz_cap, cols_unit(m->nch));
  const int cgrid = grid_for(ctx, unit_cap, t->rows_grid / ctx->sm_count);
  const int tgrid = grid_for(ctx, unit_cap, 4);   // one CTA per leaving run; there are at most unit_cap of them
  const size_t csmem = cols_smem_bytes(m->nch, m->kp, sizeof(T));
  RFM_DISPATCH_TPR(m->nch, {
    auto fm_cols = fm_cols_kernel<T, TPR, NCV, DP ? OUT_GRAD : OUT_SGD, false>;
    if (csmem > 48 * 1024)
      RFM_CUDA(cudaFuncSetAttribute(fm_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
    c.pdl_release = pdl_on(ctx) ? 1 : 0;
    RFM_LAUNCH_PDL(ctx, fm_cols, cgrid, ROWS_THREADS, csmem, c);
    c.pdl_release = 0;     // what follows the fix-up is the caller's business (exchange, optimizer, loss pass, memsets)
  });
  RFM_DISPATCH_NCH(m->nch, {
    auto fm_fixup = fm_fixup_kernel<T, NCH, DP ? OUT_GRAD : OUT_SGD>;
    const size_t fsmem = (size_t)FIX_WARPS * (m->kp + 2) * sizeof(T);
    RFM_LAUNCH_PDL(ctx, fm_fixup, tgrid, FIX_THREADS, fsmem, c);
  });
  if (!DP) ++m->version;
  return RFM_OK;
}

// ---- the two-level step (two_level.cuh) -------------------------------------------------------------------------
template <typename T, int NCTX>
int launch_vrows_as(rfm_ctx *ctx, int nch, int mode, bool sampled, const VRowsArgs<T> &args, int grid) {
  RFM_DISPATCH_TPR(nch, {
    if (mode == MODE_TRAIN) {
      if (sampled) {
        auto fm_vrows_train = fm_vrows_kernel<T, TPR, NCV, MODE_TRAIN, true, NCTX>;
        RFM_LAUNCH_PDL(ctx, fm_vrows_train, grid, ROWS_THREADS, 0, args);
      } else {
        auto fm_vrows_train = fm_vrows_kernel<T, TPR, NCV, MODE_TRAIN, false, NCTX>;
        RFM_LAUNCH_PDL(ctx, fm_vrows_train, grid, ROWS_THREADS, 0, args);
      }
    } else {
      auto fm_vrows_loss = fm_vrows_kernel<T, TPR, NCV, MODE_LOSS, false, NCTX>;
      RFM_LAUNCH_PDL(ctx, fm_vrows_loss, grid, ROWS_THREADS, 0, args);
    }
  });
  return RFM_OK;
}

template <typename T>
struct TlParams {
  T *Vv, *wv, *vnv, *R, *Ra, *Rc;
};
template <typename T>
TlParams<T> tl_params(const rfm_fm_trainer *t) {
  const TwoLevel &L = *t->tl;
  TlParams<T> p;
  p.Vv = reinterpret_cast<T *>(L.Vv.p);
  p.wv = reinterpret_cast<T *>(L.wv.p);
  p.vnv = reinterpret_cast<T *>(L.vnv.p);
  p.R = reinterpret_cast<T *>(L.R.p);
  p.Ra = p.R + (size_t)L.nv * t->m->kp;
  p.Rc = p.Ra + L.nv;
  return p;
}

// per-entity aggregates of the CURRENT parameters (no-op while nothing changed them since the last call)
// release_next: the caller launches a waiting kernel right after this call, with nothing in between
template <typename T>
int tl_refresh(rfm_fm_trainer *t, bool release_next = false) {
  TwoLevel &L = *t->tl;
  rfm_fm *m = t->m;
  if (L.agg_version == m->version) return RFM_OK;
  rfm_ctx *ctx = m->ctx;
  const TlParams<T> p = tl_params<T>(t);
  const int grid = grid_for(ctx, ceil_div(L.nv, units_per_block(m->nch)), 4);     // one wave
  RFM_DISPATCH_TPR(m->nch, {
    auto fm_entity_fwd = fm_entity_fwd_kernel<T, TPR, NCV>;
    RFM_LAUNCH_PDL(ctx, fm_entity_fwd, grid, ROWS_THREADS, 0, (const uint32_t *)L.ent_ptr.p, (const int32_t *)L.ent_col.p,
                   reinterpret_cast<const T *>(L.ent_val.p), L.nv, reinterpret_cast<const T *>(m->V.p),
                   reinterpret_cast<const T *>(m->w.p), reinterpret_cast<const T *>(m->vn.p), m->kp, p.Vv, p.wv, p.vnv,
                   release_next && pdl_on(ctx) ? 1 : 0);
  });
  L.agg_version = m->version;
  return RFM_OK;
}

template <typename T>
RadixSorter<T> &tl_sorter_of(TwoLevel &L);
template <>
RadixSorter<float> &tl_sorter_of<float>(TwoLevel &L) { return L.l2_32; }
template <>
RadixSorter<double> &tl_sorter_of<double>(TwoLevel &L) { return L.l2_64; }

template <typename T, bool DP>
int step_two_level(rfm_fm_trainer *t, int64_t batch, double lr, bool sampled, const FeistelKey &fkey, int64_t q0) {
  rfm_fm *m = t->m;
  rfm_ctx *ctx = m->ctx;
  TwoLevel &L = *t->tl;
  const rfm_csr *tr = t->train;
  RadixSorter<T> &sorter = sorter_of<T>(t);
  const TlParams<T> p = tl_params<T>(t);
  RFM_TRY(tl_refresh<T>(t));
  const uint32_t cnt = (uint32_t)batch * t->stride;
  if (cnt != t->count_host) {
    t->count_host = cnt;
    RFM_CUDA(cudaMemcpyAsync(t->count.p, &t->count_host, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  }
  RFM_CUDA(cudaMemsetAsync(L.R.p, 0, (size_t)L.nv * (m->kp + 2) * sizeof(T), ctx->stream));
  // every memset of the step happens here, so that its nine kernels follow each other in the stream (RFM_LAUNCH_PDL)
  RFM_TRY(sorter.prepare(ctx));
  RFM_CUDA(cudaMemsetAsync(L.n_tails2.p, 0, 2 * sizeof(uint32_t), ctx->stream));

  // level 1: the row pass on virtual rows, with the aggregated table standing in for the parameters
  const int grid = grid_for(ctx, ceil_div(batch, units_per_block(m->nch)), L.lean ? TL_VROWS_BLOCKS : t->rows_grid / ctx->sm_count);
  RFM_TRY(sorter.clear_histograms(ctx));
  Finish fin;
  if (DP) {
    T *g = reinterpret_cast<T *>(t->grad_ptr());
    RFM_CUDA(cudaMemsetAsync(g, 0, (size_t)grad_total(m->n, m->kp) * sizeof(T), ctx->stream));
    fin = make_finish(0, 1.0, g, nullptr, t->block_partials.p, t->ticket.p);
  } else {
    fin = make_finish(0, lr, m->w0.p, nullptr, t->block_partials.p, t->ticket.p);
  }
  if (L.lean) {
    VRowsArgs<T> v;
    memset(&v, 0, sizeof(v));
    v.user = tr->f_user.p;
    v.item = tr->f_item.p;
    v.ctx = reinterpret_cast<const T *>(tr->f_ctx.p);
    v.yp = reinterpret_cast<const T *>(tr->yp.p);
    v.idx = t->idx.p;
    v.n = batch;
    v.w0 = reinterpret_cast<const T *>(m->w0.p);
    v.A = p.Vv;
    v.wv = p.wv;
    v.vnv = p.vnv;
    v.n_user = (uint32_t)L.n_ent[0];
    v.ctx_col = (uint32_t)(L.n_ent[0] + L.n_ent[1]);
    v.kp = m->kp;
    v.S = reinterpret_cast<T *>(t->S.p);
    v.E = reinterpret_cast<T *>(t->E.p);
    v.stride = t->stride;
    v.sentinel = (uint32_t)L.nv;
    v.keys = sorter.keys[0].p;
    v.pos = sorter.pos[0].p;
    v.xs = sorter.val[0].p;
    v.ghist = sorter.ghist();
    v.n_passes = sorter.passes;
    v.fkey = fkey;
    v.q0 = q0;
    v.idx_out = t->idx.p;
    v.fin = fin;
    v.pdl_release = pdl_on(ctx) ? 1 : 0;     // sort, level 1, level 2 follow back to back, each waiting
    v.Cq = reinterpret_cast<T *>(L.Cq.p);
    if (tr->n_ctx) RFM_TRY((launch_vrows_as<T, 1>(ctx, m->nch, MODE_TRAIN, sampled, v, grid)));
    else RFM_TRY((launch_vrows_as<T, 0>(ctx, m->nch, MODE_TRAIN, sampled, v, grid)));
  } else {
  RowsArgs<T> a = rows_args<T>(m, tr);
  a.fac = tl_virtual_fac(tr->fac_dev(), L);
  a.V = p.Vv;
  a.w = p.wv;
  a.vn = p.vnv;
  a.idx = t->idx.p;
  a.n = batch;
  a.S = reinterpret_cast<T *>(t->S.p);
  a.E = reinterpret_cast<T *>(t->E.p);
  a.bptr = t->bptr.p;
  a.stride = t->stride;
  a.sentinel = (uint32_t)L.nv;
  a.keys = sorter.keys[0].p;
  a.pos = sorter.pos[0].p;
  a.xs = sorter.val[0].p;
  a.fkey = fkey;
  a.q0 = q0;
  a.idx_out = t->idx.p;
  a.ghist = sorter.ghist();
  a.n_passes = sorter.passes;
  a.fin = fin;
  a.pdl_release = pdl_on(ctx) ? 1 : 0;
  RFM_TRY(launch_rows<T>(ctx, m->nch, MODE_TRAIN, sampled, a, grid, true));
  }
  int sorted = 0;
  RFM_TRY(sorter.sort(ctx, t->count.p, &sorted, /*histograms_ready=*/true, /*prepared=*/true));

  ColsArgs<T> c;
  memset(&c, 0, sizeof(c));
  c.keys = sorter.keys[sorted].p;
  c.pos = sorter.pos[sorted].p;
  c.xs = sorter.val[sorted].p;
  c.count = t->count.p;
  c.sentinel = (uint32_t)L.nv;
  c.E = reinterpret_cast<const T *>(t->E.p);
  c.S = reinterpret_cast<const T *>(t->S.p);
  c.kp = m->kp;
  c.lr = static_cast<T>(lr);
  c.unit = (uint32_t)cols_unit(m->nch);
  c.carry_vec = reinterpret_cast<T *>(t->carry_vec.p);
  c.carry_ac = reinterpret_cast<T *>(t->carry_ac.p);
  c.tails = t->tails.p;
  c.n_tails = L.n_tails2.p;
  c.grad_V = p.R;
  c.grad_w = p.Ra;
  c.raw_c = p.Rc;
  const size_t csmem = cols_smem_bytes(m->nch, m->kp, sizeof(T));
  const size_t fsmem = (size_t)FIX_WARPS * (m->kp + 2) * sizeof(T);
  // lists too short to fill the machine with 32-entry chunks are walked in chunks of TPR entries (unit = 256)
  auto small_chunks = [&](int64_t entries) { return ceil_div(entries, cols_unit(m->nch)) < 2 * (int64_t)ctx->sm_count; };
  {
    const bool small = small_chunks(cnt);
    c.unit = small ? (uint32_t)ROWS_THREADS : (uint32_t)cols_unit(m->nch);
    const int64_t unit_cap = ceil_div((int64_t)cnt, (int64_t)c.unit);
    if (L.lean && tr->n_ctx == 1) {      // the dense context column is summed by extra CTAs of this launch
      c.ctx_blocks = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count, batch / 256));
      c.ctx_col = (uint32_t)(L.n_ent[0] + L.n_ent[1]);
      c.ctx_rows = batch;
      c.Cq = reinterpret_cast<const T *>(L.Cq.p);
      c.cpart = reinterpret_cast<T *>(L.cpart.p);
      c.ctx_ticket = L.ctx_ticket.p;
    }
    const int cgrid = grid_for(ctx, unit_cap, t->rows_grid / ctx->sm_count) + (int)c.ctx_blocks;
    const int tgrid = grid_for(ctx, unit_cap, 4);
    RFM_DISPATCH_TPR(m->nch, {
      auto fm_cols_level1 = small ? fm_cols_kernel<T, TPR, NCV, OUT_RAW, false, true>
                                  : fm_cols_kernel<T, TPR, NCV, OUT_RAW, false, false>;
      if (csmem > 48 * 1024)
        RFM_CUDA(cudaFuncSetAttribute(fm_cols_level1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      c.pdl_release = pdl_on(ctx) ? 1 : 0;
      RFM_LAUNCH_PDL(ctx, fm_cols_level1, cgrid, ROWS_THREADS, csmem, c);
    });
    RFM_DISPATCH_NCH(m->nch, {
      auto fm_fixup_level1 = fm_fixup_kernel<T, NCH, OUT_RAW>;
      RFM_LAUNCH_PDL(ctx, fm_fixup_level1, tgrid, FIX_THREADS, fsmem, c);
    });
  }

  // level 2: the static (real column, virtual column, x) list over level 1's sums -> the reference's update
  RadixSorter<T> &l2 = tl_sorter_of<T>(L);
  ColsArgs<T> d = c;
  d.ctx_blocks = 0;
  d.keys = l2.keys[L.l2_buf].p;
  d.pos = l2.pos[L.l2_buf].p;
  d.xs = l2.val[L.l2_buf].p;
  d.count = L.m2_dev.p;
  d.sentinel = KEY_NONE;
  d.E = nullptr;
  d.S = p.R;
  d.Ea = p.Ra;
  d.Ec = p.Rc;
  d.V = reinterpret_cast<T *>(m->V.p);
  d.w = reinterpret_cast<T *>(m->w.p);
  d.vn = reinterpret_cast<T *>(m->vn.p);
  d.grad_V = d.grad_w = d.raw_c = nullptr;
  if (DP) {
    T *g = reinterpret_cast<T *>(t->grad_ptr());
    d.grad_w = g + GRAD_W_OFF;
    d.grad_V = g + grad_v_off(m->n);
  }
  d.n_tails = L.n_tails2.p + 1;
  {
    const bool small = small_chunks(L.m2);
    d.unit = small ? (uint32_t)ROWS_THREADS : (uint32_t)cols_unit(m->nch);
    const int64_t unit_cap = ceil_div(L.m2, (int64_t)d.unit);
    const int cgrid = grid_for(ctx, unit_cap, t->rows_grid / ctx->sm_count);
    const int tgrid = grid_for(ctx, unit_cap, 4);
    RFM_DISPATCH_TPR(m->nch, {
      auto fm_cols_level2 = small ? fm_cols_kernel<T, TPR, NCV, DP ? OUT_GRAD : OUT_SGD, true, true>
                                  : fm_cols_kernel<T, TPR, NCV, DP ? OUT_GRAD : OUT_SGD, true, false>;
      if (csmem > 48 * 1024)
        RFM_CUDA(cudaFuncSetAttribute(fm_cols_level2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      d.pdl_release = pdl_on(ctx) ? 1 : 0;
      RFM_LAUNCH_PDL(ctx, fm_cols_level2, cgrid, ROWS_THREADS, csmem, d);
      d.pdl_release = 0;   // what follows the last fix-up is the caller's business
    });
    RFM_DISPATCH_NCH(m->nch, {
      auto fm_fixup_level2 = fm_fixup_kernel<T, NCH, DP ? OUT_GRAD : OUT_SGD>;
      RFM_LAUNCH_PDL(ctx, fm_fixup_level2, tgrid, FIX_THREADS, fsmem, d);
    });
  }
  if (!DP) ++m->version;
  return RFM_OK;
}

template <typename T>
int batch_and_val_losses(rfm_fm_trainer *t, int64_t batch, double scale_b, double *dst_b, int64_t val_begin,
                         int64_t val_end, double scale_v, double *dst_v);
template <typename T>
int post_update_losses(rfm_fm_trainer *t, int64_t batch, int64_t slot) {
  const int64_t nv = t->val ? t->val->n_rows : 0;
  return batch_and_val_losses<T>(t, batch, 1.0 / (double)batch, t->losses.p + slot, 0, nv,
                                 nv > 0 ? 1.0 / (double)nv : 1.0, t->losses.p + t->max_slots + slot);
}

template <typename T>
int epoch_impl(rfm_fm_trainer *t, int64_t batch, double lr, int64_t slot, bool sampled, const FeistelKey &fkey) {
  RFM_TRY((step_core<T, false>(t, batch, lr, sampled, fkey, 0)));
  return post_update_losses<T>(t, batch, slot);
}

// loss of the batch on the device (t->idx) and of val rows [val_begin, val_end) in ONE launch (the val rows ride
// along as a second row set); each result is scale * sum of the per-row terms
template <typename T>
int batch_and_val_losses(rfm_fm_trainer *t, int64_t batch, double scale_b, double *dst_b, int64_t val_begin,
                         int64_t val_end, double scale_v, double *dst_v) {
  rfm_fm *m = t->m;
  rfm_ctx *ctx = m->ctx;
  if (t->tl) {     // the batch (and val rows keyed by the same tables) as virtual rows over the aggregated table
    TwoLevel &L = *t->tl;
    // the entity forward may release the loss launch early only when nothing sits between them in the stream
    const bool direct = t->val && val_end > val_begin;      // no memset of dst_v before the loss launch
    RFM_TRY(tl_refresh<T>(t, /*release_next=*/direct));
    const TlParams<T> p = tl_params<T>(t);
    if (L.lean) {
      const rfm_csr *tr = t->train, *va = t->val;
      VRowsArgs<T> v;
      memset(&v, 0, sizeof(v));
      v.user = tr->f_user.p;
      v.item = tr->f_item.p;
      v.ctx = reinterpret_cast<const T *>(tr->f_ctx.p);
      v.yp = reinterpret_cast<const T *>(tr->yp.p);
      v.idx = t->idx.p;
      v.n = batch;
      v.w0 = reinterpret_cast<const T *>(m->w0.p);
      v.A = p.Vv;
      v.wv = p.wv;
      v.vnv = p.vnv;
      v.n_user = (uint32_t)L.n_ent[0];
      v.ctx_col = (uint32_t)(L.n_ent[0] + L.n_ent[1]);
      v.kp = m->kp;
      v.fin = make_finish(1, scale_b, nullptr, dst_b, t->block_partials.p, t->ticket.p);
      int64_t n_all = batch;
      const bool val_here = va && val_end > val_begin;
      if (val_here && L.val_virtual) {
        v.user2 = va->f_user.p + val_begin;
        v.item2 = va->f_item.p + val_begin;
        v.ctx2 = va->f_ctx.p ? reinterpret_cast<const T *>(va->f_ctx.p) + val_begin * va->n_ctx : nullptr;
        v.yp2 = reinterpret_cast<const T *>(va->yp.p) + val_begin;
        v.n2 = val_end - val_begin;
        v.fin2 = make_finish(1, scale_v, nullptr, dst_v, t->block_partials.p + t->rows_grid, t->ticket.p + 1);
        n_all += v.n2;
      } else if (!val_here) {
        RFM_CUDA(cudaMemsetAsync(dst_v, 0, sizeof(double), ctx->stream));
      }
      const int grid = grid_for(ctx, ceil_div(n_all, units_per_block(m->nch)), TL_VROWS_BLOCKS);
      if (tr->n_ctx) RFM_TRY((launch_vrows_as<T, 1>(ctx, m->nch, MODE_LOSS, false, v, grid)));
      else RFM_TRY((launch_vrows_as<T, 0>(ctx, m->nch, MODE_LOSS, false, v, grid)));
      if (val_here && !L.val_virtual)
        RFM_TRY(loss_pass<T>(t, va, nullptr, val_begin, val_end - val_begin, scale_v, dst_v));
      return RFM_OK;
    }
    RowsArgs<T> a = rows_args<T>(m, t->train);
    a.fac = tl_virtual_fac(t->train->fac_dev(), L);
    a.V = p.Vv;
    a.w = p.wv;
    a.vn = p.vnv;
    a.idx = t->idx.p;
    a.n = batch;
    a.fin = make_finish(1, scale_b, nullptr, dst_b, t->block_partials.p, t->ticket.p);
    int64_t n_all = batch;
    const bool val_here = t->val && val_end > val_begin;
    if (val_here && L.val_virtual) {
      a.fac2 = tl_virtual_fac(t->val->fac_dev(val_begin), L);
      a.yp2 = reinterpret_cast<const T *>(t->val->yp.p) + val_begin;
      a.n2 = val_end - val_begin;
      a.fin2 = make_finish(1, scale_v, nullptr, dst_v, t->block_partials.p + t->rows_grid, t->ticket.p + 1);
      n_all += a.n2;
    } else if (!val_here) {
      RFM_CUDA(cudaMemsetAsync(dst_v, 0, sizeof(double), ctx->stream));
    }
    const int grid = grid_for(ctx, ceil_div(n_all, units_per_block(m->nch)), t->rows_grid / ctx->sm_count);
    RFM_TRY(launch_rows<T>(ctx, m->nch, MODE_LOSS, false, a, grid, true));
    if (val_here && !L.val_virtual)     // other tables: the flat row pass on the real parameters
      RFM_TRY(loss_pass<T>(t, t->val, nullptr, val_begin, val_end - val_begin, scale_v, dst_v));
    return RFM_OK;
  }
  RowsArgs<T> a = rows_args<T>(m, t->train);
  a.idx = t->idx.p;
  a.n = batch;
  a.fin = make_finish(1, scale_b, nullptr, dst_b, t->block_partials.p, t->ticket.p);
  int64_t n_all = batch;
  if (t->val && val_end > val_begin) {
    if (t->val->factored) {
      a.fac2 = t->val->fac_dev(val_begin);
    } else {
      a.row_ptr2 = t->val->row_ptr.p + val_begin;         // row_ptr entries are absolute offsets into col / val
    }
    a.col2 = t->val->col.p;
    a.val2 = reinterpret_cast<const T *>(t->val->val.p);
    a.yp2 = reinterpret_cast<const T *>(t->val->yp.p) + val_begin;
    a.n2 = val_end - val_begin;
    a.fin2 = make_finish(1, scale_v, nullptr, dst_v, t->block_partials.p + t->rows_grid, t->ticket.p + 1);
    n_all += a.n2;
  } else {
    RFM_CUDA(cudaMemsetAsync(dst_v, 0, sizeof(double), ctx->stream));
  }
  const int grid = grid_for(ctx, ceil_div(n_all, units_per_block(m->nch)), t->rows_grid / ctx->sm_count);
  return launch_rows<T>(ctx, m->nch, MODE_LOSS, false, a, grid, t->train->factored);
}

int stage_batch(rfm_fm_trainer *t, const int64_t *batch_rows, int64_t batch) {
  rfm_ctx *ctx = t->m->ctx;
  const int r = t->ring_pos;
  t->ring_pos = (r + 1) % rfm_fm_trainer::RING;
  if (t->stage_used[r]) RFM_CUDA(cudaEventSynchronize(t->stage_ev[r]));
  RFM_TRY(t->stage[r].ensure((size_t)batch));
  memcpy(t->stage[r].p, batch_rows, (size_t)batch * sizeof(int64_t));
  RFM_CUDA(cudaMemcpyAsync(t->idx.p, t->stage[r].p, (size_t)batch * sizeof(int64_t), cudaMemcpyHostToDevice,
                           ctx->stream));
  RFM_CUDA(cudaEventRecord(t->stage_ev[r], ctx->stream));
  t->stage_used[r] = true;
  return RFM_OK;
}

int check_batch(const rfm_fm_trainer *t, int64_t batch, int64_t slot, const char *who) {
  RFM_REQUIRE(t != nullptr, "%s: trainer is NULL", who);
  RFM_REQUIRE(!t->broken, "%s: rfm_fm_trainer_set_two_level failed on this trainer; create a new one", who);
  RFM_REQUIRE(batch >= 1 && batch <= t->max_batch, "%s: batch %lld outside [1, %lld]", who, (long long)batch,
              (long long)t->max_batch);
  RFM_REQUIRE(batch <= t->train->n_rows,
              "Cannot sample %lld out of arrays with dim %lld when replace is False", (long long)batch,
              (long long)t->train->n_rows);
  RFM_REQUIRE(slot >= 0 && slot < t->max_slots, "%s: slot %lld outside [0, %lld)", who, (long long)slot,
              (long long)t->max_slots);
  return RFM_OK;
}

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------------
extern "C" {

// Rows [row_begin, row_end) are copied from the host; the object always has the full shape (every row pointer is
// uploaded). rfm_csr_create passes the whole range; rfm_csr_create_range leaves the rest for the caller to fill.
static int csr_create_impl(rfm_ctx *ctx, int64_t n_rows, int64_t n_cols, const void *indptr, int indptr_is_int64,
                           const int32_t *indices, const double *data, const int64_t *labels,
                           const double *pscores, int dtype, int64_t row_begin, int64_t row_end, rfm_csr **out,
                           const char *who) {
  RFM_REQUIRE(ctx && out, "%s: NULL ctx/out", who);
  *out = nullptr;
  RFM_REQUIRE(n_rows >= 0 && n_cols >= 1, "%s: bad shape (%lld, %lld)", who, (long long)n_rows,
              (long long)n_cols);
  RFM_REQUIRE(n_cols < 0xFFFFFFFFLL, "%s: too many columns", who);
  RFM_REQUIRE(indptr != nullptr, "%s: indptr is NULL", who);
  RFM_REQUIRE(dtype == RFM_F32 || dtype == RFM_F64, "%s: bad dtype %d", who, dtype);
  RFM_REQUIRE((labels == nullptr) == (pscores == nullptr), "%s: labels and pscores go together", who);
  RFM_CUDA(cudaSetDevice(ctx->device));
  const int64_t *p64 = static_cast<const int64_t *>(indptr);
  const int32_t *p32 = static_cast<const int32_t *>(indptr);
  auto ptr_at = [&](int64_t i) -> int64_t { return indptr_is_int64 ? p64[i] : (int64_t)p32[i]; };
  const int64_t nnz = ptr_at(n_rows) - ptr_at(0);
  RFM_REQUIRE(ptr_at(0) == 0 && nnz >= 0, "%s: indptr must start at 0 and be non-decreasing", who);
  RFM_REQUIRE(nnz == 0 || (indices && data), "%s: indices/data are NULL", who);
  RFM_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n_rows,
              "%s: row range [%lld, %lld) outside [0, %lld)", who, (long long)row_begin,
              (long long)row_end, (long long)n_rows);
  const int64_t z0 = ptr_at(row_begin), z1 = ptr_at(row_end), nz = z1 - z0, nr = row_end - row_begin;
  RFM_REQUIRE(z0 >= 0 && nz >= 0 && z1 <= nnz, "%s: indptr must be non-decreasing", who);
  rfm_csr *r = new (std::nothrow) rfm_csr();
  if (!r) return fail(RFM_ERR_NOMEM, "%s: out of host memory", who);
  r->ctx = ctx;
  r->dtype = dtype;
  r->n_rows = n_rows;
  r->n_cols = n_cols;
  r->nnz = nnz;
  r->has_targets = labels != nullptr;
  const size_t es = dsize(dtype);
  int rc = RFM_OK;
  auto body = [&]() -> int {
    RFM_TRY(r->row_ptr.alloc(n_rows + 1));
    RFM_TRY(r->col.alloc(nnz));
    RFM_TRY(r->val.alloc((size_t)nnz * es));
    RFM_TRY(r->yp.alloc((size_t)(n_rows ? n_rows : 1) * es));
    DevBuf<unsigned char> tmp, tmp2;
    if (indptr_is_int64) {
      RFM_TRY(upload(ctx, r->row_ptr.p, indptr, (size_t)(n_rows + 1) * 8));
    } else {
      RFM_TRY(tmp.alloc((size_t)(n_rows + 1) * 4));
      RFM_TRY(upload(ctx, tmp.p, indptr, (size_t)(n_rows + 1) * 4));
      RFM_LAUNCH(ctx, widen_i32_kernel, grid_for(ctx, ceil_div(n_rows + 1, 256), 8), 256, 0,
                 reinterpret_cast<const int32_t *>(tmp.p), r->row_ptr.p, n_rows + 1);
    }
    RFM_TRY(upload(ctx, r->col.p + z0, indices + z0, (size_t)nz * 4));
    DevBuf<int> bad;
    RFM_TRY(bad.alloc(1));
    RFM_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), ctx->stream));
    if (nz > 0)
      RFM_LAUNCH(ctx, check_columns_kernel, grid_for(ctx, ceil_div(nz, 256), 8), 256, 0, r->col.p + z0, nz, n_cols,
                 bad.p);
    if (dtype == RFM_F64) {
      RFM_TRY(upload(ctx, r->val.p + (size_t)z0 * 8, data + z0, (size_t)nz * 8));
    } else if (nz > 0) {
      RFM_TRY(tmp2.alloc((size_t)nz * 8));
      RFM_TRY(upload(ctx, tmp2.p, data + z0, (size_t)nz * 8));
      RFM_LAUNCH(ctx, convert_f64_kernel<float>, grid_for(ctx, ceil_div(nz, 256), 8), 256, 0,
                 reinterpret_cast<const double *>(tmp2.p), reinterpret_cast<float *>(r->val.p) + z0, nz);
    }
    DevBuf<int64_t> ytmp;
    DevBuf<double> pstmp;
    if (labels && nr > 0) {
      RFM_TRY(ytmp.alloc(nr));
      RFM_TRY(pstmp.alloc(nr));
      RFM_TRY(upload(ctx, ytmp.p, labels + row_begin, (size_t)nr * 8));
      RFM_TRY(upload(ctx, pstmp.p, pscores + row_begin, (size_t)nr * 8));
      const int g = grid_for(ctx, ceil_div(nr, 256), 8);
      if (dtype == RFM_F64) {
        RFM_LAUNCH(ctx, targets_kernel<double>, g, 256, 0, ytmp.p, pstmp.p,
                   reinterpret_cast<double *>(r->yp.p) + row_begin, nr);
      } else {
        RFM_LAUNCH(ctx, targets_kernel<float>, g, 256, 0, ytmp.p, pstmp.p,
                   reinterpret_cast<float *>(r->yp.p) + row_begin, nr);
      }
    } else if (!labels || n_rows == 0) {
      RFM_CUDA(cudaMemsetAsync(r->yp.p, 0, (size_t)(n_rows ? n_rows : 1) * es, ctx->stream));
    }
    // The host scan of the row pointers (monotonic? longest row?) runs while the copies enqueued above are
    // in flight: at 12 M rows it costs ~15 ms of one core, a quarter of the PCIe time of the whole upload.
    // No kernel above reads row_ptr contents and every size depends on nnz only, so a bad indptr is harmless
    // on the device; it is reported after the stream has drained.
    int64_t max_len = 0, bad_row = -1;
    if (indptr_is_int64) {
      for (int64_t i = 0; i < n_rows; ++i) {
        const int64_t len = p64[i + 1] - p64[i];
        if (len < 0 && bad_row < 0) bad_row = i;
        if (len > max_len) max_len = len;
      }
    } else {
      for (int64_t i = 0; i < n_rows; ++i) {
        const int64_t len = (int64_t)p32[i + 1] - (int64_t)p32[i];
        if (len < 0 && bad_row < 0) bad_row = i;
        if (len > max_len) max_len = len;
      }
    }
    r->max_row_len = max_len;
    int bad_host = 0;
    RFM_CUDA(cudaMemcpyAsync(&bad_host, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RFM_CUDA(cudaStreamSynchronize(ctx->stream));  // temporaries are freed on return
    RFM_REQUIRE(bad_row < 0, "%s: indptr decreases at row %lld", who, (long long)bad_row);
    RFM_REQUIRE(bad_host == 0, "%s: CSR column index out of range [0, %lld)", who, (long long)n_cols);
    return RFM_OK;
  };
  rc = body();
  if (rc != RFM_OK) {
    delete r;
    return rc;
  }
  *out = r;
  return RFM_OK;
}

int rfm_csr_create(rfm_ctx *ctx, int64_t n_rows, int64_t n_cols, const void *indptr, int indptr_is_int64,
                   const int32_t *indices, const double *data, const int64_t *labels, const double *pscores,
                   int dtype, rfm_csr **out) {
  return csr_create_impl(ctx, n_rows, n_cols, indptr, indptr_is_int64, indices, data, labels, pscores, dtype, 0,
                         n_rows, out, "rfm_csr_create");
}

int rfm_csr_create_range(rfm_ctx *ctx, int64_t n_rows, int64_t n_cols, const void *indptr, int indptr_is_int64,
                         const int32_t *indices, const double *data, const int64_t *labels, const double *pscores,
                         int dtype, int64_t row_begin, int64_t row_end, rfm_csr **out) {
  return csr_create_impl(ctx, n_rows, n_cols, indptr, indptr_is_int64, indices, data, labels, pscores, dtype,
                         row_begin, row_end, out, "rfm_csr_create_range");
}

int rfm_csr_device_ptrs(rfm_csr *rows, void **row_ptr_dev, void **col_dev, void **val_dev, void **targets_dev) {
  RFM_REQUIRE(rows, "rfm_csr_device_ptrs: rows is NULL");
  RFM_REQUIRE(!rows->factored, "rfm_csr_device_ptrs: factored rows have no CSR arrays");
  if (row_ptr_dev) *row_ptr_dev = rows->row_ptr.p;
  if (col_dev) *col_dev = rows->col.p;
  if (val_dev) *val_dev = rows->val.p;
  if (targets_dev) *targets_dev = rows->yp.p;
  return RFM_OK;
}

}  // extern "C" (pause)
// gen != NULL: ids, context values and targets are generated on the device (rfm_factored_generate) instead of uploaded
static int factored_create_impl(rfm_ctx *ctx, int64_t n_rows, const void *users, int32_t users_is_int64,
                                const void *items, int32_t items_is_int64, const rfm_rows_block *blocks,
                                int32_t n_blocks, const void *labels, int32_t label_bytes, const double *pscores,
                                const double *item_pscores, int64_t n_item_pscores, int dtype,
                                const rfm_click_model *gen, rfm_csr **out, int64_t rb = 0, int64_t re = -1) {
  // rows [rb, re) are copied from the host (re < 0: all of them); with a partial range the rest is for the caller to
  // fill on the device (rfm_rows_device_ptrs) and the row statistics wait for rfm_factored_finalize
  RFM_REQUIRE(ctx && out, "rfm_factored_create: NULL ctx/out");
  *out = nullptr;
  if (re < 0) re = n_rows;
  RFM_REQUIRE(rb >= 0 && rb <= re && re <= n_rows, "rfm_factored_create_range: bad row range [%lld, %lld) of %lld",
              (long long)rb, (long long)re, (long long)n_rows);
  const bool partial = rb > 0 || re < n_rows;
  RFM_REQUIRE(!(partial && gen), "rfm_factored_create_range: generated rows have no range");
  const int64_t nsel = re - rb;
  RFM_REQUIRE(n_rows >= 0 && n_rows < 0x7fffffffffLL, "rfm_factored_create: bad row count %lld", (long long)n_rows);
  RFM_REQUIRE(blocks && n_blocks >= 1 && n_blocks <= FAC_MAX_SEG, "rfm_factored_create: between 1 and %d blocks",
              FAC_MAX_SEG);
  RFM_REQUIRE(n_rows == 0 || gen || (users && items), "rfm_factored_create: users/items are NULL");
  RFM_REQUIRE(dtype == RFM_F32 || dtype == RFM_F64, "rfm_factored_create: bad dtype %d", dtype);
  RFM_REQUIRE((labels == nullptr) == (pscores == nullptr && item_pscores == nullptr),
              "rfm_factored_create: labels and pscores go together");
  RFM_REQUIRE(!(pscores && item_pscores), "rfm_factored_create: per-row OR per-item pscores");
  RFM_REQUIRE(!item_pscores || n_item_pscores >= 1, "rfm_factored_create: empty per-item pscore table");
  RFM_REQUIRE(!labels || label_bytes == 1 || label_bytes == 4 || label_bytes == 8,
              "rfm_factored_create: labels must be int8, int32 or int64 (label_bytes = %d)", label_bytes);
  RFM_CUDA(cudaSetDevice(ctx->device));
  // column layout and the id ranges the blocks agree on
  int64_t n_cols = 0, n_ctx = 0, limit[2] = {INT64_MAX, INT64_MAX};
  for (int b = 0; b < n_blocks; ++b) {
    const rfm_rows_block &k = blocks[b];
    RFM_REQUIRE(k.kind == RFM_BLOCK_ID || k.kind == RFM_BLOCK_TABLE || k.kind == RFM_BLOCK_CTX,
                "rfm_factored_create: block %d has unknown kind %d", b, k.kind);
    RFM_REQUIRE(k.n_cols >= 1, "rfm_factored_create: block %d has no columns", b);
    if (k.kind == RFM_BLOCK_CTX) {
      RFM_REQUIRE(k.values || n_rows == 0 || gen, "rfm_factored_create: context block %d has no values", b);
      n_ctx += k.n_cols;
      RFM_REQUIRE(n_ctx <= 32, "rfm_factored_create: more than 32 context columns in total");
    } else {
      RFM_REQUIRE(k.key == RFM_KEY_USER || k.key == RFM_KEY_ITEM, "rfm_factored_create: block %d has unknown key %d", b, k.key);
      const int64_t ne = k.kind == RFM_BLOCK_ID ? k.n_cols : k.n_entities;
      RFM_REQUIRE(ne >= 1 && ne < 0x7fffffffLL, "rfm_factored_create: block %d is indexed by %lld ids", b, (long long)ne);
      limit[k.key] = std::min(limit[k.key], ne);
      if (k.kind == RFM_BLOCK_TABLE) RFM_REQUIRE(k.indptr, "rfm_factored_create: table block %d has no indptr", b);
    }
    n_cols += k.n_cols;
  }
  if (item_pscores) limit[1] = std::min(limit[1], n_item_pscores);
  RFM_REQUIRE(n_cols < 0xFFFFFFFFLL, "rfm_factored_create: too many columns");
  rfm_csr *r = new (std::nothrow) rfm_csr();
  if (!r) return fail(RFM_ERR_NOMEM, "rfm_factored_create: out of host memory");
  r->ctx = ctx;
  r->dtype = dtype;
  r->n_rows = n_rows;
  r->n_cols = n_cols;
  r->has_targets = labels != nullptr || gen != nullptr;
  r->factored = true;
  r->n_seg = n_blocks;
  r->n_ctx = (int)n_ctx;
  const size_t es = dsize(dtype);
  auto body = [&]() -> int {
    const int64_t nr = n_rows ? n_rows : 1;
    RFM_TRY(r->f_user.alloc(nr));
    RFM_TRY(r->f_item.alloc(nr));
    RFM_TRY(r->yp.alloc((size_t)nr * es));
    if (n_ctx) RFM_TRY(r->f_ctx.alloc((size_t)nr * n_ctx * es));
    RFM_TRY(r->f_one.alloc(8));
    {
      const double one_d = 1.0;
      const float one_f = 1.f;
      RFM_CUDA(cudaMemcpyAsync(r->f_one.p, dtype == RFM_F64 ? (const void *)&one_d : (const void *)&one_f, es,
                               cudaMemcpyHostToDevice, ctx->stream));
      RFM_CUDA(cudaStreamSynchronize(ctx->stream));       // the source is on this stack frame
    }
    DevBuf<int> bad;
    RFM_TRY(bad.alloc(4));
    RFM_CUDA(cudaMemsetAsync(bad.p, 0, 4 * sizeof(int), ctx->stream));
    std::vector<DevBuf<unsigned char>> tmp(2 * n_blocks + 4);      // staging, freed (stream-ordered) on return
    int n_tmp = 0;
    const int g = grid_for(ctx, ceil_div(nr, 256), 8);
    // ids: 4 or 8 bytes per row over PCIe as the caller holds them, narrowed (and range-checked) on the device
    for (int side = 0; side < 2 && nsel > 0 && !gen; ++side) {
      const unsigned char *src = static_cast<const unsigned char *>(side == 0 ? users : items);
      const bool is64 = (side == 0 ? users_is_int64 : items_is_int64) != 0;
      int32_t *dst = (side == 0 ? r->f_user.p : r->f_item.p) + rb;
      const int64_t lim = limit[side] == INT64_MAX ? 0x7fffffffLL : limit[side];
      if (is64) {
        DevBuf<unsigned char> &st = tmp[n_tmp++];
        RFM_TRY(st.alloc((size_t)nsel * 8));
        RFM_TRY(upload(ctx, st.p, src + (size_t)rb * 8, (size_t)nsel * 8));
        RFM_LAUNCH(ctx, narrow_ids_kernel<int64_t>, g, 256, 0, reinterpret_cast<const int64_t *>(st.p), dst, nsel, lim,
                   bad.p + side);
      } else {
        RFM_TRY(upload(ctx, dst, src + (size_t)rb * 4, (size_t)nsel * 4));
        RFM_LAUNCH(ctx, narrow_ids_kernel<int32_t>, g, 256, 0, dst, dst, nsel, lim, bad.p + side);
      }
    }
    // blocks
    uint32_t col0 = 0;
    int ctx0 = 0;
    for (int b = 0; b < n_blocks; ++b) {
      const rfm_rows_block &k = blocks[b];
      rfm_csr::Seg &sg = r->seg[b];
      sg.kind = k.kind == RFM_BLOCK_ID ? SEG_ID : k.kind == RFM_BLOCK_TABLE ? SEG_TABLE : SEG_CTX;
      sg.key = k.key;
      sg.col0 = col0;
      sg.width = (int)k.n_cols;
      sg.ctx0 = ctx0;
      if (sg.kind == SEG_TABLE) {
        const int64_t ne = k.n_entities;
        const int64_t *p64 = static_cast<const int64_t *>(k.indptr);
        const int32_t *p32 = static_cast<const int32_t *>(k.indptr);
        std::vector<int32_t> ptr((size_t)ne + 1);
        for (int64_t i = 0; i <= ne; ++i) {
          const int64_t v = k.indptr_is_int64 ? p64[i] : (int64_t)p32[i];
          RFM_REQUIRE(v >= 0 && v < 0x7fffffffLL && (i == 0 ? v == 0 : v >= ptr[(size_t)i - 1]),
                      "rfm_factored_create: indptr of table block %d is not a valid row-pointer array (row %lld)", b,
                      (long long)i);
          ptr[(size_t)i] = (int32_t)v;
        }
        const int64_t tnz = ptr[(size_t)ne];
        sg.n_entities = ne;
        sg.tnz = tnz;
        RFM_REQUIRE(tnz == 0 || (k.indices && k.data), "rfm_factored_create: table block %d has no indices/data", b);
        for (int64_t z = 0; z < tnz; ++z)
          RFM_REQUIRE(k.indices[z] >= 0 && k.indices[z] < k.n_cols,
                      "rfm_factored_create: table block %d has a column index outside [0, %lld)", b, (long long)k.n_cols);
        RFM_TRY(sg.ptr.alloc((size_t)ne + 1));
        RFM_TRY(sg.col.alloc((size_t)tnz));
        RFM_TRY(sg.val.alloc((size_t)(tnz ? tnz : 1) * es));
        RFM_CUDA(cudaMemcpyAsync(sg.ptr.p, ptr.data(), ((size_t)ne + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
        RFM_CUDA(cudaStreamSynchronize(ctx->stream));       // ptr is a host temporary
        if (tnz > 0) {
          RFM_TRY(upload(ctx, sg.col.p, k.indices, (size_t)tnz * 4));
          if (dtype == RFM_F64) {
            RFM_TRY(upload(ctx, sg.val.p, k.data, (size_t)tnz * 8));
          } else {
            DevBuf<unsigned char> &st = tmp[n_tmp++];
            RFM_TRY(st.alloc((size_t)tnz * 8));
            RFM_TRY(upload(ctx, st.p, k.data, (size_t)tnz * 8));
            RFM_LAUNCH(ctx, convert_f64_kernel<float>, grid_for(ctx, ceil_div(tnz, 256), 8), 256, 0,
                       reinterpret_cast<const double *>(st.p), reinterpret_cast<float *>(sg.val.p), tnz);
          }
        }
      } else if (sg.kind == SEG_CTX && nsel > 0 && !gen) {
        const unsigned char *vsrc = reinterpret_cast<const unsigned char *>(k.values) + (size_t)rb * k.n_cols * 8;
        unsigned char *cdst = r->f_ctx.p + (size_t)rb * n_ctx * es;
        if (dtype == RFM_F64 && n_ctx == k.n_cols) {      // the only context block, already in the record's layout
          RFM_TRY(upload(ctx, cdst, vsrc, (size_t)nsel * n_ctx * 8));
        } else {
          DevBuf<unsigned char> &st = tmp[n_tmp++];
          RFM_TRY(st.alloc((size_t)nsel * k.n_cols * 8));
          RFM_TRY(upload(ctx, st.p, vsrc, (size_t)nsel * k.n_cols * 8));
          const int gg = grid_for(ctx, ceil_div(nsel * k.n_cols, 256), 8);
          if (dtype == RFM_F64) {
            RFM_LAUNCH(ctx, ctx_pack_kernel<double>, gg, 256, 0, reinterpret_cast<const double *>(st.p), nsel,
                       (int)k.n_cols, reinterpret_cast<double *>(cdst), (int)n_ctx, ctx0);
          } else {
            RFM_LAUNCH(ctx, ctx_pack_kernel<float>, gg, 256, 0, reinterpret_cast<const double *>(st.p), nsel,
                       (int)k.n_cols, reinterpret_cast<float *>(cdst), (int)n_ctx, ctx0);
          }
        }
      }
      if (sg.kind == SEG_CTX) ctx0 += (int)k.n_cols;
      col0 += (uint32_t)k.n_cols;
    }
    // targets
    if (gen) {
      if (gen->keep_labels) {
        RFM_TRY(r->g_label.alloc((size_t)nr));
        RFM_TRY(r->g_relevance.alloc((size_t)nr));
      }
      if (n_rows > 0) {
        RFM_REQUIRE(gen->n_users <= limit[0] && gen->n_items <= limit[1],
                    "rfm_factored_generate: the click model draws ids the blocks do not cover");
        RFM_TRY(rfm_synth_fill_rows(ctx, gen, n_rows, r->f_user.p, r->f_item.p, r->f_ctx.p, (int)n_ctx, r->yp.p,
                                    r->g_label.p, r->g_relevance.p, dtype));
      }
    } else if (labels && nsel > 0) {
      DevBuf<unsigned char> &ys = tmp[n_tmp++];
      DevBuf<unsigned char> &ps = tmp[n_tmp++];
      RFM_TRY(ys.alloc((size_t)nsel * label_bytes));
      RFM_TRY(upload(ctx, ys.p, static_cast<const unsigned char *>(labels) + (size_t)rb * label_bytes,
                     (size_t)nsel * label_bytes));
      if (item_pscores) {
        RFM_TRY(ps.alloc((size_t)n_item_pscores * 8));
        RFM_TRY(upload(ctx, ps.p, item_pscores, (size_t)n_item_pscores * 8));
      } else {
        RFM_TRY(ps.alloc((size_t)nsel * 8));
        RFM_TRY(upload(ctx, ps.p, pscores + rb, (size_t)nsel * 8));
      }
      const double *psd = reinterpret_cast<const double *>(ps.p);
      const int32_t *item_ids = r->f_item.p + rb;
      unsigned char *ypd = r->yp.p + (size_t)rb * es;
      const bool by_item = item_pscores != nullptr;
#define RFM_TARGETS(T, Y)                                                                                              \
  do {                                                                                                                 \
    if (by_item) {                                                                                                     \
      RFM_LAUNCH(ctx, (targets_by_item_kernel<T, Y>), g, 256, 0, reinterpret_cast<const Y *>(ys.p), item_ids, psd,     \
                 reinterpret_cast<T *>(ypd), nsel);                                                                    \
    } else {                                                                                                           \
      RFM_LAUNCH(ctx, (targets_any_kernel<T, Y>), g, 256, 0, reinterpret_cast<const Y *>(ys.p), psd,                   \
                 reinterpret_cast<T *>(ypd), nsel);                                                                    \
    }                                                                                                                  \
  } while (0)
      if (dtype == RFM_F64) {
        if (label_bytes == 8) RFM_TARGETS(double, int64_t); else if (label_bytes == 4) RFM_TARGETS(double, int32_t); else RFM_TARGETS(double, int8_t);
      } else {
        if (label_bytes == 8) RFM_TARGETS(float, int64_t); else if (label_bytes == 4) RFM_TARGETS(float, int32_t); else RFM_TARGETS(float, int8_t);
      }
#undef RFM_TARGETS
    } else {
      RFM_CUDA(cudaMemsetAsync(r->yp.p, 0, (size_t)nr * es, ctx->stream));
    }
    // row statistics (what the CSR path reads off the row pointers): total non-zeros, longest row
    DevBuf<unsigned long long> nnz_dev;
    RFM_TRY(nnz_dev.alloc(1));
    RFM_CUDA(cudaMemsetAsync(nnz_dev.p, 0, 8, ctx->stream));
    if (n_rows > 0 && !partial) RFM_LAUNCH(ctx, fac_stats_kernel, g, 256, 0, r->fac_dev(), n_rows, nnz_dev.p, bad.p + 2);
    unsigned long long nnz_host = 0;
    int bad_host[4] = {0, 0, 0, 0};
    RFM_CUDA(cudaMemcpyAsync(&nnz_host, nnz_dev.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    RFM_CUDA(cudaMemcpyAsync(bad_host, bad.p, sizeof(bad_host), cudaMemcpyDeviceToHost, ctx->stream));
    RFM_CUDA(cudaStreamSynchronize(ctx->stream));
    RFM_REQUIRE(bad_host[0] == 0, "rfm_factored_create: a user id is outside [0, %lld)", (long long)limit[0]);
    RFM_REQUIRE(bad_host[1] == 0, "rfm_factored_create: an item id is outside [0, %lld)", (long long)limit[1]);
    r->nnz = (int64_t)nnz_host;
    r->max_row_len = bad_host[2];
    return RFM_OK;
  };
  const int rc = body();
  if (rc != RFM_OK) {
    delete r;
    return rc;
  }
  *out = r;
  return RFM_OK;
}

extern "C" {

int rfm_factored_create(rfm_ctx *ctx, int64_t n_rows, const void *users, int32_t users_is_int64, const void *items,
                        int32_t items_is_int64, const rfm_rows_block *blocks, int32_t n_blocks, const void *labels,
                        int32_t label_bytes, const double *pscores, int dtype, rfm_csr **out) {
  return factored_create_impl(ctx, n_rows, users, users_is_int64, items, items_is_int64, blocks, n_blocks, labels,
                              label_bytes, pscores, nullptr, 0, dtype, nullptr, out);
}

int rfm_factored_create_item_pscores(rfm_ctx *ctx, int64_t n_rows, const void *users, int32_t users_is_int64,
                                     const void *items, int32_t items_is_int64, const rfm_rows_block *blocks,
                                     int32_t n_blocks, const void *labels, int32_t label_bytes,
                                     const double *item_pscores, int64_t n_item_pscores, int dtype, rfm_csr **out) {
  RFM_REQUIRE(item_pscores, "rfm_factored_create_item_pscores: item_pscores is NULL");
  return factored_create_impl(ctx, n_rows, users, users_is_int64, items, items_is_int64, blocks, n_blocks, labels,
                              label_bytes, nullptr, item_pscores, n_item_pscores, dtype, nullptr, out);
}

// Rows [row_begin, row_end) are copied from the host (the host pointers are those of the FULL arrays); the object has
// the full shape. The caller fills the other rows on the device (rfm_rows_device_ptrs: NVLink broadcasts from the
// ranks that uploaded them, rfm_b200.dist.sharded_factored_rows) and then calls rfm_factored_finalize.
int rfm_factored_create_range(rfm_ctx *ctx, int64_t n_rows, const void *users, int32_t users_is_int64, const void *items,
                              int32_t items_is_int64, const rfm_rows_block *blocks, int32_t n_blocks, const void *labels,
                              int32_t label_bytes, const double *pscores, const double *item_pscores,
                              int64_t n_item_pscores, int dtype, int64_t row_begin, int64_t row_end, rfm_csr **out) {
  RFM_REQUIRE(ctx && out, "rfm_factored_create_range: NULL ctx/out");
  RFM_REQUIRE(row_end >= 0, "rfm_factored_create_range: bad row range");
  return factored_create_impl(ctx, n_rows, users, users_is_int64, items, items_is_int64, blocks, n_blocks, labels,
                              label_bytes, pscores, item_pscores, n_item_pscores, dtype, nullptr, out, row_begin, row_end);
}

int rfm_rows_device_ptrs(rfm_csr *rows, void **user_dev, void **item_dev, void **ctx_dev, void **targets_dev) {
  RFM_REQUIRE(rows && rows->factored, "rfm_rows_device_ptrs: not factored rows");
  if (user_dev) *user_dev = rows->f_user.p;
  if (item_dev) *item_dev = rows->f_item.p;
  if (ctx_dev) *ctx_dev = rows->n_ctx ? rows->f_ctx.p : nullptr;
  if (targets_dev) *targets_dev = rows->yp.p;
  return RFM_OK;
}

// row statistics of factored rows whose arrays were completed on the device: total non-zeros, longest row
int rfm_factored_finalize(rfm_csr *rows) {
  RFM_REQUIRE(rows && rows->factored, "rfm_factored_finalize: not factored rows");
  rfm_ctx *ctx = rows->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  DevBuf<unsigned long long> nnz_dev;
  DevBuf<int> mx;
  RFM_TRY(nnz_dev.alloc(1));
  RFM_TRY(mx.alloc(1));
  RFM_CUDA(cudaMemsetAsync(nnz_dev.p, 0, 8, ctx->stream));
  RFM_CUDA(cudaMemsetAsync(mx.p, 0, sizeof(int), ctx->stream));
  if (rows->n_rows > 0)
    RFM_LAUNCH(ctx, fac_stats_kernel, grid_for(ctx, ceil_div(rows->n_rows, 256), 8), 256, 0, rows->fac_dev(), rows->n_rows,
               nnz_dev.p, mx.p);
  unsigned long long nnz_host = 0;
  int mx_host = 0;
  RFM_CUDA(cudaMemcpyAsync(&nnz_host, nnz_dev.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(&mx_host, mx.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  rows->nnz = (int64_t)nnz_host;
  rows->max_row_len = mx_host;
  return RFM_OK;
}

int rfm_factored_generate(rfm_ctx *ctx, int64_t n_rows, const rfm_click_model *model, const rfm_rows_block *blocks,
                          int32_t n_blocks, int dtype, rfm_csr **out) {
  RFM_REQUIRE(model, "rfm_factored_generate: model is NULL");
  return factored_create_impl(ctx, n_rows, nullptr, 0, nullptr, 0, blocks, n_blocks, nullptr, 8, nullptr, nullptr, 0,
                              dtype, model, out);
}

// Factored rows -> the stacked CSR the reference would have built, assembled on the device (no PCIe): the row
// kernels run ~10 % faster on resident CSR rows than on factored ones where everything is cached, so a
// long fit uploads factored and trains stacked. Same entries in the same order: results do not change by a bit.
int rfm_rows_materialize(const rfm_csr *rows, rfm_csr **out) {
  RFM_REQUIRE(rows && out, "rfm_rows_materialize: NULL argument");
  *out = nullptr;
  RFM_REQUIRE(rows->factored, "rfm_rows_materialize: the rows are a stacked CSR already");
  RFM_REQUIRE(rows->nnz < 0xFFFFFFF0LL, "rfm_rows_materialize: %lld non-zeros do not fit 32-bit offsets",
              (long long)rows->nnz);
  rfm_ctx *ctx = rows->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  rfm_csr *r = new (std::nothrow) rfm_csr();
  if (!r) return fail(RFM_ERR_NOMEM, "rfm_rows_materialize: out of host memory");
  r->ctx = ctx;
  r->dtype = rows->dtype;
  r->n_rows = rows->n_rows;
  r->n_cols = rows->n_cols;
  r->nnz = rows->nnz;
  r->max_row_len = rows->max_row_len;
  r->has_targets = rows->has_targets;
  const size_t es = dsize(rows->dtype);
  const int64_t n = rows->n_rows;
  auto body = [&]() -> int {
    RFM_TRY(r->row_ptr.alloc(n + 1));
    RFM_TRY(r->col.alloc((size_t)rows->nnz));
    RFM_TRY(r->val.alloc((size_t)rows->nnz * es));
    RFM_TRY(r->yp.alloc((size_t)(n ? n : 1) * es));
    RFM_CUDA(cudaMemcpyAsync(r->yp.p, rows->yp.p, (size_t)n * es, cudaMemcpyDeviceToDevice, ctx->stream));
    DevBuf<uint32_t> len, start, sums, total;
    RFM_TRY(len.alloc((size_t)(n ? n : 1)));
    RFM_TRY(start.alloc((size_t)(n ? n : 1)));
    RFM_TRY(sums.alloc((size_t)ceil_div(n ? n : 1, 4096) + 2));
    RFM_TRY(total.alloc(1));
    RFM_CUDA(cudaMemsetAsync(total.p, 0, 4, ctx->stream));
    const int g = grid_for(ctx, ceil_div(n + 1, 256), 8);
    if (n > 0) {
      RFM_LAUNCH(ctx, fac_all_row_len_kernel, g, 256, 0, rows->fac_dev(), n, len.p);
      RFM_TRY(exclusive_scan_u32(ctx, len.p, start.p, n, sums.p, total.p));
    }
    if (rows->dtype == RFM_F64) {
      RFM_LAUNCH(ctx, fac_fill_csr_kernel<double>, g, 256, 0, rows->fac_dev(), n, start.p, (uint32_t)rows->nnz,
                 r->row_ptr.p, r->col.p, reinterpret_cast<double *>(r->val.p));
    } else {
      RFM_LAUNCH(ctx, fac_fill_csr_kernel<float>, g, 256, 0, rows->fac_dev(), n, start.p, (uint32_t)rows->nnz,
                 r->row_ptr.p, r->col.p, reinterpret_cast<float *>(r->val.p));
    }
    RFM_CUDA(cudaStreamSynchronize(ctx->stream));
    return RFM_OK;
  };
  const int rc = body();
  if (rc != RFM_OK) {
    delete r;
    return rc;
  }
  *out = r;
  return RFM_OK;
}

// ids, context values, targets (and, for generated rows that kept them, click / relevance labels) of rows
// [first, first + n) back on the host: what a test compares with the generator's specification
int rfm_rows_download(rfm_csr *rows, int64_t first, int64_t n, int32_t *users, int32_t *items, double *ctx_values,
                      double *targets, signed char *labels, signed char *relevance) {
  RFM_REQUIRE(rows && rows->factored, "rfm_rows_download: factored rows only");
  RFM_REQUIRE(first >= 0 && n >= 0 && first + n <= rows->n_rows, "rfm_rows_download: row range out of bounds");
  RFM_REQUIRE((!labels && !relevance) || rows->g_label.p, "rfm_rows_download: these rows kept no labels");
  rfm_ctx *ctx = rows->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return RFM_OK;
  const size_t es = dsize(rows->dtype);
  if (users) RFM_CUDA(cudaMemcpyAsync(users, rows->f_user.p + first, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (items) RFM_CUDA(cudaMemcpyAsync(items, rows->f_item.p + first, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (labels) RFM_CUDA(cudaMemcpyAsync(labels, rows->g_label.p + first, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  if (relevance)
    RFM_CUDA(cudaMemcpyAsync(relevance, rows->g_relevance.p + first, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  DevBuf<double> tmp;
  const size_t nc = (size_t)n * rows->n_ctx;
  RFM_TRY(tmp.alloc(nc + (size_t)n));
  if (rows->dtype == RFM_F64) {
    if (ctx_values && nc)
      RFM_CUDA(cudaMemcpyAsync(ctx_values, rows->f_ctx.p + (size_t)first * rows->n_ctx * es, nc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (targets) RFM_CUDA(cudaMemcpyAsync(targets, rows->yp.p + (size_t)first * es, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    if (ctx_values && nc) {
      RFM_LAUNCH(ctx, widen_to_f64_kernel<float>, grid_for(ctx, ceil_div((int64_t)nc, 256), 8), 256, 0,
                 reinterpret_cast<const float *>(rows->f_ctx.p) + (size_t)first * rows->n_ctx, tmp.p, (int64_t)nc);
      RFM_CUDA(cudaMemcpyAsync(ctx_values, tmp.p, nc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (targets) {
      RFM_LAUNCH(ctx, widen_to_f64_kernel<float>, grid_for(ctx, ceil_div(n, 256), 8), 256, 0,
                 reinterpret_cast<const float *>(rows->yp.p) + first, tmp.p + nc, n);
      RFM_CUDA(cudaMemcpyAsync(targets, tmp.p + nc, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

// Overwrite the per-row targets with y/pscore computed by the caller in float64 (fractional labels: the reference
// divides whatever `labels` holds, src/fm.py:80, while rfm_csr_create takes integer labels).
int rfm_csr_set_targets(rfm_csr *rows, const double *targets) {
  RFM_REQUIRE(rows && targets, "rfm_csr_set_targets: NULL argument");
  rfm_ctx *ctx = rows->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  if (rows->n_rows == 0) return RFM_OK;
  if (rows->dtype == RFM_F64) {
    RFM_TRY(upload(ctx, rows->yp.p, targets, (size_t)rows->n_rows * 8));
  } else {
    DevBuf<double> tmp;
    RFM_TRY(tmp.alloc(rows->n_rows));
    RFM_TRY(upload(ctx, tmp.p, targets, (size_t)rows->n_rows * 8));
    RFM_LAUNCH(ctx, convert_f64_kernel<float>, grid_for(ctx, ceil_div(rows->n_rows, 256), 8), 256, 0, tmp.p,
               reinterpret_cast<float *>(rows->yp.p), rows->n_rows);
  }
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  rows->has_targets = true;
  return RFM_OK;
}

int rfm_csr_destroy(rfm_csr *rows) {
  if (rows) {
    cudaSetDevice(rows->ctx->device);
    cudaStreamSynchronize(rows->ctx->stream);
    delete rows;
  }
  return RFM_OK;
}

int rfm_fm_create(rfm_ctx *ctx, int64_t n_features, int32_t n_factors, int dtype, rfm_fm **out) {
  RFM_REQUIRE(ctx && out, "rfm_fm_create: NULL ctx/out");
  *out = nullptr;
  RFM_REQUIRE(n_features >= 1 && n_factors >= 1, "rfm_fm_create: bad shape (%lld, %d)", (long long)n_features,
              n_factors);
  RFM_REQUIRE(n_factors <= 512, "rfm_fm_create: n_factors %d > 512 is not supported", n_factors);
  RFM_REQUIRE(dtype == RFM_F32 || dtype == RFM_F64, "rfm_fm_create: bad dtype %d", dtype);
  RFM_CUDA(cudaSetDevice(ctx->device));
  rfm_fm *m = new (std::nothrow) rfm_fm();
  if (!m) return fail(RFM_ERR_NOMEM, "rfm_fm_create: out of host memory");
  m->ctx = ctx;
  m->dtype = dtype;
  m->n = n_features;
  m->k = n_factors;
  m->nch = (n_factors + 63) / 64;
  m->kp = m->nch * 64;
  const size_t es = dsize(dtype);
  int rc = m->w0.alloc(es);
  if (rc == RFM_OK) rc = m->w.alloc((size_t)n_features * es);
  if (rc == RFM_OK) rc = m->V.alloc((size_t)n_features * m->kp * es);
  if (rc == RFM_OK) rc = m->vn.alloc((size_t)n_features * es);
  if (rc == RFM_OK) {
    cudaMemsetAsync(m->vn.p, 0, (size_t)n_features * es, ctx->stream);
    cudaMemsetAsync(m->w0.p, 0, es, ctx->stream);
    cudaMemsetAsync(m->w.p, 0, (size_t)n_features * es, ctx->stream);
    cudaMemsetAsync(m->V.p, 0, (size_t)n_features * m->kp * es, ctx->stream);
  }
  if (rc != RFM_OK) {
    delete m;
    return rc;
  }
  *out = m;
  return RFM_OK;
}

int rfm_fm_destroy(rfm_fm *m) {
  if (m) {
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    delete m;
  }
  return RFM_OK;
}

int rfm_fm_set_params(rfm_fm *m, const double *w0, const double *w, const double *V) {
  RFM_REQUIRE(m && w0 && w && V, "rfm_fm_set_params: NULL argument");
  rfm_ctx *ctx = m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double> tmp;
  RFM_TRY(tmp.alloc((size_t)m->n * m->k + m->n + 1));
  double *tV = tmp.p, *tw = tmp.p + (size_t)m->n * m->k, *tw0 = tw + m->n;
  RFM_TRY(upload(ctx, tV, V, (size_t)m->n * m->k * 8));
  RFM_TRY(upload(ctx, tw, w, (size_t)m->n * 8));
  RFM_TRY(upload(ctx, tw0, w0, 8));
  const int g = grid_for(ctx, ceil_div(m->n * m->kp, 256), 8);
  if (m->dtype == RFM_F64) {
    RFM_LAUNCH(ctx, pad_rows_kernel<double>, g, 256, 0, tV, reinterpret_cast<double *>(m->V.p), m->n, m->k, m->kp);
    RFM_LAUNCH(ctx, convert_f64_kernel<double>, g, 256, 0, tw, reinterpret_cast<double *>(m->w.p), m->n);
    RFM_LAUNCH(ctx, convert_f64_kernel<double>, 1, 32, 0, tw0, reinterpret_cast<double *>(m->w0.p), (int64_t)1);
    RFM_LAUNCH(ctx, row_norms_kernel<double>, grid_for(ctx, ceil_div(m->n, ROWS_WARPS), 8), ROWS_THREADS, 0,
               reinterpret_cast<const double *>(m->V.p), reinterpret_cast<double *>(m->vn.p), m->n, m->kp);
  } else {
    RFM_LAUNCH(ctx, pad_rows_kernel<float>, g, 256, 0, tV, reinterpret_cast<float *>(m->V.p), m->n, m->k, m->kp);
    RFM_LAUNCH(ctx, convert_f64_kernel<float>, g, 256, 0, tw, reinterpret_cast<float *>(m->w.p), m->n);
    RFM_LAUNCH(ctx, convert_f64_kernel<float>, 1, 32, 0, tw0, reinterpret_cast<float *>(m->w0.p), (int64_t)1);
    RFM_LAUNCH(ctx, row_norms_kernel<float>, grid_for(ctx, ceil_div(m->n, ROWS_WARPS), 8), ROWS_THREADS, 0,
               reinterpret_cast<const float *>(m->V.p), reinterpret_cast<float *>(m->vn.p), m->n, m->kp);
  }
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  ++m->version;
  return RFM_OK;
}


int rfm_fm_get_params(rfm_fm *m, double *w0, double *w, double *V) {
  RFM_REQUIRE(m && w0 && w && V, "rfm_fm_get_params: NULL argument");
  rfm_ctx *ctx = m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  DevBuf<double> tmp;
  RFM_TRY(tmp.alloc((size_t)m->n * m->k + m->n + 1));
  double *tV = tmp.p, *tw = tmp.p + (size_t)m->n * m->k, *tw0 = tw + m->n;
  const int g = grid_for(ctx, ceil_div(m->n * m->k, 256), 8);
  if (m->dtype == RFM_F64) {
    RFM_LAUNCH(ctx, unpad_rows_kernel<double>, g, 256, 0, reinterpret_cast<const double *>(m->V.p), tV, m->n, m->k,
               m->kp);
    RFM_LAUNCH(ctx, widen_to_f64_kernel<double>, g, 256, 0, reinterpret_cast<const double *>(m->w.p), tw, m->n);
    RFM_LAUNCH(ctx, widen_to_f64_kernel<double>, 1, 32, 0, reinterpret_cast<const double *>(m->w0.p), tw0,
               (int64_t)1);
  } else {
    RFM_LAUNCH(ctx, unpad_rows_kernel<float>, g, 256, 0, reinterpret_cast<const float *>(m->V.p), tV, m->n, m->k,
               m->kp);
    RFM_LAUNCH(ctx, widen_to_f64_kernel<float>, g, 256, 0, reinterpret_cast<const float *>(m->w.p), tw, m->n);
    RFM_LAUNCH(ctx, widen_to_f64_kernel<float>, 1, 32, 0, reinterpret_cast<const float *>(m->w0.p), tw0, (int64_t)1);
  }
  RFM_CUDA(cudaMemcpyAsync(V, tV, (size_t)m->n * m->k * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(w, tw, (size_t)m->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaMemcpyAsync(w0, tw0, 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

}  // extern "C" (pause)
namespace {
int check_rows(const rfm_fm *m, const rfm_csr *rows, const char *who) {
  RFM_REQUIRE(m && rows, "%s: NULL argument", who);
  RFM_REQUIRE(rows->ctx == m->ctx, "%s: rows and model live in different contexts", who);
  RFM_REQUIRE(rows->dtype == m->dtype, "%s: rows dtype %d != model dtype %d", who, rows->dtype, m->dtype);
  RFM_REQUIRE(rows->n_cols == m->n, "%s: rows have %lld columns, model has %lld features", who,
              (long long)rows->n_cols, (long long)m->n);
  return RFM_OK;
}

// scores of every row into device memory (asynchronous): the evaluation chain of fit(evaluator=...) never leaves the GPU
template <typename T>
int predict_dev_impl(rfm_fm *m, const rfm_csr *rows, double *out_dev) {
  rfm_ctx *ctx = m->ctx;
  if (rows->n_rows == 0) return RFM_OK;
  RowsArgs<T> a = rows_args<T>(m, rows);
  a.n = rows->n_rows;
  a.out = out_dev;
  return launch_rows<T>(ctx, m->nch, MODE_PREDICT, false, a, grid_for(ctx, ceil_div(rows->n_rows, units_per_block(m->nch)), 6), rows->factored);
}

template <typename T>
int predict_impl(rfm_fm *m, const rfm_csr *rows, double *out_host) {
  rfm_ctx *ctx = m->ctx;
  if (rows->n_rows == 0) return RFM_OK;
  DevBuf<double> out;
  RFM_TRY(out.alloc(rows->n_rows));
  RowsArgs<T> a = rows_args<T>(m, rows);
  a.n = rows->n_rows;
  a.out = out.p;
  RFM_TRY(launch_rows<T>(ctx, m->nch, MODE_PREDICT, false, a, grid_for(ctx, ceil_div(rows->n_rows, units_per_block(m->nch)), 6), rows->factored));
  RFM_CUDA(cudaMemcpyAsync(out_host, out.p, (size_t)rows->n_rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

template <typename T>
int logloss_impl(rfm_fm *m, const rfm_csr *rows, double *out_host) {
  rfm_ctx *ctx = m->ctx;
  const int grid = grid_for(ctx, ceil_div(rows->n_rows, units_per_block(m->nch)), 6);
  DevBuf<double> partials, res;
  DevBuf<uint32_t> ticket;
  RFM_TRY(partials.alloc((size_t)grid));
  RFM_TRY(res.alloc(1));
  RFM_TRY(ticket.alloc(1));
  RFM_CUDA(cudaMemsetAsync(ticket.p, 0, sizeof(uint32_t), ctx->stream));
  RowsArgs<T> a = rows_args<T>(m, rows);
  a.n = rows->n_rows;
  a.fin = make_finish(1, 1.0 / (double)rows->n_rows, nullptr, res.p, partials.p, ticket.p);
  RFM_TRY(launch_rows<T>(ctx, m->nch, MODE_LOSS, false, a, grid, rows->factored));
  RFM_CUDA(cudaMemcpyAsync(out_host, res.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}
}  // namespace
extern "C" {

int rfm_fm_predict(rfm_fm *m, const rfm_csr *rows, double *out_scores) {
  RFM_TRY(check_rows(m, rows, "rfm_fm_predict"));
  RFM_REQUIRE(out_scores || rows->n_rows == 0, "rfm_fm_predict: out_scores is NULL");
  RFM_CUDA(cudaSetDevice(m->ctx->device));
  return m->dtype == RFM_F64 ? predict_impl<double>(m, rows, out_scores) : predict_impl<float>(m, rows, out_scores);
}

int rfm_fm_predict_dev(rfm_fm *m, const rfm_csr *rows, double *out_scores_dev) {
  RFM_TRY(check_rows(m, rows, "rfm_fm_predict_dev"));
  RFM_REQUIRE(out_scores_dev || rows->n_rows == 0, "rfm_fm_predict_dev: out_scores_dev is NULL");
  RFM_CUDA(cudaSetDevice(m->ctx->device));
  return m->dtype == RFM_F64 ? predict_dev_impl<double>(m, rows, out_scores_dev)
                             : predict_dev_impl<float>(m, rows, out_scores_dev);
}

int rfm_fm_logloss(rfm_fm *m, const rfm_csr *rows, double *out_loss) {
  RFM_TRY(check_rows(m, rows, "rfm_fm_logloss"));
  RFM_REQUIRE(out_loss, "rfm_fm_logloss: out_loss is NULL");
  RFM_REQUIRE(rows->has_targets, "rfm_fm_logloss: rows were created without labels/pscores");
  RFM_REQUIRE(rows->n_rows > 0, "rfm_fm_logloss: no rows");
  RFM_CUDA(cudaSetDevice(m->ctx->device));
  return m->dtype == RFM_F64 ? logloss_impl<double>(m, rows, out_loss) : logloss_impl<float>(m, rows, out_loss);
}

int rfm_fm_trainer_create(rfm_fm *m, const rfm_csr *train, const rfm_csr *val, int64_t max_batch,
                          int64_t max_slots, rfm_fm_trainer **out) {
  RFM_REQUIRE(out, "rfm_fm_trainer_create: out is NULL");
  *out = nullptr;
  RFM_TRY(check_rows(m, train, "rfm_fm_trainer_create(train)"));
  if (val) RFM_TRY(check_rows(m, val, "rfm_fm_trainer_create(val)"));
  RFM_REQUIRE(train->has_targets && (!val || val->has_targets), "rfm_fm_trainer_create: rows need labels/pscores");
  RFM_REQUIRE(!val || val->factored == train->factored,
              "rfm_fm_trainer_create: train and val rows must both be CSR or both be factored");
  RFM_REQUIRE(max_batch >= 1 && max_slots >= 1, "rfm_fm_trainer_create: bad max_batch/max_slots");
  RFM_REQUIRE(max_batch <= 0x7fffffffLL, "rfm_fm_trainer_create: max_batch too large");
  rfm_ctx *ctx = m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  if (max_batch > train->n_rows) max_batch = train->n_rows > 0 ? train->n_rows : 1;
  // Triple layout: when rows are of near-uniform length, row q of the batch owns a fixed stride of
  // max_row_len slots (unused ones hold a sentinel key that sorts last); no row-length scan is needed
  // and the element count is known on the host. Otherwise slots are packed with an exclusive scan.
  const double mean_len = train->n_rows > 0 ? (double)train->nnz / (double)train->n_rows : 0.0;
  const bool fixed_stride = train->max_row_len >= 1 && (double)train->max_row_len <= 1.25 * mean_len + 1.0;
  int64_t nnz_cap = max_batch * train->max_row_len;
  if (!fixed_stride && nnz_cap > train->nnz) nnz_cap = train->nnz;
  if (nnz_cap < 1) nnz_cap = 1;
  RFM_REQUIRE(nnz_cap < 0xFFFFFFF0LL, "rfm_fm_trainer_create: batch holds too many non-zeros (%lld)",
              (long long)nnz_cap);
  rfm_fm_trainer *t = new (std::nothrow) rfm_fm_trainer();
  if (!t) return fail(RFM_ERR_NOMEM, "rfm_fm_trainer_create: out of host memory");
  t->m = m;
  t->train = train;
  t->val = val;
  t->max_batch = max_batch;
  t->max_slots = max_slots;
  t->nnz_cap = nnz_cap;
  t->stride = fixed_stride ? (uint32_t)train->max_row_len : 0u;
  const size_t es = dsize(m->dtype);
  auto body = [&]() -> int {
    // resident CTAs per SM for the row/column kernels: 6 x 256 threads keeps 48 warps in flight
    t->rows_grid = ctx->sm_count * 6;
    RFM_TRY(t->idx.alloc(max_batch));
    RFM_TRY(t->row_len.alloc(max_batch));
    RFM_TRY(t->bptr.alloc(max_batch));
    RFM_TRY(t->scan_tmp.alloc(ceil_div(max_batch, 4096) + 2));
    RFM_TRY(t->count.alloc(1));
    RFM_TRY(t->S.alloc((size_t)max_batch * m->kp * es));
    RFM_TRY(t->E.alloc((size_t)max_batch * es));
    const int64_t chunk_cap = ceil_div(nnz_cap, cols_unit(m->nch)) + 1;   // carry records: one per column-pass unit
    RFM_TRY(t->carry_vec.alloc((size_t)chunk_cap * 2 * m->kp * es));
    RFM_TRY(t->carry_ac.alloc((size_t)chunk_cap * 4 * es));
    RFM_TRY(t->block_partials.alloc((size_t)t->rows_grid * 2));
    RFM_TRY(t->ticket.alloc(2));
    RFM_CUDA(cudaMemsetAsync(t->ticket.p, 0, 2 * sizeof(uint32_t), ctx->stream));
    RFM_TRY(t->tails.alloc((size_t)chunk_cap));
    RFM_TRY(t->n_tails.alloc(1));
    RFM_TRY(t->losses.alloc((size_t)max_slots * 2));
    RFM_TRY(t->loss_sums.alloc(4));
    RFM_CUDA(cudaMemsetAsync(t->losses.p, 0, (size_t)max_slots * 2 * sizeof(double), ctx->stream));
    // keys go up to n (the sentinel of the fixed-stride layout), hence n + 1 key values
    if (m->dtype == RFM_F64) RFM_TRY(t->sort64.init(nnz_cap, m->n + 1)); else RFM_TRY(t->sort32.init(nnz_cap, m->n + 1));
    for (int r = 0; r < rfm_fm_trainer::RING; ++r) RFM_CUDA(cudaEventCreateWithFlags(&t->stage_ev[r], cudaEventDisableTiming));
    return RFM_OK;
  };
  const int rc = body();
  if (rc != RFM_OK) {
    rfm_fm_trainer_destroy(t);
    return rc;
  }
  *out = t;
  return RFM_OK;
}

// Two-level step for factored train rows (two_level.cuh). mode 1: switch it on; 2: only where the cost model
// (gathered parameter rows per pass: batch x mean row length, against batch x virtual row length + the entity
// tables' entries) predicts at least 1.5 x. Call before the first epoch. *enabled reports the outcome.
int rfm_fm_trainer_set_two_level(rfm_fm_trainer *t, int32_t mode, int32_t *enabled) {
  RFM_REQUIRE(t, "rfm_fm_trainer_set_two_level: trainer is NULL");
  if (enabled) *enabled = t->tl ? 1 : 0;
  RFM_REQUIRE(mode == 1 || mode == 2, "rfm_fm_trainer_set_two_level: mode must be 1 (on) or 2 (auto)");
  if (t->tl) return RFM_OK;
  const rfm_csr *tr = t->train;
  if (!tr->factored) {
    RFM_REQUIRE(mode == 2, "rfm_fm_trainer_set_two_level: the train rows are not factored");
    return RFM_OK;
  }
  rfm_fm *m = t->m;
  rfm_ctx *ctx = m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  TwoLevel *Lp = new (std::nothrow) TwoLevel();
  if (!Lp) return fail(RFM_ERR_NOMEM, "rfm_fm_trainer_set_two_level: out of host memory");
  TwoLevel &L = *Lp;
  int64_t lim[2] = {INT64_MAX, INT64_MAX}, m2_est = tr->n_ctx;
  int n_key_segs[2] = {0, 0};
  for (int s = 0; s < tr->n_seg; ++s) {
    const rfm_csr::Seg &g = tr->seg[s];
    if (g.kind == SEG_CTX) {
      for (int j = 0; j < g.width; ++j) L.ctx_cols.c[g.ctx0 + j] = g.col0 + (uint32_t)j;
      continue;
    }
    lim[g.key] = std::min<int64_t>(lim[g.key], g.kind == SEG_ID ? g.width : g.n_entities);
    ++n_key_segs[g.key];
  }
  for (int key = 0; key < 2; ++key) L.n_ent[key] = n_key_segs[key] ? lim[key] : 0;
  for (int s = 0; s < tr->n_seg; ++s) {
    const rfm_csr::Seg &g = tr->seg[s];
    if (g.kind == SEG_ID) m2_est += L.n_ent[g.key];
    if (g.kind == SEG_TABLE) m2_est += g.tnz;
  }
  L.nv = L.n_ent[0] + L.n_ent[1] + tr->n_ctx;
  // the lean row kernel covers [user | item | at most one context column]; that column is dense, so it skips the sort
  // (extra CTAs of the level-1 column pass sum it) and a virtual row has two sorted entries. RFM_TL_GENERIC=1 keeps
  // the generic factored row kernel with every context column in the sort.
  L.lean = L.n_ent[0] > 0 && L.n_ent[1] > 0 && tr->n_ctx <= 1 && getenv("RFM_TL_GENERIC") == nullptr;
  L.stride = L.lean ? 2u : (uint32_t)((L.n_ent[0] > 0) + (L.n_ent[1] > 0) + tr->n_ctx);
  const double mean_len = tr->n_rows > 0 ? (double)tr->nnz / (double)tr->n_rows : 0.0;
  const double flat = (double)t->max_batch * mean_len;
  const double two = (double)t->max_batch * (double)L.stride + (double)m2_est;
  const int64_t cap1 = t->max_batch * (int64_t)L.stride;
  const bool fits = L.nv >= 1 && L.nv < 0xFFFFFFF0LL && cap1 < 0xFFFFFFF0LL && m2_est < 0x7FFFFFF0LL;
  if (mode == 2 && (!fits || flat < 1.5 * two)) {
    delete Lp;
    return RFM_OK;
  }
  const size_t es = dsize(m->dtype);
  auto body = [&]() -> int {
    RFM_REQUIRE(fits, "rfm_fm_trainer_set_two_level: too many entities or batch entries for 32-bit offsets");
    const FacDev f = tr->fac_dev();
    const int g = grid_for(ctx, ceil_div(L.nv, 256), 8);
    // entity lists: lengths -> scan -> fill (also the triples of the level-2 list, sorted by real column below)
    DevBuf<uint32_t> len, scan_tmp;
    RFM_TRY(len.alloc((size_t)L.nv));
    RFM_TRY(scan_tmp.alloc((size_t)ceil_div(L.nv, 4096) + 2));
    RFM_TRY(L.ent_ptr.alloc((size_t)L.nv + 1));
    RFM_TRY(L.m2_dev.alloc(1));
    RFM_LAUNCH(ctx, tl_ent_len_kernel, g, 256, 0, f, L.n_ent[0], L.n_ent[1], L.nv, len.p);
    RFM_TRY(exclusive_scan_u32(ctx, len.p, L.ent_ptr.p, L.nv, scan_tmp.p, L.m2_dev.p));
    uint32_t m2 = 0;
    RFM_CUDA(cudaMemcpyAsync(L.ent_ptr.p + L.nv, L.m2_dev.p, sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    RFM_CUDA(cudaMemcpyAsync(&m2, L.m2_dev.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    RFM_CUDA(cudaStreamSynchronize(ctx->stream));
    L.m2 = m2;
    RFM_REQUIRE(L.m2 >= 1, "rfm_fm_trainer_set_two_level: the blocks hold no entries");
    RFM_TRY(L.ent_col.alloc((size_t)L.m2));
    RFM_TRY(L.ent_val.alloc((size_t)L.m2 * es));
    int sorted = 0;
    if (m->dtype == RFM_F64) {
      RFM_TRY(L.l2_64.init(L.m2, m->n));
      RFM_LAUNCH(ctx, tl_ent_fill_kernel<double>, g, 256, 0, f, L.n_ent[0], L.n_ent[1], L.nv, L.ctx_cols, L.ent_ptr.p,
                 L.ent_col.p, reinterpret_cast<double *>(L.ent_val.p), L.l2_64.keys[0].p, L.l2_64.pos[0].p,
                 L.l2_64.val[0].p);
      RFM_TRY(L.l2_64.sort(ctx, L.m2_dev.p, &sorted));
      RFM_TRY(L.l2_64.check(ctx));
    } else {
      RFM_TRY(L.l2_32.init(L.m2, m->n));
      RFM_LAUNCH(ctx, tl_ent_fill_kernel<float>, g, 256, 0, f, L.n_ent[0], L.n_ent[1], L.nv, L.ctx_cols, L.ent_ptr.p,
                 L.ent_col.p, reinterpret_cast<float *>(L.ent_val.p), L.l2_32.keys[0].p, L.l2_32.pos[0].p,
                 L.l2_32.val[0].p);
      RFM_TRY(L.l2_32.sort(ctx, L.m2_dev.p, &sorted));
      RFM_TRY(L.l2_32.check(ctx));
    }
    L.l2_buf = sorted;
    RFM_TRY(L.Vv.alloc((size_t)L.nv * m->kp * es));
    RFM_TRY(L.wv.alloc((size_t)L.nv * es));
    RFM_TRY(L.vnv.alloc((size_t)L.nv * es));
    RFM_TRY(L.R.alloc((size_t)L.nv * (m->kp + 2) * es));
    RFM_TRY(L.n_tails2.alloc(2));
    // val rows ride in the virtual loss launch when they are keyed by the same tables (same blocks, same bytes)
    const rfm_csr *va = t->val;
    bool same = va && va->factored && va->n_seg == tr->n_seg && va->n_ctx == tr->n_ctx;
    DevBuf<int> differ;
    RFM_TRY(differ.alloc(1));
    RFM_CUDA(cudaMemsetAsync(differ.p, 0, sizeof(int), ctx->stream));
    for (int s = 0; same && s < tr->n_seg; ++s) {
      const rfm_csr::Seg &a = tr->seg[s], &b = va->seg[s];
      same = a.kind == b.kind && a.key == b.key && a.col0 == b.col0 && a.width == b.width && a.ctx0 == b.ctx0 &&
             a.n_entities == b.n_entities && a.tnz == b.tnz;
      if (!same || a.kind != SEG_TABLE) continue;
      const struct { const void *x, *y; int64_t words; } arrays[3] = {
          {a.ptr.p, b.ptr.p, a.n_entities + 1}, {a.col.p, b.col.p, a.tnz}, {a.val.p, b.val.p, a.tnz * (int64_t)(es / 4)}};
      for (const auto &ar : arrays)
        if (ar.words > 0 && ar.x != ar.y)
          RFM_LAUNCH(ctx, tl_compare_kernel, grid_for(ctx, ceil_div(ar.words, 256), 8), 256, 0,
                     static_cast<const uint32_t *>(ar.x), static_cast<const uint32_t *>(ar.y), ar.words, differ.p);
    }
    int differ_host = 0;
    RFM_CUDA(cudaMemcpyAsync(&differ_host, differ.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RFM_CUDA(cudaStreamSynchronize(ctx->stream));
    L.val_virtual = same && differ_host == 0;
    if (L.lean && tr->n_ctx == 1) {
      RFM_TRY(L.Cq.alloc((size_t)t->max_batch * es));
      RFM_TRY(L.cpart.alloc((size_t)ctx->sm_count * (m->kp + 2) * es));
      RFM_TRY(L.ctx_ticket.alloc(1));
      RFM_CUDA(cudaMemsetAsync(L.ctx_ticket.p, 0, sizeof(uint32_t), ctx->stream));
    }
    // level 1 sorts batch x stride virtual entries by virtual column; the carry records serve both levels.
    // From here on the trainer's own buffers change: a failure leaves it unusable (broken), not silently flat.
    t->broken = true;
    t->stride = L.stride;
    t->nnz_cap = cap1;
    t->count_host = 0;
    if (m->dtype == RFM_F64) RFM_TRY(t->sort64.init(cap1, L.nv + 1)); else RFM_TRY(t->sort32.init(cap1, L.nv + 1));
    const int64_t chunk_cap = ceil_div(std::max<int64_t>(cap1, L.m2), (int64_t)ROWS_THREADS) + 1;   // the smaller unit
    RFM_TRY(t->carry_vec.alloc((size_t)chunk_cap * 2 * m->kp * es));
    RFM_TRY(t->carry_ac.alloc((size_t)chunk_cap * 4 * es));
    RFM_TRY(t->tails.alloc((size_t)chunk_cap));
    return RFM_OK;
  };
  const int rc = body();
  if (rc != RFM_OK) {
    delete Lp;
    return rc;
  }
  t->tl = Lp;
  t->broken = false;
  if (enabled) *enabled = 1;
  return RFM_OK;
}

int rfm_fm_trainer_destroy(rfm_fm_trainer *t) {
  if (t) {
    cudaSetDevice(t->m->ctx->device);
    cudaStreamSynchronize(t->m->ctx->stream);
    for (int r = 0; r < rfm_fm_trainer::RING; ++r)
      if (t->stage_ev[r]) cudaEventDestroy(t->stage_ev[r]);
    if (t->xchg_borrowed) {
      t->m->ctx->dp.borrowers = 0;
    } else {
      for (int q = 0; q < t->dp_world; ++q)
        if (q != t->dp_rank && t->peer_base[q]) cudaIpcCloseMemHandle(t->peer_base[q]);
      if (t->xchg) cudaFree(t->xchg);
    }
    delete t->tl;
    delete t;
  }
  return RFM_OK;
}

int rfm_fm_train_epoch(rfm_fm_trainer *t, const int64_t *batch_rows, int64_t batch, double lr, int64_t slot) {
  RFM_TRY(check_batch(t, batch, slot, "rfm_fm_train_epoch"));
  RFM_REQUIRE(batch_rows, "rfm_fm_train_epoch: batch_rows is NULL");
  RFM_CUDA(cudaSetDevice(t->m->ctx->device));
  for (int64_t q = 0; q < batch; ++q)
    RFM_REQUIRE(batch_rows[q] >= 0 && batch_rows[q] < t->train->n_rows, "rfm_fm_train_epoch: row id %lld out of range",
                (long long)batch_rows[q]);
  RFM_TRY(stage_batch(t, batch_rows, batch));
  const FeistelKey none = make_feistel_key(1, 0, 0);
  return t->m->dtype == RFM_F64 ? epoch_impl<double>(t, batch, lr, slot, false, none)
                                : epoch_impl<float>(t, batch, lr, slot, false, none);
}

int rfm_fm_train_epoch_sampled(rfm_fm_trainer *t, uint32_t seed, uint32_t epoch, int64_t batch, double lr,
                               int64_t slot) {
  RFM_TRY(check_batch(t, batch, slot, "rfm_fm_train_epoch_sampled"));
  rfm_ctx *ctx = t->m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  RFM_REQUIRE(t->train->n_rows <= (1LL << 32), "rfm_fm_train_epoch_sampled: at most 2^32 rows");
  const FeistelKey key = make_feistel_key((uint64_t)t->train->n_rows, seed, epoch);
  return t->m->dtype == RFM_F64 ? epoch_impl<double>(t, batch, lr, slot, true, key)
                                : epoch_impl<float>(t, batch, lr, slot, true, key);
}

int rfm_fm_grad_size(rfm_fm_trainer *t, int64_t *n_scalars) {
  RFM_REQUIRE(t && n_scalars, "rfm_fm_grad_size: NULL argument");
  *n_scalars = grad_total(t->m->n, t->m->kp);
  return RFM_OK;
}

int rfm_fm_grad_ptr_dev(rfm_fm_trainer *t, void **grad_dev) {
  RFM_REQUIRE(t && grad_dev, "rfm_fm_grad_ptr_dev: NULL argument");
  RFM_CUDA(cudaSetDevice(t->m->ctx->device));
  if (!t->xchg && !t->grad.p) {
    int64_t n = 0;
    rfm_fm_grad_size(t, &n);
    RFM_TRY(t->grad.alloc((size_t)n * dsize(t->m->dtype)));
  }
  *grad_dev = t->grad_ptr();
  return RFM_OK;
}

}  // extern "C" (pause)
namespace {
template <typename T>
__global__ void store_sum_e_kernel(const double *__restrict__ src, T *__restrict__ grad0) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *grad0 = static_cast<T>(*src);
}
}  // namespace
extern "C" {

static int grad_epoch_tail(rfm_fm_trainer *t, int64_t batch, bool sampled, const FeistelKey &key, int64_t q0) {
  void *g = nullptr;
  RFM_TRY(rfm_fm_grad_ptr_dev(t, &g));
  if (t->m->dtype == RFM_F64) return step_core<double, true>(t, batch, 0.0, sampled, key, q0);
  return step_core<float, true>(t, batch, 0.0, sampled, key, q0);
}

int rfm_fm_grad_epoch(rfm_fm_trainer *t, const int64_t *batch_rows, int64_t batch) {
  RFM_TRY(check_batch(t, batch, 0, "rfm_fm_grad_epoch"));
  RFM_REQUIRE(batch_rows, "rfm_fm_grad_epoch: batch_rows is NULL");
  RFM_CUDA(cudaSetDevice(t->m->ctx->device));
  for (int64_t q = 0; q < batch; ++q)
    RFM_REQUIRE(batch_rows[q] >= 0 && batch_rows[q] < t->train->n_rows, "rfm_fm_grad_epoch: row id %lld out of range",
                (long long)batch_rows[q]);
  RFM_TRY(stage_batch(t, batch_rows, batch));
  return grad_epoch_tail(t, batch, false, make_feistel_key(1, 0, 0), 0);
}

int rfm_fm_grad_epoch_sampled(rfm_fm_trainer *t, uint32_t seed, uint32_t epoch, int64_t q_begin, int64_t batch) {
  RFM_TRY(check_batch(t, batch, 0, "rfm_fm_grad_epoch_sampled"));
  rfm_ctx *ctx = t->m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  RFM_REQUIRE(q_begin >= 0 && q_begin + batch <= t->train->n_rows,
              "Cannot sample %lld out of arrays with dim %lld when replace is False", (long long)(q_begin + batch),
              (long long)t->train->n_rows);
  RFM_REQUIRE(t->train->n_rows <= (1LL << 32), "rfm_fm_grad_epoch_sampled: at most 2^32 rows");
  const FeistelKey key = make_feistel_key((uint64_t)t->train->n_rows, seed, epoch);
  return grad_epoch_tail(t, batch, true, key, q_begin);
}

int rfm_fm_apply_grad(rfm_fm_trainer *t, double lr) {
  RFM_REQUIRE(t && t->grad_ptr(), "rfm_fm_apply_grad: no gradient buffer (call rfm_fm_grad_epoch first)");
  rfm_fm *m = t->m;
  rfm_ctx *ctx = m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  const int g = grid_for(ctx, ceil_div(m->n, ROWS_WARPS), 8);
  if (m->dtype == RFM_F64) {
    double *gr = reinterpret_cast<double *>(t->grad_ptr());
    RFM_LAUNCH(ctx, apply_grad_rows_kernel<double>, g, ROWS_THREADS, 0, reinterpret_cast<double *>(m->w0.p),
               reinterpret_cast<double *>(m->w.p), reinterpret_cast<double *>(m->V.p),
               reinterpret_cast<double *>(m->vn.p), gr, gr + GRAD_W_OFF, gr + grad_v_off(m->n), m->n, m->kp, lr);
  } else {
    float *gr = reinterpret_cast<float *>(t->grad_ptr());
    RFM_LAUNCH(ctx, apply_grad_rows_kernel<float>, g, ROWS_THREADS, 0, reinterpret_cast<float *>(m->w0.p),
               reinterpret_cast<float *>(m->w.p), reinterpret_cast<float *>(m->V.p),
               reinterpret_cast<float *>(m->vn.p), gr, gr + GRAD_W_OFF, gr + grad_v_off(m->n), m->n, m->kp,
               (float)lr);
  }
  ++m->version;
  return RFM_OK;
}

}  // extern "C" (pause)
namespace {

// ---- data-parallel step over NVLink peer memory ---------------------------------------------------------------
// One kernel per step and rank replaces {all-reduce, dense apply}: after a cross-GPU barrier every rank sums ITS
// slice of the gradient over all ranks' buffers (peer loads, fixed rank order, so every rank later sees the same
// bits) and writes it back in place; after a second barrier every rank reads each slice from its owner and
// applies it to its replica of the parameters. 2 (N - 1) / N of the gradient crosses NVLink per rank, nothing is
// staged, and the buffers alternate between two parities so that a fast rank never overwrites what a slow one
// still reads. Flags are monotonically increasing step numbers written with system-scope release stores.
constexpr int DPX_THREADS = 256;
constexpr uint32_t DPX_SPIN_LIMIT = 1u << 23;   // a few seconds of polling, then the status flag is raised

struct DpxArgs {
  const unsigned char *peer[rfm_fm_trainer::DP_MAX_WORLD];   // every rank's exchange region (own included)
  size_t grad_off;          // byte offset of this step's parity inside a region
  size_t flags_off;         // byte offset of the flag block: uint32 [2 phases][DP_MAX_WORLD]
  int rank, world;
  uint32_t seq;             // this step's number (>= 1)
  int64_t total, slice;     // scalars in a gradient buffer; scalars per rank slice (a multiple of 2)
  uint32_t *local;          // [arrive counter, go flag, error flag] of this rank
  const double *loss_in;    // this rank's previous-step loss sums, or nullptr
  double *loss_out;         // global loss sums of the previous step (read back from the header)
  unsigned long long *trace;   // RFM_DPX_TRACE=1: %globaltimer of CTA 0 at [start, barrier 0 passed, slice reduced,
                               // barrier 1 passed, applied] (rfm_fm_dp_trace), else nullptr
};
__device__ __forceinline__ void dpx_stamp(const DpxArgs &a, int i) {
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[i] = t;
  }
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of this rank and all ranks: CTA 0 collects the local arrivals, exchanges flags with the peers, then
// releases the local CTAs. go values are 2 * seq + phase, so they only ever grow. Returns false (to every thread of
// every CTA of this rank) when a wait ran out of patience or an earlier barrier did: the status word local[2] is
// sticky, the caller skips its reduce / apply, and the host raises when it reads the word. The leader's waits are
// bounded by DPX_SPIN_LIMIT polls; the followers wait for the leader's release, which always comes (with the
// status set if need be), so their own, much larger bound only guards against a leader that never ran.
__device__ bool dpx_barrier(const DpxArgs &a, int phase) {
  __syncthreads();
  const uint32_t go = 2u * a.seq + (uint32_t)phase;
  if (blockIdx.x == 0) {     // block-uniform branch
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(a.local, 1u);
      uint32_t spins = 0;
      const uint32_t want = gridDim.x * (2u * (a.seq - 1u) + (uint32_t)phase + 1u);   // the counter never resets
      while (ld_acquire_gpu(a.local) < want)
        if (++spins > DPX_SPIN_LIMIT) { atomicCAS(a.local + 2, 0u, 1u); break; }
      __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x < a.world) {   // one thread per peer: the stores and the polls overlap
      const int q = threadIdx.x;
      uint32_t *f = reinterpret_cast<uint32_t *>(const_cast<unsigned char *>(a.peer[q]) + a.flags_off);
      st_release_sys(f + phase * rfm_fm_trainer::DP_MAX_WORLD + a.rank, a.seq);
      const uint32_t *mine = reinterpret_cast<const uint32_t *>(a.peer[a.rank] + a.flags_off);
      uint32_t spins = 0;
      while (ld_acquire_sys(mine + phase * rfm_fm_trainer::DP_MAX_WORLD + q) < a.seq)
        if (++spins > DPX_SPIN_LIMIT) { atomicCAS(a.local + 2, 0u, 2u + (uint32_t)q); break; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      st_release_sys(a.local + 1, go);
    }
  } else if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(a.local, 1u);
    uint64_t spins = 0;
    while (ld_acquire_gpu(a.local + 1) < go)
      if (++spins > 64ull * DPX_SPIN_LIMIT) { atomicCAS(a.local + 2, 0u, 100u); break; }
  }
  const uint32_t status = threadIdx.x == 0 ? ld_acquire_gpu(a.local + 2) : 0u;
  return __syncthreads_or(status != 0u) == 0;
}

template <typename T>
__global__ void __launch_bounds__(DPX_THREADS, 4)
fm_dp_exchange_apply_kernel(const DpxArgs a, T *__restrict__ w0, T *__restrict__ w, T *__restrict__ V,
                            T *__restrict__ vn, int64_t n, int kp, T lr) {
  T *mine = reinterpret_cast<T *>(const_cast<unsigned char *>(a.peer[a.rank]) + a.grad_off);
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.loss_in) {   // the previous step's loss sums ride in the header
    mine[1] = static_cast<T>(a.loss_in[0]);
    mine[2] = static_cast<T>(a.loss_in[1]);
  }
  dpx_stamp(a, 0);
  if (!dpx_barrier(a, 0)) return;   // every rank's gradient is complete (or the exchange is abandoned: status set)
  dpx_stamp(a, 1);
  {
    const int64_t lo = min(a.total, (int64_t)a.rank * a.slice), hi = min(a.total, lo + a.slice);
    using V2 = typename Vec2<T>::type;
    const int64_t pairs = (hi - lo) / 2;       // total and slice are even
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int W = rfm_fm_trainer::DP_MAX_WORLD;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pairs; i += stride) {
      // the loads from all ranks are issued before the first add: one NVLink round trip (~2-3 us) per element
      // instead of one per rank; the sum itself stays in rank order (identical bits whoever owns the slice)
      V2 x[W];
#pragma unroll
      for (int q = 0; q < W; ++q)
        if (q < a.world)
          x[q] = reinterpret_cast<const V2 *>(reinterpret_cast<const T *>(a.peer[q] + a.grad_off) + lo)[i];
      V2 acc = x[0];
#pragma unroll
      for (int q = 1; q < W; ++q)
        if (q < a.world) {
          acc.x += x[q].x;
          acc.y += x[q].y;
        }
      reinterpret_cast<V2 *>(mine + lo)[i] = acc;
    }
  }
  dpx_stamp(a, 2);
  if (!dpx_barrier(a, 1)) return;   // every slice is reduced
  dpx_stamp(a, 3);
  auto at = [&](int64_t i) -> const T * {      // element i of the reduced gradient, in its owner's buffer
    const int64_t owner = min((int64_t)a.world - 1, i / a.slice);
    return reinterpret_cast<const T *>(a.peer[owner] + a.grad_off) + i;
  };
  const int lane = lane_id();
  const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (gwarp == 0 && lane == 0) {
    *w0 += lr * *at(0);
    if (a.loss_out) {
      a.loss_out[0] = static_cast<double>(*at(1));
      a.loss_out[1] = static_cast<double>(*at(2));
    }
  }
  const int64_t v_off = grad_v_off_dev(n);
  constexpr int R = 4;                             // feature rows in flight per warp (peer loads)
  for (int64_t j0 = gwarp; j0 < n; j0 += R * nw) {
    T gwv[R];                                      // the rows' w gradients: peer loads issued with the first V loads
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t j = j0 + r * nw;
      gwv[r] = (lane == 0 && j < n) ? *at(GRAD_W_OFF + j) : T(0);
    }
    for (int f = lane * 2; f < kp; f += 64) {      // same association as row_norms_kernel / apply_grad_rows_kernel
      T g0[R], g1[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t j = j0 + r * nw;
        if (j < n) {
          const T *g = at(v_off + j * kp + f);     // a pair never straddles a slice boundary (both are even)
          g0[r] = g[0];
          g1[r] = g[1];
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t j = j0 + r * nw;
        if (j < n) {
          V[j * kp + f] += lr * g0[r];
          V[j * kp + f + 1] += lr * g1[r];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t j = j0 + r * nw;
      if (j >= n) continue;
      T s = T(0);
      for (int f = lane * 2; f < kp; f += 64) {
        const T v0 = V[j * kp + f], v1 = V[j * kp + f + 1];   // this lane's own writes
        s += v0 * v0 + v1 * v1;
      }
      s = warp_sum(s);
      if (lane == 0) {
        vn[j] = s;
        w[j] += lr * gwv[r];
      }
    }
  }
  dpx_stamp(a, 4);
}

}  // namespace
extern "C" {

int rfm_fm_dp_export(rfm_fm_trainer *t, void *handle_out) {
  RFM_REQUIRE(t && handle_out, "rfm_fm_dp_export: NULL argument");
  rfm_ctx *ctx = t->m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  RFM_REQUIRE(!t->xchg, "rfm_fm_dp_export: already exported");
  static_assert(sizeof(cudaIpcMemHandle_t) == RFM_DP_HANDLE_BYTES, "IPC handle size");
  int64_t n = 0;
  rfm_fm_grad_size(t, &n);
  t->xchg_grad_bytes = ((size_t)n * dsize(t->m->dtype) + 255) / 256 * 256;
  const size_t bytes = 2 * t->xchg_grad_bytes + 256;
  // cudaMalloc, not the stream-ordered pool: pool memory cannot be exported through legacy CUDA IPC
  cudaIpcMemHandle_t h;
  if (ctx->dp_cache && ctx->dp.borrowers == 0) {
    // Reuse is safe: the previous borrower's last exchange kernel has finished on EVERY rank before any rank
    // could read its final losses (they come out of a collective ordered after it), and no rank starts stepping
    // before the post-connect barrier, i.e. after every rank has re-zeroed its region here.
    rfm_ctx::DpRegion &c = ctx->dp;
    if (c.base && c.bytes < bytes) {   // outgrown: peers may still map it, so it is retired, not freed
      c.retired.push_back(c.base);
      c.base = nullptr;
      c.bytes = 0;
    }
    if (!c.base) {
      RFM_CUDA(cudaMalloc(reinterpret_cast<void **>(&c.base), bytes));
      c.bytes = bytes;
      RFM_CUDA(cudaIpcGetMemHandle(&h, c.base));
      memcpy(c.own_handle, &h, sizeof(h));
    }
    memcpy(&h, c.own_handle, sizeof(h));
    t->xchg = c.base;
    t->xchg_borrowed = true;
    c.borrowers = 1;
  } else {
    RFM_CUDA(cudaMalloc(reinterpret_cast<void **>(&t->xchg), bytes));
    RFM_CUDA(cudaIpcGetMemHandle(&h, t->xchg));
  }
  RFM_CUDA(cudaMemsetAsync(t->xchg, 0, bytes, ctx->stream));
  RFM_TRY(t->dp_prev_loss.alloc(2));
  RFM_TRY(t->dp_local.alloc(4 + 16));     // + 8 x uint64 of RFM_DPX_TRACE time stamps
  RFM_CUDA(cudaMemsetAsync(t->dp_local.p, 0, (4 + 16) * sizeof(uint32_t), ctx->stream));
  RFM_CUDA(cudaMemsetAsync(t->dp_prev_loss.p, 0, 2 * sizeof(double), ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(handle_out, &h, sizeof(h));
  return RFM_OK;
}

int rfm_fm_dp_connect(rfm_fm_trainer *t, int32_t rank, int32_t world, const void *all_handles) {
  RFM_REQUIRE(t && all_handles, "rfm_fm_dp_connect: NULL argument");
  RFM_REQUIRE(t->xchg, "rfm_fm_dp_connect: call rfm_fm_dp_export first");
  RFM_REQUIRE(world >= 1 && world <= rfm_fm_trainer::DP_MAX_WORLD && rank >= 0 && rank < world,
              "rfm_fm_dp_connect: rank %d / world %d out of range (at most %d ranks)", rank, world,
              rfm_fm_trainer::DP_MAX_WORLD);
  RFM_REQUIRE(t->dp_world == 0, "rfm_fm_dp_connect: already connected");
  RFM_CUDA(cudaSetDevice(t->m->ctx->device));
  rfm_ctx::DpRegion &c = t->m->ctx->dp;
  if (t->xchg_borrowed && (c.world != world || c.rank != rank)) {   // another group layout: drop every mapping
    for (int q = 0; q < rfm_ctx::DpRegion::MAX_WORLD; ++q) {
      if (q != c.rank && c.peer[q]) cudaIpcCloseMemHandle(c.peer[q]);
      c.peer[q] = nullptr;
    }
    cudaGetLastError();
    c.world = world;
    c.rank = rank;
  }
  for (int q = 0; q < world; ++q) {
    if (q == rank) {
      t->peer_base[q] = t->xchg;
      if (t->xchg_borrowed) c.peer[q] = t->xchg;
      continue;
    }
    const unsigned char *hq = static_cast<const unsigned char *>(all_handles) + (size_t)q * sizeof(cudaIpcMemHandle_t);
    if (t->xchg_borrowed && c.peer[q]) {
      if (memcmp(hq, c.peer_handle[q], sizeof(cudaIpcMemHandle_t)) == 0) {   // same region as last time: stay mapped
        t->peer_base[q] = c.peer[q];
        continue;
      }
      cudaIpcCloseMemHandle(c.peer[q]);                                       // the peer re-allocated
      cudaGetLastError();
      c.peer[q] = nullptr;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, hq, sizeof(h));
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(RFM_ERR_CUDA, "rfm_fm_dp_connect: cudaIpcOpenMemHandle(rank %d) failed: %s", q, cudaGetErrorString(e));
    }
    t->peer_base[q] = static_cast<unsigned char *>(p);
    if (t->xchg_borrowed) {
      c.peer[q] = t->peer_base[q];
      memcpy(c.peer_handle[q], hq, sizeof(cudaIpcMemHandle_t));
    }
  }
  t->dp_rank = rank;
  t->dp_world = world;
  return RFM_OK;
}

int rfm_fm_dp_exchange_apply(rfm_fm_trainer *t, double lr) {
  RFM_REQUIRE(t && t->dp_world > 0, "rfm_fm_dp_exchange_apply: call rfm_fm_dp_connect first");
  rfm_fm *m = t->m;
  rfm_ctx *ctx = m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  DpxArgs a;
  memset(&a, 0, sizeof(a));
  for (int q = 0; q < t->dp_world; ++q) a.peer[q] = t->peer_base[q];
  a.grad_off = (size_t)t->dp_parity * t->xchg_grad_bytes;
  a.flags_off = 2 * t->xchg_grad_bytes;
  a.rank = t->dp_rank;
  a.world = t->dp_world;
  a.seq = ++t->dp_seq;
  a.total = grad_total(m->n, m->kp);
  a.slice = (ceil_div(a.total, (int64_t)t->dp_world) + 1) / 2 * 2;
  a.local = t->dp_local.p;
  a.trace = getenv("RFM_DPX_TRACE") ? reinterpret_cast<unsigned long long *>(t->dp_local.p + 4) : nullptr;
  a.loss_in = t->dp_pending_loss ? t->loss_sums.p : nullptr;
  a.loss_out = t->dp_prev_loss.p;
  // every CTA must be resident at once (the kernel spins on flags): one wave, four CTAs per SM
  const int g = (int)std::min<int64_t>(4 * (int64_t)ctx->sm_count, std::max<int64_t>(1, ceil_div(m->n, DPX_THREADS / 32)));
  if (m->dtype == RFM_F64) {
    auto fm_dp_exchange_apply = fm_dp_exchange_apply_kernel<double>;
    RFM_LAUNCH(ctx, fm_dp_exchange_apply, g, DPX_THREADS, 0, a, reinterpret_cast<double *>(m->w0.p),
               reinterpret_cast<double *>(m->w.p), reinterpret_cast<double *>(m->V.p),
               reinterpret_cast<double *>(m->vn.p), m->n, m->kp, lr);
  } else {
    auto fm_dp_exchange_apply = fm_dp_exchange_apply_kernel<float>;
    RFM_LAUNCH(ctx, fm_dp_exchange_apply, g, DPX_THREADS, 0, a, reinterpret_cast<float *>(m->w0.p),
               reinterpret_cast<float *>(m->w.p), reinterpret_cast<float *>(m->V.p),
               reinterpret_cast<float *>(m->vn.p), m->n, m->kp, (float)lr);
  }
  ++m->version;
  t->dp_parity ^= 1;           // the next gradient goes to the other buffer
  t->dp_pending_loss = true;   // rfm_fm_loss_sums of this step will ride in the next exchange
  return RFM_OK;
}

int rfm_fm_dp_trace(rfm_fm_trainer *t, uint64_t *stamps_ns) {
  RFM_REQUIRE(t && stamps_ns && t->dp_local.p, "rfm_fm_dp_trace: not exported");
  rfm_ctx *ctx = t->m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  RFM_CUDA(cudaMemcpyAsync(stamps_ns, t->dp_local.p + 4, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return RFM_OK;
}

int rfm_fm_dp_prev_loss_ptr_dev(rfm_fm_trainer *t, void **sums_dev, void **status_dev) {
  RFM_REQUIRE(t && sums_dev && t->dp_prev_loss.p, "rfm_fm_dp_prev_loss_ptr_dev: not exported");
  *sums_dev = t->dp_prev_loss.p;
  if (status_dev) *status_dev = t->dp_local.p + 2;
  return RFM_OK;
}

}  // extern "C" (pause)
namespace {
template <typename T>
int opt_epoch_impl(rfm_fm_trainer *t, int64_t batch, int64_t slot, bool sampled, const FeistelKey &key,
                   const rfm_optimizer &opt) {
  rfm_fm *m = t->m;
  rfm_ctx *ctx = m->ctx;
  RFM_TRY((step_core<T, true>(t, batch, 0.0, sampled, key, 0)));       // dense gradient of the whole batch
  const bool adam = opt.kind == RFM_OPT_ADAM;
  const size_t total = (size_t)grad_total(m->n, m->kp);
  if (adam && !t->adam_m.p) {
    RFM_TRY(t->adam_m.alloc(total * sizeof(T)));
    RFM_TRY(t->adam_v.alloc(total * sizeof(T)));
    RFM_CUDA(cudaMemsetAsync(t->adam_m.p, 0, total * sizeof(T), ctx->stream));
    RFM_CUDA(cudaMemsetAsync(t->adam_v.p, 0, total * sizeof(T), ctx->stream));
  }
  OptArgs<T> o;
  o.lr = (T)opt.lr;
  o.l2 = (T)opt.l2;
  o.b1 = (T)opt.beta1;
  o.b2 = (T)opt.beta2;
  o.eps = (T)opt.eps;
  o.c1 = adam ? (T)(1.0 / (1.0 - std::pow(opt.beta1, (double)opt.step))) : T(1);
  o.c2 = adam ? (T)(1.0 / (1.0 - std::pow(opt.beta2, (double)opt.step))) : T(1);
  const int g = grid_for(ctx, ceil_div(m->n, ROWS_WARPS), 8);
  T *gr = reinterpret_cast<T *>(t->grad_ptr());
  if (adam) {
    auto fm_adam_step = optimizer_rows_kernel<T, true>;
    RFM_LAUNCH(ctx, fm_adam_step, g, ROWS_THREADS, 0, reinterpret_cast<T *>(m->w0.p), reinterpret_cast<T *>(m->w.p),
               reinterpret_cast<T *>(m->V.p), reinterpret_cast<T *>(m->vn.p), gr, reinterpret_cast<T *>(t->adam_m.p),
               reinterpret_cast<T *>(t->adam_v.p), m->n, m->k, m->kp, GRAD_W_OFF, grad_v_off(m->n), o);
  } else {
    auto fm_sgd_l2_step = optimizer_rows_kernel<T, false>;
    RFM_LAUNCH(ctx, fm_sgd_l2_step, g, ROWS_THREADS, 0, reinterpret_cast<T *>(m->w0.p), reinterpret_cast<T *>(m->w.p),
               reinterpret_cast<T *>(m->V.p), reinterpret_cast<T *>(m->vn.p), gr, (T *)nullptr, (T *)nullptr, m->n,
               m->k, m->kp, GRAD_W_OFF, grad_v_off(m->n), o);
  }
  ++m->version;
  return post_update_losses<T>(t, batch, slot);
}
}  // namespace
extern "C" {

int rfm_fm_train_epoch_opt(rfm_fm_trainer *t, const int64_t *batch_rows, uint32_t seed, uint32_t epoch, int64_t batch,
                           int64_t slot, const rfm_optimizer *opt) {
  RFM_TRY(check_batch(t, batch, slot, "rfm_fm_train_epoch_opt"));
  RFM_REQUIRE(opt, "rfm_fm_train_epoch_opt: optimizer is NULL");
  RFM_REQUIRE(opt->kind == RFM_OPT_SGD || opt->kind == RFM_OPT_ADAM, "rfm_fm_train_epoch_opt: unknown optimizer kind %d",
              opt->kind);
  RFM_REQUIRE(opt->l2 >= 0.0, "rfm_fm_train_epoch_opt: l2 must be >= 0");
  if (opt->kind == RFM_OPT_ADAM)
    RFM_REQUIRE(opt->step >= 1 && opt->beta1 >= 0.0 && opt->beta1 < 1.0 && opt->beta2 >= 0.0 && opt->beta2 < 1.0 &&
                    opt->eps > 0.0,
                "rfm_fm_train_epoch_opt: Adam needs step >= 1, betas in [0, 1), eps > 0");
  rfm_ctx *ctx = t->m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  void *g = nullptr;
  RFM_TRY(rfm_fm_grad_ptr_dev(t, &g));
  FeistelKey key = make_feistel_key(1, 0, 0);
  const bool sampled = batch_rows == nullptr;
  if (sampled) {
    RFM_REQUIRE(t->train->n_rows <= (1LL << 32), "rfm_fm_train_epoch_opt: at most 2^32 rows");
    key = make_feistel_key((uint64_t)t->train->n_rows, seed, epoch);
  } else {
    for (int64_t q = 0; q < batch; ++q)
      RFM_REQUIRE(batch_rows[q] >= 0 && batch_rows[q] < t->train->n_rows,
                  "rfm_fm_train_epoch_opt: row id %lld out of range", (long long)batch_rows[q]);
    RFM_TRY(stage_batch(t, batch_rows, batch));
  }
  return t->m->dtype == RFM_F64 ? opt_epoch_impl<double>(t, batch, slot, sampled, key, *opt)
                                : opt_epoch_impl<float>(t, batch, slot, sampled, key, *opt);
}

int rfm_fm_loss_sums_ptr_dev(rfm_fm_trainer *t, void **sums_dev) {
  RFM_REQUIRE(t && sums_dev, "rfm_fm_loss_sums_ptr_dev: NULL argument");
  *sums_dev = t->loss_sums.p;
  return RFM_OK;
}

int rfm_fm_loss_sums(rfm_fm_trainer *t, const int64_t *batch_rows, int64_t batch, int64_t val_begin,
                     int64_t val_end) {
  RFM_REQUIRE(t, "rfm_fm_loss_sums: trainer is NULL");
  rfm_ctx *ctx = t->m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  RFM_REQUIRE(batch >= 0 && batch <= t->max_batch, "rfm_fm_loss_sums: bad batch");
  const int64_t nv = t->val ? t->val->n_rows : 0;
  RFM_REQUIRE(val_begin >= 0 && val_begin <= val_end && val_end <= nv, "rfm_fm_loss_sums: bad val range");
  if (batch > 0 && batch_rows) RFM_TRY(stage_batch(t, batch_rows, batch));  // NULL: reuse the batch on the device
  if (batch > 0) {       // one launch for both sums
    if (t->m->dtype == RFM_F64)
      return batch_and_val_losses<double>(t, batch, 1.0, t->loss_sums.p, val_begin, val_end, 1.0, t->loss_sums.p + 1);
    return batch_and_val_losses<float>(t, batch, 1.0, t->loss_sums.p, val_begin, val_end, 1.0, t->loss_sums.p + 1);
  }
  if (t->m->dtype == RFM_F64) {
    RFM_TRY(loss_pass<double>(t, t->train, t->idx.p, 0, batch, 1.0, t->loss_sums.p));
    RFM_TRY(loss_pass<double>(t, t->val ? t->val : t->train, nullptr, val_begin, val_end - val_begin, 1.0,
                              t->loss_sums.p + 1));
  } else {
    RFM_TRY(loss_pass<float>(t, t->train, t->idx.p, 0, batch, 1.0, t->loss_sums.p));
    RFM_TRY(loss_pass<float>(t, t->val ? t->val : t->train, nullptr, val_begin, val_end - val_begin, 1.0,
                             t->loss_sums.p + 1));
  }
  return RFM_OK;
}

int rfm_fm_trainer_losses(rfm_fm_trainer *t, int64_t first_slot, int64_t n_slots, double *train_loss,
                          double *val_loss) {
  RFM_REQUIRE(t, "rfm_fm_trainer_losses: trainer is NULL");
  RFM_REQUIRE(first_slot >= 0 && n_slots >= 0 && first_slot + n_slots <= t->max_slots,
              "rfm_fm_trainer_losses: slot range out of bounds");
  rfm_ctx *ctx = t->m->ctx;
  RFM_CUDA(cudaSetDevice(ctx->device));
  if (train_loss)
    RFM_CUDA(cudaMemcpyAsync(train_loss, t->losses.p + first_slot, (size_t)n_slots * 8, cudaMemcpyDeviceToHost,
                             ctx->stream));
  if (val_loss)
    RFM_CUDA(cudaMemcpyAsync(val_loss, t->losses.p + t->max_slots + first_slot, (size_t)n_slots * 8,
                             cudaMemcpyDeviceToHost, ctx->stream));
  RFM_CUDA(cudaStreamSynchronize(ctx->stream));
  return t->m->dtype == RFM_F64 ? t->sort64.check(ctx) : t->sort32.check(ctx);
}

}  // extern "C"
