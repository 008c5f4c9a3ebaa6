// two_level.cuh -- the two-level FM step for factored rows (included by fm.cu, inside its anonymous namespace).
//
// A factored row is x_t = [ blocks keyed by the row's user | blocks keyed by its item | per-row context values ]
// (rows.cuh; the reference stacks exactly these pieces: utils/dataloader/coat/_preparer.py:154-170,
// kuairec/_feature.py:169-209). Everything the FM step (src/fm.py:80-88, 125-132, 156-187) computes from the
// user-keyed part of a row depends on the user alone, and likewise for the item:
//
//   forward   s_t = A_u + C_i + sum_ctx x v,       A_u = sum_{j in user blocks} x_uj v_j   (one k-vector per USER)
//             x_t.w = alpha_u + beta_i + ...,      alpha_u = sum x_uj w_j,  q_u = sum x_uj^2 ||v_j||^2
//   backward  grad v_j = sum_t e_t (x_tj s_t - x_tj^2 v_j)
//                      = sum_u x_uj R_u - v_j sum_u x_uj^2 E_u,   R_u = sum_{t of u} e_t s_t,  E_u = sum_{t of u} e_t
//
// so a step touches 2 + n_ctx "virtual columns" per interaction (the user, the item, the context columns) instead of
// the m stored non-zeros, plus one pass over the entity tables per step that does not grow with the batch:
//
//   entity forward (A, alpha, q for every user / item; context columns are copied)     fm_entity_fwd_kernel
//   row pass on virtual rows [one-hot user | one-hot item | context] with (A, alpha, q) as the parameter table
//   sort by virtual column, level-1 column pass: R_v, sum e x, sum e x^2 per virtual column        (OUT_RAW)
//   level-2 column pass over the STATIC list (real column j, virtual column v, x), sorted by j once at set-up:
//     segmented sums of x R_v -> the reference's gradient of column j -> SGD (or the data-parallel gradient buffer)
//   entity forward again (post-update parameters; also the next step's), loss pass on the virtual rows
//
// The sums are the same real numbers as the flat step's, associated differently (per entity first): results agree
// with the flat path to rounding (1e-13 relative over a fit), not bit for bit; every association is still fixed, so
// a fit is bit-reproducible run to run. It pays when batch * (m - 2 - n_ctx) exceeds the entity tables' entries
// (KuaiRec shape: 1.08 M gathered rows per pass -> 0.20 M + 0.14 M); rfm_fm_trainer_set_two_level decides.
#pragma once

struct TlCtxCols {
  uint32_t c[32];       // real column of context column i (position inside a row's context record)
};

struct TwoLevel {
  int64_t n_ent[2] = {0, 0};   // users, items the blocks agree on
  int64_t nv = 0;              // virtual columns: [users | items | context columns]
  uint32_t stride = 0;         // entries of a virtual row
  int64_t m2 = 0;              // entries of the entity lists = of the static level-2 list
  TlCtxCols ctx_cols;
  bool lean = false;           // fm_vrows_kernel serves the row passes (both keys present, n_ctx <= 1)
  bool val_virtual = false;    // val rows are keyed by the same tables: their loss rides in the virtual loss launch
  DevBuf<unsigned char> Vv, wv, vnv;    // aggregated parameter table [nv][kp], [nv], [nv]
  DevBuf<unsigned char> R;              // level-1 sums [nv][kp] | a [nv] | c [nv]   (zeroed every step)
  DevBuf<unsigned char> Cq, cpart;      // lean, one context column: c_q per batch position; per-CTA sums of the column
  DevBuf<uint32_t> ctx_ticket, n_tails2;   // n_tails2: tail counters of level 1 and level 2 (both zeroed at the step's start)
  DevBuf<uint32_t> ent_ptr, m2_dev;     // entity lists: CSR by virtual column
  DevBuf<int32_t> ent_col;
  DevBuf<unsigned char> ent_val;
  RadixSorter<float> l2_32;             // the same entries sorted (stably) by real column
  RadixSorter<double> l2_64;
  int l2_buf = 0;
  uint64_t agg_version = ~0ull;         // rfm_fm::version the aggregated table corresponds to
};

// descriptor of the virtual rows over the same (user, item, context) records
inline FacDev tl_virtual_fac(const FacDev &f, const TwoLevel &L) {
  FacDev v = f;
  memset(v.seg, 0, sizeof(v.seg));
  int s = 0;
  uint32_t col0 = 0;
  for (int key = 0; key < 2; ++key) {
    if (L.n_ent[key] == 0) continue;
    v.seg[s].kind = SEG_ID;
    v.seg[s].key = key;
    v.seg[s].col0 = col0;
    v.seg[s].width = (int)L.n_ent[key];
    col0 += (uint32_t)L.n_ent[key];
    ++s;
  }
  if (f.n_ctx > 0) {
    v.seg[s].kind = SEG_CTX;
    v.seg[s].col0 = col0;
    v.seg[s].width = f.n_ctx;
    v.seg[s].ctx0 = 0;
    ++s;
  }
  v.n_seg = s;
  return v;
}

__device__ __forceinline__ void tl_decode(int64_t v, int64_t n_user, int64_t n_item, int &key, int &e) {
  if (v < n_user) {
    key = 0;
    e = static_cast<int>(v);
  } else if (v < n_user + n_item) {
    key = 1;
    e = static_cast<int>(v - n_user);
  } else {
    key = 2;
    e = static_cast<int>(v - n_user - n_item);
  }
}

__global__ void tl_ent_len_kernel(const FacDev f, int64_t n_user, int64_t n_item, int64_t nv, uint32_t *__restrict__ len) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    int key, e;
    tl_decode(v, n_user, n_item, key, e);
    uint32_t n = 0;
    if (key == 2) {
      n = 1;
    } else {
      for (int s = 0; s < f.n_seg; ++s) {
        const FacSegDev &g = f.seg[s];
        if (g.kind == SEG_CTX || g.key != key) continue;
        n += g.kind == SEG_ID ? 1u : static_cast<uint32_t>(g.ptr[e + 1] - g.ptr[e]);
      }
    }
    len[v] = n;
  }
}

// entity v's entries in block order (the order the stacked row holds them): (real column, x); the same entries
// as (key = real column, pos = v, x) triples for the one-time sort by real column
template <typename T>
__global__ void tl_ent_fill_kernel(const FacDev f, int64_t n_user, int64_t n_item, int64_t nv, const TlCtxCols cc,
                                   const uint32_t *__restrict__ ent_ptr, int32_t *__restrict__ ent_col,
                                   T *__restrict__ ent_val, uint32_t *__restrict__ keys, uint32_t *__restrict__ pos,
                                   T *__restrict__ xs) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    int key, e;
    tl_decode(v, n_user, n_item, key, e);
    uint32_t at = ent_ptr[v];
    auto put = [&](uint32_t c, T x) {
      ent_col[at] = static_cast<int32_t>(c);
      ent_val[at] = x;
      keys[at] = c;
      pos[at] = static_cast<uint32_t>(v);
      xs[at] = x;
      ++at;
    };
    if (key == 2) {
      put(cc.c[e], T(1));
      continue;
    }
    for (int s = 0; s < f.n_seg; ++s) {
      const FacSegDev &g = f.seg[s];
      if (g.kind == SEG_CTX || g.key != key) continue;
      if (g.kind == SEG_ID) {
        put(g.col0 + static_cast<uint32_t>(e), T(1));
      } else {
        const T *val = static_cast<const T *>(g.val);
        for (int p = g.ptr[e]; p < g.ptr[e + 1]; ++p) put(g.col0 + static_cast<uint32_t>(g.col[p]), val[p]);
      }
    }
  }
}

// word-wise comparison of two device arrays (are the val rows keyed by the same tables as the train rows?)
__global__ void tl_compare_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, int64_t n_words,
                                  int *__restrict__ differ) {
  bool mine = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x)
    mine |= a[i] != b[i];
  if (mine) atomicOr(differ, 1);
}

// Entity forward: a group of TPR lanes per virtual column sums x v_j, x w_j and x^2 ||v_j||^2 over the entity's
// entries, in entry order (the same lane layout and per-entry order as the row pass).
template <typename T, int TPR, int NCV>
__global__ void __launch_bounds__(ROWS_THREADS, NCV <= 4 ? 4 : 1)
fm_entity_fwd_kernel(const uint32_t *__restrict__ ent_ptr, const int32_t *__restrict__ ent_col,
                     const T *__restrict__ ent_val, int64_t nv, const T *__restrict__ V, const T *__restrict__ w,
                     const T *__restrict__ vn, int kp, T *__restrict__ Vv, T *__restrict__ wv, T *__restrict__ vnv,
                     int pdl_release) {
  pdl_wait_and_release(pdl_release != 0);
  using V2 = typename Vec2<T>::type;
  constexpr int GPW = 32 / TPR;
  const int lane = lane_id(), g = lane % TPR, grp = lane / TPR;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t vb = gw * GPW; vb < nv; vb += nw * GPW) {     // warp-uniform trip count
    const int64_t v = vb + grp;
    const bool active = v < nv;
    const uint32_t beg = active ? ent_ptr[v] : 0u;
    const int len = active ? static_cast<int>(ent_ptr[v + 1] - beg) : 0;
    const int maxlen = __reduce_max_sync(FULL, len);
    V2 acc[NCV];
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) acc[ch].x = acc[ch].y = T(0);
    T sl = T(0), sq = T(0);
    for (int off0 = 0; off0 < maxlen; off0 += TPR) {
      const int off = off0 + g;
      int c = 0;
      T x = T(0);
      if (off < len) {
        c = ent_col[beg + off];
        x = ent_val[beg + off];
        sl += x * __ldg(w + c);
        sq += (x * x) * __ldg(vn + c);
      }
      const int cnt = maxlen - off0 < TPR ? maxlen - off0 : TPR;
      RFM_UNROLL(RFM_ROWS_UNROLL)
      for (int i = 0; i < cnt; ++i) {
        const int cj = __shfl_sync(FULL, c, i, TPR);
        const T xj = __shfl_sync(FULL, x, i, TPR);
        const V2 *vrow = reinterpret_cast<const V2 *>(V + (size_t)cj * kp) + g;
#pragma unroll
        for (int ch = 0; ch < NCV; ++ch) {
          const V2 r = __ldg(vrow + ch * TPR);
          acc[ch].x += xj * r.x;
          acc[ch].y += xj * r.y;
        }
      }
    }
    sl = group_sum<TPR>(sl, FULL);
    sq = group_sum<TPR>(sq, FULL);
    if (active) {
      V2 *out = reinterpret_cast<V2 *>(Vv + (size_t)v * kp) + g;
#pragma unroll
      for (int ch = 0; ch < NCV; ++ch) out[ch * TPR] = acc[ch];
      if (g == 0) {
        wv[v] = sl;
        vnv[v] = sq;
      }
    }
  }
}

// ---- the row pass on virtual rows [user | item | at most one context column] ---------------------------------------
// The generic factored row pass (fm_rows_kernel<FAC>) spends most of its instructions decoding block layouts; a
// virtual row needs none of that: two table rows (A_u, C_i) are added, the context column's vector is scaled in, and
// the scalar part is three look-ups. Same lane layout, same order of every floating-point operation as the generic
// kernel on the virtual descriptor (entries in the order user, item, context), so the two are interchangeable bit
// for bit; the generic one remains for rows with several context columns or a single key.
template <typename T>
struct VRowsArgs {
  const int32_t *user, *item;     // per train row
  const T *ctx;                   // [n_rows] (NCTX == 1)
  const T *yp;
  const int64_t *idx;             // batch row ids, or nullptr: rows [row0, row0 + n)
  int64_t row0, n;
  const T *w0, *A, *wv, *vnv;     // aggregated table: users [0, U), items [U, U + I), the context column at U + I
  uint32_t n_user, ctx_col;       // ctx_col = U + I
  int kp;
  T *S, *E;
  T *Cq;                          // NCTX == 1, train: the row's context value by batch position (summed in level 1)
  uint32_t stride, sentinel;
  uint32_t *keys, *pos;
  T *xs;
  uint32_t *ghist;
  int n_passes;
  FeistelKey fkey;
  int64_t q0;
  int64_t *idx_out;
  Finish fin;
  int pdl_release;                // the next operation of the stream is a kernel that waits
  // MODE_LOSS: optional second row set (val rows keyed by the same tables)
  const int32_t *user2, *item2;
  const T *ctx2, *yp2;
  int64_t n2;
  Finish fin2;
};

#ifndef RFM_TL_VROWS_BLOCKS
#define RFM_TL_VROWS_BLOCKS 4
#endif
constexpr int TL_VROWS_BLOCKS = RFM_TL_VROWS_BLOCKS;     // resident CTAs per SM; the grid is exactly one wave of them

template <typename T, int TPR, int NCV, int MODE, bool SAMPLED, int NCTX>
__global__ void __launch_bounds__(ROWS_THREADS, NCV <= 4 ? TL_VROWS_BLOCKS : 1)
fm_vrows_kernel(const VRowsArgs<T> a) {
  pdl_wait_and_release(a.pdl_release != 0);
  using V2 = typename Vec2<T>::type;
  constexpr int GPW = 32 / TPR;
  const int lane = lane_id(), g = lane % TPR, grp = lane / TPR;
  const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_groups = nw * GPW;
  const T w0 = __ldg(a.w0);
  __shared__ uint32_t hist[MODE == MODE_TRAIN ? RS_MAX_PASSES * RS_RADIX : 1];
  if (MODE == MODE_TRAIN) {
    for (int i = threadIdx.x; i < RS_MAX_PASSES * RS_RADIX; i += ROWS_THREADS) hist[i] = 0;
    __syncthreads();
  }
  // the context column's vector and scalars stay in registers for the whole kernel
  V2 vc[NCTX ? NCV : 1];
  T wc = T(0), vnc = T(0);
  if (NCTX) {
    const V2 *crow = reinterpret_cast<const V2 *>(a.A + (size_t)a.ctx_col * a.kp) + g;
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) vc[ch] = __ldg(crow + ch * TPR);
    wc = __ldg(a.wv + a.ctx_col);
    vnc = __ldg(a.vnv + a.ctx_col);
  }
  double partial = 0.0, partial2 = 0.0;
  const int64_t n_total = a.n + (MODE == MODE_LOSS ? a.n2 : 0);
  struct RowMeta {
    int u, i;
    T c, ypv;
    bool active, second;
  };
  auto fetch = [&](int64_t qq) {
    RowMeta r;
    r.u = r.i = 0;
    r.c = r.ypv = T(0);
    r.active = qq < n_total;
    r.second = MODE == MODE_LOSS && qq >= a.n;
    if (r.active) {
      if (r.second) {
        const int64_t t = qq - a.n;
        r.u = a.user2[t];
        r.i = a.item2[t];
        if (NCTX) r.c = a.ctx2[t];
        r.ypv = a.yp2[t];
      } else {
        int64_t t;
        if (SAMPLED) {
          t = static_cast<int64_t>(feistel_permute(static_cast<uint64_t>(a.q0 + qq), a.fkey));
          if (g == 0) a.idx_out[qq] = t;
        } else {
          t = a.idx ? a.idx[qq] : a.row0 + qq;
        }
        r.u = a.user[t];
        r.i = a.item[t];
        if (NCTX) r.c = a.ctx[t];
        if (MODE != MODE_PREDICT) r.ypv = a.yp[t];
      }
    }
    return r;
  };
  RowMeta nxt = fetch(gw * GPW + grp);
  for (int64_t qb = gw * GPW; qb < n_total; qb += n_groups) {   // warp-uniform trip count
    const int64_t q = qb + grp;
    const RowMeta cur = nxt;
    nxt = fetch(q + n_groups);
    const uint32_t cu = static_cast<uint32_t>(cur.u), ci = a.n_user + static_cast<uint32_t>(cur.i);
    const V2 *ur = reinterpret_cast<const V2 *>(a.A + (size_t)cu * a.kp) + g;
    const V2 *ir = reinterpret_cast<const V2 *>(a.A + (size_t)ci * a.kp) + g;
    V2 acc[NCV];
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) {
      const V2 x = __ldg(ur + ch * TPR), y = __ldg(ir + ch * TPR);
      acc[ch].x = (T(0) + x.x) + y.x;          // entries in the order user, item, context (x = 1, 1, c)
      acc[ch].y = (T(0) + x.y) + y.y;
      if (NCTX) {
        acc[ch].x += cur.c * vc[ch].x;
        acc[ch].y += cur.c * vc[ch].y;
      }
    }
    // scalar part of the logit: lane g of the group owns entry g, as in the generic kernel
    T sl = T(0);
    if (g == 0) sl += T(1) * __ldg(a.wv + cu) - T(0.5) * (T(1) * T(1)) * __ldg(a.vnv + cu);
    if (g == 1) sl += T(1) * __ldg(a.wv + ci) - T(0.5) * (T(1) * T(1)) * __ldg(a.vnv + ci);
    if (NCTX && g == 2 && cur.c != T(0)) sl += cur.c * wc - T(0.5) * (cur.c * cur.c) * vnc;
    T ss = T(0);
#pragma unroll
    for (int ch = 0; ch < NCV; ++ch) ss += acc[ch].x * acc[ch].x + acc[ch].y * acc[ch].y;
    const T z = w0 + group_sum<TPR>(sl + T(0.5) * ss, FULL);
    const double p = sigmoid_ref(static_cast<double>(z));
    if (cur.active) {
      if (MODE == MODE_TRAIN) {
        const double e = static_cast<double>(cur.ypv) - p;
        V2 *srow = reinterpret_cast<V2 *>(a.S + (size_t)q * a.kp) + g;
#pragma unroll
        for (int ch = 0; ch < NCV; ++ch) srow[ch * TPR] = acc[ch];
        if (g == 0) {
          a.E[q] = static_cast<T>(e);
          partial += e;
        }
        if (g < 2) {                                      // the sort sees the user and the item entry (x = 1); the
          const uint32_t key = g == 0 ? cu : ci;          // context column is dense and summed by position in level 1
          const uint32_t o = static_cast<uint32_t>(q) * 2u + static_cast<uint32_t>(g);
          a.keys[o] = key;
          a.pos[o] = static_cast<uint32_t>(q);
          a.xs[o] = T(1);
          for (int ps = 0; ps < a.n_passes; ++ps) atomicAdd(&hist[ps * RS_RADIX + ((key >> (8 * ps)) & 0xFF)], 1u);
        }
        if (NCTX && g == 2) a.Cq[q] = cur.c;
      } else if (MODE == MODE_LOSS) {
        if (g == 0) {
          const double r = static_cast<double>(cur.ypv);
          const double term = r * log(p + 1e-8) + (1.0 - r) * log(1.0 - p + 1e-8);
          if (cur.second) partial2 -= term; else partial -= term;
        }
      }
    }
  }
  if (MODE == MODE_TRAIN) {
    __syncthreads();
    for (int i = threadIdx.x; i < a.n_passes * RS_RADIX; i += ROWS_THREADS)
      if (hist[i]) atomicAdd(a.ghist + i, hist[i]);
  }
  block_finish<T>(warp_sum(partial), a.fin);
  if (MODE == MODE_LOSS && a.n2 > 0) block_finish<T>(warp_sum(partial2), a.fin2);
}
