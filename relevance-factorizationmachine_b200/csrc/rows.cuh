// rows.cuh -- the row-set handle shared by fm.cu (training / scoring kernels) and synth.cu (device-side generator).
#pragma once
#include "common.cuh"

using namespace rfm;

// Factored rows: what the reference's data layer holds BEFORE scipy.sparse.hstack
// (utils/dataloader/coat/_preparer.py:154-170, kuairec/_feature.py:169-209) -- per-entity feature tables plus one
// (user, item[, context]) record per interaction. A row is the concatenation, in column order, of up to FAC_MAX_SEG
// blocks: the one-hot of an id, the table row of an id, or dense per-row context values. The row kernels assemble
// x_t on the fly, so an interaction costs 8 + 8 n_ctx + 8 bytes of HBM instead of 12 m + 16.
constexpr int FAC_MAX_SEG = 6;
enum FacKind { SEG_ID = 0, SEG_TABLE = 1, SEG_CTX = 2 };
struct FacSegDev {
  int kind, key;            // key: 0 = the row's user id, 1 = its item id (SEG_ID, SEG_TABLE)
  uint32_t col0;            // first global column of the block
  int width;                // SEG_CTX: number of columns
  const int32_t *ptr;       // SEG_TABLE: CSR of the table, columns local to the block
  const int32_t *col;
  const void *val;          // T
  int ctx0;                 // SEG_CTX: first column of the block inside a row's context record
  int pad;
};
struct FacDev {
  int n_seg, n_ctx, es, pad;  // es: bytes per value (4 or 8)
  const int32_t *user, *item;
  const void *ctx;          // T [n_rows][n_ctx]
  const void *one;          // T [1] = 1.0: what an id block's entry loads as its value (keeps the entry fetch branch-free)
  FacSegDev seg[FAC_MAX_SEG];
};

struct rfm_csr {
  rfm_ctx *ctx = nullptr;
  int dtype = RFM_F64;
  int64_t n_rows = 0, n_cols = 0, nnz = 0, max_row_len = 0;
  bool has_targets = false;
  DevBuf<int64_t> row_ptr;
  DevBuf<int32_t> col;
  DevBuf<unsigned char> val, yp;
  // factored rows (rfm_factored_create): the CSR buffers above stay empty
  bool factored = false;
  int n_seg = 0, n_ctx = 0;
  DevBuf<int32_t> f_user, f_item;
  DevBuf<unsigned char> f_ctx, f_one;
  DevBuf<signed char> g_label, g_relevance;   // generated rows only (rfm_factored_generate with keep_labels)
  struct Seg {
    int kind = 0, key = 0, width = 0, ctx0 = 0;
    uint32_t col0 = 0;
    int64_t n_entities = 0, tnz = 0;   // SEG_TABLE: rows and stored entries of the table
    DevBuf<int32_t> ptr, col;
    DevBuf<unsigned char> val;
  } seg[FAC_MAX_SEG];
  FacDev fac_dev(int64_t row_offset = 0) const {
    FacDev f;
    memset(&f, 0, sizeof(f));
    f.n_seg = n_seg;
    f.n_ctx = n_ctx;
    f.es = dtype == RFM_F64 ? 8 : 4;
    f.user = f_user.p + row_offset;
    f.item = f_item.p + row_offset;
    f.ctx = f_ctx.p ? f_ctx.p + (size_t)row_offset * n_ctx * (dtype == RFM_F64 ? 8 : 4) : nullptr;
    f.one = f_one.p;
    for (int s = 0; s < n_seg; ++s) {
      f.seg[s].kind = seg[s].kind;
      f.seg[s].key = seg[s].key;
      f.seg[s].col0 = seg[s].col0;
      f.seg[s].width = seg[s].width;
      f.seg[s].ptr = seg[s].ptr.p;
      f.seg[s].col = seg[s].col.p;
      f.seg[s].val = seg[s].val.p;
      f.seg[s].ctx0 = seg[s].ctx0;
    }
    return f;
  }
};


// synth.cu: fills user / item / context / target arrays of n_rows generated interactions (rfm_factored_generate)
int rfm_synth_fill_rows(rfm_ctx *ctx, const rfm_click_model *model, int64_t n_rows, int32_t *user_dev, int32_t *item_dev,
                        void *ctx_dev, int n_ctx, void *targets_dev, signed char *labels_dev, signed char *relevance_dev,
                        int dtype);
