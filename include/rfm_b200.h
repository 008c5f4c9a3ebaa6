/* rfm_b200.h -- C ABI of the B200-native FM / MF training and ranking-evaluation hot path.
 *
 * The reference (tatsuki1107/Relevance-FactorizationMachine) is pure Python and has no FFI:
 * its boundary is the Python class API. Each entry point below names the reference
 * interface (file:line under the reference root) it stands in for; the Python shim in
 * relevance-factorizationmachine_b200/rfm_b200/ binds them with ctypes and re-creates the
 * reference's classes on top.
 *
 * Conventions
 *   - every function returns an int status (RFM_OK == 0); rfm_last_error() gives the text of
 *     the last failure on the calling thread. Nothing here aborts the process.
 *   - pointers are HOST pointers unless the name ends in _dev. Handles own device memory.
 *   - all kernels are enqueued on the context's CUDA stream; calls that return host values
 *     synchronise that stream, the others are asynchronous.
 *   - there is no CPU fallback: without a CUDA device rfm_ctx_create fails with
 *     RFM_ERR_NO_DEVICE.
 */
#ifndef RFM_B200_H_
#define RFM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFM_ABI_VERSION 6   /* 6: rfm_factored_create_range, rfm_rows_device_ptrs, rfm_factored_finalize. 5: rfm_fm_trainer_set_two_level, rfm_fm_dp_trace. 4: rfm_*_set_targets; factored rows, device-chained evaluation, sharded top-K exchange (round 2). 3: rfm_csr_create_range, rfm_csr_device_ptrs added. 2: rfm_topk_run stats grew to int64[4]; rfm_topk_result_host, rfm_fm_dp_*, rfm_fm_train_epoch_opt added */

enum rfm_status {
  RFM_OK = 0,
  RFM_ERR_INVALID = 1,   /* bad argument (mirrors the reference's ValueError paths)      */
  RFM_ERR_CUDA = 2,      /* a CUDA runtime call or kernel failed; see rfm_last_error()    */
  RFM_ERR_NOMEM = 3,
  RFM_ERR_NO_DEVICE = 4
};

/* arithmetic type of parameters and of every kernel that touches them. The reference is
 * float64 throughout; RFM_F64 is the parity mode. */
enum rfm_dtype { RFM_F32 = 0, RFM_F64 = 1 };

typedef struct rfm_ctx rfm_ctx;
typedef struct rfm_csr rfm_csr;
typedef struct rfm_fm rfm_fm;
typedef struct rfm_fm_trainer rfm_fm_trainer;
typedef struct rfm_pairs rfm_pairs;
typedef struct rfm_mf rfm_mf;
typedef struct rfm_ranker rfm_ranker;

/* ---- library / context ------------------------------------------------------------------ */
int rfm_abi_version(void);
const char *rfm_last_error(void);
int rfm_device_count(int *out);

/* cuda_stream: a cudaStream_t (as void*) the caller owns, or NULL for the legacy default
 * stream (which is what torch uses unless told otherwise). */
int rfm_ctx_create(int device, void *cuda_stream, rfm_ctx **out);
int rfm_ctx_destroy(rfm_ctx *ctx);
int rfm_ctx_synchronize(rfm_ctx *ctx);
/* number of kernels this library has launched through ctx (bench.py's gpu_launches). */
int rfm_ctx_launch_count(rfm_ctx *ctx, int64_t *out);
/* device-side timing on the context's stream (cudaEvent pair). */
int rfm_ctx_timer_start(rfm_ctx *ctx);
int rfm_ctx_timer_stop_ms(rfm_ctx *ctx, double *ms);

/* per-kernel timing: between begin and end every launch is bracketed by a cudaEvent pair on the
 * context's stream; end writes "kernel-name<TAB>launches<TAB>total-ms" lines into out_text. */
int rfm_ctx_profile_begin(rfm_ctx *ctx);
int rfm_ctx_profile_end(rfm_ctx *ctx, char *out_text, size_t capacity);

/* page-locked host memory for the shim's staging buffers; register/unregister pin an existing
 * allocation (e.g. a NumPy array) in place so uploads run at full PCIe speed. */
int rfm_host_alloc(size_t bytes, void **out);
int rfm_host_free(void *p);
int rfm_host_register(void *p, size_t bytes);
int rfm_host_unregister(void *p);

/* ---- batch samplers ---------------------------------------------------------------------
 * rfm_legacy_batch: sklearn.utils.resample(replace=False, n_samples=B, random_state=epoch) as
 *   called at src/fm.py:72-79 and src/mf.py:88-95, i.e. RandomState(epoch).shuffle(arange(N))[:B]
 *   (MT19937 + NumPy's legacy Fisher-Yates with masked rejection). Host code, bit-identical
 *   to NumPy. scratch may be NULL or an int32[N] buffer to avoid the allocation.
 * rfm_feistel_batch: this build's perf-mode sampler (not in the reference): first B images of
 *   a keyed bijection of [0,N). Host statement of what the device kernel computes. */
int rfm_legacy_batch(int64_t n_rows, int64_t batch, uint32_t epoch, int64_t *out_rows, int32_t *scratch);
int rfm_feistel_batch(int64_t n_rows, int64_t batch, uint32_t seed, uint32_t epoch, int64_t *out_rows);
/* Host helpers of the "upload only what the fit samples" mode (no reference counterpart). rfm_feistel_batches:
 * positions [begin, begin+count) of the batches of epochs epoch0 .. epoch0+n_epochs-1, epoch-major, on
 * n_threads host threads (0 = all). rfm_csr_gather_rows: the CSR made of the listed rows, in that order
 * (scipy's X[rows]); called with out_indices == NULL it only fills out_indptr[n_sel+1] so that the caller can
 * size the other outputs (out_indptr[n_sel] non-zeros). Pure host code: both run without a GPU. */
int rfm_feistel_batches(int64_t n_rows, int64_t batch, uint32_t seed, uint32_t epoch0, int32_t n_epochs,
                        int64_t begin, int64_t count, int64_t *out_rows, int32_t n_threads);
int rfm_csr_gather_rows(int64_t n_rows, const void *indptr, int indptr_is_int64, const int32_t *indices,
                        const double *data, const int64_t *labels, const double *pscores, const int64_t *rows,
                        int64_t n_sel, int64_t *out_indptr, int32_t *out_indices, double *out_data,
                        int64_t *out_labels, double *out_pscores, int32_t n_threads);

/* ---- FM rows: scipy.sparse.csr_matrix + labels + pscores (the train/val dicts of
 * src/fm.py:55-70; layout from utils/dataloader/coat/_preparer.py:154-170) --------------- */
int rfm_csr_create(rfm_ctx *ctx, int64_t n_rows, int64_t n_cols,
                   const void *indptr, int indptr_is_int64, const int32_t *indices,
                   const double *data,
                   const int64_t *labels,  /* may be NULL (predict-only rows)            */
                   const double *pscores,  /* may be NULL                                */
                   int dtype, rfm_csr **out);
/* Data-parallel upload (no reference counterpart; SURVEY section 8e "train CSR replicated"): the object has the
 * full shape and every row pointer, but only rows [row_begin, row_end) are copied from the host, so G ranks move
 * 1/G of the train set over PCIe each. The caller fills the other ranges of col / val / targets through the
 * device pointers (rfm_b200.dist.sharded_csr_rows: NCCL broadcasts over NVLink) before the rows are used.
 * val elements have the object's dtype; targets are y/pscore in that dtype, one per row. */
int rfm_csr_create_range(rfm_ctx *ctx, int64_t n_rows, int64_t n_cols,
                         const void *indptr, int indptr_is_int64, const int32_t *indices, const double *data,
                         const int64_t *labels, const double *pscores, int dtype,
                         int64_t row_begin, int64_t row_end, rfm_csr **out);
int rfm_csr_device_ptrs(rfm_csr *rows, void **row_ptr_dev /* int64[n_rows+1] */, void **col_dev /* int32[nnz] */,
                        void **val_dev, void **targets_dev);
/* Factored rows: what the reference's data layer holds BEFORE scipy.sparse.hstack --
 * utils/dataloader/coat/_preparer.py:154-170 stacks [onehot_user_ids[u] | user_features[u] | onehot_item_ids[i] |
 * item_features[i]], utils/dataloader/kuairec/_feature.py:169-209 stacks [I_user | I_item | interaction columns |
 * user table | video table]. The rows object is described by those blocks, in column order, plus one
 * (user, item) pair per interaction; the row kernels assemble x_t on the fly, in the hstacked matrix's column
 * order, so every entry point that takes an rfm_csr gives the same bits for both representations while an
 * interaction costs ~40 bytes of PCIe / HBM instead of 12 m + 16.
 *   RFM_BLOCK_ID:    one-hot of the row's user / item id, n_cols ids wide (value 1.0)
 *   RFM_BLOCK_TABLE: row `id` of a CSR table with n_entities rows and n_cols columns (indices local to the block,
 *                    ascending inside a row, as scipy's canonical CSR)
 *   RFM_BLOCK_CTX:   n_cols dense per-interaction values, row-major [n_rows][n_cols] (at most 32 context columns
 *                    over all blocks); a zero contributes no entry, as in scipy's csr_matrix(dense)
 * users / items: int32 or int64 ids, one per row; labels: int8, int32 or int64 (label_bytes). At most 6 blocks. */
#define RFM_BLOCK_ID 0
#define RFM_BLOCK_TABLE 1
#define RFM_BLOCK_CTX 2
#define RFM_KEY_USER 0
#define RFM_KEY_ITEM 1
typedef struct rfm_rows_block {
  int32_t kind, key;
  int64_t n_cols;
  int64_t n_entities;          /* TABLE: rows of the table (ID: n_cols is used) */
  const void *indptr;          /* TABLE */
  int32_t indptr_is_int64, reserved;
  const int32_t *indices;      /* TABLE */
  const double *data;          /* TABLE */
  const double *values;        /* CTX */
} rfm_rows_block;
int rfm_factored_create(rfm_ctx *ctx, int64_t n_rows, const void *users, int32_t users_is_int64, const void *items,
                        int32_t items_is_int64, const rfm_rows_block *blocks, int32_t n_blocks,
                        const void *labels /* may be NULL */, int32_t label_bytes, const double *pscores, int dtype,
                        rfm_csr **out);
/* The same with the propensity given per ITEM instead of per row: pscore[t] = item_pscores[item[t]]. In both of the
 * reference's loaders the propensity is an item-level quantity gathered per row (coat/_preparer.py:56-62,
 * kuairec/loader.py:160-168), so the factored hand-over can stop before that gather too. */
int rfm_factored_create_item_pscores(rfm_ctx *ctx, int64_t n_rows, const void *users, int32_t users_is_int64,
                                     const void *items, int32_t items_is_int64, const rfm_rows_block *blocks,
                                     int32_t n_blocks, const void *labels, int32_t label_bytes,
                                     const double *item_pscores, int64_t n_item_pscores, int dtype, rfm_csr **out);
/* The same rows when each data-parallel rank uploads 1/G of them (no reference counterpart): rows
 * [row_begin, row_end) are copied from the host -- the host pointers are those of the FULL arrays, pscores OR
 * item_pscores is NULL --, the object has the full shape, the caller fills the other rows on the device through
 * rfm_rows_device_ptrs (user int32[n_rows], item int32[n_rows], ctx T[n_rows][n_ctx] or NULL, targets T[n_rows];
 * rfm_b200.dist.sharded_factored_rows broadcasts every slice from its owner over NVLink) and then calls
 * rfm_factored_finalize (row statistics: total non-zeros, longest row). Byte-identical to rfm_factored_create. */
int rfm_factored_create_range(rfm_ctx *ctx, int64_t n_rows, const void *users, int32_t users_is_int64, const void *items,
                              int32_t items_is_int64, const rfm_rows_block *blocks, int32_t n_blocks, const void *labels,
                              int32_t label_bytes, const double *pscores, const double *item_pscores,
                              int64_t n_item_pscores, int dtype, int64_t row_begin, int64_t row_end, rfm_csr **out);
int rfm_rows_device_ptrs(rfm_csr *rows, void **user_dev, void **item_dev, void **ctx_dev, void **targets_dev);
int rfm_factored_finalize(rfm_csr *rows);
/* Device-side generator of semi-synthetic interactions, mirroring the reference's
 * simulation utils/dataloader/kuairec/_click.py:148-235: relevance gamma = clip(watch_ratio / relevance_clip, 0, 1)
 * (:148-171), exposure theta_i = max(sigmoid(3 z_i - 1) ** exposure_bias, eps) per item (:173-205, computed by the
 * caller: item_exposure), R ~ Be(gamma), O ~ Be(theta), click Y = O * R (:207-235); pscore = theta ** pow_used
 * (kuairec/loader.py:167). The reference draws from NumPy's legacy global stream (sequential); this generator is
 * counter-based (Philox4x32-10 keyed by seed, counter = global row index), so any shard of the log can be produced
 * on any GPU: row g draws its user from user_cdf and its item from item_cdf (inclusive cumulative distributions,
 * binary search), its context values ~ N(0, 1), its watch ratio = exp(hidden_scale <p_u, q_i> + noise_scale eps
 * + watch_shift) with rank-n_hidden factors p_u, q_i ~ N(0, 1) that are themselves Philox functions of the id.
 * Specification (bit-exact for ids, 1e-12 for values): oracle/clicks_oracle.py. rfm_factored_generate returns
 * factored rows (blocks as rfm_factored_create; context blocks need no values) holding rows
 * [row0, row0 + n_rows) of the log with targets y / pscore; keep_labels also keeps Y and R (rfm_rows_download). */
typedef struct rfm_click_model {
  uint64_t seed;
  int64_t row0;
  int64_t n_users, n_items;
  const double *user_cdf;        /* [n_users], non-decreasing, last entry 1.0 */
  const double *item_cdf;        /* [n_items] */
  const double *item_exposure;   /* [n_items] theta_i in (0, 1] */
  const double *item_pscore;     /* [n_items] theta_i ** pow_used, as the caller's NumPy computes it */
  double pow_used;               /* informational */
  int32_t n_hidden, keep_labels;
  double hidden_scale, noise_scale, watch_shift, relevance_clip;
} rfm_click_model;
int rfm_factored_generate(rfm_ctx *ctx, int64_t n_rows, const rfm_click_model *model, const rfm_rows_block *blocks,
                          int32_t n_blocks, int dtype, rfm_csr **out);
/* Factored rows -> the stacked CSR the reference would have built, assembled on the device (no PCIe traffic): the
 * same entries in the same order, so nothing changes by a bit; the row kernels run ~10 % faster on resident CSR
 * rows where everything is cached, at 12 m + 16 bytes of HBM per interaction instead of ~24. */
int rfm_rows_materialize(const rfm_csr *rows, rfm_csr **out);
/* Rows [first, first + n) of factored rows back on the host (any output may be NULL): ids, context values
 * [n][n_ctx], targets y / pscore, and for generated rows with keep_labels the click and relevance labels. */
int rfm_rows_download(rfm_csr *rows, int64_t first, int64_t n, int32_t *users, int32_t *items, double *ctx_values,
                      double *targets, signed char *labels, signed char *relevance);
/* Replace the per-row targets y/pscore with values the caller computed in float64 (fractional labels: the
 * reference divides whatever `labels` holds, src/fm.py:80; rfm_csr_create takes integer labels). */
int rfm_csr_set_targets(rfm_csr *rows, const double *targets /* [n_rows] */);
int rfm_csr_destroy(rfm_csr *rows);

/* ---- FM model: w0, w, V (src/fm.py:31-53); parameter holders of utils/optimizer.py:10-64 */
int rfm_fm_create(rfm_ctx *ctx, int64_t n_features, int32_t n_factors, int dtype, rfm_fm **out);
int rfm_fm_destroy(rfm_fm *m);
/* V is row-major (n_features, n_factors) float64, exactly the ndarray SGD.params holds. */
int rfm_fm_set_params(rfm_fm *m, const double *w0, const double *w, const double *V);
int rfm_fm_get_params(rfm_fm *m, double *w0, double *w, double *V);
/* FactorizationMachines.predict, src/fm.py:114-133 (+ _sigmoid, src/base.py:63-66). */
int rfm_fm_predict(rfm_fm *m, const rfm_csr *rows, double *out_scores);
/* The same scores left in DEVICE memory (double [n_rows]), asynchronously on the context's stream: what
 * fit(evaluator=...) chains into rfm_ranker_evaluate_dev every epoch (src/fm.py:104-110) without a host round trip. */
int rfm_fm_predict_dev(rfm_fm *m, const rfm_csr *rows, double *out_scores_dev);
/* _cross_entropy_loss(labels, predict(rows), pscores), src/base.py:37-61. */
int rfm_fm_logloss(rfm_fm *m, const rfm_csr *rows, double *out_loss);

/* ---- FM training: one call == one reference "epoch" (one minibatch), src/fm.py:71-102 ----
 * forward + IPS residual (fm.py:80), _update_w0/_update_w/_update_V (fm.py:135-187) as one
 * simultaneous step with a deterministic segmented reduction per feature column, then the
 * post-update batch loss (fm.py:90-96) and the full val loss (fm.py:98-102). Losses are
 * kept on the device in slot `slot` and read back with rfm_fm_trainer_losses. */
int rfm_fm_trainer_create(rfm_fm *m, const rfm_csr *train, const rfm_csr *val /* may be NULL */,
                          int64_t max_batch, int64_t max_slots, rfm_fm_trainer **out);
int rfm_fm_trainer_destroy(rfm_fm_trainer *t);
/* Two-level step for FACTORED train rows (csrc/two_level.cuh). The same update as above (src/fm.py:80-88, 135-187),
 * computed per entity first: the user-keyed part of x_t depends on the user alone (the reference stacks
 * onehot[user], user_table[user], onehot[item], item_table[item]: coat/_preparer.py:154-170,
 * kuairec/_feature.py:169-209), so s_t = A_user + C_item + context terms and
 * grad v_j = sum_users x_uj (sum_{t of u} e_t s_t) - v_j sum_users x_uj^2 (sum_{t of u} e_t). A step then gathers
 * 2 + n_ctx parameter rows per interaction instead of m, plus one pass over the entity tables. Same sums, associated
 * per entity first: equal to the flat step to rounding, bit-reproducible run to run. mode 1 = on, 2 = on where the
 * cost model predicts >= 1.5 x fewer gathered rows per step; call before the first epoch; *enabled = outcome. */
int rfm_fm_trainer_set_two_level(rfm_fm_trainer *t, int32_t mode, int32_t *enabled);
int rfm_fm_train_epoch(rfm_fm_trainer *t, const int64_t *batch_rows, int64_t batch, double lr,
                       int64_t slot);
/* same step, batch drawn on the device by the Feistel sampler (perf mode). */
int rfm_fm_train_epoch_sampled(rfm_fm_trainer *t, uint32_t seed, uint32_t epoch, int64_t batch,
                               double lr, int64_t slot);
/* The same epoch with a dense optimizer step instead of the reference's fused SGD (SURVEY.md section 8 row f5;
 * not in the reference, utils/optimizer.py ends at :64 -- specification: oracle/optimizer_oracle.py). The
 * batch gradient g of the IPS logloss is formed as in src/fm.py:80-88,135-187, then
 *   RFM_OPT_SGD:  theta -= lr (g + l2 theta)
 *   RFM_OPT_ADAM: g += l2 theta; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 *                 theta -= lr (m / (1-b1^step)) / (sqrt(v / (1-b2^step)) + eps), step = 1, 2, ...
 * on w0, w and V; the moments live in the trainer and start at zero. batch_rows == NULL draws the batch on
 * the device (Feistel sampler, seed/epoch as in rfm_fm_train_epoch_sampled). Losses as rfm_fm_train_epoch. */
#define RFM_OPT_SGD 0
#define RFM_OPT_ADAM 1
typedef struct rfm_optimizer {
  int32_t kind;
  int32_t reserved;
  double lr, l2, beta1, beta2, eps;
  int64_t step;
} rfm_optimizer;
int rfm_fm_train_epoch_opt(rfm_fm_trainer *t, const int64_t *batch_rows /* may be NULL */, uint32_t seed,
                           uint32_t epoch, int64_t batch, int64_t slot, const rfm_optimizer *opt);
/* Data-parallel split of the same step: rank-local gradient of a batch
 * slice into a dense buffer [sum_e, 3 pad | dw (n, padded to a multiple of 4) | dV (n x kpad)]
 * (rfm_fm_grad_size scalars of the model dtype), to be all-reduced by the caller, then applied
 * identically on every rank. */
int rfm_fm_grad_size(rfm_fm_trainer *t, int64_t *n_scalars);
int rfm_fm_grad_ptr_dev(rfm_fm_trainer *t, void **grad_dev);
int rfm_fm_grad_epoch(rfm_fm_trainer *t, const int64_t *batch_rows, int64_t batch);
/* same, the slice [q_begin, q_begin + batch) of the epoch's Feistel permutation drawn on the device. */
int rfm_fm_grad_epoch_sampled(rfm_fm_trainer *t, uint32_t seed, uint32_t epoch, int64_t q_begin,
                              int64_t batch);
int rfm_fm_apply_grad(rfm_fm_trainer *t, double lr);
/* The same exchange over NVLink peer memory, one kernel per step and rank instead of {all-reduce, apply}
 * (SURVEY.md section 8e "fuse ... deterministic: fixed rank-order summation"). rfm_fm_dp_export moves the
 * gradient buffer into a CUDA-IPC region (call it before the first rfm_fm_grad_*; rfm_fm_grad_ptr_dev
 * then changes every step) and returns its handle; the caller gathers the handles of all ranks (at most
 * 8, one process per GPU of one node) and passes them, in rank order, to rfm_fm_dp_connect.
 * rfm_fm_dp_exchange_apply then: barrier -> every rank sums its slice over all ranks in rank order, in
 * place -> barrier -> every rank applies the reduced gradient read from the slice owners. The loss sums
 * of the previous step (rfm_fm_loss_sums) ride in the buffer header; their global values land in the
 * device doubles returned by rfm_fm_dp_prev_loss_ptr_dev (status: 0 = ok, else a barrier timed out). */
#define RFM_DP_HANDLE_BYTES 64
int rfm_fm_dp_export(rfm_fm_trainer *t, void *handle_out /* RFM_DP_HANDLE_BYTES */);
int rfm_fm_dp_connect(rfm_fm_trainer *t, int32_t rank, int32_t world,
                      const void *all_handles /* world x RFM_DP_HANDLE_BYTES */);
int rfm_fm_dp_exchange_apply(rfm_fm_trainer *t, double lr);
/* Diagnostic (RFM_DPX_TRACE=1): %globaltimer (ns) of the last exchange kernel's CTA 0 at [start, first barrier passed,
 * own slice reduced, second barrier passed, parameters applied]; stamps_ns: uint64 [8]. */
int rfm_fm_dp_trace(rfm_fm_trainer *t, uint64_t *stamps_ns);
int rfm_fm_dp_prev_loss_ptr_dev(rfm_fm_trainer *t, void **sums_dev /* double[2] */,
                                void **status_dev /* uint32, may be NULL */);
/* post-update loss of a batch slice / of val rows [row_begin, row_end): SUM of the per-row
 * terms (not divided) written to device slots so ranks can all-reduce them. batch_rows == NULL
 * reuses the batch that the last grad call left on the device. */
int rfm_fm_loss_sums_ptr_dev(rfm_fm_trainer *t, void **sums_dev /* double[2]: batch, val */);
int rfm_fm_loss_sums(rfm_fm_trainer *t, const int64_t *batch_rows, int64_t batch,
                     int64_t val_begin, int64_t val_end);
int rfm_fm_trainer_losses(rfm_fm_trainer *t, int64_t first_slot, int64_t n_slots,
                          double *train_loss, double *val_loss);

/* ---- MF rows and model (src/mf.py) ---------------------------------------------------- */
int rfm_pairs_create(rfm_ctx *ctx, int64_t n_rows, const int64_t *user_item /* (n_rows,2) */,
                     const int64_t *labels, const double *pscores, int dtype, rfm_pairs **out);
int rfm_pairs_set_targets(rfm_pairs *rows, const double *targets /* [n_rows]: y/pscore, src/mf.py:99 */);
int rfm_pairs_destroy(rfm_pairs *rows);
int rfm_mf_create(rfm_ctx *ctx, int64_t n_users, int64_t n_items, int32_t n_factors, int dtype,
                  rfm_mf **out);
int rfm_mf_destroy(rfm_mf *m);
int rfm_mf_set_params(rfm_mf *m, const double *P, const double *Q, const double *b_u,
                      const double *b_i, double b);
int rfm_mf_get_params(rfm_mf *m, double *P, double *Q, double *b_u, double *b_i);
/* LogisticMatrixFactorization.predict, src/mf.py:136-170. */
int rfm_mf_predict(rfm_mf *m, const rfm_pairs *rows, double *out_scores);
int rfm_mf_predict_dev(rfm_mf *m, const rfm_pairs *rows, double *out_scores_dev);
int rfm_mf_logloss(rfm_mf *m, const rfm_pairs *rows, double *out_loss);
/* one reference epoch, src/mf.py:97-124: strictly sequential per-sample SGD semantics
 * (P then Q-with-new-P then b_u, b_i; residual taken first) executed as a wavefront
 * schedule, then post-update batch loss and val loss. */
int rfm_mf_train_epoch(rfm_mf *m, const rfm_pairs *train, const rfm_pairs *val /* may be NULL */,
                       const int64_t *batch_rows, int64_t batch, double lr, double reg,
                       double *train_loss, double *val_loss);

/* ---- ranking evaluation (utils/evaluate.py, utils/metrics.py) ---------------------------
 * rfm_ranker_create groups the rows of an interaction frame by user like
 * interaction_df.groupby("user") (evaluate.py:129-156, 209-239).
 * rfm_ranker_evaluate ranks every user's candidates by score (canonical order: score
 * descending, later row first among exact ties -- argsort(kind="stable")[::-1]) and returns,
 * per K[j]:  sums and valid-user counts for DCG (metrics.py:83-107), IPS-DCG (:53-80),
 * ME (:110-127), Recall (:32-50), MAP (:9-29), and the per-item hit counts of the top-K[j]
 * lists (CatalogCoverage :152-166 == non-zero count / n_items; Gini :130-149). Users whose
 * labels sum to zero are skipped (evaluate.py:98-99, 201-202). */
int rfm_ranker_create(rfm_ctx *ctx, int64_t n_rows, const int64_t *users, const int64_t *items,
                      const double *labels, const double *pscores, int64_t n_items,
                      rfm_ranker **out);
int rfm_ranker_destroy(rfm_ranker *r);
int rfm_ranker_num_users(rfm_ranker *r, int64_t *out);
/* out_metrics: double[n_k][RFM_RANK_NCOLS]; out_item_hits: int32[n_k][n_items] or NULL;
 * out_top_rows: int64[n_users][max K] row ids (-1 padded) or NULL. */
#define RFM_RANK_NCOLS 12
enum rfm_rank_col {
  RFM_RANK_DCG_SUM = 0, RFM_RANK_IPSDCG_SUM = 1, RFM_RANK_ME_SUM = 2, RFM_RANK_ME_COUNT = 3,
  RFM_RANK_RECALL_SUM = 4, RFM_RANK_MAP_SUM = 5, RFM_RANK_USERS = 6, RFM_RANK_COVERED = 7
};
/* Optional: label totals per user (users in ascending id order, rfm_ranker_num_users entries) to use for
 * the skip rule and Recall's denominator instead of the sum over the rows given -- for callers that
 * hand over only the top of every user's candidate list (full-catalog evaluation). NULL clears it. */
int rfm_ranker_set_user_totals(rfm_ranker *r, const double *totals);
int rfm_ranker_evaluate(rfm_ranker *r, const double *scores, const int32_t *K, int32_t n_k,
                        double *out_metrics, int32_t *out_item_hits, int64_t *out_top_rows);

/* Device-chained evaluation (the epoch search, utils/search_params.py:79-152, calls
 * ValEvaluator.evaluate after every epoch, src/fm.py:104-110). rfm_ranker_scores_ptr_dev exposes the ranker's own
 * score buffer (double [n_rows], original row order) so that rfm_fm_predict_dev / rfm_mf_predict_dev can write
 * into it; rfm_ranker_evaluate_dev ranks them (scores_dev == NULL or that buffer: in place) and stores the metric
 * rows of up to 4 ranking positions in history slot `slot` ON THE DEVICE -- no synchronisation, no copy;
 * rfm_ranker_read_slots returns double[n_slots][n_k][RFM_RANK_NCOLS] after one synchronisation per fit. The
 * covered-items column is not filled on this path. */
int rfm_ranker_scores_ptr_dev(rfm_ranker *r, void **scores_dev);
int rfm_ranker_evaluate_dev(rfm_ranker *r, const double *scores_dev, const int32_t *K, int32_t n_k, int64_t slot,
                            int64_t max_slots);
int rfm_ranker_read_slots(rfm_ranker *r, int64_t first_slot, int64_t n_slots, int32_t n_k, double *out_metrics);

/* Full-catalog evaluation on the device (utils/evaluate.py:80-127 on the Cartesian-product frame: every item is a
 * candidate, the label is the held-out label where one exists and 0 elsewhere, the pscore the item's exposure).
 * The held-out labels come as a CSR by user (items strictly ascending inside a user). rfm_catalog_eval_run takes
 * ranked lists as rfm_topk_run / rfm_topk_run_sharded leave them on the device (int32 [n_rows][k_list], best
 * first, -1 padded; row r is user user_begin + r), looks the labels up, and returns the same metric rows and
 * per-item hit counts as rfm_ranker_evaluate (sums over the given users: ranks all-reduce them). */
typedef struct rfm_catalog_eval rfm_catalog_eval;
int rfm_catalog_eval_create(rfm_ctx *ctx, int64_t n_users, int64_t n_items, const int64_t *label_indptr,
                            const int32_t *label_items, const double *label_values, const double *item_pscores,
                            rfm_catalog_eval **out);
int rfm_catalog_eval_destroy(rfm_catalog_eval *e);
int rfm_catalog_eval_run(rfm_catalog_eval *e, const int32_t *lists_dev, int32_t k_list, int64_t user_begin,
                         int64_t n_rows, const int32_t *K, int32_t n_k, double *out_metrics, int32_t *out_item_hits);

/* ---- full-catalog scoring + exact top-K (new capability) ------------
 * score(u, i) = bias + alpha[u] + beta[i] + <A_u, C_i>: MF with A = P, C = Q, alpha = b_u, beta = b_i,
 * bias = b (src/mf.py:165-170); FM with the per-user / per-item sums of x_j v_j and their scalar parts
 * (src/fm.py:125-132) when rows are [user features | item features]. The reference never scores the
 * full grid (it ranks only the rows it is given, utils/evaluate.py:80-127); parity is "equal to
 * predict on the Cartesian-product rows + per-user argsort".
 * rfm_topk_run returns, for every user, the K best items of the catalog range [item_begin, item_end)
 * in the canonical order (score descending, larger item id first among exact ties) with their exact
 * float64 scores. mode 0: two bf16 tcgen05 GEMM passes with fused epilogues -- the first finds, per
 * user, a threshold that provably lies below the K-th best score (minus the bf16 error bound), the
 * second collects every item that reaches it -- then exact float64 re-scoring of the collected items
 * (users whose candidate buffer overflowed are ranked exactly); mode 1: exact float64 only.
 * stats (int64[4], may be NULL): [0] = 1 if the tensor-core path ran, [1] = users that needed the
 * exact fallback, [2] = candidates collected over all users, [3] = tile stride of the first pass.
 * A, C are row-major float64 (n_users, k), (n_items, k); alpha / beta may be NULL. */
typedef struct rfm_topk rfm_topk;
int rfm_topk_create(rfm_ctx *ctx, int64_t n_users, int64_t n_items, int32_t n_factors, rfm_topk **out);
int rfm_topk_destroy(rfm_topk *t);
int rfm_topk_set_factors(rfm_topk *t, const double *A, const double *C, const double *alpha,
                         const double *beta, double bias);
int rfm_topk_run(rfm_topk *t, int32_t K, int32_t mode, int64_t item_begin, int64_t item_end,
                 int32_t *out_items, double *out_scores, int64_t *stats);
/* The last run's result in the library's own page-locked host buffers (int32 [n_users][K], double
 * [n_users][K]), valid until the next rfm_topk_run / rfm_topk_result_host / rfm_topk_destroy on this handle:
 * saves the copy into caller memory when the caller only reads the result (call rfm_topk_run with NULL outputs). */
int rfm_topk_result_host(rfm_topk *t, int32_t K, const int32_t **items_host, const double **scores_host);
/* Item-sharded runs: rfm_topk_run with out_items == out_scores == NULL leaves the
 * shard's result on the device; rfm_topk_result_ptr_dev exposes it (int32 [n_users][K], double
 * [n_users][K]) for the caller's all-gather; rfm_topk_merge_dev merges n_lists gathered lists
 * ([n_lists][n_users][K], device pointers) into the global top-K (host outputs), canonical order. */
int rfm_topk_result_ptr_dev(rfm_topk *t, void **items_dev, void **scores_dev);
/* The same exchange over NVLink peer memory instead of an all-gather ("all-to-all by user
 * range ... then K-way merge"). One process per GPU of one node, at most 8 ranks. rfm_topk_dp_export allocates the
 * scorer's exchange region (per-rank lists and per-rank group maxima for up to k_cap entries per user) and returns
 * its CUDA-IPC handle; the caller gathers the handles of all ranks and passes them, in rank order, to
 * rfm_topk_dp_connect. rfm_topk_run_sharded is then a COLLECTIVE call (same K and mode on every rank): rank r
 * ranks catalog tiles slice(r) for every user -- the collect threshold is GLOBAL: the K-th largest of the union of
 * all ranks' sampled group maxima, read from the peers' regions, so re-scoring work shrinks with the number of
 * ranks too -- and merges the sorted per-rank lists of the users it owns, users [r U/G, (r+1) U/G) with the
 * remainder spread over the first ranks, reading the peers' lists directly. out_items / out_scores (host,
 * [owned users][K], may be NULL: the result stays on the device, rfm_topk_result_ptr_dev) receive the owned users'
 * global top-K; user_range[2] (may be NULL) their [begin, end). Cross-GPU barriers are bounded: a rank that does
 * not arrive within 20 s fails the call with RFM_ERR_CUDA instead of hanging. stats as rfm_topk_run. */
int rfm_topk_dp_export(rfm_topk *t, int32_t k_cap, void *handle_out /* RFM_DP_HANDLE_BYTES */);
int rfm_topk_dp_connect(rfm_topk *t, int32_t rank, int32_t world, const void *all_handles);
int rfm_topk_run_sharded(rfm_topk *t, int32_t K, int32_t mode, int32_t *out_items, double *out_scores,
                         int64_t *user_range, int64_t *stats);
int rfm_topk_merge_dev(rfm_ctx *ctx, int64_t n_users, int32_t K, int32_t n_lists, const int32_t *items_dev,
                       const double *scores_dev, int32_t *out_items, double *out_scores);

#ifdef __cplusplus
}
#endif
#endif /* RFM_B200_H_ */
