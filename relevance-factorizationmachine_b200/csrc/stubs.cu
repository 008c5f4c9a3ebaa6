// stubs.cu -- TEMPORARY: entry points declared in include/rfm_b200.h that are not implemented
// yet return an error (never a silent fallback). Each is removed when its subsystem lands.
#include "common.cuh"
using namespace rfm;
#define NOT_YET(name) return fail(RFM_ERR_INVALID, name ": not implemented yet")
extern "C" {
int rfm_pairs_create(rfm_ctx *, int64_t, const int64_t *, const int64_t *, const double *, int, rfm_pairs **) { NOT_YET("rfm_pairs_create"); }
int rfm_pairs_destroy(rfm_pairs *) { return RFM_OK; }
int rfm_mf_create(rfm_ctx *, int64_t, int64_t, int32_t, int, rfm_mf **) { NOT_YET("rfm_mf_create"); }
int rfm_mf_destroy(rfm_mf *) { return RFM_OK; }
int rfm_mf_set_params(rfm_mf *, const double *, const double *, const double *, const double *, double) { NOT_YET("rfm_mf_set_params"); }
int rfm_mf_get_params(rfm_mf *, double *, double *, double *, double *) { NOT_YET("rfm_mf_get_params"); }
int rfm_mf_predict(rfm_mf *, const rfm_pairs *, double *) { NOT_YET("rfm_mf_predict"); }
int rfm_mf_logloss(rfm_mf *, const rfm_pairs *, double *) { NOT_YET("rfm_mf_logloss"); }
int rfm_mf_train_epoch(rfm_mf *, const rfm_pairs *, const rfm_pairs *, const int64_t *, int64_t, double, double, double *, double *) { NOT_YET("rfm_mf_train_epoch"); }
int rfm_ranker_create(rfm_ctx *, int64_t, const int64_t *, const int64_t *, const double *, const double *, int64_t, rfm_ranker **) { NOT_YET("rfm_ranker_create"); }
int rfm_ranker_destroy(rfm_ranker *) { return RFM_OK; }
int rfm_ranker_num_users(rfm_ranker *, int64_t *) { NOT_YET("rfm_ranker_num_users"); }
int rfm_ranker_evaluate(rfm_ranker *, const double *, const int32_t *, int32_t, double *, int32_t *, int64_t *) { NOT_YET("rfm_ranker_evaluate"); }
}
