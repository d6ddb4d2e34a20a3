// temporary stubs until k2/k3 land
#include "mtt_common.cuh"
#define NI(ctx) mtt_fail(ctx, MTT_ERR_STATE, "not implemented yet")
extern "C" {
int mtt_unif_set_models(mtt_ctx *c, int, int, const double *, int) { return NI(c); }
int mtt_unif_get_matrices(mtt_ctx *c, int, double *, double *, double *) { return NI(c); }
int mtt_unif_rpow(mtt_ctx *c, int, int, double *, int *) { return NI(c); }
int mtt_unif_expm_batch(mtt_ctx *c, int64_t, const int32_t *, const double *, double *) { return NI(c); }
int mtt_unif_branch_logp_batch(mtt_ctx *c, int64_t, const int32_t *, const double *, const double *, const uint8_t *, const uint8_t *, const int64_t *, const double *, const uint8_t *, double *, int32_t *) { return NI(c); }
int mtt_unif_retype_batch(mtt_ctx *c, int64_t, const int64_t *, const int32_t *, const double *, const double *, const uint8_t *, const uint8_t *, const int32_t *, const double *, const int64_t *, int, int32_t *, double *, uint8_t *, double *, int32_t *, int32_t *, int64_t *) { return NI(c); }
int mtt_peel_create(mtt_ctx *c, int, int64_t, int, mtt_peel **) { return NI(c); }
int mtt_peel_destroy(mtt_peel *) { return MTT_ERR_STATE; }
int mtt_peel_set_tip_states(mtt_peel *, const int32_t *) { return MTT_ERR_STATE; }
int mtt_peel_set_pattern_weights(mtt_peel *, const int32_t *) { return MTT_ERR_STATE; }
int mtt_peel_set_site_model(mtt_peel *, const double *, const double *, const double *, double, double) { return MTT_ERR_STATE; }
int mtt_peel_set_hky(mtt_peel *, double) { return MTT_ERR_STATE; }
int mtt_peel_logl_batch(mtt_peel *, int64_t, const int32_t *, const double *, double *) { return MTT_ERR_STATE; }
int mtt_peel_logl_batch_mats(mtt_peel *, int64_t, const int32_t *, const double *, double *) { return MTT_ERR_STATE; }
int mtt_peel_logl_batch_dev(mtt_peel *, int64_t, const int32_t *, const double *, double *) { return MTT_ERR_STATE; }
int mtt_peel_pattern_logl(mtt_peel *, const int32_t *, const double *, double *) { return MTT_ERR_STATE; }
}
