// method_fitc (R/gp_fit.R:101-123,163-188; R/laplace.R:122-152,299-325; R/ep.R:148-185,333-374;
// R/linear_algebra.R:3-126; R/gp_pred.R:234-296).
#include "fit.cuh"

using namespace gpl;

extern "C" {

int gpl_fit_fitc(gpl_ctx* h, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                 const double* x, int64_t n, int32_t d, const double* zind, int64_t m,
                 const double* y, const double* trials, const double* offset, double jitter,
                 gpl_fit_out* out, double* fmean, double* z, double* V, double* alpha, double* fvar,
                 double* cavity_mean, double* cavity_var, gpl_fit** fit) {
  set_error("gpl_fit_fitc: not implemented yet");
  return GPL_ERR_BAD_ARG;
}

int gpl_pred_fitc(gpl_ctx* h, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt,
                  int64_t nt, uint64_t term_mask, double* mean, double* var) {
  set_error("gpl_pred_fitc: not implemented yet");
  return GPL_ERR_BAD_ARG;
}

}  // extern "C"
