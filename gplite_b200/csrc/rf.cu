// method_rf: random-feature (linearised) GP, SURVEY 8f-3.
//   feature maps        R/cf.R:532-690   (rf_featmap.*; check_num_basis R/gp.R:205-228)
//   fit                 R/gp_fit.R:125-132, R/laplace.R:241-262,327-350 (Laplace in weight space)
//   prediction          R/gp_pred.R:298-315 (get_w_mean / get_w_cov with cfind, R/gp.R:230-262)
//
// Phi (n x M) is built by one HBM-write-bound kernel from the prepared inputs (the same feature
// preparation as the K build: vars subset, normalisation, the periodic sin/cos embedding); the
// random frequencies / hidden-layer weights are drawn by the caller (R's rnorm stream cannot be
// reproduced outside R) and passed in.  H = Phi^T V^-1 Phi + I is one DMMA GEMM with k = n, the
// M x M factorisation and inverse reuse the dense path.
#include "fit.cuh"

#include <math.h>
#include <string.h>

#include <vector>

namespace gpl {

namespace {

constexpr int RF_MAX_TERMS = 16;

struct RfFactor {
  int kind, term, off, nf;
  int mw;            // columns of W (sexp: num_basis / 2, nn: num_basis)
  int64_t woff;      // offset of W (nf x mw, column-major) in the weight buffer; nn: w0 follows W
  double scale;      // magn * sqrt(1 / mw) (sexp), magn / sqrt(mw) (nn), magn (lin, const)
};

struct RfPlan {
  int nfactors, nterms, M;
  int col0[RF_MAX_TERMS], ncols[RF_MAX_TERMS];
  RfFactor f[MAX_FACTORS];
};

// Phi[i, j] = prod over the factors of column j's term of that factor's feature j
__global__ void rf_features_kernel(const __grid_constant__ RfPlan p, int64_t n,
                                   const double* __restrict__ z, int64_t ldz,
                                   const double* __restrict__ w, unsigned long long term_mask,
                                   double* __restrict__ phi, int64_t ldp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= n) return;
  int t = 0;
  while (t + 1 < p.nterms && j >= p.col0[t + 1]) ++t;
  const int jl = j - p.col0[t];
  double v = 1.0;
  if (term_mask && !((term_mask >> t) & 1ull)) {
    v = 0.0;                                     // cfind: unselected summands drop out
  } else {
    for (int q = 0; q < p.nfactors; ++q) {
      const RfFactor& f = p.f[q];
      if (f.term != t) continue;
      switch (f.kind) {
        case GPL_CF_CONST:
          v *= f.scale;
          break;
        case GPL_CF_LIN:
          v *= f.scale * z[(int64_t)(f.off + jl) * ldz + i];
          break;
        case GPL_CF_SEXP: {
          const int jj = jl < f.mw ? jl : jl - f.mw;
          const double* wc = w + f.woff + (int64_t)jj * f.nf;
          double h = 0.0;
          for (int k = 0; k < f.nf; ++k) h = fma(z[(int64_t)(f.off + k) * ldz + i], wc[k], h);
          v *= f.scale * (jl < f.mw ? cos(h) : sin(h));
          break;
        }
        case GPL_CF_NN: {
          const double* wc = w + f.woff + (int64_t)jl * f.nf;
          double h = w[f.woff + (int64_t)f.mw * f.nf + jl];
          for (int k = 0; k < f.nf; ++k) h = fma(z[(int64_t)(f.off + k) * ldz + i], wc[k], h);
          v *= f.scale * erf(h);
          break;
        }
      }
    }
  }
  phi[(int64_t)j * ldp + i] = v;
}

__global__ void mirror_lower_kernel(int64_t n, double* __restrict__ A, int64_t lda) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i < n && i > j) A[i * lda + j] = A[j * lda + i];
}

int make_rf_plan(const gpl_cf_model* cf, const HostPlan& hp, const gpl_rf_spec* rf, RfPlan* out) {
  RfPlan& p = *out;
  memset(&p, 0, sizeof(p));
  if (!rf || !rf->num_basis || rf->nterms != cf->nterms || rf->nterms > RF_MAX_TERMS) {
    set_error("method_rf: num_basis must have one entry per covariance function (at most %d)",
              RF_MAX_TERMS);
    return GPL_ERR_BAD_ARG;
  }
  p.nfactors = hp.plan.nfactors;
  p.nterms = cf->nterms;
  int col = 0;
  for (int t = 0; t < p.nterms; ++t) {
    p.col0[t] = col;
    p.ncols[t] = rf->num_basis[t];
    if (p.ncols[t] <= 0) {
      set_error("method_rf: num_basis[%d] = %d", t, p.ncols[t]);
      return GPL_ERR_BAD_ARG;
    }
    col += p.ncols[t];
  }
  p.M = col;
  if (p.M > 65535) {        // basis functions map to gridDim.y in the feature kernel
    set_error("method_rf: at most 65535 basis functions in total (got %d)", p.M);
    return GPL_ERR_BAD_ARG;
  }
  int64_t woff = 0;
  for (int q = 0; q < p.nfactors; ++q) {
    const FactorDev& fd = hp.plan.f[q];
    const gpl_cf_factor& src = cf->factors[q];
    RfFactor& f = p.f[q];
    f.kind = fd.kind;
    f.term = fd.term;
    f.off = fd.off;
    f.nf = fd.nf;
    const int nbt = p.ncols[fd.term];
    int members = 0;
    for (int o = 0; o < p.nfactors; ++o) members += hp.plan.f[o].term == fd.term;
    switch (fd.kind) {
      case GPL_CF_CONST:
      case GPL_CF_LIN:
        if (members > 1) {                       // R/cf.R:677-680
          set_error("Random features for product kernel containing constant or linear kernel not "
                    "implemented yet.");
          return GPL_ERR_BAD_ARG;
        }
        if (nbt != (fd.kind == GPL_CF_CONST ? 1 : fd.nf)) {
          set_error("method_rf: cf_const needs 1 basis function and cf_lin one per input "
                    "(check_num_basis, R/gp.R:205-228); got %d", nbt);
          return GPL_ERR_BAD_ARG;
        }
        f.scale = src.p[0];
        break;
      case GPL_CF_SEXP:
        if (nbt % 2) {                           // R/cf.R:589-591
          set_error("number of features must be an even number.");
          return GPL_ERR_BAD_ARG;
        }
        f.mw = nbt / 2;
        f.scale = src.p[1] * sqrt(1.0 / f.mw);
        f.woff = woff;
        woff += (int64_t)f.nf * f.mw;
        break;
      case GPL_CF_NN:
        f.mw = nbt;
        f.scale = src.p[2] / sqrt((double)f.mw);
        f.woff = woff;
        woff += (int64_t)f.nf * f.mw + f.mw;
        break;
      default:                                   // R/cf.R:612-620
        set_error("Random Fourier features for Matern kernels not implemented yet.");
        return GPL_ERR_BAD_ARG;
    }
  }
  if (woff != rf->nweights || (woff > 0 && !rf->weights)) {
    set_error("method_rf: expected %lld random weights, got %lld", (long long)woff,
              (long long)rf->nweights);
    return GPL_ERR_BAD_ARG;
  }
  return GPL_OK;
}

int rf_features(Ctx* ctx, const RfPlan& rp, const HostPlan& hp, const double* d_x, int64_t n,
                const double* d_w, uint64_t term_mask, double* phi, int64_t ldp) {
  const int F = hp.plan.nfeat > 0 ? hp.plan.nfeat : 1;
  DevBuf Z;
  GPL_TRY(Z.alloc(ctx, sizeof(double) * n * F));
  GPL_TRY(prep_features(ctx, hp, d_x, n, n, Z.d(), n));
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)rp.M);
  rf_features_kernel<<<grid, 256, 0, ctx->stream>>>(rp, n, Z.d(), n, d_w, term_mask, phi, ldp);
  GPL_LAUNCH_CHECK(ctx);
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));   // Z is released at scope exit
  return GPL_OK;
}

int up(Ctx* ctx, DevBuf& b, const double* h, size_t count) {
  GPL_TRY(b.alloc(ctx, sizeof(double) * (count ? count : 1)));
  if (count)
    GPL_CUDA(cudaMemcpyAsync(b.p, h, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
  return GPL_OK;
}

int down(Ctx* ctx, double* h, const double* d, size_t count) {
  if (h && count)
    GPL_CUDA(cudaMemcpyAsync(h, d, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
  return GPL_OK;
}

}  // namespace

int fit_rf_impl(Ctx* ctx, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                const gpl_rf_spec* rf, const double* x, int64_t n, int d, const double* y,
                const double* trials, const double* offset, gpl_fit_out* out, double* h_fmean,
                double* h_z, double* h_V, double* h_wmean, Fit** fit) {
  if (n <= 0 || d <= 0 || !x || !y || !out || !approx) {
    set_error("gpl_fit_rf: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  if (approx->kind != GPL_APPROX_LAPLACE) {
    set_error("No EP implementation for method_rf yet.");          // R/gp_fit.R:146-148
    return GPL_ERR_BAD_ARG;
  }
  if ((lik->kind == GPL_LIK_BINOMIAL || lik->kind == GPL_LIK_BETABINOM) && !trials) {
    set_error("trials must be provided for the binomial / beta-binomial likelihood");
    return GPL_ERR_BAD_ARG;
  }
  memset(out, 0, sizeof(*out));
  out->quad_ok = 1;
  HostPlan hp;
  GPL_TRY(make_plan(cf, d, 0, &hp));
  RfPlan rp;
  GPL_TRY(make_rf_plan(cf, hp, rf, &rp));
  const int64_t M = rp.M, ldp = round_up(n, 16), ldh = round_up(M, 16);
  DevBuf dx, dy, dtr, doff, dw, Phi, PhiV, H, Hd;
  DevBuf f, fnew, z, V, iV, zv, rhs, t1, what;
  GPL_TRY(up(ctx, dx, x, (size_t)n * d));
  GPL_TRY(up(ctx, dy, y, n));
  GPL_TRY(up(ctx, dw, rf->weights, rf->nweights));
  LikDev ld{lik->kind, lik->link, lik->param, nullptr, nullptr};
  if (trials) { GPL_TRY(up(ctx, dtr, trials, n)); ld.trials = dtr.d(); }
  if (offset) { GPL_TRY(up(ctx, doff, offset, n)); ld.offset = doff.d(); }
  GPL_TRY(Phi.alloc(ctx, sizeof(double) * ldp * M));
  GPL_TRY(PhiV.alloc(ctx, sizeof(double) * ldp * M));
  GPL_TRY(H.alloc(ctx, sizeof(double) * ldh * M));
  GPL_TRY(Hd.alloc(ctx, sizeof(double) * ((M + NB - 1) / NB) * NB * NB));
  for (DevBuf* b : {&f, &fnew, &z, &V, &iV, &zv}) GPL_TRY(b->alloc(ctx, sizeof(double) * n));
  for (DevBuf* b : {&rhs, &t1, &what}) GPL_TRY(b->alloc(ctx, sizeof(double) * M));
  GPL_TRY(rf_features(ctx, rp, hp, dx.d(), n, dw.d(), 0, Phi.d(), ldp));
  GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, f.d(), 0.0));
  int maxiter = approx->maxiter < 1 ? 1 : approx->maxiter;
  if (lik->kind == GPL_LIK_GAUSSIAN) maxiter = 1;                     // R/laplace.R:334-336
  int it = 0;
  double log_evidence = 0.0, sc[4];
  for (it = 1; it <= maxiter; ++it) {
    // laplace_iter.method_rf, R/laplace.R:241-262
    GPL_TRY(lik_pseudodata_la(ctx, ld, n, f.d(), dy.d(), z.d(), V.d()));
    GPL_TRY(vec_op(ctx, VOP_INV, n, V.d(), nullptr, iV.d()));
    GPL_TRY(vec_op(ctx, VOP_MUL, n, z.d(), iV.d(), zv.d()));
    GPL_CUDA(cudaMemcpy2DAsync(PhiV.d(), sizeof(double) * ldp, Phi.d(), sizeof(double) * ldp,
                               sizeof(double) * n, M, cudaMemcpyDeviceToDevice, ctx->stream));
    GPL_TRY(scale_rows(ctx, n, M, PhiV.d(), ldp, iV.d()));
    GPL_TRY(dgemm(ctx, true, false, M, M, n, 1.0, Phi.d(), ldp, PhiV.d(), ldp, 0.0, H.d(), ldh, true));
    GPL_TRY(scale_block(ctx, M, M, H.d(), ldh, nullptr, nullptr, 1.0, 1.0, 0));     // + wprec_prior
    int32_t info = 0;
    double logdetH = 0.0;
    GPL_TRY(dpotrf_with(ctx, Hd.d(), M, H.d(), ldh, &info, &logdetH));
    GPL_TRY(gemv_t(ctx, n, M, Phi.d(), ldp, zv.d(), rhs.d()));
    GPL_TRY(trsv_lower(ctx, Hd.d(), false, M, H.d(), ldh, rhs.d(), t1.d()));
    GPL_TRY(trsv_lower(ctx, Hd.d(), true, M, H.d(), ldh, t1.d(), what.d()));
    GPL_TRY(gemv_n(ctx, n, M, Phi.d(), ldp, what.d(), fnew.d()));
    GPL_TRY(reduce(ctx, RED_SUMSQ, M, what.d(), nullptr, 0));
    GPL_TRY(lik_loglik_sum(ctx, ld, n, fnew.d(), dy.d(), 1));
    GPL_TRY(reduce(ctx, RED_MAXABSDIFF, n, f.d(), fnew.d(), 2));
    GPL_TRY(fetch_scalars(ctx, 3, sc));
    const double log_prior = -0.5 * (double)M * log(2.0 * M_PI) - 0.5 * sc[0];
    log_evidence = 0.5 * (double)n * log(2.0 * M_PI) - 0.5 * logdetH + log_prior + sc[1];
    std::swap(f.p, fnew.p);
    if (sc[2] < approx->tol) break;
  }
  if (it > maxiter) it = maxiter;
  out->log_evidence = log_evidence;
  out->iters = it;
  out->accepted = it;
  out->converged = !(maxiter > 1 && it == maxiter);
  GPL_TRY(down(ctx, h_fmean, f.d(), n));
  GPL_TRY(down(ctx, h_z, z.d(), n));
  GPL_TRY(down(ctx, h_V, V.d(), n));
  GPL_TRY(down(ctx, h_wmean, what.d(), M));
  // wcov = H^-1, full symmetric
  GPL_TRY(dpotri_with(ctx, Hd.d(), M, H.d(), ldh));
  dim3 grid((unsigned)((M + 255) / 256), (unsigned)M);
  mirror_lower_kernel<<<grid, 256, 0, ctx->stream>>>(M, H.d(), ldh);
  GPL_LAUNCH_CHECK(ctx);
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (fit) {
    Fit* F = new Fit();
    F->method = 2;
    F->n = n;
    F->m = M;
    F->d = d;
    F->ldm = ldh;
    auto take = [&](DevBuf& dst, DevBuf& src) { dst.ctx = ctx; dst.p = src.p; src.p = nullptr; };
    take(F->Lm, H);            // wcov
    take(F->alpha, what);      // wmean
    take(F->zind, dw);         // random weights
    F->rf_num_basis.assign(rf->num_basis, rf->num_basis + rf->nterms);
    F->rf_nweights = rf->nweights;
    *fit = F;
  }
  return GPL_OK;
}

int pred_rf_impl(Ctx* ctx, const Fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                 uint64_t term_mask, double* mean, double* var) {
  if (!fit || fit->method != 2 || nt < 0 || !mean || !xt) {
    set_error("gpl_pred_rf: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  if (nt == 0) return GPL_OK;
  const int d = fit->d;
  const int64_t M = fit->m;
  HostPlan hp;
  GPL_TRY(make_plan(cf, d, 0, &hp));
  gpl_rf_spec rf;
  rf.nterms = (int32_t)fit->rf_num_basis.size();
  rf.num_basis = fit->rf_num_basis.data();
  rf.weights = fit->zind.d();               // device pointer: only its presence is checked
  rf.nweights = fit->rf_nweights;
  RfPlan rp;
  GPL_TRY(make_rf_plan(cf, hp, &rf, &rp));
  if (rp.M != M) {
    set_error("gpl_pred_rf: the model has %d basis functions, the fit %lld", rp.M, (long long)M);
    return GPL_ERR_BAD_ARG;
  }
  DevBuf dxt, Phi, T, dm, dv;
  const int64_t ldp = round_up(nt, 16);
  GPL_TRY(up(ctx, dxt, xt, (size_t)nt * d));
  GPL_TRY(Phi.alloc(ctx, sizeof(double) * ldp * M));
  GPL_TRY(dm.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(rf_features(ctx, rp, hp, dxt.d(), nt, fit->zind.d(), term_mask, Phi.d(), ldp));
  GPL_TRY(gemv_n(ctx, nt, M, Phi.d(), ldp, fit->alpha.d(), dm.d()));
  GPL_TRY(down(ctx, mean, dm.d(), nt));
  if (var) {
    // diag(zt wcov zt^T): rows of (Phi wcov) . Phi   (R/gp_pred.R:309-312)
    GPL_TRY(T.alloc(ctx, sizeof(double) * ldp * M));
    GPL_TRY(dv.alloc(ctx, sizeof(double) * nt));
    GPL_TRY(dgemm(ctx, false, false, nt, M, M, 1.0, Phi.d(), ldp, fit->Lm.d(), fit->ldm, 0.0, T.d(),
                  ldp, false));
    GPL_TRY(rowdot(ctx, nt, M, T.d(), ldp, Phi.d(), ldp, dv.d()));
    GPL_TRY(down(ctx, var, dv.d(), nt));
  }
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPL_OK;
}

// Phi_t (nt x M, column-major, host) for a fitted method_rf model: the featuremap(xt, cfind) of
// gp_draw_analytic.method_rf (R/gp_draw.R:145-163); unselected summands give zero columns.
int rf_features_impl(Ctx* ctx, const Fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                     uint64_t term_mask, double* phi_out) {
  if (!fit || fit->method != 2 || nt <= 0 || !xt || !phi_out) {
    set_error("gpl_rf_features: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  HostPlan hp;
  GPL_TRY(make_plan(cf, fit->d, 0, &hp));
  gpl_rf_spec rf;
  rf.nterms = (int32_t)fit->rf_num_basis.size();
  rf.num_basis = fit->rf_num_basis.data();
  rf.weights = fit->zind.d();
  rf.nweights = fit->rf_nweights;
  RfPlan rp;
  GPL_TRY(make_rf_plan(cf, hp, &rf, &rp));
  DevBuf dxt, Phi;
  const int64_t ldp = round_up(nt, 16);
  GPL_TRY(up(ctx, dxt, xt, (size_t)nt * fit->d));
  GPL_TRY(Phi.alloc(ctx, sizeof(double) * ldp * rp.M));
  GPL_TRY(rf_features(ctx, rp, hp, dxt.d(), nt, fit->zind.d(), term_mask, Phi.d(), ldp));
  GPL_CUDA(cudaMemcpy2DAsync(phi_out, sizeof(double) * nt, Phi.d(), sizeof(double) * ldp,
                             sizeof(double) * nt, rp.M, cudaMemcpyDeviceToHost, ctx->stream));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPL_OK;
}

}  // namespace gpl

extern "C" {

using namespace gpl;

int gpl_fit_rf(gpl_ctx* h, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
               const gpl_rf_spec* rf, const double* x, int64_t n, int32_t d, const double* y,
               const double* trials, const double* offset, gpl_fit_out* out, double* fmean, double* z,
               double* V, double* wmean, gpl_fit** fit) {
  Fit* f = nullptr;
  if (fit) *fit = nullptr;
  int rc = fit_rf_impl(reinterpret_cast<Ctx*>(h), cf, lik, approx, rf, x, n, d, y, trials, offset, out,
                       fmean, z, V, wmean, fit ? &f : nullptr);
  if (fit) *fit = reinterpret_cast<gpl_fit*>(f);
  if (rc != GPL_OK) cudaStreamSynchronize(reinterpret_cast<Ctx*>(h)->stream);
  return rc;
}

int gpl_rf_features(gpl_ctx* h, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                    uint64_t term_mask, double* phi) {
  return rf_features_impl(reinterpret_cast<Ctx*>(h), reinterpret_cast<const Fit*>(fit), cf, xt, nt,
                          term_mask, phi);
}

int gpl_pred_rf(gpl_ctx* h, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                uint64_t term_mask, double* mean, double* var) {
  return pred_rf_impl(reinterpret_cast<Ctx*>(h), reinterpret_cast<const Fit*>(fit), cf, xt, nt,
                      term_mask, mean, var);
}

}  // extern "C"
