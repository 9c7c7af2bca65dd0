// Device-resident fit loops for method_full: Laplace/Newton (R/laplace.R:13-38,274-297), parallel
// EP (R/ep.R:16-49,286-330), the Gaussian energy + analytic gradient (gpl_energy_grad) and
// prediction (R/gp_pred.R:202-232).  K is built once per fit, B = I + V^-1/2 K V^-1/2 is
// refactorised per Newton step / EP sweep by the DMMA Cholesky, and only a handful of scalars
// cross PCIe per iteration.
#include "fit.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

namespace gpl {

namespace {

// GPL_TIMING=1: per-phase wall times of gpl_energy_grad on stderr (each phase fenced by a stream
// synchronise, so only meaningful for profiling runs)
struct PhaseTimer {
  Ctx* ctx;
  bool on;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseTimer(Ctx* c) : ctx(c) {
    static const int env = getenv("GPL_TIMING") ? atoi(getenv("GPL_TIMING")) : 0;
    on = env != 0;
    if (on) {
      cudaStreamSynchronize(ctx->stream);
      t0 = std::chrono::steady_clock::now();
    }
  }
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[gpl timing] %-28s %9.3f ms\n", what,
            std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

__global__ void diag_extract_kernel(int64_t n, const double* __restrict__ A, int64_t lda,
                                    double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = A[i * lda + i];
}

__global__ void copy_lower_kernel(int64_t n, const double* __restrict__ A, int64_t lda,
                                  double* __restrict__ B, int64_t ldb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j0 = (int64_t)blockIdx.y * 8;
  if ((int64_t)(blockIdx.x + 1) * blockDim.x <= j0 || i >= n) return;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int64_t j = j0 + c;
    if (j < n && i >= j) B[j * ldb + i] = A[j * lda + i];
  }
}

// out (n x n, ld n) = diag(sqrtV) L, upper triangle zero  (gp$fit$C_chol, R/laplace.R:28)
__global__ void cchol_kernel(int64_t n, const double* __restrict__ L, int64_t ldl,
                             const double* __restrict__ sqrtV, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i >= n) return;
  out[j * n + i] = (i >= j) ? sqrtV[i] * L[j * ldl + i] : 0.0;
}

int diag_extract(Ctx* ctx, int64_t n, const double* A, int64_t lda, double* out) {
  if (n <= 0) return GPL_OK;
  diag_extract_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, A, lda, out);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int copy_lower(Ctx* ctx, int64_t n, const double* A, int64_t lda, double* B, int64_t ldb) {
  if (n <= 0) return GPL_OK;
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)((n + 7) / 8));
  copy_lower_kernel<<<grid, 256, 0, ctx->stream>>>(n, A, lda, B, ldb);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int upload(Ctx* ctx, DevBuf& buf, const double* host, size_t count) {
  GPL_TRY(buf.alloc(ctx, sizeof(double) * count));
  if (count)
    GPL_CUDA(cudaMemcpyAsync(buf.p, host, sizeof(double) * count, cudaMemcpyHostToDevice,
                             ctx->stream));
  return GPL_OK;
}

int download(Ctx* ctx, double* host, const double* dev, size_t count) {
  if (host && count)
    GPL_CUDA(cudaMemcpyAsync(host, dev, sizeof(double) * count, cudaMemcpyDeviceToHost,
                             ctx->stream));
  return GPL_OK;
}

// alpha = C^-1 z with C = diag(sqrtV) Lb Lb^T diag(sqrtV)   (R/laplace.R:29)
int solve_alpha(Ctx* ctx, int64_t n, const double* Lb, int64_t ld, const double* dinv,
                const double* sqrtV, const double* z, double* alpha, double* tmp1, double* tmp2) {
  GPL_TRY(vec_op(ctx, VOP_DIV, n, z, sqrtV, tmp1));
  GPL_TRY(trsv_lower(ctx, dinv, false, n, Lb, ld, tmp1, tmp2));
  GPL_TRY(trsv_lower(ctx, dinv, true, n, Lb, ld, tmp2, tmp1));
  GPL_TRY(vec_op(ctx, VOP_DIV, n, tmp1, sqrtV, alpha));
  return GPL_OK;
}

struct FullState {
  int64_t n = 0, ld = 0;
  int d = 0;
  DevBuf x, y, trials, offset;
  DevBuf Z;         // features, n x F
  DevBuf Zaux;
  DevBuf K;         // n x n (ld), symmetric, jitter on the diagonal
  DevBuf B;         // n x n (ld) work / factor
  DevBuf dinv;      // block inverses of the factor in B
  DevBuf v[16];     // n-vectors
  HostPlan hp;
  LikDev lik;
};

enum { V_F = 0, V_Z, V_V, V_S, V_SQRTV, V_ALPHA, V_FNEW, V_T1, V_T2, V_DIAG, V_FVAR, V_P, V_ZOLD,
       V_CM, V_CV, V_LOGZ };

int full_setup(Ctx* ctx, FullState& st, const gpl_cf_model* cf, const gpl_lik* lik, const void* x,
               int64_t n, int d, const void* y, const double* trials, const double* offset,
               bool on_device, double jitter, bool want_aux) {
  st.n = n;
  st.d = d;
  st.ld = round_up(n, 16);
  GPL_TRY(st.x.alloc(ctx, sizeof(double) * n * d));
  GPL_TRY(st.y.alloc(ctx, sizeof(double) * n));
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  GPL_CUDA(cudaMemcpyAsync(st.x.p, x, sizeof(double) * n * d, kind, ctx->stream));
  GPL_CUDA(cudaMemcpyAsync(st.y.p, y, sizeof(double) * n, kind, ctx->stream));
  st.lik.kind = lik->kind;
  st.lik.link = lik->link;
  st.lik.param = lik->param;
  st.lik.trials = nullptr;
  st.lik.offset = nullptr;
  if (trials) {
    GPL_TRY(upload(ctx, st.trials, trials, n));
    st.lik.trials = st.trials.d();
  }
  if (offset) {
    GPL_TRY(upload(ctx, st.offset, offset, n));
    st.lik.offset = st.offset.d();
  }
  if ((lik->kind == GPL_LIK_BINOMIAL || lik->kind == GPL_LIK_BETABINOM) && !trials) {
    set_error("trials must be provided for the binomial / beta-binomial likelihood");
    return GPL_ERR_BAD_ARG;
  }
  GPL_TRY(make_plan(cf, d, 0, &st.hp, want_aux));
  const int F = st.hp.plan.nfeat, FA = st.hp.plan.naux;
  GPL_TRY(st.Z.alloc(ctx, sizeof(double) * n * (F > 0 ? F : 1)));
  if (FA > 0) GPL_TRY(st.Zaux.alloc(ctx, sizeof(double) * n * FA));
  GPL_TRY(prep_features(ctx, st.hp, st.x.d(), n, n, st.Z.d(), n, FA > 0 ? st.Zaux.d() : nullptr, n));
  GPL_TRY(st.K.alloc(ctx, sizeof(double) * st.ld * n));
  GPL_TRY(cf_pairwise(ctx, st.hp.plan, st.Z.d(), n, n, nullptr, n, n, true, jitter, st.K.d(), st.ld));
  GPL_TRY(st.B.alloc(ctx, sizeof(double) * st.ld * n));
  GPL_TRY(st.dinv.alloc(ctx, sizeof(double) * ((n + NB - 1) / NB) * NB * NB));
  for (int i = 0; i < 16; ++i) GPL_TRY(st.v[i].alloc(ctx, sizeof(double) * n));
  return GPL_OK;
}

// one "K, z, V -> B factor, alpha, f = K alpha" step shared by Laplace and EP
int factor_and_solve(Ctx* ctx, FullState& st) {
  const int64_t n = st.n;
  GPL_TRY(vec_op(ctx, VOP_RSQRT_INV, n, st.v[V_V].d(), nullptr, st.v[V_S].d()));
  GPL_TRY(vec_op(ctx, VOP_SQRT, n, st.v[V_V].d(), nullptr, st.v[V_SQRTV].d()));
  GPL_TRY(build_B(ctx, n, st.K.d(), st.ld, st.v[V_S].d(), st.B.d(), st.ld));
  int32_t info = 0;
  GPL_TRY(dpotrf_with(ctx, st.dinv.d(), n, st.B.d(), st.ld, &info, nullptr));
  GPL_TRY(solve_alpha(ctx, n, st.B.d(), st.ld, st.dinv.d(), st.v[V_SQRTV].d(), st.v[V_Z].d(),
                      st.v[V_ALPHA].d(), st.v[V_T1].d(), st.v[V_T2].d()));
  GPL_TRY(gemv_t(ctx, n, n, st.K.d(), st.ld, st.v[V_ALPHA].d(), st.v[V_FNEW].d()));
  GPL_TRY(diag_extract(ctx, n, st.B.d(), st.ld, st.v[V_DIAG].d()));
  return GPL_OK;
}

int save_fit(Ctx* ctx, FullState& st, Fit** out) {
  if (!out) return GPL_OK;
  Fit* f = new Fit();
  f->method = 0;
  f->n = st.n;
  f->d = st.d;
  f->ld = st.ld;
  f->x.ctx = ctx; f->x.p = st.x.p; st.x.p = nullptr;
  f->Lb.ctx = ctx; f->Lb.p = st.B.p; st.B.p = nullptr;
  f->dinv.ctx = ctx; f->dinv.p = st.dinv.p; st.dinv.p = nullptr;
  f->sqrtV.ctx = ctx; f->sqrtV.p = st.v[V_SQRTV].p; st.v[V_SQRTV].p = nullptr;
  f->alpha.ctx = ctx; f->alpha.p = st.v[V_ALPHA].p; st.v[V_ALPHA].p = nullptr;
  *out = f;
  return GPL_OK;
}

}  // namespace

int fit_full_impl(Ctx* ctx, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                  const double* x, int64_t n, int d, const double* y, const double* trials,
                  const double* offset, double jitter, gpl_fit_out* out, double* h_fmean,
                  double* h_z, double* h_V, double* h_alpha, double* h_fvar, double* h_cm,
                  double* h_cv, Fit** fit) {
  if (n <= 0 || d <= 0 || !x || !y || !out) {
    set_error("gpl_fit_full: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  FullState st;
  GPL_TRY(full_setup(ctx, st, cf, lik, x, n, d, y, trials, offset, false, jitter, false));
  memset(out, 0, sizeof(*out));
  out->quad_ok = 1;
  double sc[8];

  if (approx->kind == GPL_APPROX_LAPLACE) {
    int maxiter = approx->maxiter;
    if (lik->kind == GPL_LIK_GAUSSIAN) maxiter = 1;                        // R/laplace.R:281-283
    if (maxiter < 1) maxiter = 1;
    GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, st.v[V_F].d(), 0.0));
    int it = 0;
    double diff = 0.0, log_evidence = 0.0;
    for (it = 1; it <= maxiter; ++it) {
      GPL_TRY(lik_pseudodata_la(ctx, st.lik, n, st.v[V_F].d(), st.y.d(), st.v[V_Z].d(),
                                st.v[V_V].d()));
      GPL_TRY(reduce(ctx, RED_MIN, n, st.v[V_V].d(), nullptr, 0));
      GPL_TRY(fetch_scalars(ctx, 1, sc));
      if (!(sc[0] > 0.0)) {
        set_error("Laplace: pseudo-variance <= 0 (non-log-concave site); the reference's bad-site "
                  "branch (R/laplace.R:40-117) is not implemented");
        return GPL_ERR_NOT_LOGCONCAVE;
      }
      GPL_TRY(factor_and_solve(ctx, st));
      GPL_TRY(reduce(ctx, RED_SUMLOG_PROD, n, st.v[V_SQRTV].d(), st.v[V_DIAG].d(), 0));
      GPL_TRY(reduce(ctx, RED_SUMLOG, n, st.v[V_V].d(), nullptr, 1));
      GPL_TRY(reduce(ctx, RED_DOT, n, st.v[V_FNEW].d(), st.v[V_ALPHA].d(), 2));
      GPL_TRY(lik_loglik_sum(ctx, st.lik, n, st.v[V_FNEW].d(), st.y.d(), 3));
      GPL_TRY(reduce(ctx, RED_MAXABSDIFF, n, st.v[V_F].d(), st.v[V_FNEW].d(), 4));
      GPL_TRY(fetch_scalars(ctx, 5, sc));
      const double C_logdet = 2.0 * sc[0], V_logdet = sc[1];
      log_evidence = sc[3] - 0.5 * sc[2] - 0.5 * (C_logdet - V_logdet);   // R/laplace.R:37
      diff = sc[4];
      std::swap(st.v[V_F].p, st.v[V_FNEW].p);
      if (diff < approx->tol) break;
    }
    if (it > maxiter) it = maxiter;
    out->log_evidence = log_evidence;
    out->iters = it;
    out->converged = !(maxiter > 1 && it == maxiter);
    GPL_TRY(download(ctx, h_fmean, st.v[V_F].d(), n));
    GPL_TRY(download(ctx, h_z, st.v[V_Z].d(), n));
    GPL_TRY(download(ctx, h_V, st.v[V_V].d(), n));
    GPL_TRY(download(ctx, h_alpha, st.v[V_ALPHA].d(), n));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    return save_fit(ctx, st, fit);
  }

  // ---- parallel EP (R/ep.R:286-330) ----
  if (!approx->quad_x || !approx->quad_w || approx->quad_order < 1) {
    set_error("EP needs Gauss-Hermite nodes and weights");
    return GPL_ERR_BAD_ARG;
  }
  DevBuf qx, qw, X, flags, fmean_old, fvar_old, z_old, P_old;
  GPL_TRY(upload(ctx, qx, approx->quad_x, approx->quad_order));
  GPL_TRY(upload(ctx, qw, approx->quad_w, approx->quad_order));
  GPL_TRY(X.alloc(ctx, sizeof(double) * st.ld * n));
  GPL_TRY(flags.alloc(ctx, sizeof(int) * 4));
  GPL_TRY(fmean_old.alloc(ctx, sizeof(double) * n));
  GPL_TRY(fvar_old.alloc(ctx, sizeof(double) * n));
  GPL_TRY(z_old.alloc(ctx, sizeof(double) * n));
  GPL_TRY(P_old.alloc(ctx, sizeof(double) * n));
  GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, fmean_old.d(), 0.0));
  GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, z_old.d(), 0.0));
  GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, P_old.d(), 0.0));
  GPL_TRY(diag_extract(ctx, n, st.K.d(), st.ld, fvar_old.d()));
  double diff_old = INFINITY, damp = approx->damping, log_evidence = 0.0, diff_new = 0.0;
  const int maxiter = approx->maxiter < 1 ? 1 : approx->maxiter;
  int it = 0, quad_ok = 1;
  for (it = 1; it <= maxiter; ++it) {
    EpSiteArgs ea;
    ea.fmean = fmean_old.d(); ea.fvar = fvar_old.d(); ea.z_old = z_old.d(); ea.P_old = P_old.d();
    ea.y = st.y.d();
    ea.z_new = st.v[V_Z].d(); ea.V_new = st.v[V_V].d(); ea.P_new = st.v[V_P].d();
    ea.cav_mean = st.v[V_CM].d(); ea.cav_var = st.v[V_CV].d(); ea.logZ = st.v[V_LOGZ].d();
    ea.damping = damp;
    ea.quad_order = approx->quad_order;
    ea.quad_x = qx.d(); ea.quad_w = qw.d();
    GPL_TRY(lik_ep_sites(ctx, st.lik, n, ea, static_cast<int*>(flags.p)));
    int hflags = 0;
    GPL_CUDA(cudaMemcpyAsync(&hflags, flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    quad_ok = (hflags & 1) ? 0 : 1;
    if (hflags & 2) {                                                      // R/ep.R:30-32
      out->log_evidence = -INFINITY;
      out->iters = it;
      out->converged = 0;
      out->quad_ok = quad_ok;
      set_error("EP: NaN pseudo-data");
      return GPL_ERR_NONFINITE;
    }
    if (hflags & 4) {
      set_error("EP: pseudo-variance <= 0 (non-log-concave site); the reference's bad-site branch "
                "(R/ep.R:50-138) is not implemented");
      return GPL_ERR_NOT_LOGCONCAVE;
    }
    GPL_TRY(factor_and_solve(ctx, st));
    // posterior marginal variances: diag(V C^-1 K) = V (1 - diag(B^-1)), diag(B^-1)_i = |L^-1 e_i|^2
    GPL_TRY(copy_lower(ctx, n, st.B.d(), st.ld, X.d(), st.ld));
    GPL_TRY(dtrtri_with(ctx, st.dinv.d(), n, X.d(), st.ld));
    GPL_TRY(lower_colsumsq(ctx, n, X.d(), st.ld, st.v[V_T1].d()));
    GPL_TRY(vec_op(ctx, VOP_MUL, n, st.v[V_V].d(), st.v[V_T1].d(), st.v[V_T2].d()));
    GPL_TRY(vec_op(ctx, VOP_SUB, n, st.v[V_V].d(), st.v[V_T2].d(), st.v[V_FVAR].d()));
    GPL_TRY(reduce(ctx, RED_SUMLOG_PROD, n, st.v[V_SQRTV].d(), st.v[V_DIAG].d(), 0));
    GPL_TRY(reduce(ctx, RED_DOT, n, st.v[V_Z].d(), st.v[V_ALPHA].d(), 1));
    GPL_TRY(ep_evidence_terms(ctx, n, st.v[V_LOGZ].d(), st.v[V_CV].d(), st.v[V_CM].d(),
                              st.v[V_Z].d(), st.v[V_V].d(), 2));
    GPL_TRY(reduce(ctx, RED_MAXABSDIFF, n, fmean_old.d(), st.v[V_FNEW].d(), 3));
    GPL_TRY(fetch_scalars(ctx, 4, sc));
    log_evidence = -0.5 * (2.0 * sc[0]) - 0.5 * sc[1] + sc[2];            // R/ep.R:45-48
    diff_new = sc[3];
    if (diff_new > diff_old) {                                             // R/ep.R:306-310
      damp *= approx->damping;
      continue;
    }
    damp = damp > approx->damping_min ? damp : approx->damping_min;
    GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[V_FNEW].d(), nullptr, fmean_old.d()));
    GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[V_FVAR].d(), nullptr, fvar_old.d()));
    GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[V_Z].d(), nullptr, z_old.d()));
    GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[V_P].d(), nullptr, P_old.d()));
    diff_old = diff_new;
    if (diff_new < approx->tol) break;
  }
  if (it > maxiter) it = maxiter;
  out->iters = it;
  out->converged = !(maxiter > 1 && it == maxiter);
  out->quad_ok = quad_ok;
  if (maxiter > 1 && it == maxiter) log_evidence = -INFINITY;              // R/ep.R:322-325
  if (!quad_ok) log_evidence = -INFINITY;                                  // R/ep.R:326-328
  out->log_evidence = log_evidence;
  GPL_TRY(download(ctx, h_fmean, st.v[V_FNEW].d(), n));
  GPL_TRY(download(ctx, h_fvar, st.v[V_FVAR].d(), n));
  GPL_TRY(download(ctx, h_z, st.v[V_Z].d(), n));
  GPL_TRY(download(ctx, h_V, st.v[V_V].d(), n));
  GPL_TRY(download(ctx, h_alpha, st.v[V_ALPHA].d(), n));
  GPL_TRY(download(ctx, h_cm, st.v[V_CM].d(), n));
  GPL_TRY(download(ctx, h_cv, st.v[V_CV].d(), n));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  return save_fit(ctx, st, fit);
}

int energy_grad_impl(Ctx* ctx, const gpl_cf_model* cf, const gpl_lik* lik, const void* x, int64_t n,
                     int d, const void* y, bool on_device, double jitter, int nparam,
                     const int32_t* param_factor, const int32_t* param_slot, double* energy,
                     double* grad, Fit** fit) {
  if (lik->kind != GPL_LIK_GAUSSIAN) {
    set_error("gpl_energy_grad: analytic gradient is implemented for lik_gaussian + method_full");
    return GPL_ERR_BAD_ARG;
  }
  if (n <= 0 || d <= 0 || !x || !y || !energy || nparam < 0 || nparam > MAX_GRAD_PARAMS) {
    set_error("gpl_energy_grad: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  bool want_aux = false;
  for (int j = 0; j < nparam; ++j)
    if (param_factor[j] >= 0 && param_slot[j] == 3) want_aux = true;
  FullState st;
  PhaseTimer tm(ctx);
  GPL_TRY(full_setup(ctx, st, cf, lik, x, n, d, y, nullptr, nullptr, on_device, jitter, want_aux));
  tm.mark("upload + features + K build");
  // map model factor indices to plan factor indices (identical: term_mask == 0 keeps all)
  GradParam gp[MAX_GRAD_PARAMS];
  int ncf = 0;
  int cf_pos[MAX_GRAD_PARAMS];
  for (int j = 0; j < nparam; ++j) {
    if (param_factor[j] < 0) continue;
    if (param_factor[j] >= st.hp.plan.nfactors) {
      set_error("gpl_energy_grad: parameter %d refers to factor %d of %d", j, param_factor[j],
                st.hp.plan.nfactors);
      return GPL_ERR_BAD_ARG;
    }
    const gpl_cf_factor& f = cf->factors[param_factor[j]];
    if (param_slot[j] == 3 && (!f.periodic || f.kind == GPL_CF_NN || f.kind == GPL_CF_CONST)) {
      set_error("gpl_energy_grad: period gradient needs a periodic sexp/matern/lin factor");
      return GPL_ERR_BAD_ARG;
    }
    gp[ncf].factor = param_factor[j];
    gp[ncf].slot = param_slot[j];
    cf_pos[ncf] = j;
    ++ncf;
  }
  double sc[8];
  GPL_TRY(lik_pseudodata_la(ctx, st.lik, n, st.v[V_F].d(), st.y.d(), st.v[V_Z].d(), st.v[V_V].d()));
  GPL_TRY(factor_and_solve(ctx, st));
  tm.mark("B build + potrf + alpha + K alpha");
  GPL_TRY(reduce(ctx, RED_SUMLOG_PROD, n, st.v[V_SQRTV].d(), st.v[V_DIAG].d(), 0));
  GPL_TRY(reduce(ctx, RED_SUMLOG, n, st.v[V_V].d(), nullptr, 1));
  GPL_TRY(reduce(ctx, RED_DOT, n, st.v[V_FNEW].d(), st.v[V_ALPHA].d(), 2));
  GPL_TRY(lik_loglik_sum(ctx, st.lik, n, st.v[V_FNEW].d(), st.y.d(), 3));
  GPL_TRY(reduce(ctx, RED_SUMSQ, n, st.v[V_ALPHA].d(), nullptr, 4));
  if (nparam > 0 && grad) {
    DevBuf keep;
    if (fit) {
      GPL_TRY(keep.alloc(ctx, sizeof(double) * st.ld * n));
      GPL_TRY(copy_lower(ctx, n, st.B.d(), st.ld, keep.d(), st.ld));
    }
    // B^-1 = X^T X with X = L_B^-1: trtri in place, then ONE triangular-aware SYRK into the K
    // buffer (K is dead from here on: the trace kernel regenerates dK/dtheta from the features)
    GPL_TRY(dtrtri_with(ctx, st.dinv.d(), n, st.B.d(), st.ld));
    tm.mark("trtri");
    GPL_TRY(lauum_oop(ctx, n, st.B.d(), st.ld, st.K.d(), st.ld));
    tm.mark("lauum (single SYRK)");
    const double* Binv = st.K.d();
    DevBuf gout;
    GPL_TRY(gout.alloc(ctx, sizeof(double) * (MAX_GRAD_PARAMS + 8)));
    GPL_TRY(grad_trace(ctx, st.hp.plan, st.Z.d(), n, st.Zaux.d(), n, Binv, st.ld,
                       st.v[V_S].d(), st.v[V_ALPHA].d(), ncf, gp, gout.d()));
    // sigma: dC/dlog sigma = 2 sigma^2 I  ->  tr(W) = sum alpha^2 - sum s_i^2 Binv_ii
    GPL_TRY(diag_extract(ctx, n, Binv, st.ld, st.v[V_T1].d()));
    GPL_TRY(vec_op(ctx, VOP_MUL, n, st.v[V_S].d(), st.v[V_S].d(), st.v[V_T2].d()));
    GPL_TRY(reduce(ctx, RED_DOT, n, st.v[V_T1].d(), st.v[V_T2].d(), 5));
    GPL_TRY(fetch_scalars(ctx, 6, sc));
    double hg[MAX_GRAD_PARAMS];
    GPL_CUDA(cudaMemcpyAsync(hg, gout.p, sizeof(double) * (ncf > 0 ? ncf : 1),
                             cudaMemcpyDeviceToHost, ctx->stream));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    tm.mark("gradient trace + reductions");
    for (int q = 0; q < ncf; ++q) grad[cf_pos[q]] = -0.5 * hg[q];
    const double sigma2 = lik->param * lik->param;
    for (int j = 0; j < nparam; ++j)
      if (param_factor[j] < 0) grad[j] = -0.5 * 2.0 * sigma2 * (sc[4] - sc[5]);
    if (fit) {
      std::swap(st.B.p, keep.p);
    }
  } else {
    GPL_TRY(fetch_scalars(ctx, 5, sc));
  }
  const double log_evidence = sc[3] - 0.5 * sc[2] - 0.5 * (2.0 * sc[0] - sc[1]);
  *energy = -log_evidence;
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (fit) return save_fit(ctx, st, fit);
  return GPL_OK;
}

int pred_full_impl(Ctx* ctx, const Fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                   uint64_t term_mask, double* mean, double* var) {
  if (!fit || fit->method != 0 || nt < 0 || !mean) {
    set_error("gpl_pred_full: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  if (nt == 0) return GPL_OK;
  const int64_t n = fit->n;
  const int d = fit->d;
  HostPlan hp;
  GPL_TRY(make_plan(cf, d, term_mask, &hp));
  const int F = hp.plan.nfeat > 0 ? hp.plan.nfeat : 1;
  DevBuf dxt, Zx, Zt, Kt, dmean, dvar, tmp, sinv;
  GPL_TRY(upload(ctx, dxt, xt, (size_t)nt * d));
  GPL_TRY(Zx.alloc(ctx, sizeof(double) * n * F));
  GPL_TRY(Zt.alloc(ctx, sizeof(double) * nt * F));
  GPL_TRY(prep_features(ctx, hp, fit->x.d(), n, n, Zx.d(), n));
  GPL_TRY(prep_features(ctx, hp, dxt.d(), nt, nt, Zt.d(), nt));
  // chunk the test points so that the nt x n cross-covariance stays <= ~4 GB
  int64_t chunk = (int64_t)(4e9 / (8.0 * (double)n));
  chunk = chunk / 128 * 128;
  if (chunk < 128) chunk = 128;
  if (chunk > nt) chunk = nt;
  const int64_t ldk = round_up(chunk, 16);
  GPL_TRY(Kt.alloc(ctx, sizeof(double) * ldk * n));
  GPL_TRY(dmean.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(dvar.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(tmp.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(sinv.alloc(ctx, sizeof(double) * n));
  if (var) GPL_TRY(vec_op(ctx, VOP_INV, n, fit->sqrtV.d(), nullptr, sinv.d()));
  for (int64_t r0 = 0; r0 < nt; r0 += chunk) {
    const int64_t c = (nt - r0) < chunk ? (nt - r0) : chunk;
    GPL_TRY(cf_pairwise(ctx, hp.plan, Zt.d() + r0, c, nt, Zx.d(), n, n, false, 0.0, Kt.d(), ldk));
    GPL_TRY(gemv_n(ctx, c, n, Kt.d(), ldk, fit->alpha.d(), dmean.d() + r0));
    if (var) {
      GPL_TRY(cf_diag(ctx, hp.plan, Zt.d() + r0, nt, Zt.d() + r0, nt, c, 0.0, dvar.d() + r0));
      GPL_TRY(scale_cols(ctx, c, n, Kt.d(), ldk, sinv.d()));
      GPL_TRY(dtrsm_right_with(ctx, fit->dinv.d(), true, c, n, fit->Lb.d(), fit->ld, Kt.d(), ldk));
      GPL_TRY(rowsumsq(ctx, c, n, Kt.d(), ldk, tmp.d() + r0));
    }
  }
  if (var) {
    GPL_TRY(vec_op(ctx, VOP_SUB, nt, dvar.d(), tmp.d(), dvar.d()));
    GPL_TRY(download(ctx, var, dvar.d(), nt));
  }
  GPL_TRY(download(ctx, mean, dmean.d(), nt));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPL_OK;
}

int fit_get_impl(Ctx* ctx, const Fit* fit, const char* what, double* out) {
  if (!fit || !what || !out) return GPL_ERR_BAD_ARG;
  const int64_t n = fit->n;
  if (!strcmp(what, "alpha")) {
    GPL_TRY(download(ctx, out, fit->alpha.d(), n));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    return GPL_OK;
  }
  if (!strcmp(what, "C_chol") && fit->method == 0) {
    DevBuf tmp;
    GPL_TRY(tmp.alloc(ctx, sizeof(double) * n * n));
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
    cchol_kernel<<<grid, 256, 0, ctx->stream>>>(n, fit->Lb.d(), fit->ld, fit->sqrtV.d(), tmp.d());
    GPL_LAUNCH_CHECK(ctx);
    GPL_TRY(download(ctx, out, tmp.d(), (size_t)n * n));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    return GPL_OK;
  }
  if (fit->method == 1 && (!strcmp(what, "Kxz") || !strcmp(what, "Kz_chol"))) {
    const bool kxz = !strcmp(what, "Kxz");
    const int64_t rows = kxz ? fit->n : fit->m, cols = fit->m;
    const double* src = kxz ? fit->Kxz.d() : fit->Kz_chol.d();
    const int64_t lds = kxz ? fit->ldkxz : fit->ldm;
    GPL_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * rows, src, sizeof(double) * lds,
                               sizeof(double) * rows, cols, cudaMemcpyDeviceToHost, ctx->stream));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!kxz) {          // the strict upper triangle of the factor buffer is scratch: return zeros
      for (int64_t j = 1; j < cols; ++j)
        for (int64_t i = 0; i < j; ++i) out[j * rows + i] = 0.0;
    }
    return GPL_OK;
  }
  set_error("gpl_fit_get: unknown member '%s'", what);
  return GPL_ERR_BAD_ARG;
}

}  // namespace gpl

using namespace gpl;

extern "C" {

int gpl_fit_full(gpl_ctx* h, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                 const double* x, int64_t n, int32_t d, const double* y, const double* trials,
                 const double* offset, double jitter, gpl_fit_out* out, double* fmean, double* z,
                 double* V, double* alpha, double* fvar, double* cavity_mean, double* cavity_var,
                 gpl_fit** fit) {
  Fit* f = nullptr;
  if (fit) *fit = nullptr;
  int rc = fit_full_impl(reinterpret_cast<Ctx*>(h), cf, lik, approx, x, n, d, y, trials, offset,
                         jitter, out, fmean, z, V, alpha, fvar, cavity_mean, cavity_var,
                         fit ? &f : nullptr);
  if (fit) *fit = reinterpret_cast<gpl_fit*>(f);
  if (rc != GPL_OK) cudaStreamSynchronize(reinterpret_cast<Ctx*>(h)->stream);
  return rc;
}

int gpl_energy_grad(gpl_ctx* h, const gpl_cf_model* cf, const gpl_lik* lik, const void* x, int64_t n,
                    int32_t d, const void* y, int32_t on_device, double jitter, int32_t nparam,
                    const int32_t* param_factor, const int32_t* param_slot, double* energy,
                    double* grad, gpl_fit** fit) {
  Fit* f = nullptr;
  if (fit) *fit = nullptr;
  int rc = energy_grad_impl(reinterpret_cast<Ctx*>(h), cf, lik, x, n, d, y, on_device != 0, jitter,
                            nparam, param_factor, param_slot, energy, grad, fit ? &f : nullptr);
  if (fit) *fit = reinterpret_cast<gpl_fit*>(f);
  if (rc != GPL_OK) cudaStreamSynchronize(reinterpret_cast<Ctx*>(h)->stream);
  return rc;
}

int gpl_pred_full(gpl_ctx* h, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt,
                  int64_t nt, uint64_t term_mask, double* mean, double* var) {
  return pred_full_impl(reinterpret_cast<Ctx*>(h), reinterpret_cast<const Fit*>(fit), cf, xt, nt,
                        term_mask, mean, var);
}

int gpl_fit_get(gpl_ctx* h, const gpl_fit* fit, const char* what, double* out) {
  return fit_get_impl(reinterpret_cast<Ctx*>(h), reinterpret_cast<const Fit*>(fit), what, out);
}

int gpl_fit_free(gpl_ctx* h, gpl_fit* fit) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  Fit* f = reinterpret_cast<Fit*>(fit);
  if (f) {
    delete f;
    cudaStreamSynchronize(c->stream);
  }
  return GPL_OK;
}

}  // extern "C"
