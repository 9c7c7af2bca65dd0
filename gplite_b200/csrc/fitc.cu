// method_fitc (R/gp_fit.R:101-123,163-188; R/laplace.R:122-152,299-325; R/ep.R:148-185,333-374;
// R/linear_algebra.R:3-126; R/gp_pred.R:234-296).
//
// n x m work (Kxz, 65.5 GB at n = 2M, m = 4096) is built once and kept resident; everything
// that the reference does with further n x m temporaries (forwardsolve(Kz_chol, t(Kxz)),
// inv_D_U, Kt_aux1 ...) is streamed through one row-chunk workspace:
//   D_i     = k_ii + jitter - |Lz^-1 k_zi|^2          right-TRSM on a chunk + row sums of squares
//   A       = Kz + Kzx S^-1 Kxz                        chunk scaled by S^-1/2, SYRK-accumulated (DMMA)
//   alpha   = z/S - S^-1 Kxz A^-1 Kzx (z/S)            two GEMVs over Kxz + m-sized solves
//   fvar_i  = V_i D_i / S_i + (V_i/S_i)^2 |L^-1 k_zi|^2   (EP; == diag1 + diag2 of R/ep.R:173-177)
// Prediction needs no n-sized object at all: with beta = Kz^-1 Kzx alpha cached at fit time,
//   mean_t = k_tz beta,  var_t = k_tt - |Lz^-1 k_zt|^2 + |L^-1 k_zt|^2
// which is R/gp_pred.R:262-291 after substituting the inversion lemma.
#include "fit.cuh"

#include <math.h>
#include <string.h>

namespace gpl {

namespace {

__global__ void scaled_copy_rows_kernel(int64_t rows, int64_t m, const double* __restrict__ src,
                                        int64_t lds, const double* __restrict__ s,
                                        double* __restrict__ dst, int64_t ldd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i >= rows || j >= m) return;
  const double v = src[j * lds + i];
  dst[j * ldd + i] = s ? v * s[i] : v;
}

__global__ void fitc_vec_kernel(int op, int64_t n, const double* __restrict__ a,
                                const double* __restrict__ b, const double* __restrict__ c,
                                const double* __restrict__ d, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  switch (op) {
    case 0: out[i] = a[i] - b[i] / c[i]; break;                 // alpha = t - (Kxz v) / S
    case 1: out[i] = a[i] + b[i] * c[i]; break;                 // f = Kxz w2 + D alpha
    case 2: {                                                   // fvar = V D / S + (V/S)^2 w
      const double r = a[i] / c[i];
      out[i] = r * b[i] + r * r * d[i];
      break;
    }
    case 3: out[i] = 1.0 / sqrt(a[i]); break;                   // S^-1/2
  }
}

int fvec(Ctx* ctx, int op, int64_t n, const double* a, const double* b, const double* c,
         const double* d, double* out) {
  if (n <= 0) return GPL_OK;
  fitc_vec_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(op, n, a, b, c, d, out);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int scaled_copy_rows(Ctx* ctx, int64_t rows, int64_t m, const double* src, int64_t lds,
                     const double* s, double* dst, int64_t ldd) {
  if (rows <= 0 || m <= 0) return GPL_OK;
  dim3 grid((unsigned)((rows + 255) / 256), (unsigned)m);
  scaled_copy_rows_kernel<<<grid, 256, 0, ctx->stream>>>(rows, m, src, lds, s, dst, ldd);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

__global__ void copy_lower_add_kernel(int64_t n, const double* __restrict__ A, int64_t lda,
                                      double* __restrict__ B, int64_t ldb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i < n && j < n && i >= j) B[j * ldb + i] = A[j * lda + i];
}

struct FitcState {
  int64_t n = 0, m = 0, ldx = 0, ldm = 0, chunk = 0, ldw = 0;
  int d = 0;
  DevBuf x, zind, y, trials, offset, Zx, Zz;
  DevBuf Kz;        // m x m with jitter (symmetric, kept)
  DevBuf Lz, Lz_dinv;
  DevBuf A, A_dinv; // m x m work / factor of Kz + Kzx S^-1 Kxz
  DevBuf Kxz;       // n x m
  DevBuf W;         // chunk x m workspace
  DevBuf Kdiag, Dg, S, rs;     // n
  DevBuf v[14];                // n-vectors
  DevBuf mv[6];                // m-vectors
  HostPlan hp;
  LikDev lik;
  double logdet_z = 0.0;
};

enum { F_F = 0, F_Z, F_V, F_ALPHA, F_FNEW, F_T1, F_T2, F_FVAR, F_P, F_CM, F_CV, F_LOGZ, F_W, F_T3 };

// rows-of-(Kxz L^-T) squared norms, streamed through the chunk workspace
int row_norms_via(Ctx* ctx, FitcState& st, const double* L, const double* dinv, double* out) {
  for (int64_t r0 = 0; r0 < st.n; r0 += st.chunk) {
    const int64_t c = (st.n - r0) < st.chunk ? (st.n - r0) : st.chunk;
    GPL_TRY(scaled_copy_rows(ctx, c, st.m, st.Kxz.d() + r0, st.ldx, nullptr, st.W.d(), st.ldw));
    GPL_TRY(dtrsm_right_with(ctx, const_cast<double*>(dinv), true, c, st.m, L, st.ldm, st.W.d(), st.ldw));
    GPL_TRY(rowsumsq(ctx, c, st.m, st.W.d(), st.ldw, out + r0));
  }
  return GPL_OK;
}

int fitc_setup(Ctx* ctx, FitcState& st, const gpl_cf_model* cf, const gpl_lik* lik, const double* x,
               int64_t n, int d, const double* zind, int64_t m, const double* y,
               const double* trials, const double* offset, double jitter) {
  st.n = n; st.m = m; st.d = d;
  st.ldx = round_up(n, 16);
  st.ldm = round_up(m, 16);
  GPL_TRY(st.x.alloc(ctx, sizeof(double) * n * d));
  GPL_TRY(st.zind.alloc(ctx, sizeof(double) * m * d));
  GPL_TRY(st.y.alloc(ctx, sizeof(double) * n));
  GPL_CUDA(cudaMemcpyAsync(st.x.p, x, sizeof(double) * n * d, cudaMemcpyHostToDevice, ctx->stream));
  GPL_CUDA(cudaMemcpyAsync(st.zind.p, zind, sizeof(double) * m * d, cudaMemcpyHostToDevice, ctx->stream));
  GPL_CUDA(cudaMemcpyAsync(st.y.p, y, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  st.lik.kind = lik->kind; st.lik.link = lik->link; st.lik.param = lik->param;
  st.lik.trials = nullptr; st.lik.offset = nullptr;
  if (trials) {
    GPL_TRY(st.trials.alloc(ctx, sizeof(double) * n));
    GPL_CUDA(cudaMemcpyAsync(st.trials.p, trials, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    st.lik.trials = st.trials.d();
  }
  if (offset) {
    GPL_TRY(st.offset.alloc(ctx, sizeof(double) * n));
    GPL_CUDA(cudaMemcpyAsync(st.offset.p, offset, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    st.lik.offset = st.offset.d();
  }
  if ((lik->kind == GPL_LIK_BINOMIAL || lik->kind == GPL_LIK_BETABINOM) && !trials) {
    set_error("trials must be provided for the binomial / beta-binomial likelihood");
    return GPL_ERR_BAD_ARG;
  }
  GPL_TRY(make_plan(cf, d, 0, &st.hp));
  const int F = st.hp.plan.nfeat > 0 ? st.hp.plan.nfeat : 1;
  GPL_TRY(st.Zx.alloc(ctx, sizeof(double) * n * F));
  GPL_TRY(st.Zz.alloc(ctx, sizeof(double) * m * F));
  GPL_TRY(prep_features(ctx, st.hp, st.x.d(), n, n, st.Zx.d(), n));
  GPL_TRY(prep_features(ctx, st.hp, st.zind.d(), m, m, st.Zz.d(), m));
  const size_t mm = sizeof(double) * st.ldm * m;
  GPL_TRY(st.Kz.alloc(ctx, mm));
  GPL_TRY(st.Lz.alloc(ctx, mm));
  GPL_TRY(st.A.alloc(ctx, mm));
  const size_t dv = sizeof(double) * ((m + NB - 1) / NB) * NB * NB;
  GPL_TRY(st.Lz_dinv.alloc(ctx, dv));
  GPL_TRY(st.A_dinv.alloc(ctx, dv));
  GPL_TRY(st.Kxz.alloc(ctx, sizeof(double) * st.ldx * m));
  // chunk: ~2 GB workspace, multiple of 128 rows
  int64_t chunk = (int64_t)(2.0e9 / (8.0 * (double)m));
  chunk = chunk / 128 * 128;
  if (chunk < 1024) chunk = 1024;
  if (chunk > n) chunk = n;
  st.chunk = chunk;
  st.ldw = round_up(chunk, 16);
  GPL_TRY(st.W.alloc(ctx, sizeof(double) * st.ldw * m));
  GPL_TRY(st.Kdiag.alloc(ctx, sizeof(double) * n));
  GPL_TRY(st.Dg.alloc(ctx, sizeof(double) * n));
  GPL_TRY(st.S.alloc(ctx, sizeof(double) * n));
  GPL_TRY(st.rs.alloc(ctx, sizeof(double) * n));
  for (auto& b : st.v) GPL_TRY(b.alloc(ctx, sizeof(double) * n));
  for (auto& b : st.mv) GPL_TRY(b.alloc(ctx, sizeof(double) * m));

  // Kz (+ jitter), its factor, Kxz, diag K(x, x) + jitter          R/gp_fit.R:105-110
  GPL_TRY(cf_pairwise(ctx, st.hp.plan, st.Zz.d(), m, m, nullptr, m, m, true, jitter, st.Kz.d(), st.ldm));
  dim3 g2((unsigned)((m + 255) / 256), (unsigned)m);
  copy_lower_add_kernel<<<g2, 256, 0, ctx->stream>>>(m, st.Kz.d(), st.ldm, st.Lz.d(), st.ldm);
  GPL_LAUNCH_CHECK(ctx);
  int32_t info = 0;
  GPL_TRY(dpotrf_with(ctx, st.Lz_dinv.d(), m, st.Lz.d(), st.ldm, &info, &st.logdet_z));
  GPL_TRY(cf_pairwise(ctx, st.hp.plan, st.Zx.d(), n, n, st.Zz.d(), m, m, false, 0.0, st.Kxz.d(), st.ldx));
  GPL_TRY(cf_diag(ctx, st.hp.plan, st.Zx.d(), n, st.Zx.d(), n, n, jitter, st.Kdiag.d()));
  // D = K_diag - colSums(forwardsolve(Kz_chol, t(Kxz))^2)
  GPL_TRY(row_norms_via(ctx, st, st.Lz.d(), st.Lz_dinv.d(), st.v[F_T1].d()));
  GPL_TRY(vec_op(ctx, VOP_SUB, n, st.Kdiag.d(), st.v[F_T1].d(), st.Dg.d()));
  return GPL_OK;
}

// inv_lemma_get + alpha + f for the current z (v[F_Z]) and V (v[F_V]);   R/linear_algebra.R:3-43,
// R/laplace.R:139-141.  Leaves: S, factor in A (+A_dinv), alpha, fnew, and scalars
//   slot 0: sum log S, slot 1: 2 sum log diag L (via logdetA), slot 2: f . alpha
int fitc_core(Ctx* ctx, FitcState& st, double* logdetA) {
  const int64_t n = st.n, m = st.m;
  GPL_TRY(vec_op(ctx, VOP_ADD, n, st.Dg.d(), st.v[F_V].d(), st.S.d()));
  GPL_TRY(fvec(ctx, 3, n, st.S.d(), nullptr, nullptr, nullptr, st.rs.d()));
  // A = Kz + Kzx S^-1 Kxz  (lower)
  dim3 g2((unsigned)((m + 255) / 256), (unsigned)m);
  copy_lower_add_kernel<<<g2, 256, 0, ctx->stream>>>(m, st.Kz.d(), st.ldm, st.A.d(), st.ldm);
  GPL_LAUNCH_CHECK(ctx);
  for (int64_t r0 = 0; r0 < n; r0 += st.chunk) {
    const int64_t c = (n - r0) < st.chunk ? (n - r0) : st.chunk;
    GPL_TRY(scaled_copy_rows(ctx, c, m, st.Kxz.d() + r0, st.ldx, st.rs.d() + r0, st.W.d(), st.ldw));
    GPL_TRY(dgemm(ctx, true, false, m, m, c, 1.0, st.W.d(), st.ldw, st.W.d(), st.ldw, 1.0, st.A.d(),
                  st.ldm, true));
  }
  int32_t info = 0;
  GPL_TRY(dpotrf_with(ctx, st.A_dinv.d(), m, st.A.d(), st.ldm, &info, logdetA));
  // alpha = z/S - S^-1 Kxz A^-1 Kzx (z/S)
  GPL_TRY(vec_op(ctx, VOP_DIV, n, st.v[F_Z].d(), st.S.d(), st.v[F_T1].d()));
  GPL_TRY(gemv_t(ctx, n, m, st.Kxz.d(), st.ldx, st.v[F_T1].d(), st.mv[0].d()));
  GPL_TRY(trsv_lower(ctx, st.A_dinv.d(), false, m, st.A.d(), st.ldm, st.mv[0].d(), st.mv[1].d()));
  GPL_TRY(trsv_lower(ctx, st.A_dinv.d(), true, m, st.A.d(), st.ldm, st.mv[1].d(), st.mv[0].d()));
  GPL_TRY(gemv_n(ctx, n, m, st.Kxz.d(), st.ldx, st.mv[0].d(), st.v[F_T2].d()));
  GPL_TRY(fvec(ctx, 0, n, st.v[F_T1].d(), st.v[F_T2].d(), st.S.d(), nullptr, st.v[F_ALPHA].d()));
  // f = Kxz Kz^-1 Kzx alpha + D alpha ;  beta = Kz^-1 Kzx alpha is kept in mv[2]
  GPL_TRY(gemv_t(ctx, n, m, st.Kxz.d(), st.ldx, st.v[F_ALPHA].d(), st.mv[3].d()));
  GPL_TRY(trsv_lower(ctx, st.Lz_dinv.d(), false, m, st.Lz.d(), st.ldm, st.mv[3].d(), st.mv[4].d()));
  GPL_TRY(trsv_lower(ctx, st.Lz_dinv.d(), true, m, st.Lz.d(), st.ldm, st.mv[4].d(), st.mv[2].d()));
  GPL_TRY(gemv_n(ctx, n, m, st.Kxz.d(), st.ldx, st.mv[2].d(), st.v[F_T2].d()));
  GPL_TRY(fvec(ctx, 1, n, st.v[F_T2].d(), st.Dg.d(), st.v[F_ALPHA].d(), nullptr, st.v[F_FNEW].d()));
  return GPL_OK;
}

int save_fitc(Ctx* ctx, FitcState& st, Fit** out) {
  if (!out) return GPL_OK;
  Fit* f = new Fit();
  f->method = 1;
  f->n = st.n; f->m = st.m; f->d = st.d;
  f->ld = st.ldx; f->ldm = st.ldm; f->ldkxz = st.ldx;
  auto take = [&](DevBuf& dst, DevBuf& src) { dst.ctx = ctx; dst.p = src.p; src.p = nullptr; };
  take(f->x, st.x);
  take(f->zind, st.zind);
  take(f->Kxz, st.Kxz);
  take(f->Kz_chol, st.Lz);
  take(f->Kz_dinv, st.Lz_dinv);
  take(f->Lm, st.A);
  take(f->Lm_dinv, st.A_dinv);
  take(f->alpha, st.v[F_ALPHA]);
  take(f->Dg, st.Dg);
  take(f->S, st.S);
  take(f->sqrtV, st.mv[2]);      // beta = Kz^-1 Kzx alpha (m-vector) rides in the sqrtV slot
  *out = f;
  return GPL_OK;
}

}  // namespace

int fit_fitc_impl(Ctx* ctx, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                  const double* x, int64_t n, int d, const double* zind, int64_t m, const double* y,
                  const double* trials, const double* offset, double jitter, gpl_fit_out* out,
                  double* h_fmean, double* h_z, double* h_V, double* h_alpha, double* h_fvar,
                  double* h_cm, double* h_cv, Fit** fit) {
  if (n <= 0 || d <= 0 || m <= 0 || !x || !y || !zind || !out) {
    set_error("gpl_fit_fitc: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  FitcState st;
  GPL_TRY(fitc_setup(ctx, st, cf, lik, x, n, d, zind, m, y, trials, offset, jitter));
  memset(out, 0, sizeof(*out));
  out->quad_ok = 1;
  double sc[8];
  auto dl = [&](double* h, const double* dptr) -> int {
    if (h) GPL_CUDA(cudaMemcpyAsync(h, dptr, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    return GPL_OK;
  };

  if (approx->kind == GPL_APPROX_LAPLACE) {                                  // R/laplace.R:299-325
    int maxiter = approx->maxiter;
    if (lik->kind == GPL_LIK_GAUSSIAN) maxiter = 1;
    if (maxiter < 1) maxiter = 1;
    GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, st.v[F_F].d(), 0.0));
    double log_post_old = -INFINITY, log_evidence = 0.0;
    int it = 0;
    for (it = 1; it <= maxiter; ++it) {
      GPL_TRY(lik_pseudodata_la(ctx, st.lik, n, st.v[F_F].d(), st.y.d(), st.v[F_Z].d(), st.v[F_V].d()));
      GPL_TRY(reduce(ctx, RED_MIN, n, st.v[F_V].d(), nullptr, 0));
      GPL_TRY(fetch_scalars(ctx, 1, sc));
      if (!(sc[0] > 0.0)) {
        set_error("Laplace (FITC): pseudo-variance <= 0; the reference's bad-site branch "
                  "(R/laplace.R:153-232) is not implemented");
        return GPL_ERR_NOT_LOGCONCAVE;
      }
      double logdetA = 0.0;
      GPL_TRY(fitc_core(ctx, st, &logdetA));
      GPL_TRY(reduce(ctx, RED_SUMLOG, n, st.S.d(), nullptr, 0));
      GPL_TRY(reduce(ctx, RED_SUMLOG, n, st.v[F_V].d(), nullptr, 1));
      GPL_TRY(reduce(ctx, RED_DOT, n, st.v[F_FNEW].d(), st.v[F_ALPHA].d(), 2));
      GPL_TRY(lik_loglik_sum(ctx, st.lik, n, st.v[F_FNEW].d(), st.y.d(), 3));
      GPL_TRY(fetch_scalars(ctx, 4, sc));
      const double C_logdet = sc[0] - st.logdet_z + logdetA;               // R/linear_algebra.R:32
      log_evidence = sc[3] - 0.5 * sc[2] + 0.5 * (sc[1] - C_logdet);       // R/laplace.R:143-150
      const double log_post = sc[3] - 0.5 * sc[2];
      std::swap(st.v[F_F].p, st.v[F_FNEW].p);
      const double diff = log_post - log_post_old;
      log_post_old = log_post;
      if (fabs(diff) < approx->tol) break;
    }
    if (it > maxiter) it = maxiter;
    out->log_evidence = log_evidence;
    out->iters = it;
    out->converged = !(maxiter > 1 && it == maxiter);
    GPL_TRY(dl(h_fmean, st.v[F_F].d()));
    GPL_TRY(dl(h_z, st.v[F_Z].d()));
    GPL_TRY(dl(h_V, st.v[F_V].d()));
    GPL_TRY(dl(h_alpha, st.v[F_ALPHA].d()));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    return save_fitc(ctx, st, fit);
  }

  // ---- EP (R/ep.R:333-374) ----
  if (!approx->quad_x || !approx->quad_w || approx->quad_order < 1) {
    set_error("EP needs Gauss-Hermite nodes and weights");
    return GPL_ERR_BAD_ARG;
  }
  DevBuf qx, qw, flags, fmean_old, fvar_old, z_old, P_old;
  GPL_TRY(qx.alloc(ctx, sizeof(double) * approx->quad_order));
  GPL_TRY(qw.alloc(ctx, sizeof(double) * approx->quad_order));
  GPL_CUDA(cudaMemcpyAsync(qx.p, approx->quad_x, sizeof(double) * approx->quad_order,
                           cudaMemcpyHostToDevice, ctx->stream));
  GPL_CUDA(cudaMemcpyAsync(qw.p, approx->quad_w, sizeof(double) * approx->quad_order,
                           cudaMemcpyHostToDevice, ctx->stream));
  GPL_TRY(flags.alloc(ctx, sizeof(int) * 4));
  GPL_TRY(fmean_old.alloc(ctx, sizeof(double) * n));
  GPL_TRY(fvar_old.alloc(ctx, sizeof(double) * n));
  GPL_TRY(z_old.alloc(ctx, sizeof(double) * n));
  GPL_TRY(P_old.alloc(ctx, sizeof(double) * n));
  GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, fmean_old.d(), 0.0));
  GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, z_old.d(), 0.0));
  GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, P_old.d(), 0.0));
  GPL_TRY(vec_op(ctx, VOP_COPY, n, st.Kdiag.d(), nullptr, fvar_old.d()));
  double diff_old = INFINITY, damp = approx->damping, log_evidence = 0.0, diff_new = 0.0;
  const int maxiter = approx->maxiter < 1 ? 1 : approx->maxiter;
  int it = 0, quad_ok = 1;
  for (it = 1; it <= maxiter; ++it) {
    EpSiteArgs ea;
    ea.fmean = fmean_old.d(); ea.fvar = fvar_old.d(); ea.z_old = z_old.d(); ea.P_old = P_old.d();
    ea.y = st.y.d();
    ea.z_new = st.v[F_Z].d(); ea.V_new = st.v[F_V].d(); ea.P_new = st.v[F_P].d();
    ea.cav_mean = st.v[F_CM].d(); ea.cav_var = st.v[F_CV].d(); ea.logZ = st.v[F_LOGZ].d();
    ea.damping = damp;
    ea.quad_order = approx->quad_order;
    ea.quad_x = qx.d(); ea.quad_w = qw.d();
    GPL_TRY(lik_ep_sites(ctx, st.lik, n, ea, static_cast<int*>(flags.p)));
    int hflags = 0;
    GPL_CUDA(cudaMemcpyAsync(&hflags, flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    quad_ok = (hflags & 1) ? 0 : 1;
    if (hflags & 2) {
      out->log_evidence = -INFINITY;
      out->iters = it;
      out->quad_ok = quad_ok;
      set_error("EP (FITC): NaN pseudo-data");
      return GPL_ERR_NONFINITE;
    }
    if (hflags & 4) {
      set_error("EP (FITC): pseudo-variance <= 0; the reference's bad-site branch "
                "(R/ep.R:186-272) is not implemented");
      return GPL_ERR_NOT_LOGCONCAVE;
    }
    double logdetA = 0.0;
    GPL_TRY(fitc_core(ctx, st, &logdetA));
    // fvar = V D / S + (V / S)^2 |L^-1 k_zi|^2
    GPL_TRY(row_norms_via(ctx, st, st.A.d(), st.A_dinv.d(), st.v[F_W].d()));
    GPL_TRY(fvec(ctx, 2, n, st.v[F_V].d(), st.Dg.d(), st.S.d(), st.v[F_W].d(), st.v[F_FVAR].d()));
    GPL_TRY(reduce(ctx, RED_SUMLOG, n, st.S.d(), nullptr, 0));
    GPL_TRY(reduce(ctx, RED_DOT, n, st.v[F_Z].d(), st.v[F_ALPHA].d(), 1));
    GPL_TRY(ep_evidence_terms(ctx, n, st.v[F_LOGZ].d(), st.v[F_CV].d(), st.v[F_CM].d(), st.v[F_Z].d(),
                              st.v[F_V].d(), 2));
    GPL_TRY(reduce(ctx, RED_MAXABSDIFF, n, fmean_old.d(), st.v[F_FNEW].d(), 3));
    GPL_TRY(fetch_scalars(ctx, 4, sc));
    const double C_logdet = sc[0] - st.logdet_z + logdetA;
    log_evidence = -0.5 * C_logdet - 0.5 * sc[1] + sc[2];                  // R/ep.R:182-184
    diff_new = sc[3];
    if (diff_new > diff_old) {
      damp *= approx->damping;
      continue;
    }
    damp = damp > approx->damping_min ? damp : approx->damping_min;
    GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[F_FNEW].d(), nullptr, fmean_old.d()));
    GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[F_FVAR].d(), nullptr, fvar_old.d()));
    GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[F_Z].d(), nullptr, z_old.d()));
    GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[F_P].d(), nullptr, P_old.d()));
    diff_old = diff_new;
    if (diff_new < approx->tol) break;
  }
  if (it > maxiter) it = maxiter;
  out->iters = it;
  out->converged = !(maxiter > 1 && it == maxiter);
  out->quad_ok = quad_ok;
  if (maxiter > 1 && it == maxiter) log_evidence = -INFINITY;
  if (!quad_ok) log_evidence = -INFINITY;
  out->log_evidence = log_evidence;
  GPL_TRY(dl(h_fmean, st.v[F_FNEW].d()));
  GPL_TRY(dl(h_fvar, st.v[F_FVAR].d()));
  GPL_TRY(dl(h_z, st.v[F_Z].d()));
  GPL_TRY(dl(h_V, st.v[F_V].d()));
  GPL_TRY(dl(h_alpha, st.v[F_ALPHA].d()));
  GPL_TRY(dl(h_cm, st.v[F_CM].d()));
  GPL_TRY(dl(h_cv, st.v[F_CV].d()));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  return save_fitc(ctx, st, fit);
}

int pred_fitc_impl(Ctx* ctx, const Fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                   uint64_t term_mask, double* mean, double* var) {
  if (!fit || fit->method != 1 || nt < 0 || !mean) {
    set_error("gpl_pred_fitc: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  if (nt == 0) return GPL_OK;
  const int64_t m = fit->m;
  const int d = fit->d;
  HostPlan hp;
  GPL_TRY(make_plan(cf, d, term_mask, &hp));
  const int F = hp.plan.nfeat > 0 ? hp.plan.nfeat : 1;
  DevBuf dxt, Zz, Zt, Ktz, W, dmean, dvar, t1, t2;
  GPL_TRY(dxt.alloc(ctx, sizeof(double) * nt * d));
  GPL_CUDA(cudaMemcpyAsync(dxt.p, xt, sizeof(double) * nt * d, cudaMemcpyHostToDevice, ctx->stream));
  GPL_TRY(Zz.alloc(ctx, sizeof(double) * m * F));
  GPL_TRY(Zt.alloc(ctx, sizeof(double) * nt * F));
  GPL_TRY(prep_features(ctx, hp, fit->zind.d(), m, m, Zz.d(), m));
  GPL_TRY(prep_features(ctx, hp, dxt.d(), nt, nt, Zt.d(), nt));
  int64_t chunk = (int64_t)(2.0e9 / (8.0 * (double)m));
  chunk = chunk / 128 * 128;
  if (chunk < 128) chunk = 128;
  if (chunk > nt) chunk = nt;
  const int64_t ldk = round_up(chunk, 16);
  GPL_TRY(Ktz.alloc(ctx, sizeof(double) * ldk * m));
  GPL_TRY(W.alloc(ctx, sizeof(double) * ldk * m));
  GPL_TRY(dmean.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(dvar.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(t1.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(t2.alloc(ctx, sizeof(double) * nt));
  const double* beta = fit->sqrtV.d();
  for (int64_t r0 = 0; r0 < nt; r0 += chunk) {
    const int64_t c = (nt - r0) < chunk ? (nt - r0) : chunk;
    GPL_TRY(cf_pairwise(ctx, hp.plan, Zt.d() + r0, c, nt, Zz.d(), m, m, false, 0.0, Ktz.d(), ldk));
    GPL_TRY(gemv_n(ctx, c, m, Ktz.d(), ldk, beta, dmean.d() + r0));
    if (var) {
      GPL_TRY(cf_diag(ctx, hp.plan, Zt.d() + r0, nt, Zt.d() + r0, nt, c, 0.0, dvar.d() + r0));
      GPL_TRY(scaled_copy_rows(ctx, c, m, Ktz.d(), ldk, nullptr, W.d(), ldk));
      GPL_TRY(dtrsm_right_with(ctx, fit->Kz_dinv.d(), true, c, m, fit->Kz_chol.d(), fit->ldm, W.d(), ldk));
      GPL_TRY(rowsumsq(ctx, c, m, W.d(), ldk, t1.d() + r0));
      GPL_TRY(dtrsm_right_with(ctx, fit->Lm_dinv.d(), true, c, m, fit->Lm.d(), fit->ldm, Ktz.d(), ldk));
      GPL_TRY(rowsumsq(ctx, c, m, Ktz.d(), ldk, t2.d() + r0));
    }
  }
  GPL_CUDA(cudaMemcpyAsync(mean, dmean.p, sizeof(double) * nt, cudaMemcpyDeviceToHost, ctx->stream));
  if (var) {
    // var = (k_tt - |Lz^-1 k_zt|^2) + |L^-1 k_zt|^2
    GPL_TRY(vec_op(ctx, VOP_SUB, nt, dvar.d(), t1.d(), dvar.d()));
    GPL_TRY(vec_op(ctx, VOP_ADD, nt, dvar.d(), t2.d(), dvar.d()));
    GPL_CUDA(cudaMemcpyAsync(var, dvar.p, sizeof(double) * nt, cudaMemcpyDeviceToHost, ctx->stream));
  }
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPL_OK;
}

}  // namespace gpl

using namespace gpl;

extern "C" {

int gpl_fit_fitc(gpl_ctx* h, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                 const double* x, int64_t n, int32_t d, const double* zind, int64_t m,
                 const double* y, const double* trials, const double* offset, double jitter,
                 gpl_fit_out* out, double* fmean, double* z, double* V, double* alpha, double* fvar,
                 double* cavity_mean, double* cavity_var, gpl_fit** fit) {
  Fit* f = nullptr;
  if (fit) *fit = nullptr;
  int rc = fit_fitc_impl(reinterpret_cast<Ctx*>(h), cf, lik, approx, x, n, d, zind, m, y, trials,
                         offset, jitter, out, fmean, z, V, alpha, fvar, cavity_mean, cavity_var,
                         fit ? &f : nullptr);
  if (fit) *fit = reinterpret_cast<gpl_fit*>(f);
  if (rc != GPL_OK) cudaStreamSynchronize(reinterpret_cast<Ctx*>(h)->stream);
  return rc;
}

int gpl_pred_fitc(gpl_ctx* h, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt,
                  int64_t nt, uint64_t term_mask, double* mean, double* var) {
  return pred_fitc_impl(reinterpret_cast<Ctx*>(h), reinterpret_cast<const Fit*>(fit), cf, xt, nt,
                        term_mask, mean, var);
}

}  // extern "C"
