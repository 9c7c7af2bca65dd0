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
#include <cmath>
#include <vector>

namespace gpl {

// ---- nb x nb algebra of the non-log-concave branch, shared by method_full and method_fitc ------
int bad_small_alloc(Ctx* ctx, BadSmall& sm, int64_t nb, const double* V_b) {
  sm.nb = nb;
  sm.ldp = round_up(nb, 16);
  GPL_TRY(sm.invVb.alloc(ctx, sizeof(double) * nb));
  for (int i = 0; i < 8; ++i) GPL_TRY(sm.b[i].alloc(ctx, sizeof(double) * nb));
  GPL_TRY(vec_op(ctx, VOP_INV, nb, V_b, nullptr, sm.invVb.d()));
  const size_t pp = sizeof(double) * sm.ldp * nb, pd = sizeof(double) * ((nb + NB - 1) / NB) * NB * NB;
  GPL_TRY(sm.P.alloc(ctx, pp));
  GPL_TRY(sm.Lk.alloc(ctx, pp));
  GPL_TRY(sm.T.alloc(ctx, pp));
  GPL_TRY(sm.A.alloc(ctx, pp));
  GPL_TRY(sm.Pd.alloc(ctx, pd));
  GPL_TRY(sm.Ad.alloc(ctx, pd));
  return GPL_OK;
}

// P holds pred_cov (lower triangle) on entry.  Lk = chol(pred_cov) -- failure means the posterior
// of the whole data is improper and is an error (R/laplace.R:69-71); A = I + Lk^T V_b^-1 Lk and its
// factor -- failure sets a_fail (R/laplace.R:73-80, R/ep.R:90-97).
int bad_small_factor(Ctx* ctx, BadSmall& sm) {
  const int64_t nb = sm.nb;
  int32_t info = 0;
  GPL_TRY(dpotrf_with(ctx, sm.Pd.d(), nb, sm.P.d(), sm.ldp, &info, nullptr));
  GPL_TRY(tril_scale(ctx, nb, sm.P.d(), sm.ldp, nullptr, sm.Lk.d(), sm.ldp));
  GPL_TRY(tril_scale(ctx, nb, sm.P.d(), sm.ldp, sm.invVb.d(), sm.T.d(), sm.ldp));
  GPL_TRY(dgemm(ctx, true, false, nb, nb, nb, 1.0, sm.Lk.d(), sm.ldp, sm.T.d(), sm.ldp, 0.0, sm.A.d(),
                sm.ldp, false));
  GPL_TRY(scale_block(ctx, nb, nb, sm.A.d(), sm.ldp, nullptr, nullptr, 1.0, 1.0, 0));
  sm.a_fail = false;
  const int rc = dpotrf_with(ctx, sm.Ad.d(), nb, sm.A.d(), sm.ldp, &info, &sm.logdetA);
  if (rc == GPL_ERR_NOT_PD || (rc == GPL_OK && !std::isfinite(sm.logdetA))) {
    sm.a_fail = true;
    return GPL_OK;
  }
  return rc;
}

// out = A^-1 c  (A = Ac Ac^T); c is destroyed; tmp, out distinct nb-vectors
static int bad_small_solve_A(Ctx* ctx, BadSmall& sm, double* c, double* tmp, double* out) {
  GPL_TRY(trsv_lower(ctx, sm.Ad.d(), false, sm.nb, sm.A.d(), sm.ldp, c, tmp));
  GPL_TRY(trsv_lower(ctx, sm.Ad.d(), true, sm.nb, sm.A.d(), sm.ldp, tmp, out));
  return GPL_OK;
}

// out = S^-1 v, S = pred_cov + V_bad (Woodbury through A); v is kept
int bad_small_solve_S(Ctx* ctx, BadSmall& sm, const double* v, double* out) {
  const int64_t nb = sm.nb;
  double *t = sm.b[4].d(), *u = sm.b[5].d(), *u2 = sm.b[6].d(), *w = sm.b[7].d();
  GPL_TRY(vec_op(ctx, VOP_MUL, nb, v, sm.invVb.d(), t));
  GPL_TRY(gemv_t(ctx, nb, nb, sm.Lk.d(), sm.ldp, t, u));
  GPL_TRY(bad_small_solve_A(ctx, sm, u, u2, w));
  GPL_TRY(gemv_n(ctx, nb, nb, sm.Lk.d(), sm.ldp, w, u));
  GPL_TRY(vec_op(ctx, VOP_MUL, nb, u, sm.invVb.d(), u2));
  GPL_TRY(vec_op(ctx, VOP_SUB, nb, t, u2, out));
  return GPL_OK;
}

// Laplace terms of the bad sites (R/laplace.R:81-99): f2hat = Lk w, w = A^-1 (Lk^T (z_b / V_b) +
// Lk^-1 pred_mean); |Lk^-1 (f2hat - pred_mean)|^2 -> slot_df2, loglik(f2hat, y_b) -> slot_ll
int bad_small_laplace(Ctx* ctx, BadSmall& sm, const double* pm, const double* z_b, const double* y_b,
                      const LikDev& lik_b, int slot_df2, int slot_ll) {
  const int64_t nb = sm.nb;
  double *c1 = sm.b[0].d(), *c2 = sm.b[1].d(), *c = sm.b[2].d(), *w = sm.b[3].d();
  double *t = sm.b[4].d(), *u = sm.b[5].d();
  GPL_TRY(vec_op(ctx, VOP_MUL, nb, z_b, sm.invVb.d(), t));
  GPL_TRY(gemv_t(ctx, nb, nb, sm.Lk.d(), sm.ldp, t, c1));
  GPL_TRY(vec_op(ctx, VOP_COPY, nb, pm, nullptr, t));
  GPL_TRY(trsv_lower(ctx, sm.Pd.d(), false, nb, sm.P.d(), sm.ldp, t, c2));
  GPL_TRY(vec_op(ctx, VOP_ADD, nb, c1, c2, c));
  GPL_TRY(bad_small_solve_A(ctx, sm, c, u, w));
  GPL_TRY(gemv_n(ctx, nb, nb, sm.Lk.d(), sm.ldp, w, t));              // f2hat
  GPL_TRY(vec_op(ctx, VOP_SUB, nb, w, c2, u));                        // Lk^-1 (f2hat - pred_mean)
  GPL_TRY(reduce(ctx, RED_SUMSQ, nb, u, nullptr, slot_df2));
  GPL_TRY(lik_loglik_sum(ctx, lik_b, nb, t, y_b, slot_ll));
  return GPL_OK;
}

// EP: (pred_mean - z_b)^T S^-1 (pred_mean - z_b) -> slot (R/ep.R:107-110)
int bad_small_ep(Ctx* ctx, BadSmall& sm, const double* pm, const double* z_b, int slot_df2) {
  double *dz = sm.b[0].d(), *sdz = sm.b[1].d();
  GPL_TRY(vec_op(ctx, VOP_SUB, sm.nb, pm, z_b, dz));
  GPL_TRY(bad_small_solve_S(ctx, sm, dz, sdz));
  GPL_TRY(reduce(ctx, RED_DOT, sm.nb, dz, sdz, slot_df2));
  return GPL_OK;
}

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
  if (i >= n) return;
  for (int64_t j = blockIdx.y; j < n; j += gridDim.y)
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
  double jitter = 0.0;
  bool K_permuted = false;   // K currently holds the bad-site ordering (rebuild before normal use)
};

enum { V_F = 0, V_Z, V_V, V_S, V_SQRTV, V_ALPHA, V_FNEW, V_T1, V_T2, V_DIAG, V_FVAR, V_P, V_ZOLD,
       V_CM, V_CV, V_LOGZ };

int full_setup(Ctx* ctx, FullState& st, const gpl_cf_model* cf, const gpl_lik* lik, const void* x,
               int64_t n, int d, const void* y, const double* trials, const double* offset,
               bool on_device, double jitter, bool want_aux) {
  st.n = n;
  st.d = d;
  st.jitter = jitter;
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

// ---------------------------------------------------------------------------------------------
// Non-log-concave sites (some pseudo-variances <= 0): R/laplace.R:40-117, R/ep.R:50-138.
// The reference evaluates p(y) = p(y_good) p(y_bad | y_good) and then refits all sites with a
// pivoted LU of K + diag(V).  Here the sites are permuted good-first (K is simply rebuilt from the
// permuted features: 2 ms at n = 32768), the good block is factored by Cholesky as usual, and the
// bad block is eliminated through its nb x nb Schur complement S = pred_cov + V_bad, whose inverse
// follows from the two small Cholesky factors the evidence needs anyway
// (S^-1 = V_b^-1 - V_b^-1 Lk A^-1 Lk^T V_b^-1, pred_cov = Lk Lk^T, A = I + Lk^T V_b^-1 Lk).
// Same alpha / fmean / fvar / predictions as the LU (block elimination of the same system) at a
// third of its flops and without a pivoted panel kernel.
struct BadSite {
  int64_t n = 0, ng = 0, nb = 0, ldw = 0;
  bool active = false;     // the last iteration took this branch
  BadSmall sm;             // nb x nb algebra (shared with method_fitc)
  std::vector<int32_t> perm_h;
  DevBuf perm, Zp, xp;
  DevBuf yp, trp, offp, zp, Vp, sp, sqp, alpha_p;
  DevBuf a1, f1, pm, ab;
  DevBuf W;
  DevBuf g[4];             // n-sized scratch
  LikDev lik_g, lik_b;
};

int copy_block(Ctx* ctx, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst,
               int64_t ldd) {
  if (rows <= 0 || cols <= 0) return GPL_OK;
  GPL_CUDA(cudaMemcpy2DAsync(dst, sizeof(double) * ldd, src, sizeof(double) * lds,
                             sizeof(double) * rows, cols, cudaMemcpyDeviceToDevice, ctx->stream));
  return GPL_OK;
}

// permutation, permuted K and site vectors, small buffers
int bad_prepare(Ctx* ctx, FullState& st, BadSite& bs) {
  const int64_t n = st.n;
  std::vector<double> hV(n);
  GPL_CUDA(cudaMemcpyAsync(hV.data(), st.v[V_V].d(), sizeof(double) * n, cudaMemcpyDeviceToHost,
                           ctx->stream));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  bs.n = n;
  bs.perm_h.resize(n);
  int64_t ng = 0;
  for (int64_t i = 0; i < n; ++i)
    if (hV[i] > 0.0) bs.perm_h[ng++] = (int32_t)i;
  int64_t k = ng;
  for (int64_t i = 0; i < n; ++i)
    if (!(hV[i] > 0.0)) bs.perm_h[k++] = (int32_t)i;
  bs.ng = ng;
  bs.nb = n - ng;
  if (ng == 0) {
    set_error("all pseudo-variances are <= 0: the posterior is not proper");
    return GPL_ERR_NOT_PD;
  }
  const int64_t nb = bs.nb;
  bs.ldw = round_up(nb, 16);
  GPL_TRY(bs.perm.alloc(ctx, sizeof(int32_t) * n));
  GPL_CUDA(cudaMemcpyAsync(bs.perm.p, bs.perm_h.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice,
                           ctx->stream));
  const int32_t* perm = static_cast<const int32_t*>(bs.perm.p);
  const int F = st.hp.plan.nfeat > 0 ? st.hp.plan.nfeat : 1;
  GPL_TRY(bs.Zp.alloc(ctx, sizeof(double) * n * F));
  GPL_TRY(bs.xp.alloc(ctx, sizeof(double) * n * st.d));
  GPL_TRY(gather_rows(ctx, n, st.hp.plan.nfeat, st.Z.d(), n, perm, bs.Zp.d(), n));
  GPL_TRY(gather_rows(ctx, n, st.d, st.x.d(), n, perm, bs.xp.d(), n));
  GPL_TRY(cf_pairwise(ctx, st.hp.plan, bs.Zp.d(), n, n, nullptr, n, n, true, st.jitter, st.K.d(),
                      st.ld));
  st.K_permuted = true;
  DevBuf* nv[] = {&bs.yp, &bs.zp, &bs.Vp, &bs.sp, &bs.sqp, &bs.alpha_p, &bs.a1, &bs.f1,
                  &bs.g[0], &bs.g[1], &bs.g[2], &bs.g[3]};
  for (DevBuf* b : nv) GPL_TRY(b->alloc(ctx, sizeof(double) * n));
  GPL_TRY(gather_vec(ctx, n, st.y.d(), perm, bs.yp.d()));
  GPL_TRY(gather_vec(ctx, n, st.v[V_Z].d(), perm, bs.zp.d()));
  GPL_TRY(gather_vec(ctx, n, st.v[V_V].d(), perm, bs.Vp.d()));
  bs.lik_g = st.lik;
  bs.lik_b = st.lik;
  if (st.lik.trials) {
    GPL_TRY(bs.trp.alloc(ctx, sizeof(double) * n));
    GPL_TRY(gather_vec(ctx, n, st.lik.trials, perm, bs.trp.d()));
    bs.lik_g.trials = bs.trp.d();
    bs.lik_b.trials = bs.trp.d() + ng;
  }
  if (st.lik.offset) {
    GPL_TRY(bs.offp.alloc(ctx, sizeof(double) * n));
    GPL_TRY(gather_vec(ctx, n, st.lik.offset, perm, bs.offp.d()));
    bs.lik_g.offset = bs.offp.d();
    bs.lik_b.offset = bs.offp.d() + ng;
  }
  GPL_TRY(vec_op(ctx, VOP_RSQRT_INV, ng, bs.Vp.d(), nullptr, bs.sp.d()));
  GPL_TRY(vec_op(ctx, VOP_SQRT, ng, bs.Vp.d(), nullptr, bs.sqp.d()));
  GPL_TRY(bs.pm.alloc(ctx, sizeof(double) * nb));
  GPL_TRY(bs.ab.alloc(ctx, sizeof(double) * nb));
  GPL_TRY(bad_small_alloc(ctx, bs.sm, nb, bs.Vp.d() + ng));
  GPL_TRY(bs.W.alloc(ctx, sizeof(double) * bs.ldw * ng));
  return GPL_OK;
}

// fit of the good sites (R/laplace.R:52-62 -> :27-30) and the predictive "prior" of the bad ones
// given them (:66-80): alpha1, pred_mean, W = K_bg C_chol^-T, Lk, A_chol
int bad_factor(Ctx* ctx, FullState& st, BadSite& bs) {
  const int64_t ng = bs.ng, nb = bs.nb, ld = st.ld;
  const double* K = st.K.d();
  GPL_TRY(build_B(ctx, ng, K, ld, bs.sp.d(), st.B.d(), ld));
  int32_t info = 0;
  GPL_TRY(dpotrf_with(ctx, st.dinv.d(), ng, st.B.d(), ld, &info, nullptr));
  GPL_TRY(solve_alpha(ctx, ng, st.B.d(), ld, st.dinv.d(), bs.sqp.d(), bs.zp.d(), bs.a1.d(),
                      bs.g[0].d(), bs.g[1].d()));
  GPL_TRY(gemv_t(ctx, ng, ng, K, ld, bs.a1.d(), bs.f1.d()));
  GPL_TRY(diag_extract(ctx, ng, st.B.d(), ld, st.v[V_DIAG].d()));
  const double* Kbg = K + ng;                          // rows ng.., columns 0..ng-1
  GPL_TRY(gemv_n(ctx, nb, ng, Kbg, ld, bs.a1.d(), bs.pm.d()));
  GPL_TRY(copy_block(ctx, nb, ng, Kbg, ld, bs.W.d(), bs.ldw));
  GPL_TRY(scale_cols(ctx, nb, ng, bs.W.d(), bs.ldw, bs.sp.d()));
  GPL_TRY(dtrsm_right_with(ctx, st.dinv.d(), true, nb, ng, st.B.d(), ld, bs.W.d(), bs.ldw));
  GPL_TRY(copy_block(ctx, nb, nb, K + ng + ng * ld, ld, bs.sm.P.d(), bs.sm.ldp));
  GPL_TRY(dgemm(ctx, false, true, nb, nb, ng, -1.0, bs.W.d(), bs.ldw, bs.W.d(), bs.ldw, 1.0,
                bs.sm.P.d(), bs.sm.ldp, true));
  return bad_small_factor(ctx, bs.sm);
}

int bad_solve_S(Ctx* ctx, BadSite& bs, const double* v, double* out) {
  return bad_small_solve_S(ctx, bs.sm, v, out);
}

// out_i = r_i S^-1 r_i^T for the rows of R (rows x nb, ld ldr); R is kept
int bad_quad(Ctx* ctx, const Fit* f, int64_t rows, const double* R, int64_t ldr, double* out,
             double* tmp_rows) {
  const int64_t nb = f->nb;
  if (rows <= 0) return GPL_OK;
  DevBuf G, G2;
  GPL_TRY(G.alloc(ctx, sizeof(double) * ldr * nb));
  GPL_TRY(G2.alloc(ctx, sizeof(double) * ldr * nb));
  GPL_TRY(copy_block(ctx, rows, nb, R, ldr, G.d(), ldr));
  GPL_TRY(scale_cols(ctx, rows, nb, G.d(), ldr, f->invVb.d()));
  GPL_TRY(dgemm(ctx, false, false, rows, nb, nb, 1.0, G.d(), ldr, f->Lk.d(), f->ldp, 0.0, G2.d(),
                ldr, false));
  GPL_TRY(dtrsm_right_with(ctx, f->Ac_dinv.d(), true, rows, nb, f->Ac.d(), f->ldp, G2.d(), ldr));
  GPL_TRY(rowsumsq(ctx, rows, nb, G2.d(), ldr, tmp_rows));
  GPL_TRY(rowsumsq_w(ctx, rows, nb, R, ldr, f->invVb.d(), out));
  GPL_TRY(vec_op(ctx, VOP_SUB, rows, out, tmp_rows, out));
  return GPL_OK;
}

// view of the small factors as a (non-owning) Fit so that fit and predict share bad_quad
struct FitView {
  Fit f;
  explicit FitView(BadSite& bs) {
    f.nb = bs.nb; f.ng = bs.ng; f.ldp = bs.sm.ldp; f.ldw = bs.ldw;
    f.Lk.p = bs.sm.Lk.p; f.Ac.p = bs.sm.A.p; f.Ac_dinv.p = bs.sm.Ad.p; f.invVb.p = bs.sm.invVb.p;
  }
  ~FitView() { f.Lk.p = f.Ac.p = f.Ac_dinv.p = f.invVb.p = nullptr; }
};

// alpha of the full system (K + diag(V)) alpha = z by block elimination, f = K alpha; results are
// scattered back to the original site order into st.v[V_ALPHA], st.v[V_FNEW]
int bad_final_solve(Ctx* ctx, FullState& st, BadSite& bs) {
  const int64_t n = st.n, ng = bs.ng, nb = bs.nb;
  const int32_t* perm = static_cast<const int32_t*>(bs.perm.p);
  double* rhs = bs.sm.b[0].d();
  GPL_TRY(vec_op(ctx, VOP_SUB, nb, bs.zp.d() + ng, bs.pm.d(), rhs));
  GPL_TRY(bad_solve_S(ctx, bs, rhs, bs.ab.d()));
  GPL_TRY(gemv_t(ctx, nb, ng, bs.W.d(), bs.ldw, bs.ab.d(), bs.g[0].d()));
  GPL_TRY(trsv_lower(ctx, st.dinv.d(), true, ng, st.B.d(), st.ld, bs.g[0].d(), bs.g[1].d()));
  GPL_TRY(vec_op(ctx, VOP_MUL, ng, bs.g[1].d(), bs.sp.d(), bs.g[0].d()));
  GPL_TRY(vec_op(ctx, VOP_SUB, ng, bs.a1.d(), bs.g[0].d(), bs.alpha_p.d()));
  GPL_TRY(vec_op(ctx, VOP_COPY, nb, bs.ab.d(), nullptr, bs.alpha_p.d() + ng));
  GPL_TRY(gemv_t(ctx, n, n, st.K.d(), st.ld, bs.alpha_p.d(), bs.g[2].d()));
  GPL_TRY(scatter_vec(ctx, n, bs.alpha_p.d(), perm, st.v[V_ALPHA].d()));
  GPL_TRY(scatter_vec(ctx, n, bs.g[2].d(), perm, st.v[V_FNEW].d()));
  return GPL_OK;
}

// one Laplace step with bad sites; *log_evidence = -Inf and bs.sm.a_fail when the posterior is improper
int bad_laplace_iter(Ctx* ctx, FullState& st, BadSite& bs, double* log_evidence) {
  double sc[8];
  GPL_TRY(bad_prepare(ctx, st, bs));
  GPL_TRY(bad_factor(ctx, st, bs));
  bs.active = true;
  if (bs.sm.a_fail) {
    *log_evidence = -INFINITY;
    return GPL_OK;
  }
  const int64_t ng = bs.ng, nb = bs.nb;
  // evidence of the good sites (R/laplace.R:33-38 on the subset)
  GPL_TRY(reduce(ctx, RED_SUMLOG_PROD, ng, bs.sqp.d(), st.v[V_DIAG].d(), 0));
  GPL_TRY(reduce(ctx, RED_SUMLOG, ng, bs.Vp.d(), nullptr, 1));
  GPL_TRY(reduce(ctx, RED_DOT, ng, bs.f1.d(), bs.a1.d(), 2));
  GPL_TRY(lik_loglik_sum(ctx, bs.lik_g, ng, bs.f1.d(), bs.yp.d(), 3));
  GPL_TRY(bad_small_laplace(ctx, bs.sm, bs.pm.d(), bs.zp.d() + ng, bs.yp.d() + ng, bs.lik_b, 4, 5));
  GPL_TRY(fetch_scalars(ctx, 6, sc));
  const double le1 = sc[3] - 0.5 * sc[2] - 0.5 * (2.0 * sc[0] - sc[1]);
  const double le2 = sc[5] - 0.5 * sc[4] + 0.5 * bs.sm.logdetA;          // R/laplace.R:100
  *log_evidence = le1 + le2;
  return bad_final_solve(ctx, st, bs);
}

// one EP sweep with bad sites (R/ep.R:50-138): evidence, alpha, fmean and the marginal variances
// fvar = V - V^2 diag(C^-1) from the block inverse.  X: n x n scratch (ld st.ld).
int bad_ep_iter(Ctx* ctx, FullState& st, BadSite& bs, double* X, double* log_evidence) {
  double sc[8];
  GPL_TRY(bad_prepare(ctx, st, bs));
  GPL_TRY(bad_factor(ctx, st, bs));
  bs.active = true;
  if (bs.sm.a_fail) {
    *log_evidence = -INFINITY;
    return GPL_OK;
  }
  const int64_t n = st.n, ng = bs.ng, nb = bs.nb, ld = st.ld;
  const int32_t* perm = static_cast<const int32_t*>(bs.perm.p);
  DevBuf logZ, vc, mc;
  GPL_TRY(logZ.alloc(ctx, sizeof(double) * n));
  GPL_TRY(vc.alloc(ctx, sizeof(double) * n));
  GPL_TRY(mc.alloc(ctx, sizeof(double) * n));
  GPL_TRY(gather_vec(ctx, n, st.v[V_LOGZ].d(), perm, logZ.d()));
  GPL_TRY(gather_vec(ctx, n, st.v[V_CV].d(), perm, vc.d()));
  GPL_TRY(gather_vec(ctx, n, st.v[V_CM].d(), perm, mc.d()));
  GPL_TRY(reduce(ctx, RED_SUMLOG_PROD, ng, bs.sqp.d(), st.v[V_DIAG].d(), 0));
  GPL_TRY(reduce(ctx, RED_DOT, ng, bs.zp.d(), bs.a1.d(), 1));
  GPL_TRY(ep_evidence_terms(ctx, ng, logZ.d(), vc.d(), mc.d(), bs.zp.d(), bs.Vp.d(), 2));
  GPL_TRY(bad_small_ep(ctx, bs.sm, bs.pm.d(), bs.zp.d() + ng, 3));     // R/ep.R:107-110
  GPL_TRY(ep_evidence_terms_bad(ctx, nb, logZ.d() + ng, vc.d() + ng, mc.d() + ng, bs.zp.d() + ng,
                                bs.Vp.d() + ng, 4));
  GPL_TRY(fetch_scalars(ctx, 5, sc));
  const double le1 = -0.5 * (2.0 * sc[0]) - 0.5 * sc[1] + sc[2];
  const double le2 = -0.5 * sc[3] + 0.5 * bs.sm.logdetA + sc[4];           // R/ep.R:112-114
  *log_evidence = le1 + le2;
  GPL_TRY(bad_final_solve(ctx, st, bs));
  // marginal variances.  diag(C^-1)_good = s^2 colsumsq(L^-1) + rows of U S^-1 U^T with
  // U = C_gg^-1 K_gb = S L^-T W^T;  diag(C^-1)_bad = diag(S^-1)
  FitView view(bs);
  GPL_TRY(copy_lower(ctx, ng, st.B.d(), ld, X, ld));
  GPL_TRY(dtrtri_with(ctx, st.dinv.d(), ng, X, ld));
  GPL_TRY(lower_colsumsq(ctx, ng, X, ld, bs.g[3].d()));
  DevBuf Ut, U;
  const int64_t ldu = round_up(ng, 16);
  GPL_TRY(Ut.alloc(ctx, sizeof(double) * bs.ldw * ng));
  GPL_TRY(U.alloc(ctx, sizeof(double) * ldu * nb));
  GPL_TRY(copy_block(ctx, nb, ng, bs.W.d(), bs.ldw, Ut.d(), bs.ldw));
  GPL_TRY(dtrsm_right_with(ctx, st.dinv.d(), false, nb, ng, st.B.d(), ld, Ut.d(), bs.ldw));
  GPL_TRY(scale_cols(ctx, nb, ng, Ut.d(), bs.ldw, bs.sp.d()));
  GPL_TRY(transpose(ctx, nb, ng, Ut.d(), bs.ldw, U.d(), ldu));
  GPL_TRY(bad_quad(ctx, &view.f, ng, U.d(), ldu, bs.g[0].d(), bs.g[1].d()));
  double* fv = bs.g[2].d();
  GPL_TRY(vec_op(ctx, VOP_MUL, ng, bs.Vp.d(), bs.g[3].d(), bs.g[1].d()));     // V T1
  GPL_TRY(vec_op(ctx, VOP_SUB, ng, bs.Vp.d(), bs.g[1].d(), fv));              // V - V T1
  GPL_TRY(vec_op(ctx, VOP_MUL, ng, bs.Vp.d(), bs.Vp.d(), bs.g[1].d()));       // V^2
  GPL_TRY(vec_op(ctx, VOP_MUL, ng, bs.g[1].d(), bs.g[0].d(), bs.g[1].d()));   // V^2 q
  GPL_TRY(vec_op(ctx, VOP_SUB, ng, fv, bs.g[1].d(), fv));
  GPL_CUDA(cudaMemsetAsync(bs.sm.T.d(), 0, sizeof(double) * bs.sm.ldp * nb, ctx->stream));
  GPL_TRY(scale_block(ctx, nb, nb, bs.sm.T.d(), bs.sm.ldp, nullptr, nullptr, 1.0, 1.0, 0));   // I
  GPL_TRY(bad_quad(ctx, &view.f, nb, bs.sm.T.d(), bs.sm.ldp, bs.sm.b[0].d(), bs.sm.b[1].d()));
  GPL_TRY(vec_op(ctx, VOP_MUL, nb, bs.Vp.d() + ng, bs.Vp.d() + ng, bs.sm.b[1].d()));
  GPL_TRY(vec_op(ctx, VOP_MUL, nb, bs.sm.b[1].d(), bs.sm.b[0].d(), bs.sm.b[1].d()));
  GPL_TRY(vec_op(ctx, VOP_SUB, nb, bs.Vp.d() + ng, bs.sm.b[1].d(), fv + ng));
  GPL_TRY(scatter_vec(ctx, n, fv, perm, st.v[V_FVAR].d()));
  return GPL_OK;
}

int save_fit(Ctx* ctx, FullState& st, Fit** out, BadSite* bs = nullptr) {
  if (!out) return GPL_OK;
  Fit* f = new Fit();
  auto take = [&](DevBuf& dst, DevBuf& src) {
    dst.ctx = ctx;
    dst.p = src.p;
    src.p = nullptr;
  };
  f->method = 0;
  f->n = st.n;
  f->d = st.d;
  f->ld = st.ld;
  f->ng = st.n;
  take(f->x, st.x);
  take(f->Lb, st.B);
  take(f->dinv, st.dinv);
  take(f->alpha, st.v[V_ALPHA]);
  if (bs && bs->active) {          // block factor of the non-log-concave fit
    f->ng = bs->ng;
    f->nb = bs->nb;
    f->ldw = bs->ldw;
    f->ldp = bs->sm.ldp;
    take(f->sqrtV, bs->sqp);
    take(f->perm, bs->perm);
    take(f->xp, bs->xp);
    take(f->alpha_p, bs->alpha_p);
    take(f->W, bs->W);
    take(f->Lk, bs->sm.Lk);
    take(f->Lk_dinv, bs->sm.Pd);
    take(f->Ac, bs->sm.A);
    take(f->Ac_dinv, bs->sm.Ad);
    take(f->invVb, bs->sm.invVb);
  } else {
    take(f->sqrtV, st.v[V_SQRTV]);
  }
  *out = f;
  return GPL_OK;
}

// Newton iterations of laplace.method_full (R/laplace.R:274-297) on a set-up state.  Leaves f in
// v[V_F], the pseudo-data z, V of the last step, alpha, the factor of B in st.B (or the block factor
// in bs when the last step had non-log-concave sites).
int laplace_loop(Ctx* ctx, FullState& st, BadSite& bs, const gpl_lik* lik, const gpl_approx* approx,
               gpl_fit_out* out) {
  const int64_t n = st.n;
  double sc[8];
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
    if (!(sc[0] > 0.0)) {                       // non-log-concave sites: R/laplace.R:40-117
      GPL_TRY(bad_laplace_iter(ctx, st, bs, &log_evidence));
      if (bs.sm.a_fail) {                          // fit = list(fmean = fhat_old, log_evidence = -Inf)
        out->log_evidence = -INFINITY;
        out->iters = it;
        out->converged = 0;
        set_error("Laplace: the posterior with non-log-concave sites is not proper");
        return GPL_ERR_NONFINITE;
      }
      GPL_TRY(reduce(ctx, RED_MAXABSDIFF, n, st.v[V_F].d(), st.v[V_FNEW].d(), 0));
      GPL_TRY(fetch_scalars(ctx, 1, sc));
      diff = sc[0];
      std::swap(st.v[V_F].p, st.v[V_FNEW].p);
      if (diff < approx->tol) break;
      continue;
    }
    bs.active = false;
    if (st.K_permuted) {                        // a previous step left K in bad-site order
      GPL_TRY(cf_pairwise(ctx, st.hp.plan, st.Z.d(), n, n, nullptr, n, n, true, st.jitter,
                          st.K.d(), st.ld));
      st.K_permuted = false;
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
  out->accepted = it;
  out->converged = !(maxiter > 1 && it == maxiter);
  out->n_bad = bs.active ? (int32_t)bs.nb : 0;
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
  BadSite bs;

  if (approx->kind == GPL_APPROX_LAPLACE) {
    GPL_TRY(laplace_loop(ctx, st, bs, lik, approx, out));
    GPL_TRY(download(ctx, h_fmean, st.v[V_F].d(), n));
    GPL_TRY(download(ctx, h_z, st.v[V_Z].d(), n));
    GPL_TRY(download(ctx, h_V, st.v[V_V].d(), n));
    GPL_TRY(download(ctx, h_alpha, st.v[V_ALPHA].d(), n));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    return save_fit(ctx, st, fit, &bs);
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
  int it = 0, quad_ok = 1, accepted = 0;
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
    if (hflags & 4) {                                                      // R/ep.R:50-138
      GPL_TRY(bad_ep_iter(ctx, st, bs, X.d(), &log_evidence));
      if (bs.sm.a_fail) {
        // the reference returns fmean = Inf here and lets its loop run on NaNs; report it instead
        out->log_evidence = -INFINITY;
        out->iters = it;
        out->converged = 0;
        out->quad_ok = quad_ok;
        set_error("EP: the posterior with non-log-concave sites is not proper");
        return GPL_ERR_NONFINITE;
      }
      GPL_TRY(reduce(ctx, RED_MAXABSDIFF, n, fmean_old.d(), st.v[V_FNEW].d(), 0));
      GPL_TRY(fetch_scalars(ctx, 1, sc));
      diff_new = sc[0];
    } else {
      bs.active = false;
      if (st.K_permuted) {
        GPL_TRY(cf_pairwise(ctx, st.hp.plan, st.Z.d(), n, n, nullptr, n, n, true, st.jitter,
                            st.K.d(), st.ld));
        st.K_permuted = false;
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
    }
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
    ++accepted;
    if (diff_new < approx->tol) break;
  }
  if (it > maxiter) it = maxiter;
  out->accepted = accepted;
  out->damping_final = damp;
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
  out->n_bad = bs.active ? (int32_t)bs.nb : 0;
  return save_fit(ctx, st, fit, &bs);
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

// Energy and its gradient for approx_laplace with a non-Gaussian likelihood, method_full (new
// capability, as gpl_energy_grad is for lik_gaussian; Rasmussen & Williams Alg. 5.1 on the reference's
// objective).  At the mode, with a = alpha, R = (K + W^-1)^-1 = S B^-1 S, the posterior variances
// pv = V (1 - diag(B^-1)) and d3 = d/df of the second derivative of log p(y|f),
//   s2 = pv d3 / 2,  u = (I - R K) s2,
//   dE/dlog(theta_j) = -1/2 sum_ik [(a+u)_i (a+u)_k - u_i u_k - R_ik] dK_ik/dlog(theta_j)
// (the u terms carry the dependence of the mode on theta).  d3 is a central difference of the
// reference's own finite-difference curvature (get_loglik_d2, R/lik.R:449-454), so no likelihood
// needs new derivative code.  Only covariance-function parameters; the Newton loop should be run
// to a tight tolerance (the implicit-function argument assumes the mode).
int energy_grad_laplace_impl(Ctx* ctx, const gpl_cf_model* cf, const gpl_lik* lik,
                             const gpl_approx* approx, const double* x, int64_t n, int d,
                             const double* y, const double* trials, const double* offset,
                             double jitter, int nparam, const int32_t* param_factor,
                             const int32_t* param_slot, double* energy, double* grad,
                             gpl_fit_out* out) {
  if (n <= 0 || d <= 0 || !x || !y || !energy || !out || !approx || nparam < 0 ||
      nparam > MAX_GRAD_PARAMS || (nparam > 0 && !grad)) {
    set_error("gpl_energy_grad_laplace: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  if (approx->kind != GPL_APPROX_LAPLACE) {
    set_error("gpl_energy_grad_laplace: approx_laplace only");
    return GPL_ERR_BAD_ARG;
  }
  bool want_aux = false;
  for (int j = 0; j < nparam; ++j) {
    if (param_factor[j] < 0) {
      set_error("gpl_energy_grad_laplace: only covariance-function parameters have a gradient here "
                "(fix the likelihood's parameters)");
      return GPL_ERR_BAD_ARG;
    }
    if (param_slot[j] == 3) want_aux = true;
  }
  FullState st;
  GPL_TRY(full_setup(ctx, st, cf, lik, x, n, d, y, trials, offset, false, jitter, want_aux));
  memset(out, 0, sizeof(*out));
  out->quad_ok = 1;
  BadSite bs;
  GPL_TRY(laplace_loop(ctx, st, bs, lik, approx, out));
  *energy = -out->log_evidence;
  if (bs.active) {
    set_error("gpl_energy_grad_laplace: non-log-concave sites at the mode; no gradient");
    return GPL_ERR_NOT_LOGCONCAVE;
  }
  if (nparam == 0) return GPL_OK;
  GradParam gp[MAX_GRAD_PARAMS];
  for (int j = 0; j < nparam; ++j) {
    if (param_factor[j] >= st.hp.plan.nfactors) {
      set_error("gpl_energy_grad_laplace: parameter %d refers to factor %d of %d", j, param_factor[j],
                st.hp.plan.nfactors);
      return GPL_ERR_BAD_ARG;
    }
    gp[j].factor = param_factor[j];
    gp[j].slot = param_slot[j];
  }
  // X = L^-1 in a third buffer (the factor is still needed for R t), diag(B^-1), pv
  DevBuf X, Vp, Vm, s2, u, p, gout;
  GPL_TRY(X.alloc(ctx, sizeof(double) * st.ld * n));
  for (DevBuf* b : {&Vp, &Vm, &s2, &u, &p}) GPL_TRY(b->alloc(ctx, sizeof(double) * n));
  GPL_TRY(copy_lower(ctx, n, st.B.d(), st.ld, X.d(), st.ld));
  GPL_TRY(dtrtri_with(ctx, st.dinv.d(), n, X.d(), st.ld));
  GPL_TRY(lower_colsumsq(ctx, n, X.d(), st.ld, st.v[V_T1].d()));                   // diag(B^-1)
  GPL_TRY(vec_op(ctx, VOP_MUL, n, st.v[V_V].d(), st.v[V_T1].d(), st.v[V_T2].d()));
  GPL_TRY(vec_op(ctx, VOP_SUB, n, st.v[V_V].d(), st.v[V_T2].d(), st.v[V_FVAR].d()));   // pv
  // d3 = (g2(f + h) - g2(f - h)) / 2h with g2 = -1 / V(f)
  const double h = 1e-4;
  GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, st.v[V_T1].d(), h));
  GPL_TRY(vec_op(ctx, VOP_ADD, n, st.v[V_F].d(), st.v[V_T1].d(), st.v[V_T2].d()));
  GPL_TRY(lik_pseudodata_la(ctx, st.lik, n, st.v[V_T2].d(), st.y.d(), st.v[V_ZOLD].d(), Vp.d()));
  GPL_TRY(vec_op(ctx, VOP_SUB, n, st.v[V_F].d(), st.v[V_T1].d(), st.v[V_T2].d()));
  GPL_TRY(lik_pseudodata_la(ctx, st.lik, n, st.v[V_T2].d(), st.y.d(), st.v[V_ZOLD].d(), Vm.d()));
  GPL_TRY(vec_op(ctx, VOP_INV, n, Vp.d(), nullptr, Vp.d()));
  GPL_TRY(vec_op(ctx, VOP_INV, n, Vm.d(), nullptr, Vm.d()));
  GPL_TRY(vec_op(ctx, VOP_SUB, n, Vm.d(), Vp.d(), st.v[V_T1].d()));                // g2+ - g2- = 1/V- - 1/V+
  GPL_TRY(vec_op(ctx, VOP_MUL, n, st.v[V_T1].d(), st.v[V_FVAR].d(), st.v[V_T2].d()));
  GPL_TRY(vec_op(ctx, VOP_AXPBY, n, st.v[V_T2].d(), st.v[V_T2].d(), s2.d(), 0.25 / h, 0.0));   // pv d3 / 2
  // u = s2 - R (K s2),  R t = C^-1 t by the factor;  p = alpha + u
  GPL_TRY(gemv_t(ctx, n, n, st.K.d(), st.ld, s2.d(), st.v[V_P].d()));
  GPL_TRY(solve_alpha(ctx, n, st.B.d(), st.ld, st.dinv.d(), st.v[V_SQRTV].d(), st.v[V_P].d(),
                      st.v[V_CM].d(), st.v[V_T1].d(), st.v[V_T2].d()));
  GPL_TRY(vec_op(ctx, VOP_SUB, n, s2.d(), st.v[V_CM].d(), u.d()));
  GPL_TRY(vec_op(ctx, VOP_ADD, n, st.v[V_ALPHA].d(), u.d(), p.d()));
  // B^-1 = X^T X into the K buffer (K is dead from here on), then the fused trace
  GPL_TRY(lauum_oop(ctx, n, X.d(), st.ld, st.K.d(), st.ld));
  GPL_TRY(gout.alloc(ctx, sizeof(double) * (MAX_GRAD_PARAMS + 8)));
  GPL_TRY(grad_trace(ctx, st.hp.plan, st.Z.d(), n, st.Zaux.d(), n, st.K.d(), st.ld, st.v[V_S].d(),
                     p.d(), nparam, gp, gout.d(), -1, 0, u.d()));
  double hg[MAX_GRAD_PARAMS];
  GPL_CUDA(cudaMemcpyAsync(hg, gout.p, sizeof(double) * nparam, cudaMemcpyDeviceToHost, ctx->stream));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int j = 0; j < nparam; ++j) grad[j] = -0.5 * hg[j];
  return GPL_OK;
}

int grad_trace_block_impl(Ctx* ctx, const gpl_cf_model* cf, const double* x, int64_t n, int d,
                          int64_t col0, int64_t ncols, const double* Cinv, int64_t ld,
                          const double* s, const double* alpha, int nparam,
                          const int32_t* param_factor, const int32_t* param_slot, double* out,
                          double* diag_sum) {
  if (!cf || !x || n <= 0 || d <= 0 || !Cinv || !s || !alpha || nparam < 0 ||
      nparam > MAX_GRAD_PARAMS || !out) {
    set_error("gpl_grad_trace_block_dev: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  bool want_aux = false;
  for (int j = 0; j < nparam; ++j)
    if (param_factor[j] >= 0 && param_slot[j] == 3) want_aux = true;
  HostPlan hp;
  GPL_TRY(make_plan(cf, d, 0, &hp, want_aux));
  const int F = hp.plan.nfeat > 0 ? hp.plan.nfeat : 1, FA = hp.plan.naux;
  DevBuf Z, Zaux, gout, dg, s2;
  GPL_TRY(Z.alloc(ctx, sizeof(double) * n * F));
  if (FA > 0) GPL_TRY(Zaux.alloc(ctx, sizeof(double) * n * FA));
  GPL_TRY(prep_features(ctx, hp, x, n, n, Z.d(), n, FA > 0 ? Zaux.d() : nullptr, n));
  GradParam gp[MAX_GRAD_PARAMS];
  int pos[MAX_GRAD_PARAMS], ncf = 0;
  for (int j = 0; j < nparam; ++j) {
    out[j] = 0.0;
    if (param_factor[j] < 0) continue;
    if (param_factor[j] >= hp.plan.nfactors) {
      set_error("gpl_grad_trace_block_dev: parameter %d refers to factor %d of %d", j,
                param_factor[j], hp.plan.nfactors);
      return GPL_ERR_BAD_ARG;
    }
    gp[ncf].factor = param_factor[j];
    gp[ncf].slot = param_slot[j];
    pos[ncf++] = j;
  }
  GPL_TRY(gout.alloc(ctx, sizeof(double) * (MAX_GRAD_PARAMS + 8)));
  GPL_TRY(grad_trace(ctx, hp.plan, Z.d(), n, Zaux.d(), n, Cinv, ld, s, alpha, ncf, gp, gout.d(),
                     col0, ncols));
  // diagonal of the block: element (col0 + j, j) of the slab
  GPL_TRY(dg.alloc(ctx, sizeof(double) * ncols));
  GPL_TRY(s2.alloc(ctx, sizeof(double) * ncols));
  GPL_TRY(diag_extract(ctx, ncols, Cinv + col0, ld, dg.d()));
  GPL_TRY(vec_op(ctx, VOP_MUL, ncols, s + col0, s + col0, s2.d()));
  GPL_TRY(reduce(ctx, RED_DOT, ncols, dg.d(), s2.d(), 0));
  double sc[1];
  GPL_TRY(fetch_scalars(ctx, 1, sc));
  if (diag_sum) *diag_sum = sc[0];
  double hg[MAX_GRAD_PARAMS];
  if (ncf > 0) {
    GPL_CUDA(cudaMemcpyAsync(hg, gout.p, sizeof(double) * ncf, cudaMemcpyDeviceToHost, ctx->stream));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < ncf; ++q) out[pos[q]] = hg[q];
  }
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
  // non-log-concave fit: training sites in the permuted (good-first) order of the block factor
  const int64_t ng = fit->ng, nb = fit->nb;
  const double* xtrain = nb > 0 ? fit->xp.d() : fit->x.d();
  const double* alpha = nb > 0 ? fit->alpha_p.d() : fit->alpha.d();
  DevBuf dxt, Zx, Zt, Kt, dmean, dvar, tmp, tmp2, sinv, scratch;
  GPL_TRY(upload(ctx, dxt, xt, (size_t)nt * d));
  GPL_TRY(Zx.alloc(ctx, sizeof(double) * n * F));
  GPL_TRY(Zt.alloc(ctx, sizeof(double) * nt * F));
  GPL_TRY(prep_features(ctx, hp, xtrain, n, n, Zx.d(), n));
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
  GPL_TRY(tmp2.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(sinv.alloc(ctx, sizeof(double) * n));
  GPL_TRY(scratch.alloc(ctx, sizeof(double) * (chunk > 0 ? chunk : 1)));
  if (var) GPL_TRY(vec_op(ctx, VOP_INV, ng, fit->sqrtV.d(), nullptr, sinv.d()));
  for (int64_t r0 = 0; r0 < nt; r0 += chunk) {
    const int64_t c = (nt - r0) < chunk ? (nt - r0) : chunk;
    GPL_TRY(cf_pairwise(ctx, hp.plan, Zt.d() + r0, c, nt, Zx.d(), n, n, false, 0.0, Kt.d(), ldk));
    GPL_TRY(gemv_n(ctx, c, n, Kt.d(), ldk, alpha, dmean.d() + r0));
    if (var) {
      // var = diag(Ktt) - |Kt_g C_chol^-T|^2 [- r S^-1 r^T, r = Kt_b - (Kt_g C_chol^-T) W^T]
      GPL_TRY(cf_diag(ctx, hp.plan, Zt.d() + r0, nt, Zt.d() + r0, nt, c, 0.0, dvar.d() + r0));
      GPL_TRY(scale_cols(ctx, c, ng, Kt.d(), ldk, sinv.d()));
      GPL_TRY(dtrsm_right_with(ctx, fit->dinv.d(), true, c, ng, fit->Lb.d(), fit->ld, Kt.d(), ldk));
      GPL_TRY(rowsumsq(ctx, c, ng, Kt.d(), ldk, tmp.d() + r0));
      if (nb > 0) {                                    // LU branch of R/gp_pred.R:220-223
        double* R = Kt.d() + ng * ldk;
        GPL_TRY(dgemm(ctx, false, true, c, nb, ng, -1.0, Kt.d(), ldk, fit->W.d(), fit->ldw, 1.0, R,
                      ldk, false));
        GPL_TRY(bad_quad(ctx, fit, c, R, ldk, tmp2.d() + r0, scratch.d()));
      }
    }
  }
  if (var) {
    GPL_TRY(vec_op(ctx, VOP_SUB, nt, dvar.d(), tmp.d(), dvar.d()));
    if (nb > 0) GPL_TRY(vec_op(ctx, VOP_SUB, nt, dvar.d(), tmp2.d(), dvar.d()));
    GPL_TRY(download(ctx, var, dvar.d(), nt));
  }
  GPL_TRY(download(ctx, mean, dmean.d(), nt));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPL_OK;
}

// ---- joint prediction / draws (gp_draw, R/gp_draw.R; gp_pred_post(cov = TRUE),
// R/gp_pred.R:202-232).  d_mean: nt, d_cov: nt x nt (ld ldc) on the device, cov includes jitter.
int pred_full_cov_dev(Ctx* ctx, const Fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                      uint64_t term_mask, double jitter, double* d_mean, double* d_cov, int64_t ldc) {
  const int64_t n = fit->n, ng = fit->ng, nb = fit->nb;
  const int d = fit->d;
  if ((double)nt * (double)n * 8.0 > 48e9) {
    set_error("joint prediction of %lld points against n = %lld needs a %0.f GB cross-covariance; "
              "predict fewer points at a time", (long long)nt, (long long)n, 8e-9 * nt * n);
    return GPL_ERR_NOMEM;
  }
  HostPlan hp;
  GPL_TRY(make_plan(cf, d, term_mask, &hp));
  const int F = hp.plan.nfeat > 0 ? hp.plan.nfeat : 1;
  const double* xtrain = nb > 0 ? fit->xp.d() : fit->x.d();
  const double* alpha = nb > 0 ? fit->alpha_p.d() : fit->alpha.d();
  DevBuf dxt, Zx, Zt, Kt, sinv;
  const int64_t ldk = round_up(nt, 16);
  GPL_TRY(upload(ctx, dxt, xt, (size_t)nt * d));
  GPL_TRY(Zx.alloc(ctx, sizeof(double) * n * F));
  GPL_TRY(Zt.alloc(ctx, sizeof(double) * nt * F));
  GPL_TRY(prep_features(ctx, hp, xtrain, n, n, Zx.d(), n));
  GPL_TRY(prep_features(ctx, hp, dxt.d(), nt, nt, Zt.d(), nt));
  GPL_TRY(Kt.alloc(ctx, sizeof(double) * ldk * n));
  GPL_TRY(sinv.alloc(ctx, sizeof(double) * n));
  GPL_TRY(cf_pairwise(ctx, hp.plan, Zt.d(), nt, nt, Zx.d(), n, n, false, 0.0, Kt.d(), ldk));
  GPL_TRY(gemv_n(ctx, nt, n, Kt.d(), ldk, alpha, d_mean));
  GPL_TRY(cf_pairwise(ctx, hp.plan, Zt.d(), nt, nt, nullptr, nt, nt, true, jitter, d_cov, ldc));
  GPL_TRY(vec_op(ctx, VOP_INV, ng, fit->sqrtV.d(), nullptr, sinv.d()));
  GPL_TRY(scale_cols(ctx, nt, ng, Kt.d(), ldk, sinv.d()));
  GPL_TRY(dtrsm_right_with(ctx, fit->dinv.d(), true, nt, ng, fit->Lb.d(), fit->ld, Kt.d(), ldk));
  GPL_TRY(dgemm(ctx, false, true, nt, nt, ng, -1.0, Kt.d(), ldk, Kt.d(), ldk, 1.0, d_cov, ldc, false));
  if (nb > 0) {
    // cov -= r S^-1 r^T = (r V_b^-1) r^T - H H^T,  H = (r V_b^-1) Lk A_chol^-T   (cf. bad_quad)
    double* R = Kt.d() + ng * ldk;
    GPL_TRY(dgemm(ctx, false, true, nt, nb, ng, -1.0, Kt.d(), ldk, fit->W.d(), fit->ldw, 1.0, R, ldk,
                  false));
    DevBuf G, H;
    GPL_TRY(G.alloc(ctx, sizeof(double) * ldk * nb));
    GPL_TRY(H.alloc(ctx, sizeof(double) * ldk * nb));
    GPL_TRY(copy_block(ctx, nt, nb, R, ldk, G.d(), ldk));
    GPL_TRY(scale_cols(ctx, nt, nb, G.d(), ldk, fit->invVb.d()));
    GPL_TRY(dgemm(ctx, false, false, nt, nb, nb, 1.0, G.d(), ldk, fit->Lk.d(), fit->ldp, 0.0, H.d(),
                  ldk, false));
    GPL_TRY(dtrsm_right_with(ctx, fit->Ac_dinv.d(), true, nt, nb, fit->Ac.d(), fit->ldp, H.d(), ldk));
    GPL_TRY(dgemm(ctx, false, true, nt, nt, nb, -1.0, G.d(), ldk, R, ldk, 1.0, d_cov, ldc, false));
    GPL_TRY(dgemm(ctx, false, true, nt, nt, nb, 1.0, H.d(), ldk, H.d(), ldk, 1.0, d_cov, ldc, false));
  }
  return GPL_OK;
}

int pred_full_cov_impl(Ctx* ctx, const Fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                       uint64_t term_mask, double jitter, double* mean, double* cov) {
  if (!fit || fit->method > 1 || nt <= 0 || !xt || !mean || !cov) {
    set_error("gpl_pred_full_cov: bad arguments (method_full and method_fitc fits)");
    return GPL_ERR_BAD_ARG;
  }
  DevBuf dm, dc;
  const int64_t ldc = round_up(nt, 16);
  GPL_TRY(dm.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(dc.alloc(ctx, sizeof(double) * ldc * nt));
  if (fit->method == 0)
    GPL_TRY(pred_full_cov_dev(ctx, fit, cf, xt, nt, term_mask, jitter, dm.d(), dc.d(), ldc));
  else
    GPL_TRY(pred_fitc_cov_dev(ctx, fit, cf, xt, nt, term_mask, jitter, dm.d(), dc.d(), ldc));
  GPL_TRY(download(ctx, mean, dm.d(), nt));
  GPL_CUDA(cudaMemcpy2DAsync(cov, sizeof(double) * nt, dc.d(), sizeof(double) * ldc,
                             sizeof(double) * nt, nt, cudaMemcpyDeviceToHost, ctx->stream));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPL_OK;
}

// out (nt x draws) = mean + chol(cov) normals.  fit == NULL: draws from the prior at xt (nt x d),
// mean 0, cov = K(xt, xt) + jitter I (gp_draw_prior.method_full, R/gp_draw.R:66-78).
int draw_full_impl(Ctx* ctx, const Fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                   int d_prior, uint64_t term_mask, double jitter, const double* normals,
                   int64_t draws, double* out) {
  if ((fit && fit->method > 1) || nt <= 0 || draws <= 0 || !xt || !normals || !out ||
      (!fit && d_prior <= 0)) {
    set_error("gpl_draw_full: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  DevBuf dm, dc, dcd, Lc, dn, dout;
  const int64_t ldc = round_up(nt, 16);
  GPL_TRY(dm.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(dc.alloc(ctx, sizeof(double) * ldc * nt));
  GPL_TRY(Lc.alloc(ctx, sizeof(double) * ldc * nt));
  GPL_TRY(dcd.alloc(ctx, sizeof(double) * ((nt + NB - 1) / NB) * NB * NB));
  if (fit && fit->method == 1) {
    GPL_TRY(pred_fitc_cov_dev(ctx, fit, cf, xt, nt, term_mask, jitter, dm.d(), dc.d(), ldc));
  } else if (fit) {
    GPL_TRY(pred_full_cov_dev(ctx, fit, cf, xt, nt, term_mask, jitter, dm.d(), dc.d(), ldc));
  } else {
    HostPlan hp;
    GPL_TRY(make_plan(cf, d_prior, term_mask, &hp));
    const int F = hp.plan.nfeat > 0 ? hp.plan.nfeat : 1;
    DevBuf dxt, Zt;
    GPL_TRY(upload(ctx, dxt, xt, (size_t)nt * d_prior));
    GPL_TRY(Zt.alloc(ctx, sizeof(double) * nt * F));
    GPL_TRY(prep_features(ctx, hp, dxt.d(), nt, nt, Zt.d(), nt));
    GPL_TRY(cf_pairwise(ctx, hp.plan, Zt.d(), nt, nt, nullptr, nt, nt, true, jitter, dc.d(), ldc));
    GPL_TRY(vec_op(ctx, VOP_FILL, nt, nullptr, nullptr, dm.d(), 0.0));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));     // dxt, Zt are released at scope exit
  }
  int32_t info = 0;
  GPL_TRY(dpotrf_with(ctx, dcd.d(), nt, dc.d(), ldc, &info, nullptr));       // t(chol(pred$cov))
  GPL_TRY(tril_scale(ctx, nt, dc.d(), ldc, nullptr, Lc.d(), ldc));
  const int64_t ldn = round_up(nt, 16);
  GPL_TRY(dn.alloc(ctx, sizeof(double) * ldn * draws));
  GPL_TRY(dout.alloc(ctx, sizeof(double) * ldn * draws));
  GPL_CUDA(cudaMemcpy2DAsync(dn.d(), sizeof(double) * ldn, normals, sizeof(double) * nt,
                             sizeof(double) * nt, draws, cudaMemcpyHostToDevice, ctx->stream));
  GPL_TRY(dgemm(ctx, false, false, nt, draws, nt, 1.0, Lc.d(), ldc, dn.d(), ldn, 0.0, dout.d(), ldn,
                false));
  // + mean on every column: rows scaled by 1 and shifted -- reuse scale_block's machinery via axpy
  for (int64_t j = 0; j < draws; ++j)
    GPL_TRY(vec_op(ctx, VOP_ADD, nt, dout.d() + j * ldn, dm.d(), dout.d() + j * ldn));
  GPL_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * nt, dout.d(), sizeof(double) * ldn,
                             sizeof(double) * nt, draws, cudaMemcpyDeviceToHost, ctx->stream));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPL_OK;
}

int fit_get_impl(Ctx* ctx, const Fit* fit, const char* what, double* out) {
  if (!fit || !what || !out) return GPL_ERR_BAD_ARG;
  const int64_t n = fit->n;
  if (!strcmp(what, "alpha") && fit->method != 2) {
    GPL_TRY(download(ctx, out, fit->alpha.d(), n));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    return GPL_OK;
  }
  if (!strcmp(what, "C_chol") && fit->method == 0 && fit->nb > 0) {
    set_error("gpl_fit_get: this fit has non-log-concave sites (the reference stores C_lu, not "
              "C_chol); its block factor stays on the device");
    return GPL_ERR_BAD_ARG;
  }
  if (!strcmp(what, "C_chol") && fit->method == 0) {
    DevBuf tmp;
    GPL_TRY(tmp.alloc(ctx, sizeof(double) * n * n));
    dim3 grid((unsigned)((n + 255) / 256), grid_y_cap(n));
    cchol_kernel<<<grid, 256, 0, ctx->stream>>>(n, fit->Lb.d(), fit->ld, fit->sqrtV.d(), tmp.d());
    GPL_LAUNCH_CHECK(ctx);
    GPL_TRY(download(ctx, out, tmp.d(), (size_t)n * n));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    return GPL_OK;
  }
  if (fit->method == 1 && (!strcmp(what, "Kxz") || !strcmp(what, "Kz_chol"))) {
    const bool kxz = !strcmp(what, "Kxz");
    if (kxz && !fit->Kxz.p) {
      set_error("gpl_fit_get: this fit streamed Kxz (never materialised); refit with GPL_FITC_STREAM_GB raised");
      return GPL_ERR_BAD_ARG;
    }
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
  if (fit->method == 1 && !strcmp(what, "var_train")) return fitc_var_train(ctx, fit, out);
  if (fit->method == 2 && (!strcmp(what, "wcov") || !strcmp(what, "wmean"))) {
    const int64_t M = fit->m;
    if (!strcmp(what, "wmean")) {
      GPL_TRY(download(ctx, out, fit->alpha.d(), M));
    } else {
      GPL_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * M, fit->Lm.d(), sizeof(double) * fit->ldm,
                                 sizeof(double) * M, M, cudaMemcpyDeviceToHost, ctx->stream));
    }
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
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

int gpl_energy_grad_laplace(gpl_ctx* h, const gpl_cf_model* cf, const gpl_lik* lik,
                            const gpl_approx* approx, const double* x, int64_t n, int32_t d,
                            const double* y, const double* trials, const double* offset, double jitter,
                            int32_t nparam, const int32_t* param_factor, const int32_t* param_slot,
                            double* energy, double* grad, gpl_fit_out* out) {
  int rc = gpl::energy_grad_laplace_impl(reinterpret_cast<gpl::Ctx*>(h), cf, lik, approx, x, n, d, y,
                                         trials, offset, jitter, nparam, param_factor, param_slot,
                                         energy, grad, out);
  if (rc != GPL_OK) cudaStreamSynchronize(reinterpret_cast<gpl::Ctx*>(h)->stream);
  return rc;
}

int gpl_grad_trace_block_dev(gpl_ctx* h, const gpl_cf_model* cf, const double* x, int64_t n,
                             int32_t d, int64_t col0, int64_t ncols, const double* Cinv_blk,
                             int64_t ld, const double* s, const double* alpha, int32_t nparam,
                             const int32_t* param_factor, const int32_t* param_slot, double* out,
                             double* diag_sum) {
  return gpl::grad_trace_block_impl(reinterpret_cast<gpl::Ctx*>(h), cf, x, n, d, col0, ncols,
                                    Cinv_blk, ld, s, alpha, nparam, param_factor, param_slot, out,
                                    diag_sum);
}

int gpl_pred_full_cov(gpl_ctx* h, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt,
                      int64_t nt, uint64_t term_mask, double jitter, double* mean, double* cov) {
  return gpl::pred_full_cov_impl(reinterpret_cast<gpl::Ctx*>(h), reinterpret_cast<const gpl::Fit*>(fit),
                                 cf, xt, nt, term_mask, jitter, mean, cov);
}

int gpl_draw_full(gpl_ctx* h, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                  int32_t d, uint64_t term_mask, double jitter, const double* normals, int64_t draws,
                  double* out) {
  return gpl::draw_full_impl(reinterpret_cast<gpl::Ctx*>(h), reinterpret_cast<const gpl::Fit*>(fit), cf,
                             xt, nt, d, term_mask, jitter, normals, draws, out);
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
