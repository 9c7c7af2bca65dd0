/*
 * gplite_b200.h -- C ABI of libgplite_b200.so: the B200 (sm_100a) implementation of the
 * dense-GP hot path of the R package gplite (jpiironen/gplite v0.13.0).
 *
 * This is the drop-in boundary.  Everything here is `extern "C"`, plain pointers, sizes and
 * POD structs: no R, Rcpp, torch or C++ types cross it.  R glue (.Call stubs, see
 * INTEGRATION.md) or any FFI (the repo's own host layer uses Python ctypes) binds these
 * symbols directly.  All matrices are IEEE fp64, COLUMN-MAJOR (R layout), dense, with the
 * leading dimension equal to the row count unless an explicit `ld` argument says otherwise.
 *
 * What each entry point replaces in the reference (paths relative to the gplite repo):
 *   gpl_cf_eval / gpl_cf_eval_dev     eval_cf.*            R/cf.R:402-497  and the native loops
 *                                     cf_sexp_c/cf_matern32_c/cf_matern52_c/cf_nn_c
 *                                     src/cf.cpp:9,48,86,125 (bound in src/RcppExports.cpp:16-87,
 *                                     R/RcppExports.R:4-18); prepare_inputmat R/misc.R:17-31
 *   gpl_cf_{sexp,matern32,matern52,nn}_c   the four .Call targets, same argument meaning
 *   gpl_fit_full                      fit_laplace.method_full R/gp_fit.R:91-99 + laplace.method_full
 *                                     R/laplace.R:13-38,274-297; fit_ep.method_full R/gp_fit.R:150-161
 *                                     + ep.method_full R/ep.R:16-49,286-330
 *   gpl_fit_fitc                      fit_laplace.method_fitc R/gp_fit.R:101-123, fit_ep.method_fitc
 *                                     :163-188, laplace/ep .method_fitc R/laplace.R:122-152,299-325,
 *                                     R/ep.R:148-185,333-374, inv_lemma_* R/linear_algebra.R:3-126
 *   gpl_energy_grad                   gp_energy R/gp.R:138-147 evaluated through gp_fit, PLUS the
 *                                     analytic gradient (new capability: the reference's gp_optim is
 *                                     derivative-free, R/gp_optim.R:94)
 *   gpl_pred_full / gpl_pred_fitc     gp_pred_post.method_full/fitc R/gp_pred.R:202-296
 *   gpl_dpotrf_dev, gpl_dgemm_dev ... base-R chol()/forwardsolve()/backsolve()/%*% call sites
 *                                     (R/laplace.R:27-30, R/ep.R:38-42, R/gp_pred.R:218)
 *
 * Error model: every function returns a gpl_status.  A failure message for the calling thread
 * is available from gpl_last_error().  Non-positive-definite factorisations return
 * GPL_ERR_NOT_PD (R's chol() error); the glue decides whether that aborts (method_full,
 * R/gp_fit.R:95-97) or becomes log_evidence = -Inf (FITC, R/gp_fit.R:113-121).
 * There is NO CPU fallback: without a CUDA device every compute entry returns GPL_ERR_CUDA.
 */
#ifndef GPLITE_B200_H
#define GPLITE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPL_ABI_VERSION 2   /* 2: gpl_fit_out grew accepted / damping_final; gpl_dist_* entries */

typedef enum gpl_status {
  GPL_OK = 0,
  GPL_ERR_BAD_ARG = 1,
  GPL_ERR_NOT_PD = 2,       /* Cholesky pivot <= 0 or NaN; gpl_last_error() names the pivot */
  GPL_ERR_NONFINITE = 3,    /* NaN pseudo-data in EP (R/ep.R:30-32) */
  GPL_ERR_CUDA = 4,
  GPL_ERR_NCCL = 5,
  GPL_ERR_NOMEM = 6,
  GPL_ERR_NOT_LOGCONCAVE = 7 /* some pseudo-variance V <= 0 on a path without the non-log-concave
                                branch (the multi-GPU drivers); gpl_fit_full / gpl_fit_fitc handle
                                V <= 0 themselves and report gpl_fit_out.n_bad */
} gpl_status;

/* ---- model descriptors ---------------------------------------------------------------- */

typedef enum gpl_cf_kind {
  GPL_CF_CONST = 0,     /* p = {magn}                      R/cf.R:413-424 */
  GPL_CF_LIN = 1,       /* p = {magn}                      R/cf.R:426-435 */
  GPL_CF_SEXP = 2,      /* p = {lscale, magn}              src/cf.cpp:9-42 */
  GPL_CF_MATERN32 = 3,  /* p = {lscale, magn}              src/cf.cpp:48-80 */
  GPL_CF_MATERN52 = 4,  /* p = {lscale, magn}              src/cf.cpp:86-119 */
  GPL_CF_NN = 5         /* p = {sigma0, sigma, magn}       src/cf.cpp:125-162 */
} gpl_cf_kind;

/* One multiplicative factor of one summand.  gp$cfs (a list, summed: R/cf.R:402-411) whose
 * members may be cf_prod (multiplied: R/cf.R:491-497) flattens to a list of factors, each
 * tagged with the index `term` of the summand it multiplies into. */
typedef struct gpl_cf_factor {
  int32_t kind;          /* gpl_cf_kind of the (base) covariance function */
  int32_t term;          /* 0-based index of the summand in gp$cfs */
  int32_t periodic;      /* 1: cf_periodic -- the kind is applied to cbind(sin(w x), cos(w x)),
                            w = 2*pi/period (R/cf.R:478-489); never normalised (R/cf.R:852-855) */
  int32_t normalize;     /* 1: inputs are (x - means) / scales (R/misc.R:25-29) */
  int32_t nvars;         /* number of selected input columns; 0 = all d columns (vars = NULL) */
  int32_t reserved;
  const int32_t* vars;   /* nvars 0-based column indices, or NULL */
  const double* means;   /* per selected column, used iff normalize (learn_scales, R/cf.R:836-845) */
  const double* scales;
  double p[3];           /* parameters on the natural (not log) scale, see gpl_cf_kind */
  double period;
} gpl_cf_factor;

typedef struct gpl_cf_model {
  int32_t nfactors;
  int32_t nterms;
  const gpl_cf_factor* factors;   /* ordered by term */
} gpl_cf_model;

typedef enum gpl_lik_kind {
  GPL_LIK_GAUSSIAN = 0,   /* param = sigma                   R/lik.R:188-192,301-308 */
  GPL_LIK_BERNOULLI = 1,  /*                                 R/lik.R:310-312,372-374 */
  GPL_LIK_BINOMIAL = 2,   /* needs trials                    R/lik.R:314-330,376-397 */
  GPL_LIK_BETABINOM = 3,  /* param = phi, needs trials       R/lik.R:332-353,400-424 */
  GPL_LIK_POISSON = 4     /* log link                        R/lik.R:355-366,426-443 */
} gpl_lik_kind;

typedef enum gpl_link { GPL_LINK_IDENTITY = 0, GPL_LINK_LOGIT = 1, GPL_LINK_PROBIT = 2,
                        GPL_LINK_LOG = 3 } gpl_link;

typedef struct gpl_lik {
  int32_t kind;
  int32_t link;
  double param;           /* sigma (gaussian) or phi (betabinom) */
} gpl_lik;

typedef enum gpl_approx_kind { GPL_APPROX_LAPLACE = 0, GPL_APPROX_EP = 1 } gpl_approx_kind;

typedef struct gpl_approx {
  int32_t kind;
  int32_t maxiter;        /* approx_laplace: 30; approx_ep: 100       R/approx.R:60,68 */
  double tol;             /* laplace: 1e-4; EP full 1e-4, EP fitc 1e-3 R/ep.R:286,335 */
  double damping;         /* EP, default 0.9 */
  double damping_min;     /* EP, 0.1 (R/gp_fit.R:158) */
  int32_t quad_order;     /* EP Gauss-Hermite order, default 11 */
  int32_t reserved;
  const double* quad_x;   /* quad_order nodes of probabilists' He_p (R/quadrature.R:28-34) */
  const double* quad_w;   /* quad_order weights, summing to 1 */
} gpl_approx;

/* ---- context and handles --------------------------------------------------------------- */

typedef struct gpl_ctx gpl_ctx;   /* one per (process, device): stream, workspaces */
typedef struct gpl_fit gpl_fit;   /* device-resident result of a fit (C_chol, alpha, Kxz ...) */

int gpl_abi_version(void);
const char* gpl_last_error(void);
int gpl_device_count(void);

/* stream: 0 -> the library creates its own (non-blocking) stream; otherwise a cudaStream_t cast to
 * uintptr_t on which all work is enqueued.  To share the legacy default stream (what
 * torch.cuda.current_stream().cuda_stream == 0 denotes) pass cudaStreamLegacy, i.e. 1. */
int gpl_ctx_create(int device, uint64_t stream, gpl_ctx** out);
int gpl_ctx_destroy(gpl_ctx* ctx);
int gpl_ctx_sync(gpl_ctx* ctx);
/* Give the device memory cached by the context back to the driver (the stream-ordered pool keeps the
 * multi-GB buffers of finished fits for reuse).  Live gpl_fit handles are not affected. */
int gpl_ctx_release_cache(gpl_ctx* ctx);
/* number of kernels this context has launched so far (for bench.py's gpu_launches) */
int64_t gpl_ctx_launch_count(gpl_ctx* ctx);
/* Optional per-launch CUDA-event timing of the FP64 tensor-core GEMM/SYRK kernel (the dominant
 * kernel of every factorisation).  enable != 0 starts a fresh accumulation; _read harvests the
 * events recorded so far: summed kernel milliseconds, executed flops (2mnk, halved for
 * lower-only updates) and launch count. */
int gpl_ctx_profile(gpl_ctx* ctx, int32_t enable);
int gpl_ctx_profile_read(gpl_ctx* ctx, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches);
/* Tuning knob for the GEMM tile choice of gpl_dgemm_dev and everything built on it: 0 = automatic
 * (default), 1 / 2 / 3 = force the 128x128 / 64x64 / 32x32 CTA tile (tools/gemm_tune.py). */
int gpl_ctx_gemm_config(gpl_ctx* ctx, int32_t cfg);
/* plain device memory helpers so that an FFI without its own CUDA binding can stage data */
int gpl_dev_alloc(gpl_ctx* ctx, uint64_t bytes, void** dptr);
int gpl_dev_free(gpl_ctx* ctx, void* dptr);
int gpl_memcpy_h2d(gpl_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int gpl_memcpy_d2h(gpl_ctx* ctx, void* dst, const void* src, uint64_t bytes);

/* ---- inducing-point selection ---------------------------------------------- */

/* k-means (Lloyd) on x (n x d host, column-major), as used by get_inducing (R/gp_fit.R:190-234;
 * the reference calls stats::kmeans on the standardised inputs -- standardising is the caller's
 * job).  init_idx: m distinct rows of x used as initial centres (the reference draws them with R's
 * RNG, which cannot be reproduced outside R).  Stops when no assignment changes or after maxiter
 * (stats::kmeans: 10).  centers: m x d column-major; labels (may be NULL): n; iters (may be NULL).
 * Empty clusters keep their centre; ties go to the lowest centre index. */
int gpl_kmeans(gpl_ctx* ctx, const double* x, int64_t n, int32_t d, const int64_t* init_idx, int64_t m,
               int32_t maxiter, double* centers, int32_t* labels, int32_t* iters);

/* ---- covariance matrices -------------------------------------------- */

/* K = eval_cf(cfs, x1, x2, cfind)  [+ jitter on the diagonal when `same` != 0].
 * Host buffers.  x1: n1 x d, x2: n2 x d (ignored when same != 0: x2 := x1 and the symmetric
 * fast path is used), out: n1 x n2, or min(n1,n2) when diag_only.  term_mask bit k selects
 * summand k (cfind); pass 0 for all. */
int gpl_cf_eval(gpl_ctx* ctx, const gpl_cf_model* cf, const double* x1, int64_t n1,
                const double* x2, int64_t n2, int32_t d, int32_t same, uint64_t term_mask,
                int32_t diag_only, double jitter, double* out);

/* Same on device-resident buffers; K has leading dimension ldk >= n1 (ldk even). */
int gpl_cf_eval_dev(gpl_ctx* ctx, const gpl_cf_model* cf, const double* d_x1, int64_t n1,
                    const double* d_x2, int64_t n2, int32_t d, int32_t same, uint64_t term_mask,
                    int32_t diag_only, double jitter, double* d_out, int64_t ldk);

/* Rectangular block of K on device buffers whose row counts differ from their leading dimensions
 * (sub-blocks of a replicated x): out[i, j] = k(x1_i, x2_j), i < n1, j < n2.  Used by the
 * multi-GPU path, where each rank builds only the block columns it owns. */
int gpl_cf_eval_block_dev(gpl_ctx* ctx, const gpl_cf_model* cf, const double* d_x1, int64_t n1,
                          int64_t ldx1, const double* d_x2, int64_t n2, int64_t ldx2, int32_t d,
                          double* d_out, int64_t ldk);

/* The reference's four .Call targets with their argument meaning (src/RcppExports.cpp:16-74):
 * x1 (n1 x d), x2 (n2 x d) host column-major, out n1 x n2 (or min(n1,n2) if diag_only). */
int gpl_cf_sexp_c(gpl_ctx* ctx, const double* x1, int64_t n1, const double* x2, int64_t n2,
                  int32_t d, double lscale, double magn, int32_t diag_only, double* out);
int gpl_cf_matern32_c(gpl_ctx* ctx, const double* x1, int64_t n1, const double* x2, int64_t n2,
                      int32_t d, double lscale, double magn, int32_t diag_only, double* out);
int gpl_cf_matern52_c(gpl_ctx* ctx, const double* x1, int64_t n1, const double* x2, int64_t n2,
                      int32_t d, double lscale, double magn, int32_t diag_only, double* out);
int gpl_cf_nn_c(gpl_ctx* ctx, const double* x1, int64_t n1, const double* x2, int64_t n2,
                int32_t d, double sigma0, double sigma, double magn, int32_t diag_only,
                double* out);

/* ---- dense fp64 linear algebra on device buffers (SURVEY "library calls that act as kernels")
 * op(A): 0 = A, 1 = A^T.  Column-major, leading dimensions must be even. ------------------ */

/* C = alpha op(A) op(B) + beta C;  lower != 0: only the lower triangle of (square) C is
 * computed/written (SYRK-shaped trailing update). */
int gpl_dgemm_dev(gpl_ctx* ctx, int32_t transa, int32_t transb, int64_t m, int64_t n, int64_t k,
                  double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                  double beta, double* C, int64_t ldc, int32_t lower);
/* In-place lower Cholesky A = L L^T (R: t(chol(A))).  Strict upper triangle is scratch.
 * On GPL_ERR_NOT_PD *info is the 1-based failing pivot.  logdet (may be NULL) receives
 * 2*sum(log(diag(L))). */
int gpl_dpotrf_dev(gpl_ctx* ctx, int64_t n, double* A, int64_t lda, int32_t* info, double* logdet);
/* B <- B L^-T (trans=1) or B L^-1 (trans=0), B is m x n, L from gpl_dpotrf_dev on the same ctx
 * (uses the diagonal-block inverses cached by the factorisation). */
int gpl_dtrsm_right_dev(gpl_ctx* ctx, int32_t trans, int64_t m, int64_t n, const double* L,
                        int64_t ldl, double* B, int64_t ldb);
/* Same with an explicit block-inverse buffer (from gpl_dpotrf_get_dinv_dev), so that several
 * factors can be alive on one context (FITC keeps chol(Kz) and chol(Kz + Kzx S^-1 Kxz)). */
int gpl_dtrsm_right_dinv_dev(gpl_ctx* ctx, int32_t trans, int64_t m, int64_t n, const double* L,
                             int64_t ldl, const double* d_dinv, double* B, int64_t ldb);
/* out_i = sum_j A_ij^2 (A is m x n): column norms of forwardsolve(L, t(A)) once A <- A L^-T. */
int gpl_rowsumsq_dev(gpl_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, double* out);
/* A <- (L L^T)^-1 given L from gpl_dpotrf_dev, lower triangle (dpotri). */
int gpl_dpotri_dev(gpl_ctx* ctx, int64_t n, double* A, int64_t lda);
/* Copy out the diagonal-block inverses cached by the last gpl_dpotrf_dev on this ctx
 * (ceil(n/128) blocks of 128 x 128 doubles) so that several factors can be kept alive. */
int gpl_dpotrf_get_dinv_dev(gpl_ctx* ctx, int64_t n, double* d_dinv_out);
/* x <- L^-1 b (trans = 0) or L^-T b (trans = 1) for one right-hand side; dinv from
 * gpl_dpotrf_get_dinv_dev (NULL: the ctx cache).  b is destroyed, x must not alias b. */
int gpl_dtrsv_dev(gpl_ctx* ctx, int32_t trans, int64_t n, const double* L, int64_t ldl,
                  const double* d_dinv, double* b, double* x);
/* y = A x (trans = 0, A is m x n) or y = A^T x (trans = 1, y has n entries). */
int gpl_dgemv_dev(gpl_ctx* ctx, int32_t trans, int64_t m, int64_t n, const double* A, int64_t lda,
                  const double* x, double* y);
/* A[i, j] <- alpha * r_i * A[i, j] * c_j  (+ diag_add where i == j + diag_offset); r, c may be NULL.
 * Builds B = I + V^-1/2 K V^-1/2 (R/laplace.R:27) in place on a block of columns. */
int gpl_scale_block_dev(gpl_ctx* ctx, int64_t m, int64_t n, double* A, int64_t lda, const double* r,
                        const double* c, double alpha, double diag_add, int64_t diag_offset);
/* Site kernels on device n-vectors (the multi-GPU Newton loop keeps f, z, V replicated per rank).
 * get_pseudodata_la (R/lik.R:188-201): z, V from f, y; trials / offset may be NULL. */
int gpl_lik_pseudodata_la_dev(gpl_ctx* ctx, const gpl_lik* lik, int64_t n, const double* f,
                              const double* y, const double* trials, const double* offset,
                              double* z, double* V);
/* sum_i log p(y_i | f_i) (get_loglik, R/lik.R:297-366) returned to the host. */
int gpl_lik_loglik_dev(gpl_ctx* ctx, const gpl_lik* lik, int64_t n, const double* f, const double* y,
                       const double* trials, const double* offset, double* loglik_host);
/* EP site update on device n-vectors (get_pseudodata_ep + get_tilted_moments, R/lik.R:207-270):
 * cavity from (fmean, fvar, z_old, P_old), Gauss-Hermite tilted moments with the device arrays
 * quad_x / quad_w (quad_order entries), damped new sites z_new, V_new = 1 / P_new, P_new, plus the
 * cavity mean / variance and log Z.  *flags: bit 0 quadrature not ok, bit 1 NaN pseudo-data
 * (R/ep.R:30-32), bit 2 some V_new <= 0. */
int gpl_lik_ep_sites_dev(gpl_ctx* ctx, const gpl_lik* lik, int64_t n, const double* fmean,
                         const double* fvar, const double* z_old, const double* P_old, const double* y,
                         const double* trials, const double* offset, double damping, int32_t quad_order,
                         const double* quad_x, const double* quad_w, double* z_new, double* V_new,
                         double* P_new, double* cavity_mean, double* cavity_var, double* logZ,
                         int32_t* flags);
/* out_j = sum_i A_ij^2 for the n columns of A (m x n): diag(B^-1) from X = L^-1. */
int gpl_colsumsq_dev(gpl_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, double* out);
/* One block column's share of the gradient trace (multi-GPU energy gradient; the single-GPU path is
 * gpl_energy_grad).  Cinv_blk: columns [col0, col0 + ncols) of the lower triangle of
 * Binv = (I + S K S)^-1, all n rows with leading dimension ld (rows above the block's diagonal are
 * not read); col0 a multiple of 128.  s = V^-1/2 and alpha are device n-vectors, x is n x d on the
 * device.  out[p] = sum over the block's lower-triangular entries (mirror entries counted) of
 * (alpha_i alpha_j - s_i Binv_ij s_j) dK_ij/dlog(theta_p); *diag_sum = sum_j s_j^2 Binv_jj.
 * param_factor / param_slot as in gpl_energy_grad (entries with factor < 0 are skipped: out = 0). */
int gpl_grad_trace_block_dev(gpl_ctx* ctx, const gpl_cf_model* cf, const double* x, int64_t n,
                             int32_t d, int64_t col0, int64_t ncols, const double* Cinv_blk,
                             int64_t ld, const double* s, const double* alpha, int32_t nparam,
                             const int32_t* param_factor, const int32_t* param_slot, double* out,
                             double* diag_sum);

/* ---- fits ---------------------------------------------------------- */

typedef struct gpl_fit_out {
  double log_evidence;    /* gp$fit$log_evidence; -inf when EP hit maxiter / quad not ok */
  int32_t iters;          /* Newton iterations / EP sweeps performed */
  int32_t converged;
  int32_t quad_ok;
  int32_t n_bad;          /* sites with pseudo-variance <= 0 in the returned fit (0: C_chol; > 0: the
                           * reference's C_lu case, R/laplace.R:40-117, R/ep.R:50-138) */
  int32_t accepted;       /* EP: sweeps that were accepted (iters - accepted were discarded because
                           * max|delta fmean| grew, R/ep.R:306-310); Laplace: = iters */
  int32_t reserved;
  double damping_final;   /* EP: damp_value after the last sweep (R/ep.R:308,311); Laplace: 0 */
} gpl_fit_out;

/* method_full.  x: n x d host, y: n, trials/offset: n or NULL.  n-vectors are returned eagerly
 * into the caller's buffers (any may be NULL): fmean, z, V, alpha, and for EP fvar,
 * cavity_mean, cavity_var.  The n x n factor stays on the device behind *fit. */
int gpl_fit_full(gpl_ctx* ctx, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                 const double* x, int64_t n, int32_t d, const double* y, const double* trials,
                 const double* offset, double jitter, gpl_fit_out* out, double* fmean, double* z,
                 double* V, double* alpha, double* fvar, double* cavity_mean, double* cavity_var,
                 gpl_fit** fit);

/* method_fitc with explicit inducing points zind (m x d). */
int gpl_fit_fitc(gpl_ctx* ctx, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                 const double* x, int64_t n, int32_t d, const double* zind, int64_t m,
                 const double* y, const double* trials, const double* offset, double jitter,
                 gpl_fit_out* out, double* fmean, double* z, double* V, double* alpha, double* fvar,
                 double* cavity_mean, double* cavity_var, gpl_fit** fit);

/* Energy (= -log_evidence, without the prior term, which stays on the host: R/gp.R:142-146)
 * and its gradient w.r.t. the LOG of the listed parameters, lik_gaussian + method_full.
 * Parameter j is identified by (param_factor[j], param_slot[j]): factor index into cf->factors
 * and slot into p[] (slot 3 = period); param_factor[j] = -1 selects the likelihood sigma.
 * d_x / d_y are DEVICE pointers when on_device != 0, host pointers otherwise. */
int gpl_energy_grad(gpl_ctx* ctx, const gpl_cf_model* cf, const gpl_lik* lik, const void* x,
                    int64_t n, int32_t d, const void* y, int32_t on_device, double jitter,
                    int32_t nparam, const int32_t* param_factor, const int32_t* param_slot,
                    double* energy, double* grad, gpl_fit** fit /* may be NULL */);

/* ---- method_rf: random-feature (linearised) GP, SURVEY 8f-3 --------------------------------- */

/* Basis sizes and the random draws of rf_featmap (R/cf.R:532-690).  num_basis: one entry per
 * summand after check_num_basis (R/gp.R:205-228: cf_const 1, cf_lin one per input column; even for
 * cf_sexp).  weights: for each factor in model order that has random features --
 *   cf_sexp (also as the base of cf_periodic): W, nf x num_basis/2 column-major, = N(0,1) / lscale,
 *     nf = selected inputs (twice that for a periodic factor);
 *   cf_nn: W, nf x num_basis column-major, = N(0,1) (the prepared inputs already carry sigma),
 *     followed by w0, num_basis values = N(0,1) * sigma0 * sigma;
 * nothing for cf_lin / cf_const.  cf_matern32/52 are rejected as in the reference. */
typedef struct gpl_rf_spec {
  int32_t nterms;
  const int32_t* num_basis;
  const double* weights;
  int64_t nweights;
} gpl_rf_spec;

/* fit_laplace.method_rf (R/gp_fit.R:125-132; laplace_iter.method_rf R/laplace.R:241-262,327-350).
 * Returns fmean, z, V (n) and wmean (sum num_basis); wcov stays behind *fit ("wcov" of gpl_fit_get). */
int gpl_fit_rf(gpl_ctx* ctx, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
               const gpl_rf_spec* rf, const double* x, int64_t n, int32_t d, const double* y,
               const double* trials, const double* offset, gpl_fit_out* out, double* fmean, double* z,
               double* V, double* wmean, gpl_fit** fit);
/* featuremap(xt, cfind) of a fitted method_rf model: phi is nt x sum(num_basis), column-major, host;
 * columns of unselected summands are zero (gp_draw_analytic.method_rf, R/gp_draw.R:145-163, draws the
 * weights on the host and needs the features of the test points). */
int gpl_rf_features(gpl_ctx* ctx, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                    uint64_t term_mask, double* phi);
/* gp_pred_post.method_rf (R/gp_pred.R:298-315); term_mask = cfind. */
int gpl_pred_rf(gpl_ctx* ctx, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                uint64_t term_mask, double* mean, double* var);

/* Energy (without priors) and its gradient for approx_laplace with a non-Gaussian likelihood,
 * method_full: runs the Newton loop of gpl_fit_full, then the implicit-function gradient (mode
 * dependence included).  Covariance-function parameters only (param_factor >= 0); use a tight
 * approx->tol.  Not in the reference (its gp_optim is derivative-free, R/gp_optim.R:94).  out (iters,
 * converged, log_evidence) must not be NULL.  GPL_ERR_NOT_LOGCONCAVE when the mode has sites with
 * pseudo-variance <= 0 (energy is still returned). */
int gpl_energy_grad_laplace(gpl_ctx* ctx, const gpl_cf_model* cf, const gpl_lik* lik,
                            const gpl_approx* approx, const double* x, int64_t n, int32_t d,
                            const double* y, const double* trials, const double* offset, double jitter,
                            int32_t nparam, const int32_t* param_factor, const int32_t* param_slot,
                            double* energy, double* grad, gpl_fit_out* out);

/* ---- prediction ---------------------------------------------------- */

/* mean[nt] (and var[nt] if var != NULL) of the latent at xt (nt x d host). */
int gpl_pred_full(gpl_ctx* ctx, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt,
                  int64_t nt, uint64_t term_mask, double* mean, double* var);
int gpl_pred_fitc(gpl_ctx* ctx, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt,
                  int64_t nt, uint64_t term_mask, double* mean, double* var);

/* Joint prediction for method_full and method_fitc fits: mean[nt] and cov[nt x nt] (column-major)
 * of the latent at xt, jitter on the diagonal of cov (gp_pred_post.method_full / .method_fitc with
 * cov = TRUE, R/gp_pred.R:202-232, 262-273). */
int gpl_pred_full_cov(gpl_ctx* ctx, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt,
                      int64_t nt, uint64_t term_mask, double jitter, double* mean, double* cov);
/* gp_draw (R/gp_draw.R:66-78,105-123, mvnrnd :183-207): out (nt x draws, column-major) =
 * mean + t(chol(cov)) %*% normals, normals nt x draws standard normal variates supplied by the
 * caller (R's stats::rnorm stream is not reproducible outside R).  fit != NULL: posterior of a
 * method_full or method_fitc fit (d ignored); fit == NULL: method_full prior, mean 0,
 * cov = K(xt, xt) + jitter I, xt nt x d.
 * Offsets, the response transform and generate_target stay with the caller. */
int gpl_draw_full(gpl_ctx* ctx, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                  int32_t d, uint64_t term_mask, double jitter, const double* normals, int64_t draws,
                  double* out);

/* Materialise members of gp$fit on demand: what = "C_chol" (n x n), "Kxz" (n x m),
 * "Kz_chol" (m x m), "alpha" (n); method_fitc also "var_train" (n):
 * gp_pred_post.method_fitc(train = TRUE)$var at the training sites as gp_loo needs it (R/gp_loo.R:123;
 * R/gp_pred.R:245-256,280-286 including that code's jitter-free Dt), computed on demand by one more
 * pass over Kxz; method_rf: "wmean" (M), "wcov" (M x M).  out must hold rows*cols doubles. */
int gpl_fit_get(gpl_ctx* ctx, const gpl_fit* fit, const char* what, double* out);
int gpl_fit_free(gpl_ctx* ctx, gpl_fit* fit);

/* ---- multi-GPU full GP (BASELINE config 5) ---------------------------------------
 * One process per GPU.  The NCCL communicator lives inside the library: the host only has to carry a
 * 128-byte id from rank 0 to the other ranks (any channel: a file, a socket, MPI, torch.distributed)
 * and call gpl_dist_init on every rank.  K is built directly in a 1-D block-cyclic column layout (block
 * columns of width nb, block b on rank b % world), factorised right-looking with panel broadcasts
 * (ncclBroadcast over NVLink) inside a look-ahead window of panel buffers; alpha by distributed block substitution;
 * (K + V)^-1 for the gradient / EP variances by two more panel sweeps.  Every rank passes the SAME
 * host data and receives the SAME results.  world = 1 is allowed (no NCCL needed).
 * NCCL is bound at run time (libnccl.so.2; GPL_NCCL_LIB overrides the path); failures return
 * GPL_ERR_NCCL. */
typedef struct gpl_dist gpl_dist;
#define GPL_DIST_ID_BYTES 128
/* rank 0: fill id128 (ncclGetUniqueId); ship it to every other rank */
int gpl_dist_unique_id(void* id128);
/* collective over all ranks of the job; ctx must be a context on this rank's GPU */
int gpl_dist_init(gpl_ctx* ctx, const void* id128, int32_t rank, int32_t world, gpl_dist** out);
int gpl_dist_destroy(gpl_dist* dist);
/* free the cached n x n / world slabs (kept between calls so that Newton / optimiser iterations reuse them) */
int gpl_dist_release(gpl_dist* dist);
/* diagnostics: average seconds of one ncclBroadcast of `count` doubles on the library's communicator and
 * stream (roots rotate over the ranks), nothing else running */
int gpl_dist_bcast_bench(gpl_dist* dist, int64_t count, int32_t reps, double* seconds);
int gpl_dist_rank(const gpl_dist* dist);
int gpl_dist_world(const gpl_dist* dist);

typedef struct gpl_dist_timings {   /* seconds on this rank (host clock around stream syncs) */
  double kbuild_s;    /* upload, features (first call: allocations) */
  double factor_s;    /* K build in the block-cyclic layout + factorisation */
  double solve_s;     /* alpha, f, log evidence */
  double trtri_s;     /* X = L^-1 sweep (+ diag(B^-1) for EP) */
  double lauum_s;     /* B^-1 = X^T X sweep */
  double trace_s;     /* gradient trace + all-reduce */
  double total_s;
  double nccl_bytes;  /* payload bytes this rank put through NCCL collectives */
} gpl_dist_timings;

/* fit_laplace / fit_ep .method_full across the ranks (R/gp_fit.R:91-99,150-161; R/laplace.R:13-38,
 * 274-297; R/ep.R:16-49,286-330), any likelihood with log-concave sites (GPL_ERR_NOT_LOGCONCAVE
 * otherwise: that branch is single-GPU).  x: n x d host, replicated; outputs as gpl_fit_full (any may
 * be NULL); nb: block width, a multiple of 128 (0 = 1024). */
int gpl_dist_fit_full(gpl_dist* dist, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                      const double* x, int64_t n, int32_t d, const double* y, const double* trials,
                      const double* offset, double jitter, int64_t nb, gpl_fit_out* out, double* fmean,
                      double* z, double* V, double* alpha, double* fvar, double* cavity_mean,
                      double* cavity_var, gpl_dist_timings* timings);
/* gpl_energy_grad across the ranks (lik_gaussian): energy without priors and, when nparam > 0, its
 * gradient w.r.t. the log-parameters (param_factor / param_slot as in gpl_energy_grad).  nparam = 0:
 * energy only (one n x n / world slab per GPU instead of two).  alpha (n, may be NULL). */
int gpl_dist_energy_grad(gpl_dist* dist, const gpl_cf_model* cf, const gpl_lik* lik, const double* x,
                         int64_t n, int32_t d, const double* y, double jitter, int64_t nb, int32_t nparam,
                         const int32_t* param_factor, const int32_t* param_slot, double* energy,
                         double* grad, double* alpha, gpl_dist_timings* timings);

/* fit_laplace / fit_ep .method_fitc across the ranks (SURVEY 8e "FITC: one reduction"; R/gp_fit.R:101-123,
 * 163-188, R/laplace.R:122-152,299-325, R/ep.R:148-185,333-374, R/linear_algebra.R:3-126): rank r works on
 * rows [n r / P, n (r + 1) / P) of x (every rank passes the SAME full host arrays), Kz and the two m x m
 * factorisations are replicated, and per Newton step / EP sweep the data plane is ONE all-reduce of the
 * m x m matrix Kzx S^-1 Kxz plus a few m-vectors and scalars.  Outputs as gpl_fit_fitc, whole n-vectors,
 * identical on every rank; *fit (may be NULL) is a rank-local handle good for gpl_pred_fitc.
 * GPL_ERR_NOT_LOGCONCAVE when a pseudo-variance is <= 0 (that branch is single-GPU). */
int gpl_dist_fit_fitc(gpl_dist* dist, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                      const double* x, int64_t n, int32_t d, const double* zind, int64_t m,
                      const double* y, const double* trials, const double* offset, double jitter,
                      gpl_fit_out* out, double* fmean, double* z, double* V, double* alpha, double* fvar,
                      double* cavity_mean, double* cavity_var, gpl_fit** fit);

#ifdef __cplusplus
}
#endif
#endif /* GPLITE_B200_H */
