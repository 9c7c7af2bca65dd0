// Vector / elementwise / likelihood kernels and the device-resident fit object.
#pragma once

#include "cf.cuh"

#include <vector>

namespace gpl {

// chol.cu (explicit block-inverse buffer variants)
int dpotrf_with(Ctx* ctx, double* dinv, int64_t n, double* A, int64_t lda, int32_t* info_host,
                double* logdet_host);
int dtrsm_right_with(Ctx* ctx, double* dinv, bool trans, int64_t m, int64_t n, const double* L,
                     int64_t ldl, double* B, int64_t ldb);
int dpotri_with(Ctx* ctx, double* dinv, int64_t n, double* A, int64_t lda);
int dtrtri_with(Ctx* ctx, double* dinv, int64_t n, double* A, int64_t lda);

// RAII device buffer on the context's stream-ordered pool
struct DevBuf {
  Ctx* ctx = nullptr;
  void* p = nullptr;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  int alloc(Ctx* c, size_t bytes);
  void release();
  double* d() const { return static_cast<double*>(p); }
};

// ---- vec_kernels.cu -------------------------------------------------------------------------
// B_ij = delta_ij + s_i (K_ij s_j), lower triangle (R/laplace.R:27)
int build_B(Ctx* ctx, int64_t n, const double* K, int64_t ldk, const double* s, double* B,
            int64_t ldb);
// y = A^T x (A is m x n, y has n entries); for symmetric A this is A x
int gemv_t(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const double* x, double* y);
// y = A x  (A is m x n)
int gemv_n(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const double* x, double* y);
// x <- L^-1 b  /  x <- L^-T b using the factor's block inverses; b is destroyed, x != b
int trsv_lower(Ctx* ctx, const double* dinv, bool trans, int64_t n, const double* L, int64_t ldl,
               double* b, double* x);
// out_j = sum_{i >= j} X_ij^2  (column sums of squares of a lower-triangular matrix)
int lower_colsumsq(Ctx* ctx, int64_t n, const double* X, int64_t ldx, double* out);
// out_i = sum_j A_ij^2 (row sums of squares, A m x n)
int rowsumsq(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, double* out);
// A[:, j] *= s_j
int scale_cols(Ctx* ctx, int64_t m, int64_t n, double* A, int64_t lda, const double* s);
// A[i, :] *= s_i
int scale_rows(Ctx* ctx, int64_t m, int64_t n, double* A, int64_t lda, const double* s);
// A[i, j] <- alpha r_i A[i, j] c_j (+ diag_add where i == j + diag_offset)
int scale_block(Ctx* ctx, int64_t m, int64_t n, double* A, int64_t lda, const double* r,
                const double* c, double alpha, double diag_add, int64_t diag_offset);

// ---- helpers of the non-log-concave ("bad site") branch (R/laplace.R:40-117, R/ep.R:50-138) ----
int gather_vec(Ctx* ctx, int64_t n, const double* src, const int32_t* perm, double* dst);   // dst_i = src[perm_i]
int scatter_vec(Ctx* ctx, int64_t n, const double* src, const int32_t* perm, double* dst);  // dst[perm_i] = src_i
int gather_rows(Ctx* ctx, int64_t n, int64_t ncols, const double* src, int64_t lds,
                const int32_t* perm, double* dst, int64_t ldd);
int transpose(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, double* B, int64_t ldb);
// out = tril(L) with rows scaled by r (r may be null)
int tril_scale(Ctx* ctx, int64_t n, const double* L, int64_t ldl, const double* r, double* out,
               int64_t ldo);
// out_i = sum_j A_ij^2 w_j
int rowsumsq_w(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const double* w,
               double* out);

// out_j = sum_i A_ij^2
int colsumsq(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, double* out);
// out_i = sum_j A_ij B_ij
int rowdot(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const double* B, int64_t ldb,
           double* out);

enum VecOp {
  VOP_DIV = 0,        // out = a / b
  VOP_MUL = 1,        // out = a * b
  VOP_RSQRT_INV = 2,  // out = sqrt(1 / a)
  VOP_SQRT = 3,       // out = sqrt(a)
  VOP_FILL = 4,       // out = alpha
  VOP_INV = 5,        // out = 1 / a
  VOP_SUB = 6,        // out = a - b
  VOP_ADD = 7,        // out = a + b
  VOP_AXPBY = 8,      // out = alpha a + beta b
  VOP_COPY = 9
};
int vec_op(Ctx* ctx, VecOp op, int64_t n, const double* a, const double* b, double* out,
           double alpha = 0.0, double beta = 0.0);

enum RedOp {
  RED_SUM = 0,        // sum a
  RED_DOT = 1,        // sum a*b
  RED_SUMLOG = 2,     // sum log a
  RED_MAXABSDIFF = 3, // max |a - b|
  RED_MIN = 4,        // min a
  RED_COUNT_NAN = 5,  // number of NaNs in a (and b if given)
  RED_SUMSQ = 6,      // sum a^2
  RED_SUMLOG_PROD = 7 // sum log(a*b)
};
// result lands in ctx->d_scal[slot]; fetch_scalars() copies the first `count` slots to the host
int reduce(Ctx* ctx, RedOp op, int64_t n, const double* a, const double* b, int slot);
int fetch_scalars(Ctx* ctx, int count, double* host_out);

// ---- likelihood kernels (R/lik.R) -------------------------------------------------------------
struct LikDev {
  int kind, link;
  double param;
  const double* trials;   // device, may be null
  const double* offset;   // device, may be null
};
// z, V from f, y  (get_pseudodata_la, R/lik.R:188-201)
int lik_pseudodata_la(Ctx* ctx, const LikDev& lik, int64_t n, const double* f, const double* y,
                      double* z, double* V);
// sum_i loglik(f_i, y_i) into d_scal[slot] (get_loglik, R/lik.R:297-366)
int lik_loglik_sum(Ctx* ctx, const LikDev& lik, int64_t n, const double* f, const double* y,
                   int slot);
// EP site update (get_pseudodata_ep + get_tilted_moments, R/lik.R:207-270)
struct EpSiteArgs {
  const double* fmean; const double* fvar; const double* z_old; const double* P_old;
  const double* y;
  double* z_new; double* V_new; double* P_new;
  double* cav_mean; double* cav_var; double* logZ;
  double damping;
  int quad_order;
  const double* quad_x;   // device
  const double* quad_w;   // device
};
int lik_ep_sites(Ctx* ctx, const LikDev& lik, int64_t n, const EpSiteArgs& a, int* d_flags);
// EP evidence terms: sum logZ + 0.5 sum log(vc + V) + 0.5 sum (mc - z)^2 / (vc + V) -> slot
int ep_evidence_terms(Ctx* ctx, int64_t n, const double* logZ, const double* vc, const double* mc,
                      const double* z, const double* V, int slot);

// same with log((vc + V) / V) in the middle term (bad sites, R/ep.R:112-114)
int ep_evidence_terms_bad(Ctx* ctx, int64_t n, const double* logZ, const double* vc,
                          const double* mc, const double* z, const double* V, int slot);

// ---- gradient trace (new capability; see gpl_energy_grad) -----------------------------------
struct GradParam {
  int factor;   // index into plan.f, or -1 for the likelihood sigma (handled outside)
  int slot;     // which parameter of the factor: 0.. as in gpl_cf_factor.p ; 3 = period
};
constexpr int MAX_GRAD_PARAMS = 24;
// out[j] = sum_{i,k} (alpha_i alpha_k - s_i Binv_ik s_k) dK_ik/dlog(theta_j), Binv lower triangle
int grad_trace(Ctx* ctx, const CfPlan& plan, const double* z, int64_t ldz, const double* zaux,
               int64_t n, const double* Binv, int64_t ldb, const double* s, const double* alpha,
               int nparam, const GradParam* params, double* d_out /* device, nparam doubles */,
               int64_t col0 = -1, int64_t ncols = 0, const double* u = nullptr);
// u != NULL adds - u_i u_k to the weight (the implicit term of the Laplace gradient, where alpha
// holds a + u: (a+u)(a+u)^T - u u^T = a a^T + a u^T + u a^T).
// col0 >= 0: block mode -- Binv holds only the block column [col0, col0 + ncols) (all n rows, rows
// above the block's diagonal are never read); the result is that block column's share of the sum.

// ---- nb x nb algebra of the non-log-concave branch (R/laplace.R:66-100, R/ep.R:81-114), shared by
// method_full (fit.cu) and method_fitc (fitc.cu) ------------------------------------------------
struct BadSmall {
  int64_t nb = 0, ldp = 0;
  bool a_fail = false;     // A not positive definite: improper posterior (log_evidence = -Inf)
  double logdetA = 0.0;
  DevBuf P, Pd;            // pred_cov -> its factor (+ block inverses)
  DevBuf Lk, T;            // clean lower factor of pred_cov; scratch
  DevBuf A, Ad;            // factor of A = I + Lk^T V_b^-1 Lk
  DevBuf invVb;
  DevBuf b[8];             // nb-sized scratch
};
int bad_small_alloc(Ctx* ctx, BadSmall& sm, int64_t nb, const double* V_b);
int bad_small_factor(Ctx* ctx, BadSmall& sm);
int bad_small_solve_S(Ctx* ctx, BadSmall& sm, const double* v, double* out);
int bad_small_laplace(Ctx* ctx, BadSmall& sm, const double* pm, const double* z_b, const double* y_b,
                      const LikDev& lik_b, int slot_df2, int slot_ll);
int bad_small_ep(Ctx* ctx, BadSmall& sm, const double* pm, const double* z_b, int slot_df2);

// ---- fit object --------------------------------------------------------------------------------
struct Fit {
  int method = 0;        // 0 full, 1 fitc, 2 rf (m = basis functions, Lm = wcov, alpha = wmean,
                         //                        zind = random weights)
  int64_t n = 0, m = 0;
  int d = 0;
  int64_t ld = 0;        // leading dimension of Lb
  DevBuf x;              // n x d training inputs (gp$x)
  DevBuf Lb;             // factor of B = I + V^-1/2 K V^-1/2  (C_chol = diag(sqrtV) Lb)
  DevBuf dinv;           // its diagonal-block inverses
  DevBuf sqrtV;          // n
  DevBuf alpha;          // n
  // FITC
  DevBuf zind;           // m x d
  DevBuf Kxz;            // n x m (ld = ldx_)
  int64_t ldkxz = 0;
  DevBuf Kz_chol, Kz_dinv;   // m x m factor of Kz (+ jitter)
  DevBuf Lm, Lm_dinv;        // m x m factor of Kz + Kzx S^-1 Kxz  (inv_lemma chol)
  int64_t ldm = 0;
  DevBuf S;              // D + V
  DevBuf Dg;             // FITC diagonal correction D
  // Non-log-concave fit (some pseudo-variances <= 0; the reference keeps an LU of K + diag(V),
  // R/laplace.R:113, R/ep.R:122).  Here: sites permuted good-first; Lb/dinv/sqrtV factor the good
  // block (size ng); the bad block is eliminated through its nb x nb Schur complement
  // S = pred_cov + V_bad with pred_cov = Lk Lk^T and A = I + Lk^T V_bad^-1 Lk = Ac Ac^T.
  int64_t ng = 0, nb = 0;
  DevBuf perm;           // int32[n]: permuted position -> original index
  DevBuf xp;             // n x d inputs, permuted
  DevBuf alpha_p;        // n, permuted
  DevBuf W;              // nb x ng: K_bg C_chol_gg^-T
  int64_t ldw = 0;
  DevBuf Lk, Lk_dinv;    // nb x nb (clean lower triangle), ld ldp
  DevBuf Ac, Ac_dinv;    // nb x nb factor of A
  int64_t ldp = 0;
  DevBuf invVb;          // nb: 1 / V_bad
  double jitter = 0.0;   // FITC: the jitter of the fit (fitc_var_train)
  // method_rf
  std::vector<int32_t> rf_num_basis;
  int64_t rf_nweights = 0;
};

int fitc_var_train(Ctx* ctx, const Fit* fit, double* out);

// ---- dist.cu: the communicator object behind gpl_dist, as seen by the row-sharded FITC fit ----------
struct Dist;
Ctx* dist_ctx(Dist* d);
int dist_rank(const Dist* d);
int dist_world(const Dist* d);
// in-place all-reduce of device doubles on ctx->stream: op 0 sum, 1 max, 2 min
int dist_allreduce_f64(Dist* d, Ctx* ctx, double* buf, int64_t count, int op);
int dist_allreduce_i32_max(Dist* d, Ctx* ctx, int* buf, int64_t count);
// joint predictions on the device (mean nt, cov nt x nt with jitter on the diagonal)
int pred_full_cov_dev(Ctx* ctx, const Fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                      uint64_t term_mask, double jitter, double* d_mean, double* d_cov, int64_t ldc);
int pred_fitc_cov_dev(Ctx* ctx, const Fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                      uint64_t term_mask, double jitter, double* d_mean, double* d_cov, int64_t ldc);

}  // namespace gpl
