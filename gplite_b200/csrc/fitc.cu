// method_fitc (R/gp_fit.R:101-123,163-188; R/laplace.R:122-152,299-325; R/ep.R:148-185,333-374;
// R/linear_algebra.R:3-126; R/gp_pred.R:234-296).
//
// Kxz (n x m; 65.5 GB at n = 2M, m = 4096) is either kept resident (small problems: every pass reads
// it back) or, beyond GPL_FITC_STREAM_GB (default 8), NEVER MATERIALISED: each pass regenerates the
// row chunk it works on from the prepared inputs with the tensor-path K build (cf_gram.cu) -- a chunk
// costs about as much to rebuild as to read, and the fit then needs two m x m matrices, one row-chunk
// workspace and a handful of n-vectors instead of 65.5 GB.  Everything that the
// reference does with further n x m temporaries (forwardsolve(Kz_chol, t(Kxz)), inv_D_U, Kt_aux1 ...)
// goes through that one row-chunk workspace:
//   D_i     = k_ii + jitter - |Lz^-1 k_zi|^2          right-TRSM on a chunk + row sums of squares
//   A       = Kz + Kzx S^-1 Kxz                        chunk scaled by S^-1/2, SYRK-accumulated (DMMA)
//   alpha   = z/S - S^-1 Kxz A^-1 Kzx (z/S)            two GEMVs over Kxz + m-sized solves
//   fvar_i  = V_i D_i / S_i + (V_i/S_i)^2 |L^-1 k_zi|^2   (EP; == diag1 + diag2 of R/ep.R:173-177)
// Prediction needs no n-sized object at all: with beta = Kz^-1 Kzx alpha cached at fit time,
//   mean_t = k_tz beta,  var_t = k_tt - |Lz^-1 k_zt|^2 + |L^-1 k_zt|^2
// which is R/gp_pred.R:262-291 after substituting the inversion lemma.
#include "fit.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

namespace gpl {

namespace {

// K builds of the FITC path (Kz, Kxz chunks, Ktz): see Ctx::cf_direct.  GPL_FITC_DIRECT_K=0 / 1 overrides.
int fitc_cf_pairwise(Ctx* ctx, const CfPlan& plan, const double* z1, int64_t n1, int64_t ldz1, const double* z2,
                     int64_t n2, int64_t ldz2, bool same, double jitter, double* K, int64_t ldk) {
  static const int direct = getenv("GPL_FITC_DIRECT_K") ? atoi(getenv("GPL_FITC_DIRECT_K")) : 0;
  const bool prev = ctx->cf_direct;
  ctx->cf_direct = direct != 0;
  const int rc = cf_pairwise(ctx, plan, z1, n1, ldz1, z2, n2, ldz2, same, jitter, K, ldk);
  ctx->cf_direct = prev;
  return rc;
}


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
    case 4: out[i] = b[i] > 0.0 ? 1.0 / sqrt(a[i]) : 0.0; break; // S^-1/2 on the good sites (V > 0), else 0
    case 5: out[i] = b[i] > 0.0 ? 1.0 : 0.0; break;              // good-site mask
    case 6: out[i] = a[i] * b[i]; break;
  }
}

// gp_pred_post.method_fitc(train = TRUE)$var at the training sites as the reference computes it
// (R/gp_pred.R:245-246,253-256,280-286): its Dt carries no jitter while the fit's D does, so with
// the exact marginal variance v = V D / S + (V / S)^2 w and c = (C^-1)_ii = 1 / S - w / S^2 the
// reference's value is v + jitter - (2 jitter V + jitter^2) c.
__global__ void fitc_train_var_kernel(int64_t n, const double* __restrict__ V,
                                      const double* __restrict__ D, const double* __restrict__ S,
                                      const double* __restrict__ w, double jitter,
                                      double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = V[i] / S[i];
  const double v = r * D[i] + r * r * w[i];
  const double c = 1.0 / S[i] - w[i] / (S[i] * S[i]);
  out[i] = v + jitter - (2.0 * jitter * V[i] + jitter * jitter) * c;
}

__global__ void add_diag_vec_kernel(int64_t n, double* __restrict__ A, int64_t lda,
                                    const double* __restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[i * lda + i] += v[i];
}

__global__ void mirror_lower_to_upper_kernel(int64_t n, double* __restrict__ A, int64_t lda) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i < n && i > j) A[i * lda + j] = A[j * lda + i];
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

// GPL_TIMING=1: per-phase wall times on stderr (each phase fenced by a stream synchronise)
struct FitcTimer {
  Ctx* ctx;
  bool on;
  std::chrono::steady_clock::time_point t0;
  explicit FitcTimer(Ctx* c) : ctx(c) {
    on = getenv("GPL_TIMING") && atoi(getenv("GPL_TIMING")) != 0;
    if (on) {
      cudaStreamSynchronize(ctx->stream);
      t0 = std::chrono::steady_clock::now();
    }
  }
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[gpl timing fitc] %-40s %9.3f ms\n", what,
            std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

struct FitcState {
  // multi-GPU (gpl_dist_fit_fitc): this rank holds rows [row0, row0 + n) of the n_total training points;
  // Kz and the two m x m factors are replicated, the m x m sum and the m-vectors / scalars are all-reduced
  Dist* dist = nullptr;
  int64_t n_total = 0, row0 = 0;
  int64_t n = 0, m = 0, ldx = 0, ldm = 0, chunk = 0, ldw = 0;
  int d = 0;
  DevBuf x, zind, y, trials, offset, Zx, Zz;
  DevBuf Kz;        // m x m with jitter (symmetric, kept)
  DevBuf Lz, Lz_dinv;
  DevBuf A, A_dinv; // m x m work / factor of Kz + Kzx S^-1 Kxz
  DevBuf Kxz;       // n x m (resident mode only)
  bool streamed = false;   // Kxz is regenerated chunk by chunk, never materialised
  DevBuf W;         // chunk x m workspace
  DevBuf Kdiag, Dg, S, rs;     // n
  DevBuf v[14];                // n-vectors
  DevBuf mv[6];                // m-vectors
  HostPlan hp;
  LikDev lik;
  double logdet_z = 0.0;
  double jitter = 0.0;
};

enum { F_F = 0, F_Z, F_V, F_ALPHA, F_FNEW, F_T1, F_T2, F_FVAR, F_P, F_CM, F_CV, F_LOGZ, F_W, F_T3 };

// the first nsum scalar slots are sums over the sites, the next nmax maxima, the next nmin minima: combined
// over the ranks (when sharded) and fetched
int fetch_site_scalars(Ctx* ctx, FitcState& st, int nsum, int nmax, int nmin, double* host) {
  if (st.dist) {
    if (nsum) GPL_TRY(dist_allreduce_f64(st.dist, ctx, ctx->d_scal, nsum, 0));
    if (nmax) GPL_TRY(dist_allreduce_f64(st.dist, ctx, ctx->d_scal + nsum, nmax, 1));
    if (nmin) GPL_TRY(dist_allreduce_f64(st.dist, ctx, ctx->d_scal + nsum + nmax, nmin, 2));
  }
  return fetch_scalars(ctx, nsum + nmax + nmin, host);
}

__global__ void spread_flags_kernel(int* f) {       // f[0] bitmask -> f[1..3] one int per bit
  if (threadIdx.x < 3) f[1 + threadIdx.x] = (f[0] >> threadIdx.x) & 1;
}
__global__ void gather_flags_kernel(int* f) {
  if (threadIdx.x == 0) f[0] = (f[1] ? 1 : 0) | (f[2] ? 2 : 0) | (f[3] ? 4 : 0);
}
// EP site flags (bit 0 quadrature, bit 1 NaN, bit 2 V <= 0): OR over the ranks
int combine_flags(Ctx* ctx, FitcState& st, int* d_flags) {
  if (!st.dist) return GPL_OK;
  spread_flags_kernel<<<1, 32, 0, ctx->stream>>>(d_flags);
  GPL_LAUNCH_CHECK(ctx);
  GPL_TRY(dist_allreduce_i32_max(st.dist, ctx, d_flags + 1, 3));
  gather_flags_kernel<<<1, 32, 0, ctx->stream>>>(d_flags);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

// W <- rows [r0, r0 + c) of Kxz, row i scaled by s[r0 + i] when s != NULL: a scaled copy of the resident
// matrix, or regenerated from the prepared inputs (K(x_chunk, z): R/gp_fit.R:106)
int chunk_into_W(Ctx* ctx, FitcState& st, int64_t r0, int64_t c, const double* s) {
  if (!st.streamed)
    return scaled_copy_rows(ctx, c, st.m, st.Kxz.d() + r0, st.ldx, s ? s + r0 : nullptr, st.W.d(), st.ldw);
  GPL_TRY(fitc_cf_pairwise(ctx, st.hp.plan, st.Zx.d() + r0, c, st.n, st.Zz.d(), st.m, st.m, false, 0.0, st.W.d(),
                      st.ldw));
  if (s) GPL_TRY(scale_rows(ctx, c, st.m, st.W.d(), st.ldw, s + r0));
  return GPL_OK;
}

// y (m) = Kzx x (n)
int kxz_gemv_t(Ctx* ctx, FitcState& st, const double* x, double* y) {
  if (!st.streamed) {
    if (st.n > 0) GPL_TRY(gemv_t(ctx, st.n, st.m, st.Kxz.d(), st.ldx, x, y));
    else GPL_TRY(vec_op(ctx, VOP_FILL, st.m, nullptr, nullptr, y, 0.0));
  } else {
    GPL_TRY(vec_op(ctx, VOP_FILL, st.m, nullptr, nullptr, y, 0.0));
    for (int64_t r0 = 0; r0 < st.n; r0 += st.chunk) {
      const int64_t c = (st.n - r0) < st.chunk ? (st.n - r0) : st.chunk;
      GPL_TRY(chunk_into_W(ctx, st, r0, c, nullptr));
      GPL_TRY(gemv_t(ctx, c, st.m, st.W.d(), st.ldw, x + r0, st.mv[5].d()));
      GPL_TRY(vec_op(ctx, VOP_ADD, st.m, y, st.mv[5].d(), y));
    }
  }
  if (st.dist) GPL_TRY(dist_allreduce_f64(st.dist, ctx, y, st.m, 0));      // sum over the row shards
  return GPL_OK;
}

// y (n) = Kxz x (m)
int kxz_gemv_n(Ctx* ctx, FitcState& st, const double* x, double* y) {
  if (st.n <= 0) return GPL_OK;
  if (!st.streamed) return gemv_n(ctx, st.n, st.m, st.Kxz.d(), st.ldx, x, y);
  for (int64_t r0 = 0; r0 < st.n; r0 += st.chunk) {
    const int64_t c = (st.n - r0) < st.chunk ? (st.n - r0) : st.chunk;
    GPL_TRY(chunk_into_W(ctx, st, r0, c, nullptr));
    GPL_TRY(gemv_n(ctx, c, st.m, st.W.d(), st.ldw, x, y + r0));
  }
  return GPL_OK;
}

// rows-of-(Kxz L^-T) squared norms, streamed through the chunk workspace
int row_norms_via(Ctx* ctx, FitcState& st, const double* L, const double* dinv, double* out) {
  for (int64_t r0 = 0; r0 < st.n; r0 += st.chunk) {
    const int64_t c = (st.n - r0) < st.chunk ? (st.n - r0) : st.chunk;
    GPL_TRY(chunk_into_W(ctx, st, r0, c, nullptr));
    GPL_TRY(dtrsm_right_with(ctx, const_cast<double*>(dinv), true, c, st.m, L, st.ldm, st.W.d(), st.ldw));
    GPL_TRY(rowsumsq(ctx, c, st.m, st.W.d(), st.ldw, out + r0));
  }
  return GPL_OK;
}

int fitc_setup(Ctx* ctx, FitcState& st, const gpl_cf_model* cf, const gpl_lik* lik, const double* x,
               int64_t n_total, int d, const double* zind, int64_t m, const double* y,
               const double* trials, const double* offset, double jitter, Dist* dist) {
  FitcTimer tm(ctx);
  st.dist = dist;
  st.n_total = n_total;
  st.row0 = 0;
  int64_t n = n_total;
  if (dist) {                                      // this rank's rows of x, y, trials, offset
    const int64_t P = dist_world(dist), r = dist_rank(dist);
    st.row0 = n_total * r / P;
    n = n_total * (r + 1) / P - st.row0;
    x += st.row0; y += st.row0;
    if (trials) trials += st.row0;
    if (offset) offset += st.row0;
  }
  st.n = n; st.m = m; st.d = d;
  st.jitter = jitter;
  st.ldx = round_up(n > 0 ? n : 1, 16);
  st.ldm = round_up(m, 16);
  const int64_t na = n > 0 ? n : 1;               // (a rank may own no rows)
  GPL_TRY(st.x.alloc(ctx, sizeof(double) * na * d));
  GPL_TRY(st.zind.alloc(ctx, sizeof(double) * m * d));
  GPL_TRY(st.y.alloc(ctx, sizeof(double) * na));
  if (n > 0) {
    GPL_CUDA(cudaMemcpy2DAsync(st.x.p, sizeof(double) * n, x, sizeof(double) * n_total, sizeof(double) * n, d,
                               cudaMemcpyHostToDevice, ctx->stream));
    GPL_CUDA(cudaMemcpyAsync(st.y.p, y, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  }
  GPL_CUDA(cudaMemcpyAsync(st.zind.p, zind, sizeof(double) * m * d, cudaMemcpyHostToDevice, ctx->stream));
  st.lik.kind = lik->kind; st.lik.link = lik->link; st.lik.param = lik->param;
  st.lik.trials = nullptr; st.lik.offset = nullptr;
  if (trials) {
    GPL_TRY(st.trials.alloc(ctx, sizeof(double) * na));
    if (n > 0) GPL_CUDA(cudaMemcpyAsync(st.trials.p, trials, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    st.lik.trials = st.trials.d();
  }
  if (offset) {
    GPL_TRY(st.offset.alloc(ctx, sizeof(double) * na));
    if (n > 0) GPL_CUDA(cudaMemcpyAsync(st.offset.p, offset, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
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
  {
    const double stream_gb = getenv("GPL_FITC_STREAM_GB") ? atof(getenv("GPL_FITC_STREAM_GB")) : 8.0;
    st.streamed = 8.0 * (double)n * (double)m > stream_gb * 1e9;
  }
  if (!st.streamed) GPL_TRY(st.Kxz.alloc(ctx, sizeof(double) * st.ldx * m));
  // chunk: ~2 GB workspace, multiple of 128 rows
  int64_t chunk = (int64_t)(2.0e9 / (8.0 * (double)m));
  chunk = chunk / 128 * 128;
  if (chunk < 1024) chunk = 1024;
  if (getenv("GPL_FITC_CHUNK")) chunk = round_up(atoll(getenv("GPL_FITC_CHUNK")) > 0 ? atoll(getenv("GPL_FITC_CHUNK")) : 1, 128);
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

  tm.mark("upload + features + allocations");
  // Kz (+ jitter), its factor, Kxz, diag K(x, x) + jitter          R/gp_fit.R:105-110
  GPL_TRY(fitc_cf_pairwise(ctx, st.hp.plan, st.Zz.d(), m, m, nullptr, m, m, true, jitter, st.Kz.d(), st.ldm));
  dim3 g2((unsigned)((m + 255) / 256), (unsigned)m);
  copy_lower_add_kernel<<<g2, 256, 0, ctx->stream>>>(m, st.Kz.d(), st.ldm, st.Lz.d(), st.ldm);
  GPL_LAUNCH_CHECK(ctx);
  int32_t info = 0;
  GPL_TRY(dpotrf_with(ctx, st.Lz_dinv.d(), m, st.Lz.d(), st.ldm, &info, &st.logdet_z));
  if (!st.streamed)
    GPL_TRY(fitc_cf_pairwise(ctx, st.hp.plan, st.Zx.d(), n, n, st.Zz.d(), m, m, false, 0.0, st.Kxz.d(), st.ldx));
  GPL_TRY(cf_diag(ctx, st.hp.plan, st.Zx.d(), n, st.Zx.d(), n, n, jitter, st.Kdiag.d()));
  tm.mark(st.streamed ? "Kz, chol(Kz), diag K (Kxz streamed)" : "Kz, chol(Kz), Kxz (resident), diag K");
  // D = K_diag - colSums(forwardsolve(Kz_chol, t(Kxz))^2)
  GPL_TRY(row_norms_via(ctx, st, st.Lz.d(), st.Lz_dinv.d(), st.v[F_T1].d()));
  GPL_TRY(vec_op(ctx, VOP_SUB, n, st.Kdiag.d(), st.v[F_T1].d(), st.Dg.d()));
  tm.mark("D: chunk -> TRSM -> row norms (m^2 n flops)");
  return GPL_OK;
}

int fitc_alpha_f(Ctx* ctx, FitcState& st, const double* mask);

// inv_lemma_get + alpha + f for the current z (v[F_Z]) and V (v[F_V]);   R/linear_algebra.R:3-43,
// R/laplace.R:139-141.  Leaves: S, factor in A (+A_dinv), alpha, fnew, and scalars
//   slot 0: sum log S, slot 1: 2 sum log diag L (via logdetA), slot 2: f . alpha
int fitc_core(Ctx* ctx, FitcState& st, double* logdetA) {
  const int64_t n = st.n, m = st.m;
  FitcTimer tm(ctx);
  GPL_TRY(vec_op(ctx, VOP_ADD, n, st.Dg.d(), st.v[F_V].d(), st.S.d()));
  GPL_TRY(fvec(ctx, 3, n, st.S.d(), nullptr, nullptr, nullptr, st.rs.d()));
  // A = Kz + Kzx S^-1 Kxz  (lower).  Sharded: every rank sums its rows, rank 0 contributes Kz, ONE all-reduce
  // of the m x m matrix (134 MB at m = 4096) assembles it everywhere.
  dim3 g2((unsigned)((m + 255) / 256), (unsigned)m);
  if (st.dist) GPL_CUDA(cudaMemsetAsync(st.A.p, 0, sizeof(double) * st.ldm * m, ctx->stream));
  if (!st.dist || dist_rank(st.dist) == 0) {
    copy_lower_add_kernel<<<g2, 256, 0, ctx->stream>>>(m, st.Kz.d(), st.ldm, st.A.d(), st.ldm);
    GPL_LAUNCH_CHECK(ctx);
  }
  for (int64_t r0 = 0; r0 < n; r0 += st.chunk) {
    const int64_t c = (n - r0) < st.chunk ? (n - r0) : st.chunk;
    GPL_TRY(chunk_into_W(ctx, st, r0, c, st.rs.d()));
    GPL_TRY(dgemm(ctx, true, false, m, m, c, 1.0, st.W.d(), st.ldw, st.W.d(), st.ldw, 1.0, st.A.d(),
                  st.ldm, true));
  }
  if (st.dist) GPL_TRY(dist_allreduce_f64(st.dist, ctx, st.A.d(), st.ldm * m, 0));
  tm.mark("A = Kz + Kzx S^-1 Kxz: chunk -> SYRK (n m^2 flops)");
  int32_t info = 0;
  GPL_TRY(dpotrf_with(ctx, st.A_dinv.d(), m, st.A.d(), st.ldm, &info, logdetA));
  const int rc = fitc_alpha_f(ctx, st, nullptr);
  tm.mark("chol(A), alpha, f (4 passes over Kxz)");
  return rc;
}

// alpha = C^-1 z by the inversion lemma from the factor in A (any non-zero S), f = Kxz Kz^-1 Kzx alpha
// + D alpha (R/linear_algebra.R:45-126, R/laplace.R:140-141).  mask (0/1 per site, may be NULL)
// restricts the system to a subset of the sites: alpha is zero elsewhere.
int fitc_alpha_f(Ctx* ctx, FitcState& st, const double* mask) {
  const int64_t n = st.n, m = st.m;
  // alpha = z/S - S^-1 Kxz A^-1 Kzx (z/S)
  GPL_TRY(vec_op(ctx, VOP_DIV, n, st.v[F_Z].d(), st.S.d(), st.v[F_T1].d()));
  if (mask) GPL_TRY(vec_op(ctx, VOP_MUL, n, st.v[F_T1].d(), mask, st.v[F_T1].d()));
  GPL_TRY(kxz_gemv_t(ctx, st, st.v[F_T1].d(), st.mv[0].d()));
  GPL_TRY(trsv_lower(ctx, st.A_dinv.d(), false, m, st.A.d(), st.ldm, st.mv[0].d(), st.mv[1].d()));
  GPL_TRY(trsv_lower(ctx, st.A_dinv.d(), true, m, st.A.d(), st.ldm, st.mv[1].d(), st.mv[0].d()));
  GPL_TRY(kxz_gemv_n(ctx, st, st.mv[0].d(), st.v[F_T2].d()));
  GPL_TRY(fvec(ctx, 0, n, st.v[F_T1].d(), st.v[F_T2].d(), st.S.d(), nullptr, st.v[F_ALPHA].d()));
  if (mask) GPL_TRY(vec_op(ctx, VOP_MUL, n, st.v[F_ALPHA].d(), mask, st.v[F_ALPHA].d()));
  // f = Kxz Kz^-1 Kzx alpha + D alpha ;  beta = Kz^-1 Kzx alpha is kept in mv[2]
  GPL_TRY(kxz_gemv_t(ctx, st, st.v[F_ALPHA].d(), st.mv[3].d()));
  GPL_TRY(trsv_lower(ctx, st.Lz_dinv.d(), false, m, st.Lz.d(), st.ldm, st.mv[3].d(), st.mv[4].d()));
  GPL_TRY(trsv_lower(ctx, st.Lz_dinv.d(), true, m, st.Lz.d(), st.ldm, st.mv[4].d(), st.mv[2].d()));
  GPL_TRY(kxz_gemv_n(ctx, st, st.mv[2].d(), st.v[F_T2].d()));
  GPL_TRY(fvec(ctx, 1, n, st.v[F_T2].d(), st.Dg.d(), st.v[F_ALPHA].d(), nullptr, st.v[F_FNEW].d()));
  return GPL_OK;
}

// ---------------------------------------------------------------------------------------------
// Non-log-concave sites with FITC (R/laplace.R:153-232, R/ep.R:186-272): p(y) = p(y_good) p(y_bad |
// y_good), then the fit of all sites by the inversion lemma with S = D + V of mixed sign.
// Everything n-sized is a masked pass over Kxz; the bad block lives in the m-dimensional space of
// U = Kxz Lz^-T (K_21 = U_b U_g^T is never formed):
//   E = sum_good k_i k_i^T / S_i,  A1 = Kz + E = L1 L1^T            (fit of the good sites)
//   K21 C1^-1 K12 = U_b G U_b^T,   G = T - Y^T Y,  T = Lz^-1 E Lz^-T,  Y = L1^-1 E Lz^-T
//   pred_cov = U_b (I - G) U_b^T + diag(D_b),  pred_mean = U_b Lz^-1 Kzx alpha1
//   final factor: A = A1 + K_b^T diag(1 / S_b) K_b                  (rank-nb update, S_b may be < 0)
struct FitcBad {
  int64_t ng = 0, nb = 0;
  BadSmall sm;
};

int fitc_bad_iter(Ctx* ctx, FitcState& st, bool ep, FitcBad& fb, double* log_evidence,
                  double* log_post_unnorm) {
  if (st.streamed) {
    set_error("FITC: pseudo-variance <= 0 and Kxz is streamed; the non-log-concave branch (R/laplace.R:153-232, "
              "R/ep.R:186-272) needs the resident matrix: raise GPL_FITC_STREAM_GB");
    return GPL_ERR_NOT_LOGCONCAVE;
  }
  const int64_t n = st.n, m = st.m, ldm = st.ldm;
  double sc[8];
  // ---- permutation good-first (only n-vectors are permuted)
  std::vector<double> hV(n);
  GPL_CUDA(cudaMemcpyAsync(hV.data(), st.v[F_V].d(), sizeof(double) * n, cudaMemcpyDeviceToHost,
                           ctx->stream));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int32_t> perm(n);
  int64_t ng = 0;
  for (int64_t i = 0; i < n; ++i)
    if (hV[i] > 0.0) perm[ng++] = (int32_t)i;
  int64_t k = ng;
  for (int64_t i = 0; i < n; ++i)
    if (!(hV[i] > 0.0)) perm[k++] = (int32_t)i;
  const int64_t nb = n - ng;
  fb.ng = ng;
  fb.nb = nb;
  if (ng == 0) {
    set_error("all pseudo-variances are <= 0: the posterior is not proper");
    return GPL_ERR_NOT_PD;
  }
  DevBuf dperm, mask, yp, trp, offp, zp, Vp, Sp, f1p, Dp, lz, cvp, cmp;
  GPL_TRY(dperm.alloc(ctx, sizeof(int32_t) * n));
  GPL_CUDA(cudaMemcpyAsync(dperm.p, perm.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  const int32_t* pp = static_cast<const int32_t*>(dperm.p);
  for (DevBuf* b : {&mask, &yp, &zp, &Vp, &Sp, &f1p, &Dp}) GPL_TRY(b->alloc(ctx, sizeof(double) * n));
  GPL_TRY(vec_op(ctx, VOP_ADD, n, st.Dg.d(), st.v[F_V].d(), st.S.d()));
  GPL_TRY(fvec(ctx, 5, n, nullptr, st.v[F_V].d(), nullptr, nullptr, mask.d()));
  GPL_TRY(fvec(ctx, 4, n, st.S.d(), st.v[F_V].d(), nullptr, nullptr, st.rs.d()));
  // ---- fit of the good sites: E, A1 = Kz + E, alpha1, f1
  DevBuf E;
  GPL_TRY(E.alloc(ctx, sizeof(double) * ldm * m));
  for (int64_t r0 = 0; r0 < n; r0 += st.chunk) {
    const int64_t c = (n - r0) < st.chunk ? (n - r0) : st.chunk;
    GPL_TRY(scaled_copy_rows(ctx, c, m, st.Kxz.d() + r0, st.ldx, st.rs.d() + r0, st.W.d(), st.ldw));
    GPL_TRY(dgemm(ctx, true, false, m, m, c, 1.0, st.W.d(), st.ldw, st.W.d(), st.ldw, r0 == 0 ? 0.0 : 1.0,
                  E.d(), ldm, true));
  }
  dim3 g2((unsigned)((m + 255) / 256), (unsigned)m);
  mirror_lower_to_upper_kernel<<<g2, 256, 0, ctx->stream>>>(m, E.d(), ldm);
  GPL_LAUNCH_CHECK(ctx);
  copy_lower_add_kernel<<<g2, 256, 0, ctx->stream>>>(m, st.Kz.d(), ldm, st.A.d(), ldm);
  GPL_LAUNCH_CHECK(ctx);
  // A1 = Kz + E (lower): E is full, adding everywhere is harmless (the factorisation reads the lower part)
  GPL_TRY(vec_op(ctx, VOP_ADD, ldm * m, st.A.d(), E.d(), st.A.d()));
  int32_t info = 0;
  double logdetA1 = 0.0;
  GPL_TRY(dpotrf_with(ctx, st.A_dinv.d(), m, st.A.d(), ldm, &info, &logdetA1));
  GPL_TRY(fitc_alpha_f(ctx, st, mask.d()));                 // alpha1 (zero on bad sites), f1
  // ---- evidence of the good sites
  GPL_TRY(gather_vec(ctx, n, st.y.d(), pp, yp.d()));
  GPL_TRY(gather_vec(ctx, n, st.v[F_Z].d(), pp, zp.d()));
  GPL_TRY(gather_vec(ctx, n, st.v[F_V].d(), pp, Vp.d()));
  GPL_TRY(gather_vec(ctx, n, st.S.d(), pp, Sp.d()));
  GPL_TRY(gather_vec(ctx, n, st.Dg.d(), pp, Dp.d()));
  GPL_TRY(gather_vec(ctx, n, st.v[F_FNEW].d(), pp, f1p.d()));
  LikDev lik_g = st.lik, lik_b = st.lik;
  if (st.lik.trials) {
    GPL_TRY(trp.alloc(ctx, sizeof(double) * n));
    GPL_TRY(gather_vec(ctx, n, st.lik.trials, pp, trp.d()));
    lik_g.trials = trp.d();
    lik_b.trials = trp.d() + ng;
  }
  if (st.lik.offset) {
    GPL_TRY(offp.alloc(ctx, sizeof(double) * n));
    GPL_TRY(gather_vec(ctx, n, st.lik.offset, pp, offp.d()));
    lik_g.offset = offp.d();
    lik_b.offset = offp.d() + ng;
  }
  GPL_TRY(reduce(ctx, RED_SUMLOG, ng, Sp.d(), nullptr, 0));
  GPL_TRY(reduce(ctx, RED_SUMLOG, ng, Vp.d(), nullptr, 1));
  double le1 = 0.0, lpu1 = 0.0;
  if (!ep) {
    GPL_TRY(reduce(ctx, RED_DOT, n, st.v[F_FNEW].d(), st.v[F_ALPHA].d(), 2));      // alpha1 = 0 on bad sites
    GPL_TRY(lik_loglik_sum(ctx, lik_g, ng, f1p.d(), yp.d(), 3));
    GPL_TRY(fetch_scalars(ctx, 4, sc));
    const double C_logdet = sc[0] - st.logdet_z + logdetA1;
    le1 = sc[3] - 0.5 * sc[2] + 0.5 * (sc[1] - C_logdet);
    lpu1 = sc[3] - 0.5 * sc[2];
  } else {
    GPL_TRY(lz.alloc(ctx, sizeof(double) * n));
    GPL_TRY(cvp.alloc(ctx, sizeof(double) * n));
    GPL_TRY(cmp.alloc(ctx, sizeof(double) * n));
    GPL_TRY(gather_vec(ctx, n, st.v[F_LOGZ].d(), pp, lz.d()));
    GPL_TRY(gather_vec(ctx, n, st.v[F_CV].d(), pp, cvp.d()));
    GPL_TRY(gather_vec(ctx, n, st.v[F_CM].d(), pp, cmp.d()));
    GPL_TRY(reduce(ctx, RED_DOT, n, st.v[F_Z].d(), st.v[F_ALPHA].d(), 2));
    GPL_TRY(ep_evidence_terms(ctx, ng, lz.d(), cvp.d(), cmp.d(), zp.d(), Vp.d(), 3));
    GPL_TRY(fetch_scalars(ctx, 4, sc));
    const double C_logdet = sc[0] - st.logdet_z + logdetA1;
    le1 = -0.5 * C_logdet - 0.5 * sc[2] + sc[3];
  }
  // ---- the bad block in U-space
  const int64_t ldb = round_up(nb, 16);
  DevBuf Kb, Ub, R, Rt, Tm, Yt, H, pm, pvec;
  GPL_TRY(Kb.alloc(ctx, sizeof(double) * ldb * m));
  GPL_TRY(Ub.alloc(ctx, sizeof(double) * ldb * m));
  GPL_TRY(H.alloc(ctx, sizeof(double) * ldb * m));
  for (DevBuf* b : {&R, &Rt, &Tm, &Yt}) GPL_TRY(b->alloc(ctx, sizeof(double) * ldm * m));
  GPL_TRY(pm.alloc(ctx, sizeof(double) * nb));
  GPL_TRY(pvec.alloc(ctx, sizeof(double) * m));
  GPL_TRY(gather_rows(ctx, nb, m, st.Kxz.d(), st.ldx, pp + ng, Kb.d(), ldb));
  GPL_CUDA(cudaMemcpyAsync(Ub.d(), Kb.d(), sizeof(double) * ldb * m, cudaMemcpyDeviceToDevice, ctx->stream));
  GPL_TRY(dtrsm_right_with(ctx, st.Lz_dinv.d(), true, nb, m, st.Lz.d(), ldm, Ub.d(), ldb));
  // pred_mean = U_b Lz^-1 (Kzx alpha1)
  GPL_TRY(gemv_t(ctx, n, m, st.Kxz.d(), st.ldx, st.v[F_ALPHA].d(), st.mv[3].d()));
  GPL_TRY(trsv_lower(ctx, st.Lz_dinv.d(), false, m, st.Lz.d(), ldm, st.mv[3].d(), pvec.d()));
  GPL_TRY(gemv_n(ctx, nb, m, Ub.d(), ldb, pvec.d(), pm.d()));
  // R = E Lz^-T, T = Lz^-1 E Lz^-T = R^T Lz^-T, Y^T = R^T L1^-T
  GPL_CUDA(cudaMemcpyAsync(R.d(), E.d(), sizeof(double) * ldm * m, cudaMemcpyDeviceToDevice, ctx->stream));
  GPL_TRY(dtrsm_right_with(ctx, st.Lz_dinv.d(), true, m, m, st.Lz.d(), ldm, R.d(), ldm));
  GPL_TRY(transpose(ctx, m, m, R.d(), ldm, Rt.d(), ldm));
  GPL_CUDA(cudaMemcpyAsync(Tm.d(), Rt.d(), sizeof(double) * ldm * m, cudaMemcpyDeviceToDevice, ctx->stream));
  GPL_TRY(dtrsm_right_with(ctx, st.Lz_dinv.d(), true, m, m, st.Lz.d(), ldm, Tm.d(), ldm));
  GPL_CUDA(cudaMemcpyAsync(Yt.d(), Rt.d(), sizeof(double) * ldm * m, cudaMemcpyDeviceToDevice, ctx->stream));
  GPL_TRY(dtrsm_right_with(ctx, st.A_dinv.d(), true, m, m, st.A.d(), ldm, Yt.d(), ldm));
  // Tm <- I - T + Y^T Y
  GPL_TRY(scale_block(ctx, m, m, Tm.d(), ldm, nullptr, nullptr, -1.0, 1.0, 0));
  GPL_TRY(dgemm(ctx, false, true, m, m, m, 1.0, Yt.d(), ldm, Yt.d(), ldm, 1.0, Tm.d(), ldm, false));
  // pred_cov = U_b (I - G) U_b^T + diag(D_b)
  GPL_TRY(bad_small_alloc(ctx, fb.sm, nb, Vp.d() + ng));
  GPL_TRY(dgemm(ctx, false, false, nb, m, m, 1.0, Ub.d(), ldb, Tm.d(), ldm, 0.0, H.d(), ldb, false));
  GPL_TRY(dgemm(ctx, false, true, nb, nb, m, 1.0, H.d(), ldb, Ub.d(), ldb, 0.0, fb.sm.P.d(), fb.sm.ldp, false));
  add_diag_vec_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(nb, fb.sm.P.d(), fb.sm.ldp,
                                                                             Dp.d() + ng);
  GPL_LAUNCH_CHECK(ctx);
  GPL_TRY(bad_small_factor(ctx, fb.sm));
  if (fb.sm.a_fail) {
    *log_evidence = -INFINITY;
    *log_post_unnorm = -INFINITY;
    return GPL_OK;
  }
  if (!ep) {
    GPL_TRY(bad_small_laplace(ctx, fb.sm, pm.d(), zp.d() + ng, yp.d() + ng, lik_b, 0, 1));
    GPL_TRY(fetch_scalars(ctx, 2, sc));
    const double lpu2 = sc[1] - 0.5 * sc[0];
    *log_post_unnorm = lpu1 + lpu2;
    *log_evidence = le1 + lpu2 + 0.5 * fb.sm.logdetA;                        // R/laplace.R:216-221
  } else {
    GPL_TRY(bad_small_ep(ctx, fb.sm, pm.d(), zp.d() + ng, 0));
    GPL_TRY(ep_evidence_terms_bad(ctx, nb, lz.d() + ng, cvp.d() + ng, cmp.d() + ng, zp.d() + ng,
                                  Vp.d() + ng, 1));
    GPL_TRY(fetch_scalars(ctx, 2, sc));
    *log_evidence = le1 - 0.5 * sc[0] + 0.5 * fb.sm.logdetA + sc[1];         // R/ep.R:249-254
    *log_post_unnorm = 0.0;
  }
  // ---- fit of all sites: A = Kz + E + K_b^T diag(1/S_b) K_b, then alpha, f by the inversion lemma
  copy_lower_add_kernel<<<g2, 256, 0, ctx->stream>>>(m, st.Kz.d(), ldm, st.A.d(), ldm);
  GPL_LAUNCH_CHECK(ctx);
  GPL_TRY(vec_op(ctx, VOP_ADD, ldm * m, st.A.d(), E.d(), st.A.d()));
  DevBuf iSb;
  GPL_TRY(iSb.alloc(ctx, sizeof(double) * nb));
  GPL_TRY(vec_op(ctx, VOP_INV, nb, Sp.d() + ng, nullptr, iSb.d()));
  GPL_CUDA(cudaMemcpyAsync(H.d(), Kb.d(), sizeof(double) * ldb * m, cudaMemcpyDeviceToDevice, ctx->stream));
  GPL_TRY(scale_rows(ctx, nb, m, H.d(), ldb, iSb.d()));
  GPL_TRY(dgemm(ctx, true, false, m, m, nb, 1.0, H.d(), ldb, Kb.d(), ldb, 1.0, st.A.d(), ldm, true));
  double logdetA = 0.0;
  GPL_TRY(dpotrf_with(ctx, st.A_dinv.d(), m, st.A.d(), ldm, &info, &logdetA));
  GPL_TRY(fitc_alpha_f(ctx, st, nullptr));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));            // local buffers are released at scope exit
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
  f->jitter = st.jitter;
  *out = f;
  return GPL_OK;
}

}  // namespace

int fit_fitc_impl(Ctx* ctx, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                  const double* x, int64_t n, int d, const double* zind, int64_t m, const double* y,
                  const double* trials, const double* offset, double jitter, gpl_fit_out* out,
                  double* h_fmean, double* h_z, double* h_V, double* h_alpha, double* h_fvar,
                  double* h_cm, double* h_cv, Fit** fit, Dist* dist) {
  if (n <= 0 || d <= 0 || m <= 0 || !x || !y || !zind || !out) {
    set_error("gpl_fit_fitc: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  if (dist && n < dist_world(dist)) {
    set_error("gpl_dist_fit_fitc: fewer training points than ranks");
    return GPL_ERR_BAD_ARG;
  }
  if (m > 65535) {          // the m-indexed helper kernels map inducing points to gridDim.y
    set_error("gpl_fit_fitc: at most 65535 inducing points (got %lld)", (long long)m);
    return GPL_ERR_BAD_ARG;
  }
  FitcState st;
  GPL_TRY(fitc_setup(ctx, st, cf, lik, x, n, d, zind, m, y, trials, offset, jitter, dist));
  const int64_t n_total = n;
  n = st.n;                                    // from here on: this rank's rows
  memset(out, 0, sizeof(*out));
  out->quad_ok = 1;
  double sc[8];
  DevBuf gather;
  if (dist) GPL_TRY(gather.alloc(ctx, sizeof(double) * n_total));
  // n-vectors go back whole on every rank: each rank drops its rows into a zeroed n_total-vector, one sum
  auto dl = [&](double* h, const double* dptr) -> int {
    if (!h) return GPL_OK;
    if (!dist) {
      GPL_CUDA(cudaMemcpyAsync(h, dptr, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
      return GPL_OK;
    }
    GPL_CUDA(cudaMemsetAsync(gather.p, 0, sizeof(double) * n_total, ctx->stream));
    GPL_CUDA(cudaMemcpyAsync(gather.d() + st.row0, dptr, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    GPL_TRY(dist_allreduce_f64(dist, ctx, gather.d(), n_total, 0));
    GPL_CUDA(cudaMemcpyAsync(h, gather.p, sizeof(double) * n_total, cudaMemcpyDeviceToHost, ctx->stream));
    return GPL_OK;
  };

  if (approx->kind == GPL_APPROX_LAPLACE) {                                  // R/laplace.R:299-325
    int maxiter = approx->maxiter;
    if (lik->kind == GPL_LIK_GAUSSIAN) maxiter = 1;
    if (maxiter < 1) maxiter = 1;
    GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, st.v[F_F].d(), 0.0));
    double log_post_old = -INFINITY, log_evidence = 0.0;
    int it = 0, n_bad = 0;
    for (it = 1; it <= maxiter; ++it) {
      GPL_TRY(lik_pseudodata_la(ctx, st.lik, n, st.v[F_F].d(), st.y.d(), st.v[F_Z].d(), st.v[F_V].d()));
      GPL_TRY(reduce(ctx, RED_MIN, n, st.v[F_V].d(), nullptr, 0));
      GPL_TRY(fetch_site_scalars(ctx, st, 0, 0, 1, sc));
      if (!(sc[0] > 0.0) && dist) {
        set_error("pseudo-variance <= 0: the non-log-concave branch (R/laplace.R:153-232) is single-GPU only");
        return GPL_ERR_NOT_LOGCONCAVE;
      }
      if (!(sc[0] > 0.0)) {                                   // R/laplace.R:153-232
        FitcBad fb;
        double log_post = 0.0;
        GPL_TRY(fitc_bad_iter(ctx, st, false, fb, &log_evidence, &log_post));
        n_bad = (int)fb.nb;
        if (fb.sm.a_fail) {                                   // list(fmean = fhat_old, log_evidence = -Inf)
          out->log_evidence = -INFINITY;
          out->iters = it;
          set_error("Laplace (FITC): the posterior with non-log-concave sites is not proper");
          return GPL_ERR_NONFINITE;
        }
        std::swap(st.v[F_F].p, st.v[F_FNEW].p);
        const double diff = log_post - log_post_old;
        log_post_old = log_post;
        if (fabs(diff) < approx->tol) break;
        continue;
      }
      n_bad = 0;
      double logdetA = 0.0;
      GPL_TRY(fitc_core(ctx, st, &logdetA));
      GPL_TRY(reduce(ctx, RED_SUMLOG, n, st.S.d(), nullptr, 0));
      GPL_TRY(reduce(ctx, RED_SUMLOG, n, st.v[F_V].d(), nullptr, 1));
      GPL_TRY(reduce(ctx, RED_DOT, n, st.v[F_FNEW].d(), st.v[F_ALPHA].d(), 2));
      GPL_TRY(lik_loglik_sum(ctx, st.lik, n, st.v[F_FNEW].d(), st.y.d(), 3));
      GPL_TRY(fetch_site_scalars(ctx, st, 4, 0, 0, sc));
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
    out->accepted = it;
    out->converged = !(maxiter > 1 && it == maxiter);
    out->n_bad = n_bad;
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
  int it = 0, quad_ok = 1, n_bad = 0, accepted = 0;
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
    GPL_TRY(combine_flags(ctx, st, static_cast<int*>(flags.p)));
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
    if ((hflags & 4) && dist) {
      set_error("pseudo-variance <= 0: the non-log-concave branch (R/ep.R:186-272) is single-GPU only");
      return GPL_ERR_NOT_LOGCONCAVE;
    }
    if (hflags & 4) {                                                        // R/ep.R:186-272
      FitcBad fb;
      double unused = 0.0;
      GPL_TRY(fitc_bad_iter(ctx, st, true, fb, &log_evidence, &unused));
      n_bad = (int)fb.nb;
      if (fb.sm.a_fail) {
        out->log_evidence = -INFINITY;
        out->iters = it;
        out->quad_ok = quad_ok;
        set_error("EP (FITC): the posterior with non-log-concave sites is not proper");
        return GPL_ERR_NONFINITE;
      }
      GPL_TRY(row_norms_via(ctx, st, st.A.d(), st.A_dinv.d(), st.v[F_W].d()));
      GPL_TRY(fvec(ctx, 2, n, st.v[F_V].d(), st.Dg.d(), st.S.d(), st.v[F_W].d(), st.v[F_FVAR].d()));
      GPL_TRY(reduce(ctx, RED_MAXABSDIFF, n, fmean_old.d(), st.v[F_FNEW].d(), 0));
      GPL_TRY(fetch_scalars(ctx, 1, sc));
      diff_new = sc[0];
    } else {
      n_bad = 0;
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
      GPL_TRY(fetch_site_scalars(ctx, st, 3, 1, 0, sc));
      const double C_logdet = sc[0] - st.logdet_z + logdetA;
      log_evidence = -0.5 * C_logdet - 0.5 * sc[1] + sc[2];                  // R/ep.R:182-184
      diff_new = sc[3];
    }
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
    ++accepted;
    if (diff_new < approx->tol) break;
  }
  if (it > maxiter) it = maxiter;
  out->accepted = accepted;
  out->damping_final = damp;
  out->iters = it;
  out->converged = !(maxiter > 1 && it == maxiter);
  out->quad_ok = quad_ok;
  if (maxiter > 1 && it == maxiter) log_evidence = -INFINITY;
  if (!quad_ok) log_evidence = -INFINITY;
  out->log_evidence = log_evidence;
  out->n_bad = n_bad;
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
    GPL_TRY(fitc_cf_pairwise(ctx, hp.plan, Zt.d() + r0, c, nt, Zz.d(), m, m, false, 0.0, Ktz.d(), ldk));
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

// gp_pred_post.method_fitc(train = TRUE)$var at the training sites, on demand (gp_loo with the
// Laplace approximation, R/gp_loo.R:123): one more streamed pass over Kxz (n m^2 flops), so it is
// not part of the fit.  out: n doubles on the host.
int fitc_var_train(Ctx* ctx, const Fit* fit, double* out) {
  const int64_t n = fit->n, m = fit->m;
  if (!fit->Kxz.p) {
    set_error("var_train: this fit streamed Kxz (never materialised); refit with GPL_FITC_STREAM_GB raised");
    return GPL_ERR_BAD_ARG;
  }
  int64_t chunk = (int64_t)(2.0e9 / (8.0 * (double)m));
  chunk = chunk / 128 * 128;
  if (chunk < 1024) chunk = 1024;
  if (chunk > n) chunk = n;
  const int64_t ldw = round_up(chunk, 16);
  DevBuf W, w, V, res;
  GPL_TRY(W.alloc(ctx, sizeof(double) * ldw * m));
  GPL_TRY(w.alloc(ctx, sizeof(double) * n));
  GPL_TRY(V.alloc(ctx, sizeof(double) * n));
  GPL_TRY(res.alloc(ctx, sizeof(double) * n));
  for (int64_t r0 = 0; r0 < n; r0 += chunk) {
    const int64_t c = (n - r0) < chunk ? (n - r0) : chunk;
    GPL_TRY(scaled_copy_rows(ctx, c, m, fit->Kxz.d() + r0, fit->ldkxz, nullptr, W.d(), ldw));
    GPL_TRY(dtrsm_right_with(ctx, fit->Lm_dinv.d(), true, c, m, fit->Lm.d(), fit->ldm, W.d(), ldw));
    GPL_TRY(rowsumsq(ctx, c, m, W.d(), ldw, w.d() + r0));
  }
  GPL_TRY(vec_op(ctx, VOP_SUB, n, fit->S.d(), fit->Dg.d(), V.d()));          // V = S - D
  fitc_train_var_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
      n, V.d(), fit->Dg.d(), fit->S.d(), w.d(), fit->jitter, res.d());
  GPL_LAUNCH_CHECK(ctx);
  GPL_CUDA(cudaMemcpyAsync(out, res.d(), sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPL_OK;
}

// Joint prediction for method_fitc (gp_pred_post.method_fitc with cov = TRUE, R/gp_pred.R:262-273):
// substituting the inversion lemma, prior_cov - Kt_aux2^T (Kt_aux1 C^-1 Kt_aux1^T) Kt_aux2 collapses to
//   cov = diag(k_tt - |Lz^-1 k_zt|^2) + (Ktz L^-T)(Ktz L^-T)^T + jitter I,   L L^T = Kz + Kzx S^-1 Kxz
// (the same identity as the variances above).  d_mean: nt, d_cov: nt x nt (ld ldc), on the device.
int pred_fitc_cov_dev(Ctx* ctx, const Fit* fit, const gpl_cf_model* cf, const double* xt, int64_t nt,
                      uint64_t term_mask, double jitter, double* d_mean, double* d_cov, int64_t ldc) {
  const int64_t m = fit->m;
  const int d = fit->d;
  if ((double)nt * (double)m * 8.0 > 48e9) {
    set_error("joint prediction of %lld points with m = %lld is too large; predict fewer points at a time",
              (long long)nt, (long long)m);
    return GPL_ERR_NOMEM;
  }
  HostPlan hp;
  GPL_TRY(make_plan(cf, d, term_mask, &hp));
  const int F = hp.plan.nfeat > 0 ? hp.plan.nfeat : 1;
  DevBuf dxt, Zz, Zt, Ktz, W, dt, t1;
  const int64_t ldk = round_up(nt, 16);
  GPL_TRY(dxt.alloc(ctx, sizeof(double) * nt * d));
  GPL_CUDA(cudaMemcpyAsync(dxt.p, xt, sizeof(double) * nt * d, cudaMemcpyHostToDevice, ctx->stream));
  GPL_TRY(Zz.alloc(ctx, sizeof(double) * m * F));
  GPL_TRY(Zt.alloc(ctx, sizeof(double) * nt * F));
  GPL_TRY(prep_features(ctx, hp, fit->zind.d(), m, m, Zz.d(), m));
  GPL_TRY(prep_features(ctx, hp, dxt.d(), nt, nt, Zt.d(), nt));
  GPL_TRY(Ktz.alloc(ctx, sizeof(double) * ldk * m));
  GPL_TRY(W.alloc(ctx, sizeof(double) * ldk * m));
  GPL_TRY(dt.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(t1.alloc(ctx, sizeof(double) * nt));
  GPL_TRY(fitc_cf_pairwise(ctx, hp.plan, Zt.d(), nt, nt, Zz.d(), m, m, false, 0.0, Ktz.d(), ldk));
  GPL_TRY(gemv_n(ctx, nt, m, Ktz.d(), ldk, fit->sqrtV.d(), d_mean));          // beta rides in sqrtV
  GPL_TRY(cf_diag(ctx, hp.plan, Zt.d(), nt, Zt.d(), nt, nt, 0.0, dt.d()));
  GPL_TRY(scaled_copy_rows(ctx, nt, m, Ktz.d(), ldk, nullptr, W.d(), ldk));
  GPL_TRY(dtrsm_right_with(ctx, fit->Kz_dinv.d(), true, nt, m, fit->Kz_chol.d(), fit->ldm, W.d(), ldk));
  GPL_TRY(rowsumsq(ctx, nt, m, W.d(), ldk, t1.d()));
  GPL_TRY(vec_op(ctx, VOP_SUB, nt, dt.d(), t1.d(), dt.d()));                   // Dt
  GPL_TRY(dtrsm_right_with(ctx, fit->Lm_dinv.d(), true, nt, m, fit->Lm.d(), fit->ldm, Ktz.d(), ldk));
  GPL_TRY(dgemm(ctx, false, true, nt, nt, m, 1.0, Ktz.d(), ldk, Ktz.d(), ldk, 0.0, d_cov, ldc, false));
  GPL_TRY(vec_op(ctx, VOP_FILL, nt, nullptr, nullptr, t1.d(), jitter));
  GPL_TRY(vec_op(ctx, VOP_ADD, nt, dt.d(), t1.d(), dt.d()));
  add_diag_vec_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(nt, d_cov, ldc, dt.d());
  GPL_LAUNCH_CHECK(ctx);
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));     // local buffers are released at scope exit
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
                         fit ? &f : nullptr, nullptr);
  if (fit) *fit = reinterpret_cast<gpl_fit*>(f);
  if (rc != GPL_OK) cudaStreamSynchronize(reinterpret_cast<Ctx*>(h)->stream);
  return rc;
}

int gpl_dist_fit_fitc(gpl_dist* dh, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                      const double* x, int64_t n, int32_t d, const double* zind, int64_t m,
                      const double* y, const double* trials, const double* offset, double jitter,
                      gpl_fit_out* out, double* fmean, double* z, double* V, double* alpha, double* fvar,
                      double* cavity_mean, double* cavity_var, gpl_fit** fit) {
  Dist* dist = reinterpret_cast<Dist*>(dh);
  if (!dist) {
    set_error("gpl_dist_fit_fitc: no communicator");
    return GPL_ERR_BAD_ARG;
  }
  Ctx* ctx = dist_ctx(dist);
  Fit* f = nullptr;
  if (fit) *fit = nullptr;
  int rc = fit_fitc_impl(ctx, cf, lik, approx, x, n, d, zind, m, y, trials, offset, jitter, out, fmean, z, V,
                         alpha, fvar, cavity_mean, cavity_var, fit ? &f : nullptr, dist_world(dist) > 1 ? dist : nullptr);
  if (fit) *fit = reinterpret_cast<gpl_fit*>(f);
  if (rc != GPL_OK) cudaStreamSynchronize(ctx->stream);
  return rc;
}

int gpl_pred_fitc(gpl_ctx* h, const gpl_fit* fit, const gpl_cf_model* cf, const double* xt,
                  int64_t nt, uint64_t term_mask, double* mean, double* var) {
  return pred_fitc_impl(reinterpret_cast<Ctx*>(h), reinterpret_cast<const Fit*>(fit), cf, xt, nt,
                        term_mask, mean, var);
}

}  // extern "C"
