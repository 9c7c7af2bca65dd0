// HBM-bound vector / matrix-vector / elementwise kernels of the fit loops, the likelihood maths
// (R/lik.R) and the EP site sweep.  All of these are single-pass, coalesced along the
// column-major leading dimension, and leave their scalar results in ctx->d_scal so that one
// small D2H copy per Newton iteration / EP sweep carries every convergence quantity.
#include "fit.cuh"

#include <math.h>

namespace gpl {

int DevBuf::alloc(Ctx* c, size_t bytes) {
  release();
  ctx = c;
  cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 16, c->stream);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    p = nullptr;
    set_error("device allocation of %zu bytes failed", bytes);
    return GPL_ERR_NOMEM;
  }
  GPL_CUDA(e);
  return GPL_OK;
}

void DevBuf::release() {
  if (p && ctx) cudaFreeAsync(p, ctx->stream);
  p = nullptr;
}

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

__global__ void build_B_kernel(int64_t n, const double* __restrict__ K, int64_t ldk,
                               const double* __restrict__ s, double* __restrict__ B, int64_t ldb) {
  const int64_t j0 = (int64_t)blockIdx.y * 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((int64_t)(blockIdx.x + 1) * blockDim.x <= j0) return;   // block entirely above the diagonal
  if (i >= n) return;
  const double si = s[i];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int64_t j = j0 + c;
    if (j < n && i >= j) {
      const double v = si * (K[j * ldk + i] * s[j]);
      B[j * ldb + i] = (i == j) ? 1.0 + v : v;
    }
  }
}

__global__ void gemv_t_kernel(int64_t m, int64_t n, const double* __restrict__ A, int64_t lda,
                              const double* __restrict__ x, double* __restrict__ y, double alpha,
                              double beta) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  const double* col = A + j * lda;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int64_t i = lane;
  for (; i + 96 < m; i += 128) {
    s0 = fma(col[i], x[i], s0);
    s1 = fma(col[i + 32], x[i + 32], s1);
    s2 = fma(col[i + 64], x[i + 64], s2);
    s3 = fma(col[i + 96], x[i + 96], s3);
  }
  for (; i < m; i += 32) s0 = fma(col[i], x[i], s0);
  const double s = warp_sum((s0 + s1) + (s2 + s3));
  if (lane == 0) y[j] = (beta != 0.0) ? beta * y[j] + alpha * s : alpha * s;
}

// partial[split][m]: row-wise reduction over a column range. MODE 0: sum A_ij x_j, 1: sum A_ij^2
template <int MODE>
__global__ void row_reduce_kernel(int64_t m, int64_t n, const double* __restrict__ A, int64_t lda,
                                  const double* __restrict__ x, double* __restrict__ partial,
                                  int64_t cols_per_split) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int64_t j0 = (int64_t)blockIdx.y * cols_per_split;
  int64_t j1 = j0 + cols_per_split;
  if (j1 > n) j1 = n;
  double s0 = 0.0, s1 = 0.0;
  int64_t j = j0;
  for (; j + 1 < j1; j += 2) {
    const double a0 = A[j * lda + i], a1 = A[(j + 1) * lda + i];
    if (MODE == 0) {
      s0 = fma(a0, x[j], s0);
      s1 = fma(a1, x[j + 1], s1);
    } else {
      s0 = fma(a0, a0, s0);
      s1 = fma(a1, a1, s1);
    }
  }
  if (j < j1) {
    const double a0 = A[j * lda + i];
    s0 = (MODE == 0) ? fma(a0, x[j], s0) : fma(a0, a0, s0);
  }
  partial[(int64_t)blockIdx.y * m + i] = s0 + s1;
}

__global__ void sum_partials_kernel(int64_t m, int nsplit, const double* __restrict__ partial,
                                    double* __restrict__ y, double alpha, double beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double s = 0.0;
  for (int k = 0; k < nsplit; ++k) s += partial[(int64_t)k * m + i];
  y[i] = (beta != 0.0) ? beta * y[i] + alpha * s : alpha * s;
}

// x <- Linv b (trans = 0) or Linv^T b (trans = 1) for one <= 128 base block, in place.  Linv is the
// full 128 x 128 tile with a zero upper triangle, so fixed trip counts keep the loads pipelined.
__global__ void trsv_block_kernel(int nbk, const double* __restrict__ Linv, double* __restrict__ b,
                                  int trans) {
  __shared__ double bk[NB];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < NB) bk[tid] = (tid < nbk) ? b[tid] : 0.0;
  __syncthreads();
  if (!trans) {
    if (tid < NB) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 8
      for (int c = 0; c < NB; c += 4) {
        s0 = fma(Linv[(c + 0) * NB + tid], bk[c + 0], s0);
        s1 = fma(Linv[(c + 1) * NB + tid], bk[c + 1], s1);
        s2 = fma(Linv[(c + 2) * NB + tid], bk[c + 2], s2);
        s3 = fma(Linv[(c + 3) * NB + tid], bk[c + 3], s3);
      }
      if (tid < nbk) b[tid] = (s0 + s1) + (s2 + s3);
    }
  } else {
    // x_c = sum_r Linv[r][c] b_r: one warp per column (coalesced along r)
    for (int c = wid; c < nbk; c += (blockDim.x >> 5)) {
      const double* col = Linv + c * NB;
      double s = fma(col[lane], bk[lane], 0.0);
      s = fma(col[lane + 32], bk[lane + 32], s);
      s = fma(col[lane + 64], bk[lane + 64], s);
      s = fma(col[lane + 96], bk[lane + 96], s);
      s = warp_sum(s);
      if (lane == 0) b[c] = s;
    }
  }
}

__global__ void lower_colsumsq_kernel(int64_t n, const double* __restrict__ X, int64_t ldx,
                                      double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  const double* col = X + j * ldx;
  double s0 = 0.0, s1 = 0.0;
  int64_t i = j + lane;
  for (; i + 32 < n; i += 64) {
    const double a = col[i], b = col[i + 32];
    s0 = fma(a, a, s0);
    s1 = fma(b, b, s1);
  }
  if (i < n) s0 = fma(col[i], col[i], s0);
  const double s = warp_sum(s0 + s1);
  if (lane == 0) out[j] = s;
}

__global__ void scale_kernel(int64_t m, int64_t n, double* __restrict__ A, int64_t lda,
                             const double* __restrict__ s, int by_col) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (int64_t j = blockIdx.y; j < n; j += gridDim.y) A[j * lda + i] *= by_col ? s[j] : s[i];
}

__global__ void scale_block_kernel(int64_t m, int64_t n, double* __restrict__ A, int64_t lda,
                                   const double* __restrict__ r, const double* __restrict__ c,
                                   double alpha, double diag_add, int64_t diag_offset) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (int64_t j = blockIdx.y; j < n; j += gridDim.y) {
    double v = A[j * lda + i];
    if (c) v *= c[j];
    if (r) v *= r[i];
    v *= alpha;
    if (i == j + diag_offset) v += diag_add;
    A[j * lda + i] = v;
  }
}

__global__ void vec_op_kernel(int op, int64_t n, const double* __restrict__ a,
                              const double* __restrict__ b, double* __restrict__ out, double alpha,
                              double beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = 0.0;
  switch (op) {
    case VOP_DIV: v = a[i] / b[i]; break;
    case VOP_MUL: v = a[i] * b[i]; break;
    case VOP_RSQRT_INV: v = sqrt(1.0 / a[i]); break;
    case VOP_SQRT: v = sqrt(a[i]); break;
    case VOP_FILL: v = alpha; break;
    case VOP_INV: v = 1.0 / a[i]; break;
    case VOP_SUB: v = a[i] - b[i]; break;
    case VOP_ADD: v = a[i] + b[i]; break;
    case VOP_AXPBY: v = alpha * a[i] + beta * b[i]; break;
    case VOP_COPY: v = a[i]; break;
  }
  out[i] = v;
}

__global__ void reduce_kernel(int op, int64_t n, const double* __restrict__ a,
                              const double* __restrict__ b, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = (op == RED_MIN) ? INFINITY : 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    switch (op) {
      case RED_SUM: acc += a[i]; break;
      case RED_DOT: acc = fma(a[i], b[i], acc); break;
      case RED_SUMLOG: acc += log(a[i]); break;
      case RED_MAXABSDIFF: {
        const double dlt = fabs(a[i] - b[i]);
        acc = (dlt > acc || dlt != dlt) ? dlt : acc;    // NaN propagates
        break;
      }
      case RED_MIN: acc = (a[i] < acc || a[i] != a[i]) ? a[i] : acc; break;
      case RED_COUNT_NAN: acc += (a[i] != a[i]) ? 1.0 : 0.0; if (b) acc += (b[i] != b[i]) ? 1.0 : 0.0; break;
      case RED_SUMSQ: acc = fma(a[i], a[i], acc); break;
      case RED_SUMLOG_PROD: acc += log(a[i] * b[i]); break;
    }
  }
  if (op == RED_MAXABSDIFF || op == RED_MIN) {
    const bool is_max = (op == RED_MAXABSDIFF);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, acc, o);
      if (other != other) acc = other;
      else if (acc == acc) acc = is_max ? fmax(acc, other) : fmin(acc, other);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = acc;
    __syncthreads();
    if (w == 0) {
      acc = (lane < (int)(blockDim.x >> 5)) ? red[lane] : (is_max ? 0.0 : INFINITY);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, acc, o);
        if (other != other) acc = other;
        else if (acc == acc) acc = is_max ? fmax(acc, other) : fmin(acc, other);
      }
      if (lane == 0) out[0] = acc;
    }
  } else {
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = s;
  }
}

// ---- likelihood maths (R/lik.R) ----------------------------------------------------------------

__device__ __forceinline__ double digamma_dev(double x) {
  // recurrence up to x >= 10, then the asymptotic series
  double r = 0.0;
  while (x < 10.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  const double xi = 1.0 / x, x2 = xi * xi;
  const double ser = x2 * (1.0 / 12 - x2 * (1.0 / 120 - x2 * (1.0 / 252 - x2 * (1.0 / 240 -
                     x2 * (1.0 / 132 - x2 * (691.0 / 32760 - x2 * (1.0 / 12)))))));
  return r + log(x) - 0.5 * xi - ser;
}

__device__ __forceinline__ double xlogy_dev(double x, double y) { return x == 0.0 ? 0.0 : x * log(y); }
__device__ __forceinline__ double xlog1py_dev(double x, double y) { return x == 0.0 ? 0.0 : x * log1p(y); }

struct LikP {
  int kind, link;
  double param;
};

__device__ __forceinline__ double response_dev(const LikP& l, double f) {   // R/lik.R:458-498
  if (l.link == GPL_LINK_LOGIT) return 1.0 / (1.0 + exp(-f));
  if (l.link == GPL_LINK_PROBIT) return normcdf(f);
  if (l.link == GPL_LINK_LOG) return exp(f);
  return f;
}

// f already includes the offset
__device__ __forceinline__ double loglik_dev(const LikP& l, double f, double y, double t) {
  switch (l.kind) {
    case GPL_LIK_GAUSSIAN: {                                                // R/lik.R:301-308
      const double r = (y - f) / l.param;
      return -(0.918938533204672741780329736406 + log(l.param) + 0.5 * r * r);
    }
    case GPL_LIK_BERNOULLI:
    case GPL_LIK_BINOMIAL: {                                                // R/lik.R:314-330
      const double mu = response_dev(l, f);
      const double lch = lgamma(t + 1.0) - lgamma(y + 1.0) - lgamma(t - y + 1.0);
      return lch + xlogy_dev(y, mu) + xlog1py_dev(t - y, -mu);
    }
    case GPL_LIK_BETABINOM: {                                               // R/lik.R:332-353
      const double mu = response_dev(l, f);
      const double a = mu / l.param, b = (1.0 - mu) / l.param;
      const double t1 = lgamma(t + 1.0) - lgamma(y + 1.0) - lgamma(t - y + 1.0);
      const double t2 = lgamma(y + a) + lgamma(t - y + b) - lgamma(t + a + b);
      const double t3 = lgamma(a + b) - lgamma(a) - lgamma(b);
      return t1 + t2 + t3;
    }
    case GPL_LIK_POISSON: {                                                 // R/lik.R:355-366
      const double mu = response_dev(l, f);
      return xlogy_dev(y, mu) - mu - lgamma(y + 1.0);
    }
  }
  return 0.0;
}

__device__ __forceinline__ double loglik_d_dev(const LikP& l, double f, double y, double t) {
  const double mu = response_dev(l, f);
  if (l.kind == GPL_LIK_POISSON) return (y / mu - 1.0) * mu;                // R/lik.R:441
  double dmu;
  if (l.link == GPL_LINK_PROBIT) dmu = 0.398942280401432677939946059934 * exp(-0.5 * f * f);
  else dmu = mu * (1.0 - mu);
  if (l.kind == GPL_LIK_BETABINOM) {                                        // R/lik.R:400-424
    const double phi = l.param, a = mu / phi, b = (1.0 - mu) / phi;
    const double t2 = (dmu / phi) * (digamma_dev(y + a) - digamma_dev(t - y + b));
    const double t3 = (dmu / phi) * (-digamma_dev(a) + digamma_dev(b));
    return t2 + t3;
  }
  return dmu * (y / mu - (t - y) / (1.0 - mu));                             // R/lik.R:396
}

__global__ void pseudodata_la_kernel(LikP l, int64_t n, const double* __restrict__ f,
                                     const double* __restrict__ y, const double* __restrict__ trials,
                                     const double* __restrict__ offset, double* __restrict__ z,
                                     double* __restrict__ V) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (l.kind == GPL_LIK_GAUSSIAN) {                                         // R/lik.R:188-192
    z[i] = y[i];
    V[i] = l.param * l.param * 1.0;
    return;
  }
  const double t = (l.kind == GPL_LIK_BERNOULLI) ? 1.0 : (trials ? trials[i] : 1.0);
  const double off = offset ? offset[i] : 0.0;
  const double fi = f[i];
  const double eps = 1e-6;                                                  // R/lik.R:449
  const double g = loglik_d_dev(l, fi + off, y[i], t);
  const double ge = loglik_d_dev(l, (fi + eps) + off, y[i], t);
  double g2 = (ge - g) / eps;
  const double sgn = (g2 > 0.0) ? 1.0 : ((g2 < 0.0) ? -1.0 : 0.0);
  g2 = sgn * fmax(fabs(g2), 1e-12);                                         // R/lik.R:198
  z[i] = fi - g / g2;
  V[i] = -1.0 / g2;
}

__global__ void loglik_sum_kernel(LikP l, int64_t n, const double* __restrict__ f,
                                  const double* __restrict__ y, const double* __restrict__ trials,
                                  const double* __restrict__ offset, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double t = (l.kind == GPL_LIK_BERNOULLI) ? 1.0 : (trials ? trials[i] : 1.0);
    const double off = offset ? offset[i] : 0.0;
    acc += loglik_dev(l, f[i] + off, y[i], t);
  }
  const double s = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = s;
}

constexpr int MAX_QUAD = 64;

__global__ void ep_sites_kernel(LikP l, int64_t n, EpSiteArgs a, const double* __restrict__ trials,
                                const double* __restrict__ offset, int* __restrict__ flags) {
  __shared__ double xq[MAX_QUAD], wq[MAX_QUAD];
  if (threadIdx.x < a.quad_order) {
    xq[threadIdx.x] = a.quad_x[threadIdx.x];
    wq[threadIdx.x] = a.quad_w[threadIdx.x];
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = (l.kind == GPL_LIK_BERNOULLI) ? 1.0 : (trials ? trials[i] : 1.0);
  const double off = offset ? offset[i] : 0.0;
  const double yi = a.y[i];
  // R/lik.R:239-240
  const double post_prec = 1.0 / a.fvar[i];
  const double P_old = a.P_old[i], z_old = a.z_old[i];
  const double cav_prec = post_prec - P_old;
  const double cav_mean = 1.0 / cav_prec * (a.fmean[i] * post_prec - z_old * P_old);
  const double scale = sqrt(1.0 / cav_prec);
  const int Q = a.quad_order;
  // pass 1: log-likelihood on the grid, its max (R/lik.R:243-259, R/gp_loo.R:136-144)
  double cmax = -INFINITY;
  bool ok = true;
  for (int q = 0; q < Q; ++q) {
    const double fg = xq[q] * scale + cav_mean;
    double ll = loglik_dev(l, fg + off, yi, t);
    if (ll != ll) { ll = -INFINITY; ok = false; }
    cmax = fmax(cmax, ll);
  }
  double sumw = 0.0, m1 = 0.0;
  for (int q = 0; q < Q; ++q) {
    const double fg = xq[q] * scale + cav_mean;
    double ll = loglik_dev(l, fg + off, yi, t);
    double lk = exp(ll);
    if (ll != ll) { ll = -INFINITY; lk = 0.0; }
    sumw += wq[q] * exp(ll - cmax);
    m1 += (fg * lk) * wq[q];
  }
  const double logC = cmax + log(sumw);
  const double C = exp(logC);
  const double mean_t = (1.0 / C) * m1;
  double m2 = 0.0;
  for (int q = 0; q < Q; ++q) {
    const double fg = xq[q] * scale + cav_mean;
    double ll = loglik_dev(l, fg + off, yi, t);
    double lk = exp(ll);
    if (ll != ll) lk = 0.0;
    const double dlt = fg - mean_t;
    m2 += ((dlt * dlt) * lk) * wq[q];
  }
  const double s2_t = (1.0 / C) * m2;
  const double var_cav = 1.0 / cav_prec;
  // R/lik.R:218-226
  double P_new = 1.0 / s2_t - 1.0 / var_cav;
  double nu_new = mean_t / s2_t - cav_mean / var_cav;
  const double nu_old = z_old * P_old;
  nu_new = a.damping * nu_new + (1.0 - a.damping) * nu_old;
  P_new = a.damping * P_new + (1.0 - a.damping) * P_old;
  const double zi = (1.0 / P_new) * nu_new;
  const double Vi = 1.0 / P_new;
  a.z_new[i] = zi;
  a.V_new[i] = Vi;
  a.P_new[i] = 1.0 / Vi;           // fit$P = 1 / V (R/ep.R:143)
  a.cav_mean[i] = cav_mean;
  a.cav_var[i] = var_cav;
  a.logZ[i] = logC;
  int fl = 0;
  if (!ok) fl |= 1;
  if (zi != zi || Vi != Vi) fl |= 2;
  else if (!(Vi > 0.0)) fl |= 4;
  if (fl) atomicOr(flags, fl);
}

__global__ void ep_evidence_kernel(int64_t n, const double* __restrict__ logZ,
                                   const double* __restrict__ vc, const double* __restrict__ mc,
                                   const double* __restrict__ z, const double* __restrict__ V,
                                   double* __restrict__ out) {
  __shared__ double red[32];
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double tot = vc[i] + V[i];
    const double dm = mc[i] - z[i];
    s1 += logZ[i];
    s2 += log(tot);
    s3 += dm * dm / tot;
  }
  const double a1 = block_sum(s1, red);
  const double a2 = block_sum(s2, red);
  const double a3 = block_sum(s3, red);
  if (threadIdx.x == 0) out[0] = a1 + 0.5 * a2 + 0.5 * a3;   // R/ep.R:47-48
}

// ---- small helpers of the non-log-concave ("bad site") branch ---------------------------------
__global__ void gather_vec_kernel(int64_t n, const double* __restrict__ src,
                                  const int32_t* __restrict__ perm, double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

__global__ void scatter_vec_kernel(int64_t n, const double* __restrict__ src,
                                   const int32_t* __restrict__ perm, double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[perm[i]] = src[i];
}

__global__ void gather_rows_kernel(int64_t n, int64_t ncols, const double* __restrict__ src,
                                   int64_t lds, const int32_t* __restrict__ perm,
                                   double* __restrict__ dst, int64_t ldd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t pi = perm[i];
  for (int64_t j = blockIdx.y; j < ncols; j += gridDim.y) dst[j * ldd + i] = src[j * lds + pi];
}

// B (n x m) = A^T, A m x n; 32 x 32 tiles through shared memory
__global__ void transpose_kernel(int64_t m, int64_t n, const double* __restrict__ A, int64_t lda,
                                 double* __restrict__ B, int64_t ldb) {
  __shared__ double tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
  for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
    const int64_t i = i0 + threadIdx.x, j = j0 + jj;
    tile[jj][threadIdx.x] = (i < m && j < n) ? A[j * lda + i] : 0.0;
  }
  __syncthreads();
  for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
    const int64_t j = j0 + threadIdx.x, i = i0 + ii;
    if (i < m && j < n) B[i * ldb + j] = tile[threadIdx.x][ii];
  }
}

// out(i,j) = (i >= j) ? r_i L(i,j) : 0   (clean lower-triangular copy, optional row scaling)
__global__ void tril_scale_kernel(int64_t n, const double* __restrict__ L, int64_t ldl,
                                  const double* __restrict__ r, double* __restrict__ out,
                                  int64_t ldo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int64_t j = blockIdx.y; j < n; j += gridDim.y)
    out[j * ldo + i] = (i >= j) ? (r ? r[i] : 1.0) * L[j * ldl + i] : 0.0;
}

// out_i = sum_j A_ij^2 w_j   (one warp per row; A m x n with small n)
__global__ void rowsumsq_w_kernel(int64_t m, int64_t n, const double* __restrict__ A, int64_t lda,
                                  const double* __restrict__ w, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double s = 0.0;
  for (int64_t j = 0; j < n; ++j) {
    const double a = A[j * lda + i];
    s += a * a * w[j];
  }
  out[i] = s;
}

// out_j = sum_i A_ij^2, one CTA per column
__global__ void colsumsq_kernel(int64_t m, const double* __restrict__ A, int64_t lda,
                                double* __restrict__ out) {
  __shared__ double red[32];
  const double* col = A + (int64_t)blockIdx.x * lda;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) s = fma(col[i], col[i], s);
  const double t = block_sum(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = t;
}

// out_i = sum_j A_ij B_ij
__global__ void rowdot_kernel(int64_t m, int64_t n, const double* __restrict__ A, int64_t lda,
                              const double* __restrict__ B, int64_t ldb, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double s = 0.0;
  for (int64_t j = 0; j < n; ++j) s = fma(A[j * lda + i], B[j * ldb + i], s);
  out[i] = s;
}

// bad-site EP evidence terms (R/ep.R:112-114): sum logZ + 0.5 sum log((vc+V)/V) + 0.5 sum (mc-z)^2/(vc+V)
__global__ void ep_evidence_bad_kernel(int64_t n, const double* __restrict__ logZ,
                                       const double* __restrict__ vc, const double* __restrict__ mc,
                                       const double* __restrict__ z, const double* __restrict__ V,
                                       double* __restrict__ out) {
  __shared__ double red[32];
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double tot = vc[i] + V[i];
    const double dm = mc[i] - z[i];
    s1 += logZ[i];
    s2 += log(tot / V[i]);
    s3 += dm * dm / tot;
  }
  const double a1 = block_sum(s1, red);
  const double a2 = block_sum(s2, red);
  const double a3 = block_sum(s3, red);
  if (threadIdx.x == 0) out[0] = a1 + 0.5 * a2 + 0.5 * a3;
}

}  // namespace

int build_B(Ctx* ctx, int64_t n, const double* K, int64_t ldk, const double* s, double* B,
            int64_t ldb) {
  if (n <= 0) return GPL_OK;
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)((n + 7) / 8));
  build_B_kernel<<<grid, 256, 0, ctx->stream>>>(n, K, ldk, s, B, ldb);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int gemv_t_ab(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const double* x,
              double* y, double alpha, double beta) {
  if (n <= 0) return GPL_OK;
  gemv_t_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(m, n, A, lda, x, y, alpha, beta);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int gemv_t(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const double* x, double* y) {
  return gemv_t_ab(ctx, m, n, A, lda, x, y, 1.0, 0.0);
}

template <int MODE>
static int row_reduce(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const double* x,
                      double* y, double alpha = 1.0, double beta = 0.0) {
  if (m <= 0) return GPL_OK;
  const int64_t row_blocks = (m + 255) / 256;
  int64_t nsplit = (4 * (int64_t)ctx->sm_count + row_blocks - 1) / row_blocks;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > (n + 63) / 64) nsplit = (n + 63) / 64;
  if (nsplit < 1) nsplit = 1;
  const int64_t cps = (n + nsplit - 1) / nsplit;
  DevBuf part;
  GPL_TRY(part.alloc(ctx, sizeof(double) * nsplit * m));
  dim3 grid((unsigned)row_blocks, (unsigned)nsplit);
  row_reduce_kernel<MODE><<<grid, 256, 0, ctx->stream>>>(m, n, A, lda, x, part.d(), cps);
  GPL_LAUNCH_CHECK(ctx);
  sum_partials_kernel<<<(unsigned)row_blocks, 256, 0, ctx->stream>>>(m, (int)nsplit, part.d(), y,
                                                                      alpha, beta);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int gemv_n(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const double* x, double* y) {
  return row_reduce<0>(ctx, m, n, A, lda, x, y);
}

int gemv_n_ab(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const double* x,
              double* y, double alpha, double beta) {
  return row_reduce<0>(ctx, m, n, A, lda, x, y, alpha, beta);
}

int rowsumsq(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, double* out) {
  return row_reduce<1>(ctx, m, n, A, lda, nullptr, out);
}

// recursive in-place triangular solve: the block products are plain GEMVs over L21 (HBM-bound,
// n^2/2 * 8 bytes in total), the 128-blocks use the factor's cached block inverses
static int trsv_rec(Ctx* ctx, const double* dinv, bool trans, int64_t n, const double* L,
                    int64_t ldl, double* x, int64_t blk0) {
  if (n <= NB) {
    trsv_block_kernel<<<1, 256, 0, ctx->stream>>>((int)n, dinv + blk0 * NB * NB, x, trans ? 1 : 0);
    GPL_LAUNCH_CHECK(ctx);
    return GPL_OK;
  }
  const int64_t nblk = (n + NB - 1) / NB;
  const int64_t n1 = (nblk / 2) * NB, n2 = n - n1;
  const double* L21 = L + n1;
  const double* L22 = L + n1 * ldl + n1;
  if (!trans) {
    GPL_TRY(trsv_rec(ctx, dinv, false, n1, L, ldl, x, blk0));
    GPL_TRY(gemv_n_ab(ctx, n2, n1, L21, ldl, x, x + n1, -1.0, 1.0));
    GPL_TRY(trsv_rec(ctx, dinv, false, n2, L22, ldl, x + n1, blk0 + n1 / NB));
  } else {
    GPL_TRY(trsv_rec(ctx, dinv, true, n2, L22, ldl, x + n1, blk0 + n1 / NB));
    GPL_TRY(gemv_t_ab(ctx, n2, n1, L21, ldl, x + n1, x, -1.0, 1.0));
    GPL_TRY(trsv_rec(ctx, dinv, true, n1, L, ldl, x, blk0));
  }
  return GPL_OK;
}

int trsv_lower(Ctx* ctx, const double* dinv, bool trans, int64_t n, const double* L, int64_t ldl,
               double* b, double* x) {
  if (n <= 0) return GPL_OK;
  if (b == x) {
    set_error("trsv_lower: work vector and solution must be distinct buffers");
    return GPL_ERR_BAD_ARG;
  }
  GPL_CUDA(cudaMemcpyAsync(x, b, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  return trsv_rec(ctx, dinv, trans, n, L, ldl, x, 0);
}

int lower_colsumsq(Ctx* ctx, int64_t n, const double* X, int64_t ldx, double* out) {
  if (n <= 0) return GPL_OK;
  lower_colsumsq_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(n, X, ldx, out);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int scale_cols(Ctx* ctx, int64_t m, int64_t n, double* A, int64_t lda, const double* s) {
  if (m <= 0 || n <= 0) return GPL_OK;
  dim3 grid((unsigned)((m + 255) / 256), grid_y_cap(n));
  scale_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, A, lda, s, 1);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int scale_rows(Ctx* ctx, int64_t m, int64_t n, double* A, int64_t lda, const double* s) {
  if (m <= 0 || n <= 0) return GPL_OK;
  dim3 grid((unsigned)((m + 255) / 256), grid_y_cap(n));
  scale_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, A, lda, s, 0);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int scale_block(Ctx* ctx, int64_t m, int64_t n, double* A, int64_t lda, const double* r,
                const double* c, double alpha, double diag_add, int64_t diag_offset) {
  if (m <= 0 || n <= 0) return GPL_OK;
  dim3 grid((unsigned)((m + 255) / 256), grid_y_cap(n));
  scale_block_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, A, lda, r, c, alpha, diag_add, diag_offset);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int vec_op(Ctx* ctx, VecOp op, int64_t n, const double* a, const double* b, double* out,
           double alpha, double beta) {
  if (n <= 0) return GPL_OK;
  vec_op_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((int)op, n, a, b, out, alpha,
                                                                     beta);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int reduce(Ctx* ctx, RedOp op, int64_t n, const double* a, const double* b, int slot) {
  reduce_kernel<<<1, 1024, 0, ctx->stream>>>((int)op, n, a, b, ctx->d_scal + slot);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int fetch_scalars(Ctx* ctx, int count, double* host_out) {
  GPL_CUDA(cudaMemcpyAsync(ctx->h_scal, ctx->d_scal, sizeof(double) * count, cudaMemcpyDeviceToHost,
                           ctx->stream));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < count; ++i) host_out[i] = ctx->h_scal[i];
  return GPL_OK;
}

int lik_pseudodata_la(Ctx* ctx, const LikDev& lik, int64_t n, const double* f, const double* y,
                      double* z, double* V) {
  if (n <= 0) return GPL_OK;
  LikP l{lik.kind, lik.link, lik.param};
  pseudodata_la_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
      l, n, f, y, lik.trials, lik.offset, z, V);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int lik_loglik_sum(Ctx* ctx, const LikDev& lik, int64_t n, const double* f, const double* y,
                   int slot) {
  LikP l{lik.kind, lik.link, lik.param};
  loglik_sum_kernel<<<1, 1024, 0, ctx->stream>>>(l, n, f, y, lik.trials, lik.offset,
                                                 ctx->d_scal + slot);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int lik_ep_sites(Ctx* ctx, const LikDev& lik, int64_t n, const EpSiteArgs& a, int* d_flags) {
  if (a.quad_order > MAX_QUAD || a.quad_order < 1) {
    set_error("EP quadrature order %d outside [1, %d]", a.quad_order, MAX_QUAD);
    return GPL_ERR_BAD_ARG;
  }
  LikP l{lik.kind, lik.link, lik.param};
  GPL_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int), ctx->stream));
  ep_sites_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(l, n, a, lik.trials,
                                                                        lik.offset, d_flags);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int ep_evidence_terms(Ctx* ctx, int64_t n, const double* logZ, const double* vc, const double* mc,
                      const double* z, const double* V, int slot) {
  ep_evidence_kernel<<<1, 1024, 0, ctx->stream>>>(n, logZ, vc, mc, z, V, ctx->d_scal + slot);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int gather_vec(Ctx* ctx, int64_t n, const double* src, const int32_t* perm, double* dst) {
  if (n <= 0) return GPL_OK;
  gather_vec_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, src, perm, dst);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int scatter_vec(Ctx* ctx, int64_t n, const double* src, const int32_t* perm, double* dst) {
  if (n <= 0) return GPL_OK;
  scatter_vec_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, src, perm, dst);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int gather_rows(Ctx* ctx, int64_t n, int64_t ncols, const double* src, int64_t lds,
                const int32_t* perm, double* dst, int64_t ldd) {
  if (n <= 0 || ncols <= 0) return GPL_OK;
  dim3 grid((unsigned)((n + 255) / 256), grid_y_cap(ncols));
  gather_rows_kernel<<<grid, 256, 0, ctx->stream>>>(n, ncols, src, lds, perm, dst, ldd);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int transpose(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, double* B, int64_t ldb) {
  if (m <= 0 || n <= 0) return GPL_OK;
  dim3 grid((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(m, n, A, lda, B, ldb);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int tril_scale(Ctx* ctx, int64_t n, const double* L, int64_t ldl, const double* r, double* out,
               int64_t ldo) {
  if (n <= 0) return GPL_OK;
  dim3 grid((unsigned)((n + 255) / 256), grid_y_cap(n));
  tril_scale_kernel<<<grid, 256, 0, ctx->stream>>>(n, L, ldl, r, out, ldo);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int rowsumsq_w(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const double* w,
               double* out) {
  if (m <= 0) return GPL_OK;
  rowsumsq_w_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(m, n, A, lda, w, out);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int colsumsq(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, double* out) {
  if (n <= 0) return GPL_OK;
  colsumsq_kernel<<<(unsigned)n, 256, 0, ctx->stream>>>(m, A, lda, out);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int rowdot(Ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const double* B, int64_t ldb,
           double* out) {
  if (m <= 0) return GPL_OK;
  rowdot_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(m, n, A, lda, B, ldb, out);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int ep_evidence_terms_bad(Ctx* ctx, int64_t n, const double* logZ, const double* vc,
                          const double* mc, const double* z, const double* V, int slot) {
  ep_evidence_bad_kernel<<<1, 1024, 0, ctx->stream>>>(n, logZ, vc, mc, z, V, ctx->d_scal + slot);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

}  // namespace gpl
