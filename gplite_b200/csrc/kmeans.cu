// Inducing-point selection for method_fitc: k-means on the standardised inputs,
// the step before the FITC hot path (get_inducing, R/gp_fit.R:190-234; stats::kmeans).
//
// Lloyd iterations on the device: an assignment kernel (thread = point, centres staged through
// shared memory in tiles, nearest centre by direct squared distance, lowest index wins ties) that
// also accumulates the per-cluster sums with fp64 atomics, and a centre update kernel.  The
// reference's Hartigan-Wong variant and its R RNG stream are not reproducible outside R; the
// initial centres are rows picked by the caller, so the result is a deterministic function of
// (x, init) that the tests restate in numpy.
#include "fit.cuh"

#include <vector>

namespace gpl {

namespace {

constexpr int KM_THREADS = 256;
constexpr int KM_TILE = 128;     // centres per shared-memory tile
constexpr int KM_MAXD = 32;

// x: n x d column-major (ld n); centres: m x d column-major (ld m)
template <int D>
__global__ void __launch_bounds__(KM_THREADS)
kmeans_assign_kernel(int64_t n, int d_rt, const double* __restrict__ x, int64_t m,
                     const double* __restrict__ cen, int32_t* __restrict__ label,
                     double* __restrict__ sums, double* __restrict__ counts,
                     unsigned long long* __restrict__ changed) {
  __shared__ double sc[KM_TILE * KM_MAXD];
  const int d = D > 0 ? D : d_rt;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double xi[D > 0 ? D : KM_MAXD];
#pragma unroll
  for (int k = 0; k < (D > 0 ? D : KM_MAXD); ++k) xi[k] = (k < d && i < n) ? x[(int64_t)k * n + i] : 0.0;
  double best = INFINITY;
  int32_t arg = 0;
  for (int64_t c0 = 0; c0 < m; c0 += KM_TILE) {
    const int tc = (int)((m - c0) < KM_TILE ? (m - c0) : KM_TILE);
    __syncthreads();
    for (int idx = threadIdx.x; idx < tc * d; idx += KM_THREADS) {
      const int k = idx / tc, c = idx - k * tc;
      sc[c * KM_MAXD + k] = cen[(int64_t)k * m + c0 + c];
    }
    __syncthreads();
    for (int c = 0; c < tc; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < (D > 0 ? D : KM_MAXD); ++k) {
        if (k < d) {
          const double t = xi[k] - sc[c * KM_MAXD + k];
          s = fma(t, t, s);
        }
      }
      if (s < best) {
        best = s;
        arg = (int32_t)(c0 + c);
      }
    }
  }
  if (i >= n) return;
  if (label[i] != arg) {
    label[i] = arg;
    atomicAdd(changed, 1ULL);
  }
  for (int k = 0; k < d; ++k) atomicAdd(&sums[(int64_t)k * m + arg], xi[k]);
  atomicAdd(&counts[arg], 1.0);
}

// empty clusters keep their centre
__global__ void kmeans_update_kernel(int64_t m, int d, const double* __restrict__ sums,
                                     const double* __restrict__ counts, double* __restrict__ cen) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  const double cnt = counts[c];
  if (cnt > 0.0)
    for (int k = 0; k < d; ++k) cen[(int64_t)k * m + c] = sums[(int64_t)k * m + c] / cnt;
}

__global__ void kmeans_init_kernel(int64_t n, int64_t m, int d, const double* __restrict__ x,
                                   const int64_t* __restrict__ idx, double* __restrict__ cen) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  for (int k = 0; k < d; ++k) cen[(int64_t)k * m + c] = x[(int64_t)k * n + idx[c]];
}

}  // namespace

int kmeans_impl(Ctx* ctx, const double* x, int64_t n, int d, const int64_t* init_idx, int64_t m,
                int maxiter, double* centers, int32_t* labels, int32_t* iters_out) {
  if (!x || !init_idx || !centers || n <= 0 || m <= 0 || m > n || d <= 0 || d > KM_MAXD ||
      maxiter < 1) {
    set_error("gpl_kmeans: bad arguments (need 0 < m <= n, 0 < d <= %d, maxiter >= 1)", KM_MAXD);
    return GPL_ERR_BAD_ARG;
  }
  for (int64_t c = 0; c < m; ++c)
    if (init_idx[c] < 0 || init_idx[c] >= n) {
      set_error("gpl_kmeans: initial centre %lld refers to row %lld of %lld", (long long)c,
                (long long)init_idx[c], (long long)n);
      return GPL_ERR_BAD_ARG;
    }
  DevBuf dx, dc, dsum, dcnt, dlab, didx, dchg;
  GPL_TRY(dx.alloc(ctx, sizeof(double) * n * d));
  GPL_TRY(dc.alloc(ctx, sizeof(double) * m * d));
  GPL_TRY(dsum.alloc(ctx, sizeof(double) * m * d));
  GPL_TRY(dcnt.alloc(ctx, sizeof(double) * m));
  GPL_TRY(dlab.alloc(ctx, sizeof(int32_t) * n));
  GPL_TRY(didx.alloc(ctx, sizeof(int64_t) * m));
  GPL_TRY(dchg.alloc(ctx, sizeof(unsigned long long)));
  cudaStream_t s = ctx->stream;
  GPL_CUDA(cudaMemcpyAsync(dx.p, x, sizeof(double) * n * d, cudaMemcpyHostToDevice, s));
  GPL_CUDA(cudaMemcpyAsync(didx.p, init_idx, sizeof(int64_t) * m, cudaMemcpyHostToDevice, s));
  GPL_CUDA(cudaMemsetAsync(dlab.p, 0xff, sizeof(int32_t) * n, s));      // label -1: everything "changes"
  kmeans_init_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(
      n, m, d, dx.d(), static_cast<const int64_t*>(didx.p), dc.d());
  GPL_LAUNCH_CHECK(ctx);
  int it = 0;
  for (it = 1; it <= maxiter; ++it) {
    GPL_CUDA(cudaMemsetAsync(dsum.p, 0, sizeof(double) * m * d, s));
    GPL_CUDA(cudaMemsetAsync(dcnt.p, 0, sizeof(double) * m, s));
    GPL_CUDA(cudaMemsetAsync(dchg.p, 0, sizeof(unsigned long long), s));
    const unsigned grid = (unsigned)((n + KM_THREADS - 1) / KM_THREADS);
    int32_t* lab = static_cast<int32_t*>(dlab.p);
    unsigned long long* chg = static_cast<unsigned long long*>(dchg.p);
#define KM_LAUNCH(DD)                                                                          \
  kmeans_assign_kernel<DD><<<grid, KM_THREADS, 0, s>>>(n, d, dx.d(), m, dc.d(), lab, dsum.d(), \
                                                       dcnt.d(), chg)
    switch (d) {
      case 1: KM_LAUNCH(1); break;
      case 2: KM_LAUNCH(2); break;
      case 3: KM_LAUNCH(3); break;
      case 4: KM_LAUNCH(4); break;
      case 6: KM_LAUNCH(6); break;
      case 8: KM_LAUNCH(8); break;
      case 10: KM_LAUNCH(10); break;
      default: KM_LAUNCH(0); break;
    }
#undef KM_LAUNCH
    GPL_LAUNCH_CHECK(ctx);
    unsigned long long changed = 0;
    GPL_CUDA(cudaMemcpyAsync(&changed, dchg.p, sizeof(changed), cudaMemcpyDeviceToHost, s));
    GPL_CUDA(cudaStreamSynchronize(s));
    if (changed == 0) break;                       // assignments stable: centres are already their means
    kmeans_update_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(m, d, dsum.d(), dcnt.d(), dc.d());
    GPL_LAUNCH_CHECK(ctx);
  }
  if (it > maxiter) it = maxiter;
  GPL_CUDA(cudaMemcpyAsync(centers, dc.p, sizeof(double) * m * d, cudaMemcpyDeviceToHost, s));
  if (labels)
    GPL_CUDA(cudaMemcpyAsync(labels, dlab.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  GPL_CUDA(cudaStreamSynchronize(s));
  if (iters_out) *iters_out = it;
  return GPL_OK;
}

}  // namespace gpl

extern "C" int gpl_kmeans(gpl_ctx* h, const double* x, int64_t n, int32_t d, const int64_t* init_idx,
                          int64_t m, int32_t maxiter, double* centers, int32_t* labels,
                          int32_t* iters) {
  return gpl::kmeans_impl(reinterpret_cast<gpl::Ctx*>(h), x, n, d, init_idx, m, maxiter, centers,
                          labels, iters);
}
