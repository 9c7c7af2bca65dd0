// Blocked FP64 Cholesky, triangular solves and SPD inverse, all built on dgemm() (DMMA).
//
// The factorisation is the recursive right-looking form: factor the leading half, solve the
// panel below it, SYRK-update the trailing half, recurse.  Every flop above the 128x128 base
// block is a large DMMA GEMM/SYRK (K = n/2, n/4, ...), so trailing updates run at tensor-pipe
// speed and the trailing matrix is re-read only O(log n) times.  The base block is factored
// in one CTA's shared memory, which also emits the block's explicit inverse; triangular solves
// against a base block then become a K=128 GEMM (in place: a tile only overwrites rows/cols it
// alone read).  The same block inverses seed trtri, and lauum finishes (L L^T)^-1.
//
// Replaces base-R chol() / forwardsolve() / backsolve() (LAPACK dpotrf, BLAS dtrsm) at
// R/laplace.R:28-29, R/ep.R:39-42, R/gp_fit.R:108-110, R/linear_algebra.R:28-32,
// R/gp_pred.R:218.
#include "common.cuh"
#include <atomic>

#include <math.h>
#include <stdlib.h>

namespace gpl {

namespace {

constexpr int PLD = NB + 1;
constexpr int PW = 16;                 // panel width of the in-CTA factorisation
constexpr int NPANEL = NB / PW;
constexpr int POTF2_THREADS = 512;
constexpr int POTF2_SMEM = (NB * PLD + NB) * (int)sizeof(double);

__device__ __forceinline__ double shfl_d(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

// Factor one nb x nb (nb <= 128) diagonal block in shared memory and emit its inverse.
//   A    : in/out, lower triangle replaced by L
//   Linv : out, 128 x 128 (ld 128), lower = L^-1, strict upper = 0, identity padding beyond nb
//   info : device int; set to the smallest 1-based global pivot index that was not > 0
//
// Blocked by 16-column panels so that every flop outside the 16 x 16 diagonal blocks is a
// register-tiled rank-16 update (a rank-1 formulation needs 3 shared-memory accesses per FMA and
// ran at ~250 us per block).  Per panel: (A) warp 0 factors the 16 x 16 diagonal block in
// registers with shuffles, (B) one thread per row below solves its 16 unknowns by substitution,
// (C) all 16 warps apply the rank-16 update: warp <-> 8 columns, lane <-> rows {lane + 32 r},
// so panel reads are conflict-free (rows) or broadcasts (columns).  The inverse uses the same
// three primitives on the block recurrence X_IJ = -X_II sum_K L_IK X_KJ.
__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* __restrict__ A, int64_t lda, int nb, double* __restrict__ Linv,
                 int* __restrict__ info, int pivot_base) {
  extern __shared__ double S[];          // S[c * PLD + r] = element (r, c)
  double* rdiag = S + NB * PLD;          // 1 / L_jj
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int idx = tid; idx < NB * NB; idx += POTF2_THREADS) {
    const int i = idx & (NB - 1), c = idx >> 7;
    double v = (i == c) ? 1.0 : 0.0;
    if (i < nb && c < nb && i >= c) v = A[(int64_t)c * lda + i];
    S[c * PLD + i] = v;
  }
  __syncthreads();

  // ------------------------------------------------------------------ factorisation
  for (int p = 0; p < NPANEL; ++p) {
    const int c0 = p * PW;
    // (A) diagonal block, one warp, lane r (mod 16) owns row r
    if (warp == 0) {
      const int r = lane & 15;
      double a[PW];
#pragma unroll
      for (int k = 0; k < PW; ++k) a[k] = S[(c0 + k) * PLD + c0 + r];
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        const double djj = shfl_d(a[j], j);
        if (lane == 0 && !(djj > 0.0) && c0 + j < nb) atomicMin(info, pivot_base + c0 + j + 1);
        const double rinv = rsqrt(djj);
        a[j] = (r == j) ? djj * rinv : a[j] * rinv;
        if (lane == 0) rdiag[c0 + j] = rinv;
#pragma unroll
        for (int k = j + 1; k < PW; ++k) {
          const double lkj = shfl_d(a[j], k);
          a[k] = fma(-a[j], lkj, a[k]);
        }
      }
      // lanes 16..31 read the same elements as lanes 0..15 above; the shuffles already order those
      // reads before the write-back, the explicit barrier states it (and satisfies racecheck)
      __syncwarp();
      if (lane < PW) {
#pragma unroll
        for (int k = 0; k < PW; ++k)
          if (k <= r) S[(c0 + k) * PLD + c0 + r] = a[k];
      }
    }
    __syncthreads();
    // (B) rows below the diagonal block: x L_D^T = a, one thread per row
    {
      const int i = c0 + PW + tid;
      if (i < NB) {
        double x[PW];
#pragma unroll
        for (int j = 0; j < PW; ++j) {
          double sacc = S[(c0 + j) * PLD + i];
#pragma unroll
          for (int k = 0; k < j; ++k) sacc = fma(-x[k], S[(c0 + k) * PLD + c0 + j], sacc);
          x[j] = sacc * rdiag[c0 + j];
        }
#pragma unroll
        for (int j = 0; j < PW; ++j) S[(c0 + j) * PLD + i] = x[j];
      }
    }
    __syncthreads();
    // (C) trailing update S[i][k] -= sum_c P[i][c] P[k][c], lower triangle
    {
      const int kc0 = c0 + PW + 8 * warp;
      if (kc0 < NB) {
        double acc[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[r][q] = 0.0;
#pragma unroll 4
        for (int c = 0; c < PW; ++c) {
          const double* col = S + (c0 + c) * PLD;
          double ai[4], ak[8];
#pragma unroll
          for (int r = 0; r < 4; ++r) ai[r] = col[lane + 32 * r];
#pragma unroll
          for (int q = 0; q < 8; ++q) ak[q] = col[kc0 + q];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[r][q] = fma(ai[r], ak[q], acc[r][q]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int i = lane + 32 * r;
            if (i >= kc0 + q) S[(kc0 + q) * PLD + i] -= acc[r][q];
          }
      }
    }
    __syncthreads();
  }

  // L back to global
  for (int idx = tid; idx < NB * NB; idx += POTF2_THREADS) {
    const int i = idx & (NB - 1), c = idx >> 7;
    if (i < nb && c < nb && i >= c) A[(int64_t)c * lda + i] = S[c * PLD + i];
  }
  __syncthreads();

  // ------------------------------------------------------------------ inverse
  // (A') all eight 16 x 16 diagonal blocks at once, one warp each: lane c owns column c of the
  // inverse, L[i][k] comes by shuffle from the lane that owns row i.
  if (warp < NPANEL) {
    const int c0 = warp * PW;
    const int r = lane & 15;
    double a[PW], x[PW];
#pragma unroll
    for (int k = 0; k < PW; ++k) a[k] = S[(c0 + k) * PLD + c0 + r];
#pragma unroll
    for (int i = 0; i < PW; ++i) {
      double sacc = (i == r) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < i; ++k) {
        const double lik = shfl_d(a[k], i);
        sacc = fma(-lik, x[k], sacc);
      }
      x[i] = sacc * rdiag[c0 + i];
    }
    __syncwarp();
    if (lane < PW) {
#pragma unroll
      for (int i = 0; i < PW; ++i)
        if (i >= r) S[(c0 + r) * PLD + c0 + i] = x[i];
    }
  }
  __syncthreads();

  for (int K = 0; K < NPANEL; ++K) {
    const int k0 = K * PW;
    // (B') finalise block row K: X_KJ = -X_KK R_KJ for the columns left of the diagonal block
    if (tid < k0) {
      double* col = S + tid * PLD + k0;
      double rv[PW], ov[PW];
#pragma unroll
      for (int c = 0; c < PW; ++c) rv[c] = col[c];
#pragma unroll
      for (int r = 0; r < PW; ++r) {
        double sacc = 0.0;
#pragma unroll
        for (int c = 0; c <= r; ++c) sacc = fma(S[(k0 + c) * PLD + k0 + r], rv[c], sacc);
        ov[r] = -sacc;
      }
#pragma unroll
      for (int r = 0; r < PW; ++r) col[r] = ov[r];
    }
    __syncthreads();
    // (C') rows below block row K: R_IJ += L_IK X_KJ (J < K) and R_IK = L_IK X_KK.  The second
    // overwrites its own input, so everything is accumulated first and written after a barrier.
    const int kc0 = 8 * warp;                       // this warp's 8 columns
    const bool active = (kc0 < k0 + PW) && (k0 + PW < NB);
    double acc[4][8];
    if (active) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[r][q] = 0.0;
#pragma unroll 4
      for (int c = 0; c < PW; ++c) {
        double ai[4], xk[8];
#pragma unroll
        for (int r = 0; r < 4; ++r) ai[r] = S[(k0 + c) * PLD + lane + 32 * r];      // L[i][k0 + c]
#pragma unroll
        for (int q = 0; q < 8; ++q) xk[q] = (kc0 + q <= k0 + c) ? S[(kc0 + q) * PLD + k0 + c] : 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[r][q] = fma(ai[r], xk[q], acc[r][q]);
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = lane + 32 * r;
          if (i >= k0 + PW) {
            double* dst = S + (kc0 + q) * PLD + i;
            *dst = (kc0 + q >= k0) ? acc[r][q] : *dst + acc[r][q];
          }
        }
    }
    __syncthreads();
  }

  for (int idx = tid; idx < NB * NB; idx += POTF2_THREADS) {
    const int i = idx & (NB - 1), c = idx >> 7;
    Linv[c * NB + i] = (i >= c) ? S[c * PLD + i] : 0.0;
  }
}

__global__ void diag_logsum_kernel(const double* __restrict__ A, int64_t lda, int64_t n,
                                   double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += log(A[i * lda + i]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[0] = 2.0 * s;
  }
}

__global__ void copy_block_kernel(const double* __restrict__ src, int64_t lds, double* dst,
                                  int64_t ldd, int rows, int cols) {
  for (int idx = threadIdx.x + blockIdx.x * blockDim.x; idx < rows * cols;
       idx += blockDim.x * gridDim.x) {
    const int i = idx % rows, c = idx / rows;
    dst[(int64_t)c * ldd + i] = src[(int64_t)c * lds + i];
  }
}

__global__ void zero_upper_kernel(double* __restrict__ A, int64_t lda, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i < j && i < n) A[j * lda + i] = 0.0;
}

inline int64_t split(int64_t n) {
  const int64_t nblk = (n + NB - 1) / NB;
  return (nblk / 2) * NB;
}

struct Fac {
  Ctx* ctx;
  double* dinv;        // block inverses of the factor being built / used
  int* info = nullptr; // device int receiving the smallest failing pivot (default: ctx->d_info)
  int pivot0 = 0;      // added to the reported pivot (a diagonal block of a larger matrix)
};

int potf2(const Fac& f, double* A, int64_t lda, int64_t n, int64_t blk0) {
  static std::atomic<bool> configured[64];       // see gemm_f64.cu launch(): idempotent, release / acquire
  std::atomic<bool>& cfgd = configured[f.ctx->device & 63];
  if (!cfgd.load(std::memory_order_acquire)) {
    GPL_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  POTF2_SMEM));
    cfgd.store(true, std::memory_order_release);
  }
  potf2_inv_kernel<<<1, POTF2_THREADS, POTF2_SMEM, f.ctx->stream>>>(
      A, lda, (int)n, f.dinv + blk0 * NB * NB, f.info ? f.info : f.ctx->d_info,
      f.pivot0 + (int)(blk0 * NB));
  GPL_LAUNCH_CHECK(f.ctx);
  return GPL_OK;
}

// B (m x n) <- B L^-T   (trans) or   B L^-1   (!trans); L is n x n lower with block inverses
int trsm_right_rec(const Fac& f, bool trans, int64_t m, int64_t n, const double* L, int64_t ldl,
                   double* B, int64_t ldb, int64_t blk0) {
  if (m <= 0 || n <= 0) return GPL_OK;
  if (n <= NB) {
    return dgemm(f.ctx, false, trans, m, n, n, 1.0, B, ldb, f.dinv + blk0 * NB * NB, NB, 0.0, B,
                 ldb, false, 1);
  }
  const int64_t n1 = split(n), n2 = n - n1;
  const double* L21 = L + n1;
  const double* L22 = L + n1 * ldl + n1;
  double* B2 = B + n1 * ldb;
  if (trans) {
    GPL_TRY(trsm_right_rec(f, true, m, n1, L, ldl, B, ldb, blk0));
    GPL_TRY(dgemm(f.ctx, false, true, m, n2, n1, -1.0, B, ldb, L21, ldl, 1.0, B2, ldb, false));
    GPL_TRY(trsm_right_rec(f, true, m, n2, L22, ldl, B2, ldb, blk0 + n1 / NB));
  } else {
    GPL_TRY(trsm_right_rec(f, false, m, n2, L22, ldl, B2, ldb, blk0 + n1 / NB));
    GPL_TRY(dgemm(f.ctx, false, false, m, n1, n2, -1.0, B2, ldb, L21, ldl, 1.0, B, ldb, false));
    GPL_TRY(trsm_right_rec(f, false, m, n1, L, ldl, B, ldb, blk0));
  }
  return GPL_OK;
}

int potrf_rec(const Fac& f, double* A, int64_t lda, int64_t n, int64_t blk0) {
  if (n <= NB) return potf2(f, A, lda, n, blk0);
  const int64_t n1 = split(n), n2 = n - n1;
  double* A21 = A + n1;
  double* A22 = A + n1 * lda + n1;
  GPL_TRY(potrf_rec(f, A, lda, n1, blk0));
  GPL_TRY(trsm_right_rec(f, true, n2, n1, A, lda, A21, lda, blk0));
  GPL_TRY(dgemm(f.ctx, false, true, n2, n2, n1, -1.0, A21, lda, A21, lda, 1.0, A22, lda, true));
  GPL_TRY(potrf_rec(f, A22, lda, n2, blk0 + n1 / NB));
  return GPL_OK;
}

// B (m x n) <- B X,  X n x n lower triangular with clean (zero) upper parts in its diagonal blocks
int trmm_right_lower(const Fac& f, int64_t m, int64_t n, const double* X, int64_t ldx, double* B,
                     int64_t ldb) {
  if (m <= 0 || n <= 0) return GPL_OK;
  if (n <= NB) return dgemm(f.ctx, false, false, m, n, n, 1.0, B, ldb, X, ldx, 0.0, B, ldb, false, 1);
  const int64_t n1 = split(n), n2 = n - n1;
  double* B2 = B + n1 * ldb;
  GPL_TRY(trmm_right_lower(f, m, n1, X, ldx, B, ldb));
  GPL_TRY(dgemm(f.ctx, false, false, m, n1, n2, 1.0, B2, ldb, X + n1, ldx, 1.0, B, ldb, false));
  GPL_TRY(trmm_right_lower(f, m, n2, X + n1 * ldx + n1, ldx, B2, ldb));
  return GPL_OK;
}

// B (n x m) <- alpha X B
int trmm_left_lower(const Fac& f, double alpha, int64_t n, int64_t m, const double* X, int64_t ldx,
                    double* B, int64_t ldb) {
  if (m <= 0 || n <= 0) return GPL_OK;
  if (n <= NB) return dgemm(f.ctx, false, false, n, m, n, alpha, X, ldx, B, ldb, 0.0, B, ldb, false, 2);
  const int64_t n1 = split(n), n2 = n - n1;
  double* B2 = B + n1;
  GPL_TRY(trmm_left_lower(f, alpha, n2, m, X + n1 * ldx + n1, ldx, B2, ldb));
  GPL_TRY(dgemm(f.ctx, false, false, n2, m, n1, alpha, X + n1, ldx, B, ldb, 1.0, B2, ldb, false));
  GPL_TRY(trmm_left_lower(f, alpha, n1, m, X, ldx, B, ldb));
  return GPL_OK;
}

// B (n x m) <- X^T B
int trmm_left_lowerT(const Fac& f, int64_t n, int64_t m, const double* X, int64_t ldx, double* B,
                     int64_t ldb) {
  if (m <= 0 || n <= 0) return GPL_OK;
  if (n <= NB) return dgemm(f.ctx, true, false, n, m, n, 1.0, X, ldx, B, ldb, 0.0, B, ldb, false, 2);
  const int64_t n1 = split(n), n2 = n - n1;
  double* B2 = B + n1;
  GPL_TRY(trmm_left_lowerT(f, n1, m, X, ldx, B, ldb));
  GPL_TRY(dgemm(f.ctx, true, false, n1, m, n2, 1.0, X + n1, ldx, B2, ldb, 1.0, B, ldb, false));
  GPL_TRY(trmm_left_lowerT(f, n2, m, X + n1 * ldx + n1, ldx, B2, ldb));
  return GPL_OK;
}

// L (n x n lower, factor with block inverses) <- L^-1 in place
int trtri_rec(const Fac& f, double* L, int64_t ldl, int64_t n, int64_t blk0) {
  if (n <= NB) {
    copy_block_kernel<<<16, 256, 0, f.ctx->stream>>>(f.dinv + blk0 * NB * NB, NB, L, ldl, (int)n,
                                                     (int)n);
    GPL_LAUNCH_CHECK(f.ctx);
    return GPL_OK;
  }
  const int64_t n1 = split(n), n2 = n - n1;
  double* L21 = L + n1;
  double* L22 = L + n1 * ldl + n1;
  GPL_TRY(trtri_rec(f, L, ldl, n1, blk0));
  GPL_TRY(trtri_rec(f, L22, ldl, n2, blk0 + n1 / NB));
  GPL_TRY(trmm_right_lower(f, n2, n1, L, ldl, L21, ldl));          // L21 X11
  GPL_TRY(trmm_left_lower(f, -1.0, n2, n1, L22, ldl, L21, ldl));   // -X22 (L21 X11)
  return GPL_OK;
}

// X (n x n lower) <- X^T X (lower triangle; diagonal blocks come out full)
int lauum_rec(const Fac& f, double* X, int64_t ldx, int64_t n) {
  if (n <= NB) return dgemm(f.ctx, true, false, n, n, n, 1.0, X, ldx, X, ldx, 0.0, X, ldx, false, 3);
  const int64_t n1 = split(n), n2 = n - n1;
  double* X21 = X + n1;
  double* X22 = X + n1 * ldx + n1;
  GPL_TRY(lauum_rec(f, X, ldx, n1));
  GPL_TRY(dgemm(f.ctx, true, false, n1, n1, n2, 1.0, X21, ldx, X21, ldx, 1.0, X, ldx, true));
  GPL_TRY(trmm_left_lowerT(f, n2, n1, X22, ldx, X21, ldx));
  GPL_TRY(lauum_rec(f, X22, ldx, n2));
  return GPL_OK;
}

// Left-looking block-row inverse of a lower-triangular factor for large n.  The recursive form
// spends a third of its time in mid-size GEMMs near the diagonal (measured 28.5 TFLOP/s at
// n = 32768); here every block row is two large triangular-aware GEMMs:
//   T = L[i, 0:i0] X[0:i0, 0:i0]     (X lower triangular: k starts at the tile's first column)
//   X[i, 0:i0] = -X_ii T              (X_ii lower triangular: k ends at the tile's last row)
// after the LL_NB x LL_NB diagonal blocks have been inverted by the recursive routine.
constexpr int64_t LL_NB = 1024;

int trtri_ll(const Fac& f, double* L, int64_t ldl, int64_t n, double* T, int64_t ldt) {
  // The diagonal-block inversions are independent chains of small, latency-bound kernels (~60
  // launches of ~10 us per block, a few SMs each): spread them over the side streams.
  Ctx* ctx = f.ctx;
  static const int no_aux = getenv("GPL_NO_AUX_STREAMS") ? 1 : 0;
  const bool fork = !no_aux && !ctx->profile && ctx->aux[0] != nullptr && n > 2 * LL_NB;
  cudaStream_t main = ctx->stream;
  if (fork) {
    GPL_CUDA(cudaEventRecord(ctx->ev_col, main));
    for (int i = 0; i < Ctx::NAUX; ++i) GPL_CUDA(cudaStreamWaitEvent(ctx->aux[i], ctx->ev_col, 0));
  }
  int rc = GPL_OK, b = 0;
  for (int64_t i0 = 0; i0 < n && rc == GPL_OK; i0 += LL_NB, ++b) {
    const int64_t w = (n - i0) < LL_NB ? (n - i0) : LL_NB;
    if (fork) ctx->stream = ctx->aux[b % Ctx::NAUX];
    rc = trtri_rec(f, L + i0 * ldl + i0, ldl, w, i0 / NB);
  }
  ctx->stream = main;
  if (fork) {
    for (int i = 0; i < Ctx::NAUX; ++i) {
      GPL_CUDA(cudaEventRecord(ctx->aux_ev[i], ctx->aux[i]));
      GPL_CUDA(cudaStreamWaitEvent(main, ctx->aux_ev[i], 0));
    }
  }
  GPL_TRY(rc);
  for (int64_t i0 = LL_NB; i0 < n; i0 += LL_NB) {
    const int64_t w = (n - i0) < LL_NB ? (n - i0) : LL_NB;
    GPL_TRY(dgemm(f.ctx, false, false, w, i0, i0, 1.0, L + i0, ldl, L, ldl, 0.0, T, ldt, false, 0, 2));
    GPL_TRY(dgemm(f.ctx, false, false, w, i0, w, -1.0, L + i0 * ldl + i0, ldl, T, ldt, 0.0, L + i0, ldl,
                  false, 0, 3));
  }
  return GPL_OK;
}

// Left-looking blocked factorisation for large n: block column j is updated by ONE large GEMM
// against everything to its left (K = j0), then its diagonal block is factored recursively and the
// rows below are solved.  Same n^3/3 flops as the recursion, but no mid-size trailing updates.
int potrf_ll(const Fac& f, double* A, int64_t lda, int64_t n) {
  for (int64_t j0 = 0; j0 < n; j0 += LL_NB) {
    const int64_t w = (n - j0) < LL_NB ? (n - j0) : LL_NB;
    double* Ajj = A + j0 * lda + j0;
    if (j0 > 0) {
      GPL_TRY(dgemm(f.ctx, false, true, n - j0, w, j0, -1.0, A + j0, lda, A + j0, lda, 1.0, Ajj, lda,
                    false));
    }
    GPL_TRY(potrf_rec(f, Ajj, lda, w, j0 / NB));
    if (n - j0 - w > 0)
      GPL_TRY(trsm_right_rec(f, true, n - j0 - w, w, Ajj, lda, Ajj + w, lda, j0 / NB));
  }
  return GPL_OK;
}

// Blocked right-looking factorisation with one step of look-ahead for large n.  Block columns of
// LA_NB: after panel k has been solved, block column k+1 is updated first and its diagonal block is
// factored (the recursive routine: a latency-bound chain of small kernels) on a high-priority side
// stream while the main stream applies panel k to the rest of the trailing matrix -- so the
// small-kernel chain leaves the critical path.
// Panel width: measured (tools/potrf_variants.sh) 1024 wins up to n = 16384 (53.7 vs 56.9 ms), 2048
// at n = 32768 (372.3 vs 374.6 ms; plain recursion 383.9 ms).  GPL_LA_NB overrides (multiple of 128).
static int64_t la_nb_for(int64_t n) {
  static const int64_t forced = getenv("GPL_LA_NB") ? atoll(getenv("GPL_LA_NB")) : 0;
  if (forced > 0) return forced;
  return n >= 24576 ? 2048 : 1024;
}

int potrf_lookahead(const Fac& f, double* A, int64_t lda, int64_t n) {
  const int64_t LA_NB = la_nb_for(n);
  Ctx* ctx = f.ctx;
  cudaStream_t main = ctx->stream, side = ctx->panel_stream;
  const int64_t nblk = (n + LA_NB - 1) / LA_NB;
  // order the side stream after whatever produced A on the main stream
  GPL_CUDA(cudaEventRecord(ctx->ev_col, main));
  GPL_CUDA(cudaStreamWaitEvent(side, ctx->ev_col, 0));
  ctx->stream = side;
  int rc = potrf_rec(f, A, lda, LA_NB < n ? LA_NB : n, 0);
  ctx->stream = main;
  GPL_TRY(rc);
  GPL_CUDA(cudaEventRecord(ctx->ev_diag, side));
  for (int64_t k = 0; k < nblk; ++k) {
    const int64_t k0 = k * LA_NB;
    const int64_t w = (n - k0) < LA_NB ? (n - k0) : LA_NB;
    const int64_t k1 = k0 + w;
    if (k1 >= n) {
      GPL_CUDA(cudaStreamWaitEvent(main, ctx->ev_diag, 0));
      break;
    }
    double* Akk = A + k0 * lda + k0;
    double* P = A + k0 * lda + k1;                    // panel rows k1.., columns k0..k1
    const int64_t mp = n - k1;
    GPL_CUDA(cudaStreamWaitEvent(main, ctx->ev_diag, 0));
    GPL_TRY(trsm_right_rec(f, true, mp, w, Akk, lda, P, lda, k0 / NB));
    const int64_t w1 = (n - k1) < LA_NB ? (n - k1) : LA_NB;
    const int64_t k2 = k1 + w1;
    // block column k+1 (rows k1.., columns k1..k2): its diagonal block is needed first
    GPL_TRY(dgemm(ctx, false, true, mp, w1, w, -1.0, P, lda, P, lda, 1.0, A + k1 * lda + k1, lda, false));
    GPL_CUDA(cudaEventRecord(ctx->ev_col, main));
    GPL_CUDA(cudaStreamWaitEvent(side, ctx->ev_col, 0));
    ctx->stream = side;
    rc = potrf_rec(f, A + k1 * lda + k1, lda, w1, k1 / NB);
    ctx->stream = main;
    GPL_TRY(rc);
    GPL_CUDA(cudaEventRecord(ctx->ev_diag, side));
    // rest of the trailing matrix (lower triangle), concurrently with the side stream
    if (k2 < n) {
      double* P2 = A + k0 * lda + k2;
      GPL_TRY(dgemm(ctx, false, true, n - k2, n - k2, w, -1.0, P2, lda, P2, lda, 1.0,
                    A + k2 * lda + k2, lda, true));
    }
  }
  return GPL_OK;
}

}  // namespace

int ensure_dinv(Ctx* ctx, int64_t n) {
  const int64_t nblk = (n + NB - 1) / NB;
  if (nblk > ctx->dinv_blocks) {
    if (ctx->dinv) GPL_CUDA(cudaFree(ctx->dinv));
    ctx->dinv = nullptr;
    ctx->dinv_blocks = 0;
    GPL_CUDA(cudaMalloc(&ctx->dinv, (size_t)nblk * NB * NB * sizeof(double)));
    ctx->dinv_blocks = nblk;
  }
  return GPL_OK;
}

int dpotrf_with(Ctx* ctx, double* dinv, int64_t n, double* A, int64_t lda, int32_t* info_host,
                double* logdet_host) {
  if (n <= 0) {
    if (info_host) *info_host = 0;
    if (logdet_host) *logdet_host = 0.0;
    return GPL_OK;
  }
  if ((lda & 1) || ((uintptr_t)A & 15)) {
    set_error("dpotrf: A must be 16-byte aligned with an even leading dimension");
    return GPL_ERR_BAD_ARG;
  }
  const int32_t big = 0x7fffffff;
  GPL_CUDA(cudaMemcpyAsync(ctx->d_info, &big, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  Fac f{ctx, dinv};
  // Look-ahead (diagonal-block chain of step k+1 on a high-priority side stream) is the default:
  // 372 ms vs 384 ms for the plain recursion at n = 32768 once the GEMM epilogue stopped penalising
  // the K = 1024..2048 blocked updates (it had measured 409 vs 403 ms before); GPL_LOOKAHEAD=0
  // disables it.
  static const int lookahead_on = getenv("GPL_LOOKAHEAD") ? atoi(getenv("GPL_LOOKAHEAD")) : 1;
  static const int ll_off = getenv("GPL_NO_LEFTLOOKING") ? 1 : 0;
  // Left-looking potrf measured 436 ms vs 403 ms for the recursion at n = 32768 (its panel TRSMs and
  // 1024-block factorisations are exposed), so it is opt-in; left-looking trtri is a win (370 vs 411 ms)
  // and is the default.
  static const int potrf_ll_on = getenv("GPL_POTRF_LL") ? atoi(getenv("GPL_POTRF_LL")) : 0;
  if (n >= 4 * la_nb_for(n) && ctx->panel_stream && lookahead_on && !ctx->profile) {
    GPL_TRY(potrf_lookahead(f, A, lda, n));
  } else if (n >= 4 * LL_NB && !ll_off && potrf_ll_on) {
    GPL_TRY(potrf_ll(f, A, lda, n));
  } else {
    GPL_TRY(potrf_rec(f, A, lda, n, 0));
  }
  diag_logsum_kernel<<<1, 1024, 0, ctx->stream>>>(A, lda, n, ctx->d_scal);
  GPL_LAUNCH_CHECK(ctx);
  int32_t info = 0;
  double ld = 0.0;
  GPL_CUDA(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  GPL_CUDA(cudaMemcpyAsync(&ld, ctx->d_scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (info == big) info = 0;
  if (info_host) *info_host = info;
  if (logdet_host) *logdet_host = ld;
  if (info != 0) {
    set_error("the leading minor of order %d is not positive definite", info);
    return GPL_ERR_NOT_PD;
  }
  return GPL_OK;
}

// Factor an n x n diagonal block on ctx->stream WITHOUT any host synchronisation: a failing pivot is
// reported as pivot0 + (1-based index inside the block) by atomicMin into the device int `info`
// (which the caller initialises to INT_MAX and reads when it chooses to).  Used by the multi-GPU
// panel step, which must not stall the host between panels.
int dpotrf_nosync(Ctx* ctx, double* dinv, int64_t n, double* A, int64_t lda, int* d_info, int pivot0) {
  if (n <= 0) return GPL_OK;
  if ((lda & 1) || ((uintptr_t)A & 15)) {
    set_error("dpotrf: A must be 16-byte aligned with an even leading dimension");
    return GPL_ERR_BAD_ARG;
  }
  Fac f{ctx, dinv, d_info, pivot0};
  return potrf_rec(f, A, lda, n, 0);
}

// out (n x n, ld ldo) <- L^-1 for a factored diagonal block (exact zeros above the diagonal), from its
// 128-block inverses: the recursive triangular inverse on a copy
int dtrtri_block(Ctx* ctx, double* dinv, int64_t n, const double* L, int64_t ldl, double* out, int64_t ldo) {
  if (n <= 0) return GPL_OK;
  GPL_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * ldo, L, sizeof(double) * ldl, sizeof(double) * n, n,
                             cudaMemcpyDeviceToDevice, ctx->stream));
  Fac f{ctx, dinv};
  GPL_TRY(trtri_rec(f, out, ldo, n, 0));
  zero_upper_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)n), 256, 0, ctx->stream>>>(out, ldo, n);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

int dpotrf(Ctx* ctx, int64_t n, double* A, int64_t lda, int32_t* info_host, double* logdet_host) {
  GPL_TRY(ensure_dinv(ctx, n));
  return dpotrf_with(ctx, ctx->dinv, n, A, lda, info_host, logdet_host);
}

int dtrsm_right_with(Ctx* ctx, double* dinv, bool trans, int64_t m, int64_t n, const double* L,
                     int64_t ldl, double* B, int64_t ldb) {
  Fac f{ctx, dinv};
  return trsm_right_rec(f, trans, m, n, L, ldl, B, ldb, 0);
}

int dtrsm_right(Ctx* ctx, bool trans, int64_t m, int64_t n, const double* L, int64_t ldl,
                double* B, int64_t ldb, int64_t blk0) {
  Fac f{ctx, ctx->dinv};
  return trsm_right_rec(f, trans, m, n, L, ldl, B, ldb, blk0);
}

int dtrtri_with(Ctx* ctx, double* dinv, int64_t n, double* A, int64_t lda);

int dpotri_with(Ctx* ctx, double* dinv, int64_t n, double* A, int64_t lda) {
  Fac f{ctx, dinv};
  GPL_TRY(dtrtri_with(ctx, dinv, n, A, lda));
  GPL_TRY(lauum_rec(f, A, lda, n));
  return GPL_OK;
}

int dtrtri_with(Ctx* ctx, double* dinv, int64_t n, double* A, int64_t lda) {
  Fac f{ctx, dinv};
  static const int ll_off = getenv("GPL_NO_LEFTLOOKING") ? 1 : 0;
  if (n >= 4 * LL_NB && !ll_off) {
    double* T = nullptr;
    const int64_t ldt = LL_NB;
    GPL_CUDA(cudaMallocAsync(&T, sizeof(double) * ldt * n, ctx->stream));
    int rc = trtri_ll(f, A, lda, n, T, ldt);
    cudaFreeAsync(T, ctx->stream);
    return rc;
  }
  return trtri_rec(f, A, lda, n, 0);
}

int lauum_oop(Ctx* ctx, int64_t n, const double* X, int64_t ldx, double* Y, int64_t ldy) {
  // Y_ij = sum_{k >= i} X_ki X_kj (i >= j): rows k < i of column-block i are zero
  return dgemm(ctx, true, false, n, n, n, 1.0, X, ldx, X, ldx, 0.0, Y, ldy, true, 0, 1);
}

int dpotri(Ctx* ctx, int64_t n, double* A, int64_t lda) {
  return dpotri_with(ctx, ctx->dinv, n, A, lda);
}

}  // namespace gpl
