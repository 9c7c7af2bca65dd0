// FP64 GEMM / SYRK-shaped update on the Blackwell FP64 tensor path.
//
// On sm_100a the only FP64 tensor primitive is the warp-level mma.sync m8n8k4 (SASS
// DMMA.8x8x4); tcgen05.mma has no f64 kind, so accumulators live in registers and operands
// come from shared memory with ordinary LDS.64.  The kernel is a 128x128x16 CTA tile,
// 8 warps (4 along m x 2 along n, 32x64 per warp = 32 DMMA per k4 step), fed by a 4-stage
// cp.async ring.  Both operand layouts are supported per side (contiguous along m/n or
// along k) so that NN, NT, TN products of column-major matrices need no transposition pass;
// shared-memory row strides (132 / 20 doubles) make every fragment load conflict-free.
//
// Replaces the reference's BLAS calls behind chol()/forwardsolve()/%*% (R/laplace.R:27-30,
// R/ep.R:38-42, R/linear_algebra.R:24, R/gp_pred.R:218).
#include "common.cuh"

namespace gpl {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, THREADS = 256;
constexpr int LDM = 132;            // smem row stride of the [k][mn] layout (== 4 mod 16)
constexpr int LDK = 20;             // smem row stride of the [mn][k] layout (== 4 mod 16)
constexpr int OPER = 128 * LDK;     // doubles per operand per stage (>= 16*LDM)
constexpr int SMEM_BYTES = STAGES * 2 * OPER * (int)sizeof(double);

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, int nbytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(nbytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Stage one 128 x 16 operand tile.  KCONTIG: global is contiguous along k (element (mn,k) at
// g[k + mn*ld]) -> smem [mn][k].  Otherwise contiguous along mn (g[mn + k*ld]) -> smem [k][mn].
// Out-of-range parts are zero-filled by cp.async's src-size.
template <bool KCONTIG>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ g, int64_t ld,
                                          int mn0, int mn_lim, int k0, int k_lim, int tid) {
  if (KCONTIG) {
    const int k2 = tid & 7;
    const int k = k0 + 2 * k2;
    int kb = (k_lim - k) * 8;
    kb = kb < 0 ? 0 : (kb > 16 ? 16 : kb);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (tid >> 3) + 32 * i;
      const int mn = mn0 + r;
      const int nb = (mn < mn_lim) ? kb : 0;
      const double* src = nb ? (g + (int64_t)mn * ld + k) : g;
      cp_async16(s + r * LDK + 2 * k2, src, nb);
    }
  } else {
    const int m2 = tid & 63;
    const int mn = mn0 + 2 * m2;
    int mb = (mn_lim - mn) * 8;
    mb = mb < 0 ? 0 : (mb > 16 ? 16 : mb);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kr = (tid >> 6) + 4 * i;
      const int k = k0 + kr;
      const int nb = (k < k_lim) ? mb : 0;
      const double* src = nb ? (g + (int64_t)k * ld + mn) : g;
      cp_async16(s + kr * LDM + 2 * m2, src, nb);
    }
  }
}

struct GemmArgs {
  int M, N, K;
  const double* A;
  int64_t lda;
  const double* B;
  int64_t ldb;
  double* C;
  int64_t ldc;
  double alpha, beta;
  int lower;
  int tiles_m, tiles_n;
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(THREADS, 1) gemm_f64_kernel(const GemmArgs p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;

  int bm, bn;
  if (p.lower) {
    const int L = blockIdx.x;
    int bi = (int)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
    while ((int64_t)bi * (bi + 1) / 2 > L) --bi;
    while ((int64_t)(bi + 1) * (bi + 2) / 2 <= L) ++bi;
    bm = bi;
    bn = L - bi * (bi + 1) / 2;
  } else {
    bm = blockIdx.x % p.tiles_m;
    bn = blockIdx.x / p.tiles_m;
  }
  const int m0 = bm * BM, n0 = bn * BN;

  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int ktiles = (p.K + BK - 1) / BK;

  auto issue = [&](int kt) {
    double* sA = smem + (kt % STAGES) * 2 * OPER;
    double* sB = sA + OPER;
    // A operand: TA -> stored K x M, contiguous along k
    load_tile<TA>(sA, p.A, p.lda, m0, p.M, kt * BK, p.K, tid);
    // B operand of the product is op(B)(k,n): !TB -> stored K x N, contiguous along k
    load_tile<!TB>(sB, p.B, p.ldb, n0, p.N, kt * BK, p.K, tid);
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ktiles) issue(s);
    cp_async_commit();
  }

  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kt + STAGES - 1 < ktiles) issue(kt + STAGES - 1);
    cp_async_commit();

    const double* sA = smem + (kt % STAGES) * 2 * OPER;
    const double* sB = sA + OPER;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      const int kk = ks * 4 + t;
      double a[4], b[8];
#pragma unroll
      for (int im = 0; im < 4; ++im) {
        const int m = wm * 32 + im * 8 + g;
        a[im] = TA ? sA[m * LDK + kk] : sA[kk * LDM + m];
      }
#pragma unroll
      for (int in = 0; in < 8; ++in) {
        const int n = wn * 64 + in * 8 + g;
        b[in] = TB ? sB[kk * LDM + n] : sB[n * LDK + kk];
      }
#pragma unroll
      for (int im = 0; im < 4; ++im)
#pragma unroll
        for (int in = 0; in < 8; ++in) dmma(acc[im][in][0], acc[im][in][1], a[im], b[in]);
    }
  }
  cp_async_wait<0>();

  // epilogue: thread owns C[row = .. + g][col = .. + 2t, 2t+1] of each 8x8 sub-tile
  const bool diag_tile = p.lower && (bm == bn);
#pragma unroll
  for (int im = 0; im < 4; ++im) {
    const int row = m0 + wm * 32 + im * 8 + g;
    if (row >= p.M) continue;
#pragma unroll
    for (int in = 0; in < 8; ++in) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = n0 + wn * 64 + in * 8 + 2 * t + c;
        if (col >= p.N) continue;
        if (diag_tile && col > row) continue;
        double* dst = p.C + (int64_t)col * p.ldc + row;
        double v = p.alpha * acc[im][in][c];
        if (p.beta != 0.0) v += p.beta * (*dst);
        *dst = v;
      }
    }
  }
}

template <bool TA, bool TB>
int launch(Ctx* ctx, const GemmArgs& a, int grid) {
  static bool configured = false;
  if (!configured) {
    GPL_CUDA(cudaFuncSetAttribute(gemm_f64_kernel<TA, TB>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->profile) {
    while (ctx->ev_pool.size() < ctx->ev_used + 2) {
      cudaEvent_t e;
      GPL_CUDA(cudaEventCreate(&e));
      ctx->ev_pool.push_back(e);
    }
    e0 = ctx->ev_pool[ctx->ev_used++];
    e1 = ctx->ev_pool[ctx->ev_used++];
    GPL_CUDA(cudaEventRecord(e0, ctx->stream));
  }
  gemm_f64_kernel<TA, TB><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(a);
  GPL_LAUNCH_CHECK(ctx);
  if (ctx->profile) {
    GPL_CUDA(cudaEventRecord(e1, ctx->stream));
    ctx->prof_flops += 2.0 * a.M * a.N * (double)a.K * (a.lower ? 0.5 : 1.0);
    ctx->prof_launches++;
  }
  return GPL_OK;
}

}  // namespace

int dgemm(Ctx* ctx, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha,
          const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
          int64_t ldc, bool lower) {
  if (m <= 0 || n <= 0) return GPL_OK;
  if ((lda & 1) || (ldb & 1) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) {
    set_error("dgemm: operands must be 16-byte aligned with even leading dimensions "
              "(lda=%lld ldb=%lld)", (long long)lda, (long long)ldb);
    return GPL_ERR_BAD_ARG;
  }
  if (lower && m != n) {
    set_error("dgemm: lower-only update needs a square C");
    return GPL_ERR_BAD_ARG;
  }
  GemmArgs a;
  a.M = (int)m; a.N = (int)n; a.K = (int)k;
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc;
  a.alpha = alpha; a.beta = beta; a.lower = lower ? 1 : 0;
  a.tiles_m = (int)((m + BM - 1) / BM);
  a.tiles_n = (int)((n + BN - 1) / BN);
  int grid = lower ? a.tiles_m * (a.tiles_m + 1) / 2 : a.tiles_m * a.tiles_n;
  if (ta && tb) {
    set_error("dgemm: TT product not implemented (never needed on the hot path)");
    return GPL_ERR_BAD_ARG;
  }
  if (!ta && !tb) return launch<false, false>(ctx, a, grid);
  if (!ta && tb) return launch<false, true>(ctx, a, grid);
  return launch<true, false>(ctx, a, grid);
}

}  // namespace gpl
