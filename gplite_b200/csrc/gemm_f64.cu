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

#include <stdlib.h>

#include <atomic>
#include <type_traits>

namespace gpl {

namespace {


// CTA tile BM x BN, warp tile WM x WN.  Three configurations: L (128x128, 8 warps of 32x64) for
// large updates, M (64x64, 4 warps of 32x32) and S (32x32, 4 warps of 16x16) so that the small
// GEMMs near the diagonal of the recursive factorisation still spread over many SMs -- one SM
// only issues 64 fp64 FMA/clk, so a 128^3 product on a single CTA costs ~17 us.
template <int BM_, int BN_, int WM_, int WN_, int BK_, int STAGES_>
struct Cfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, BK = BK_, STAGES = STAGES_;
  static constexpr int LDK = BK + 4;                   // [mn][k] stride (== 4 mod 16)
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int MI = WM / 8, NI = WN / 8;
  static constexpr int LDA = BM + 4, LDB = BN + 4;     // [k][mn] strides (== 4 mod 16)
  static constexpr int OPER_A = (BM * LDK > BK * LDA) ? BM * LDK : BK * LDA;
  static constexpr int OPER_B = (BN * LDK > BK * LDB) ? BN * LDK : BK * LDB;
  static constexpr int SMEM_BYTES = STAGES * (OPER_A + OPER_B) * (int)sizeof(double);
};
using CfgL = Cfg<128, 128, 32, 64, 32, 3>;
using CfgL16 = Cfg<128, 128, 32, 32, 32, 3>;   // same CTA tile, 16 warps (4 per scheduler) of 32x32
// (A 128 x 64 tile with two CTAs per SM -- one CTA's prologue / epilogue under the other's main loop -- was
// measured in round 2 and did not pay: potrf 395-412 vs 392-403 ms, lauum 330-331 vs 332 ms, trtri slower;
// profiles/r02_phase_timing_l2x_experiment.log.  The lower-triangular mode below accepts BM != BN since then.)
using CfgM = Cfg<64, 64, 32, 32, 16, 4>;
using CfgS = Cfg<32, 32, 16, 16, 16, 4>;
// In-place products (C aliases an operand) are only race-free when one CTA owns the whole
// K-extent of what it overwrites: full-width row panels (C aliases A, N <= 128) or full-height
// column panels (C aliases B, M <= 128).
using CfgR = Cfg<32, 128, 16, 32, 16, 4>;
using CfgC = Cfg<128, 32, 32, 16, 16, 4>;

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

// Per-thread cp.async plan for one R x BK operand tile (NT threads).  KCONTIG: global is
// contiguous along k (element (mn,k) at g[k + mn*ld]) -> smem [mn][k], row stride BK + 4.
// Otherwise contiguous along mn (g[mn + k*ld]) -> smem [k][mn], row stride R + 4.  All address
// arithmetic that does not depend on the k-tile is done once in init(); per k-tile a chunk
// costs one 64-bit add.  Out-of-range parts are zero-filled by cp.async's src-size.
template <bool KCONTIG, int R, int NT, int BK>
struct TileLoader {
  static constexpr int CPR = KCONTIG ? BK / 2 : R / 2;      // 16-byte chunks per smem row
  static constexpr int ROWS = KCONTIG ? R : BK;             // smem rows
  static constexpr int RSTEP = NT / CPR;                    // rows covered per iteration
  static constexpr int ITERS = ROWS / RSTEP;
  static constexpr int LDS_ = KCONTIG ? BK + 4 : R + 4;
  static_assert(NT % CPR == 0 && ROWS % RSTEP == 0, "tile/thread mismatch");
  const double* base;
  const double* p;        // this thread's chunk of iteration 0 in the current k-tile
  int64_t step;           // doubles between iterations
  int64_t kadv;           // doubles per k-tile
  int soff;               // smem offset (doubles) of iteration 0
  int r0, c2;             // first row, chunk column (in doubles)
  int fixed_bytes;        // !KCONTIG: valid bytes along mn (constant)
  unsigned rowmask;       // KCONTIG: bit i set iff row of iteration i is inside the matrix

  __device__ __forceinline__ void init(const double* g, int64_t ld, int mn0, int mn_lim, int tid) {
    base = g;
    r0 = tid / CPR;
    c2 = 2 * (tid % CPR);
    soff = r0 * LDS_ + c2;
    if (KCONTIG) {
      p = g + (int64_t)(mn0 + r0) * ld + c2;
      step = (int64_t)RSTEP * ld;
      kadv = BK;
      rowmask = 0;
#pragma unroll
      for (int i = 0; i < ITERS; ++i)
        if (mn0 + r0 + RSTEP * i < mn_lim) rowmask |= 1u << i;
      fixed_bytes = 0;
    } else {
      p = g + (int64_t)r0 * ld + mn0 + c2;
      step = (int64_t)RSTEP * ld;
      kadv = (int64_t)BK * ld;
      int mb = (mn_lim - (mn0 + c2)) * 8;
      fixed_bytes = mb < 0 ? 0 : (mb > 16 ? 16 : mb);
      rowmask = 0;
    }
  }
  // Chunk i of the current k-tile.  FAST: the CTA's tile is interior and K is a multiple of BK, so
  // no bounds logic is needed (one LDGSTS + one 64-bit add).  krem = K - k0 (> 0).
  template <bool FAST>
  __device__ __forceinline__ void issue_chunk(double* s, int krem, int i, const double*& q) const {
    int nb = 16;
    if (!FAST) {
      if (KCONTIG) {
        int kb = (krem - c2) * 8;
        kb = kb < 0 ? 0 : (kb > 16 ? 16 : kb);
        nb = ((rowmask >> i) & 1u) ? kb : 0;
      } else {
        nb = (r0 + RSTEP * i < krem) ? fixed_bytes : 0;
      }
    }
    cp_async16(s + soff + i * RSTEP * LDS_, (FAST || nb) ? q : base, nb);
    q += step;
  }
  __device__ __forceinline__ void advance() { p += kadv; }
};

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
  int kskip;      // 0: full k range; 1: k starts at the tile's first row (A^T lower-triangular in k);
                  // 2: k starts at the tile's first column (B lower-triangular);
                  // 3: k ends at the tile's last row (A lower-triangular: A[m][k] = 0 for k > m)
  int tiles_m, tiles_n;
  // Staircase mode (stair_nb > 0), used by the block-cyclic multi-GPU updates: C's columns come in
  // blocks of stair_nb; block c only exists from row stair_first + c * stair_step on (the rows above
  // are the never-referenced upper part of a block column of the slab), and its B rows sit at
  // n + c * stair_bskip in the B operand (the panel rows of every P-th block, read in place).
  int stair_nb, stair_first, stair_step, stair_bskip;
};

template <bool TA, bool TB, class CF>
__global__ void __launch_bounds__(CF::THREADS) gemm_f64_kernel(const GemmArgs p) {
  extern __shared__ __align__(16) double smem[];
  constexpr int BM = CF::BM, BN = CF::BN, MI = CF::MI, NI = CF::NI;
  constexpr int BK = CF::BK, STAGES = CF::STAGES, LDK = CF::LDK;
  constexpr int LDA = CF::LDA, LDB = CF::LDB;
  constexpr int STAGE = CF::OPER_A + CF::OPER_B;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp % CF::WARPS_M, wn = warp / CF::WARPS_M;

  // Grouped rasterisation: consecutive CTAs walk GROUP_M row-tiles (1024 rows) x all column-tiles, so
  // the ~148 tiles resident at any time share ~8 + 18 operand panels instead of 1 + 148 (DRAM traffic of the
  // plain column-major order was 811 GB per energy+gradient evaluation at n = 32768).  In lower
  // mode the same order runs over the full square of tiles and CTAs above the diagonal exit.
  int bm, bn;
  {
    constexpr int GROUP_M = 1024 / BM;     // 1024 rows per group for every tile size
    const int L = blockIdx.x;
    const int per_group = GROUP_M * p.tiles_n;
    const int group = L / per_group;
    const int first_m = group * GROUP_M;
    const int gsz = (p.tiles_m - first_m) < GROUP_M ? (p.tiles_m - first_m) : GROUP_M;
    const int r = L - group * per_group;
    bm = first_m + r % gsz;
    bn = r / gsz;
  }
  const int m0 = bm * BM, n0 = bn * BN;
  if (p.lower && n0 >= m0 + BM) return;      // the whole tile lies above the diagonal (BM != BN allowed)
  int nb0 = n0, nb_lim = p.N;        // B-operand coordinates of this tile's columns
  if (p.stair_nb > 0) {
    const int cblk = n0 / p.stair_nb;
    if (m0 + BM <= p.stair_first + cblk * p.stair_step) return;
    nb0 = n0 + cblk * p.stair_bskip;
    nb_lim = p.N + cblk * p.stair_bskip;
  }

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  int ktiles = (p.K + BK - 1) / BK;
  // triangular operands: the leading (or trailing) k-tiles of this output tile only multiply zeros
  const int kt0 = p.kskip == 1 ? m0 / BK : (p.kskip == 2 ? n0 / BK : 0);
  if (p.kskip == 3) {
    const int kend = (m0 + BM + BK - 1) / BK;
    ktiles = kend < ktiles ? kend : ktiles;
  }

  // A operand: TA -> stored K x M, contiguous along k.  The B operand of the product is
  // op(B)(k,n): !TB -> stored K x N, contiguous along k.
  TileLoader<TA, BM, CF::THREADS, BK> ldA;
  TileLoader<!TB, BN, CF::THREADS, BK> ldB;
  ldA.init(p.A, p.lda, m0, p.M, tid);
  ldB.init(p.B, p.ldb, nb0, nb_lim, tid);
  ldA.p += (int64_t)kt0 * ldA.kadv;
  ldB.p += (int64_t)kt0 * ldB.kadv;
  // interior tile and K a multiple of BK: no bounds logic in the loads (CTA-uniform)
  const bool fast = (m0 + BM <= p.M) && (n0 + BN <= p.N) && (p.K % BK == 0);
  constexpr int KS = BK / 4;                       // k4 steps per tile
  constexpr int ITA = decltype(ldA)::ITERS, ITB = decltype(ldB)::ITERS;

  // whole-tile issue (prologue)
  auto issue_tile = [&](int kt) {
    double* sA = smem + ((kt - kt0) % STAGES) * STAGE;
    double* sB = sA + CF::OPER_A;
    const int krem = p.K - kt * BK;
    const double* qa = ldA.p;
    const double* qb = ldB.p;
#pragma unroll
    for (int i = 0; i < ITA; ++i) ldA.template issue_chunk<false>(sA, krem, i, qa);
#pragma unroll
    for (int i = 0; i < ITB; ++i) ldB.template issue_chunk<false>(sB, krem, i, qb);
    ldA.advance();
    ldB.advance();
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (kt0 + s < ktiles) issue_tile(kt0 + s);
    cp_async_commit();
  }

  // main loop: the cp.async chunks of tile kt + STAGES - 1 are spread over the k4 steps of tile kt,
  // so that the two warps of a scheduler do not both leave the DMMA pipe idle after the barrier
  auto mainloop = [&](auto fast_tag) {
    constexpr bool FAST = decltype(fast_tag)::value;
    for (int kt = kt0; kt < ktiles; ++kt) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      const int ktn = kt + STAGES - 1;
      const bool do_load = ktn < ktiles;
      double* nA = smem + ((ktn - kt0) % STAGES) * STAGE;
      double* nB = nA + CF::OPER_A;
      const int krem = p.K - ktn * BK;
      const double* qa = ldA.p;
      const double* qb = ldB.p;

      const double* sA = smem + ((kt - kt0) % STAGES) * STAGE;
      const double* sB = sA + CF::OPER_A;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        if (do_load) {
#pragma unroll
          for (int i = 0; i < ITA; ++i)
            if (i * KS / ITA == ks) ldA.template issue_chunk<FAST>(nA, krem, i, qa);
#pragma unroll
          for (int i = 0; i < ITB; ++i)
            if (i * KS / ITB == ks) ldB.template issue_chunk<FAST>(nB, krem, i, qb);
        }
        const int kk = ks * 4 + t;
        double a[MI], b[NI];
#pragma unroll
        for (int im = 0; im < MI; ++im) {
          const int m = wm * CF::WM + im * 8 + g;
          a[im] = TA ? sA[m * LDK + kk] : sA[kk * LDA + m];
        }
#pragma unroll
        for (int in = 0; in < NI; ++in) {
          const int n = wn * CF::WN + in * 8 + g;
          b[in] = TB ? sB[kk * LDB + n] : sB[n * LDK + kk];
        }
#pragma unroll
        for (int im = 0; im < MI; ++im)
#pragma unroll
          for (int in = 0; in < NI; ++in) dmma(acc[im][in][0], acc[im][in][1], a[im], b[in]);
      }
      if (do_load) {
        ldA.advance();
        ldB.advance();
      }
      cp_async_commit();
    }
  };
  if (fast) mainloop(std::true_type{});
  else mainloop(std::false_type{});
  cp_async_wait<0>();

  // epilogue: thread owns C[row = .. + g][col = .. + 2t, 2t+1] of each 8x8 sub-tile.  The C values of
  // a batch of sub-tiles are all loaded before the first store: interleaved load/store pairs cannot
  // be reordered by the compiler (possible aliasing), which serialised 64 L2 round trips per thread
  // (~16 us per 128x128 tile, as much as 7 k-tiles of the main loop).
  const bool diag_tile = p.lower && (n0 + BN > m0);      // the tile reaches above the diagonal
  const bool rmw = p.beta != 0.0;
  constexpr int IMB = (MI % 2 == 0) ? 2 : 1;
#pragma unroll
  for (int im0 = 0; im0 < MI; im0 += IMB) {
    double cv[IMB][NI][2];
    unsigned ok = 0;
#pragma unroll
    for (int ii = 0; ii < IMB; ++ii) {
      const int row = m0 + wm * CF::WM + (im0 + ii) * 8 + g;
#pragma unroll
      for (int in = 0; in < NI; ++in) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = n0 + wn * CF::WN + in * 8 + 2 * t + c;
          const bool v = row < p.M && col < p.N && !(diag_tile && col > row);
          if (v) ok |= 1u << ((ii * NI + in) * 2 + c);
          cv[ii][in][c] = (v && rmw) ? p.C[(int64_t)col * p.ldc + row] : 0.0;
        }
      }
    }
#pragma unroll
    for (int ii = 0; ii < IMB; ++ii) {
      const int row = m0 + wm * CF::WM + (im0 + ii) * 8 + g;
#pragma unroll
      for (int in = 0; in < NI; ++in) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = n0 + wn * CF::WN + in * 8 + 2 * t + c;
          if ((ok >> ((ii * NI + in) * 2 + c)) & 1u)
            p.C[(int64_t)col * p.ldc + row] = p.alpha * acc[im0 + ii][in][c] + p.beta * cv[ii][in][c];
        }
      }
    }
  }
}

template <bool TA, bool TB, class CF>
int launch(Ctx* ctx, GemmArgs a) {
  // per device (the attribute is per-context).  Atomic flags: two host threads driving two contexts may
  // both set the attribute (idempotent), but neither can observe the flag before its own call returned.
  static std::atomic<bool> configured[64];
  std::atomic<bool>& cfgd = configured[ctx->device & 63];
  if (!cfgd.load(std::memory_order_acquire)) {
    GPL_CUDA(cudaFuncSetAttribute(gemm_f64_kernel<TA, TB, CF>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
    cfgd.store(true, std::memory_order_release);
  }
  a.tiles_m = (a.M + CF::BM - 1) / CF::BM;
  a.tiles_n = (a.N + CF::BN - 1) / CF::BN;
  const int grid = a.tiles_m * a.tiles_n;
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
  gemm_f64_kernel<TA, TB, CF><<<grid, CF::THREADS, CF::SMEM_BYTES, ctx->stream>>>(a);
  GPL_LAUNCH_CHECK(ctx);
  if (ctx->profile) {
    GPL_CUDA(cudaEventRecord(e1, ctx->stream));
    ctx->prof_flops += 2.0 * a.M * a.N * (double)a.K * (a.lower ? 0.5 : 1.0) *
                       (a.kskip ? (a.lower ? 1.0 / 3.0 : 0.5) : 1.0);
    ctx->prof_launches++;
    ctx->prof_recs.push_back({a.M, a.N, a.K, a.lower, a.kskip, CF::BM * 1000 + CF::BN});
  }
  return GPL_OK;
}

template <bool TA, bool TB>
int dispatch(Ctx* ctx, const GemmArgs& a, int alias) {
  if (alias) {
    if (((alias & 1) && a.N > 128) || ((alias & 2) && a.M > 128)) {
      set_error("dgemm: in-place product needs the aliased extent <= 128");
      return GPL_ERR_BAD_ARG;
    }
    if (alias == 3) return launch<TA, TB, CfgL>(ctx, a);
    if (alias == 1)
      return (a.M >= 64 * ctx->sm_count) ? launch<TA, TB, CfgL>(ctx, a) : launch<TA, TB, CfgR>(ctx, a);
    return (a.N >= 64 * ctx->sm_count) ? launch<TA, TB, CfgL>(ctx, a) : launch<TA, TB, CfgC>(ctx, a);
  }
  auto ntiles = [&](int bm, int bn) {
    const int64_t tm = (a.M + bm - 1) / bm, tn = (a.N + bn - 1) / bn;
    return a.lower ? tm * (tm + 1) / 2 : tm * tn;
  };
  // Tile choice by a small analytical cost model, fitted to a measured table of 480 shapes x 3 tiles
  // (tools/gemm_tune.py, tools/gemm_model.py; regret vs the per-shape best: 0.4 %).  Per config:
  // r co-resident CTAs per SM, s1 = us per unit of k for one CTA alone on an SM, rate = flop/us per
  // SM at full residency, t0 = prologue + epilogue that nothing overlaps.  A single wave spreads
  // breadth-first over the SMs; after that, slots free up r at a time, so L and M run in waves of
  // r * SMs tiles while the many small S CTAs de-synchronise into a near-continuous refill.
  const int64_t sms = ctx->sm_count;
  if (ctx->gemm_force == 1) return launch<TA, TB, CfgL>(ctx, a);
  if (ctx->gemm_force == 2) return launch<TA, TB, CfgM>(ctx, a);
  if (ctx->gemm_force == 3) return launch<TA, TB, CfgS>(ctx, a);
  const double k = a.kskip ? 0.5 * a.K : (double)a.K;
  auto cost = [&](int b, int r, double s1, double rate, double t0, bool continuous) {
    const int64_t T = ntiles(b, b);
    const double w = 2.0 * b * b;
    if (T <= r * sms) {
      const double q = (double)((T + sms - 1) / sms);
      return t0 + k * fmax(s1, q * w / rate);
    }
    if (continuous) return t0 + k * (double)((T + sms - 1) / sms) * w / rate;
    const double waves = (double)((T + r * sms - 1) / (r * sms));
    return t0 + waves * k * r * w / rate + (r == 1 ? t0 * (waves - 1.0) : 0.0);
  };
  const double cost_l = cost(128, 1, 0.139, 234e3, 9.0, false);
  const double cost_m = cost(64, 2, 0.044, 231e3, 4.5, false);
  const double cost_s = cost(32, 5, 0.016, 220e3, 3.0, true);
  static const int use_l16 = getenv("GPL_GEMM_L16") ? atoi(getenv("GPL_GEMM_L16")) : 0;
  if (cost_l <= cost_m && cost_l <= cost_s)
    return use_l16 ? launch<TA, TB, CfgL16>(ctx, a) : launch<TA, TB, CfgL>(ctx, a);
  if (cost_m <= cost_s) return launch<TA, TB, CfgM>(ctx, a);
  return launch<TA, TB, CfgS>(ctx, a);
}

}  // namespace

int dgemm(Ctx* ctx, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha,
          const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
          int64_t ldc, bool lower, int alias, int kskip, const GemmStair* stair) {
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
  a.kskip = kskip;
  a.tiles_m = a.tiles_n = 0;
  a.stair_nb = a.stair_first = a.stair_step = a.stair_bskip = 0;
  if (stair && stair->nb > 0) {
    if (stair->nb % 128 || alias || lower || kskip) {
      set_error("dgemm: staircase mode needs column blocks that are multiples of 128 and a plain product");
      return GPL_ERR_BAD_ARG;
    }
    a.stair_nb = (int)stair->nb; a.stair_first = (int)stair->first; a.stair_step = (int)stair->step;
    a.stair_bskip = (int)stair->bskip;
  }
  if (ta && tb) {
    set_error("dgemm: TT product not implemented (never needed on the hot path)");
    return GPL_ERR_BAD_ARG;
  }
  if (!ta && !tb) return dispatch<false, false>(ctx, a, alias);
  if (!ta && tb) return dispatch<false, true>(ctx, a, alias);
  return dispatch<true, false>(ctx, a, alias);
}

}  // namespace gpl
