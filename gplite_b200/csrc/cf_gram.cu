// Covariance-matrix construction, second generation: EVERY factor's pairwise statistic through the FP64
// tensor pipe, for any sum of products of {cf_const, cf_lin, cf_sexp, cf_matern32, cf_matern52, cf_nn}
// and for rectangular blocks (Kt, Kxz) as well as symmetric K.  (cf_kernels.cu keeps the first
// generation -- direct differences on the fp64 pipe -- as the fallback for very wide feature sets.)
//
// The first K build was bound by fp64 instruction ISSUE, not by HBM: ~28 fp64 instructions per entry for
// cf_matern52 (distance fix-up, sqrt, exp, polynomial) against 62 per clock per SM.  What this kernel does
// about it, per 128 x 128 tile of K (256 threads, warp tile 32 x 32 or 32 x 16):
//   * stationary kinds: the staged rows carry two extra "feature" columns, [a, |a|^2, 1] . [-2 b, 1, |b|^2],
//     so one DMMA.8x8x4 chain delivers |a - b|^2 itself -- no add / fma fix-up per entry -- and the rows are
//     pre-scaled (1 / (sqrt2 l), sqrt3 / l, sqrt5 / l) so that the contraction is directly the argument the
//     kernel function needs (-log k for sexp, t^2 for the Matern kinds): no per-entry scaling multiply;
//   * cf_lin / cf_nn: their dot products are Gram contractions on the same staged tile (src/cf.cpp:149-158
//     recomputes the two norms per PAIR; here they are per row, once per tile);
//   * exp(-t): 256-entry table 2^(j/256) in shared memory, degree-3 minimax polynomial on |r| <= ln2/512
//     (error 2.4e-14), exponent applied by an integer add; clamps done on the integer pipe;
//   * sqrt: MUFU.RSQ64H seed + one third-order step (5 fp64 instructions).
// cf_matern52 drops to ~16 fp64 instructions per entry.  Products and sums are folded in registers, one
// factor at a time, so nothing n x n is ever materialised besides K (the reference materialises one
// matrix per cf plus one per `*` / `+`, R/cf.R:408,494).
//
// Accuracy: rows are shifted by a common reference row first (distances are translation invariant) and a
// tile whose scaled norms could cost more than ~5e-13 relative in K through cancellation falls back to
// direct differences (CTA-uniform branch), so the 1e-12 parity with src/cf.cpp:9-162 holds.
#include "cf.cuh"
#include "gram_math.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

namespace gpl {

namespace {

constexpr int TILE = 128;
constexpr int GTHREADS = 256;
constexpr int GLD = TILE + 4;        // smem row stride, == 4 mod 16: conflict-free fragment loads
constexpr int MAX_ROWS = 48;         // staged rows per side: 2 CTAs / SM

struct GFactor {
  int kind;
  int zoff, nf;       // feature columns in Z
  int goff, gk;       // staged rows [goff, goff + gk) take part in the contraction, gk % 4 == 0; row
                      // goff + gk is an auxiliary per-row scalar (cf_nn)
  int first, last;    // first / last factor of its summand
  int pad;
  double sc;          // staging scale of the stationary kinds
  double magn2, c0, m3;
};

struct GPlan {
  int nfactors, rows;
  GFactor f[MAX_FACTORS];
};

struct GArgs {
  const double* z1;
  const double* z2;
  double* K;
  int64_t ldz1, ldz2, ldk;
  int n1, n2;
  int same;
  int tiles_i, tiles_j;
  double jitter;
};

template <int KIND>
__device__ __forceinline__ double gvalue(const GFactor& f, double s, double rb, double rc,
                                         const double* __restrict__ tab) {
  if (KIND == GPL_CF_CONST) return f.magn2;
  if (KIND == GPL_CF_LIN) return f.magn2 * s;
  if (KIND == GPL_CF_SEXP) return f.magn2 * exp_m(clamp_arg(s), tab);
  if (KIND == GPL_CF_MATERN32) {
    const double u = clamp_lo(s);                          // u = t^2 may be large; only exp's argument is capped
    const double t = clamp_hi(sqrt_pos(u));
    return fma(f.magn2, t, f.magn2) * exp_m(t, tab);
  }
  if (KIND == GPL_CF_MATERN52) {
    const double u = clamp_lo(s);
    const double t = clamp_hi(sqrt_pos(u));
    return fma(f.m3, u, fma(f.magn2, t, f.magn2)) * exp_m(t, tab);
  }
  // cf_nn (src/cf.cpp:149-158): magn^2 (2/pi) asin(2a / sqrt((1 + 2b)(1 + 2c))), a = c0 + x1.x2;
  // rb * rc first: the symmetric matrix is then symmetric bit for bit
  return f.m3 * asin(2.0 * (f.c0 + s) * (rb * rc));
}

__device__ __forceinline__ bool is_stationary(int kind) {
  return kind == GPL_CF_SEXP || kind == GPL_CF_MATERN32 || kind == GPL_CF_MATERN52;
}

// MODE >= 0: the model is ONE factor of that kind.  MODE == -1: any sum of products.  MODE == -2: ONE product
// term (BASELINE config 3's cf_sexp * cf_lin): the running product is the result, so the per-entry `out`
// accumulators of the general case are not needed (a third fewer live registers per entry).
// NSUB: 8-column sub-tiles per warp and pass (4: warp tile 32 x 32, two passes; 2: 32 x 16, four passes).
template <int MODE, int NSUB, int MINB>
__global__ void __launch_bounds__(GTHREADS, MINB)
gram_kernel(const __grid_constant__ GPlan plan, const GArgs a) {
  extern __shared__ double sm[];
  const int R = plan.rows;
  double* Zs1 = sm;                  // [R][GLD]
  double* Zs2 = sm + R * GLD;
  double* tab = sm + 2 * R * GLD;    // [256]
  __shared__ int s_bad, s_nan;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  tab[tid] = kExpTab256[tid];
  if (tid == 0) s_bad = s_nan = 0;
  __syncthreads();

  int bi, bj;
  if (a.same) {
    const int L = blockIdx.x;
    int b = (int)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
    while ((int64_t)b * (b + 1) / 2 > L) --b;
    while ((int64_t)(b + 1) * (b + 2) / 2 <= L) ++b;
    bi = b;
    bj = L - b * (b + 1) / 2;
  } else {
    bi = blockIdx.x % a.tiles_i;
    bj = blockIdx.x / a.tiles_i;
  }
  const int i0 = bi * TILE, j0 = bj * TILE;
  const bool mirror = a.same && (bi != bj);
  const bool diag_tile = a.same && (bi == bj);
  const bool interior = (i0 + TILE <= a.n1) && (j0 + TILE <= a.n2);
  const bool vec_ok = ((a.ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.K) & 15) == 0);

  // ---- stage: features (shifted and scaled for the stationary kinds; -2 folded into the column side)
  for (int fi = 0; fi < plan.nfactors; ++fi) {
    const GFactor& f = plan.f[fi];
    const bool stat = is_stationary(f.kind);
    // stationary kinds: rows goff, goff + 1 are the two norm columns.  They come FIRST so that they are
    // summed into a zero accumulator, |a_i|^2 + |b_j|^2 = |a_j|^2 + |b_i|^2 bit for bit, and the
    // symmetric matrix is exactly symmetric also inside its diagonal tiles.
    const int aug = stat ? 2 : 0;
    for (int idx = tid; idx < (f.gk + 1 - aug) * TILE; idx += GTHREADS) {
      const int j = idx >> 7, r = idx & (TILE - 1);
      double v1 = 0.0, v2 = 0.0;
      if (j < f.nf) {
        const int64_t c1 = (int64_t)(f.zoff + j) * a.ldz1, c2 = (int64_t)(f.zoff + j) * a.ldz2;
        const double ref = stat ? a.z1[c1] : 0.0;
        const double sc = stat ? f.sc : 1.0;
        const int gi = i0 + r, gj = j0 + r;
        if (gi < a.n1) v1 = (a.z1[c1 + gi] - ref) * sc;
        if (gj < a.n2) v2 = (a.z2[c2 + gj] - ref) * sc;
        if (stat) v2 *= -2.0;
        if (v1 != v1 || v2 != v2) s_nan = 1;               // NaN input: the whole tile becomes NaN (below)
      }
      Zs1[(f.goff + aug + j) * GLD + r] = v1;
      Zs2[(f.goff + aug + j) * GLD + r] = v2;
    }
  }
  __syncthreads();
  // ---- per-row norms: the two augmented columns (stationary) or the cf_nn row scalars
  for (int fi = 0; fi < plan.nfactors; ++fi) {
    const GFactor& f = plan.f[fi];
    const bool stat = is_stationary(f.kind);
    if (!stat && f.kind != GPL_CF_NN) continue;
    double* Zs = (tid < TILE) ? Zs1 : Zs2;
    const int r = tid & (TILE - 1);
    const int aug = stat ? 2 : 0;
    double sacc = 0.0;
    for (int j = 0; j < f.nf; ++j) {
      const double v = Zs[(f.goff + aug + j) * GLD + r];
      sacc = fma(v, v, sacc);
    }
    if (stat) {
      if (tid >= TILE) sacc *= 0.25;                      // the column side holds -2 b
      Zs[f.goff * GLD + r] = (tid < TILE) ? sacc : 1.0;
      Zs[(f.goff + 1) * GLD + r] = (tid < TILE) ? 1.0 : sacc;
      // cancellation guard: the relative error of K is ~ 4 eps * (scaled |a|^2 + |b|^2) * |dlog k / du|
      // (d log k / du is 1 for sexp and at most 1/2 for the Matern kinds, also as u -> 0)
      if (!(sacc < 250.0)) s_bad = 1;
    } else {
      Zs[(f.goff + f.gk) * GLD + r] = rsqrt(1.0 + 2.0 * (f.c0 + sacc));
    }
  }
  __syncthreads();
  const bool bad = s_bad != 0;

  constexpr int PASS_COLS = 16 * NSUB;       // two warps along n
#pragma unroll 1
  for (int pass = 0; pass < TILE / PASS_COLS; ++pass) {
    const int cbase = pass * PASS_COLS + wn * (8 * NSUB);
    double out[4][NSUB][2], prod[4][NSUB][2];
    if (MODE == -1) {
#pragma unroll
      for (int im = 0; im < 4; ++im)
#pragma unroll
        for (int in = 0; in < NSUB; ++in) out[im][in][0] = out[im][in][1] = 0.0;
    }
#pragma unroll 1
    for (int fi = 0; fi < (MODE < 0 ? plan.nfactors : 1); ++fi) {
      const GFactor& f = plan.f[fi];
      const int kind = MODE < 0 ? f.kind : MODE;
      const bool stat = is_stationary(kind);
      double acc[4][NSUB][2];
#pragma unroll
      for (int im = 0; im < 4; ++im)
#pragma unroll
        for (int in = 0; in < NSUB; ++in) acc[im][in][0] = acc[im][in][1] = 0.0;
      if (!(bad && stat)) {
        for (int ks = 0; ks < f.gk; ks += 4) {
          const int kk = f.goff + ks + t;
          double av[4], bv[NSUB];
#pragma unroll
          for (int im = 0; im < 4; ++im) av[im] = Zs1[kk * GLD + wm * 32 + im * 8 + g];
#pragma unroll
          for (int in = 0; in < NSUB; ++in) bv[in] = Zs2[kk * GLD + cbase + in * 8 + g];
#pragma unroll
          for (int im = 0; im < 4; ++im)
#pragma unroll
            for (int in = 0; in < NSUB; ++in) dmma_k4(acc[im][in][0], acc[im][in][1], av[im], bv[in]);
        }
      } else {
        // direct differences from the staged (already scaled) rows: a - b = Zs1 + Zs2 / 2
#pragma unroll
        for (int im = 0; im < 4; ++im)
#pragma unroll
          for (int in = 0; in < NSUB; ++in)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int row = wm * 32 + im * 8 + g, col = cbase + in * 8 + 2 * t + c;
              double r2 = 0.0;
              for (int j = 2; j < f.nf + 2; ++j) {
                const double dlt = fma(0.5, Zs2[(f.goff + j) * GLD + col], Zs1[(f.goff + j) * GLD + row]);
                r2 = fma(dlt, dlt, r2);
              }
              acc[im][in][c] = r2;
            }
      }
      // values of this factor for the thread's entries, folded into the running product / sum.  The kind is
      // dispatched ONCE per factor and pass (a CTA-uniform switch around fully unrolled, branch-free entry
      // loops), never per entry.
      auto fold = [&](auto kind_tag) {
        constexpr int K = decltype(kind_tag)::value;
        constexpr bool kstat = K == GPL_CF_SEXP || K == GPL_CF_MATERN32 || K == GPL_CF_MATERN52;
#pragma unroll
        for (int im = 0; im < 4; ++im) {
          const int row = wm * 32 + im * 8 + g;
          const double rb = K == GPL_CF_NN ? Zs1[(f.goff + f.gk) * GLD + row] : 0.0;
#pragma unroll
          for (int in = 0; in < NSUB; ++in) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int col = cbase + in * 8 + 2 * t + c;
              const double rc = K == GPL_CF_NN ? Zs2[(f.goff + f.gk) * GLD + col] : 0.0;
              double s = acc[im][in][c];
              if (kstat && diag_tile && row == col) s = 0.0;        // k(x, x) exactly
              const double v = gvalue<K>(f, s, rb, rc, tab);
              if (MODE == -1) {
                const double pr = f.first ? v : prod[im][in][c] * v;
                prod[im][in][c] = pr;
                if (f.last) out[im][in][c] += pr;
              } else if (MODE == -2) {
                out[im][in][c] = (fi == 0) ? v : out[im][in][c] * v;
              } else {
                out[im][in][c] = v;
              }
            }
          }
        }
      };
      if (MODE >= 0) {
        fold(std::integral_constant<int, (MODE < 0 ? 0 : MODE)>{});
      } else {
        switch (kind) {
          case GPL_CF_CONST: fold(std::integral_constant<int, GPL_CF_CONST>{}); break;
          case GPL_CF_LIN: fold(std::integral_constant<int, GPL_CF_LIN>{}); break;
          case GPL_CF_SEXP: fold(std::integral_constant<int, GPL_CF_SEXP>{}); break;
          case GPL_CF_MATERN32: fold(std::integral_constant<int, GPL_CF_MATERN32>{}); break;
          case GPL_CF_MATERN52: fold(std::integral_constant<int, GPL_CF_MATERN52>{}); break;
          default: fold(std::integral_constant<int, GPL_CF_NN>{}); break;
        }
      }
    }
    // ---- store: thread owns row g, columns 2t, 2t + 1 of each 8 x 8 sub-tile.  The common case (interior
    // tile off the diagonal) is one uniform branch around fully unrolled stores with immediate offsets;
    // the mirrored entry pair is one 16-byte store.
    if (interior && !diag_tile && vec_ok) {
      const int64_t cs = a.ldk;
      double* p0 = a.K + (int64_t)(j0 + cbase + 2 * t) * cs + (i0 + wm * 32 + g);
#pragma unroll
      for (int in = 0; in < NSUB; ++in) {
        double* pc = p0 + (int64_t)(in * 8) * cs;
#pragma unroll
        for (int im = 0; im < 4; ++im) {
          pc[im * 8] = out[im][in][0];
          pc[cs + im * 8] = out[im][in][1];
        }
      }
      if (mirror) {
        double* q0 = a.K + (int64_t)(i0 + wm * 32 + g) * cs + (j0 + cbase + 2 * t);
#pragma unroll
        for (int im = 0; im < 4; ++im) {
          double* qr = q0 + (int64_t)(im * 8) * cs;
#pragma unroll
          for (int in = 0; in < NSUB; ++in)
            *reinterpret_cast<double2*>(qr + in * 8) = make_double2(out[im][in][0], out[im][in][1]);
        }
      }
    } else {
#pragma unroll
      for (int im = 0; im < 4; ++im) {
        const int gi = i0 + wm * 32 + im * 8 + g;
#pragma unroll
        for (int in = 0; in < NSUB; ++in) {
          const int gj = j0 + cbase + in * 8 + 2 * t;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            double v = out[im][in][c];
            if (diag_tile && gi == gj + c) v += a.jitter;
            if (gi < a.n1 && gj + c < a.n2) {
              a.K[(int64_t)(gj + c) * a.ldk + gi] = v;
              if (mirror) a.K[(int64_t)gi * a.ldk + gj + c] = v;
            }
          }
        }
      }
    }
  }
  if (s_nan) {
    // a NaN among this tile's inputs (the clamps above do not propagate it): R's chol() must see it
    __syncthreads();
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int idx = tid; idx < TILE * TILE; idx += GTHREADS) {
      const int gi = i0 + (idx & (TILE - 1)), gj = j0 + (idx >> 7);
      if (gi < a.n1 && gj < a.n2) {
        a.K[(int64_t)gj * a.ldk + gi] = qnan;
        if (mirror) a.K[(int64_t)gi * a.ldk + gj] = qnan;
      }
    }
  }
}

template <int MODE, int NSUB, int MINB>
int launch_gram1(Ctx* ctx, const GPlan& gp, const GArgs& a, int grid) {
  const int smem = (2 * gp.rows * GLD + 256) * (int)sizeof(double);
  GPL_CUDA(cudaFuncSetAttribute(gram_kernel<MODE, NSUB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  gram_kernel<MODE, NSUB, MINB><<<grid, GTHREADS, smem, ctx->stream>>>(gp, a);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

// tuning knob GPL_GRAM_VARIANT: 0 = warp tile 32 x 32, 2 CTAs / SM (default); 1 = 32 x 16, 2 CTAs / SM;
// 2 = 32 x 32, 1 CTA / SM (255 registers); 3 = 32 x 16, 3 CTAs / SM; 4 = 32 x 16, 4 CTAs / SM
template <int MODE, int NSUB>
int launch_gram(Ctx* ctx, const GPlan& gp, const GArgs& a, int grid) {
  // Measured at n = 32768, d = 8 (profiles/r02_kbuild_variants.jsonl): the kernel is latency bound, so
  // occupancy beats per-thread tile size -- cf_matern52 symmetric 2.29 / 1.98 / 2.88 / 1.71 / 1.58 ms for
  // variants 0..4; 4 (warp tile 32 x 16, 64 registers, 4 CTAs = 32 warps per SM) is the default.
  static const int variant = getenv("GPL_GRAM_VARIANT") ? atoi(getenv("GPL_GRAM_VARIANT")) : 4;
  const int smem = (2 * gp.rows * GLD + 256) * (int)sizeof(double);
  if constexpr (MODE >= 0) {
    if (variant == 0) return launch_gram1<MODE, 4, 2>(ctx, gp, a, grid);
    if (variant == 1) return launch_gram1<MODE, 2, 2>(ctx, gp, a, grid);
    if (variant == 2) return launch_gram1<MODE, 4, 1>(ctx, gp, a, grid);
    if (variant == 3 || 4 * smem > 220 * 1024) return launch_gram1<MODE, 2, 3>(ctx, gp, a, grid);
    return launch_gram1<MODE, 2, 4>(ctx, gp, a, grid);
  } else {
    // sums of products keep acc + prod + out per entry (one product term: acc + out): warp tile 32 x 8 to stay
    // at 3-4 CTAs per SM; GPL_GRAM_VARIANT 0 / 1: warp tile 32 x 16 at 2 / 3 CTAs per SM
    if (variant == 0) return launch_gram1<MODE, 2, 2>(ctx, gp, a, grid);
    if (variant == 1 && 3 * smem <= 220 * 1024) return launch_gram1<MODE, 2, 3>(ctx, gp, a, grid);
    if (variant == 3 || 4 * smem > 220 * 1024) return launch_gram1<MODE, 1, 3>(ctx, gp, a, grid);
    return launch_gram1<MODE, 1, 4>(ctx, gp, a, grid);
  }
}

}  // namespace

// Returns GPL_OK with *done = 1 when the tensor-path kernel handled the request, *done = 0 when the
// model is outside its shared-memory budget (the caller falls back to cf_kernels.cu).
int cf_pairwise_gram(Ctx* ctx, const CfPlan& plan, const double* z1, int64_t n1, int64_t ldz1,
                     const double* z2, int64_t n2, int64_t ldz2, bool same, double jitter, double* K,
                     int64_t ldk, int* done) {
  *done = 0;
  static const int off = getenv("GPL_KBUILD_V1") ? atoi(getenv("GPL_KBUILD_V1")) : 0;
  if (off || plan.nfactors < 1) return GPL_OK;
  GPlan gp;
  memset(&gp, 0, sizeof(gp));
  int rows = 0;
  for (int i = 0; i < plan.nfactors; ++i) {
    const FactorDev& f = plan.f[i];
    GFactor& q = gp.f[i];
    q.kind = f.kind;
    q.zoff = f.off;
    q.nf = f.nf;
    q.first = (i == 0) || (plan.f[i - 1].term != f.term);
    q.last = (i == plan.nfactors - 1) || (plan.f[i + 1].term != f.term);
    q.magn2 = f.magn2;
    q.c0 = f.c0;
    q.sc = 1.0;
    q.m3 = 0.0;
    int k = f.nf;
    switch (f.kind) {
      case GPL_CF_CONST: k = 0; break;
      case GPL_CF_LIN: break;
      case GPL_CF_SEXP: q.sc = sqrt(-f.c0); k += 2; break;            // c0 = -1 / (2 l^2): u = r^2 / (2 l^2)
      case GPL_CF_MATERN32: q.sc = f.c0; k += 2; break;               // c0 = sqrt3 / l: u = t^2
      case GPL_CF_MATERN52: q.sc = f.c0; q.m3 = f.magn2 / 3.0; k += 2; break;
      case GPL_CF_NN: q.m3 = f.magn2 * f.c1; break;                   // c1 = 2 / pi
      default: return GPL_OK;
    }
    q.gk = (k + 3) & ~3;
    q.goff = rows;
    rows += q.gk + 1;
  }
  if (rows > MAX_ROWS) return GPL_OK;
  gp.nfactors = plan.nfactors;
  gp.rows = rows;
  GArgs a;
  a.z1 = z1; a.z2 = same ? z1 : z2; a.K = K;
  a.ldz1 = ldz1; a.ldz2 = same ? ldz1 : ldz2; a.ldk = ldk;
  a.n1 = (int)n1; a.n2 = same ? (int)n1 : (int)n2;
  a.same = same ? 1 : 0;
  a.tiles_i = (int)((n1 + TILE - 1) / TILE);
  a.tiles_j = (int)((a.n2 + TILE - 1) / TILE);
  a.jitter = same ? jitter : 0.0;
  const int grid = same ? a.tiles_i * (a.tiles_i + 1) / 2 : a.tiles_i * a.tiles_j;
  *done = 1;
  if (plan.nfactors == 1) {
    switch (plan.f[0].kind) {
      case GPL_CF_SEXP: return launch_gram<GPL_CF_SEXP, 4>(ctx, gp, a, grid);
      case GPL_CF_MATERN32: return launch_gram<GPL_CF_MATERN32, 4>(ctx, gp, a, grid);
      case GPL_CF_MATERN52: return launch_gram<GPL_CF_MATERN52, 4>(ctx, gp, a, grid);
      case GPL_CF_LIN: return launch_gram<GPL_CF_LIN, 4>(ctx, gp, a, grid);
      case GPL_CF_NN: return launch_gram<GPL_CF_NN, 4>(ctx, gp, a, grid);
      default: break;
    }
  }
  // one product term (every factor in the same term): bit-identical values from fewer registers (measured at
  // n = 32768: cf_sexp * cf_lin 2.67 -> 2.39 ms symmetric, 4.25 -> 3.77 ms rectangular).  GPL_GRAM_ONE_TERM=0
  // keeps it on the general kernel.
  static const int one_term_ok = getenv("GPL_GRAM_ONE_TERM") ? atoi(getenv("GPL_GRAM_ONE_TERM")) : 1;
  bool one_term = one_term_ok != 0;
  for (int i = 0; i < plan.nfactors; ++i)
    if (plan.f[i].term != plan.f[0].term) one_term = false;
  if (one_term) return launch_gram<-2, 2>(ctx, gp, a, grid);
  return launch_gram<-1, 2>(ctx, gp, a, grid);
}

}  // namespace gpl
