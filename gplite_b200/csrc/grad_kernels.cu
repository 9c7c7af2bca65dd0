// Fused gradient trace (north-star item 3; NOT in the reference, whose gp_optim is
// derivative-free, R/gp_optim.R:94):
//
//   out_p = sum_{i,k} W_ik * dK_ik / dlog(theta_p),   W = alpha alpha^T - S Binv S
//
// where S Binv S = (K + V)^-1.  dK/dlog(theta) is regenerated tile by tile from the staged
// feature rows, exactly like the K build, and never materialised; Binv is read once (lower
// triangle, coalesced along columns).  One partial per CTA and a deterministic second-stage sum.
#include "fit.cuh"
#include "gram_math.cuh"

#include <math.h>
#include <stdlib.h>

namespace gpl {

namespace {

constexpr int GT = 128;          // tile edge
constexpr int GTHREADS = 256;
constexpr int GP = 8;            // parameters per launch

struct GradArgs {
  const double* z;
  const double* zaux;
  const double* Binv;
  const double* s;
  const double* alpha;
  const double* u;     // optional: W = alpha alpha^T - u u^T - S Binv S (Laplace implicit term), or null
  double* partial;     // [grid][GP]
  int64_t ldz, ldaux, ldb;
  int n;
  int nparam;
  // block mode (multi-GPU: one block column of the inverse per call): Binv holds columns
  // [col_tile0 * GT, ...) only, grid = (row tiles, column tiles of the block)
  int block_mode;
  int col_tile0;
  int col_end;        // one past the last column this call covers
  GradParam prm[GP];
};

// d value / d log(param `slot`) of one factor.  st: s (r2 or dot), sb, sc as in factor_value;
// dsdlp = d s / d log(period) for periodic factors.
__device__ __forceinline__ double factor_dlog(const FactorDev& f, int slot, double v, double s,
                                              double sb, double sc, double dsdlp) {
  switch (f.kind) {
    case GPL_CF_CONST:
    case GPL_CF_LIN:
      if (slot == 0) return 2.0 * v;
      if (slot == 3 && f.kind == GPL_CF_LIN) return f.magn2 * dsdlp;
      return 0.0;
    case GPL_CF_SEXP:
      if (slot == 1) return 2.0 * v;
      if (slot == 0) return v * (-2.0 * f.c0 * s);
      if (slot == 3) return f.c0 * v * dsdlp;
      return 0.0;
    case GPL_CF_MATERN32: {
      if (slot == 1) return 2.0 * v;
      const double t = f.c0 * sqrt(s);
      const double e = f.magn2 * exp(-t);
      if (slot == 0) return e * t * t;
      if (slot == 3) return -0.5 * e * f.c0 * f.c0 * dsdlp;
      return 0.0;
    }
    case GPL_CF_MATERN52: {
      if (slot == 1) return 2.0 * v;
      const double t = f.c0 * sqrt(s);
      const double e = f.magn2 * exp(-t) * (1.0 + t);
      if (slot == 0) return e * (t * t) * (1.0 / 3.0);
      if (slot == 3) return -e * (f.c0 * f.c0) * (1.0 / 6.0) * dsdlp;
      return 0.0;
    }
    case GPL_CF_NN: {
      if (slot == 2) return 2.0 * v;
      const double A = f.c0 + s, B = f.c0 + sb, C = f.c0 + sc;
      const double den = sqrt((1.0 + 2.0 * B) * (1.0 + 2.0 * C));
      const double u = 2.0 * A / den;
      const double dasin = 1.0 / sqrt(1.0 - u * u);
      double dA, dB, dC;
      if (slot == 0) { dA = dB = dC = 2.0 * f.c0; }
      else if (slot == 1) { dA = 2.0 * A; dB = 2.0 * B; dC = 2.0 * C; }
      else return 0.0;
      const double du = 2.0 * dA / den - u * (dB / (1.0 + 2.0 * B) + dC / (1.0 + 2.0 * C));
      return f.magn2 * f.c1 * dasin * du;
    }
  }
  return 0.0;
}

__global__ void __launch_bounds__(GTHREADS, 2)
grad_trace_kernel(const __grid_constant__ CfPlan plan, const GradArgs a) {
  extern __shared__ double sm[];
  const int F = plan.nfeat, FA = plan.naux;
  double* Z1 = sm;
  double* Z2 = Z1 + F * GT;
  double* A1 = Z2 + F * GT;
  double* A2 = A1 + FA * GT;
  double* sv1 = A2 + FA * GT;    // s rows, alpha rows
  double* al1 = sv1 + GT;
  double* sv2 = al1 + GT;
  double* al2 = sv2 + GT;
  double* uu1 = al2 + GT;
  double* uu2 = uu1 + GT;
  __shared__ double red[GTHREADS / 32][GP];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int bi, bj;
  if (a.block_mode) {
    bi = blockIdx.x;
    bj = a.col_tile0 + blockIdx.y;
    if (bi < bj) {                               // above the diagonal: nothing to add
      if (tid < GP) a.partial[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * GP + tid] = 0.0;
      return;
    }
  } else {
    const int L = blockIdx.x;
    bi = (int)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
    while ((int64_t)bi * (bi + 1) / 2 > L) --bi;
    while ((int64_t)(bi + 1) * (bi + 2) / 2 <= L) ++bi;
    bj = L - bi * (bi + 1) / 2;
  }
  const int i0 = bi * GT, j0 = bj * GT;
  const int64_t bcol0 = a.block_mode ? (int64_t)a.col_tile0 * GT : 0;

  for (int idx = tid; idx < F * GT; idx += GTHREADS) {
    const int f = idx >> 7, r = idx & (GT - 1);
    Z1[idx] = (i0 + r < a.n) ? a.z[(int64_t)f * a.ldz + i0 + r] : 0.0;
    Z2[idx] = (j0 + r < a.n) ? a.z[(int64_t)f * a.ldz + j0 + r] : 0.0;
  }
  for (int idx = tid; idx < FA * GT; idx += GTHREADS) {
    const int f = idx >> 7, r = idx & (GT - 1);
    A1[idx] = (i0 + r < a.n) ? a.zaux[(int64_t)f * a.ldaux + i0 + r] : 0.0;
    A2[idx] = (j0 + r < a.n) ? a.zaux[(int64_t)f * a.ldaux + j0 + r] : 0.0;
  }
  if (tid < GT) {
    sv1[tid] = (i0 + tid < a.n) ? a.s[i0 + tid] : 0.0;
    al1[tid] = (i0 + tid < a.n) ? a.alpha[i0 + tid] : 0.0;
    sv2[tid] = (j0 + tid < a.n) ? a.s[j0 + tid] : 0.0;
    al2[tid] = (j0 + tid < a.n) ? a.alpha[j0 + tid] : 0.0;
    uu1[tid] = (a.u && i0 + tid < a.n) ? a.u[i0 + tid] : 0.0;
    uu2[tid] = (a.u && j0 + tid < a.n) ? a.u[j0 + tid] : 0.0;
  }
  __syncthreads();

  double G[GP];
#pragma unroll
  for (int p = 0; p < GP; ++p) G[p] = 0.0;

  // warp w owns columns [16 w, 16 w + 16); lanes own rows lane + 32 r
  for (int c = 0; c < 16; ++c) {
    const int jl = warp * 16 + c;
    const int gj = j0 + jl;
    if (gj >= a.col_end) break;
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
      const int il = lane + 32 * r;
      const int gi = i0 + il;
      if (gi >= a.n || gi < gj) continue;          // lower triangle only
      const double wgt = (gi == gj) ? 1.0 : 2.0;   // mirror entry
      const double W = wgt * (al1[il] * al2[jl] - uu1[il] * uu2[jl] -
                              sv1[il] * a.Binv[((int64_t)gj - bcol0) * a.ldb + gi] * sv2[jl]);
      // walk the terms
      int fi = 0;
      while (fi < plan.nfactors) {
        int fe = fi;
        while (fe < plan.nfactors && plan.f[fe].term == plan.f[fi].term) ++fe;
        // pass 1: values of the factors of this term
        double vals[MAX_FACTORS], st0[MAX_FACTORS], st1[MAX_FACTORS], st2[MAX_FACTORS],
            st3[MAX_FACTORS];
        for (int q = fi; q < fe; ++q) {
          const FactorDev& f = plan.f[q];
          double s = 0.0, sb = 0.0, sc = 0.0, dsdlp = 0.0;
          const bool stationary =
              f.kind == GPL_CF_SEXP || f.kind == GPL_CF_MATERN32 || f.kind == GPL_CF_MATERN52;
          for (int k = f.off; k < f.off + f.nf; ++k) {
            const double av = Z1[k * GT + il], bv = Z2[k * GT + jl];
            if (stationary) {
              const double dlt = av - bv;
              s = fma(dlt, dlt, s);
            } else {
              s = fma(av, bv, s);
              sb = fma(av, av, sb);
              sc = fma(bv, bv, sc);
            }
          }
          if (f.auxoff >= 0) {
            // periodic: features are [sin block | cos block] of the angles in Zaux.
            const int nsel = f.nf / 2;
            for (int k = 0; k < nsel; ++k) {
              const double si = Z1[(f.off + k) * GT + il], ci = Z1[(f.off + nsel + k) * GT + il];
              const double sj = Z2[(f.off + k) * GT + jl], cj = Z2[(f.off + nsel + k) * GT + jl];
              const double ai = A1[(f.auxoff + k) * GT + il], aj = A2[(f.auxoff + k) * GT + jl];
              if (stationary) {
                // r2 = sum 2 - 2 cos(ai - aj); d/dlog p = -2 sin(ai - aj) (ai - aj)
                dsdlp += -2.0 * (si * cj - ci * sj) * (ai - aj);
              } else {
                // dot = sum cos(ai - aj); d/dlog p = sin(ai - aj) (ai - aj)   (cf_lin base)
                dsdlp += (si * cj - ci * sj) * (ai - aj);
              }
            }
          }
          vals[q - fi] = factor_value(f, s, sb, sc);
          st0[q - fi] = s; st1[q - fi] = sb; st2[q - fi] = sc; st3[q - fi] = dsdlp;
        }
#pragma unroll
        for (int p = 0; p < GP; ++p) {
          if (p < a.nparam) {
            const int q = a.prm[p].factor;
            if (q >= fi && q < fe) {
              double dv = factor_dlog(plan.f[q], a.prm[p].slot, vals[q - fi], st0[q - fi],
                                      st1[q - fi], st2[q - fi], st3[q - fi]);
              for (int o = fi; o < fe; ++o)
                if (o != q) dv *= vals[o - fi];
              G[p] = fma(W, dv, G[p]);
            }
          }
        }
        fi = fe;
      }
    }
  }

#pragma unroll
  for (int p = 0; p < GP; ++p) {
    double v = G[p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][p] = v;
  }
  __syncthreads();
  if (tid < GP) {
    double v = 0.0;
    for (int w = 0; w < GTHREADS / 32; ++w) v += red[w][tid];
    a.partial[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * GP + tid] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Fast path: ONE stationary factor (cf_sexp / cf_matern32 / cf_matern52 -- BASELINE configs 2 and 5),
// parameters among {lscale, magn}.  Same machinery as the K build (cf_gram.cu): the scaled squared
// distance of a 128 x 128 tile comes out of one DMMA.8x8x4 chain over the staged, shifted, augmented
// feature rows; exp / sqrt are the branch-free replacements; both derivative sums stay in registers.
// Per lower-triangle entry: ~25 fp64 instructions and one 8-byte read of Binv, against ~200
// instructions (libm exp / sqrt, per-entry arrays) in the generic kernel below -- 15.1 ms -> the HBM
// roofline of the Binv read is 0.66 ms at n = 32768.
constexpr int FLD = GT + 4;

struct FastArgs {
  const double* z;
  const double* Binv;
  const double* s;
  const double* alpha;
  const double* u;
  double* partial;     // [grid][GP]: slot 0 = sum W dK/dlog(lscale), slot 1 = sum W 2 K
  int64_t ldz, ldb;
  int n, nf, gk;       // gk = pad4(nf + 2)
  int block_mode, col_tile0, col_end;
  double sc, magn2, m3;
};

template <int KIND, int MINB, int NS>
__global__ void __launch_bounds__(GTHREADS, MINB)
grad_trace_fast_kernel(const FastArgs a) {
  extern __shared__ double sm[];
  double* Zs1 = sm;                    // [gk][FLD]
  double* Zs2 = sm + a.gk * FLD;
  double* tab = sm + 2 * a.gk * FLD;   // [256]
  double* vec = tab + 256;             // s, alpha, u for rows then columns: [6][GT]
  __shared__ double red[GTHREADS / 32][2];
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  tab[tid] = kExpTab256[tid];
  if (tid == 0) s_bad = 0;
  int bi, bj;
  int64_t pslot;
  if (a.block_mode) {
    bi = blockIdx.x;
    bj = a.col_tile0 + blockIdx.y;
    pslot = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    if (bi < bj) {                               // above the diagonal: nothing to add
      if (tid < GP) a.partial[pslot * GP + tid] = 0.0;
      return;
    }
  } else {
    const int L = blockIdx.x;
    bi = (int)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
    while ((int64_t)bi * (bi + 1) / 2 > L) --bi;
    while ((int64_t)(bi + 1) * (bi + 2) / 2 <= L) ++bi;
    bj = L - bi * (bi + 1) / 2;
    pslot = L;
  }
  const int i0 = bi * GT, j0 = bj * GT;
  const int64_t bcol0 = a.block_mode ? (int64_t)a.col_tile0 * GT : 0;

  for (int idx = tid; idx < (a.gk - 2) * GT; idx += GTHREADS) {
    const int j = idx >> 7, r = idx & (GT - 1);
    double v1 = 0.0, v2 = 0.0;
    if (j < a.nf) {
      const double ref = a.z[(int64_t)j * a.ldz];
      if (i0 + r < a.n) v1 = (a.z[(int64_t)j * a.ldz + i0 + r] - ref) * a.sc;
      if (j0 + r < a.n) v2 = (a.z[(int64_t)j * a.ldz + j0 + r] - ref) * a.sc * -2.0;
    }
    Zs1[(j + 2) * FLD + r] = v1;          // rows 0, 1: the two norm columns (as in cf_gram.cu)
    Zs2[(j + 2) * FLD + r] = v2;
  }
  if (tid < GT) {
    const bool ok = i0 + tid < a.n;
    vec[0 * GT + tid] = ok ? a.s[i0 + tid] : 0.0;
    vec[1 * GT + tid] = ok ? a.alpha[i0 + tid] : 0.0;
    vec[2 * GT + tid] = (ok && a.u) ? a.u[i0 + tid] : 0.0;
  } else {
    const int r = tid - GT;
    const bool ok = j0 + r < a.n;
    vec[3 * GT + r] = ok ? a.s[j0 + r] : 0.0;
    vec[4 * GT + r] = ok ? a.alpha[j0 + r] : 0.0;
    vec[5 * GT + r] = (ok && a.u) ? a.u[j0 + r] : 0.0;
  }
  __syncthreads();
  {
    double* Zs = (tid < GT) ? Zs1 : Zs2;
    const int r = tid & (GT - 1);
    double sacc = 0.0;
    for (int j = 2; j < a.nf + 2; ++j) {
      const double v = Zs[j * FLD + r];
      sacc = fma(v, v, sacc);
    }
    if (tid >= GT) sacc *= 0.25;
    Zs[r] = (tid < GT) ? sacc : 1.0;
    Zs[FLD + r] = (tid < GT) ? 1.0 : sacc;
    if (!(sacc < 250.0)) s_bad = 1;
  }
  __syncthreads();
  const bool bad = s_bad != 0;
  const bool diag_tile = bi == bj;

  double G0 = 0.0, G1 = 0.0;
#pragma unroll 1
  for (int pass = 0; pass < GT / (16 * NS); ++pass) {   // warp tile 32 x (8 NS) per pass
    const int cbase = pass * (16 * NS) + wn * (8 * NS);
    double acc[4][NS][2];
#pragma unroll
    for (int im = 0; im < 4; ++im)
#pragma unroll
      for (int in = 0; in < NS; ++in) acc[im][in][0] = acc[im][in][1] = 0.0;
    if (!bad) {
      for (int ks = 0; ks < a.gk; ks += 4) {
        const int kk = ks + t;
        double av[4], bv[NS];
#pragma unroll
        for (int im = 0; im < 4; ++im) av[im] = Zs1[kk * FLD + wm * 32 + im * 8 + g];
#pragma unroll
        for (int in = 0; in < NS; ++in) bv[in] = Zs2[kk * FLD + cbase + in * 8 + g];
#pragma unroll
        for (int im = 0; im < 4; ++im)
#pragma unroll
          for (int in = 0; in < NS; ++in) dmma_k4(acc[im][in][0], acc[im][in][1], av[im], bv[in]);
      }
    }
#pragma unroll
    for (int im = 0; im < 4; ++im) {
      const int row = wm * 32 + im * 8 + g;
      const int gi = i0 + row;
      const double si = vec[0 * GT + row], ai = vec[1 * GT + row], ui = vec[2 * GT + row];
#pragma unroll
      for (int in = 0; in < NS; ++in) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = cbase + in * 8 + 2 * t + c;
          const int gj = j0 + col;
          // lower triangle of the matrix, inside the covered columns
          const bool live = gi < a.n && gj < a.col_end && gi >= gj;
          double uarg;
          if (!bad) {
            uarg = acc[im][in][c];
          } else {
            uarg = 0.0;
            for (int j = 2; j < a.nf + 2; ++j) {
              const double dlt = fma(0.5, Zs2[j * FLD + col], Zs1[j * FLD + row]);
              uarg = fma(dlt, dlt, uarg);
            }
          }
          if (diag_tile && row == col) uarg = 0.0;
          double kv, dl;                            // K and d K / d log(lscale)
          if (KIND == GPL_CF_SEXP) {
            uarg = clamp_arg(uarg);
            kv = a.magn2 * exp_m(uarg, tab);
            dl = kv * (2.0 * uarg);
          } else {
            const double tt = clamp_hi(sqrt_pos(clamp_lo(uarg)));   // only exp's argument is capped
            const double t2 = tt * tt;
            const double e = a.magn2 * exp_m(tt, tab);
            if (KIND == GPL_CF_MATERN32) {
              kv = fma(e, tt, e);
              dl = e * t2;
            } else {
              const double e1 = fma(e, tt, e);       // magn^2 (1 + t) e^-t
              kv = fma(e * (1.0 / 3.0), t2, e1);
              dl = e1 * (t2 * (1.0 / 3.0));
            }
          }
          const double bij = live ? a.Binv[((int64_t)gj - bcol0) * a.ldb + gi] : 0.0;
          const double sk = vec[3 * GT + col], ak = vec[4 * GT + col], uk = vec[5 * GT + col];
          double W = fma(ai, ak, -(si * bij) * sk);
          W = fma(-ui, uk, W);
          W = live ? ((gi == gj) ? W : 2.0 * W) : 0.0;
          G0 = fma(W, dl, G0);
          G1 = fma(W, kv, G1);
        }
      }
    }
  }
  G1 *= 2.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    G0 += __shfl_xor_sync(0xffffffffu, G0, o);
    G1 += __shfl_xor_sync(0xffffffffu, G1, o);
  }
  if (lane == 0) {
    red[warp][0] = G0;
    red[warp][1] = G1;
  }
  __syncthreads();
  if (tid < GP) {
    double v = 0.0;
    if (tid < 2)
      for (int w = 0; w < GTHREADS / 32; ++w) v += red[w][tid];
    a.partial[pslot * GP + tid] = v;
  }
}

template <int KIND, int MINB>
int launch_fast1(Ctx* ctx, const FastArgs& a, dim3 grid, int smem) {
  constexpr int NS = MINB >= 3 ? 1 : 2;
  GPL_CUDA(cudaFuncSetAttribute(grad_trace_fast_kernel<KIND, MINB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem));
  grad_trace_fast_kernel<KIND, MINB, NS><<<grid, GTHREADS, smem, ctx->stream>>>(a);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

// like the K build this kernel is latency bound: occupancy first (GPL_GRAD_MINB = 2 / 3 / 4 CTAs per SM)
template <int KIND>
int launch_fast(Ctx* ctx, const FastArgs& a, dim3 grid) {
  static const int minb = getenv("GPL_GRAD_MINB") ? atoi(getenv("GPL_GRAD_MINB")) : 3;
  const int smem = (2 * a.gk * FLD + 256 + 6 * GT) * (int)sizeof(double);
  if (minb >= 4 && 4 * smem <= 220 * 1024) return launch_fast1<KIND, 4>(ctx, a, grid, smem);
  if (minb >= 3 && 3 * smem <= 220 * 1024) return launch_fast1<KIND, 3>(ctx, a, grid, smem);
  return launch_fast1<KIND, 2>(ctx, a, grid, smem);
}

__global__ void sum_grad_partials_kernel(const double* __restrict__ partial, int grid, int nparam,
                                         double* __restrict__ out) {
  __shared__ double red[32];
  const int p = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < grid; i += blockDim.x) acc += partial[(int64_t)i * GP + p];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0 && p < nparam) out[p] = acc;
  }
}

}  // namespace

int grad_trace(Ctx* ctx, const CfPlan& plan, const double* z, int64_t ldz, const double* zaux,
               int64_t n, const double* Binv, int64_t ldb, const double* s, const double* alpha,
               int nparam, const GradParam* params, double* d_out, int64_t col0, int64_t ncols,
               const double* u) {
  if (nparam <= 0 || n <= 0) return GPL_OK;
  const int tiles = (int)((n + GT - 1) / GT);
  const bool block_mode = col0 >= 0;
  if (block_mode && (col0 % GT != 0 || ncols <= 0 || col0 + ncols > n)) {
    set_error("grad_trace: the block column must start at a multiple of %d inside the matrix", GT);
    return GPL_ERR_BAD_ARG;
  }
  const int ctiles = block_mode ? (int)((ncols + GT - 1) / GT) : 0;
  const int grid = block_mode ? tiles * ctiles : tiles * (tiles + 1) / 2;
  const int smem = (2 * (plan.nfeat + plan.naux) * GT + 6 * GT) * (int)sizeof(double);
  if (smem > 227 * 1024) {
    set_error("grad_trace: too many feature columns (%d) for shared memory", plan.nfeat);
    return GPL_ERR_BAD_ARG;
  }
  GPL_CUDA(cudaFuncSetAttribute(grad_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem));
  DevBuf part;
  GPL_TRY(part.alloc(ctx, sizeof(double) * (size_t)grid * GP));
  // single stationary factor, parameters among {lscale, magn}: the tensor-path kernel
  static const int no_fast = getenv("GPL_GRAD_GENERIC") ? atoi(getenv("GPL_GRAD_GENERIC")) : 0;
  bool fast_ok = !no_fast && plan.nfactors == 1 && plan.f[0].auxoff < 0 && nparam <= 2 && plan.f[0].nf >= 1 &&
                 plan.f[0].nf <= 38 &&
                 (plan.f[0].kind == GPL_CF_SEXP || plan.f[0].kind == GPL_CF_MATERN32 ||
                  plan.f[0].kind == GPL_CF_MATERN52);
  for (int p = 0; p < nparam && fast_ok; ++p)
    if (params[p].factor != 0 || params[p].slot < 0 || params[p].slot > 1) fast_ok = false;
  if (fast_ok && nparam == 2 && params[0].slot == params[1].slot) fast_ok = false;
  if (fast_ok) {
    const FactorDev& f = plan.f[0];
    FastArgs a;
    a.z = z + (int64_t)f.off * ldz; a.Binv = Binv; a.s = s; a.alpha = alpha; a.u = u; a.partial = part.d();
    a.ldz = ldz; a.ldb = ldb;
    a.n = (int)n; a.nf = f.nf; a.gk = (f.nf + 2 + 3) & ~3;
    a.block_mode = block_mode ? 1 : 0;
    a.col_tile0 = block_mode ? (int)(col0 / GT) : 0;
    a.col_end = block_mode ? (int)(col0 + ncols) : (int)n;
    a.sc = f.kind == GPL_CF_SEXP ? sqrt(-f.c0) : f.c0;
    a.magn2 = f.magn2;
    a.m3 = f.magn2 / 3.0;
    const dim3 g = block_mode ? dim3((unsigned)tiles, (unsigned)ctiles) : dim3((unsigned)grid);
    if (f.kind == GPL_CF_SEXP) GPL_TRY(launch_fast<GPL_CF_SEXP>(ctx, a, g));
    else if (f.kind == GPL_CF_MATERN32) GPL_TRY(launch_fast<GPL_CF_MATERN32>(ctx, a, g));
    else GPL_TRY(launch_fast<GPL_CF_MATERN52>(ctx, a, g));
    // partial slot 0 holds the lscale sum, slot 1 the magn sum: route them to the callers' order
    for (int p = 0; p < nparam; ++p) {
      sum_grad_partials_kernel<<<1, 256, 0, ctx->stream>>>(part.d() + params[p].slot, grid, 1, d_out + p);
      GPL_LAUNCH_CHECK(ctx);
    }
    return GPL_OK;
  }
  for (int p0 = 0; p0 < nparam; p0 += GP) {
    GradArgs a;
    a.z = z; a.zaux = zaux; a.Binv = Binv; a.s = s; a.alpha = alpha; a.u = u; a.partial = part.d();
    a.ldz = ldz; a.ldaux = ldz; a.ldb = ldb;
    // block mode: columns beyond the block belong to other calls
    a.n = (int)n;
    a.block_mode = block_mode ? 1 : 0;
    a.col_tile0 = block_mode ? (int)(col0 / GT) : 0;
    a.col_end = block_mode ? (int)(col0 + ncols) : (int)n;
    a.nparam = nparam - p0 < GP ? nparam - p0 : GP;
    for (int p = 0; p < GP; ++p) {
      a.prm[p].factor = -2;
      a.prm[p].slot = 0;
      if (p < a.nparam) a.prm[p] = params[p0 + p];
    }
    const dim3 g = block_mode ? dim3((unsigned)tiles, (unsigned)ctiles) : dim3((unsigned)grid);
    grad_trace_kernel<<<g, GTHREADS, smem, ctx->stream>>>(plan, a);
    GPL_LAUNCH_CHECK(ctx);
    sum_grad_partials_kernel<<<a.nparam, 256, 0, ctx->stream>>>(part.d(), grid, a.nparam,
                                                                d_out + p0);
    GPL_LAUNCH_CHECK(ctx);
  }
  return GPL_OK;
}

}  // namespace gpl
