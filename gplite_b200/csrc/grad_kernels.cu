// Fused gradient trace (north-star item 3; NOT in the reference, whose gp_optim is
// derivative-free, R/gp_optim.R:94):
//
//   out_p = sum_{i,k} W_ik * dK_ik / dlog(theta_p),   W = alpha alpha^T - S Binv S
//
// where S Binv S = (K + V)^-1.  dK/dlog(theta) is regenerated tile by tile from the staged
// feature rows, exactly like the K build, and never materialised; Binv is read once (lower
// triangle, coalesced along columns).  One partial per CTA and a deterministic second-stage sum.
#include "fit.cuh"

#include <math.h>

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
