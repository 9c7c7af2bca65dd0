// Covariance-matrix construction: eval_cf.* (R/cf.R:402-497) and the native loops of
// src/cf.cpp, as tiled sm_100a kernels.
//
// A flattened model (sum over terms of products of factors) is evaluated in ONE pass: raw
// inputs are first mapped to a small feature matrix Z (column subset, normalisation,
// cf_periodic's sin/cos map, cf_nn's sigma scaling -- O(n d) work), then each CTA stages
// 128 rows of Z1 and 128 rows of Z2 in shared memory and produces a 128 x 128 tile of K.
// Each thread owns rows {lane, lane+32, lane+64, lane+96} x CJ columns, so Z1 reads are
// conflict-free, Z2 reads are warp broadcasts, and every global store instruction writes
// 256 contiguous bytes of one column of K (column-major, as R expects).  When x1 is x2 only
// tiles on or below the diagonal are computed; the mirror tile is transposed through shared
// memory and written with the same coalescing.  The diagonal jitter (R/gp_fit.R:95) is fused.
// No per-cf n x n temporaries exist (the reference materialises one per cf plus one per
// `*`/`+`, R/cf.R:408,494).
#include "cf.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace gpl {

namespace {

constexpr int TILE = 128;
constexpr int PAIR_THREADS = 256;
constexpr int TLD = TILE + 1;

struct PairArgs {
  const double* z1;
  const double* z2;
  double* K;
  int64_t ldz1, ldz2, ldk;
  int n1, n2;
  int same;
  int tiles_i, tiles_j;
  double jitter;
};

template <int CJ>
struct Acc {
  double s[4][CJ];
};

// KIND >= 0: single-factor fast path of that kind.  KIND == -1: generic sum of products.
template <int CJ, int KIND>
__global__ void __launch_bounds__(PAIR_THREADS, 2)
pair_kernel(const __grid_constant__ CfPlan plan, const PairArgs a) {
  extern __shared__ double sm[];
  const int F = plan.nfeat;
  double* Zs1 = sm;                 // [F][128]
  double* Zs2 = sm + F * TILE;      // [F][128]
  double* T = sm + 2 * F * TILE;    // [8*CJ][129], only used on the mirrored path
  constexpr int PASSCOLS = 8 * CJ;
  __shared__ double etab[64];       // 2^(j/64) for exp_neg_fast
  if (threadIdx.x < 64) etab[threadIdx.x] = kExpTab[threadIdx.x];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

  for (int idx = tid; idx < F * TILE; idx += PAIR_THREADS) {
    const int f = idx >> 7, r = idx & (TILE - 1);
    const int gi = i0 + r, gj = j0 + r;
    Zs1[idx] = (gi < a.n1) ? a.z1[(int64_t)f * a.ldz1 + gi] : 0.0;
    Zs2[idx] = (gj < a.n2) ? a.z2[(int64_t)f * a.ldz2 + gj] : 0.0;
  }
  __syncthreads();

#pragma unroll 1
  for (int pass = 0; pass < TILE / PASSCOLS; ++pass) {
    const int jl0 = pass * PASSCOLS + warp * CJ;     // first local column of this thread
    double out[4][CJ];

    if (KIND >= 0) {
      const FactorDev& f = plan.f[0];
      double s[4][CJ], sb[4], sc[CJ];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        sb[r] = 0.0;
#pragma unroll
        for (int c = 0; c < CJ; ++c) s[r][c] = 0.0;
      }
#pragma unroll
      for (int c = 0; c < CJ; ++c) sc[c] = 0.0;
      for (int k = f.off; k < f.off + f.nf; ++k) {
        double av[4], bv[CJ];
#pragma unroll
        for (int r = 0; r < 4; ++r) av[r] = Zs1[k * TILE + lane + 32 * r];
#pragma unroll
        for (int c = 0; c < CJ; ++c) bv[c] = Zs2[k * TILE + jl0 + c];
        if (KIND == GPL_CF_SEXP || KIND == GPL_CF_MATERN32 || KIND == GPL_CF_MATERN52) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < CJ; ++c) {
              const double dlt = av[r] - bv[c];
              s[r][c] = fma(dlt, dlt, s[r][c]);
            }
        } else {
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < CJ; ++c) s[r][c] = fma(av[r], bv[c], s[r][c]);
          if (KIND == GPL_CF_NN) {
#pragma unroll
            for (int r = 0; r < 4; ++r) sb[r] = fma(av[r], av[r], sb[r]);
#pragma unroll
            for (int c = 0; c < CJ; ++c) sc[c] = fma(bv[c], bv[c], sc[c]);
          }
        }
      }
      FactorDev fk = f;
      fk.kind = KIND;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < CJ; ++c) out[r][c] = factor_value(fk, s[r][c], sb[r], sc[c], etab);
    } else {
      double prod[4][CJ];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < CJ; ++c) out[r][c] = 0.0;
      for (int fi = 0; fi < plan.nfactors; ++fi) {
        const FactorDev& f = plan.f[fi];
        double s[4][CJ], sb[4], sc[CJ];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          sb[r] = 0.0;
#pragma unroll
          for (int c = 0; c < CJ; ++c) s[r][c] = 0.0;
        }
#pragma unroll
        for (int c = 0; c < CJ; ++c) sc[c] = 0.0;
        const bool stationary =
            f.kind == GPL_CF_SEXP || f.kind == GPL_CF_MATERN32 || f.kind == GPL_CF_MATERN52;
        for (int k = f.off; k < f.off + f.nf; ++k) {
          double av[4], bv[CJ];
#pragma unroll
          for (int r = 0; r < 4; ++r) av[r] = Zs1[k * TILE + lane + 32 * r];
#pragma unroll
          for (int c = 0; c < CJ; ++c) bv[c] = Zs2[k * TILE + jl0 + c];
          if (stationary) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int c = 0; c < CJ; ++c) {
                const double dlt = av[r] - bv[c];
                s[r][c] = fma(dlt, dlt, s[r][c]);
              }
          } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              sb[r] = fma(av[r], av[r], sb[r]);
#pragma unroll
              for (int c = 0; c < CJ; ++c) s[r][c] = fma(av[r], bv[c], s[r][c]);
            }
#pragma unroll
            for (int c = 0; c < CJ; ++c) sc[c] = fma(bv[c], bv[c], sc[c]);
          }
        }
        const bool first = (fi == 0) || (plan.f[fi - 1].term != f.term);
        const bool last = (fi == plan.nfactors - 1) || (plan.f[fi + 1].term != f.term);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < CJ; ++c) {
            const double v = factor_value(f, s[r][c], sb[r], sc[c], etab);
            prod[r][c] = first ? v : prod[r][c] * v;
            if (last) out[r][c] += prod[r][c];
          }
      }
    }

    // direct store: lanes -> consecutive rows of one column (256 contiguous bytes per warp store)
    if (interior && !diag_tile) {
      double* Kp = a.K + (int64_t)(j0 + jl0) * a.ldk + i0 + lane;
#pragma unroll
      for (int c = 0; c < CJ; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) Kp[(int64_t)c * a.ldk + 32 * r] = out[r][c];
    } else {
#pragma unroll
      for (int c = 0; c < CJ; ++c) {
        const int gj = j0 + jl0 + c;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int gi = i0 + lane + 32 * r;
          double v = out[r][c];
          if (a.same && gi == gj) v += a.jitter;
          out[r][c] = v;
          if (gi < a.n1 && gj < a.n2) a.K[(int64_t)gj * a.ldk + gi] = v;
        }
      }
    }
    if (mirror) {
      // transpose this 128 x PASSCOLS slab through shared memory, then lanes -> consecutive
      // rows (original columns) of the mirrored tile
#pragma unroll
      for (int c = 0; c < CJ; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) T[(warp * CJ + c) * TLD + lane + 32 * r] = out[r][c];
      __syncthreads();
      if (interior) {
        double* Kp = a.K + (int64_t)i0 * a.ldk + j0 + pass * PASSCOLS + lane;
#pragma unroll 4
        for (int il = warp; il < TILE; il += PAIR_THREADS / 32) {
#pragma unroll
          for (int h = 0; h < PASSCOLS / 32; ++h)
            Kp[(int64_t)il * a.ldk + 32 * h] = T[(lane + 32 * h) * TLD + il];
        }
      } else {
        for (int il = warp; il < TILE; il += PAIR_THREADS / 32) {
          const int gcol = i0 + il;            // column of the mirrored tile
          if (gcol >= a.n1) break;
#pragma unroll
          for (int h = 0; h < PASSCOLS / 32; ++h) {
            const int cl = lane + 32 * h;
            const int grow = j0 + pass * PASSCOLS + cl;
            if (grow < a.n2) a.K[(int64_t)gcol * a.ldk + grow] = T[cl * TLD + il];
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stationary single-factor fast path: squared distances through the FP64 tensor pipe.
//   r2_ij = |a_i|^2 + |b_j|^2 - 2 a_i . b_j,  a . b by DMMA.8x8x4 (a Gram contraction over features)
// which halves the fp64-pipe time of the distance part (one FMA per feature instead of a subtract
// and an FMA) and removes it from the issue stream (2 DMMA per 8 x 8 entries at d = 8).  All rows
// are first shifted by one common reference row (distances are translation invariant) and a tile
// whose |a|^2 + |b|^2 is still so large that the cancellation could cost more than ~5e-13 relative
// in K falls back to direct differences (CTA-uniform branch).  Warp tile 32 x 32 per 64-column pass;
// the accumulator layout (row g, columns 2t, 2t+1) makes the mirrored store a 16-byte vector store
// and the direct stores 64-byte row segments, so no shared-memory transpose is needed.
constexpr int GLD = TILE + 4;       // smem row stride, == 4 mod 16: conflict-free fragment loads

__device__ __forceinline__ void dmma_k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int KIND>
__global__ void __launch_bounds__(PAIR_THREADS, 2)
pair_gram_kernel(const __grid_constant__ CfPlan plan, const PairArgs a) {
  extern __shared__ double sm[];
  const int F = plan.nfeat, F4 = (F + 3) & ~3;
  double* Zs1 = sm;                      // [F4][GLD]
  double* Zs2 = sm + F4 * GLD;
  double* nrm1 = sm + 2 * F4 * GLD;      // [TILE]
  double* nrm2 = nrm1 + TILE;
  __shared__ double etab[64];
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  if (tid < 64) etab[tid] = kExpTab[tid];
  if (tid == 0) s_bad = 0;

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

  for (int idx = tid; idx < F4 * TILE; idx += PAIR_THREADS) {
    const int f = idx >> 7, r = idx & (TILE - 1);
    const int gi = i0 + r, gj = j0 + r;
    double v1 = 0.0, v2 = 0.0;
    if (f < F) {
      const double ref = a.z1[(int64_t)f * a.ldz1];
      if (gi < a.n1) v1 = a.z1[(int64_t)f * a.ldz1 + gi] - ref;
      if (gj < a.n2) v2 = a.z2[(int64_t)f * a.ldz2 + gj] - ref;
    }
    Zs1[f * GLD + r] = v1;
    Zs2[f * GLD + r] = v2;
  }
  __syncthreads();
  FactorDev fk = plan.f[0];
  fk.kind = KIND;
  {
    const double* Zs = (tid < TILE) ? Zs1 : Zs2;
    const int r = tid & (TILE - 1);
    double sacc = 0.0;
    for (int f = 0; f < F4; ++f) {
      const double v = Zs[f * GLD + r];
      sacc = fma(v, v, sacc);
    }
    ((tid < TILE) ? nrm1 : nrm2)[r] = sacc;
    // cancellation guard: relative error of K ~ cs * 4 eps (|a|^2 + |b|^2), cs ~ 1 / lscale^2
    const double cs = (KIND == GPL_CF_SEXP) ? fabs(fk.c0) : fk.c0 * fk.c0;
    if (!(2.0 * sacc * cs < 500.0)) s_bad = 1;       // also catches NaN
  }
  __syncthreads();
  const bool bad = s_bad != 0;

#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    double acc[4][4][2];
#pragma unroll
    for (int im = 0; im < 4; ++im)
#pragma unroll
      for (int in = 0; in < 4; ++in) acc[im][in][0] = acc[im][in][1] = 0.0;
    const int cbase = pass * 64 + wn * 32;
    if (!bad) {
      for (int ks = 0; ks < F4; ks += 4) {
        const int kk = ks + t;
        double av[4], bv[4];
#pragma unroll
        for (int im = 0; im < 4; ++im) av[im] = Zs1[kk * GLD + wm * 32 + im * 8 + g];
#pragma unroll
        for (int in = 0; in < 4; ++in) bv[in] = Zs2[kk * GLD + cbase + in * 8 + g];
#pragma unroll
        for (int im = 0; im < 4; ++im)
#pragma unroll
          for (int in = 0; in < 4; ++in) dmma_k4(acc[im][in][0], acc[im][in][1], av[im], bv[in]);
      }
    }
#pragma unroll
    for (int im = 0; im < 4; ++im) {
      const int row = wm * 32 + im * 8 + g;
      const int gi = i0 + row;
      const double n1v = nrm1[row];
#pragma unroll
      for (int in = 0; in < 4; ++in) {
        const int col = cbase + in * 8 + 2 * t;
        const int gj = j0 + col;
        double v[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double r2;
          if (!bad) {
            r2 = fma(-2.0, acc[im][in][c], n1v + nrm2[col + c]);
            r2 = 0.5 * (r2 + fabs(r2));                       // clamp at 0, NaN stays NaN
          } else {
            r2 = 0.0;
            for (int f = 0; f < F; ++f) {
              const double dlt = Zs1[f * GLD + row] - Zs2[f * GLD + col + c];
              r2 = fma(dlt, dlt, r2);
            }
          }
          if (diag_tile && gi == gj + c) r2 = 0.0;
          double val = factor_value(fk, r2, 0.0, 0.0, etab);
          if (diag_tile && gi == gj + c) val += a.jitter;
          v[c] = val;
        }
        if (interior) {
          a.K[(int64_t)gj * a.ldk + gi] = v[0];
          a.K[(int64_t)(gj + 1) * a.ldk + gi] = v[1];
          if (mirror) {
            double* dst = a.K + (int64_t)gi * a.ldk + gj;
            if (vec_ok) {
              *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
            } else {
              dst[0] = v[0];
              dst[1] = v[1];
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (gi < a.n1 && gj + c < a.n2) {
              a.K[(int64_t)(gj + c) * a.ldk + gi] = v[c];
              if (mirror) a.K[(int64_t)gi * a.ldk + gj + c] = v[c];
            }
          }
        }
      }
    }
  }
}

template <int KIND>
int launch_pair_gram(Ctx* ctx, const CfPlan& plan, const PairArgs& a, int grid) {
  const int F4 = (plan.nfeat + 3) & ~3;
  const int smem = (2 * F4 * GLD + 2 * TILE) * (int)sizeof(double);
  GPL_CUDA(cudaFuncSetAttribute(pair_gram_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem));
  pair_gram_kernel<KIND><<<grid, PAIR_THREADS, smem, ctx->stream>>>(plan, a);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

struct DiagArgs {
  const double* z1;
  const double* z2;
  int64_t ldz1, ldz2;
  int64_t n;
  double add;
  double* out;
};

__global__ void diag_kernel(const __grid_constant__ CfPlan plan, const DiagArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  double total = 0.0, prod = 0.0;
  for (int fi = 0; fi < plan.nfactors; ++fi) {
    const FactorDev& f = plan.f[fi];
    double s = 0.0, sb = 0.0, sc = 0.0;
    const bool stationary =
        f.kind == GPL_CF_SEXP || f.kind == GPL_CF_MATERN32 || f.kind == GPL_CF_MATERN52;
    for (int k = f.off; k < f.off + f.nf; ++k) {
      const double av = a.z1[(int64_t)k * a.ldz1 + i], bv = a.z2[(int64_t)k * a.ldz2 + i];
      if (stationary) {
        const double dlt = av - bv;
        s = fma(dlt, dlt, s);
      } else {
        s = fma(av, bv, s);
        sb = fma(av, av, sb);
        sc = fma(bv, bv, sc);
      }
    }
    const double v = factor_value(f, s, sb, sc);
    const bool first = (fi == 0) || (plan.f[fi - 1].term != f.term);
    const bool last = (fi == plan.nfactors - 1) || (plan.f[fi + 1].term != f.term);
    prod = first ? v : prod * v;
    if (last) total += prod;
  }
  a.out[i] = total + a.add;
}

struct PrepColDev {
  int src, dst, dst2, flags;   // flags bit0 normalize, bit1 periodic
  int dst_aux, pad;
  double mean, scale, w, mul;
};

__global__ void prep_kernel(const double* __restrict__ x, int64_t n, int64_t ldx,
                            double* __restrict__ z, int64_t ldz, double* __restrict__ zaux,
                            int64_t ldzaux, const PrepColDev* __restrict__ cols, int ncols) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int c = 0; c < ncols; ++c) {
    const PrepColDev pc = cols[c];
    double v = x[(int64_t)pc.src * ldx + i];
    if (pc.flags & 1) v = (v - pc.mean) / pc.scale;
    if (pc.flags & 2) {
      const double ang = pc.w * v;
      if (zaux && pc.dst_aux >= 0) zaux[(int64_t)pc.dst_aux * ldzaux + i] = ang;
      z[(int64_t)pc.dst * ldz + i] = sin(ang) * pc.mul;
      z[(int64_t)pc.dst2 * ldz + i] = cos(ang) * pc.mul;
    } else {
      z[(int64_t)pc.dst * ldz + i] = v * pc.mul;
    }
  }
}

template <int CJ, int KIND>
int launch_pair(Ctx* ctx, const CfPlan& plan, const PairArgs& a, int grid) {
  const int smem = (2 * plan.nfeat * TILE + 8 * CJ * TLD) * (int)sizeof(double);
  if (smem > 227 * 1024 / 2) {
    // more than half an SM's shared memory: still legal up to 227 KB, one CTA per SM
    if (smem > 227 * 1024) {
      set_error("cf_pairwise: %d feature columns need %d bytes of shared memory (max 232448)",
                plan.nfeat, smem);
      return GPL_ERR_BAD_ARG;
    }
  }
  GPL_CUDA(cudaFuncSetAttribute(pair_kernel<CJ, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem));
  pair_kernel<CJ, KIND><<<grid, PAIR_THREADS, smem, ctx->stream>>>(plan, a);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

}  // namespace

int make_plan(const gpl_cf_model* cf, int d, uint64_t term_mask, HostPlan* out, bool want_aux) {
  if (!cf || cf->nfactors < 0 || (cf->nfactors > 0 && !cf->factors)) {
    set_error("cf model: null descriptor");
    return GPL_ERR_BAD_ARG;
  }
  CfPlan& p = out->plan;
  memset(&p, 0, sizeof(p));
  out->cols.clear();
  p.term_mask = term_mask;
  int nfeat = 0, nf_out = 0, naux = 0;
  for (int i = 0; i < cf->nfactors; ++i) {
    const gpl_cf_factor& f = cf->factors[i];
    if (term_mask && f.term < 64 && !((term_mask >> f.term) & 1ull)) continue;
    if (nf_out >= MAX_FACTORS) {
      set_error("cf model: more than %d factors", MAX_FACTORS);
      return GPL_ERR_BAD_ARG;
    }
    if (f.kind < GPL_CF_CONST || f.kind > GPL_CF_NN) {
      set_error("cf model: unknown kind %d", f.kind);
      return GPL_ERR_BAD_ARG;
    }
    if (i > 0 && f.term < cf->factors[i - 1].term) {
      set_error("cf model: factors must be ordered by term");
      return GPL_ERR_BAD_ARG;
    }
    FactorDev& fd = p.f[nf_out++];
    fd.kind = f.kind;
    fd.term = f.term;
    fd.off = nfeat;
    fd.auxoff = -1;
    const int nsel = f.nvars > 0 ? f.nvars : d;
    double mul = 1.0;
    switch (f.kind) {
      case GPL_CF_CONST:
      case GPL_CF_LIN:
        fd.magn2 = f.p[0] * f.p[0];
        break;
      case GPL_CF_SEXP:
        fd.c0 = -0.5 / (f.p[0] * f.p[0]);
        fd.magn2 = f.p[1] * f.p[1];
        break;
      case GPL_CF_MATERN32:
        fd.c0 = sqrt(3.0) / f.p[0];
        fd.magn2 = f.p[1] * f.p[1];
        break;
      case GPL_CF_MATERN52:
        fd.c0 = sqrt(5.0) / f.p[0];
        fd.c1 = 5.0 / (3.0 * (f.p[0] * f.p[0]));
        fd.magn2 = f.p[1] * f.p[1];
        fd.c2 = fd.magn2 * fd.c1;
        break;
      case GPL_CF_NN:
        fd.c0 = f.p[0] * f.p[0] * f.p[1] * f.p[1];
        fd.c1 = 2.0 / M_PI;
        fd.magn2 = f.p[2] * f.p[2];
        mul = f.p[1];
        break;
    }
    if (f.kind == GPL_CF_CONST) {
      fd.nf = 0;
      continue;
    }
    fd.nf = f.periodic ? 2 * nsel : nsel;
    if (f.periodic && want_aux) fd.auxoff = naux;
    for (int c = 0; c < nsel; ++c) {
      PrepCol pc;
      memset(&pc, 0, sizeof(pc));
      pc.src = f.nvars > 0 ? f.vars[c] : c;
      if (pc.src < 0 || pc.src >= d) {
        set_error("cf model: column index %d out of range (d = %d)", pc.src, d);
        return GPL_ERR_BAD_ARG;
      }
      pc.dst = nfeat + c;
      pc.dst2 = f.periodic ? nfeat + nsel + c : -1;
      pc.dst_aux = (f.periodic && want_aux) ? naux + c : -1;
      pc.normalize = (f.normalize && !f.periodic && f.means && f.scales) ? 1 : 0;
      pc.periodic = f.periodic ? 1 : 0;
      pc.mean = pc.normalize ? f.means[c] : 0.0;
      pc.scale = pc.normalize ? f.scales[c] : 1.0;
      pc.w = f.periodic ? (2.0 * M_PI / f.period) : 1.0;
      pc.mul = mul;
      out->cols.push_back(pc);
    }
    nfeat += fd.nf;
    if (f.periodic && want_aux) naux += nsel;
  }
  p.nfactors = nf_out;
  p.nfeat = nfeat;
  p.naux = naux;
  return GPL_OK;
}

int prep_features(Ctx* ctx, const HostPlan& hp, const double* d_x, int64_t n, int64_t ldx,
                  double* d_z, int64_t ldz, double* d_zaux, int64_t ldzaux) {
  const int ncols = (int)hp.cols.size();
  if (ncols == 0 || n == 0) return GPL_OK;
  std::vector<PrepColDev> h(ncols);
  for (int c = 0; c < ncols; ++c) {
    const PrepCol& pc = hp.cols[c];
    h[c].src = pc.src; h[c].dst = pc.dst; h[c].dst2 = pc.dst2; h[c].dst_aux = pc.dst_aux; h[c].pad = 0;
    h[c].flags = (pc.normalize ? 1 : 0) | (pc.periodic ? 2 : 0);
    h[c].mean = pc.mean; h[c].scale = pc.scale; h[c].w = pc.w; h[c].mul = pc.mul;
  }
  PrepColDev* d_cols = nullptr;
  GPL_CUDA(cudaMallocAsync(&d_cols, sizeof(PrepColDev) * ncols, ctx->stream));
  GPL_CUDA(cudaMemcpyAsync(d_cols, h.data(), sizeof(PrepColDev) * ncols, cudaMemcpyHostToDevice,
                           ctx->stream));
  const int threads = 256;
  prep_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, ctx->stream>>>(
      d_x, n, ldx, d_z, ldz, d_zaux, ldzaux, d_cols, ncols);
  GPL_LAUNCH_CHECK(ctx);
  // h goes out of scope after the (pageable -> staged) copy has been issued, which is safe:
  // cudaMemcpyAsync from pageable memory returns after staging.
  GPL_CUDA(cudaFreeAsync(d_cols, ctx->stream));
  return GPL_OK;
}

int cf_pairwise(Ctx* ctx, const CfPlan& plan, const double* z1, int64_t n1, int64_t ldz1,
                const double* z2, int64_t n2, int64_t ldz2, bool same, double jitter, double* K,
                int64_t ldk) {
  if (n1 <= 0 || n2 <= 0) return GPL_OK;
  if (!ctx->cf_direct) {
    int done = 0;
    GPL_TRY(cf_pairwise_gram(ctx, plan, z1, n1, ldz1, z2, n2, ldz2, same, jitter, K, ldk, &done));
    if (done) return GPL_OK;
  }
  PairArgs a;
  a.z1 = z1; a.z2 = same ? z1 : z2; a.K = K;
  a.ldz1 = ldz1; a.ldz2 = same ? ldz1 : ldz2; a.ldk = ldk;
  a.n1 = (int)n1; a.n2 = same ? (int)n1 : (int)n2;
  a.same = same ? 1 : 0;
  a.tiles_i = (int)((n1 + TILE - 1) / TILE);
  a.tiles_j = (int)((a.n2 + TILE - 1) / TILE);
  a.jitter = same ? jitter : 0.0;
  const int grid = same ? a.tiles_i * (a.tiles_i + 1) / 2 : a.tiles_i * a.tiles_j;
  static const int direct_only = getenv("GPL_KBUILD_DIRECT") ? atoi(getenv("GPL_KBUILD_DIRECT")) : 0;
  if (plan.nfactors == 1 && plan.nfeat >= 1 && plan.nfeat <= 32 && !direct_only && !ctx->cf_direct) {
    switch (plan.f[0].kind) {
      case GPL_CF_SEXP: return launch_pair_gram<GPL_CF_SEXP>(ctx, plan, a, grid);
      case GPL_CF_MATERN32: return launch_pair_gram<GPL_CF_MATERN32>(ctx, plan, a, grid);
      case GPL_CF_MATERN52: return launch_pair_gram<GPL_CF_MATERN52>(ctx, plan, a, grid);
      default: break;
    }
  }
  if (plan.nfactors == 1) {
    switch (plan.f[0].kind) {
      case GPL_CF_SEXP: return launch_pair<8, GPL_CF_SEXP>(ctx, plan, a, grid);
      case GPL_CF_MATERN32: return launch_pair<8, GPL_CF_MATERN32>(ctx, plan, a, grid);
      case GPL_CF_MATERN52: return launch_pair<8, GPL_CF_MATERN52>(ctx, plan, a, grid);
      case GPL_CF_NN: return launch_pair<8, GPL_CF_NN>(ctx, plan, a, grid);
      case GPL_CF_LIN: return launch_pair<8, GPL_CF_LIN>(ctx, plan, a, grid);
      default: break;
    }
  }
  return launch_pair<4, -1>(ctx, plan, a, grid);
}

int cf_diag(Ctx* ctx, const CfPlan& plan, const double* z1, int64_t ldz1, const double* z2,
            int64_t ldz2, int64_t n, double add, double* out) {
  if (n <= 0) return GPL_OK;
  DiagArgs a{z1, z2, ldz1, ldz2, n, add, out};
  diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(plan, a);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
}

}  // namespace gpl
