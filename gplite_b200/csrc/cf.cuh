// Device-side description of a flattened covariance model and the per-pair evaluation shared by
// the K-build, diagonal and gradient-trace kernels.
#pragma once

#include "common.cuh"

namespace gpl {

constexpr int MAX_FACTORS = 16;

struct FactorDev {
  int kind;       // gpl_cf_kind
  int term;       // summand index
  int off;        // first feature column in Z
  int nf;         // number of feature columns
  int auxoff;     // first column in Zaux holding the raw angles w*x of a periodic factor, or -1
  double c0, c1, c2, c3;   // precomputed constants, see make_plan()
  double magn2;
};

struct CfPlan {
  int nfactors;
  int nfeat;              // total feature columns F
  int naux;               // columns of Zaux (only when the plan was made for gradients)
  unsigned long long term_mask;   // bit per summand (0 = all)
  FactorDev f[MAX_FACTORS];
};

// Host-side: everything needed to turn raw x into the feature matrix Z
struct PrepCol {
  int src;        // source column in x
  int dst;        // destination column in Z (sin part for periodic)
  int dst2;       // cos part for periodic, -1 otherwise
  int dst_aux;    // column of Zaux receiving the angle w*x (gradient w.r.t. the period), or -1
  double mean;     // (x - mean) / scale, scale kept as a divisor: same rounding as R
  double mul;      // post-multiplier (cf_nn: sigma, as src/cf.cpp:155 scales the rows)
  double scale;
  double w;       // periodic: 2*pi/period, else unused
  int normalize, periodic;
};

struct HostPlan {
  CfPlan plan;
  std::vector<PrepCol> cols;
};

int make_plan(const gpl_cf_model* cf, int d, uint64_t term_mask, HostPlan* out,
              bool want_aux = false);

// Z (n x F, ld = ldz) from x (n x d, ld = ldx)
int prep_features(Ctx* ctx, const HostPlan& hp, const double* d_x, int64_t n, int64_t ldx,
                  double* d_z, int64_t ldz, double* d_zaux = nullptr, int64_t ldzaux = 0);

// K (n1 x n2, ld ldk) from feature matrices; same != 0 -> z2 ignored, symmetric path, jitter on diag
int cf_pairwise(Ctx* ctx, const CfPlan& plan, const double* z1, int64_t n1, int64_t ldz1,
                const double* z2, int64_t n2, int64_t ldz2, bool same, double jitter, double* K,
                int64_t ldk);
int cf_diag(Ctx* ctx, const CfPlan& plan, const double* z1, int64_t ldz1, const double* z2,
            int64_t ldz2, int64_t n, double add, double* out);

// ---- per-pair math ------------------------------------------------------------------------

// The K build is bound by fp64 instruction issue, not by HBM, unless the transcendental cost is
// cut: libdevice exp()/sqrt() cost ~40 / ~20 instructions per entry.  These two replacements
// are branch-free and ~1 ulp, far inside the 1e-12 parity budget.

// exp(x) for x <= 0.  k = rint(x log2 e) by the magic-number trick, r = x - k ln2 (two-part),
// degree-12 Taylor on |r| <= ln2/2 (truncation 1.7e-16), scaling by adding k to the exponent
// field.  Branch-free on purpose (a rare-path branch stops ptxas from interleaving the 32
// independent entries a thread owns): arguments are clamped at -708, i.e. entries whose true
// value is below 3.3e-308 come out as 3.3e-308 instead of a denormal / zero.
// Coefficients live in constant memory so that each DFMA reads its addend from the constant bank;
// as literals ptxas materialises every one of them with two UMOVs per use (~65 extra issue slots
// per K entry, measured).
static __constant__ double kExpC[16] = {
    2.08767569878680989792e-09,   // 1/12!
    2.50521083854417187751e-08,   // 1/11!
    2.75573192239858906526e-07,   // 1/10!
    2.75573192239858906526e-06,   // 1/9!
    2.48015873015873015873e-05,   // 1/8!
    1.98412698412698412698e-04,   // 1/7!
    1.38888888888888888889e-03,   // 1/6!
    8.33333333333333333333e-03,   // 1/5!
    4.16666666666666666667e-02,   // 1/4!
    1.66666666666666666667e-01,   // 1/3!
    1.4426950408889634074,        // [10] log2(e)
    6755399441055744.0,           // [11] 1.5 * 2^52
    -6.93147180369123816490e-01,  // [12] -ln2 (hi)
    -1.90821492927058770002e-10,  // [13] -ln2 (lo)
    0.0, 0.0};

__device__ __forceinline__ double exp_neg_fast(double x) {
  const double magic = kExpC[11];
  const double t = fma(x, kExpC[10], magic);
  int k = __double2loint(t);
  const double kd = t - magic;
  double r = fma(kd, kExpC[12], x);
  r = fma(kd, kExpC[13], r);
  double p = kExpC[0];
#pragma unroll
  for (int q = 1; q < 10; ++q) p = fma(p, r, kExpC[q]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  // 2^k as a double (k clamped at the smallest normal exponent), applied by a multiply so that a
  // NaN argument stays NaN (R's chol() then fails on it, as with the reference)
  k = max(k, -1022);
  return p * __hiloint2double((k + 1023) << 20, 0);
}

// sqrt(s) for s >= 0: MUFU.RSQ64H seed (2^-22) and two coupled Newton steps (~2 ulp).  1e-300 is
// added so that s == 0 gives 1e-150 (indistinguishable from 0 in every kernel formula) instead of
// 0 * inf; NaN propagates.
__device__ __forceinline__ double sqrt_fast(double s) {
  s += 1e-300;
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  double g = s * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  return g;
}

// value of one factor given its accumulated statistics
//   stationary kinds: s = r2 ; lin: s = dot ; nn: s = dot, sb = |z1|^2, sc = |z2|^2
__device__ __forceinline__ double factor_value(const FactorDev& f, double s, double sb, double sc) {
  switch (f.kind) {
    case GPL_CF_CONST:
      return f.magn2;
    case GPL_CF_LIN:
      return f.magn2 * s;
    case GPL_CF_SEXP:
      return f.magn2 * exp_neg_fast(f.c0 * s);              // c0 = -0.5 / lscale^2
    case GPL_CF_MATERN32: {
      const double t = f.c0 * sqrt_fast(s);                 // c0 = sqrt(3) / lscale
      return fma(f.magn2, t, f.magn2) * exp_neg_fast(-t);
    }
    case GPL_CF_MATERN52: {
      const double r = sqrt_fast(s);
      const double t = f.c0 * r;                            // c0 = sqrt(5) / lscale
      // magn2 (1 + t + c1 r^2), c2 = magn2 * c1 = magn2 * 5 / (3 lscale^2); r^2 == s to 2 ulp
      return fma(f.c2, s, fma(f.magn2, t, f.magn2)) * exp_neg_fast(-t);
    }
    case GPL_CF_NN: {
      const double a = f.c0 + s, b = f.c0 + sb, c = f.c0 + sc;   // c0 = sigma0^2 sigma^2
      return f.magn2 * f.c1 * asin(2.0 * a / sqrt((1.0 + 2.0 * b) * (1.0 + 2.0 * c)));  // c1 = 2/pi
    }
  }
  return 0.0;
}

}  // namespace gpl
