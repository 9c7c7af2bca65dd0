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
// cf_gram.cu: every factor through the FP64 tensor pipe (sums of products, rectangular blocks);
// *done = 0 when the model does not fit its shared-memory budget
int cf_pairwise_gram(Ctx* ctx, const CfPlan& plan, const double* z1, int64_t n1, int64_t ldz1,
                     const double* z2, int64_t n2, int64_t ldz2, bool same, double jitter, double* K,
                     int64_t ldk, int* done);
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
// exp(x) for x <= 0, table driven: k = rint(x * 64/ln2) by the magic-number trick,
// r = x - k ln2/64 (two-part, |r| <= 0.0055), exp(x) = 2^(k>>6) * T[k&63] * p5(r) with the degree-5
// Taylor polynomial (truncation 4e-17).  11 fp64 operations instead of 17 for the table-free
// degree-12 form.  Branch-free on purpose (a rare-path branch stops ptxas from interleaving the 32
// independent entries a thread owns); the exponent is clamped at the smallest normal, and the
// scale is applied by a multiply so that NaN stays NaN (the reference's chol() then fails on it).
// `tab` points at the 64-entry table 2^(j/64): shared memory in the K build, global elsewhere.
static __device__ const double kExpTab[64] = {
    1.0, 1.01088928605170046, 1.0218971486541166782, 1.0330248790212284225,
    1.0442737824274138403, 1.0556451783605571588, 1.0671404006768236182, 1.0787607977571197937,
    1.0905077326652576592, 1.1023825833078409436, 1.1143867425958925363, 1.1265216186082418998,
    1.1387886347566916537, 1.1511892299529827058, 1.1637248587775775138, 1.1763969916502812763,
    1.1892071150027210667, 1.2021567314527031421, 1.2152473599804688781, 1.2284805361068700057,
    1.2418578120734840486, 1.2553807570246910896, 1.2690509571917332226, 1.2828700160787782807,
    1.2968395546510096659, 1.3109612115247643419, 1.3252366431597412946, 1.3396675240533030054,
    1.3542555469368927283, 1.3690024229745906119, 1.3839098819638319549, 1.3989796725383111402,
    1.4142135623730950488, 1.4296133383919700112, 1.44518080697704662, 1.4609177941806469887,
    1.4768261459394993114, 1.4929077282912648492, 1.5091644275934227398, 1.5255981507445383069,
    1.5422108254079408236, 1.559004400237836967, 1.5759808451078864865, 1.5931421513422668979,
    1.6104903319492543082, 1.6280274218573477668, 1.6457554781539648445, 1.663676580326736435,
    1.6817928305074290861, 1.7001063537185234695, 1.7186192981224779156, 1.737333835273706249,
    1.7562521603732994831, 1.7753764925265212526, 1.7947090750031071864, 1.8142521755003987562,
    1.8340080864093424635, 1.8539791250833855684, 1.8741676341102999013, 1.8945759815869656413,
    1.9152065613971472939, 1.9360617934922944506, 1.957144124175400269, 1.9784560263879509683,
};
static __constant__ double kExpC[8] = {
    9.23324826168936567683e+01,    // [0] 64/ln2
    6755399441055744.0,            // [1] 1.5 * 2^52
    -1.08304172754287719727e-02,   // [2] -ln2/64 (hi, low 32 bits zero)
    -7.42082037348716703866e-09,   // [3] -ln2/64 (lo)
    8.33333333333333333333e-03,    // [4] 1/5!
    4.16666666666666666667e-02,    // [5] 1/4!
    1.66666666666666666667e-01,    // [6] 1/3!
    0.0};

__device__ __forceinline__ double exp_neg_fast(double x, const double* __restrict__ tab) {
  const double magic = kExpC[1];
  const double t = fma(x, kExpC[0], magic);
  const int k = __double2loint(t);
  const double kd = t - magic;
  double r = fma(kd, kExpC[2], x);
  r = fma(kd, kExpC[3], r);
  double p = fma(kExpC[4], r, kExpC[5]);
  p = fma(p, r, kExpC[6]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int e = max(k >> 6, -1022);
  return (tab[k & 63] * p) * __hiloint2double((e + 1023) << 20, 0);
}

// sqrt(s) for s >= 0: MUFU.RSQ64H seed y0 (2^-22), then one third-order step
// g = g0 (1 + e/2 + 3 e^2/8), g0 = s y0, e = 1 - g0 y0  (error e^3).  1e-300 is added so that
// s == 0 gives 1e-150 (indistinguishable from 0 in every kernel formula) instead of 0 * inf;
// NaN propagates.
__device__ __forceinline__ double sqrt_fast(double s) {
  s += 1e-300;
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  const double g0 = s * y;
  const double e = fma(-g0, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(g0, p * e, g0);
}

// value of one factor given its accumulated statistics
//   stationary kinds: s = r2 ; lin: s = dot ; nn: s = dot, sb = |z1|^2, sc = |z2|^2
__device__ __forceinline__ double factor_value(const FactorDev& f, double s, double sb, double sc,
                                               const double* __restrict__ tab = kExpTab) {
  switch (f.kind) {
    case GPL_CF_CONST:
      return f.magn2;
    case GPL_CF_LIN:
      return f.magn2 * s;
    case GPL_CF_SEXP:
      return f.magn2 * exp_neg_fast(f.c0 * s, tab);              // c0 = -0.5 / lscale^2
    case GPL_CF_MATERN32: {
      const double t = f.c0 * sqrt_fast(s);                 // c0 = sqrt(3) / lscale
      return fma(f.magn2, t, f.magn2) * exp_neg_fast(-t, tab);
    }
    case GPL_CF_MATERN52: {
      const double r = sqrt_fast(s);
      const double t = f.c0 * r;                            // c0 = sqrt(5) / lscale
      // magn2 (1 + t + c1 r^2), c2 = magn2 * c1 = magn2 * 5 / (3 lscale^2); r^2 == s to 2 ulp
      return fma(f.c2, s, fma(f.magn2, t, f.magn2)) * exp_neg_fast(-t, tab);
    }
    case GPL_CF_NN: {
      const double a = f.c0 + s, b = f.c0 + sb, c = f.c0 + sc;   // c0 = sigma0^2 sigma^2
      return f.magn2 * f.c1 * asin(2.0 * a / sqrt((1.0 + 2.0 * b) * (1.0 + 2.0 * c)));  // c1 = 2/pi
    }
  }
  return 0.0;
}

}  // namespace gpl
