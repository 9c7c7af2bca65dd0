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

// value of one factor given its accumulated statistics
//   stationary kinds: s = r2 ; lin: s = dot ; nn: s = dot, sb = |z1|^2, sc = |z2|^2
__device__ __forceinline__ double factor_value(const FactorDev& f, double s, double sb, double sc) {
  switch (f.kind) {
    case GPL_CF_CONST:
      return f.magn2;
    case GPL_CF_LIN:
      return f.magn2 * s;
    case GPL_CF_SEXP:
      return f.magn2 * exp(f.c0 * s);                       // c0 = -0.5 / lscale^2
    case GPL_CF_MATERN32: {
      const double t = f.c0 * sqrt(s);                      // c0 = sqrt(3) / lscale
      return f.magn2 * (1.0 + t) * exp(-t);
    }
    case GPL_CF_MATERN52: {
      const double r = sqrt(s);
      const double t = f.c0 * r;                            // c0 = sqrt(5) / lscale
      return f.magn2 * (1.0 + t + f.c1 * (r * r)) * exp(-t);   // c1 = 5 / (3 lscale^2)
    }
    case GPL_CF_NN: {
      const double a = f.c0 + s, b = f.c0 + sb, c = f.c0 + sc;   // c0 = sigma0^2 sigma^2
      return f.magn2 * f.c1 * asin(2.0 * a / sqrt((1.0 + 2.0 * b) * (1.0 + 2.0 * c)));  // c1 = 2/pi
    }
  }
  return 0.0;
}

}  // namespace gpl
