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

#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace gpl {

namespace {

constexpr int TILE = 128;
constexpr int GTHREADS = 256;
constexpr int GLD = TILE + 4;        // smem row stride, == 4 mod 16: conflict-free fragment loads
constexpr int MAX_ROWS = 48;         // staged rows per side: 2 CTAs / SM

static __device__ const double kExpTab256[256] = {
    1.0, 1.0027112750502024854, 1.0054299011128028214, 1.0081558981184175158,
    1.01088928605170046, 1.0136300849514894388, 1.0163783149109530379, 1.0191339960777379497,
    1.0218971486541166782, 1.0246677928971356451, 1.0274459491187636965, 1.0302316376860410129,
    1.0330248790212284225, 1.03582569360195712, 1.0386341019613787906, 1.0414501246883161413,
    1.0442737824274138403, 1.0471050958792898661, 1.0499440858006872661, 1.0527907730046263271,
    1.0556451783605571588, 1.0585073227945126901, 1.061377227289262081, 1.0642549128844645498,
    1.0671404006768236182, 1.0700337118202417735, 1.0729348675259755514, 1.0758438890627910378,
    1.0787607977571197937, 1.0816856149932152019, 1.0846183622133092378, 1.0875590609177696653,
    1.0905077326652576592, 1.0934643990728858542, 1.0964290818163768234, 1.0994018026302219855,
    1.1023825833078409436, 1.1053714457017412556, 1.108368411723678638, 1.1113735033448176039,
    1.1143867425958925363, 1.1174081515673691991, 1.1204377524096066844, 1.1234755673330198007,
    1.1265216186082418998, 1.129575928566288146, 1.132638519598719228, 1.1357094141578055142,
    1.1387886347566916537, 1.1418762039695616227, 1.1449721444318042194, 1.1480764788401790068,
    1.1511892299529827058, 1.1543104205902160395, 1.1574400736337510296, 1.1605782120274987464,
    1.1637248587775775138, 1.1668800369524815706, 1.1700437696832501881, 1.1732160801636372475,
    1.1763969916502812763, 1.1795865274628759455, 1.1827847109843410299, 1.1859915656609938314,
    1.1892071150027210667, 1.1924313825831512221, 1.1956643920398273746, 1.1989061670743804818,
    1.2021567314527031421, 1.2054161090051238256, 1.2086843236265815774, 1.2119613992768011945,
    1.2152473599804688781, 1.2185422298274083618, 1.2218460329727575169, 1.2251587936371454377,
    1.2284805361068700057, 1.2318112847340759359, 1.2351510639369333057, 1.2384998981998165678,
    1.2418578120734840486, 1.2452248301752579328, 1.2486009771892047366, 1.2519862778663162701,
    1.2553807570246910896, 1.2587844395497164431, 1.262197350394250708, 1.2656195145788063242,
    1.2690509571917332226, 1.2724917033894027512, 1.2759417783963921004, 1.2794012075056692269,
    1.2828700160787782807, 1.2863482295460255336, 1.2898358734066658122, 1.2933329732290894367,
    1.2968395546510096659, 1.300355643379650651, 1.3038812651919358986, 1.3074164459346772448,
    1.3109612115247643419, 1.3145155879493546585, 1.3180796012660639947, 1.3216532776031575143,
    1.3252366431597412946, 1.3288297242059543955, 1.3324325470831614494, 1.3360451382041457734,
    1.3396675240533030054, 1.3432997311868352638, 1.3469417862329458358, 1.3505937158920343914,
    1.3542555469368927283, 1.3579273062129010465, 1.3616090206382247556, 1.3653007172040118154,
    1.3690024229745906119, 1.3727141650876683693, 1.3764359707545301002, 1.3801678672602380956,
    1.3839098819638319549, 1.387662042298529159, 1.3914243757719261871, 1.3951969099662001783,
    1.3989796725383111402, 1.4027726912202047064, 1.4065759938190154425, 1.4103896082172707044,
    1.4142135623730950488, 1.4180478843204151978, 1.4218926021691655598, 1.4257477441054943068,
    1.4296133383919700112, 1.4334894133677888425, 1.4373759974489823256, 1.4412731191286256627,
    1.44518080697704662, 1.4490990896420349805, 1.4530279958490525648, 1.4569675544014438216,
    1.4609177941806469887, 1.4648787441464058266, 1.4688504333369819269, 1.472832890869367595,
    1.4768261459394993114, 1.4808302278224717705, 1.4848451658727525005, 1.4888709895243970654,
    1.4929077282912648492, 1.4969554117672354268, 1.5010140696264255203, 1.5050837316234065441,
    1.5091644275934227398, 1.5132561874526099022, 1.5173590411982146988, 1.5214730189088145835,
    1.5255981507445383069, 1.5297344669472870239, 1.5338819978409560018, 1.5380407738316569284,
    1.5422108254079408236, 1.5463921831410215549, 1.5505848776849999591, 1.5547889397770885721,
    1.559004400237836967, 1.5632312899713577041, 1.5674696399655528931, 1.5717194812923413692,
    1.5759808451078864865, 1.5802537626528245268, 1.5845382652524937301, 1.5888343843171639442,
    1.5931421513422668979, 1.5974615979086270983, 1.6017927556826933538, 1.6061356564167709257,
    1.6104903319492543082, 1.6148568142048606401, 1.6192351351948637493, 1.6236253270173288318,
    1.6280274218573477668, 1.6324414519872750703, 1.6368674497669644878, 1.6413054476440062287,
    1.6457554781539648445, 1.650217573920617751, 1.6546917676561943976, 1.6591780921616160854,
    1.663676580326736435, 1.6681872651305825067, 1.6727101796415965736, 1.6772453570178785518,
    1.6817928305074290861, 1.6863526334483932959, 1.690924799269305182, 1.695509361489332695,
    1.7001063537185234695, 1.7047158096580512237, 1.7093377631004628271, 1.7139722479299260386,
    1.7186192981224779156, 1.7232789477462738972, 1.7279512309618375622, 1.7326361820223110649,
    1.737333835273706249, 1.7420442251551564435, 1.7467673861991689401, 1.7515033530318781561,
    1.7562521603732994831, 1.7610138430375838242, 1.765788435933272821, 1.7705759740635547735,
    1.7753764925265212526, 1.780190026515424409, 1.7850166113189349801, 1.789856282321400996,
    1.7947090750031071864, 1.7995750249405350921, 1.8044541678066238803, 1.8093465393710318685,
    1.8142521755003987562, 1.8191711121586085683, 1.824103385407053311, 1.8290490314048973431,
    1.8340080864093424635, 1.8389805867758937177, 1.8439665689586259251, 1.8489660695104509284,
    1.8539791250833855684, 1.8590057724288203846, 1.8640460483977890448, 1.8690999899412385049,
    1.8741676341102999013, 1.879249018056560178, 1.8843441790323344497, 1.8894531543909391035,
    1.8945759815869656413, 1.8997126981765552647, 1.9048633418176742037, 1.9100279502703897924,
    1.9152065613971472939, 1.9203992131630474739, 1.925605943636124929, 1.9308267909876271679,
    1.9360617934922944506, 1.9413109895286403852, 1.9465744175792332863, 1.9518521162309782959,
    1.957144124175400269, 1.9624504802089274268, 1.9677712232331757781, 1.9731063922552343129,
    1.9784560263879509683, 1.9838201648502193699, 1.9891988469672663513, 1.9945921121709402525,
};

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

__device__ __forceinline__ void dmma_k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// u <- max(u, ~1e-300) on the integer pipe (negative and zero contractions become a tiny positive
// number; a canonical NaN, 0x7ff8..., compares above and survives), then u <- min(u, 800) for finite u
__device__ __forceinline__ double clamp_arg(double u) {
  int hi = __double2hiint(u);
  const int lo = __double2loint(u);
  hi = max(hi, 0x01a56e1f);
  if (hi >= 0x40890000 && hi <= 0x7ff00000) return 800.0;
  return __hiloint2double(hi, lo);
}

// exp(-u), 0 <= u <= 800: k = rint(-u * 256/ln2) by the magic-number trick, r = -u - k ln2/256 (two-part),
// exp(-u) = 2^(k >> 8) * T[k & 255] * p3(r).  8 fp64 instructions; NaN propagates (k = 0 for a NaN).
__device__ __forceinline__ double exp_m(double u, const double* __restrict__ tab) {
  const double magic = 6755399441055744.0;                 // 1.5 * 2^52
  const double t = fma(u, -369.3299304675746322841407, magic);
  const int k = __double2loint(t);
  const double kd = t - magic;
  double r = fma(kd, -0.002707604318857193, -u);
  r = fma(kd, -1.855205093371747e-09, r);
  double p = fma(0.1666666806316513, r, 0.500000063263577);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double v = tab[k & 255] * p;
  const int e = max(k >> 8, -1022);
  return __hiloint2double(__double2hiint(v) + (e << 20), __double2loint(v));
}

// sqrt(u) for u >= 1e-300: MUFU.RSQ64H seed (2^-22), one third-order step (error e^3)
__device__ __forceinline__ double sqrt_pos(double u) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u));
  const double g0 = u * y;
  const double e = fma(-g0, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(g0, p * e, g0);
}

template <int KIND>
__device__ __forceinline__ double gvalue(const GFactor& f, double s, double rb, double rc,
                                         const double* __restrict__ tab) {
  if (KIND == GPL_CF_CONST) return f.magn2;
  if (KIND == GPL_CF_LIN) return f.magn2 * s;
  if (KIND == GPL_CF_SEXP) return f.magn2 * exp_m(clamp_arg(s), tab);
  if (KIND == GPL_CF_MATERN32) {
    const double u = clamp_arg(s);
    const double t = sqrt_pos(u);
    return fma(f.magn2, t, f.magn2) * exp_m(t, tab);
  }
  if (KIND == GPL_CF_MATERN52) {
    const double u = clamp_arg(s);
    const double t = sqrt_pos(u);
    return fma(f.m3, u, fma(f.magn2, t, f.magn2)) * exp_m(t, tab);
  }
  // cf_nn (src/cf.cpp:149-158): magn^2 (2/pi) asin(2a / sqrt((1 + 2b)(1 + 2c))), a = c0 + x1.x2
  return f.m3 * asin(2.0 * (f.c0 + s) * rb * rc);
}

__device__ __forceinline__ double gvalue_dyn(const GFactor& f, double s, double rb, double rc,
                                             const double* __restrict__ tab) {
  switch (f.kind) {
    case GPL_CF_CONST: return gvalue<GPL_CF_CONST>(f, s, rb, rc, tab);
    case GPL_CF_LIN: return gvalue<GPL_CF_LIN>(f, s, rb, rc, tab);
    case GPL_CF_SEXP: return gvalue<GPL_CF_SEXP>(f, s, rb, rc, tab);
    case GPL_CF_MATERN32: return gvalue<GPL_CF_MATERN32>(f, s, rb, rc, tab);
    case GPL_CF_MATERN52: return gvalue<GPL_CF_MATERN52>(f, s, rb, rc, tab);
    default: return gvalue<GPL_CF_NN>(f, s, rb, rc, tab);
  }
}

__device__ __forceinline__ bool is_stationary(int kind) {
  return kind == GPL_CF_SEXP || kind == GPL_CF_MATERN32 || kind == GPL_CF_MATERN52;
}

// MODE >= 0: the model is ONE factor of that kind.  MODE == -1: any sum of products.
// NSUB: 8-column sub-tiles per warp and pass (4: warp tile 32 x 32, two passes; 2: 32 x 16, four passes).
template <int MODE, int NSUB>
__global__ void __launch_bounds__(GTHREADS, 2)
gram_kernel(const __grid_constant__ GPlan plan, const GArgs a) {
  extern __shared__ double sm[];
  const int R = plan.rows;
  double* Zs1 = sm;                  // [R][GLD]
  double* Zs2 = sm + R * GLD;
  double* tab = sm + 2 * R * GLD;    // [256]
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  tab[tid] = kExpTab256[tid];
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

  // ---- stage: features (shifted and scaled for the stationary kinds; -2 folded into the column side)
  for (int fi = 0; fi < plan.nfactors; ++fi) {
    const GFactor& f = plan.f[fi];
    const bool stat = is_stationary(f.kind);
    for (int idx = tid; idx < (f.gk + 1) * TILE; idx += GTHREADS) {
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
      }
      Zs1[(f.goff + j) * GLD + r] = v1;
      Zs2[(f.goff + j) * GLD + r] = v2;
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
    double sacc = 0.0;
    for (int j = 0; j < f.nf; ++j) {
      const double v = Zs[(f.goff + j) * GLD + r];
      sacc = fma(v, v, sacc);
    }
    if (stat) {
      if (tid >= TILE) sacc *= 0.25;                      // the column side holds -2 b
      Zs[(f.goff + f.nf) * GLD + r] = (tid < TILE) ? sacc : 1.0;
      Zs[(f.goff + f.nf + 1) * GLD + r] = (tid < TILE) ? 1.0 : sacc;
      // cancellation guard: the relative error of K is ~ 4 eps * (scaled |a|^2 + |b|^2) * |dlog k / du|
      // (d log k / du is 1 for sexp and at most 1/2 for the Matern kinds, also as u -> 0)
      if (!(sacc < 250.0)) s_bad = 1;                     // also catches NaN
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
    if (MODE < 0) {
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
              for (int j = 0; j < f.nf; ++j) {
                const double dlt = fma(0.5, Zs2[(f.goff + j) * GLD + col], Zs1[(f.goff + j) * GLD + row]);
                r2 = fma(dlt, dlt, r2);
              }
              acc[im][in][c] = r2;
            }
      }
      const bool is_nn = kind == GPL_CF_NN;
#pragma unroll
      for (int im = 0; im < 4; ++im) {
        const int row = wm * 32 + im * 8 + g;
        const double rb = is_nn ? Zs1[(f.goff + f.gk) * GLD + row] : 0.0;
#pragma unroll
        for (int in = 0; in < NSUB; ++in) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int col = cbase + in * 8 + 2 * t + c;
            const double rc = is_nn ? Zs2[(f.goff + f.gk) * GLD + col] : 0.0;
            double s = acc[im][in][c];
            if (diag_tile && stat && row == col) s = 0.0;          // k(x, x) exactly
            const double v = MODE < 0 ? gvalue_dyn(f, s, rb, rc, tab) : gvalue<MODE < 0 ? 0 : MODE>(f, s, rb, rc, tab);
            if (MODE < 0) {
              prod[im][in][c] = f.first ? v : prod[im][in][c] * v;
              if (f.last) out[im][in][c] += prod[im][in][c];
            } else {
              out[im][in][c] = v;
            }
          }
        }
      }
    }
    // ---- store: thread owns row g, columns 2t, 2t + 1 of each 8 x 8 sub-tile; the mirrored entry pair
    // is one 16-byte store
#pragma unroll
    for (int im = 0; im < 4; ++im) {
      const int row = wm * 32 + im * 8 + g;
      const int gi = i0 + row;
#pragma unroll
      for (int in = 0; in < NSUB; ++in) {
        const int col = cbase + in * 8 + 2 * t;
        const int gj = j0 + col;
        double v0 = out[im][in][0], v1 = out[im][in][1];
        if (diag_tile) {
          if (gi == gj) v0 += a.jitter;
          if (gi == gj + 1) v1 += a.jitter;
        }
        if (interior) {
          a.K[(int64_t)gj * a.ldk + gi] = v0;
          a.K[(int64_t)(gj + 1) * a.ldk + gi] = v1;
          if (mirror) {
            double* dst = a.K + (int64_t)gi * a.ldk + gj;
            if (vec_ok) {
              *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
            } else {
              dst[0] = v0;
              dst[1] = v1;
            }
          }
        } else {
          if (gi < a.n1 && gj < a.n2) {
            a.K[(int64_t)gj * a.ldk + gi] = v0;
            if (mirror) a.K[(int64_t)gi * a.ldk + gj] = v0;
          }
          if (gi < a.n1 && gj + 1 < a.n2) {
            a.K[(int64_t)(gj + 1) * a.ldk + gi] = v1;
            if (mirror) a.K[(int64_t)gi * a.ldk + gj + 1] = v1;
          }
        }
      }
    }
  }
}

template <int MODE, int NSUB>
int launch_gram(Ctx* ctx, const GPlan& gp, const GArgs& a, int grid) {
  const int smem = (2 * gp.rows * GLD + 256) * (int)sizeof(double);
  GPL_CUDA(cudaFuncSetAttribute(gram_kernel<MODE, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  gram_kernel<MODE, NSUB><<<grid, GTHREADS, smem, ctx->stream>>>(gp, a);
  GPL_LAUNCH_CHECK(ctx);
  return GPL_OK;
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
  return launch_gram<-1, 2>(ctx, gp, a, grid);
}

}  // namespace gpl
