// Internal declarations shared by the translation units of libgplite_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "gplite_b200.h"

namespace gpl {

constexpr int NB = 128;   // base block of the recursive factorisation / solves

void set_error(const char* fmt, ...);

// gridDim.y is limited to 65535: kernels that map columns to blockIdx.y stride over them
inline unsigned grid_y_cap(int64_t n) { return (unsigned)(n < 65535 ? (n > 0 ? n : 1) : 65535); }

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // high-priority side stream + events for the look-ahead Cholesky (diagonal-block factorisation
  // of step k+1 overlaps the trailing update of step k)
  cudaStream_t panel_stream = nullptr;
  cudaEvent_t ev_col = nullptr, ev_diag = nullptr;
  // side streams for independent latency-bound chains (the diagonal-block inversions of trtri)
  static constexpr int NAUX = 4;
  cudaStream_t aux[NAUX] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t aux_ev[NAUX] = {nullptr, nullptr, nullptr, nullptr};
  int sm_count = 148;
  int64_t launches = 0;
  // scratch: diagonal-block inverses of the most recent factorisation living in this ctx
  double* dinv = nullptr;       // nblk * NB * NB doubles
  int64_t dinv_blocks = 0;
  int32_t* d_info = nullptr;    // device int[4]
  double* d_scal = nullptr;     // device double[64] for small reductions
  double* h_scal = nullptr;     // pinned host mirror
  // optional per-launch timing of the DMMA GEMM kernel (bench.py's roofline leg): one event
  // pair per launch on the launching stream, harvested by gpl_ctx_profile_read
  bool profile = false;
  int gemm_force = 0;   // tuning knob: 0 = automatic tile choice, 1/2/3 = force the 128/64/32 tile
  // cf_pairwise: direct differences on the fp64 pipe (the reference's own arithmetic, src/cf.cpp:27-36) instead of
  // the Gram contraction on the tensor pipe.  Set by the FITC code around its K builds: Kz + jitter I is nearly
  // singular by construction, and the rounding of |a|^2 + |b|^2 - 2 a.b (1e-14 relative in K) is amplified there.
  bool cf_direct = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  double prof_flops = 0.0;
  double prof_ms = 0.0;
  int64_t prof_launches = 0;
  struct GemmRec { int m, n, k, lower, kskip, tile; };
  std::vector<GemmRec> prof_recs;   // one per profiled launch (GPL_GEMM_LOG=path dumps them with times)
};

#define GPL_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      gpl::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,                 \
                     cudaGetErrorName(e__), cudaGetErrorString(e__));               \
      return GPL_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define GPL_TRY(call)                      \
  do {                                     \
    int s__ = (call);                      \
    if (s__ != GPL_OK) return s__;         \
  } while (0)

#define GPL_LAUNCH_CHECK(ctx)                                                       \
  do {                                                                              \
    (ctx)->launches++;                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      gpl::set_error("%s:%d kernel launch failed %s: %s", __FILE__, __LINE__,       \
                     cudaGetErrorName(e__), cudaGetErrorString(e__));               \
      return GPL_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---- gemm_f64.cu ----
// Staircase products of the block-cyclic multi-GPU layout: C's columns come in blocks of `nb`
// (a multiple of 128); column block c is only touched from row first + c * step on, and its rows of
// the B operand start at c * (nb + bskip) instead of c * nb.
struct GemmStair { int64_t nb, first, step, bskip; };
int dgemm(Ctx* ctx, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha,
          const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
          int64_t ldc, bool lower, int alias = 0, int kskip = 0, const GemmStair* stair = nullptr);
// alias: 0 = C distinct from A and B; 1 = C is A (needs n <= 128); 2 = C is B (needs m <= 128);
// 3 = both (single tile)
// kskip: 1 = skip the k-tiles above each output tile's first ROW (op(A) = A^T with A lower
// triangular in k), 2 = above its first COLUMN (B lower triangular), 3 = beyond its last row (A lower
// triangular); the skipped tiles are zeros (or never-written scratch above the diagonal).

// ---- chol.cu ----
int ensure_dinv(Ctx* ctx, int64_t n);
int dpotrf(Ctx* ctx, int64_t n, double* A, int64_t lda, int32_t* info_host, double* logdet_host);
int dtrsm_right(Ctx* ctx, bool trans, int64_t m, int64_t n, const double* L, int64_t ldl,
                double* B, int64_t ldb, int64_t blk0);
int dpotri(Ctx* ctx, int64_t n, double* A, int64_t lda);
int dpotrf_nosync(Ctx* ctx, double* dinv, int64_t n, double* A, int64_t lda, int* d_info, int pivot0);
int dtrtri_block(Ctx* ctx, double* dinv, int64_t n, const double* L, int64_t ldl, double* out, int64_t ldo);
// Y (lower) <- X^T X for lower-triangular X with clean (zero) upper parts in its diagonal blocks:
// one triangular-aware SYRK launch, out of place (Y must not alias X)
int lauum_oop(Ctx* ctx, int64_t n, const double* X, int64_t ldx, double* Y, int64_t ldy);

}  // namespace gpl
