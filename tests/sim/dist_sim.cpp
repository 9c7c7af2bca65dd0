// TEST INFRASTRUCTURE ONLY -- a host back end for gplite_b200/csrc/dist_sched.hpp, so that the
// multi-GPU schedule of the product (ownership maps, staircase updates, panel packing, the order in
// which collectives are issued, the inverse sweeps) runs on CPUs: plain loops for the dense kernels,
// immediate "streams", and collectives that call back into the Python test (torch.distributed,
// gloo, world size 2 / 3).  Nothing in the product links or imports this file.
//
// Built by tests/test_dist_cpu.py:  g++ -O2 -shared -fPIC -I gplite_b200/csrc tests/sim/dist_sim.cpp
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "dist_sched.hpp"

using namespace gpl::dist;

typedef void (*bcast_cb)(double* buf, int64_t count, int root);
typedef void (*allreduce_cb)(double* buf, int64_t count);

struct SimOps {
  Layout* lay;
  const double* Bfull;      // the whole matrix B (n x n, ld n), same on every rank
  bcast_cb cb_bcast;
  allreduce_cb cb_allreduce;
  int first_bad = 0;
  int64_t collectives = 0;
  std::vector<std::vector<double>> inv;   // explicit inverse of each owned diagonal block

  int fill_block(Stream, int64_t b, double* dst, int64_t ld) {
    const int64_t c0 = lay->col0(b), w = lay->width(b), n = lay->n;
    for (int64_t j = 0; j < w; ++j)
      for (int64_t i = c0; i < n; ++i) dst[j * ld + (i - c0)] = Bfull[(c0 + j) * n + i];
    return 0;
  }
  int potrf_diag(Stream, int64_t k, int64_t w, double* A, int64_t lda, int pivot0) {
    for (int64_t j = 0; j < w; ++j) {
      double d = A[j * lda + j];
      for (int64_t p = 0; p < j; ++p) d -= A[p * lda + j] * A[p * lda + j];
      if (!(d > 0.0)) {
        if (first_bad == 0 || pivot0 + (int)j + 1 < first_bad) first_bad = pivot0 + (int)j + 1;
        d = NAN;
      }
      d = sqrt(d);
      A[j * lda + j] = d;
      for (int64_t i = j + 1; i < w; ++i) {
        double v = A[j * lda + i];
        for (int64_t p = 0; p < j; ++p) v -= A[p * lda + i] * A[p * lda + j];
        A[j * lda + i] = v / d;
      }
    }
    // explicit inverse (what the device keeps as 128-block inverses)
    if ((int64_t)inv.size() < lay->nblk) inv.resize((size_t)lay->nblk);
    std::vector<double>& X = inv[(size_t)k];
    X.assign((size_t)(w * w), 0.0);
    for (int64_t c = 0; c < w; ++c)
      for (int64_t i = c; i < w; ++i) {
        double v = (i == c) ? 1.0 : 0.0;
        for (int64_t p = c; p < i; ++p) v -= A[p * lda + i] * X[c * w + p];
        X[c * w + i] = v / A[i * lda + i];
      }
    return 0;
  }
  int trsm_below(Stream, int64_t, int64_t, int64_t, int64_t m, int64_t w, const double* L, int64_t ldl, double* B,
                 int64_t ldb) {
    for (int64_t j = 0; j < w; ++j)          // B <- B L^-T, column by column
      for (int64_t i = 0; i < m; ++i) {
        double v = B[j * ldb + i];
        for (int64_t p = 0; p < j; ++p) v -= B[p * ldb + i] * L[p * ldl + j];
        B[j * ldb + i] = v / L[j * ldl + j];
      }
    return 0;
  }
  int inv_diag(Stream, int64_t k, int64_t w, const double*, int64_t, double* out, int64_t ldo) {
    const std::vector<double>& X = inv[(size_t)k];
    for (int64_t c = 0; c < w; ++c)
      for (int64_t i = 0; i < w; ++i) out[c * ldo + i] = X[c * w + i];
    return 0;
  }
  int gemm(Stream, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
           const double* B, int64_t ldb, double beta, double* C, int64_t ldc, bool lower, const Stair* st) {
    std::vector<double> out((size_t)m);
    for (int64_t j = 0; j < n; ++j) {
      int64_t row0 = 0, jb = j;
      if (st) {
        const int64_t c = j / st->nb;
        row0 = st->first + c * st->step;        // rows above never exist for this column block
        jb = j + c * st->bskip;
      }
      if (lower && j > row0) row0 = j;
      for (int64_t i = row0; i < m; ++i) {
        double acc = 0.0;
        for (int64_t p = 0; p < k; ++p) {
          const double a = ta ? A[i * lda + p] : A[p * lda + i];
          const double b = tb ? B[p * ldb + jb] : B[jb * ldb + p];
          acc += a * b;
        }
        out[(size_t)i] = alpha * acc + (beta != 0.0 ? beta * C[j * ldc + i] : 0.0);
      }
      for (int64_t i = row0; i < m; ++i) C[j * ldc + i] = out[(size_t)i];
    }
    return 0;
  }
  int copy2d(Stream, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols) {
    std::vector<double> t((size_t)(rows * cols));           // src and dst may overlap (block-row write-back)
    for (int64_t j = 0; j < cols; ++j) memcpy(&t[(size_t)(j * rows)], src + j * lds, sizeof(double) * rows);
    for (int64_t j = 0; j < cols; ++j) memcpy(dst + j * ldd, &t[(size_t)(j * rows)], sizeof(double) * rows);
    return 0;
  }
  int zero(Stream, double* p, int64_t count) { memset(p, 0, sizeof(double) * (size_t)count); return 0; }
  int trsv(Stream, int64_t k, bool trans, int64_t w, const double*, int64_t, double* b, double* x) {
    const std::vector<double>& X = inv[(size_t)k];          // x = L^-1 b or L^-T b
    for (int64_t i = 0; i < w; ++i) {
      double v = 0.0;
      for (int64_t p = 0; p < w; ++p) v += (trans ? X[i * w + p] : X[p * w + i]) * b[p];
      x[i] = v;
    }
    return 0;
  }
  int gemv(Stream, bool trans, int64_t m, int64_t n, const double* A, int64_t lda, const double* x, double* y) {
    if (!trans) {
      for (int64_t i = 0; i < m; ++i) y[i] = 0.0;
      for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < m; ++i) y[i] += A[j * lda + i] * x[j];
    } else {
      for (int64_t j = 0; j < n; ++j) {
        double v = 0.0;
        for (int64_t i = 0; i < m; ++i) v += A[j * lda + i] * x[i];
        y[j] = v;
      }
    }
    return 0;
  }
  int axpy(Stream, int64_t n, double a, const double* x, double* y) {
    for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
    return 0;
  }
  int sub(Stream, int64_t n, const double* a, const double* b, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = a[i] - b[i];
    return 0;
  }
  int colsumsq(Stream, int64_t m, int64_t n, const double* A, int64_t lda, double* out) {
    for (int64_t j = 0; j < n; ++j) {
      double v = 0.0;
      for (int64_t i = 0; i < m; ++i) v += A[j * lda + i] * A[j * lda + i];
      out[j] = v;
    }
    return 0;
  }
  int diag_logsum(Stream, int64_t w, const double* A, int64_t lda, double* acc) {
    for (int64_t i = 0; i < w; ++i) acc[0] += 2.0 * log(A[i * lda + i]);
    return 0;
  }
  int record(Event, Stream) { return 0; }
  int wait(Stream, Event) { return 0; }
  int sync(Stream) { return 0; }
  int bcast(Stream, double* buf, int64_t count, int root) {
    ++collectives;
    if (lay->world > 1) cb_bcast(buf, count, root);
    return 0;
  }
  int allreduce_sum(Stream, double* buf, int64_t count) {
    ++collectives;
    if (lay->world > 1) cb_allreduce(buf, count);
    return 0;
  }
  int check_info(int* bad) {
    // min over ranks of the failing pivot through the sum-only callback: every rank contributes its
    // pivot (or 0) in its own slot
    std::vector<double> all((size_t)lay->world, 0.0);
    all[(size_t)lay->rank] = (double)first_bad;
    ++collectives;
    if (lay->world > 1) cb_allreduce(all.data(), lay->world);
    double best = 0.0;
    for (double v : all)
      if (v > 0.0 && (best == 0.0 || v < best)) best = v;
    *bad = (int)best;
    return 0;
  }
};

// Runs factor / solve / x sweep / diag / y sweep on this rank.  Outputs: logdet (local part, caller
// sums), solution (n), diag(B^-1) (n), and the owned block columns of L and of B^-1 (nloc * ld each,
// slab layout) for inspection.
extern "C" int sim_run(int rank, int world, int64_t n, int64_t nb, int window, const double* Bfull, const double* rhs,
                       bcast_cb cbb, allreduce_cb cba, int do_inverse, double* logdet_local, double* sol,
                       double* diag_binv, double* Lslab, double* Yslab, int64_t* nloc_out, int64_t* ld_out,
                       int* info, int64_t* ncollectives) {
  Layout lay;
  lay.init(n, nb, world, rank);
  SimOps ops;
  ops.lay = &lay;
  ops.Bfull = Bfull;
  ops.cb_bcast = cbb;
  ops.cb_allreduce = cba;
  ShardedChol<SimOps> ch;
  ch.lay = lay;
  ch.ops = &ops;
  const size_t slab = (size_t)(lay.nloc > 0 ? lay.nloc : 1) * lay.ld;
  // fill with NaN so that any read of a never-written entry shows up in the results
  std::vector<double> A(slab, NAN), X(slab, NAN);
  std::vector<std::vector<double>> pbuf((size_t)window, std::vector<double>((size_t)(nb * lay.ld), NAN));
  std::vector<double> tmp((size_t)(nb * (lay.nloc > nb ? lay.nloc : nb)), NAN);
  std::vector<double> v0((size_t)n), v1((size_t)n), v2((size_t)n), v3((size_t)n);
  double acc = 0.0;
  ch.A = A.data();
  ch.X = X.data();
  ch.window = window;
  ch.chunk_align = 32;            // the host back end solves against any sub-block
  for (int w = 0; w < window; ++w) ch.panel[w] = pbuf[(size_t)w].data();
  ch.tmp = tmp.data();
  ch.vec[0] = v0.data(); ch.vec[1] = v1.data(); ch.vec[2] = v2.data(); ch.vec[3] = v3.data();
  ch.logdet_acc = &acc;
  *nloc_out = lay.nloc;
  *ld_out = lay.ld;
  int rc = ch.factor(info);
  if (rc) return rc;
  *logdet_local = acc;
  if (*info == 0) {
    rc = ch.solve(rhs, sol);
    if (rc) return rc;
    if (do_inverse) {
      rc = ch.x_sweep();
      if (rc) return rc;
      rc = ch.diag_binv(diag_binv);
      if (rc) return rc;
      rc = ch.y_sweep();
      if (rc) return rc;
    }
  }
  memcpy(Lslab, A.data(), sizeof(double) * slab);
  memcpy(Yslab, X.data(), sizeof(double) * slab);
  *ncollectives = ops.collectives;
  return 0;
}
