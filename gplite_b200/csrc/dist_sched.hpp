// Schedule of the multi-GPU dense factorisation (BASELINE config 5): B = I + S K S built,
// factorised, solved against and inverted across the ranks of one box, one process per GPU.
//
// This header is pure C++ (no CUDA, no NCCL): it decides WHO does WHAT on WHICH stream and in which
// order collectives are issued; every device operation, stream/event primitive and collective goes
// through the `Ops` policy class.  libgplite_b200.so instantiates it with the CUDA + NCCL back end
// (dist.cu); tests/sim/dist_sim.cpp instantiates it with a host back end whose collectives are
// callbacks into torch.distributed (gloo), so that the index arithmetic and the collective order are
// exercised on CPUs with world sizes 2 and 3.  The sim back end is test infrastructure, not a
// product path.
//
// Layout: 1-D block-cyclic by block COLUMNS of width nb (block b lives on rank b % P, stored as a
// column-major slab of full height).  On NVSwitch every peer is equally far, so the 2-D layout's only
// advantage (less panel traffic per link) buys nothing on 8 GPUs: panel traffic is n^2/2 * 8 bytes per
// rank in total (160 GB at n = 200k) against n^3/(3P) flops, a few per cent of the step time and
// hidden behind the look-ahead.  What the 1-D layout gives in exchange: the panel step (potrf of the
// diagonal block + TRSM of the rows below) is local to one rank, and the whole trailing update of a
// rank is ONE staircase GEMM per panel (all its remaining block columns at once, the panel rows of
// every P-th block read in place), instead of one launch per block column.
//
// Streams: MAIN carries the bulk of the trailing updates, PANEL (high priority) the critical chain, COMM
// the panel broadcasts.  Look-ahead of depth D with W = D + 1 panel buffers: when panel k arrives, every
// rank first applies it ON THE PANEL STREAM to the one block column it owns among the next D
// ("eager" update; the owner of block k+1 then factors it right away and its broadcast starts), while
// MAIN applies panel k to everything further away.  The chain panel -> eager update -> potrf -> TRSM ->
// broadcast therefore never queues behind a bulk update, it may run up to D panels ahead of the slowest
// rank's bulk work, and the +-1 block-column imbalance between ranks at any single step (5 % at n = 100k
// on 8 GPUs) averages out over the window instead of being paid at every step.  Nothing synchronises
// with the host inside a sweep: a failing pivot is collected in a device word and looked at once at the
// end.
//
// The tail.  Once the trailing matrix is small ((n - k) < 4 nb P rows) a step's bulk update no longer
// covers the panel chain, and what remains is a latency chain through ONE GPU per step: update of the
// next block column, potrf, TRSM, broadcast (measured at n = 100k on 8 GPUs: the last 38 of 98 steps took
// 196 ms for 77 ms of arithmetic).  There the panel is produced, broadcast and consumed in MAX_CHUNKS
// column chunks: while chunk c is on the wire, the owner already solves chunk c + 1 and the next owner
// already applies chunk c - 1 to its block column (a K = nb / 4 product), so the chain per step shrinks
// from (update + TRSM + broadcast + potrf) to about (TRSM + potrf + one chunk's update and broadcast).
//
// Replaces, for matrices that do not fit one GPU, base-R chol() / forwardsolve() / backsolve() at
// R/laplace.R:27-30 and R/ep.R:38-42, and the n x n solve behind R/ep.R:42,141.
#pragma once

#include <stdint.h>

#include <vector>

namespace gpl {
namespace dist {

enum Stream { S_MAIN = 0, S_PANEL = 1, S_COMM = 2, S_AUX = 3 };
// events come in families indexed by the panel buffer (k % W) they guard
constexpr int MAX_WINDOW = 8;
enum Event { EV_PANEL_READY0 = 0,                       // owner: packed panel is in the buffer
             EV_BCAST_DONE0 = MAX_WINDOW,               // the broadcast into the buffer has completed
             EV_BUF_FREE0 = 2 * MAX_WINDOW,             // MAIN no longer reads the buffer (bulk update done)
             EV_EAGER_DONE0 = 3 * MAX_WINDOW,           // PANEL no longer reads the buffer (eager update done)
             EV_COL_READY = 4 * MAX_WINDOW, EV_MAIN_DONE, EV_AUX_GO, EV_AUX_DONE,
             EV_CHUNK_READY0,                                       // + slot * MAX_CHUNKS + chunk
             EV_CHUNK_DONE0 = EV_CHUNK_READY0 + MAX_WINDOW * 4,
             EV_COUNT = EV_CHUNK_DONE0 + MAX_WINDOW * 4 };
constexpr int MAX_CHUNKS = 4;
inline Event ev(Event family, int64_t w) { return (Event)((int)family + (int)w); }
inline Event evc(Event family, int64_t w, int c) { return (Event)((int)family + (int)w * MAX_CHUNKS + c); }

struct Stair { int64_t nb, first, step, bskip; };   // same meaning as gpl::GemmStair

// Ownership / index maps of the 1-D block-cyclic column layout.
struct Layout {
  int64_t n = 0, nb = 0, ld = 0;
  int world = 1, rank = 0;
  int64_t nblk = 0;
  int64_t nloc = 0;                 // local columns
  std::vector<int64_t> lc0;         // local first column of block b (owned blocks only, else -1)
  void init(int64_t n_, int64_t nb_, int world_, int rank_) {
    n = n_; nb = nb_; world = world_; rank = rank_;
    ld = (n + 15) / 16 * 16;
    nblk = (n + nb - 1) / nb;
    lc0.assign((size_t)nblk, -1);
    nloc = 0;
    for (int64_t b = 0; b < nblk; ++b)
      if (owner(b) == rank) { lc0[(size_t)b] = nloc; nloc += width(b); }
  }
  int owner(int64_t b) const { return (int)(b % world); }
  int64_t col0(int64_t b) const { return b * nb; }
  int64_t width(int64_t b) const { return (n - b * nb) < nb ? (n - b * nb) : nb; }
  // first block >= b owned by this rank (may be >= nblk)
  int64_t first_owned_from(int64_t b) const {
    const int64_t r = ((rank - b) % world + world) % world;
    return b + r;
  }
  int64_t ldp(int64_t k) const { return (n - col0(k) + 15) / 16 * 16; }   // packed panel k
};

// Ops must provide (all asynchronous on the given Stream unless stated):
//   int fill_block(Stream, int64_t b, double* dst /* slab column lc0[b], row col0(b) */, int64_t ld)
//        -- rows col0(b) .. n of block column b of B (diagonal block included)
//   int potrf_diag(Stream, int64_t k, int64_t w, double* A, int64_t lda, int pivot0)   (no host sync)
//   int trsm_below(Stream, int64_t k, int64_t c0, int64_t wc, int64_t m, int64_t w, const double* L, int64_t ldl,
//                  double* B, int64_t ldb)
//        -- B (m x w) <- B L^-T, L = the w x w diagonal sub-block starting at column c0 of diagonal block k
//   int inv_diag(Stream, int64_t k, int64_t w, const double* L, int64_t ldl, double* out, int64_t ldo)
//   int gemm(Stream, bool ta, bool tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, bool lower, const Stair*)
//   int copy2d(Stream, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols)
//   int zero(Stream, double* p, int64_t count)
//   int trsv(Stream, int64_t k, bool trans, int64_t w, const double* L, int64_t ldl, double* b, double* x)
//   int gemv(Stream, bool trans, int64_t m, int64_t n, const double* A, int64_t lda, const double* x, double* y)
//   int axpy(Stream, int64_t n, double a, const double* x, double* y)            y += a x
//   int sub(Stream, int64_t n, const double* a, const double* b, double* out)    out = a - b
//   int colsumsq(Stream, int64_t m, int64_t n, const double* A, int64_t lda, double* out)
//   int diag_logsum(Stream, int64_t w, const double* A, int64_t lda, double* acc)   acc += 2 sum log diag
//   int record(Event, Stream); int wait(Stream, Event); int sync(Stream)
//   int bcast(Stream, double* buf, int64_t count, int root)
//   int allreduce_sum(Stream, double* buf, int64_t count)
//   int check_info(int* first_bad_pivot)   -- syncs, all-reduces (min) the failing pivot word
// Every method returns 0 on success; any other value aborts the sweep and is returned unchanged.

#define GPL_DS(call)            \
  do {                          \
    int s__ = (call);           \
    if (s__ != 0) return s__;   \
  } while (0)

template <class Ops>
class ShardedChol {
 public:
  Layout lay;
  Ops* ops = nullptr;
  double* A = nullptr;        // slab: nloc columns, ld rows: B, then its factor L
  double* X = nullptr;        // second slab (optional): L^-1, then B^-1
  double* panel[MAX_WINDOW] = {};          // W packed-panel buffers, nb * ld doubles each
  int window = 2;                          // W; the look-ahead depth is D = min(W - 1, world)
  double* tmp = nullptr;      // nb x max(nloc, nb) scratch for the inverse sweeps
  double* vec[4] = {nullptr, nullptr, nullptr, nullptr};   // n-vectors (solve workspace)
  double* logdet_acc = nullptr;   // device scalar
  double bcast_bytes = 0.0;       // panel bytes this rank has put through the collectives

  double* slab(double* S, int64_t b) const { return S + lay.lc0[(size_t)b] * lay.ld; }
  bool chunked_tail = true;       // GPL_DIST_TAIL=0 switches the chunked tail off (tuning / diagnostics)
  int chunk_align = 128;          // a chunk must be a whole number of the back end's diagonal sub-blocks
  // panel k travels in MAX_CHUNKS column chunks (tail of the factorisation)?
  int chunks_of(int64_t k) const {
    const Layout& L = lay;
    if (!chunked_tail || L.world < 2) return 1;
    if (L.width(k) != L.nb || L.nb % (chunk_align * MAX_CHUNKS) != 0) return 1;
    return (L.n - L.col0(k)) < 4 * L.nb * (int64_t)L.world ? MAX_CHUNKS : 1;
  }

  // ---- build + factorise ---------------------------------------------------------------------
  // B's owned block columns come from ops->fill_block; returns the first failing pivot (0 = ok) in
  // *info; logdet(B) accumulates into the device scalar logdet_acc (summed over ranks by the caller
  // through finish_logdet).
  int factor(int* info) {
    const Layout& L = lay;
    const int me = L.rank;
    const int W = window < 2 ? 2 : (window > MAX_WINDOW ? MAX_WINDOW : window);
    const int64_t D = (W - 1) < L.world ? (W - 1) : L.world;         // at most one own block per window
    GPL_DS(ops->zero(S_MAIN, logdet_acc, 1));
    // the first panel's block column is built first so that its factorisation overlaps the rest of the build
    if (L.owner(0) == me) {
      GPL_DS(ops->fill_block(S_MAIN, 0, slab(A, 0), L.ld));
      GPL_DS(ops->record(EV_COL_READY, S_MAIN));
      GPL_DS(ops->wait(S_PANEL, EV_COL_READY));
    }
    for (int w = 0; w < W; ++w) {                                      // every panel buffer starts free
      GPL_DS(ops->record(ev(EV_BUF_FREE0, w), S_MAIN));
      GPL_DS(ops->record(ev(EV_EAGER_DONE0, w), S_MAIN));
    }
    if (L.owner(0) == me) GPL_DS(factor_panel(0, W));
    GPL_DS(send_panel(0, W));
    for (int64_t b = L.first_owned_from(L.owner(0) == me ? 1 : 0); b < L.nblk; b += L.world)
      GPL_DS(ops->fill_block(S_MAIN, b, slab(A, b) + L.col0(b), L.ld));
    GPL_DS(ops->record(EV_MAIN_DONE, S_MAIN));                        // all own blocks exist
    GPL_DS(ops->wait(S_PANEL, EV_MAIN_DONE));
    for (int64_t k = 0; k < L.nblk; ++k) {
      const int64_t w = k % W;
      // ---- eager update (PANEL): this rank's block column among the next D
      const int64_t b = L.first_owned_from(k + 1);
      // (the next owner consumes a chunked panel chunk by chunk; everyone else waits for all of it)
      if (!(b == k + 1 && b < L.nblk && chunks_of(k) > 1)) GPL_DS(ops->wait(S_PANEL, ev(EV_BCAST_DONE0, w)));
      if (b < L.nblk && b <= k + D) {
        // its first eager update (from panel b - D) must follow the bulk updates by all earlier panels
        if (b == k + D && k > 0) GPL_DS(ops->wait(S_PANEL, ev(EV_BUF_FREE0, (k - 1) % W)));
        if (b == k + 1) {
          // the next panel: only its diagonal block is needed before potrf can start, so the rows below
          // are updated on a second high-priority stream while the diagonal block is being factored
          GPL_DS(ops->record(EV_AUX_GO, S_PANEL));
          GPL_DS(ops->wait(S_AUX, EV_AUX_GO));
          const int C = chunks_of(k);
          const int64_t wc = L.width(k) / C;
          for (int c = 0; c < C; ++c) {                              // (C > 1: as the chunks of panel k arrive)
            if (C > 1) {
              GPL_DS(ops->wait(S_AUX, evc(EV_CHUNK_DONE0, w, c)));
              GPL_DS(ops->wait(S_PANEL, evc(EV_CHUNK_DONE0, w, c)));
            }
            GPL_DS(apply_panel_rows(S_AUX, k, b, L.width(b), L.n - L.col0(b) - L.width(b), W, c * wc, wc));
            GPL_DS(apply_panel_rows(S_PANEL, k, b, 0, L.width(b), W, c * wc, wc));
          }
          GPL_DS(ops->record(EV_AUX_DONE, S_AUX));
          GPL_DS(factor_panel(b, W, true));
        } else {
          GPL_DS(apply_panel_to(S_PANEL, k, b, 1, W));
        }
      }
      GPL_DS(ops->record(ev(EV_EAGER_DONE0, w), S_PANEL));
      if (k + 1 < L.nblk) GPL_DS(send_panel(k + 1, W));
      // ---- bulk update (MAIN): everything beyond the window, one staircase product
      GPL_DS(ops->wait(S_MAIN, ev(EV_BCAST_DONE0, w)));
      const int64_t j = L.first_owned_from(k + D + 1);
      if (j < L.nblk) {
        const int64_t cnt = (L.nblk - 1 - j) / L.world + 1;
        GPL_DS(apply_panel_to(S_MAIN, k, j, cnt, W));
      }
      GPL_DS(ops->record(ev(EV_BUF_FREE0, w), S_MAIN));
    }
    GPL_DS(ops->record(EV_COL_READY, S_PANEL));                       // join the side streams
    GPL_DS(ops->wait(S_MAIN, EV_COL_READY));
    GPL_DS(ops->record(EV_MAIN_DONE, S_COMM));
    GPL_DS(ops->wait(S_MAIN, EV_MAIN_DONE));
    return ops->check_info(info);
  }

  // ---- triangular solves: out = B^-1 rhs (replicated in, replicated out) -------------------------
  // Block forward / backward substitution; per block one nb-sized all-reduce (forward: the partial
  // sums sum_j L_kj w_j live with the owners of the columns j) or broadcast (backward).
  int solve(const double* rhs, double* out) {
    const Layout& L = lay;
    const int me = L.rank;
    const int64_t n = L.n;
    double *u = vec[0], *wv = vec[1], *t1 = vec[2], *t2 = vec[3];
    GPL_DS(ops->zero(S_MAIN, u, n));
    GPL_DS(ops->zero(S_MAIN, wv, n));
    for (int64_t k = 0; k < L.nblk; ++k) {
      const int64_t k0 = L.col0(k), w = L.width(k), below = n - k0 - w;
      GPL_DS(ops->allreduce_sum(S_MAIN, u + k0, w));
      if (L.owner(k) == me) {
        const double* Lk = slab(A, k) + k0;
        GPL_DS(ops->sub(S_MAIN, w, rhs + k0, u + k0, t1));
        GPL_DS(ops->trsv(S_MAIN, k, false, w, Lk, L.ld, t1, wv + k0));
        if (below > 0) {
          GPL_DS(ops->gemv(S_MAIN, false, below, w, Lk + w, L.ld, wv + k0, t2));
          GPL_DS(ops->axpy(S_MAIN, below, 1.0, t2, u + k0 + w));
        }
      } else if (below > 0) {
        // u[k0 + w ..] of a non-owner holds only its own partial sums: nothing to do
      }
      // the entries u[k0 .. k0 + w) were summed in place on every rank; later blocks are untouched
    }
    GPL_DS(ops->zero(S_MAIN, out, n));
    for (int64_t k = L.nblk - 1; k >= 0; --k) {
      const int64_t k0 = L.col0(k), w = L.width(k), below = n - k0 - w;
      if (L.owner(k) == me) {
        const double* Lk = slab(A, k) + k0;
        if (below > 0) {
          GPL_DS(ops->gemv(S_MAIN, true, below, w, Lk + w, L.ld, out + k0 + w, t2));
          GPL_DS(ops->sub(S_MAIN, w, wv + k0, t2, t1));
        } else {
          GPL_DS(ops->sub(S_MAIN, w, wv + k0, out + k0, t1));      // out[k0..] is still zero
        }
        GPL_DS(ops->trsv(S_MAIN, k, true, w, Lk, L.ld, t1, out + k0));
      }
      GPL_DS(ops->bcast(S_MAIN, out + k0, w, L.owner(k)));
    }
    return 0;
  }

  // ---- inverse: B^-1 = X^T X, X = L^-1 (SURVEY 8e "C^-1 + trace term") ----------------------------
  // Two more sweeps of panel broadcasts, each n^3/3 flops spread like the factorisation.
  //   X sweep: panel k = [L_kk^-1 ; L[k1:, k]] goes to every rank; for ALL its columns j < k at once a
  //            rank finishes block row k of X (X_kJ = L_kk^-1 R_kJ) and pushes the update down
  //            (R[k1:, J] -= L[k1:, k] X_kJ); the running right-hand side R lives in X's slab.
  //   Y sweep: panel i = X[i0:, i]; Y_iJ = X[i0:, i]^T X[i0:, J] for its columns j <= i, written over
  //            the block row of X that is no longer needed.
  // Panels depend only on finished data, so panel k+1 is always in flight during step k.
  int x_sweep() {
    const Layout& L = lay;
    GPL_DS(ops->zero(S_MAIN, X, L.nloc * L.ld));
    return sweep(true);
  }
  int y_sweep() { return sweep(false); }

  // diag(B^-1)_i = sum_k X_ki^2 after x_sweep: column sums of squares of the owned block columns into
  // out (n, zero elsewhere), then one all-reduce
  int diag_binv(double* out) {
    const Layout& L = lay;
    GPL_DS(ops->zero(S_MAIN, out, L.n));
    for (int64_t j = L.first_owned_from(0); j < L.nblk; j += L.world) {
      const int64_t j0 = L.col0(j);
      GPL_DS(ops->colsumsq(S_MAIN, L.n - j0, L.width(j), slab(X, j) + j0, L.ld, out + j0));
    }
    return ops->allreduce_sum(S_MAIN, out, L.n);
  }

 private:
  // owner only, on PANEL: potrf of the diagonal block, TRSM of the rows below, pack for broadcast
  int factor_panel(int64_t k, int W, bool rows_on_aux = false) {
    const Layout& L = lay;
    const int64_t k0 = L.col0(k), w = L.width(k), below = L.n - k0 - w, pb = k % W;
    double* Akk = slab(A, k) + k0;
    GPL_DS(ops->potrf_diag(S_PANEL, k, w, Akk, L.ld, (int)k0));
    GPL_DS(ops->diag_logsum(S_PANEL, w, Akk, L.ld, logdet_acc));
    if (rows_on_aux) GPL_DS(ops->wait(S_PANEL, EV_AUX_DONE));       // the rows below have their update
    GPL_DS(ops->wait(S_PANEL, ev(EV_BUF_FREE0, pb)));               // (its own eager use is earlier in PANEL order)
    const int C = chunks_of(k);
    if (C == 1) {
      if (below > 0) GPL_DS(ops->trsm_below(S_PANEL, k, 0, w, below, w, Akk, L.ld, Akk + w, L.ld));
      GPL_DS(ops->copy2d(S_PANEL, panel[pb], L.ldp(k), Akk, L.ld, L.n - k0, w));
      return ops->record(ev(EV_PANEL_READY0, pb), S_PANEL);
    }
    // tail: column chunk c of the rows below is solved (after the update by the chunks to its left),
    // packed and released to the broadcast while chunk c + 1 is being solved
    const int64_t wc = w / C, ldp = L.ldp(k);
    double* B = Akk + w;                                            // rows below the diagonal block
    for (int c = 0; c < C; ++c) {
      const int64_t c0 = c * wc;
      if (below > 0) {
        if (c > 0)
          GPL_DS(ops->gemm(S_PANEL, false, true, below, wc, c0, -1.0, B, L.ld, Akk + c0, L.ld, 1.0, B + c0 * L.ld, L.ld,
                           false, nullptr));
        GPL_DS(ops->trsm_below(S_PANEL, k, c0, wc, below, wc, Akk + c0 * L.ld + c0, L.ld, B + c0 * L.ld, L.ld));
      }
      GPL_DS(ops->copy2d(S_PANEL, panel[pb] + c0 * ldp, ldp, Akk + c0 * L.ld, L.ld, L.n - k0, wc));
      GPL_DS(ops->record(evc(EV_CHUNK_READY0, pb, c), S_PANEL));
    }
    return ops->record(ev(EV_PANEL_READY0, pb), S_PANEL);
  }

  // every rank, on COMM: broadcast of packed panel k from its owner into buffer k % W
  int send_panel(int64_t k, int W) {
    const Layout& L = lay;
    const int64_t pb = k % W;
    const int C = chunks_of(k);
    const bool own = L.owner(k) == L.rank;
    if (!own) {
      GPL_DS(ops->wait(S_COMM, ev(EV_BUF_FREE0, pb)));
      GPL_DS(ops->wait(S_COMM, ev(EV_EAGER_DONE0, pb)));
    } else if (C == 1) {
      GPL_DS(ops->wait(S_COMM, ev(EV_PANEL_READY0, pb)));
    }
    const int64_t count = L.ldp(k) * L.width(k) / C;
    for (int c = 0; c < C; ++c) {
      if (own && C > 1) GPL_DS(ops->wait(S_COMM, evc(EV_CHUNK_READY0, pb, c)));
      GPL_DS(ops->bcast(S_COMM, panel[pb] + c * count, count, L.owner(k)));
      bcast_bytes += 8.0 * (double)count;
      if (C > 1) GPL_DS(ops->record(evc(EV_CHUNK_DONE0, pb, c), S_COMM));
    }
    return ops->record(ev(EV_BCAST_DONE0, pb), S_COMM);
  }

  // rows [r0, r0 + m) (relative to the block's diagonal) of block column j -= P_k[those rows] P_k[block j rows]^T
  // (only the panel columns [kc0, kc0 + kw) when kw > 0: one chunk of a chunked panel)
  int apply_panel_rows(Stream S, int64_t k, int64_t j, int64_t r0, int64_t m, int W, int64_t kc0 = 0,
                       int64_t kw = 0) {
    const Layout& L = lay;
    if (m <= 0) return 0;
    const int64_t k0 = L.col0(k), j0 = L.col0(j), ldp = L.ldp(k);
    if (kw <= 0) kw = L.width(k);
    const double* P = panel[k % W] + kc0 * ldp;
    return ops->gemm(S, false, true, m, L.width(j), kw, -1.0, P + (j0 - k0) + r0, ldp, P + (j0 - k0), ldp, 1.0,
                     slab(A, j) + j0 + r0, L.ld, false, nullptr);
  }

  // A[j0.., blocks j, j + P, ... (cnt of them)] -= P_k[rows] P_k[block rows]^T
  int apply_panel_to(Stream S, int64_t k, int64_t j, int64_t cnt, int W) {
    const Layout& L = lay;
    const int64_t k0 = L.col0(k), w = L.width(k), j0 = L.col0(j);
    const int64_t jl = j + (cnt - 1) * L.world;                    // last block of the group
    const int64_t ncols = (cnt - 1) * L.nb + L.width(jl);
    const double* P = panel[k % W];
    const int64_t ldp = L.ldp(k);
    Stair st{L.nb, 0, (int64_t)L.world * L.nb, (int64_t)(L.world - 1) * L.nb};
    return ops->gemm(S, false, true, L.n - j0, ncols, w, -1.0, P + (j0 - k0), ldp, P + (j0 - k0), ldp,
                     1.0, slab(A, j) + j0, L.ld, false, cnt > 1 ? &st : nullptr);
  }

  // shared skeleton of the two inverse sweeps: prepare (owner, PANEL) -> broadcast (COMM) -> apply (MAIN)
  int sweep(bool xs) {
    const Layout& L = lay;
    const int me = L.rank;
    GPL_DS(ops->record(EV_COL_READY, S_MAIN));
    GPL_DS(ops->wait(S_PANEL, EV_COL_READY));
    for (int w = 0; w < 2; ++w) {
      GPL_DS(ops->record(ev(EV_BUF_FREE0, w), S_MAIN));
      GPL_DS(ops->record(ev(EV_EAGER_DONE0, w), S_MAIN));
    }
    if (L.owner(0) == me) GPL_DS(xs ? prep_L_panel(0) : prep_X_panel(0));
    GPL_DS(send_panel(0, 2));
    for (int64_t k = 0; k < L.nblk; ++k) {
      GPL_DS(ops->wait(S_MAIN, ev(EV_BCAST_DONE0, k & 1)));
      if (k + 1 < L.nblk) {
        if (L.owner(k + 1) == me) GPL_DS(xs ? prep_L_panel(k + 1) : prep_X_panel(k + 1));
        GPL_DS(send_panel(k + 1, 2));
      }
      GPL_DS(xs ? apply_L_panel(k) : apply_X_panel(k));
      GPL_DS(ops->record(ev(EV_BUF_FREE0, k & 1), S_MAIN));
    }
    GPL_DS(ops->record(EV_COL_READY, S_PANEL));
    GPL_DS(ops->wait(S_MAIN, EV_COL_READY));
    GPL_DS(ops->record(EV_MAIN_DONE, S_COMM));
    return ops->wait(S_MAIN, EV_MAIN_DONE);
  }

  int prep_L_panel(int64_t k) {
    const Layout& L = lay;
    const int64_t k0 = L.col0(k), w = L.width(k), below = L.n - k0 - w, ldp = L.ldp(k);
    const double* Akk = slab(A, k) + k0;
    double* pv = panel[k & 1];
    GPL_DS(ops->wait(S_PANEL, ev(EV_BUF_FREE0, k & 1)));
    GPL_DS(ops->copy2d(S_PANEL, pv, ldp, Akk, L.ld, L.n - k0, w));
    GPL_DS(ops->inv_diag(S_PANEL, k, w, Akk, L.ld, pv, ldp));        // diagonal block <- L_kk^-1
    // this rank's own column k of X:  X_kk = L_kk^-1,  R[k1:, k] = -L[k1:, k] L_kk^-1
    double* Xk = slab(X, k) + k0;
    GPL_DS(ops->copy2d(S_PANEL, Xk, L.ld, pv, ldp, w, w));
    if (below > 0)
      GPL_DS(ops->gemm(S_PANEL, false, false, below, w, w, -1.0, pv + w, ldp, pv, ldp, 0.0, Xk + w, L.ld,
                       false, nullptr));
    return ops->record(ev(EV_PANEL_READY0, k & 1), S_PANEL);
  }

  int apply_L_panel(int64_t k) {
    const Layout& L = lay;
    const int64_t k0 = L.col0(k), w = L.width(k), below = L.n - k0 - w, ldp = L.ldp(k);
    // local columns of the blocks j < k owned by this rank: a contiguous prefix of the slab
    const int64_t jn = L.first_owned_from(k);                       // first owned block >= k
    const int64_t nc = jn < L.nblk ? L.lc0[(size_t)jn] : L.nloc;
    if (nc == 0) return 0;
    const double* pv = panel[k & 1];
    double* Xr = X + k0;                                            // block row k of the slab
    GPL_DS(ops->gemm(S_MAIN, false, false, w, nc, w, 1.0, pv, ldp, Xr, L.ld, 0.0, tmp, w, false, nullptr));
    if (below > 0)
      GPL_DS(ops->gemm(S_MAIN, false, false, below, nc, w, -1.0, pv + w, ldp, tmp, w, 1.0, Xr + w, L.ld, false,
                       nullptr));
    return ops->copy2d(S_MAIN, Xr, L.ld, tmp, w, w, nc);
  }

  int prep_X_panel(int64_t i) {
    const Layout& L = lay;
    const int64_t i0 = L.col0(i), w = L.width(i);
    GPL_DS(ops->wait(S_PANEL, ev(EV_BUF_FREE0, i & 1)));
    GPL_DS(ops->copy2d(S_PANEL, panel[i & 1], L.ldp(i), slab(X, i) + i0, L.ld, L.n - i0, w));
    return ops->record(ev(EV_PANEL_READY0, i & 1), S_PANEL);
  }

  int apply_X_panel(int64_t i) {
    const Layout& L = lay;
    const int64_t i0 = L.col0(i), w = L.width(i), ldp = L.ldp(i);
    // columns of the blocks j <= i owned by this rank
    const int64_t jn = L.first_owned_from(i + 1);
    const int64_t nc = jn < L.nblk ? L.lc0[(size_t)jn] : L.nloc;
    if (nc == 0) return 0;
    double* Xr = X + i0;
    GPL_DS(ops->gemm(S_MAIN, true, false, w, nc, L.n - i0, 1.0, panel[i & 1], ldp, Xr, L.ld, 0.0, tmp, w, false,
                     nullptr));
    return ops->copy2d(S_MAIN, Xr, L.ld, tmp, w, w, nc);
  }
};

}  // namespace dist
}  // namespace gpl
