"""Multi-GPU full GP (BASELINE config 5: n = 200k, lik_gaussian, method_full): K is built and
factorised across the GPUs of one box, one process per GPU.

Layout: 1-D block-cyclic by block COLUMNS (block b of width nb lives on rank b % P, stored as a
local column-major slab of full height).  On NVSwitch every peer is equally far, so a 1-D layout
costs nothing in link contention and makes the panel step entirely local to one rank:

  step k:  owner(k): potrf(diag block) + TRSM of the rows below        (DMMA, local)
           broadcast the (n - k0) x nb panel to all ranks                (NCCL over NVLink)
           every rank: A[:, j] -= P[j-rows] P[j-block]^T for its own block columns j > k  (DMMA GEMM)

Panel traffic is n^2/2 * 8 bytes in total (160 GB at n = 200k, received by every rank) against
n^3/3 / P flops per rank: ~2-3 % of the step time at P = 8 even without overlap.  The triangular
solves for alpha run block by block with an 8 KB all-reduce / broadcast per block.

All n x n arithmetic goes through the C ABI (`ops`); torch.distributed is only the transport.
`TorchOps`-style stand-ins used by the CPU (gloo) tests live in tests/, not here: this module has
no CPU compute path.

Reference semantics: fit_laplace.method_full (R/gp_fit.R:91-99) + laplace.method_full /
laplace_iter.method_full (R/laplace.R:13-38, 274-297) for every likelihood (the Newton loop rebuilds
B = I + V^-1/2 K V^-1/2 from x each iteration -- the K build is ~1 % of a factorisation -- and the
site vectors f, z, V are replicated: 8n bytes), with f = K alpha evaluated as z - V alpha (the two
are identical because (K + V) alpha = z).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from . import model as M


class BlockCyclic:
    """Ownership / index maps of the 1-D block-cyclic column layout."""

    def __init__(self, n, nb, world, rank):
        self.n, self.nb, self.world, self.rank = n, nb, world, rank
        self.nblk = (n + nb - 1) // nb
        self.owned = [b for b in range(self.nblk) if b % world == rank]
        self.width = [min(nb, n - b * nb) for b in range(self.nblk)]
        self.local_col0 = {}
        c = 0
        for b in self.owned:
            self.local_col0[b] = c
            c += self.width[b]
        self.nloc = c

    def owner(self, b):
        return b % self.world

    def col0(self, b):
        return b * self.nb


class GpuOps:
    """The C-ABI calls used by the distributed driver, on torch CUDA tensors (column-major slabs:
    tensor[l, i] is element (row i, local column l))."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.lib = ctx.lib

    @staticmethod
    def _p(t, offset_elems=0):
        return C.c_void_p(t.data_ptr() + 8 * offset_elems)

    def kblock(self, pm, X_t, n, d, r0, nrows, c0, ncols, A_t, lc0):
        ld = A_t.shape[1]
        _lib.check(self.lib.gpl_cf_eval_block_dev(self.ctx.h, pm.ref(), self._p(X_t, r0), nrows, n,
                                                  self._p(X_t, c0), ncols, n, d,
                                                  self._p(A_t, lc0 * ld + r0), ld))

    def scale_block(self, A_t, lc0, ncols, r0, nrows, alpha, diag_add, diag_offset, s_t=None, c0=0):
        """A[r0:r0+nrows, lc0:lc0+ncols] <- alpha s_i A_ij s_j (+ diag_add on the block's diagonal);
        s_t: replicated n-vector (rows r0.., global columns c0..) or None"""
        ld = A_t.shape[1]
        r = self._p(s_t, r0) if s_t is not None else None
        c = self._p(s_t, c0) if s_t is not None else None
        _lib.check(self.lib.gpl_scale_block_dev(self.ctx.h, nrows, ncols, self._p(A_t, lc0 * ld + r0), ld,
                                                r, c, float(alpha), float(diag_add), diag_offset))

    def gemm(self, ta, tb, m, n, k, alpha, A, B, beta, Cm, lower=0):
        """C <- alpha op(A) op(B) + beta C on column-major views given as (tensor, element offset, ld)"""
        (At, ao, lda), (Bt, bo, ldb), (Ct, co, ldc) = A, B, Cm
        _lib.check(self.lib.gpl_dgemm_dev(self.ctx.h, ta, tb, m, n, k, float(alpha), self._p(At, ao), lda,
                                          self._p(Bt, bo), ldb, float(beta), self._p(Ct, co), ldc, lower))

    def inv_diag(self, A_t, lc0, k0, w, dinv_t, out_t, ldo):
        """out (w x w, ld ldo) <- L_kk^-1 for the diagonal block of block column lc0 (exact zeros above
        the diagonal): right-solve X L_kk = I with that block's saved 128-block inverses"""
        import torch
        ld = A_t.shape[1]
        blk = torch.as_strided(out_t, (w, w), (ldo, 1))
        blk.zero_()
        blk.diagonal().fill_(1.0)
        _lib.check(self.lib.gpl_dtrsm_right_dinv_dev(self.ctx.h, 0, w, w, self._p(A_t, lc0 * ld + k0), ld,
                                                     self._p(dinv_t), self._p(out_t), ldo))

    def grad_trace_block(self, pm, X_t, n, d, c0, w, Y_t, lc0, s_t, alpha_t, pf, ps):
        """block column's share of sum W_ij dK_ij/dlog(theta_p) and of sum_j s_j^2 Binv_jj"""
        ld = Y_t.shape[1]
        out = np.zeros(len(pf), dtype=np.float64)
        dsum = C.c_double()
        _lib.check(self.lib.gpl_grad_trace_block_dev(
            self.ctx.h, pm.ref(), self._p(X_t), n, d, c0, w, self._p(Y_t, lc0 * ld), ld, self._p(s_t),
            self._p(alpha_t), len(pf), _lib._ptr(pf), _lib._ptr(ps), _lib._ptr(out), C.byref(dsum)))
        return out, dsum.value

    def colsumsq(self, A_t, lc0, r0, rows, ncols, out_t, out_off):
        """out[out_off + j] = sum_i A[r0 + i, lc0 + j]^2"""
        ld = A_t.shape[1]
        _lib.check(self.lib.gpl_colsumsq_dev(self.ctx.h, rows, ncols, self._p(A_t, lc0 * ld + r0), ld,
                                             self._p(out_t, out_off)))

    def kdiag(self, pm, X_t, n, d, jitter, out_t):
        """diag(K) + jitter"""
        _lib.check(self.lib.gpl_cf_eval_dev(self.ctx.h, pm.ref(), self._p(X_t), n, self._p(X_t), n, d, 1, 0, 1,
                                            float(jitter), self._p(out_t), n))

    def ep_sites(self, lik, fmean, fvar, z_old, P_old, y_t, tr_t, off_t, damping, xq_t, wq_t,
                 z_t, V_t, P_t, cm_t, cv_t, logZ_t):
        """parallel EP site update on replicated device n-vectors; returns the flag word"""
        flags = C.c_int32()
        _lib.check(self.lib.gpl_lik_ep_sites_dev(
            self.ctx.h, C.byref(lik), fmean.numel(), self._p(fmean), self._p(fvar), self._p(z_old), self._p(P_old),
            self._p(y_t), self._p(tr_t) if tr_t is not None else None,
            self._p(off_t) if off_t is not None else None, float(damping), xq_t.numel(), self._p(xq_t),
            self._p(wq_t), self._p(z_t), self._p(V_t), self._p(P_t), self._p(cm_t), self._p(cv_t),
            self._p(logZ_t), C.byref(flags)))
        return flags.value

    def pseudodata(self, lik, f_t, y_t, tr_t, off_t, z_t, V_t):
        """z, V <- get_pseudodata_la(lik, f, y) on replicated device n-vectors (R/lik.R:188-201)"""
        _lib.check(self.lib.gpl_lik_pseudodata_la_dev(
            self.ctx.h, C.byref(lik), f_t.numel(), self._p(f_t), self._p(y_t),
            self._p(tr_t) if tr_t is not None else None, self._p(off_t) if off_t is not None else None,
            self._p(z_t), self._p(V_t)))

    def loglik(self, lik, f_t, y_t, tr_t, off_t):
        out = C.c_double()
        _lib.check(self.lib.gpl_lik_loglik_dev(
            self.ctx.h, C.byref(lik), f_t.numel(), self._p(f_t), self._p(y_t),
            self._p(tr_t) if tr_t is not None else None, self._p(off_t) if off_t is not None else None,
            C.byref(out)))
        return out.value

    def potrf(self, A_t, lc0, r0, nbk):
        ld = A_t.shape[1]
        info = C.c_int32()
        logdet = C.c_double()
        st = self.lib.gpl_dpotrf_dev(self.ctx.h, nbk, self._p(A_t, lc0 * ld + r0), ld, C.byref(info),
                                     C.byref(logdet))
        return st, info.value, logdet.value

    def save_dinv(self, nbk, out_t):
        _lib.check(self.lib.gpl_dpotrf_get_dinv_dev(self.ctx.h, nbk, self._p(out_t)))

    def trsm_below(self, A_t, lc0, r0, nbk, m):
        """rows r0+nbk .. r0+nbk+m of the block column <- (those rows) L^-T, L = diag block"""
        ld = A_t.shape[1]
        _lib.check(self.lib.gpl_dtrsm_right_dev(self.ctx.h, 1, m, nbk, self._p(A_t, lc0 * ld + r0), ld,
                                                self._p(A_t, lc0 * ld + r0 + nbk), ld))

    def update(self, P_t, ldp, nbk, off_rows, m, off_cols, ncols, A_t, lc0, r0):
        """A[r0:r0+m, lc0:lc0+ncols] -= P[off_rows:off_rows+m, :] P[off_cols:off_cols+ncols, :]^T"""
        ld = A_t.shape[1]
        _lib.check(self.lib.gpl_dgemm_dev(self.ctx.h, 0, 1, m, ncols, nbk, -1.0, self._p(P_t, off_rows), ldp,
                                          self._p(P_t, off_cols), ldp, 1.0, self._p(A_t, lc0 * ld + r0), ld, 0))

    def trsv(self, trans, nbk, A_t, lc0, r0, dinv_t, b_t, x_t):
        ld = A_t.shape[1]
        _lib.check(self.lib.gpl_dtrsv_dev(self.ctx.h, trans, nbk, self._p(A_t, lc0 * ld + r0), ld,
                                          self._p(dinv_t), self._p(b_t), self._p(x_t)))

    def gemv(self, trans, m, ncols, A_t, lc0, r0, x_t, x_off, y_t, y_off):
        ld = A_t.shape[1]
        _lib.check(self.lib.gpl_dgemv_dev(self.ctx.h, trans, m, ncols, self._p(A_t, lc0 * ld + r0), ld,
                                          self._p(x_t, x_off), self._p(y_t, y_off)))

    def sync(self):
        self.ctx.sync()


class ShardedFactor:
    """B = I + V^-1/2 (K + jitter I) V^-1/2 built, factorised and (optionally) inverted in the 1-D
    block-cyclic column layout.  All n x n arithmetic goes through `ops` (the C ABI); `dist` moves
    panels; replicated n-vectors are torch tensors."""

    def __init__(self, gp, x, nb, ops, dist, device, jitter, timings=None):
        import torch
        self.torch = torch
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[:, None]
        self.n, self.d = n, d = x.shape
        if nb % 128 or n % 2:
            raise ValueError("nb must be a multiple of 128 and n even")
        for cf in gp.cfs:
            M.learn_scales(cf, x)
        self.jit = 1e-6 if jitter is None else float(jitter)
        factors, _ = M.flatten_cfs(gp.cfs)
        self.pm = _lib.PackedModel(factors)
        if ops is None:
            # same stream as torch: kernels, copies and NCCL collectives are then ordered without
            # host synchronisation
            # (torch's default stream has handle 0, which the C ABI reads as "create your own": pass
            # cudaStreamLegacy = 1 for it)
            ops = GpuOps(_lib.Context(torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream or 1))
        self.ops, self.device, self.nb = ops, device, nb
        self.lay = lay = BlockCyclic(n, nb, self.world, self.rank)
        self.ld = ld = (n + 15) // 16 * 16
        self.X_t = torch.from_numpy(np.ascontiguousarray(x.T)).to(device)     # (d, n): column-major n x d
        self.A_t = torch.empty((max(lay.nloc, 1), ld), dtype=torch.float64, device=device)
        self.Xi_t = None
        self.dinv = {b: torch.empty(((lay.width[b] + 127) // 128) * 128 * 128, dtype=torch.float64, device=device)
                     for b in lay.owned}
        self.panels = [torch.empty(nb * ld, dtype=torch.float64, device=device) for _ in range(2)]
        self.tmpb = None
        self.timings = timings
        self.t_acc = {}
        self._t0 = None

    # ---- timing -------------------------------------------------------------------------------
    def tick(self):
        if self.timings is not None:
            import time
            self.ops.sync()
            if self.device == "cuda":
                self.torch.cuda.synchronize()
            self._t0 = time.perf_counter()

    def tock(self, name):
        if self.timings is not None:
            import time
            self.ops.sync()
            if self.device == "cuda":
                self.torch.cuda.synchronize()
            t = time.perf_counter()
            self.t_acc[name] = self.t_acc.get(name, 0.0) + t - self._t0
            self._t0 = t

    def vec(self, v):
        return None if v is None else self.torch.from_numpy(
            np.asarray(v, dtype=np.float64).ravel().copy()).to(self.device)

    def panel_view(self, k):
        w, mp = self.lay.width[k], self.n - self.lay.col0(k)
        return self.panels[k % 2][: w * mp].view(w, mp)

    def _sweep(self, prep, apply):
        """panel k+1 is prepared by its owner and broadcast (async) while panel k is applied"""
        lay, dist = self.lay, self.dist
        if lay.owner(0) == self.rank:
            prep(0)
        work = dist.broadcast(self.panel_view(0), src=lay.owner(0), async_op=True)
        for k in range(lay.nblk):
            if work is not None:
                work.wait()
            work = None
            if k + 1 < lay.nblk:
                if lay.owner(k + 1) == self.rank:
                    prep(k + 1)
                work = dist.broadcast(self.panel_view(k + 1), src=lay.owner(k + 1), async_op=True)
            apply(k)

    # ---- build + factorise --------------------------------------------------------------------
    def factor(self, s_t):
        """B(s) = I + S (K + jitter I) S, S = diag(s_t); returns logdet(B).  Right-looking with panel
        broadcasts and one step of look-ahead: the owner of panel k+1 updates that block column
        first, factors it and starts its broadcast (async, NCCL stream) while every rank is still
        applying panel k to its other columns, so the panel factorisation and the NVLink transfer
        leave the critical path."""
        torch, lay, ops, n, dist = self.torch, self.lay, self.ops, self.n, self.dist
        A_t = self.A_t
        self.tick()
        # K (owned block columns, rows at and below the diagonal block), scaled in place
        for b in lay.owned:
            c0, w, lc0 = lay.col0(b), lay.width[b], lay.local_col0[b]
            ops.kblock(self.pm, self.X_t, n, self.d, c0, n - c0, c0, w, A_t, lc0)
            ops.scale_block(A_t, lc0, w, c0, n - c0, 1.0, 0.0, 0, s_t, c0)
        diag_add = 1.0 + self.jit * s_t * s_t                              # replicated n-vector
        for b in lay.owned:
            c0, w, lc0 = lay.col0(b), lay.width[b], lay.local_col0[b]
            A_t[lc0:lc0 + w, c0:c0 + w].diagonal().add_(diag_add[c0:c0 + w])
        self.tock("kbuild")
        state = {"logdet": 0.0, "fail": 0}

        def factor_panel(k):
            """owner only: potrf of the diagonal block, TRSM of the rows below, pack for broadcast"""
            k0, w, lc0 = lay.col0(k), lay.width[k], lay.local_col0[k]
            st, info, ld_k = ops.potrf(A_t, lc0, k0, w)
            if st != _lib.GPL_OK:
                state["fail"] = max(state["fail"], k0 + info if st == _lib.GPL_ERR_NOT_PD else 1)
            else:
                ops.save_dinv(w, self.dinv[k])
                state["logdet"] += ld_k
                if n - k0 - w > 0:
                    ops.trsm_below(A_t, lc0, k0, w, n - k0 - w)
            self.panel_view(k).copy_(A_t[lc0:lc0 + w, k0:n])

        def apply_panel(k, j):
            k0, w = lay.col0(k), lay.width[k]
            j0, wj = lay.col0(j), lay.width[j]
            ops.update(self.panel_view(k), n - k0, w, j0 - k0, n - j0, j0 - k0, wj, A_t, lay.local_col0[j], j0)

        if lay.owner(0) == self.rank:
            factor_panel(0)
        work = dist.broadcast(self.panel_view(0), src=lay.owner(0), async_op=True)
        for k in range(lay.nblk):
            if work is not None:
                work.wait()
            nxt = k + 1
            work = None
            if nxt < lay.nblk:
                if lay.owner(nxt) == self.rank:
                    apply_panel(k, nxt)
                    factor_panel(nxt)
                work = dist.broadcast(self.panel_view(nxt), src=lay.owner(nxt), async_op=True)
            for j in lay.owned:
                if j > k and not (j == nxt):
                    apply_panel(k, j)
        red = torch.tensor([float(state["fail"])], dtype=torch.float64, device=self.device)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        if red.item() != 0:
            raise _lib.NotPositiveDefinite(_lib.GPL_ERR_NOT_PD,
                                           f"the leading minor of order {int(red.item())} is not positive definite")
        ldt = torch.tensor([state["logdet"]], dtype=torch.float64, device=self.device)
        dist.all_reduce(ldt)
        self.tock("factor")
        return float(ldt.item())

    # ---- triangular solves --------------------------------------------------------------------
    def solve(self, rhs):
        """B^-1 rhs (replicated in, replicated out): block forward / backward substitution with an
        8 KB all-reduce / broadcast per block"""
        torch, lay, ops, n, dist, A_t, nb = self.torch, self.lay, self.ops, self.n, self.dist, self.A_t, self.nb
        dev = self.device
        self.tick()
        u = torch.zeros(n, dtype=torch.float64, device=dev)       # local partial sums  sum_j L[:, j] w_j
        wv = torch.zeros(n, dtype=torch.float64, device=dev)      # forward solution (owned blocks)
        tmp = torch.empty(nb, dtype=torch.float64, device=dev)
        tmp2 = torch.empty(nb, dtype=torch.float64, device=dev)
        for k in range(lay.nblk):
            k0, w = lay.col0(k), lay.width[k]
            blk = u[k0:k0 + w].clone()
            dist.all_reduce(blk)
            if lay.owner(k) == self.rank:
                lc0 = lay.local_col0[k]
                tmp[:w] = rhs[k0:k0 + w] - blk
                ops.trsv(0, w, A_t, lc0, k0, self.dinv[k], tmp, tmp2)
                wv[k0:k0 + w] = tmp2[:w]
                if n - k0 - w > 0:
                    upd = torch.empty(n - k0 - w, dtype=torch.float64, device=dev)
                    ops.gemv(0, n - k0 - w, w, A_t, lc0, k0 + w, tmp2, 0, upd, 0)
                    u[k0 + w:] += upd
        xv = torch.zeros(n, dtype=torch.float64, device=dev)      # backward solution, replicated
        for k in reversed(range(lay.nblk)):
            k0, w = lay.col0(k), lay.width[k]
            blk = torch.zeros(w, dtype=torch.float64, device=dev)
            if lay.owner(k) == self.rank:
                lc0 = lay.local_col0[k]
                tmp[:w] = wv[k0:k0 + w]
                if n - k0 - w > 0:
                    dot = torch.empty(w, dtype=torch.float64, device=dev)
                    ops.gemv(1, n - k0 - w, w, A_t, lc0, k0 + w, xv, k0 + w, dot, 0)
                    tmp[:w] -= dot
                ops.trsv(1, w, A_t, lc0, k0, self.dinv[k], tmp, tmp2)
                blk.copy_(tmp2[:w])
            dist.broadcast(blk, src=lay.owner(k))
            xv[k0:k0 + w] = blk
        self.tock("solve")
        return xv

    # ---- inverse: B^-1 = X^T X, X = L^-1 (SURVEY 8e "C^-1 + trace term") ----------------------------
    # Two more sweeps of panel broadcasts, each n^3/3 flops spread like the factorisation:
    #   X sweep: panel k = [L_kk^-1 ; L[k1:, k]] goes to every rank; for its own columns j < k a rank
    #            finishes block row k of X (X_kj = L_kk^-1 R_kj) and pushes the update down
    #            (R[k1:, j] -= L[k1:, k] X_kj); the running right-hand side R lives in X's slab.
    #   Y sweep: panel i = X[i0:, i]; Y_ij = X[i0:, i]^T X[i0:, j] for own columns j <= i, written
    #            over the block row of X it no longer needs.
    # Panels depend only on finished data, so panel k+1 is always in flight during step k.
    def x_sweep(self):
        torch, lay, ops, n, A_t, ldn = self.torch, self.lay, self.ops, self.n, self.A_t, self.ld
        self.tick()
        if self.Xi_t is None:
            self.Xi_t = torch.empty((max(lay.nloc, 1), self.ld), dtype=torch.float64, device=self.device)
            self.tmpb = torch.empty(self.nb * self.nb, dtype=torch.float64, device=self.device)
        Xi_t, tmpb = self.Xi_t, self.tmpb
        Xi_t.zero_()

        def prep_L_panel(k):
            k0, w, lc0 = lay.col0(k), lay.width[k], lay.local_col0[k]
            pv = self.panel_view(k)                              # (w, n - k0): column-major (n-k0) x w
            pv.copy_(A_t[lc0:lc0 + w, k0:n])
            ops.inv_diag(A_t, lc0, k0, w, self.dinv[k], pv, n - k0)   # diagonal block <- L_kk^-1
            # this rank's own column k of X:  X_kk = L_kk^-1,  R[k1:, k] = -L[k1:, k] L_kk^-1
            Xi_t[lc0:lc0 + w, k0:k0 + w] = pv[:, :w]
            if n - k0 - w > 0:
                ops.gemm(0, 0, n - k0 - w, w, w, -1.0, (pv, w, n - k0), (pv, 0, n - k0), 0.0,
                         (Xi_t, lc0 * ldn + k0 + w, ldn))

        def apply_L_panel(k):
            k0, w = lay.col0(k), lay.width[k]
            pv = self.panel_view(k)
            for j in lay.owned:
                if j >= k:
                    break
                wj, lcj = lay.width[j], lay.local_col0[j]
                # X_kj = L_kk^-1 R_kj (into a scratch block, then over R_kj)
                ops.gemm(0, 0, w, wj, w, 1.0, (pv, 0, n - k0), (Xi_t, lcj * ldn + k0, ldn), 0.0, (tmpb, 0, w))
                if n - k0 - w > 0:
                    ops.gemm(0, 0, n - k0 - w, wj, w, -1.0, (pv, w, n - k0), (tmpb, 0, w), 1.0,
                             (Xi_t, lcj * ldn + k0 + w, ldn))
                Xi_t[lcj:lcj + wj, k0:k0 + w] = torch.as_strided(tmpb, (wj, w), (w, 1))

        self._sweep(prep_L_panel, apply_L_panel)
        self.tock("trtri")

    def diag_binv(self):
        """diag(B^-1)_i = sum_k X_ki^2 after x_sweep: column sums of squares of the owned block
        columns, assembled with one all-reduce of an n-vector"""
        torch, lay, n = self.torch, self.lay, self.n
        out = torch.zeros(n, dtype=torch.float64, device=self.device)
        for j in lay.owned:
            j0, w, lcj = lay.col0(j), lay.width[j], lay.local_col0[j]
            self.ops.colsumsq(self.Xi_t, lcj, j0, n - j0, w, out, j0)
        self.dist.all_reduce(out)
        self.tock("diag")
        return out

    def y_sweep(self):
        torch, lay, ops, n, ldn = self.torch, self.lay, self.ops, self.n, self.ld
        Xi_t, tmpb = self.Xi_t, self.tmpb
        self.tick()

        def prep_X_panel(i):
            i0, w, lc0 = lay.col0(i), lay.width[i], lay.local_col0[i]
            self.panel_view(i).copy_(Xi_t[lc0:lc0 + w, i0:n])

        def apply_X_panel(i):
            i0, w = lay.col0(i), lay.width[i]
            pv = self.panel_view(i)
            for j in lay.owned:
                if j > i:
                    break
                wj, lcj = lay.width[j], lay.local_col0[j]
                ops.gemm(1, 0, w, wj, n - i0, 1.0, (pv, 0, n - i0), (Xi_t, lcj * ldn + i0, ldn), 0.0,
                         (tmpb, 0, w), 1 if j == i else 0)
                Xi_t[lcj:lcj + wj, i0:i0 + w] = torch.as_strided(tmpb, (wj, w), (w, 1))

        self._sweep(prep_X_panel, apply_X_panel)
        self.tock("lauum")

    def grad_trace(self, s_t, alpha, pf, ps):
        """sum_ij (alpha_i alpha_j - s_i Binv_ij s_j) dK_ij/dlog(theta_p) and sum_j s_j^2 Binv_jj after
        y_sweep: local block traces + one all-reduce of |theta| + 1 numbers"""
        torch, lay = self.torch, self.lay
        g_loc = np.zeros(len(pf), dtype=np.float64)
        dsum_loc = 0.0
        for j in lay.owned:
            gj, dj = self.ops.grad_trace_block(self.pm, self.X_t, self.n, self.d, lay.col0(j), lay.width[j],
                                               self.Xi_t, lay.local_col0[j], s_t, alpha, pf, ps)
            g_loc += gj
            dsum_loc += dj
        red = torch.from_numpy(np.concatenate([g_loc, [dsum_loc]])).to(self.device)
        self.dist.all_reduce(red)
        red = red.cpu().numpy()
        self.tock("trace")
        return red[:-1], red[-1]

    def finish_timings(self):
        if self.timings is not None:
            self.timings.update(self.t_acc)


def dist_fit_gaussian(gp, x, y, jitter=None, nb=1024, ops=None, dist=None, device="cuda", timings=None):
    """lik_gaussian entry point kept for callers of the first version: see dist_fit_laplace."""
    if gp.lik.name != "lik_gaussian":
        raise NotImplementedError("dist_fit_gaussian is for lik_gaussian; use dist_fit_laplace")
    return dist_fit_laplace(gp, x, y, jitter=jitter, nb=nb, ops=ops, dist=dist, device=device,
                            timings=timings)


def dist_fit_laplace(gp, x, y, trials=None, offset=None, jitter=None, nb=1024, ops=None, dist=None,
                     device="cuda", timings=None, grad_params=None):
    """Distributed gp_fit for method_full + approx_laplace, any likelihood (log-concave sites).
    Every rank passes the same (replicated) x, y; returns dict(log_evidence, alpha, fmean, z, V
    (replicated torch tensors), iters, converged).

    `dist` is torch.distributed (initialised by the caller: NCCL on GPUs); `ops` defaults to the
    C-ABI GpuOps on the calling process's default context.

    grad_params=(param_factor, param_slot) (lik_gaussian only): also return `grad_raw`, the
    derivative of the energy without priors w.r.t. those log-parameters -- see dist_energy_grad."""
    import time
    import torch
    from .gp import _pack_lik
    if gp.method.name != "full" or gp.approx.name != "laplace":
        raise NotImplementedError("dist_fit_laplace covers method_full + approx_laplace (EP: dist_fit_ep)")
    if dist is None:
        import torch.distributed as dist
    t_start = time.perf_counter()
    sf = ShardedFactor(gp, x, nb, ops, dist, device, jitter, timings)
    ops, n = sf.ops, sf.n
    gaussian = gp.lik.name == "lik_gaussian"
    lik = _pack_lik(gp.lik)
    y_t, tr_t, off_t = sf.vec(y), sf.vec(trials), sf.vec(offset)
    if gp.lik.name in ("lik_binomial", "lik_betabinom") and tr_t is None:
        raise ValueError("trials must be provided for the binomial / beta-binomial likelihood")
    f_t = torch.zeros(n, dtype=torch.float64, device=device)               # R/laplace.R:280
    z_t = torch.empty(n, dtype=torch.float64, device=device)
    V_t = torch.empty(n, dtype=torch.float64, device=device)
    maxiter = 1 if gaussian else max(int(gp.approx.maxiter), 1)            # R/laplace.R:281-283
    tol = float(gp.approx.tol)
    it = 0
    for it in range(1, maxiter + 1):                                       # R/laplace.R:285-292
        # one laplace_iter.method_full (R/laplace.R:13-38) on the sharded matrix
        ops.pseudodata(lik, f_t, y_t, tr_t, off_t, z_t, V_t)
        if not bool((V_t > 0).all().item()):
            raise _lib.GplError(_lib.GPL_ERR_NOT_LOGCONCAVE,
                                "pseudo-variance <= 0: the non-log-concave branch (R/laplace.R:40-117) "
                                "is single-GPU only")
        s_t = torch.rsqrt(V_t)
        logdet = sf.factor(s_t)                                            # C_logdet - V_logdet = logdet(B)
        alpha = sf.solve(z_t * s_t) * s_t                                  # R/laplace.R:29
        fnew = z_t - V_t * alpha                                           # == K alpha
        log_lik = ops.loglik(lik, fnew, y_t, tr_t, off_t)
        log_evidence = log_lik - 0.5 * float(torch.dot(fnew, alpha).item()) - 0.5 * logdet   # R/laplace.R:33-38
        diff = float(torch.max(torch.abs(f_t - fnew)).item())
        f_t = fnew
        if diff < tol:
            break
    out = dict(log_evidence=log_evidence, alpha=alpha, fmean=f_t, z=z_t, V=V_t, iters=it,
               converged=not (maxiter > 1 and it == maxiter), layout=sf.lay)
    if grad_params is not None:
        if not gaussian:
            raise NotImplementedError("the analytic gradient is implemented for lik_gaussian")
        pf, ps = (np.ascontiguousarray(v, dtype=np.int32) for v in grad_params)
        sf.x_sweep()
        sf.y_sweep()
        g_cf, dsum = sf.grad_trace(s_t, alpha, pf, ps)
        sigma2 = gp.lik.sigma ** 2
        tr_w = float(torch.dot(alpha, alpha).item()) - dsum                # tr(alpha alpha^T - C^-1)
        out["grad_raw"] = np.where(pf < 0, -0.5 * 2.0 * sigma2 * tr_w, -0.5 * g_cf)
    sf.finish_timings()
    if timings is not None:
        for k in ("kbuild", "factor", "solve", "trtri", "lauum", "trace"):
            timings.setdefault(k, 0.0)
        ops.sync()
        timings["total"] = time.perf_counter() - t_start
    return out


def dist_fit_ep(gp, x, y, trials=None, offset=None, jitter=None, nb=1024, ops=None, dist=None,
                device="cuda", timings=None):
    """Distributed gp_fit for method_full + approx_ep (SURVEY 8e "Laplace/EP site sweeps"): parallel
    EP (R/ep.R:16-49, 286-330) over the sharded factor.  Per sweep: the site update on replicated
    n-vectors (gpl_lik_ep_sites_dev), B = I + V^-1/2 K V^-1/2 rebuilt and factorised, alpha, and the
    marginal variances from diag(B^-1) (the X sweep: fvar = V - V diag(B^-1), the same identity as on
    one GPU instead of the reference's n x n solve, R/ep.R:42,141)."""
    import torch
    from .gp import _pack_lik
    from .quadrature import gauss_hermite_points
    if gp.method.name != "full" or gp.approx.name != "ep":
        raise NotImplementedError("dist_fit_ep covers method_full + approx_ep")
    if dist is None:
        import torch.distributed as dist
    sf = ShardedFactor(gp, x, nb, ops, dist, device, jitter, timings)
    ops, n = sf.ops, sf.n
    lik = _pack_lik(gp.lik)
    y_t, tr_t, off_t = sf.vec(y), sf.vec(trials), sf.vec(offset)
    if gp.lik.name in ("lik_binomial", "lik_betabinom") and tr_t is None:
        raise ValueError("trials must be provided for the binomial / beta-binomial likelihood")
    a = gp.approx
    xq, wq = gauss_hermite_points(int(a.quad_order))
    xq_t, wq_t = sf.vec(xq), sf.vec(wq)
    new = lambda: torch.zeros(n, dtype=torch.float64, device=device)
    P_old, z_old, fmean_old = new(), new(), new()
    fvar_old = new()
    ops.kdiag(sf.pm, sf.X_t, n, sf.d, sf.jit, fvar_old)                    # fvar_old = diag(K), R/ep.R:297
    z_t, V_t, P_t, cm_t, cv_t, logZ_t = new(), new(), new(), new(), new(), new()
    maxiter = max(int(a.maxiter), 1)
    tol, damping, damping_min = 1e-4, float(a.damping), 0.1                # R/ep.R:286-287
    diff_old, damp = math.inf, damping
    quad_ok, it = True, 0
    fit = None
    for it in range(1, maxiter + 1):
        flags = ops.ep_sites(lik, fmean_old, fvar_old, z_old, P_old, y_t, tr_t, off_t, damp, xq_t, wq_t,
                             z_t, V_t, P_t, cm_t, cv_t, logZ_t)
        quad_ok = not (flags & 1)
        if flags & 2:                                                      # R/ep.R:30-32
            return dict(log_evidence=-math.inf, iters=it, converged=False, quad_ok=quad_ok)
        if flags & 4:
            raise _lib.GplError(_lib.GPL_ERR_NOT_LOGCONCAVE,
                                "pseudo-variance <= 0: the non-log-concave branch (R/ep.R:50-138) is single-GPU only")
        s_t = torch.rsqrt(V_t)
        logdet_b = sf.factor(s_t)
        alpha = sf.solve(z_t * s_t) * s_t
        fmean = z_t - V_t * alpha                                          # == K alpha
        sf.x_sweep()
        fvar = V_t - V_t * sf.diag_binv()                                  # diag(V C^-1 K)
        C_logdet = float(torch.sum(torch.log(V_t)).item()) + logdet_b
        tot = cv_t + V_t
        log_evidence = (-0.5 * C_logdet - 0.5 * float(torch.dot(z_t, alpha).item())
                        + float((torch.sum(logZ_t) + 0.5 * torch.sum(torch.log(tot))
                                 + 0.5 * torch.sum((cm_t - z_t) ** 2 / tot)).item()))       # R/ep.R:45-49
        fit = dict(fmean=fmean, fvar=fvar, z=z_t.clone(), V=V_t.clone(), P=P_t.clone(), alpha=alpha,
                   cavity_mean=cm_t.clone(), cavity_var=cv_t.clone(), log_evidence=log_evidence)
        diff_new = float(torch.max(torch.abs(fmean_old - fmean)).item())
        if diff_new > diff_old:                                            # R/ep.R:306-310
            damp *= damping
            continue
        damp = max(damp, damping_min)
        fmean_old, fvar_old, z_old, P_old = fmean, fvar, fit["z"], fit["P"]
        diff_old = diff_new
        if diff_new < tol:
            break
    fit["iters"] = it
    fit["converged"] = not (maxiter > 1 and it == maxiter)
    fit["quad_ok"] = quad_ok
    if maxiter > 1 and it == maxiter:                                      # R/ep.R:322-325
        fit["log_evidence"] = -math.inf
    if not quad_ok:
        fit["log_evidence"] = -math.inf
    fit["layout"] = sf.lay
    sf.finish_timings()
    return fit


def dist_energy_grad(gp, x, y, jitter=None, include_prior=True, nb=1024, ops=None, dist=None, device="cuda",
                     timings=None):
    """Multi-GPU gp_energy_grad (lik_gaussian + method_full): energy and its gradient w.r.t.
    get_param(gp), same conventions as gp.gp_energy_grad; every rank returns the same values."""
    items = M.get_param_objs(gp.cfs, gp.lik)
    _, where = M.flatten_cfs(gp.cfs)
    pf = np.empty(len(items), dtype=np.int32)
    ps = np.empty(len(items), dtype=np.int32)
    for j, (_, owner, attr) in enumerate(items):
        pf[j], ps[j] = (-1, 0) if owner is gp.lik else where[(id(owner), attr)]
    res = dist_fit_laplace(gp, x, y, jitter=jitter, nb=nb, ops=ops, dist=dist, device=device, timings=timings,
                           grad_params=(pf, ps))
    e = -res["log_evidence"]
    grad = res["grad_raw"].copy()
    if include_prior:
        for j, (_, owner, attr) in enumerate(items):
            lp = math.log(getattr(owner, attr))
            e -= M.lpdf_prior_one(owner.priors[attr], lp)
            grad[j] -= M.dlpdf_prior_one(owner.priors[attr], lp)
    return e, grad


# ================================================================================================
# FITC across GPUs ("shards naturally, one reduction")
# ================================================================================================

class FitcOps:
    """C-ABI calls of the row-sharded FITC driver on torch CUDA tensors.  Matrices are column-major
    slabs: a (rows x cols) matrix is a tensor of shape (cols, ld), element (i, j) at t[j, i]."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.lib = ctx.lib

    _p = staticmethod(GpuOps._p)

    def ksym(self, pm, Z_t, m, d, jitter, out_t):
        """out (m x m) = K(z, z) + jitter I"""
        _lib.check(self.lib.gpl_cf_eval_dev(self.ctx.h, pm.ref(), self._p(Z_t), m, self._p(Z_t), m, d, 1, 0, 0,
                                            float(jitter), self._p(out_t), out_t.shape[1]))

    def kblock(self, pm, X_t, nr, ldx, Z_t, m, d, out_t):
        """out (nr x m) = K(x_local, z)"""
        _lib.check(self.lib.gpl_cf_eval_block_dev(self.ctx.h, pm.ref(), self._p(X_t), nr, ldx, self._p(Z_t), m, m,
                                                  d, self._p(out_t), out_t.shape[1]))

    def kdiag(self, pm, X_t, nr, d, jitter, out_t):
        _lib.check(self.lib.gpl_cf_eval_dev(self.ctx.h, pm.ref(), self._p(X_t), nr, self._p(X_t), nr, d, 1, 0, 1,
                                            float(jitter), self._p(out_t), nr))

    def potrf(self, A_t, m, dinv_t):
        info = C.c_int32()
        logdet = C.c_double()
        st = self.lib.gpl_dpotrf_dev(self.ctx.h, m, self._p(A_t), A_t.shape[1], C.byref(info), C.byref(logdet))
        if st == _lib.GPL_OK:
            _lib.check(self.lib.gpl_dpotrf_get_dinv_dev(self.ctx.h, m, self._p(dinv_t)))
        return st, info.value, logdet.value

    def rownorms_after_solve(self, W_t, rows, m, L_t, dinv_t, out_t, out_off):
        """out[off:off+rows] = row sums of squares of W L^-T  (W is destroyed)"""
        _lib.check(self.lib.gpl_dtrsm_right_dinv_dev(self.ctx.h, 1, rows, m, self._p(L_t), L_t.shape[1],
                                                     self._p(dinv_t), self._p(W_t), W_t.shape[1]))
        _lib.check(self.lib.gpl_rowsumsq_dev(self.ctx.h, rows, m, self._p(W_t), W_t.shape[1], self._p(out_t, out_off)))

    def scale_rows(self, W_t, rows, m, s_t, s_off):
        _lib.check(self.lib.gpl_scale_block_dev(self.ctx.h, rows, m, self._p(W_t), W_t.shape[1], self._p(s_t, s_off),
                                                None, 1.0, 0.0, -(1 << 40)))

    def syrk_acc(self, W_t, rows, m, A_t):
        """A (lower) += W^T W"""
        _lib.check(self.lib.gpl_dgemm_dev(self.ctx.h, 1, 0, m, m, rows, 1.0, self._p(W_t), W_t.shape[1],
                                          self._p(W_t), W_t.shape[1], 1.0, self._p(A_t), A_t.shape[1], 1))

    def gemv(self, trans, rows, m, K_t, x_t, y_t):
        _lib.check(self.lib.gpl_dgemv_dev(self.ctx.h, trans, rows, m, self._p(K_t), K_t.shape[1], self._p(x_t),
                                          self._p(y_t)))

    pseudodata = GpuOps.pseudodata
    loglik = GpuOps.loglik

    def trsv(self, trans, m, L_t, dinv_t, b_t, x_t):
        _lib.check(self.lib.gpl_dtrsv_dev(self.ctx.h, trans, m, self._p(L_t), L_t.shape[1], self._p(dinv_t),
                                          self._p(b_t), self._p(x_t)))

    def sync(self):
        self.ctx.sync()


def dist_fit_fitc_gaussian(gp, x, y, jitter=None, ops=None, dist=None, device="cuda", chunk_rows=None):
    """lik_gaussian entry point kept for callers of the first version: see dist_fit_fitc."""
    if gp.lik.name != "lik_gaussian":
        raise NotImplementedError("dist_fit_fitc_gaussian is for lik_gaussian; use dist_fit_fitc")
    return dist_fit_fitc(gp, x, y, jitter=jitter, ops=ops, dist=dist, device=device, chunk_rows=chunk_rows)


def dist_fit_fitc(gp, x, y, trials=None, offset=None, jitter=None, ops=None, dist=None, device="cuda",
                  chunk_rows=None):
    """Row-sharded FITC fit with approx_laplace for any likelihood with log-concave sites
    (R/gp_fit.R:101-123 + R/laplace.R:122-152,299-325 + R/linear_algebra.R:3-43): rank r owns rows
    [r n/P, (r+1) n/P) of x and the site vectors of those rows; Kz and the m x m factorisations are
    replicated; per Newton step the only data-path collectives are one all-reduce of the m x m matrix
    Kzx S^-1 Kxz and a few m-vectors / scalars.  Every rank passes the full x, y."""
    import torch
    from .gp import _pack_lik
    if gp.method.name != "fitc" or gp.approx.name != "laplace":
        raise NotImplementedError("dist_fit_fitc covers method_fitc + approx_laplace")
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    y = np.asarray(y, dtype=np.float64).ravel()
    zind = np.asarray(gp.method.inducing, dtype=np.float64)
    if zind.ndim == 1:
        zind = zind[:, None]
    n, d = x.shape
    m = zind.shape[0]
    for cf in gp.cfs:
        M.learn_scales(cf, x)                       # training means / scales use ALL rows
    jit = 1e-6 if jitter is None else float(jitter)
    gaussian = gp.lik.name == "lik_gaussian"
    lik = _pack_lik(gp.lik)
    factors, _ = M.flatten_cfs(gp.cfs)
    pm = _lib.PackedModel(factors)
    if ops is None:
        ops = FitcOps(_lib.Context(torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream or 1))
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    nr = hi - lo
    f64 = dict(dtype=torch.float64, device=device)
    ldr = (max(nr, 1) + 15) // 16 * 16
    ldm = (m + 15) // 16 * 16
    loc = lambda v: None if v is None else torch.from_numpy(
        np.asarray(v, dtype=np.float64).ravel()[lo:hi].copy()).to(device)
    X_t = torch.from_numpy(np.ascontiguousarray(x[lo:hi].T)).to(device)          # (d, nr)
    Z_t = torch.from_numpy(np.ascontiguousarray(zind.T)).to(device)              # (d, m)
    y_t, tr_t, off_t = loc(y), loc(trials), loc(offset)
    if gp.lik.name in ("lik_binomial", "lik_betabinom") and tr_t is None:
        raise ValueError("trials must be provided for the binomial / beta-binomial likelihood")
    nblk = (m + 127) // 128
    Kz = torch.zeros((m, ldm), **f64)
    Lz = torch.zeros((m, ldm), **f64)
    A = torch.zeros((m, ldm), **f64)
    dinv_z = torch.empty(nblk * 128 * 128, **f64)
    dinv_a = torch.empty(nblk * 128 * 128, **f64)
    Kxz = torch.zeros((m, ldr), **f64)
    ops.ksym(pm, Z_t, m, d, jit, Kz)
    Lz.copy_(Kz)
    st, info, logdet_z = ops.potrf(Lz, m, dinv_z)
    if st != _lib.GPL_OK:
        raise _lib.NotPositiveDefinite(st, "Kz is not positive definite")
    ops.kblock(pm, X_t, nr, nr, Z_t, m, d, Kxz)
    Kdiag = torch.empty(max(nr, 1), **f64)
    ops.kdiag(pm, X_t, nr, d, jit, Kdiag)
    # D = K_diag - |Lz^-1 k_zi|^2, streamed through a row-chunk workspace
    if chunk_rows is None:
        chunk_rows = max(1024, int(2.0e9 / (8.0 * m)) // 128 * 128)
    chunk_rows = min(chunk_rows, max(nr, 1))
    ldw = (chunk_rows + 15) // 16 * 16
    W = torch.empty((m, ldw), **f64)
    q = torch.empty(max(nr, 1), **f64)
    for r0 in range(0, nr, chunk_rows):
        c = min(chunk_rows, nr - r0)
        W[:, :c].copy_(Kxz[:, r0:r0 + c])
        ops.rownorms_after_solve(W, c, m, Lz, dinv_z, q, r0)
    D = Kdiag[:nr] - q[:nr]
    f_loc = torch.zeros(nr, **f64)                                     # R/laplace.R:305
    z_loc = torch.empty(nr, **f64)
    V_loc = torch.empty(nr, **f64)
    w1 = torch.empty(m, **f64)
    maxiter = 1 if gaussian else max(int(gp.approx.maxiter), 1)        # R/laplace.R:306-308
    tol = float(gp.approx.tol)
    log_post_old = -math.inf
    it = 0
    for it in range(1, maxiter + 1):
        # laplace_iter.method_fitc, R/laplace.R:122-152, on the local rows
        if nr > 0:
            ops.pseudodata(lik, f_loc, y_t, tr_t, off_t, z_loc, V_loc)
        vmin = torch.tensor([float(V_loc.min().item()) if nr > 0 else math.inf], **f64)
        dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
        if not vmin.item() > 0:
            raise _lib.GplError(_lib.GPL_ERR_NOT_LOGCONCAVE,
                                "pseudo-variance <= 0: the non-log-concave branch (R/laplace.R:153-232) "
                                "is single-GPU only")
        S = D + V_loc
        rs = torch.rsqrt(S).contiguous()
        # A = Kz + sum over ranks of Kzx S^-1 Kxz  (lower triangle)
        part = torch.zeros((m, ldm), **f64)
        for r0 in range(0, nr, chunk_rows):
            c = min(chunk_rows, nr - r0)
            W[:, :c].copy_(Kxz[:, r0:r0 + c])
            ops.scale_rows(W, c, m, rs, r0)
            ops.syrk_acc(W, c, m, part)
        dist.all_reduce(part)
        A.copy_(Kz + part)
        st, info, logdet_a = ops.potrf(A, m, dinv_a)
        if st != _lib.GPL_OK:
            raise _lib.NotPositiveDefinite(st, "Kz + Kzx S^-1 Kxz is not positive definite")
        # alpha = z/S - S^-1 Kxz A^-1 Kzx (z/S)
        t = (z_loc / S).contiguous()
        u = torch.zeros(m, **f64)
        if nr > 0:
            ops.gemv(1, nr, m, Kxz, t, u)
        dist.all_reduce(u)
        v = torch.empty(m, **f64)
        ops.trsv(0, m, A, dinv_a, u.clone(), w1)
        ops.trsv(1, m, A, dinv_a, w1.clone(), v)
        kv = torch.zeros(max(nr, 1), **f64)
        if nr > 0:
            ops.gemv(0, nr, m, Kxz, v, kv)
        alpha = t - kv[:nr] / S
        # f = Kxz Kz^-1 Kzx alpha + D alpha
        u2 = torch.zeros(m, **f64)
        if nr > 0:
            ops.gemv(1, nr, m, Kxz, alpha.contiguous(), u2)
        dist.all_reduce(u2)
        beta = torch.empty(m, **f64)
        ops.trsv(0, m, Lz, dinv_z, u2.clone(), w1)
        ops.trsv(1, m, Lz, dinv_z, w1.clone(), beta)
        kb = torch.zeros(max(nr, 1), **f64)
        if nr > 0:
            ops.gemv(0, nr, m, Kxz, beta, kb)
        fmean = (kb[:nr] + D * alpha).contiguous()
        ll = ops.loglik(lik, fmean, y_t, tr_t, off_t) if nr > 0 else 0.0
        sc = torch.stack([torch.log(S).sum(), torch.dot(fmean, alpha), torch.log(V_loc).sum(),
                          torch.tensor(ll, **f64)])
        dist.all_reduce(sc)
        sum_log_S, f_invK_f, V_logdet, log_lik = [float(vv) for vv in sc]
        C_logdet = sum_log_S - logdet_z + logdet_a                                # R/linear_algebra.R:32
        log_evidence = log_lik - 0.5 * f_invK_f + 0.5 * (V_logdet - C_logdet)     # R/laplace.R:143-150
        log_post = log_lik - 0.5 * f_invK_f
        f_loc = fmean
        diff = log_post - log_post_old
        log_post_old = log_post
        if abs(diff) < tol:                                                       # R/laplace.R:314-318
            break
    return dict(log_evidence=log_evidence, alpha_local=alpha, fmean_local=f_loc, z_local=z_loc, V_local=V_loc,
                rows=(lo, hi), beta=beta, iters=it, converged=not (maxiter > 1 and it == maxiter))
