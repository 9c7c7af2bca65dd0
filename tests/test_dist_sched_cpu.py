"""CPU, world sizes 1-4 over gloo: the multi-GPU schedule of the product (gplite_b200/csrc/dist_sched.hpp --
block-cyclic ownership, look-ahead panel order, staircase trailing updates, distributed triangular
solves, the two inverse sweeps) instantiated with a host back end (tests/sim/dist_sim.cpp, test-only)
whose collectives are torch.distributed broadcasts / all-reduces.  The same header is what
libgplite_b200.so compiles with its CUDA + NCCL back end, so the index arithmetic and the ORDER in
which every rank issues its collectives (a mismatch deadlocks or corrupts here exactly as it would
under NCCL) are exercised without a GPU.  Results are compared with numpy on the full matrix."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_SO = os.path.join(ROOT, "tests", "sim", "_build", "libdist_sim.so")


def _build_sim():
    os.makedirs(os.path.dirname(SIM_SO), exist_ok=True)
    src = os.path.join(ROOT, "tests", "sim", "dist_sim.cpp")
    hdr = os.path.join(ROOT, "gplite_b200", "csrc", "dist_sched.hpp")
    if os.path.exists(SIM_SO) and os.path.getmtime(SIM_SO) > max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-shared", "-fPIC",
                        "-I" + os.path.join(ROOT, "gplite_b200", "csrc"), "-o", SIM_SO, src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(n, seed, pd=True):
    rng = np.random.default_rng(seed)
    Q = rng.normal(size=(n, n // 3 + 2))
    B = Q @ Q.T / Q.shape[1] + np.eye(n)          # "I + S K S"-like: SPD, eigenvalues >= 1
    if not pd:
        B[n // 2, n // 2] = -5.0                    # leading minor of order n//2 + 1 fails
    rhs = rng.normal(size=n)
    return np.asfortranarray(B), rhs


BCAST = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int64, C.c_int)
ALLRED = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int64)


def _run_rank(rank, world, n, nb, B, rhs, dist, do_inverse=True, window=2):
    import torch
    lib = C.CDLL(SIM_SO)

    def bcast(buf, count, root):
        t = torch.from_numpy(np.ctypeslib.as_array(buf, shape=(count,)))
        dist.broadcast(t, src=root)

    def allred(buf, count):
        t = torch.from_numpy(np.ctypeslib.as_array(buf, shape=(count,)))
        dist.all_reduce(t)

    cb1, cb2 = BCAST(bcast), ALLRED(allred)
    ld = (n + 15) // 16 * 16
    nblk = (n + nb - 1) // nb
    nloc = sum(min(nb, n - b * nb) for b in range(nblk) if b % world == rank)
    slab = max(nloc, 1) * ld
    logdet = C.c_double()
    sol, dg = np.zeros(n), np.zeros(n)
    Ls, Ys = np.zeros(slab), np.zeros(slab)
    nloc_o, ld_o, info, ncoll = C.c_int64(), C.c_int64(), C.c_int(), C.c_int64()
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = lib.sim_run(rank, world, C.c_int64(n), C.c_int64(nb), int(window), P(B), P(rhs), cb1, cb2, int(do_inverse),
                     C.byref(logdet), P(sol), P(dg), P(Ls), P(Ys), C.byref(nloc_o), C.byref(ld_o), C.byref(info),
                     C.byref(ncoll))
    assert rc == 0
    assert nloc_o.value == nloc and ld_o.value == ld
    return dict(rank=rank, logdet=logdet.value, sol=sol, diag=dg, L=Ls.reshape(max(nloc, 1), ld),
                Y=Ys.reshape(max(nloc, 1), ld), info=info.value, ncoll=ncoll.value)


def _check(outs, world, n, nb, B, rhs):
    L = np.linalg.cholesky(B)
    Binv = np.linalg.inv(B)
    want_logdet = 2 * np.log(np.diag(L)).sum()
    want_sol = np.linalg.solve(B, rhs)
    assert abs(sum(o["logdet"] for o in outs) - want_logdet) < 1e-10 * abs(want_logdet)
    nblk = (n + nb - 1) // nb
    assert len({o["ncoll"] for o in outs}) == 1            # every rank issued the same number of collectives
    for o in outs:
        assert o["info"] == 0
        assert np.max(np.abs(o["sol"] - want_sol)) < 1e-10 * np.max(np.abs(want_sol))
        assert np.max(np.abs(o["diag"] - np.diag(Binv))) < 1e-10 * np.max(np.diag(Binv))
        lc = 0
        for b in range(o["rank"], nblk, world):
            c0, w = b * nb, min(nb, n - b * nb)
            for j in range(w):
                got_L = o["L"][lc + j, c0 + j:n]
                assert np.max(np.abs(got_L - L[c0 + j:, c0 + j])) < 1e-10
                got_Y = o["Y"][lc + j, c0 + j:n]            # lower triangle of B^-1, this block column
                assert np.max(np.abs(got_Y - Binv[c0 + j:, c0 + j])) < 1e-10 * np.max(np.abs(Binv))
            lc += w


class _NoDist:
    @staticmethod
    def broadcast(t, src):
        raise AssertionError("no collective may be issued with world = 1")

    all_reduce = broadcast


@pytest.mark.parametrize("n,nb,window", [(640, 128, 2), (700, 128, 3), (100, 128, 2), (1030, 256, 5)])
def test_schedule_single_rank(n, nb, window):
    _build_sim()
    B, rhs = _problem(n, n)
    out = _run_rank(0, 1, n, nb, B, rhs, _NoDist, window=window)
    _check([out], 1, n, nb, B, rhs)


def test_schedule_reports_first_bad_pivot():
    _build_sim()
    n, nb = 600, 128
    B, rhs = _problem(n, 5, pd=False)
    out = _run_rank(0, 1, n, nb, B, rhs, _NoDist)
    assert out["info"] == n // 2 + 1


def _worker(rank, world, port, n, nb, q, pd, window):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, rhs = _problem(n, n + world, pd=pd)
        out = _run_rank(rank, world, n, nb, B, rhs, dist, window=window)
        q.put(out)
    finally:
        dist.destroy_process_group()


def _spawn(world, n, nb, pd=True, window=2):
    import torch.multiprocessing as mp
    _build_sim()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, nb, q, pd, window)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(outs, key=lambda o: o["rank"])


@pytest.mark.parametrize("world,n,nb,window", [(2, 1000, 128, 2), (2, 1000, 128, 3), (3, 1111, 128, 3),
                                               (4, 300, 128, 5), (2, 1300, 256, 4), (4, 1500, 128, 5),
                                               (4, 1500, 128, 3), (2, 1400, 128, 8), (3, 1700, 128, 8)])
def test_schedule_over_gloo(world, n, nb, window):
    """world 2 - 4 with ragged last blocks, panel windows W = 2 .. 5 (look-ahead depth min(W - 1, world));
    world 4 with 3 blocks (one rank owns nothing but still joins every collective); nb = 256.  The last two
    cases use the production window (8) and are long enough (11 / 14 panels) that the factorisation starts with
    whole-panel broadcasts and SWITCHES to the column-chunked tail once (n - k) < 4 nb P, as the n = 100k runs do
    (in the shorter cases every panel is already in the tail)."""
    outs = _spawn(world, n, nb, window=window)
    B, rhs = _problem(n, n + world)
    _check(outs, world, n, nb, B, rhs)


def test_schedule_over_gloo_reports_first_bad_pivot():
    outs = _spawn(2, 900, 128, pd=False)
    assert all(o["info"] == 900 // 2 + 1 for o in outs)
