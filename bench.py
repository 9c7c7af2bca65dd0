"""Headline benchmark: gp_energy + gradient evaluations / second at n = 32768, d = 8, fp64
(BASELINE.json configs[1]: synthetic regression, cf_matern52 + lik_gaussian, method_full).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--size 32768]

A "step" is ONE complete energy + gradient evaluation of the full GP: K build (+jitter) ->
B = I + K/sigma^2 -> blocked DMMA Cholesky -> alpha, f = K alpha, log evidence -> (K+V)^-1 via
trtri + lauum -> fused gradient trace.  n^3 fp64 flops per step.

* value : evals/s with x, y already resident in HBM (gpl_energy_grad on device pointers).
* e2e   : same metric through the public host API (gp_energy_grad on numpy arrays: H2D of x, y and
          D2H of energy + gradient inside the timed region).
* N > 1 : the n = 32768 problem fits one GPU (north star: "smaller problems stay on one GPU"), so for
          `value` the ranks are independent replicas evaluating different hyper-parameters (what a
          parallel simplex / multi-start gp_optim does): weak scaling, no data-path collective.
* sharded: at EVERY N (1 included) the line also carries a `sharded` record: BASELINE config 5's model
          (cf_matern32, d = 8) through the multi-GPU entries of the C ABI (gpl_dist_*: K built
          block-cyclic, NCCL panel broadcasts inside the library): an in-run parity check at
          n = 6144 (N-rank vs the single-GPU path vs the oracle), the factorisation at n = 100000
          (strong scaling across the driver's N = 1, 2, 4, 8 runs: seconds, aggregate TFLOP/s, NCCL
          bytes) and, where it fits, n = 200000 (factorisation at N >= 4, energy + gradient at N = 8).
* roofline: the dominant kernel is the FP64 tensor-core GEMM/SYRK (gemm_f64_kernel, DMMA.8x8x4).
          achieved = algorithmic n^3 flops per step / summed CUDA-event durations of that kernel
          per step (gpl_ctx_profile, recorded on the launching stream in a separate profiled
          pass so the timed steps carry no event overhead); peak = cuBLAS DGEMM measured in this
          same process (MEASURED_PEAKS.json has no fp64 entry; its bf16 figure does not bound an
          fp64 kernel).
* cpu_baseline / --impl reference: the reference path on the host cores, sampled: the K build by the
          reference's own src/cf.cpp (oracle/_ref, single thread as in the reference) on a slab of
          rows of the full n = 32768 matrix (x n / rows: entries are independent), the dense linear
          algebra (dpotrf, dpotri, trace) by scipy/LAPACK with all cores at n = 8192, scaled by
          (n / 8192)^p; cpu_baseline measures p between 8192 and 16384 instead of assuming 3.
          The line says `"extrapolated": true`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FULL, D_FULL = 32768, 8
LSCALE, MAGN, SIGMA, JITTER = 1.3, 1.0, 0.5, 1e-6


def synth(n, d, seed=2):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.normal(size=(n, d)))
    y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + X[:, 2] + 0.5 * rng.normal(size=n)
    return X, np.ascontiguousarray(y)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------- CPU arm
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core the
    reference's LAPACK could (round-1 SCALE ran it on one thread at N > 1).  Returns the thread count in force."""
    n = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
        got = [p.get("num_threads") for p in threadpoolctl.threadpool_info() if p.get("user_api") == "blas"]
        return max(got) if got else n
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", n))


def cpu_kbuild_slab(n, d, rows):
    """seconds for `rows` rows of the n x n K build by the reference's loops (src/cf.cpp:86-119), 1 thread"""
    from oracle import gplite_oracle as o
    X, _ = synth(n, d)
    t0 = time.perf_counter()
    o.cf_native("cf_matern52", X[:rows], X, LSCALE, MAGN, use_ref=o.ref_available())
    return time.perf_counter() - t0


def cpu_linalg_eval(n, d):
    """One energy + gradient evaluation by the reference path on the CPU at size n, WITHOUT the K build
    being timed (K comes from numpy here): seconds for B, dpotrf, the solves, dpotri and the trace."""
    import scipy.linalg as sla
    X, y = synth(n, d)
    sq = np.sum(X * X, 1)
    r2 = -2.0 * (X @ X.T) + sq[:, None] + sq[None, :]
    np.maximum(r2, 0, out=r2)
    tt = np.sqrt(5.0 * r2) / LSCALE
    E = np.exp(-tt)
    K = (MAGN ** 2) * (1 + tt + tt * tt / 3.0) * E
    dK_l = (MAGN ** 2) * (tt * tt / 3.0) * (1 + tt) * E
    del r2, E, tt
    t0 = time.perf_counter()
    K[np.diag_indices(n)] += JITTER
    V = SIGMA ** 2
    B = np.eye(n) + K / V                                      # R/laplace.R:27
    L = sla.cholesky(B, lower=True, check_finite=False, overwrite_a=True)   # R/laplace.R:28 (dpotrf)
    t = sla.solve_triangular(L, y / np.sqrt(V), lower=True, check_finite=False)
    alpha = sla.solve_triangular(L, t, lower=True, trans="T", check_finite=False) / np.sqrt(V)
    f = K @ alpha
    le = (-0.5 * n * np.log(2 * np.pi * V) - 0.5 * np.sum((y - f) ** 2) / V - 0.5 * f @ alpha
          - np.sum(np.log(np.diag(L))))
    Binv, info = sla.lapack.dpotri(L, lower=1, overwrite_c=1)  # gradient: (K+V)^-1 = Binv / V
    W = np.outer(alpha, alpha)
    W -= np.tril(Binv) / V
    Wl = np.tril(W, -1)
    g_l = -0.5 * (2 * np.sum(Wl * dK_l))
    g_m = -0.5 * (2 * np.sum(Wl * 2 * K) + np.sum(np.diag(W) * 2 * (np.diag(K) - JITTER)))
    g_s = -0.5 * 2 * V * np.trace(W)
    return time.perf_counter() - t0, -le, np.array([g_l, g_m, g_s])


def cpu_eval_seconds(n_full, d, rows, n_la, p):
    """(extrapolated seconds of one evaluation at n_full, K-slab seconds, linear-algebra seconds)"""
    tk = cpu_kbuild_slab(n_full, d, rows)
    tl, _, _ = cpu_linalg_eval(n_la, d)
    return tk * (n_full / rows) + tl * (n_full / n_la) ** p, tk, tl


def cpu_baseline(d):
    from oracle import gplite_oracle as o
    threads = use_all_host_threads()
    rows, n1, n2 = 1024, 8192, 16384
    tk = cpu_kbuild_slab(N_FULL, d, rows)
    t1, _, _ = cpu_linalg_eval(n1, d)
    t2, _, _ = cpu_linalg_eval(n2, d)
    p = float(np.log(t2 / t1) / np.log(n2 / n1))
    t_full = tk * (N_FULL / rows) + t2 * (N_FULL / n2) ** p
    return {"value": 1.0 / t_full, "unit": "evals/s", "cores": threads,
            "kind": "reference" if o.ref_available() else "port", "extrapolated": True,
            "exponent_measured": p, "kbuild_slab_s": tk, "linalg_s": {str(n1): t1, str(n2): t2},
            "sample": (f"K build: {rows} of the {N_FULL} rows of the n={N_FULL} matrix by the reference's src/cf.cpp on "
                       f"1 thread ({tk:.2f} s, x{N_FULL // rows}); dpotrf + solves + dpotri + trace by LAPACK on "
                       f"{threads} threads at n={n1} ({t1:.2f} s) and n={n2} ({t2:.2f} s): measured exponent "
                       f"{p:.2f}, extrapolated from n={n2} to n={N_FULL} with it")}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import gplite_oracle as o
    threads = use_all_host_threads()
    rows, n_la, p = 384, 8192, 3.0
    vals = []
    parts = []
    for i in range(args.warmup + args.steps):
        t, tk, tl = cpu_eval_seconds(args.n, D_FULL, rows, n_la, p)
        if i >= args.warmup:
            vals.append(t)
            parts.append((tk, tl))
    t = float(np.mean(vals))
    val = 1.0 / t
    tk, tl = np.mean(parts, axis=0)
    sample = (f"each step: {rows} of the {args.n} rows of the K build by the reference's src/cf.cpp on 1 thread "
              f"({tk:.2f} s, x{args.n / rows:.0f}) + dpotrf / solves / dpotri / trace by LAPACK on {threads} threads "
              f"at n={n_la} ({tl:.2f} s, x(n/{n_la})^{p:.0f}; cpu_baseline of the GPU arm measures the exponent)")
    line = {"impl": "reference", "metric": "gp_energy+grad evals/sec at n=32k fp64", "value": val,
            "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "extrapolated": True,
            "config": {"workload": f"cfg2: n={args.n}, d={D_FULL}, cf_matern52 + lik_gaussian, method_full, "
                                   "energy + analytic gradient (reference path on host cores, sampled and "
                                   "extrapolated: see cpu_baseline.sample)"},
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": threads,
                             "kind": "reference" if o.ref_available() else "port", "extrapolated": True,
                             "sample": sample},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- GPU arm
def synth5(n, d=8, seed=5):
    """SURVEY 8d cfg5 generator: as cfg2 with cf_matern32"""
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.normal(size=(n, d)))
    y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + X[:, 2] + 0.5 * rng.normal(size=n)
    return X, np.ascontiguousarray(y)


def sharded_record(ctx, rank, world, dist, nb=1024):
    """BASELINE config 5 through gpl_dist_* (NCCL inside the library).  Every rank calls this; returns
    the record on every rank (identical numbers: times are the max over ranks)."""
    import torch
    import gplite_b200 as g
    from gplite_b200.dist import DistComm, dist_energy_grad

    comm = DistComm.from_torch(dist, ctx=ctx) if world > 1 else DistComm.single(ctx)
    mk = lambda: g.gp_init(cfs=g.cf_matern32(lscale=LSCALE), lik=g.lik_gaussian(sigma=SIGMA))

    def tmax(d):
        keys = sorted(d)
        t = torch.tensor([float(d[k]) for k in keys], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return dict(zip(keys, [float(v) for v in t]))

    rec = {"layout": f"1-D block-cyclic by block columns, nb={nb}, block b on rank b % {world}; panels by "
                     "ncclBroadcast inside the library (own comm stream, communicator capped at 4 CTAs), look-ahead "
                     "window of 8 panel buffers, column-chunked panel pipeline in the tail (csrc/dist_sched.hpp)",
           "model": "cfg5: cf_matern32(lscale=1.3) + lik_gaussian(sigma=0.5), d=8, method_full, fp64",
           "n_gpus": world}
    # ---- in-run parity at a check size: N ranks vs the single-GPU path vs the oracle
    n_chk = 6144
    X, y = synth5(n_chk)
    e, gr, alpha = dist_energy_grad(mk(), X, y, jitter=JITTER, include_prior=False, nb=nb, comm=comm,
                                    want_alpha=True)
    par = {"n": n_chk, "energy": e}
    if rank == 0:
        e1, g1 = g.gp_energy_grad(mk(), X, y, jitter=JITTER, include_prior=False, ctx=ctx)
        gp1 = mk()
        g.gp_fit(gp1, X, y, jitter=JITTER, ctx=ctx)
        par.update(energy_rel_diff_vs_single_gpu=abs(e - e1) / abs(e1),
                   grad_rel_diff_vs_single_gpu=float(np.max(np.abs(gr - g1)) / np.max(np.abs(g1))),
                   alpha_scaled_diff_vs_single_gpu=float(np.max(np.abs(alpha - gp1.fit["alpha"]))
                                                         / np.max(np.abs(gp1.fit["alpha"]))))
        try:
            from oracle import gplite_oracle as o
            gpo = mk()
            o.gp_fit(gpo, X, y, jitter=JITTER)
            eo = -gpo.fit["log_evidence"]
            par.update(energy_rel_diff_vs_oracle=abs(e - eo) / abs(eo),
                       alpha_scaled_diff_vs_oracle=float(np.max(np.abs(alpha - gpo.fit["alpha"]))
                                                         / np.max(np.abs(gpo.fit["alpha"]))))
        except Exception as ex:          # the oracle is a checker; its absence must not hide the numbers
            par["oracle_error"] = repr(ex)
    rec["parity"] = par

    def run(n, grad):
        X, y = synth5(n)
        tm = {}
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e, _ = dist_energy_grad(mk(), X, y, jitter=JITTER, include_prior=False, nb=nb, comm=comm, timings=tm,
                                grad=grad)
        t = tmax(tm)
        out = {"n": n, "energy": e, "factor_s": t["factor_s"], "solve_s": t["solve_s"],
               "setup_s": t["kbuild_s"], "total_s": t["total_s"],
               "factor_tflops_aggregate": n ** 3 / 3 / t["factor_s"] / 1e12,
               "factor_tflops_per_gpu": n ** 3 / 3 / t["factor_s"] / 1e12 / world,
               "nccl_gb_per_rank": tm["nccl_bytes"] / 1e9,
               "matrix_gb_per_gpu": (2 if grad else 1) * 8.0 * n * n / world / 1e9}
        if grad:
            out.update(trtri_s=t["trtri_s"], lauum_s=t["lauum_s"], trace_s=t["trace_s"],
                       energy_grad_tflops_aggregate_of_n3=n ** 3 / (t["factor_s"] + t["solve_s"] + t["trtri_s"]
                                                                     + t["lauum_s"] + t["trace_s"]) / 1e12)
        return out

    # ---- strong scaling point: the same n = 100000 factorisation at every N the driver runs.  Two repetitions,
    # the faster one is the record (the first also pays the 80 / N GB slab allocation and NCCL's lazy
    # connection set-up: 2.5 s vs 1.4 s on 8 GPUs, profiles/r02_dist8_n100k_nccl_cta_cap_and_windows.jsonl);
    # both times are kept in the record
    reps = [run(100000, False) for _ in range(2)]
    rec["n100k_factor"] = dict(min(reps, key=lambda r: r["factor_s"]))
    rec["n100k_factor"]["factor_s_each_repetition"] = [r["factor_s"] for r in reps]
    # the N = 1 time of this same code on this pool (profiles/r02_dist_schedule_v2_window_sweep.jsonl), so that a
    # single line already shows the strong-scaling efficiency; the driver's own N = 1 line is the authority
    ref1 = 10.544
    rec["n100k_factor"]["strong_scaling_efficiency_vs_n1_reference"] = ref1 / (world * rec["n100k_factor"]["factor_s"])
    rec["n100k_factor"]["n1_reference_factor_s"] = ref1
    # ---- n = 200000 where it fits (one slab of 320 / N GB per GPU; two for the gradient)
    if world >= 4:
        rec["n200k_factor"] = run(200000, False)
    if world >= 8:
        rec["n200k_energy_grad"] = run(200000, True)
    comm.release()
    comm.close()
    return rec


def parts_record(ctx, torch, n, d, dx):
    """BASELINE metric parts (ii) K-build GB/s and (iii) Cholesky FP64 TFLOP/s at n = 32768, timed with
    CUDA events on the library's stream (x resident in HBM)."""
    import ctypes as C
    import gplite_b200 as g
    from gplite_b200 import _lib
    P = lambda t: C.c_void_p(t.data_ptr())
    ld = n
    K = torch.empty((n, ld), dtype=torch.float64, device="cuda")
    out = {}

    def time_it(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    hbm = None
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if hbm else "fallback 6650 GB/s (B200_PROFILING.md)"
    hbm = hbm or 6650.0
    cases = {"matern52 (cfg2), symmetric": ([g.cf_matern52(lscale=LSCALE)], 1),
             "matern52, rectangular (Kt / Kxz shape)": ([g.cf_matern52(lscale=LSCALE)], 0),
             "sexp*lin (cfg3), symmetric": ([g.cf_sexp(lscale=1.0) * g.cf_lin(prior_magn=g.prior_fixed())], 1),
             "matern32 (cfg5), symmetric": ([g.cf_matern32(lscale=LSCALE)], 1)}
    kb = {}
    for name, (cfs, same) in cases.items():
        factors, _ = g.model.flatten_cfs(cfs)
        pm = _lib.PackedModel(factors)
        ms = time_it(lambda: _lib.check(ctx.lib.gpl_cf_eval_dev(ctx.h, pm.ref(), P(dx), n, P(dx), n, d, same, 0, 0,
                                                                 JITTER if same else 0.0, P(K), ld)), 5)
        gbs = (8.0 * n * n + 8.0 * d * 2 * n) / (ms * 1e-3) / 1e9
        kb[name] = {"ms": ms, "gbs": gbs, "frac_of_hbm_peak": gbs / hbm}
    out["kbuild"] = {"n": n, "algorithmic_bytes": 8.0 * n * n + 8.0 * d * 2 * n, "peak_gbs": hbm,
                     "peak_source": peak_src, "cases": kb}
    out["kbuild_gbs"] = kb["matern52 (cfg2), symmetric"]["gbs"]
    # Cholesky of B = I + K / sigma^2
    factors, _ = g.model.flatten_cfs([g.cf_matern52(lscale=LSCALE)])
    pm = _lib.PackedModel(factors)
    info = C.c_int32()

    def build():
        _lib.check(ctx.lib.gpl_cf_eval_dev(ctx.h, pm.ref(), P(dx), n, P(dx), n, d, 1, 0, 0, JITTER, P(K), ld))
        _lib.check(ctx.lib.gpl_scale_block_dev(ctx.h, n, n, P(K), ld, None, None, 1.0 / SIGMA ** 2, 1.0, 0))

    best = 1e30
    for _ in range(3):
        build()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(ctx.lib.gpl_dpotrf_dev(ctx.h, n, P(K), ld, C.byref(info), None))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["potrf"] = {"n": n, "ms": best, "tflops": n ** 3 / 3 / (best * 1e-3) / 1e12}
    out["potrf_tflops"] = out["potrf"]["tflops"]
    del K
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import gplite_b200 as g
    from gplite_b200 import _lib

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # torch's default stream has handle 0 ("create your own" for gpl_ctx_create): use
    # cudaStreamLegacy (1) so that torch.cuda.Event timing sees the library's kernels
    stream = torch.cuda.current_stream().cuda_stream or 1
    ctx = _lib.Context(local_rank, stream)
    n, d = args.n, D_FULL
    X, y = synth(n, d, seed=2)
    dx = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()      # (d, n) row-major == n x d col-major
    dy = torch.from_numpy(y).cuda()
    hx = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory()
    hy = torch.from_numpy(y).pin_memory()
    Xp = hx.numpy().T                                            # F-ordered view of the pinned buffer
    yp = hy.numpy()

    def model(step):
        # each replica / step evaluates a different point of hyper-parameter space, as gp_optim does
        eps = 0.01 * ((step + 7 * rank) % 5)
        return g.gp_init(cfs=g.cf_matern52(lscale=LSCALE * (1 + eps), magn=MAGN), lik=g.lik_gaussian(sigma=SIGMA))

    def step_dev(i):
        return g.gp_energy_grad(model(i), None, None, jitter=JITTER, include_prior=False, ctx=ctx,
                                device_ptrs=(dx.data_ptr(), dy.data_ptr(), n, d))
    # (dx / dy are re-created after the sharded record; step_dev reads the current bindings)

    def step_host(i):
        return g.gp_energy_grad(model(i), Xp, yp, jitter=JITTER, include_prior=False, ctx=ctx)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record()
        for i in range(steps):
            out = fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx.launch_count() - l0, out

    for i in range(args.warmup):
        step_dev(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches, out = timed(step_dev, args.steps)
    for i in range(max(1, args.warmup // 2)):
        step_host(i)
    ms_e2e, _, _ = timed(step_host, args.steps)
    sampler.stop_flag = True

    value = world * args.steps / (ms * 1e-3)
    e2e = world * args.steps / (ms_e2e * 1e-3)
    nparam = 3
    line = {"metric": "gp_energy+grad evals/sec at n=32k fp64", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"cfg2: n={n}, d={d}, cf_matern52(lscale~{LSCALE}) + lik_gaussian(sigma={SIGMA}), "
                                   f"method_full, jitter={JITTER}: energy + analytic gradient (3 log-params)",
                       "parallelism": "1 GPU" if world == 1 else
                       f"value / e2e: {world} independent replicas of the n=32768 problem (it fits one GPU; no "
                       f"data-path collective); the `sharded` record is the n=100000 / n=200000 problem split "
                       f"across the {world} GPUs over NCCL",
                       "l2": f"working set per step 2 x {8 * n * n / 1e9:.1f} GB >> 126 MB L2, no flush needed"},
            "e2e": {"value": e2e, "unit": "evals/s", "h2d_bytes_per_step": int(8 * n * d + 8 * n),
                    "d2h_bytes_per_step": int(8 * (1 + nparam))},
            "gpu_launches": int(launches),
            "flops_per_step": float(n) ** 3, "tflops_step": float(n) ** 3 / (ms / args.steps * 1e-3) / 1e12}

    parts = None
    if rank == 0:
        try:
            parts = parts_record(ctx, torch, n, d, dx)
        except Exception as ex:
            parts = {"error": repr(ex)}
    del dx, dy
    sharded = None
    if not args.no_sharded:
        # free the replica's cached n x n buffers before the 80 GB slabs are allocated
        torch.cuda.synchronize()
        ctx.release_cache()
        torch.cuda.empty_cache()
        try:
            sharded = sharded_record(ctx, rank, world, dist)
        except Exception as ex:
            sharded = {"error": repr(ex)}
        dx = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
        dy = torch.from_numpy(y).cuda()
    else:
        dx = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
        dy = torch.from_numpy(y).cuda()
    if rank == 0:
        if sharded is not None:
            line["sharded"] = sharded
        if parts is not None:
            line["parts"] = parts
            if "potrf" in parts:
                line["kbuild_gbs"] = parts["kbuild_gbs"]
                line["potrf_tflops"] = parts["potrf_tflops"]
        # profiled pass (not timed): per-launch CUDA events around the DMMA GEMM kernel
        ctx.profile(True)
        psteps = max(1, min(2, args.steps))
        for i in range(psteps):
            step_dev(i)
        pr = ctx.profile_read()
        ctx.profile(False)
        gemm_ms_step = pr["gemm_ms"] / psteps
        achieved = float(n) ** 3 / (gemm_ms_step * 1e-3) / 1e12
        # cuBLAS DGEMM yardstick, same process, burst (best of 5)
        a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        best = 1e9
        for _ in range(6):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            torch.matmul(a, b)
            t1.record()
            torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1))
        peak = 2 * 8192.0 ** 3 / (best * 1e-3) / 1e12
        del a, b
        traffic = None
        try:      # DRAM bytes of all GEMM launches of one step, from the committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_gemm_traffic_egrad32768_v19.json")))
            if n == N_FULL:
                traffic = tj["dram_bytes_total"]
        except Exception:
            pass
        line["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak, "traffic": traffic,
                            "traffic_note": "bytes per step over all launches of the kernel (ncu dram__bytes_read+write, "
                                            "profiles/r01_gemm_traffic_egrad32768_v19.json, launch list profiles/r01_launches_egrad32768_v19.csv); achieved is per step likewise",
                            "kernel": "gemm_f64_kernel (DMMA.8x8x4 FP64 tensor path)",
                            "peak_source": "cuBLAS DGEMM 8192^3 burst measured in this run (no fp64 entry in "
                                           "MEASURED_PEAKS.json); nominal B200 FP64 tensor 37-40 TFLOP/s",
                            "gemm_ms_per_step": gemm_ms_step, "gemm_launches_per_step": pr["gemm_launches"] / psteps,
                            "gemm_share_of_step": gemm_ms_step / (ms / args.steps),
                            "executed_gemm_flops_per_step": pr["gemm_flops"] / psteps}
        line["clocks"] = sampler.summary()
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(d)
            except Exception as ex:
                line["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=N_FULL)
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded (config 5) record")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="development runs: skip the CPU leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__,
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--size", str(args.n)] + (["--no-sharded"] if args.no_sharded else [])
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
