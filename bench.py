"""Headline benchmark: gp_energy + gradient evaluations / second at n = 32768, d = 8, fp64
(BASELINE.json configs[1]: synthetic regression, cf_matern52 + lik_gaussian, method_full).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--size 32768]

A "step" is ONE complete energy + gradient evaluation of the full GP: K build (+jitter) ->
B = I + K/sigma^2 -> blocked DMMA Cholesky -> alpha, f = K alpha, log evidence -> (K+V)^-1 via
trtri + lauum -> fused gradient trace.  n^3 fp64 flops per step.

* value : evals/s with x, y already resident in HBM (gpl_energy_grad on device pointers).
* e2e   : same metric through the public host API (gp_energy_grad on numpy arrays: H2D of x, y and
          D2H of energy + gradient inside the timed region).
* N > 1 : the problem fits one GPU, so ranks are independent replicas evaluating different
          hyper-parameters (what a parallel simplex / multi-start gp_optim does): weak scaling,
          no data-path collective; torch.distributed only carries the barrier and the max-time.
* roofline: the dominant kernel is the FP64 tensor-core GEMM/SYRK (gemm_f64_kernel, DMMA.8x8x4).
          achieved = algorithmic n^3 flops per step / summed CUDA-event durations of that kernel
          per step (gpl_ctx_profile, recorded on the launching stream in a separate profiled
          pass so the timed steps carry no event overhead); peak = cuBLAS DGEMM measured in this
          same process (MEASURED_PEAKS.json has no fp64 entry; its bf16 figure does not bound an
          fp64 kernel).
* cpu_baseline / --impl reference: the reference path on the host cores -- K by the reference's
          own src/cf.cpp (oracle/_ref, single thread as in the reference), LAPACK by scipy with
          all cores -- on a bounded sample size, extrapolated by its n^2 / n^3 cost model.
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
def cpu_reference_eval(n, d, threads=None):
    """One energy + gradient evaluation by the reference path on the CPU at size n.
    Returns (seconds for the K build, seconds for the dense linear algebra + trace)."""
    import scipy.linalg as sla
    from oracle import gplite_oracle as o
    X, y = synth(n, d)
    t0 = time.perf_counter()
    K = o.cf_native("cf_matern52", X, X, LSCALE, MAGN, use_ref=o.ref_available())    # src/cf.cpp:86-119
    tk = time.perf_counter() - t0
    t0 = time.perf_counter()
    K[np.diag_indices(n)] += JITTER
    V = SIGMA ** 2
    B = np.eye(n) + K / V                                      # R/laplace.R:27
    L = sla.cholesky(B, lower=True, check_finite=False)        # R/laplace.R:28 (dpotrf)
    t = sla.solve_triangular(L, y / np.sqrt(V), lower=True, check_finite=False)
    alpha = sla.solve_triangular(L, t, lower=True, trans="T", check_finite=False) / np.sqrt(V)
    f = K @ alpha
    le = (-0.5 * n * np.log(2 * np.pi * V) - 0.5 * np.sum((y - f) ** 2) / V - 0.5 * f @ alpha
          - np.sum(np.log(np.diag(L))))
    Binv, info = sla.lapack.dpotri(L, lower=1)                 # gradient: (K+V)^-1 = Binv / V
    W = np.outer(alpha, alpha) - np.tril(Binv) / V
    # trace against dK/dlog(lscale) and dK/dlog(magn), lower triangle doubled
    r2 = -2.0 * (X @ X.T) + np.sum(X * X, 1)[:, None] + np.sum(X * X, 1)[None, :]
    np.maximum(r2, 0, out=r2)
    tt = np.sqrt(5.0 * r2) / LSCALE
    dK_l = (MAGN ** 2) * (tt * tt / 3.0) * (1 + tt) * np.exp(-tt)
    Wl = np.tril(W, -1)
    g_l = -0.5 * (2 * np.sum(Wl * dK_l))
    g_m = -0.5 * (2 * np.sum(Wl * 2 * K) + np.sum(np.diag(W) * 2 * (np.diag(K) - JITTER)))
    g_s = -0.5 * 2 * V * np.trace(W)
    tl = time.perf_counter() - t0
    return tk, tl, -le, np.array([g_l, g_m, g_s])


def cpu_baseline(sample_n, d):
    tk, tl, _, _ = cpu_reference_eval(sample_n, d)
    scale = N_FULL / sample_n
    t_full = tk * scale ** 2 + tl * scale ** 3
    from oracle import gplite_oracle as o
    return {"value": 1.0 / t_full, "unit": "evals/s", "cores": os.cpu_count(),
            "kind": "reference" if o.ref_available() else "port",
            "sample": (f"one energy+grad evaluation at n={sample_n}, d={d} (K build {tk:.2f} s on 1 thread "
                       f"by the reference's src/cf.cpp, LAPACK/BLAS part {tl:.2f} s on {os.cpu_count()} threads), "
                       f"extrapolated to n={N_FULL} as n^2 (K build) + n^3 (factorisation, inverse, trace)")}


def run_reference(args, rank, world):
    if rank != 0:
        return
    vals = []
    sample_n = args.cpu_n
    for i in range(args.warmup + args.steps):
        tk, tl, _, _ = cpu_reference_eval(sample_n, D_FULL)
        scale = args.n / sample_n
        if i >= args.warmup:
            vals.append(tk * scale ** 2 + tl * scale ** 3)
    t = float(np.mean(vals))
    from oracle import gplite_oracle as o
    val = 1.0 / t
    line = {"impl": "reference", "metric": "gp_energy+grad evals/sec at n=32k fp64", "value": val,
            "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg2: n={args.n}, d={D_FULL}, cf_matern52 + lik_gaussian, method_full, "
                                   "energy + analytic gradient (reference path on host cores)"},
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": os.cpu_count(),
                             "kind": "reference" if o.ref_available() else "port",
                             "sample": f"each step = one evaluation at n={sample_n} (K build by the reference's "
                                       f"src/cf.cpp on 1 thread, LAPACK on {os.cpu_count()} threads), "
                                       f"extrapolated to n={args.n} by n^2 / n^3"},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- GPU arm
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
                       "parallelism": "1 GPU" if world == 1 else f"{world} independent replicas (no collective)",
                       "l2": f"working set per step 2 x {8 * n * n / 1e9:.1f} GB >> 126 MB L2, no flush needed"},
            "e2e": {"value": e2e, "unit": "evals/s", "h2d_bytes_per_step": int(8 * n * d + 8 * n),
                    "d2h_bytes_per_step": int(8 * (1 + nparam))},
            "gpu_launches": int(launches),
            "flops_per_step": float(n) ** 3, "tflops_step": float(n) ** 3 / (ms / args.steps * 1e-3) / 1e12}

    if rank == 0:
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
        if world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(args.cpu_n, d)
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
    ap.add_argument("--cpu-n", type=int, default=5120, help="sample size of the CPU reference arm")
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
               "--size", str(args.n), "--cpu-n", str(args.cpu_n)]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
