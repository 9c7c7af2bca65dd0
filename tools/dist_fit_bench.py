"""Multi-GPU full GP (BASELINE config 5) -- launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_fit_bench.py --size 200000 [--nb 1024] [--check]

--check : first run n = 6000 both ways (distributed vs the single-GPU gp_fit) and compare.
Prints one JSON line from rank 0: seconds per phase (max over ranks), aggregate TFLOP/s of the
factorisation (n^3/3 flops), log evidence.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gplite_b200 as g                                 # noqa: E402
from gplite_b200.dist import dist_fit_gaussian          # noqa: E402


def data(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d))
    y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + X[:, 2] + 0.5 * rng.normal(size=n)
    return X, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=65536)
    ap.add_argument("--d", type=int, default=8)
    ap.add_argument("--nb", type=int, default=1024)
    ap.add_argument("--kind", default="matern32")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--fitc", action="store_true", help="config 4: row-sharded FITC instead of the full GP")
    ap.add_argument("--inducing", type=int, default=4096)
    ap.add_argument("--grad", action="store_true",
                    help="energy + analytic gradient (distributed inverse: two more panel sweeps + block trace)")
    ap.add_argument("--ep", action="store_true", help="with --bernoulli: parallel EP instead of Laplace")
    ap.add_argument("--bernoulli", action="store_true",
                    help="config 3 style: lik_bernoulli + cf_sexp*cf_lin, Laplace Newton loop over the sharded factor")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mk = {"matern32": g.cf_matern32, "matern52": g.cf_matern52, "sexp": g.cf_sexp}[args.kind]

    if args.check and not args.grad:
        X, y = data(6000, args.d, 5)
        gp = g.gp_init(cfs=mk(lscale=1.3), lik=g.lik_gaussian(sigma=0.5))
        res = dist_fit_gaussian(gp, X, y, nb=256)
        out = {"check_n": 6000, "dist_log_evidence": res["log_evidence"]}
        if rank == 0:
            gp1 = g.gp_init(cfs=mk(lscale=1.3), lik=g.lik_gaussian(sigma=0.5))
            g.gp_fit(gp1, X, y)
            le1 = gp1.fit["log_evidence"]
            a1 = gp1.fit["alpha"]
            out.update(single_log_evidence=le1, rel_diff=abs(res["log_evidence"] - le1) / abs(le1),
                       alpha_scaled_err=float(np.max(np.abs(res["alpha"].cpu().numpy() - a1)) / np.max(np.abs(a1))))
            print(json.dumps(out), flush=True)
        dist.barrier()

    if args.fitc:
        from gplite_b200.dist import dist_fit_fitc_gaussian
        rng = np.random.default_rng(4)
        n, m, d = args.n, args.inducing, 6
        X = rng.uniform(-3, 3, size=(n, d))
        y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.5 * rng.normal(size=n)
        Zi = X[rng.permutation(n)[:m]].copy()
        gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0), lik=g.lik_gaussian(sigma=0.5), method=g.method_fitc(inducing=Zi))
        for rep in range(2):                      # first pass warms the allocator / NCCL
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            res = dist_fit_fitc_gaussian(gp, X, y)
            torch.cuda.synchronize()
            tsec = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(tsec, op=dist.ReduceOp.MAX)
        if rank == 0:
            flops = 2.0 * n * m * m + 2 * m ** 3 / 3
            print(json.dumps({"workload": f"cfg4: FITC n={n}, m={m}, d={d}, cf_sexp + lik_gaussian, rows sharded over "
                                          f"{world} GPU(s), one {8.0 * m * m / 1e6:.0f} MB all-reduce",
                              "n_gpus": world, "seconds": float(tsec), "log_evidence": res["log_evidence"],
                              "algorithmic_tflops_aggregate": flops / float(tsec) / 1e12}), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return

    if args.bernoulli:
        import warnings
        from gplite_b200.dist import dist_fit_laplace
        rng = np.random.default_rng(3)
        n, d = args.n, 10
        X = rng.normal(size=(n, d))
        y = (rng.uniform(size=n) < 1 / (1 + np.exp(-(2 * X[:, 0] - X[:, 1])))).astype(float)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0) * g.cf_lin(prior_magn=g.prior_fixed()),
                           lik=g.lik_bernoulli(), approx=g.approx_laplace(maxiter=30, tol=1e-4))
        tm = {}
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        if args.ep:
            from gplite_b200.dist import dist_fit_ep
            gp.approx = g.approx_ep(damping=0.9, quad_order=11, maxiter=100)
            res = dist_fit_ep(gp, X, y, nb=args.nb, timings=tm)
        else:
            res = dist_fit_laplace(gp, X, y, nb=args.nb, timings=tm)
        torch.cuda.synchronize()
        vals = torch.tensor([tm["kbuild"], tm["factor"], tm["solve"] + tm.get("trtri", 0.0) + tm.get("diag", 0.0),
                             time.perf_counter() - t0],
                            dtype=torch.float64, device="cuda")
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        if rank == 0:
            kb, fa, so, tot = [float(v) for v in vals]
            print(json.dumps({"workload": f"cfg3-style: n={n}, d={d}, cf_sexp*cf_lin + lik_bernoulli, "
                                          f"{'EP (solve_s includes the X = L^-1 sweeps)' if args.ep else 'Laplace'}, "
                                          f"1-D block-cyclic (nb={args.nb}) over {world} GPU(s)",
                              "n_gpus": world, "newton_iters": res["iters"], "kbuild_s": kb, "factor_s": fa,
                              "solve_s": so, "total_s": tot,
                              "factor_tflops_aggregate": res["iters"] * n ** 3 / 3 / fa / 1e12,
                              "log_evidence": res["log_evidence"]}), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return

    X, y = data(args.n, args.d, 5)
    gp = g.gp_init(cfs=mk(lscale=1.3), lik=g.lik_gaussian(sigma=0.5))
    if args.grad:
        from gplite_b200.dist import dist_energy_grad
        tm = {}
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        e, gr = dist_energy_grad(gp, X, y, nb=args.nb, timings=tm)
        torch.cuda.synchronize()
        vals = torch.tensor([tm["kbuild"], tm["factor"], tm["solve"], tm["trtri"], tm["lauum"], tm["trace"],
                             time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        single = {}
        if args.check and rank == 0 and args.n <= 40000:      # same evaluation on one GPU
            e1, g1 = g.gp_energy_grad(g.gp_init(cfs=mk(lscale=1.3), lik=g.lik_gaussian(sigma=0.5)), X, y)
            single = {"single_gpu_energy": e1, "energy_rel_diff": abs(e - e1) / abs(e1),
                      "grad_max_rel_diff": float(np.max(np.abs(gr - g1)) / np.max(np.abs(g1)))}
        if rank == 0:
            kb, fa, so, ti, la, tr, tot = [float(v) for v in vals]
            n = args.n
            print(json.dumps({"workload": f"cfg5-style energy+gradient: full GP n={n}, d={args.d}, cf_{args.kind} + "
                                          f"lik_gaussian, 1-D block-cyclic (nb={args.nb}) over {world} GPU(s)",
                              "n_gpus": world, "kbuild_s": kb, "factor_s": fa, "solve_s": so, "trtri_s": ti,
                              "lauum_s": la, "trace_s": tr, "total_s": tot,
                              "evals_per_s": 1.0 / tot, "tflops_aggregate_of_n3": n ** 3 / tot / 1e12,
                              "energy": e, "grad": [float(v) for v in gr],
                              "matrix_gb_per_gpu": 2 * 8.0 * n * n / world / 1e9, **single}), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return
    tm = {}
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    res = dist_fit_gaussian(gp, X, y, nb=args.nb, timings=tm)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    vals = torch.tensor([tm["kbuild"], tm["factor"], tm["solve"], total], dtype=torch.float64, device="cuda")
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    if rank == 0:
        kb, fa, so, tot = [float(v) for v in vals]
        n = args.n
        print(json.dumps({"workload": f"cfg5-style: full GP n={n}, d={args.d}, cf_{args.kind} + lik_gaussian, "
                                      f"1-D block-cyclic (nb={args.nb}) over {world} GPU(s)",
                          "n_gpus": world, "kbuild_s": kb, "factor_s": fa, "solve_s": so, "total_s": tot,
                          "factor_tflops_aggregate": n ** 3 / 3 / fa / 1e12,
                          "factor_tflops_per_gpu": n ** 3 / 3 / fa / 1e12 / world,
                          "kbuild_gbs_aggregate": 8.0 * n * n / 2 / kb / 1e9,
                          "log_evidence": res["log_evidence"],
                          "matrix_gb_per_gpu": 8.0 * n * n / world / 1e9}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
