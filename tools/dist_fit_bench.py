"""Multi-GPU full GP (BASELINE config 5) through the gpl_dist_* entries -- launch with torchrun, one
rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_fit_bench.py --size 100000 [--nb 1024,2048] [--grad] [--reps 2]

Prints one JSON line per (nb, repetition) from rank 0: seconds per phase (max over ranks), aggregate
TFLOP/s of the factorisation (n^3/3 flops), NCCL payload per rank.  torch.distributed (gloo) only
carries the 128-byte NCCL id and the max-over-ranks of the timings; the data plane is the library's own
NCCL communicator.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gplite_b200 as g                                             # noqa: E402
from gplite_b200.dist import DistComm, dist_energy_grad, dist_fit_laplace, dist_fit_ep   # noqa: E402


def data(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d))
    y = np.sin(X[:, 0]) + 0.5 * X[:, 1] ** 2 + X[:, 2] + 0.5 * rng.normal(size=n)
    return X, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=65536)
    ap.add_argument("--d", type=int, default=8)
    ap.add_argument("--nb", default="1024")
    ap.add_argument("--kind", default="matern32")
    ap.add_argument("--grad", action="store_true", help="energy + analytic gradient (two more panel sweeps)")
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--bernoulli", action="store_true", help="config 3 style: lik_bernoulli + cf_sexp*cf_lin, Laplace")
    ap.add_argument("--ep", action="store_true", help="with --bernoulli: parallel EP instead of Laplace")
    ap.add_argument("--bcast-test", action="store_true", help="time bare panel broadcasts on the library's communicator")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    comm = DistComm.from_torch(dist, device=local)
    mk = {"matern32": g.cf_matern32, "matern52": g.cf_matern52, "sexp": g.cf_sexp}[args.kind]

    def tmax(tm):
        keys = sorted(tm)
        t = torch.tensor([float(tm[k]) for k in keys], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return dict(zip(keys, [float(v) for v in t]))

    n = args.n
    if args.bcast_test:
        import ctypes as C
        from gplite_b200 import _lib
        for rows in (n, n // 2, n // 4, n // 8):
            sec = C.c_double()
            _lib.check(comm.lib.gpl_dist_bcast_bench(comm.h, rows * 1024, 10, C.byref(sec)))
            if rank == 0:
                print(json.dumps({"bcast_test": True, "n_gpus": world, "panel_rows": rows, "mbytes": rows * 1024 * 8 / 1e6,
                                  "seconds": sec.value, "gbs": rows * 1024 * 8 / sec.value / 1e9}), flush=True)
        comm.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    if args.bernoulli:
        import warnings
        rng = np.random.default_rng(3)
        X = rng.normal(size=(n, 10))
        y = (rng.uniform(size=n) < 1 / (1 + np.exp(-(2 * X[:, 0] - X[:, 1])))).astype(float)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gp = g.gp_init(cfs=g.cf_sexp(lscale=1.0) * g.cf_lin(prior_magn=g.prior_fixed()), lik=g.lik_bernoulli(),
                           approx=g.approx_ep() if args.ep else g.approx_laplace())
    else:
        X, y = data(n, args.d, 5)
        gp = g.gp_init(cfs=mk(lscale=1.3), lik=g.lik_gaussian(sigma=0.5))
    for nb in [int(v) for v in args.nb.split(",")]:
        for rep in range(args.reps):
            tm = {}
            dist.barrier()
            if args.bernoulli:
                res = (dist_fit_ep if args.ep else dist_fit_laplace)(gp, X, y, nb=nb, comm=comm, timings=tm)
                extra = {"iters": res["iters"], "log_evidence": res["log_evidence"]}
                nfac = res["iters"]
            else:
                e, gr = dist_energy_grad(gp, X, y, nb=nb, comm=comm, timings=tm, grad=args.grad)
                extra = {"energy": e, "grad": None if gr is None else [float(v) for v in gr]}
                nfac = 1
            t = tmax(tm)
            if rank == 0:
                out = {"n": n, "n_gpus": world, "nb": nb, "rep": rep,
                       "workload": ("cfg3-style Bernoulli " + ("EP" if args.ep else "Laplace")) if args.bernoulli
                       else f"cfg5-style cf_{args.kind} + lik_gaussian" + (" energy+grad" if args.grad else " energy"),
                       **{k: t[k] for k in ("factor_s", "solve_s", "trtri_s", "lauum_s", "trace_s", "total_s")},
                       "factor_tflops_aggregate": nfac * n ** 3 / 3 / t["factor_s"] / 1e12,
                       "factor_tflops_per_gpu": nfac * n ** 3 / 3 / t["factor_s"] / 1e12 / world,
                       "nccl_gb_per_rank": tm["nccl_bytes"] / 1e9, **extra}
                if args.grad:
                    out["tflops_aggregate_of_n3"] = n ** 3 / (t["factor_s"] + t["solve_s"] + t["trtri_s"] + t["lauum_s"]
                                                              + t["trace_s"]) / 1e12
                print(json.dumps(out), flush=True)
    comm.release()
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
