"""CPU, world_size 2 over gloo: the only host-side step of the multi-GPU paths -- the NCCL-id rendezvous of
DistComm (rank 0 draws the id inside the library, a process group carries its 128 bytes).  Everything else
is C++ inside the library: the sharded full-GP schedule is exercised on CPUs by
tests/test_dist_sched_cpu.py (host back end, gloo collectives), the NCCL data plane by
tests/test_gpu_dist.py::test_nccl_two_ranks_equal_single_gpu and bench.py's `sharded.parity` record."""
import copy
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# ------------------------------------------------------------------------------ NCCL id rendezvous
def _id_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gplite_b200.dist import DistComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, DistComm.exchange_id(dist)))
    finally:
        dist.destroy_process_group()


def test_nccl_id_reaches_every_rank():
    """rank 0 draws the id inside the library (gpl_dist_unique_id -> ncclGetUniqueId, no GPU needed),
    the process group only carries its 128 bytes"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_id_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(out[0]) == 128 and out[0] == out[1] and any(out[0])
