"""world_size-2 test of the partition-sharded path on CPU (gloo): each rank computes its share of the partitions, the fold
scores are summed with one all-reduce, and the result equals the single-rank run.  The per-rank compute here is the CPU
oracle standing in for the GPU (the GPU equivalent is tests/test_gpu_parity.py::test_two_rank_shares_on_one_gpu); what is
exercised is the rank->partition chunking rule and the reduce."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def chunk(nParts, world, rank):
    """src/MPIHelper.cpp:29-37: nParts/world (+1 for the first nParts%world ranks), contiguous"""
    cnt = [nParts // world + (1 if (nParts % world) > r else 0) for r in range(world)]
    first = sum(cnt[:rank])
    return first, first + cnt[rank]


def _unit_worker(rank, world, port, q):
    """unit sharding (the default of the GPU path): contiguous ranges of the nFolds x nParts units, ONE all-reduce of the score arrays"""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_case
    from oracle import pyoracle as O
    from parsmurf_b200 import host_api as H
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = make_case("fy_26")
    x = O.make_x(c["X"], c["y"])
    rng = [H.part_range(True, c["nF"], c["nParts"], world, rank, f) for f in range(c["nF"])]
    rc, c1, c2 = O.o_cv_ranges(x, c["y"], c["f"], c["nF"], c["nParts"], c["fp"], c["ratio"], c["k"], c["T"], c["mtry"], 7, [a for a, b in rng], [b for a, b in rng])
    t = torch.from_numpy(np.stack([c1, c2]))
    dist.all_reduce(t)                                                  # the one cross-rank sum of a unit-sharded CV
    if rank == 0:
        q.put(t.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_case
    from oracle import pyoracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = make_case("fy_26")
    x = O.make_x(c["X"], c["y"])
    p0, p1 = chunk(c["nParts"], world, rank)
    rc, c1, c2 = O.o_cv_parts(x, c["y"], c["f"], c["nF"], c["nParts"], c["fp"], c["ratio"], c["k"], c["T"], c["mtry"], 7, p0, p1)
    t = torch.from_numpy(np.stack([c1, c2]))
    dist.all_reduce(t)
    if rank == 0:
        q.put(t.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_chunks_cover_all_partitions():
    for nParts in (1, 7, 10, 100, 1000):
        for world in (1, 2, 3, 4, 8):
            r = [chunk(nParts, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nParts and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes


def test_two_rank_gloo_allreduce_equals_single_rank():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_case
    from oracle import pyoracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = make_case("fy_26")
    x = O.make_x(c["X"], c["y"])
    rc, c1, c2 = O.o_cv(x, c["y"], c["f"], c["nF"], c["nParts"], c["fp"], c["ratio"], c["k"], c["T"], c["mtry"], 7)
    assert np.allclose(got[0], c1, rtol=0, atol=1e-15) and np.allclose(got[1], c2, rtol=0, atol=1e-15)


def test_three_rank_gloo_unit_sharding_equals_single_rank():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_case
    from oracle import pyoracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_unit_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = make_case("fy_26")
    x = O.make_x(c["X"], c["y"])
    rc, c1, c2 = O.o_cv(x, c["y"], c["f"], c["nF"], c["nParts"], c["fp"], c["ratio"], c["k"], c["T"], c["mtry"], 7)
    assert np.allclose(got[0], c1, rtol=0, atol=1e-15) and np.allclose(got[1], c2, rtol=0, atol=1e-15)
