"""world_size-2 gloo test (CPU) of the N>1 path: contiguous child shards, all-gather of per-child results,
all-reduce(min) of the incumbent.  The arithmetic provider here is the CPU oracle (the product has no CPU path);
what is under test is the host-side sharding/exchange logic that bench.py and a multi-GPU B-BB use."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bonmin_b200 import instances as I, shard
    from oracle import ipm
    m = I.syn_miqcp(16, 8, 3, 3, 21)
    root = ipm.solve_batch(m, m.xl, m.xu, m.x0)
    xs = root["x"][0]
    total = 22
    XL, XU = I.c3_children(m, xs, total)
    lo, hi = shard.shard_range(total, rank, world)
    r = ipm.solve_batch(m, XL[lo:hi], XU[lo:hi], xs)
    obj, st = shard.gather_children(torch.from_numpy(r["obj"]), torch.from_numpy(r["status"]), total)
    mine = r["obj"][r["status"] == 1]
    inc = shard.allreduce_incumbent(float(mine.min()) if len(mine) else float("inf"), "cpu")
    if rank == 0:
        full = ipm.solve_batch(m, XL, XU, xs)
        ok = bool(np.array_equal(obj.numpy(), full["obj"]) and np.array_equal(st.numpy(), full["status"])
                  and inc == full["obj"][full["status"] == 1].min() and lo % 2 == 0)
        out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_gather():
    from oracle import ipm
    ipm.build()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def test_shard_ranges_cover_and_pair():
    from bonmin_b200 import shard
    for total in (1, 2, 7, 64, 4096, 4097):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0] and a[1] % 2 == 0 or a[1] == total
