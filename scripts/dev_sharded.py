"""Multi-rank check of the engine's NCCL exchange (run under torchrun, one rank per GPU):
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dev_sharded.py
(1) bnlp_solve_batch_sharded returns on EVERY rank exactly what one GPU returns for the whole batch;
(2) bnlp_bb_run_engine_sharded reaches the same optimum, node count and incumbent on every rank as the single-GPU search;
(3) timing of a C5 node batch (64 children) and of the C2 B-BB sharded vs unsharded."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import bb, instances as I
from bonmin_b200.engine import Comm, Engine, Problem

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
for mk, B in ((lambda: I.syn_miqcp(40, 20, 8, 8, 3), 101), (I.config_c3, 1000), (I.config_c5, 64)):
    m = mk()
    eng = Engine(Problem.from_model(m), devices=[lr])
    comm = Comm.from_torch(eng, dev)
    root = eng.solve_batch(m.xl, m.xu, m.x0)
    xs = root["x"][0]
    XL, XU = I.c3_children(m, xs, B)
    one = eng.solve_batch(XL, XU, xs, want_g=False, want_duals=False)
    comm.solve_batch(XL, XU, xs)
    dist.barrier(); t0 = time.perf_counter()
    sh = comm.solve_batch(XL, XU, xs)
    dist.barrier(); ts = time.perf_counter() - t0
    t0 = time.perf_counter(); eng.solve_batch(XL, XU, xs, want_g=False, want_duals=False); t1 = time.perf_counter() - t0
    same = all(np.array_equal(one[k], sh[k]) for k in ("obj", "status", "iters", "x"))
    ok &= same
    if rank == 0:
        print(f"{m.name}: {B} children sharded over {world} ranks == one GPU bit for bit: {same}; {t1 * 1e3:.2f} ms on one GPU, "
              f"{ts * 1e3:.2f} ms sharded (exchange {comm.stats()['exchange_ms'] * 1e3:.0f} us)", flush=True)
    assert comm.allreduce_min(float(rank + 5)) == 5.0
    comm.close(); eng.close()
for mk, kw in ((lambda: I.syn_miqcp(24, 12, 5, 5, 6), dict(number_strong_branch=12, number_before_trust=1000000)),
               (I.config_c2, dict(number_strong_branch=32, number_before_trust=1000000, node_limit=400, parallel_nodes=32))):
    m = mk()
    eng = Engine(Problem.from_model(m), devices=[lr])
    comm = Comm.from_torch(eng, dev)
    a = bb.run_engine(eng, **kw)
    dist.barrier()
    b = bb.run_engine_sharded(comm, **kw)
    same = a["best_obj"] == b["best_obj"] and a["nodes"] == b["nodes"] and a["status"] == b["status"]
    ok &= same
    if rank == 0:
        print(f"{m.name}: B-BB one GPU {a['seconds']:.3f} s, sharded over {world} ranks {b['seconds']:.3f} s; nodes {a['nodes']}/{b['nodes']}, "
              f"optimum {a['best_obj']:.9f}/{b['best_obj']:.9f}, identical: {same}; collectives {comm.stats()['collectives']}", flush=True)
    comm.close(); eng.close()
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED OK" if flag.item() == 1.0 else "SHARDED MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
