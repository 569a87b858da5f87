"""BASELINE configs 4 and 5 run AS ALGORITHMS on the engine (host loops of bonmin_b200/oa.py and the C++ B-BB / dive
drivers behind the C ABI), one JSON line per run:
  C4  SYN-CVX-MINLP n=500: B-OA (OACutGenerator2::performOa), B-QG and B-Hyb style single trees (OaFeasibilityChecker /
      OaNlpOptim) — every NLP, the initial LP and every linearisation on the GPU, master MILP / node LPs by HiGHS;
  C5  SYN-PORTFOLIO n=400: root relaxation, the two NLP dives (HeuristicDive fractional / vectorLength), B-BB with 32
      strong-branching candidates per node for a bounded number of nodes; under torchrun the children of every round are
      sharded over the ranks (bnlp_solve_batch_sharded, NCCL inside libbnlp) and every rank must see the same search.
python scripts/run_algorithms.py [--limit SECONDS] [--nodes N]      |      torchrun ... scripts/run_algorithms.py"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import bb, oa, instances as I
from bonmin_b200.engine import Engine, Problem

ap = argparse.ArgumentParser()
ap.add_argument("--limit", type=float, default=120.0)
ap.add_argument("--nodes", type=int, default=200)
ap.add_argument("--skip-c4", action="store_true")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
comm = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

def emit(d):
    if rank == 0:
        print(json.dumps(d), flush=True)

class Timed(oa.EngineProvider):
    """EngineProvider that keeps the wall time spent inside the engine (everything else of a run is the host master / LPs)."""
    engine_s = 0.0

    def _t(self, f, *a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            self.engine_s += time.perf_counter() - t0

    def solve(self, *a, **k): return self._t(super().solve, *a, **k)
    def extract(self, *a, **k): return self._t(super().extract, *a, **k)
    def oa(self, *a, **k): return self._t(super().oa, *a, **k)
    def violation(self, *a, **k): return self._t(super().violation, *a, **k)


for mk, tag, lim in ((lambda: I.syn_cvx_minlp(60, 20, 12, 12, 5), "C4-shaped, reduced", args.limit), (I.config_c4, "C4", args.limit)):
    if args.skip_c4 or rank != 0:
        break
    m = mk()
    eng = Engine(Problem.from_model(m), devices=[lr])
    P = Timed(eng)
    res = {}
    for name, fn in (("B-OA", lambda: oa.run_oa(P, time_limit=lim)),
                     ("B-QG", lambda: oa.run_qg(P, nlp_solves_per_level=0.0, time_limit=lim)),
                     ("B-Hyb", lambda: oa.run_qg(P, nlp_solves_per_level=1.0, time_limit=lim))):
        l0, P.engine_s = eng.launches(), 0.0
        r = fn()
        r.pop("best_x", None)
        r["gpu_launches"] = eng.launches() - l0
        r["seconds_in_engine_calls"] = P.engine_s
        res[name] = r
        emit(dict(config=tag + " " + m.name, algorithm=name, **{k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in r.items()}))
    if tag != "C4":
        b = bb.run_engine(eng, number_strong_branch=16, time_limit=lim)
        res["B-BB"] = b
        emit(dict(config=tag + " " + m.name, algorithm="B-BB", status=b["status"], best_obj=b["best_obj"], nodes=b["nodes"], seconds=b["seconds"]))
    done = [r["best_obj"] for r in res.values() if r["status"] == 0]
    if len(done) > 1:
        emit({"config": tag, "algorithms_finished": len(done), "same_optimum_1e-6": bool(max(done) - min(done) <= 1e-6 * max(1.0, abs(done[0]))), "optima": done})
    eng.close()
if world > 1:
    dist.barrier()

m = I.config_c5()
eng = Engine(Problem.from_model(m), devices=[lr])
t0 = time.perf_counter(); root = eng.solve_batch(m.xl, m.xu, m.x0); t_root = time.perf_counter() - t0
node = dict(x=root["x"][0], g=root["g"][0], obj=root["obj"][0])
dives = {}
best = None
for rule in ("fractional", "vectorLength"):
    t0 = time.perf_counter(); d = bb.dive_engine(eng, node, rule=rule, speculate=True); dt = time.perf_counter() - t0
    dives[rule] = dict(found=bool(d["found"]), obj=float(d["obj"]), steps=int(d["steps"]), nlp_solves=int(d["nlp_solves"]), seconds=dt)
    if d["found"] and (best is None or d["obj"] < best[0]):
        best = (float(d["obj"]), d["x"].copy())
emit(dict(config="C5 " + m.name, root_obj=float(node["obj"]), root_ms=t_root * 1e3, root_iters=int(root["iters"][0]), dives=dives))
# B-Hyb's order: heuristics first, their solution is the incumbent the tree starts from
kw = dict(number_strong_branch=32, number_before_trust=1000000, node_limit=args.nodes, parallel_nodes=8, incumbent=best)
h = bb.run_engine(eng, **kw)
emit(dict(config="C5 " + m.name, algorithm="B-Hyb order: dives, then B-BB from their incumbent (32 strong-branching candidates per node, 8 nodes per round)",
          incumbent_from_dives=(best[0] if best else None), seconds=h["seconds"], nodes=h["nodes"], status=h["status"], best_obj=h["best_obj"],
          best_bound=h["best_bound"]))
kw["incumbent"] = None      # the bounded search WITHOUT the heuristic's cutoff: the batches the ranks share
a = bb.run_engine(eng, **kw)
line = dict(config="C5 " + m.name, algorithm="B-BB without a starting incumbent, 32 strong-branching candidates per node, 8 nodes per round", node_limit=args.nodes,
            one_gpu=dict(seconds=a["seconds"], nodes=a["nodes"], status=a["status"], best_obj=a["best_obj"], best_bound=a["best_bound"], node_nlps=a["nlp_solves"], sb_children=a["sb_children"], ipm_iterations=a["ipm_iterations"]))
if world > 1:
    from bonmin_b200.engine import Comm
    comm = Comm.from_torch(eng, torch.device("cuda", lr))
    dist.barrier()
    b = bb.run_engine_sharded(comm, **kw)
    same = a["best_obj"] == b["best_obj"] and a["nodes"] == b["nodes"] and a["status"] == b["status"]
    flag = torch.tensor([1.0 if same else 0.0], device=torch.device("cuda", lr))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    line["sharded"] = dict(ranks=world, seconds=b["seconds"], nodes=b["nodes"], status=b["status"], best_obj=b["best_obj"],
                           nccl_collectives=comm.stats()["collectives"], same_search_on_every_rank=bool(flag.item() == 1.0))
    comm.close()
emit(line)
eng.close()
if world > 1:
    dist.destroy_process_group()
