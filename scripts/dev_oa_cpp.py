"""B-OA through the C++ loop (bnlp_oa_run_engine) beside the Python loop (oa.run_oa) on a C4-shaped instance and on C4
itself with a bounded number of passes: same passes / cuts / optimum, time outside the master compared."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bonmin_b200 import oa, instances as I
from bonmin_b200.engine import Engine, Problem

KEYS = ("status", "best_obj", "lower_bound", "passes", "nlp_solves", "cuts", "oa_points")


class Timed(oa.EngineProvider):
    t = 0.0
    def _t(self, f, *a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); self.t += time.perf_counter() - t0; return r
    def solve(self, *a, **k): return self._t(super().solve, *a, **k)
    def extract(self, *a, **k): return self._t(super().extract, *a, **k)
    def oa(self, *a, **k): return self._t(super().oa, *a, **k)


for name, m, passes in (("C4-shaped SYN-CVX-MINLP(60,24,10,10,5)", I.syn_cvx_minlp(60, 24, 10, 10, 5), 200),
                        ("C4 SYN-CVX-MINLP(500,250,100,100,20261021)", I.config_c4(), int(os.environ.get("C4_PASSES", "2")))):
    eng = Engine(Problem.from_model(m))
    eng.solve_batch(m.xl, m.xu, m.x0)      # warm-up (engine buffers)
    t0 = time.perf_counter(); g = oa.run_oa_engine(eng, max_passes=passes); tg = time.perf_counter() - t0
    if m.n >= 500 and os.environ.get("C4_SKIP_PY"):      # the master takes minutes per MILP at this size: C++ loop only
        print(json.dumps({"instance": name, "max_passes": passes,
                          "cpp_loop": {**{k: g[k] for k in KEYS}, "wall_s": tg, "master_s": g["master_seconds"],
                                       "outside_master_s": g["seconds"] - g["master_seconds"]}}), flush=True)
        eng.close()
        continue
    P = Timed(eng)
    t0 = time.perf_counter(); p = oa.run_oa(P, max_passes=passes); tp = time.perf_counter() - t0
    print(json.dumps({"instance": name, "max_passes": passes,
                      "cpp_loop": {**{k: g[k] for k in KEYS}, "wall_s": tg, "master_s": g["master_seconds"],
                                   "outside_master_s": g["seconds"] - g["master_seconds"]},
                      "python_loop": {**{k: p[k] for k in KEYS}, "wall_s": tp, "engine_calls_s": P.t},
                      "same_search": all(g[k] == p[k] for k in KEYS)}), flush=True)
    eng.close()
