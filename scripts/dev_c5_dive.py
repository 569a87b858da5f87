"""C5 (cardinality-constrained portfolio, n = 400): root relaxation, both NLP dives, one strong-branching node, on the engine."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import bb, instances as I
from bonmin_b200.engine import Problem, Engine
m = I.config_c5()
eng = Engine(Problem.from_model(m))
t0 = time.perf_counter(); root = eng.solve_batch(m.xl, m.xu, m.x0); t_root = time.perf_counter() - t0
node = dict(x=root["x"][0], g=root["g"][0], obj=root["obj"][0])
ints = np.flatnonzero(np.asarray(m.vtype) != 0)
frac = [j for j in ints if abs(node["x"][j] - round(node["x"][j])) > 1e-6]
print(f"root: status {root['status'][0]} obj {node['obj']:.6f} iters {root['iters'][0]} in {t_root*1e3:.1f} ms; {len(frac)} fractional of {len(ints)} integers")
for rule in ("fractional", "vectorLength"):
    for spec in (True, False):
        t0 = time.perf_counter(); d = bb.dive_engine(eng, node, rule=rule, speculate=spec); dt = time.perf_counter() - t0
        print(f"dive {rule:12s} speculate={spec}: found {d['found']} obj {d['obj']:.6f} steps {d['steps']} nlps {d['nlp_solves']} batches {d['batches']} in {dt:.2f} s")
cand = frac[:32]
XL, XU = I.strong_branching_children(m, m.xl, m.xu, node["x"], cand)
t0 = time.perf_counter(); r = eng.solve_batch(XL, XU, node["x"]); dt = time.perf_counter() - t0
print(f"strong branching: {len(cand)} candidates = {len(XL)} children in {dt*1e3:.1f} ms; statuses", np.unique(r["status"], return_counts=True))
