import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
m = I.config_c2()
eng = Engine(Problem.from_model(m))
root = eng.solve_batch(m.xl, m.xu, m.x0)
xs, du = root["x"][0], root["duals"][0]
ints = np.flatnonzero(m.vtype != 0)
frac = [j for j in ints if abs(xs[j] - round(xs[j])) > 1e-6][:32]
XL, XU = I.strong_branching_children(m, m.xl, m.xu, xs, frac)
for _ in range(3):
    eng.tqp_solve_batch(xs, du, XL, XU)
t0 = time.perf_counter(); K = 20
for _ in range(K):
    q = eng.tqp_solve_batch(xs, du, XL, XU)
tq = (time.perf_counter() - t0) / K * 1e3
t0 = time.perf_counter()
for _ in range(K):
    r = eng.solve_batch(XL, XU, xs, want_g=False, want_duals=False)
tn = (time.perf_counter() - t0) / K * 1e3
print(f"C2, {len(XL)} children: QP surrogate batch {tq:.2f} ms/call (device {eng.stats()['device_ms']:.2f}), NLP batch {tn:.2f} ms/call; QP iters {q['iters'].mean():.1f} NLP iters {r['iters'].mean():.1f}")
