import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
m = I.config_c2()
eng = Engine(Problem.from_model(m))
xs = eng.solve_batch(m.xl, m.xu, m.x0)["x"][0]
for B in (8192, 16384):
    XL, XU = I.c3_children(m, xs, B)
    eng.solve_batch(XL, XU, xs, want_g=False, want_duals=False)
    t0 = time.perf_counter(); r = eng.solve_batch(XL, XU, xs, want_g=False, want_duals=False); dt = time.perf_counter() - t0
    print(f"C2 raw batch B={B}: {B/dt:.0f} relax/s e2e, device {B/eng.stats()['device_ms']*1e3:.0f}/s, mean iters {r['iters'].mean():.1f}")
