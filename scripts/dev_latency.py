"""Per-call latency of a strong-branching batch (64 children) through the C ABI: wall time vs device time."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
for name, m in (("C2 n=100", I.config_c2()), ("n=60", I.syn_miqcp(60, 30, 12, 12, 3)), ("C3 n=200", I.config_c3())):
    eng = Engine(Problem.from_model(m))
    xs = eng.solve_batch(m.xl, m.xu, m.x0)["x"][0]
    for B in (64, 2):
        XL, XU = I.c3_children(m, xs, B)
        for _ in range(3):
            eng.solve_batch(XL, XU, xs)
        t0 = time.perf_counter(); dev = 0.0; K = 50
        for _ in range(K):
            r = eng.solve_batch(XL, XU, xs, want_g=False, want_duals=False)
            dev += eng.stats()["device_ms"]
        wall = (time.perf_counter() - t0) / K * 1e3
        print(f"{name} B={B}: wall {wall:.3f} ms/call, device (copies+kernel) {dev/K:.3f} ms, max iters {r['iters'].max()}, mean iters {r['iters'].mean():.1f}")
    eng.close()
