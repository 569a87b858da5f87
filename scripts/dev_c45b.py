"""C4 / C5 relaxation throughput on one GPU (device-timed, repeated), for the large-n path."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
for which, B in (("c5", 592), ("c4", 296)):
    m = I.config_c5() if which == "c5" else I.config_c4()
    eng = Engine(Problem.from_model(m))
    root = eng.solve_batch(m.xl, m.xu, m.x0)
    xs = root["x"][0]
    XL, XU = I.c3_children(m, xs, B)
    for rep in range(3):
        r = eng.solve_batch(XL, XU, xs, want_g=False, want_duals=False)
        ms = eng.stats()["device_ms"]
        print(which, m.name, "B", B, f"device {ms:.1f} ms = {B / ms * 1e3:.0f} relax/s, mean iters {r['iters'].mean():.2f}, statuses",
              dict(zip(*np.unique(r["status"], return_counts=True))), flush=True)
    eng.close()
