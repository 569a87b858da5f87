"""C4 / C5 throughput against the number of persistent CTAs: 148 tile-major matrices of 1 MB (n = 500) do not fit the 126 MB L2."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
for which, B in (("c4", 296), ("c5", 592)):
    m = I.config_c4() if which == "c4" else I.config_c5()
    for ctas in (148, 132, 116, 100, 84, 74):
        eng = Engine(Problem.from_model(m), ctas_per_device=ctas)
        root = eng.solve_batch(m.xl, m.xu, m.x0)
        xs = root["x"][0]
        XL, XU = I.c3_children(m, xs, B)
        best = 1e30
        for rep in range(2):
            eng.solve_batch(XL, XU, xs, want_g=False, want_duals=False)
            best = min(best, eng.stats()["device_ms"])
        print(which, "ctas", ctas, f"device {best:.1f} ms = {B / best * 1e3:.0f} relax/s", flush=True)
        eng.close()
