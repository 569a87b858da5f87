"""Batch size from which the 256-thread 2-CTA/SM kernel beats the 512-thread one at C2 (BNLP_COOP256_MIN_B overrides the switch)."""
import os, subprocess, sys
code = r'''
import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
m = I.config_c2()
eng = Engine(Problem.from_model(m))
xs = eng.solve_batch(m.xl, m.xu, m.x0)["x"][0]
for B in (128, 149, 200, 256, 296, 400, 592):
    XL, XU = I.c3_children(m, xs, B)
    eng.solve_batch(XL, XU, xs, want_g=False, want_duals=False)
    ms = []
    for _ in range(5):
        eng.solve_batch(XL, XU, xs, want_g=False, want_duals=False); ms.append(eng.stats()["device_ms"])
    print(B, round(min(ms), 3), end=" | ")
print()
'''
for thr in ("100000", "1"):
    env = dict(os.environ, BNLP_COOP256_MIN_B=thr)
    print("512-thread only:" if thr == "100000" else "256-thread always:", flush=True)
    subprocess.run([sys.executable, "-c", code], env=env)
