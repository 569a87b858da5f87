"""Full-size C4 / C5 on the general kernel: root relaxation + a batch of children, checked against the CPU oracle on a sample."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
from oracle import ipm
which = sys.argv[1] if len(sys.argv) > 1 else "c5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
m = I.config_c5() if which == "c5" else I.config_c4()
t0 = time.perf_counter(); eng = Engine(Problem.from_model(m)); print("engine", time.perf_counter() - t0, "s", m.name, "n", m.n, "m", m.m)
t0 = time.perf_counter(); root = eng.solve_batch(m.xl, m.xu, m.x0); print("root", root["status"], root["obj"], root["iters"], time.perf_counter() - t0, "s")
xs = root["x"][0]
XL, XU = I.c3_children(m, xs, B)
eng.solve_batch(XL[:8], XU[:8], xs)
t0 = time.perf_counter(); r = eng.solve_batch(XL, XU, xs); dt = time.perf_counter() - t0
print(f"{B} children: {dt:.3f} s = {B/dt:.1f} relax/s; device_ms {eng.stats()['device_ms']:.1f}; statuses", np.unique(r["status"], return_counts=True), "mean iters", r["iters"].mean())
S = min(B, 16)
t0 = time.perf_counter(); o = ipm.solve_batch(m, XL[:S], XU[:S], xs, threads=16); dc = time.perf_counter() - t0
ok = o["status"] == 1
print(f"oracle {S} children {dc:.2f} s = {S/dc:.2f} relax/s; status equal", (o["status"] == r["status"][:S]).all(), "max rel obj diff",
      np.max(np.abs(o["obj"][ok] - r["obj"][:S][ok]) / np.maximum(1, np.abs(o["obj"][ok]))) if ok.any() else None)
