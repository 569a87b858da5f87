"""bnlp_solve_batch over several devices of one process (device_ids): contiguous child shards, one stream per device."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
nd = torch.cuda.device_count()
m = I.config_c3()
P = Problem.from_model(m)
e1 = Engine(P, devices=[0])
xs = e1.solve_batch(m.xl, m.xu, m.x0)["x"][0]
B = 4096
XL, XU = I.c3_children(m, xs, B)
a = e1.solve_batch(XL, XU, xs, want_g=False, want_duals=False)
t0 = time.perf_counter(); a = e1.solve_batch(XL, XU, xs, want_g=False, want_duals=False); t1 = time.perf_counter() - t0
eN = Engine(P, devices=list(range(nd)))
b = eN.solve_batch(XL, XU, xs, want_g=False, want_duals=False)
t0 = time.perf_counter(); b = eN.solve_batch(XL, XU, xs, want_g=False, want_duals=False); tN = time.perf_counter() - t0
same = all(np.array_equal(a[k], b[k]) for k in ("obj", "status", "iters", "x"))
print(f"{nd} devices in one process: 1 device {B/t1:.0f} relax/s, {nd} devices {B/tN:.0f} relax/s, identical results: {same}")
