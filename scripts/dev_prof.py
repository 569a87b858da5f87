import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
m = I.config_c3()
eng = Engine(Problem.from_model(m))
xs = eng.solve_batch(m.xl, m.xu, m.x0)["x"][0]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
XL, XU = I.c3_children(m, xs, B)
eng.solve_batch(XL, XU, xs)
r = eng.solve_batch(XL, XU, xs)
print(eng.stats(), B / eng.stats()["device_ms"] * 1e3, "relax/s")
