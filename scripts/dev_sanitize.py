"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): cooperative kernels (both builds), the general
kernel, the children expansion kernel and the OA pipeline on tiny models."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
which = sys.argv[1] if len(sys.argv) > 1 else "all"
m = I.syn_miqcp(12, 6, 3, 3, 1)
eng = Engine(Problem.from_model(m))
xs = eng.solve_batch(m.xl, m.xu, m.x0)["x"][0]
XL, XU = I.c3_children(m, xs, 24)
r = eng.solve_batch(XL, XU, xs)
print("coop512", r["status"][:8], r["obj"][:3])
if which == "all":
    XL2, XU2 = I.c3_children(m, xs, 200)            # >= 1.25 x SMs: the 256-thread build
    r2 = eng.solve_batch(XL2, XU2, xs, want_g=False, want_duals=False)
    print("coop256", (r2["status"] == 1).sum())
    cp = np.zeros(8, np.int32); cc = np.arange(8, dtype=np.int32) % 6; cw = np.arange(8, dtype=np.int32) % 2
    cb = np.where(cw == 0, np.floor(xs[cc]), np.ceil(xs[cc]))
    r3 = eng.solve_children(m.xl[None, :], m.xu[None, :], cp, cc, cw, cb, xs)
    print("children", r3["status"])
    c = eng.oa_batch(np.clip(xs + 0.1, m.xl, m.xu)[None, :].repeat(4, 0), None, m.xl, m.xu, get_obj=1)
    print("oa rows", int((c["src_row"] >= 0).sum()))
    g = I.syn_cvx_minlp(16, 6, 3, 5, 2)              # element terms: the general kernel
    e2 = Engine(Problem.from_model(g))
    x2 = e2.solve_batch(g.xl, g.xu, g.x0)
    print("general", x2["status"], x2["obj"])
