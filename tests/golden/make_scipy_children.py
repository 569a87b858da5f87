"""Generates tests/golden/scipy_children_{c2,c3}.npz: an INDEPENDENT solver's verdict on 200 child relaxations each of the
BASELINE config shapes C2 (n=100, m=40) and C3 (n=200, m=80) — SURVEY.md §7 step 1c.  scipy.optimize.minimize
(method="trust-constr", exact gradients / Jacobians / Hessians of the quadratic model) stands in for Ipopt, which cannot be
built offline.  The children are bonmin_b200.instances.c3_children(model, x*, 200) with x* = the oracle's root optimum (the
start point only: what is stored is scipy's own objective and constraint violation per child).

    python tests/golden/make_scipy_children.py [c2|c3] [processes]

Takes minutes (it is NOT run by the test-suite; tests/test_oracle_golden.py::test_oracle_matches_scipy_fixture reads the files)."""
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bonmin_b200 import instances as I  # noqa: E402
from oracle import ipm  # noqa: E402

_G = {}


def _init(which):
    m = I.config_c2() if which == "c2" else I.config_c3()
    c, A, qidx, Q = ipm.dense_arrays(m)
    _G.update(m=m, c=c, A=A, qidx=qidx, Q=Q)


def _solve(args):
    from scipy.optimize import Bounds, NonlinearConstraint, minimize
    xl, xu, x0 = args
    if (xu < xl).any():            # crossed bounds: infeasible before any solver is asked (status -1 = not solved)
        return 0.0, float(np.max(xl - xu)), -1, 0.0, 0
    m, c, A, qidx, Q = (_G[k] for k in ("m", "c", "A", "qidx", "Q"))
    n, mm = m.n, m.m
    Qk = [Q[qidx[k]] if qidx[k] >= 0 else None for k in range(mm + 1)]

    def val(k, x):
        return c[k] + A[k] @ x + (0.5 * x @ Qk[k] @ x if Qk[k] is not None else 0.0)

    def grad(k, x):
        return A[k] + (Qk[k] @ x if Qk[k] is not None else 0.0)

    def g(x):
        return np.array([val(i + 1, x) for i in range(mm)])

    def jac(x):
        return np.array([grad(i + 1, x) for i in range(mm)])

    def hess(x, v):
        H = np.zeros((n, n))
        for i in range(mm):
            if Qk[i + 1] is not None:
                H += v[i] * Qk[i + 1]
        return H
    gl = np.where(np.asarray(m.gl) <= -1e19, -np.inf, m.gl)
    gu = np.where(np.asarray(m.gu) >= 1e19, np.inf, m.gu)
    free = xu > xl                                      # fixed variables are kept out of the barrier (keep_feasible off)
    lo = np.where(free, xl, xl - 0.0)
    res = minimize(lambda x: val(0, x), np.clip(x0, xl, xu), jac=lambda x: grad(0, x),
                   hess=lambda x: (Qk[0] if Qk[0] is not None else np.zeros((n, n))),
                   bounds=Bounds(lo, xu), constraints=[NonlinearConstraint(g, gl, gu, jac=jac, hess=hess)],
                   method="trust-constr", options=dict(gtol=1e-9, xtol=1e-12, barrier_tol=1e-10, maxiter=3000))
    gx = g(res.x)
    viol = max(0.0, float(np.max(gx - gu)), float(np.max(gl - gx)), float(np.max(xl - res.x)), float(np.max(res.x - xu)))
    return float(res.fun), viol, int(res.status), float(res.optimality), int(res.nit)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    _init(which)
    m = _G["m"]
    dense = (_G["c"], _G["A"], _G["qidx"], _G["Q"])
    root = ipm.solve_batch(m, m.xl, m.xu, m.x0, dense=dense)
    xs = root["x"][0]
    B = 200
    XL, XU = I.c3_children(m, xs, B)
    with Pool(procs, initializer=_init, initargs=(which,)) as p:
        out = p.map(_solve, [(XL[b], XU[b], xs) for b in range(B)], chunksize=1)
    fun, viol, st, opt, nit = (np.array(v) for v in zip(*out))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"scipy_children_{which}.npz")
    np.savez_compressed(path, fun=fun, viol=viol, status=st, optimality=opt, nit=nit, x_start=xs,
                        model=np.array(m.name), children=np.array(B))
    print(path, "feasible:", int((viol < 1e-6).sum()), "of", B, "median nit", np.median(nit))


if __name__ == "__main__":
    main()
