"""ctypes front end of oracle/ipm_oracle.cpp (TEST INFRASTRUCTURE ONLY).

`solve_batch(model, XL, XU, x0)` solves B child relaxations of a StructuredModel
(bonmin_b200.instances) on host threads and returns the TMINLP2TNLP result fields
(src/Interfaces/BonTMINLP2TNLP.cpp:474-509)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libipm_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libipm_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def dense_arrays(model):
    """StructuredModel -> (c, A, qidx, Qfull) with every quadratic part densified."""
    n, m = model.n, model.m
    qidx = np.full(m + 1, -1, np.int32)
    mats = []
    il = np.tril_indices(n)
    for k, fn in enumerate(model.qd_fn):
        Q = np.zeros((n, n))
        Q[il] = model.qd[k]
        Q = Q + Q.T - np.diag(np.diag(Q))
        qidx[fn] = len(mats)
        mats.append(Q)
    for fn, i, j, v in zip(model.qs_fn, model.qs_i, model.qs_j, model.qs_v):
        if qidx[fn] < 0:
            qidx[fn] = len(mats)
            mats.append(np.zeros((n, n)))
        Q = mats[qidx[fn]]
        Q[i, j] += v
        if i != j:
            Q[j, i] += v
    Q = np.ascontiguousarray(np.array(mats)) if mats else np.zeros((0, n, n))
    return (np.ascontiguousarray(model.c, dtype=np.float64), np.ascontiguousarray(model.A, dtype=np.float64),
            qidx, Q)


def _p(a, t=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(t))


def solve_batch(model, XL, XU, x0, threads=1, max_iter=0, dense=None, dual0=None):
    """dual0: multipliers [z_L | z_U | lambda] of a warm start (one vector for all children or B rows) or None."""
    n, m = model.n, model.m
    XL = np.ascontiguousarray(np.atleast_2d(XL), dtype=np.float64)
    XU = np.ascontiguousarray(np.atleast_2d(XU), dtype=np.float64)
    B = XL.shape[0]
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    per_child = int(x0.ndim == 2)
    c, A, qidx, Q = dense if dense is not None else dense_arrays(model)
    xl = np.ascontiguousarray(model.xl, dtype=np.float64)
    xu = np.ascontiguousarray(model.xu, dtype=np.float64)
    gl = np.ascontiguousarray(model.gl, dtype=np.float64)
    gu = np.ascontiguousarray(model.gu, dtype=np.float64)
    OBJ = np.zeros(B)
    STATUS = np.zeros(B, np.int32)
    ITERS = np.zeros(B, np.int32)
    P1 = np.zeros(B, np.int32)
    X = np.zeros((B, n))
    G = np.zeros((B, max(m, 1)))[:, :m].copy() if m else np.zeros((B, 0))
    D = np.zeros((B, 2 * n + m))
    ci = ctypes.c_int
    tfn, tvs, tv, tos, top, tia, tda = [], [0], [], [0], [], [], []
    for fn, vars_, prog in model.get("terms", []):
        tfn.append(fn)
        tv.extend(vars_)
        tvs.append(len(tv))
        for op, ia, da in prog:
            top.append(op)
            tia.append(ia)
            tda.append(da)
        tos.append(len(top))
    ti = lambda a: np.ascontiguousarray(a if len(a) else [0], dtype=np.int32)  # noqa: E731
    tfn_a, tvs_a, tv_a, tos_a, top_a, tia_a = ti(tfn), ti(tvs), ti(tv), ti(tos), ti(top), ti(tia)
    tda_a = np.ascontiguousarray(tda if len(tda) else [0.0], dtype=np.float64)
    d0 = None if dual0 is None else np.ascontiguousarray(dual0, dtype=np.float64)
    if d0 is not None and d0.shape[-1] != 2 * n + m:
        raise ValueError("dual0 must have 2n+m entries per row")
    lib().oracle_solve_batch_warm(ci(n), ci(m), _p(xl), _p(xu), _p(gl), _p(gu), _p(c), _p(A), _p(qidx, ci), _p(Q),
                                  ci(len(tfn)), _p(tfn_a, ci), _p(tvs_a, ci), _p(tv_a, ci), _p(tos_a, ci), _p(top_a, ci),
                                  _p(tia_a, ci), _p(tda_a), ci(B), _p(XL), _p(XU), _p(x0), ci(per_child),
                                  None if d0 is None else _p(d0), ci(int(d0 is not None and d0.ndim == 2)), ci(threads),
                                  ci(max_iter), _p(OBJ), _p(STATUS, ci), _p(ITERS, ci), _p(X), _p(G), _p(D), _p(P1, ci))
    return dict(obj=OBJ, status=STATUS, iters=ITERS, x=X, g=G, duals=D, phase1_iters=P1)


def oa_batch(model, ev, X, duals=None, get_obj=1, threads=1, dense=None, as_cuts=True):
    """B x getOuterApproximation on host threads (oracle_oa_batch, the C++ restatement beside struct_eval.outer_approximation).
    ev: oracle.struct_eval.Evaluator of the model (Jacobian structure).  Returns per point the list of cuts
    (src_row, cols, vals, lb, ub) — or, with as_cuts=False, the raw fixed-capacity arrays (the timing baseline of bench.py)."""
    n, m = model.n, model.m
    X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
    B = X.shape[0]
    c, A, qidx, Q = dense if dense is not None else dense_arrays(model)
    arr = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    xl, xu, gl, gu = arr(model.xl), arr(model.xu), arr(model.gl), arr(model.gu)
    ci = ctypes.c_int
    tfn, tvs, tv, tos, top, tia, tda = [], [0], [], [0], [], [], []
    for fn, vars_, prog in model.get("terms", []):
        tfn.append(fn)
        tv.extend(vars_)
        tvs.append(len(tv))
        for op, ia, da in prog:
            top.append(op)
            tia.append(ia)
            tda.append(da)
        tos.append(len(top))
    ti = lambda a: np.ascontiguousarray(a if len(a) else [0], dtype=np.int32)  # noqa: E731
    tfn_a, tvs_a, tv_a, tos_a, top_a, tia_a = ti(tfn), ti(tvs), ti(tv), ti(tos), ti(top), ti(tia)
    tda_a = np.ascontiguousarray(tda if len(tda) else [0.0], dtype=np.float64)
    jac_start = np.zeros(m + 1, np.int32)
    for i, _ in ev.jac:
        jac_start[i + 1] += 1
    jac_start = np.cumsum(jac_start).astype(np.int32)
    jac_col = ti([j for _, j in ev.jac])
    nl_rows = ti([i for i in range(m) if ev.con_nonlinear[i]])
    mnl = int(sum(1 for i in range(m) if ev.con_nonlinear[i]))
    cap = [0]
    for s in range(mnl):
        i = int(nl_rows[s])
        cap.append(cap[-1] + int(jac_start[i + 1] - jac_start[i]))
    cap.append(cap[-1] + n + 1)
    cap_start = ti(cap)
    slots = mnl + 1
    CNT = np.zeros(B * slots, np.int32)
    COLS = np.zeros(B * cap[-1], np.int32)
    VALS = np.zeros(B * cap[-1])
    LB, UB = np.zeros(B * slots), np.zeros(B * slots)
    D = None if duals is None else np.ascontiguousarray(np.atleast_2d(duals), dtype=np.float64)
    if D is not None and D.shape[0] == 1 and B > 1:
        D = np.ascontiguousarray(np.tile(D, (B, 1)))
    L = lib()
    L.oracle_oa_batch.restype = ci
    L.oracle_oa_batch(ci(n), ci(m), _p(xl), _p(xu), _p(gl), _p(gu), _p(c), _p(A), _p(qidx, ci), _p(Q), ci(len(tfn)), _p(tfn_a, ci),
                      _p(tvs_a, ci), _p(tv_a, ci), _p(tos_a, ci), _p(top_a, ci), _p(tia_a, ci), _p(tda_a), _p(jac_start, ci),
                      _p(jac_col, ci), ci(mnl), _p(nl_rows, ci), ci(int(ev.obj_nonlinear)), _p(cap_start, ci), ci(B), _p(X),
                      None if D is None else _p(D), ci(get_obj), ci(threads), _p(CNT, ci), _p(COLS, ci), _p(VALS), _p(LB), _p(UB))
    if not as_cuts:
        return dict(cnt=CNT, cols=COLS, vals=VALS, lb=LB, ub=UB, slots=slots, cap=cap[-1], cap_start=cap_start)
    out = []
    for b in range(B):
        cuts = []
        for s in range(slots):
            k = int(CNT[b * slots + s])
            if k < 0:
                continue
            o = b * cap[-1] + cap[s]
            src = int(nl_rows[s]) if s < mnl else m
            cuts.append((src, COLS[o:o + k].tolist(), VALS[o:o + k].tolist(), float(LB[b * slots + s]), float(UB[b * slots + s])))
        out.append(cuts)
    return out
