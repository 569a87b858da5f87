"""ctypes front end of the host B-BB driver (bonmin_b200/host).  `run_engine` is the product path (C++ driver calling
the CUDA engine directly); `run_callback` runs the same driver with any Python batch solver (tests use the CPU oracle
to check that both arithmetic providers reach the same MINLP optimum)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import engine as _e

c_int, c_dbl, c_void = ctypes.c_int, ctypes.c_double, ctypes.c_void_p
P_int, P_dbl = ctypes.POINTER(c_int), ctypes.POINTER(c_dbl)


class BbOptions(ctypes.Structure):
    _fields_ = [("number_strong_branch", c_int), ("number_before_trust", c_int), ("node_limit", c_int),
                ("log_level", c_int), ("integer_tolerance", c_dbl), ("cutoff_decr", c_dbl), ("time_limit", c_dbl),
                ("parallel_nodes", c_int), ("variable_selection", c_int), ("priorities", ctypes.POINTER(c_int)),
                ("nlp_failure_behavior", c_int), ("number_look_ahead", c_int), ("incumbent_x", ctypes.POINTER(c_dbl)),
                ("incumbent_obj", c_dbl)]


class BbResult(ctypes.Structure):
    _fields_ = [("status", c_int), ("nodes", c_int), ("nlp_solves", c_int), ("sb_batches", c_int),
                ("sb_children", c_int), ("served_from_sb_cache", c_int), ("ipm_iterations", ctypes.c_longlong),
                ("best_obj", c_dbl), ("best_bound", c_dbl), ("seconds", c_dbl), ("failed_nodes", c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


CB = ctypes.CFUNCTYPE(c_int, c_void, c_int, P_dbl, P_dbl, P_dbl, P_dbl, P_int, P_int, P_dbl, P_dbl, P_dbl)
_host = None


def host_lib():
    global _host
    if _host is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbnlp_host.so")
        if not os.path.exists(path):
            from . import build
            build.build_host()
        _host = ctypes.CDLL(path)
    return _host


QCB = ctypes.CFUNCTYPE(c_int, c_void, P_dbl, P_dbl, c_int, P_dbl, P_dbl, P_dbl, P_int, P_int, P_dbl)
VARIABLE_SELECTION = {"strong-branching": 0, "qp-strong-branching": 1}


def _opts(number_strong_branch=0, number_before_trust=0, node_limit=0, time_limit=0.0, parallel_nodes=0,
          variable_selection="strong-branching", priorities=None, nlp_failure_behavior="stop", number_look_ahead=0,
          incumbent=None):
    o = BbOptions(number_strong_branch, number_before_trust, node_limit, 0, 0.0, 0.0, time_limit, parallel_nodes,
                  VARIABLE_SELECTION[variable_selection])
    o.nlp_failure_behavior = {"stop": 0, "fathom": 1}[nlp_failure_behavior]
    o.number_look_ahead = int(number_look_ahead)
    if incumbent is not None:       # (objective, x) of a heuristic's solution: the search starts from its cutoff
        o._inc = np.ascontiguousarray(incumbent[1], float)
        o.incumbent_x = o._inc.ctypes.data_as(ctypes.POINTER(c_dbl))
        o.incumbent_obj = float(incumbent[0])
    if priorities is not None:      # n ints, the LOWEST value is branched on first (kept alive on the struct)
        o._pri = np.ascontiguousarray(priorities, np.int32)
        o.priorities = o._pri.ctypes.data_as(ctypes.POINTER(c_int))
    return o


def run_engine(eng: "_e.Engine", **kw):
    res = BbResult()
    x = np.zeros(eng.n)
    L = _e.lib()
    L.bnlp_bb_run_engine.argtypes = [c_void, c_void, ctypes.POINTER(BbOptions), ctypes.POINTER(BbResult), c_void]
    o = _opts(**kw)
    rc = L.bnlp_bb_run_engine(eng.h, eng.problem.h, ctypes.byref(o), ctypes.byref(res), x.ctypes.data_as(c_void))
    if rc:
        raise _e.BnlpError("bnlp_bb_run_engine failed")
    d = res.as_dict()
    d["best_x"] = x
    return d


def run_engine_sharded(comm: "_e.Comm", **kw):
    """bnlp_bb_run_engine_sharded: the same search on every rank, every batch of children split over the ranks (NCCL)."""
    eng = comm.engine
    res = BbResult()
    x = np.zeros(eng.n)
    L = _e.lib()
    L.bnlp_bb_run_engine_sharded.argtypes = [c_void, c_void, ctypes.POINTER(BbOptions), ctypes.POINTER(BbResult), c_void]
    o = _opts(**kw)
    rc = L.bnlp_bb_run_engine_sharded(comm.h, eng.problem.h, ctypes.byref(o), ctypes.byref(res), x.ctypes.data_as(c_void))
    if rc:
        raise _e.BnlpError("bnlp_bb_run_engine_sharded failed: " + L.bnlp_last_error().decode())
    d = res.as_dict()
    d["best_x"] = x
    return d


def run_callback(model, solve_batch, tqp_solve=None, **kw):
    """solve_batch(XL, XU, x0) -> dict(obj, status, iters, x) with the meaning of Engine.solve_batch.
    tqp_solve(x_sol, duals_sol, XL, XU) -> dict(obj, status, iters, x) with the meaning of Engine.tqp_solve_batch
    (needed for variable_selection="qp-strong-branching")."""
    n, m = model.n, model.m

    def qcb(ctx, XS, DU, B, XL, XU, OBJ, ST, IT, X):
        xs = np.ctypeslib.as_array(XS, (n,)).copy()
        du = np.ctypeslib.as_array(DU, (2 * n + m,)).copy()
        xl = np.ctypeslib.as_array(XL, (B, n)).copy()
        xu = np.ctypeslib.as_array(XU, (B, n)).copy()
        r = tqp_solve(xs, du, xl, xu)
        np.ctypeslib.as_array(OBJ, (B,))[:] = r["obj"]
        np.ctypeslib.as_array(ST, (B,))[:] = r["status"]
        np.ctypeslib.as_array(IT, (B,))[:] = r["iters"]
        if X:
            np.ctypeslib.as_array(X, (B, n))[:] = r["x"]
        return 0

    failure = []

    def cb(ctx, B, XL, XU, X0, OBJ, ST, IT, X, G, D):
        xl = np.ctypeslib.as_array(XL, (B, n)).copy()
        xu = np.ctypeslib.as_array(XU, (B, n)).copy()
        x0 = np.ctypeslib.as_array(X0, (n,)).copy()
        try:
            r = solve_batch(xl, xu, x0)
        except Exception as e:      # a provider failure must reach the driver as an error code (and the caller as an exception)
            failure.append(e)
            return -1
        np.ctypeslib.as_array(OBJ, (B,))[:] = r["obj"]
        np.ctypeslib.as_array(ST, (B,))[:] = r["status"]
        np.ctypeslib.as_array(IT, (B,))[:] = r["iters"]
        if X:
            np.ctypeslib.as_array(X, (B, n))[:] = r["x"]
        if G and m:
            np.ctypeslib.as_array(G, (B, m))[:] = r["g"]
        if D:
            np.ctypeslib.as_array(D, (B, 2 * n + m))[:] = r["duals"]
        return 0

    res = BbResult()
    x = np.zeros(n)
    arr = dict(vt=np.ascontiguousarray(model.vtype, np.int32), xl=np.ascontiguousarray(model.xl, float),
               xu=np.ascontiguousarray(model.xu, float), gl=np.ascontiguousarray(model.gl, float),
               gu=np.ascontiguousarray(model.gu, float), x0=np.ascontiguousarray(model.x0, float))
    L = host_lib()
    L.bnlp_bb_run_cb2.argtypes = [c_int, c_int, c_void, c_void, c_void, c_void, c_void, c_void, CB, QCB, c_void,
                                  ctypes.POINTER(BbOptions), ctypes.POINTER(BbResult), c_void]
    o = _opts(**kw)
    keep = CB(cb)
    keepq = QCB(qcb) if tqp_solve is not None else ctypes.cast(None, QCB)
    rc = L.bnlp_bb_run_cb2(n, m, *[arr[k].ctypes.data_as(c_void) for k in ("vt", "xl", "xu", "gl", "gu", "x0")], keep,
                           keepq, None, ctypes.byref(o), ctypes.byref(res), x.ctypes.data_as(c_void))
    if rc:
        raise RuntimeError("bnlp_bb_run_cb failed" + (f": {failure[0]!r}" if failure else ""))
    d = res.as_dict()
    d["best_x"] = x
    return d


class DiveOptions(ctypes.Structure):
    _fields_ = [("rule", c_int), ("speculate", c_int), ("integer_tolerance", c_dbl), ("percentage_to_fix", c_dbl),
                ("seed", ctypes.c_uint)]


class DiveResult(ctypes.Structure):
    _fields_ = [("found", c_int), ("steps", c_int), ("nlp_solves", c_int), ("batches", c_int),
                ("ipm_iterations", ctypes.c_longlong), ("obj", c_dbl)]


GCB = ctypes.CFUNCTYPE(c_int, c_void, P_dbl, P_dbl)
DIVE_RULES = {"fractional": 0, "vectorLength": 1}


def dive_engine(eng: "_e.Engine", node, cutoff=1e100, rule="fractional", speculate=True, xl=None, xu=None):
    """HeuristicDive from a solved relaxation `node` (dict with x, g, obj of Engine.solve_batch row 0) on the engine."""
    info = eng.problem.info()
    xl = np.ascontiguousarray(info["xl"] if xl is None else xl, float)
    xu = np.ascontiguousarray(info["xu"] if xu is None else xu, float)
    xs = np.ascontiguousarray(node["x"], float)
    gs = np.ascontiguousarray(node["g"], float)
    res, x = DiveResult(), np.zeros(eng.n)
    o = DiveOptions(DIVE_RULES[rule], int(speculate), 0.0, 0.0, 1)
    L = _e.lib()
    L.bnlp_dive_run_engine.argtypes = [c_void, c_void, c_void, c_void, c_void, c_void, c_dbl, c_dbl,
                                       ctypes.POINTER(DiveOptions), ctypes.POINTER(DiveResult), c_void]
    rc = L.bnlp_dive_run_engine(eng.h, eng.problem.h, xl.ctypes.data_as(c_void), xu.ctypes.data_as(c_void),
                                xs.ctypes.data_as(c_void), gs.ctypes.data_as(c_void), float(node["obj"]), float(cutoff),
                                ctypes.byref(o), ctypes.byref(res), x.ctypes.data_as(c_void))
    if rc:
        raise _e.BnlpError("bnlp_dive_run_engine failed")
    d = {k: getattr(res, k) for k, _ in res._fields_}
    d["x"] = x
    return d


def dive_callback(model, solve_batch, grad_f, column_length, node, cutoff=1e100, rule="fractional", speculate=True):
    """Same dive with Python providers: solve_batch as in run_callback, grad_f(x) -> gradient of the objective."""
    n, m = model.n, model.m

    def cb(ctx, B, XL, XU, X0, OBJ, ST, IT, X, G, D):
        r = solve_batch(np.ctypeslib.as_array(XL, (B, n)).copy(), np.ctypeslib.as_array(XU, (B, n)).copy(),
                        np.ctypeslib.as_array(X0, (n,)).copy())
        np.ctypeslib.as_array(OBJ, (B,))[:] = r["obj"]
        np.ctypeslib.as_array(ST, (B,))[:] = r["status"]
        np.ctypeslib.as_array(IT, (B,))[:] = r["iters"]
        if X:
            np.ctypeslib.as_array(X, (B, n))[:] = r["x"]
        if G and m:
            np.ctypeslib.as_array(G, (B, m))[:] = r["g"]
        return 0

    def gcb(ctx, X, GF):
        np.ctypeslib.as_array(GF, (n,))[:] = grad_f(np.ctypeslib.as_array(X, (n,)).copy())
        return 0

    arr = dict(vt=np.ascontiguousarray(model.vtype, np.int32), xl=np.ascontiguousarray(model.xl, float),
               xu=np.ascontiguousarray(model.xu, float), gl=np.ascontiguousarray(model.gl, float),
               gu=np.ascontiguousarray(model.gu, float), xs=np.ascontiguousarray(node["x"], float),
               gs=np.ascontiguousarray(node["g"], float), cl=np.ascontiguousarray(column_length, np.int32))
    res, x = DiveResult(), np.zeros(n)
    o = DiveOptions(DIVE_RULES[rule], int(speculate), 0.0, 0.0, 1)
    L = host_lib()
    L.bnlp_dive_run_cb.argtypes = [c_int, c_int] + [c_void] * 7 + [c_dbl, c_void, c_dbl, CB, GCB, c_void,
                                                                   ctypes.POINTER(DiveOptions), ctypes.POINTER(DiveResult), c_void]
    k1, k2 = CB(cb), GCB(gcb)
    p = lambda k: arr[k].ctypes.data_as(c_void)
    rc = L.bnlp_dive_run_cb(n, m, p("vt"), p("xl"), p("xu"), p("gl"), p("gu"), p("xs"), p("gs"), float(node["obj"]), p("cl"),
                            float(cutoff), k1, k2, None, ctypes.byref(o), ctypes.byref(res), x.ctypes.data_as(c_void))
    if rc:
        raise RuntimeError("bnlp_dive_run_cb failed")
    d = {k: getattr(res, k) for k, _ in res._fields_}
    d["x"] = x
    return d


class RetryResult(ctypes.Structure):
    _fields_ = [("used", c_int), ("num_failed", c_int), ("num_infeasible", c_int), ("nlp_solves", c_int), ("mean", c_dbl),
                ("std_dev", c_dbl)]


RETRY_MODES = {"robustness": 0, "cost": 1}


def _retry_out(n, m):
    return dict(obj=c_dbl(), status=c_int(), x=np.zeros(n), g=np.zeros(max(m, 1)), duals=np.zeros(2 * n + m))


def resolve_engine(eng: "_e.Engine", mode, x_init, numsolve, xl=None, xu=None, cur_obj=1e100, cur_status=0,
                   max_random_point_radius=0.0, seed=1):
    """resolveForRobustness / resolveForCost on the engine: the restarts are ONE batch with per-child start points."""
    info = eng.problem.info()
    n, m = eng.n, eng.m
    xl = np.ascontiguousarray(info["xl"] if xl is None else xl, float)
    xu = np.ascontiguousarray(info["xu"] if xu is None else xu, float)
    xi = np.ascontiguousarray(x_init, float)
    res, o = RetryResult(), _retry_out(n, m)
    L = _e.lib()
    L.bnlp_resolve_engine.argtypes = [c_void, c_int, c_void, c_void, c_void, c_dbl, c_int, c_int, c_dbl, ctypes.c_uint,
                                      ctypes.POINTER(RetryResult), P_dbl, P_int, c_void, c_void, c_void]
    p = lambda a: a.ctypes.data_as(c_void)
    rc = L.bnlp_resolve_engine(eng.h, RETRY_MODES[mode], p(xl), p(xu), p(xi), float(cur_obj), int(cur_status), int(numsolve),
                               float(max_random_point_radius), seed, ctypes.byref(res), ctypes.byref(o["obj"]),
                               ctypes.byref(o["status"]), p(o["x"]), p(o["g"]), p(o["duals"]))
    if rc:
        raise _e.BnlpError("bnlp_resolve_engine failed")
    d = {k: getattr(res, k) for k, _ in res._fields_}
    d.update(obj=o["obj"].value, status=o["status"].value, x=o["x"], g=o["g"][:m], duals=o["duals"])
    return d


def resolve_callback(model, solve_starts, mode, x_init, numsolve, cur_obj=1e100, cur_status=0, max_random_point_radius=0.0,
                     seed=1):
    """Same with a Python provider: solve_starts(XL, XU, X0[B x n]) -> dict like Engine.solve_batch."""
    n, m = model.n, model.m

    def cb(ctx, B, XL, XU, X0, OBJ, ST, IT, X, G, D):
        r = solve_starts(np.ctypeslib.as_array(XL, (B, n)).copy(), np.ctypeslib.as_array(XU, (B, n)).copy(),
                         np.ctypeslib.as_array(X0, (B, n)).copy())
        np.ctypeslib.as_array(OBJ, (B,))[:] = r["obj"]
        np.ctypeslib.as_array(ST, (B,))[:] = r["status"]
        np.ctypeslib.as_array(IT, (B,))[:] = r["iters"]
        np.ctypeslib.as_array(X, (B, n))[:] = r["x"]
        if m:
            np.ctypeslib.as_array(G, (B, m))[:] = r["g"]
        np.ctypeslib.as_array(D, (B, 2 * n + m))[:] = r["duals"]
        return 0

    xl = np.ascontiguousarray(model.xl, float)
    xu = np.ascontiguousarray(model.xu, float)
    xi = np.ascontiguousarray(x_init, float)
    res, o = RetryResult(), _retry_out(n, m)
    L = host_lib()
    L.bnlp_resolve_cb.argtypes = [c_int, c_int, c_int, c_void, c_void, c_void, c_dbl, c_int, c_int, c_dbl, ctypes.c_uint, CB,
                                  c_void, ctypes.POINTER(RetryResult), P_dbl, P_int, c_void, c_void, c_void]
    keep = CB(cb)
    p = lambda a: a.ctypes.data_as(c_void)
    rc = L.bnlp_resolve_cb(RETRY_MODES[mode], n, m, p(xl), p(xu), p(xi), float(cur_obj), int(cur_status), int(numsolve),
                           float(max_random_point_radius), seed, keep, None, ctypes.byref(res), ctypes.byref(o["obj"]),
                           ctypes.byref(o["status"]), p(o["x"]), p(o["g"]), p(o["duals"]))
    if rc:
        raise RuntimeError("bnlp_resolve_cb failed")
    d = {k: getattr(res, k) for k, _ in res._fields_}
    d.update(obj=o["obj"].value, status=o["status"].value, x=o["x"], g=o["g"][:m], duals=o["duals"])
    return d
