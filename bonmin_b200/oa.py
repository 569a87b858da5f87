"""Host B-OA driver: the outer-approximation decomposition loop of OACutGenerator2::performOa
(src/Algorithms/OaGenerators/BonOACutGenerator2.cpp:62-284, helpers BonSolverHelp.cpp:49-174, BonOaDecBase.cpp:237-400)
around the batched engine.  Host control flow only (the reference's MILP master is Cbc; here the master is solved by
HiGHS through scipy.optimize.milp — an external solver on either side).  What runs on the device:

  * the NLP relaxation and every NLP(y) with the integers fixed      -> provider.solve   (bnlp_solve_batch)
  * the initial LP (extractLinearRelaxation + addObjectiveFunction)   -> provider.extract (bnlp_extract_linear_relaxation)
  * the linearisations of all nonlinear rows, at SEVERAL points per pass (the NLP(y) optimum and, ECP style, the master
    point itself) in one call                                          -> provider.oa      (bnlp_oa_batch)

A provider is any object with those three methods (tests run the same loop on the CPU oracle)."""
from __future__ import annotations

import time

import numpy as np

DMAX = np.finfo(float).max


class EngineProvider:
    """Arithmetic provider = the CUDA engine (product path)."""

    def __init__(self, eng, model_info=None):
        self.eng = eng
        info = model_info or eng.problem.info()
        self.n, self.m = eng.n, eng.m
        self.xl, self.xu = info["xl"], info["xu"]
        self.vtype = info["var_type"]
        self.x0 = info["x0"]
        self._oa_out = {}
        # multipliers of the LAST NLP solve: the interface's getRowPrice() (all zero before the first solve, as the
        # reference's duals_sol_: a ranged row then gets no cut side at all, :1968-1974)
        self.duals = np.zeros(2 * self.n + self.m)

    def solve(self, XL, XU, x0):
        r = self.eng.solve_batch(XL, XU, x0)
        self.duals = np.array(r["duals"][-1])
        return r

    def extract(self, x, duals):
        r = self.eng.extract_linear_relaxation(x, duals, get_obj=True)
        rows = [(int(r["src_row"][k]), r["col_idx"][r["row_start"][k]:r["row_start"][k + 1]].tolist(),
                 r["val"][r["row_start"][k]:r["row_start"][k + 1]].tolist(), float(r["row_lb"][k]), float(r["row_ub"][k]))
                for k in range(r["n_rows"])]
        return dict(rows=rows, has_alpha=r["has_alpha"], obj=r["obj"], obj_offset=r["obj_offset"])

    def violation(self, x, obj):
        """getNonLinearitiesViolation(x, obj) (BonOsiTMINLPInterface.cpp:1869-1876)."""
        v, _ = self.eng.violation_batch(np.atleast_2d(x), np.atleast_1d(obj))
        return float(v[0])

    def oa(self, X, X2=None, theta=0.0):
        """getOuterApproximation(cs, x, getObj = 1, x2, theta) at every row of X.  As in the reference the side of a ranged /
        equality nonlinear row that is linearised is picked by the multipliers of the last NLP solve
        (BonOsiTMINLPInterface.cpp:1952,1968-1974: duals = getRowPrice()): for a convex g only the binding side is a valid
        outer approximation.  X2 rows are whole LP points (alpha column included when the objective cut exists)."""
        X = np.atleast_2d(X)
        # the cut buffers of each batch size are kept and page-locked: every pass of a loop reuses them
        r = self.eng.oa_batch(X, self.duals, self.xl, self.xu, X2=X2, theta=theta, get_obj=1, out=self._oa_out.get(len(X)), pin=True)
        self._oa_out[len(X)] = r
        out = []
        for k in range(len(r["src_row"])):
            if r["src_row"][k] < 0:
                continue
            a, b = r["row_start"][k], r["row_start"][k + 1]
            out.append((int(r["src_row"][k]), r["col"][a:b].tolist(), r["val"][a:b].tolist(), float(r["lb"][k]), float(r["ub"][k])))
        return out


def _lp_point(x_lp, n):
    """An LP point as getOuterApproximation's x2: the n columns plus the alpha column (0 if the LP has none)."""
    x = np.zeros(n + 1)
    k = min(len(x_lp), n + 1)
    x[:k] = np.asarray(x_lp, float)[:k]
    return x[None, :]


def _solve_master(n_cols, rows, obj, obj_offset, col_lb, col_ub, integer, cutoff):
    """min obj.x over the cut rows with HiGHS; returns (status, x, bound): status 0 optimal, 1 infeasible / cut off."""
    from scipy.optimize import Bounds, LinearConstraint, milp
    from scipy.sparse import csr_matrix
    indptr, idx, val, lo, hi = [0], [], [], [], []
    for _, cols, vals, lb, ub in rows:
        idx.extend(cols)
        val.extend(vals)
        indptr.append(len(idx))
        lo.append(-np.inf if lb <= -1e100 else lb)
        hi.append(np.inf if ub >= 1e100 else ub)
    cons = []
    if rows:
        cons.append(LinearConstraint(csr_matrix((val, idx, indptr), shape=(len(rows), n_cols)), lo, hi))
    if cutoff < 1e99:     # OsiDualObjectiveLimit (:150): only points better than the incumbent are of interest
        cons.append(LinearConstraint(csr_matrix(np.asarray(obj)[None, :]), -np.inf, cutoff - obj_offset))
    lb = np.where(np.asarray(col_lb) <= -1e100, -np.inf, col_lb)
    ub = np.where(np.asarray(col_ub) >= 1e100, np.inf, col_ub)
    res = milp(c=np.asarray(obj, float), constraints=cons, integrality=np.asarray(integer, int), bounds=Bounds(lb, ub),
               options=dict(mip_rel_gap=0.0))
    if res.status != 0 or res.x is None:
        return 1, None, 1e200
    return 0, res.x, float(res.fun) + obj_offset


def ecp_rounds(provider, lp, rounds=5, abs_violation_tol=1e-6, rel_violation_tol=1e-4, integer=None):
    """EcpCuts::generateCuts (src/Algorithms/OaGenerators/BonEcpCuts.cpp:46-143) on an LP relaxation `lp` (what
    provider.extract returns): while the LP optimum violates the nonlinear rows (getNonLinearitiesViolation), linearise
    them AT THE LP POINT (getOuterApproximation with getObj = 1), install the cuts, re-solve.  Returns dict(obj = LP
    bound after the rounds (1e308 if the LP became infeasible), violation, rounds, cs = every cut generated (what the
    reference returns in its OsiCuts), cuts = len(cs), installed = those that went into the LP, rows = the enlarged row
    list)."""
    P = provider
    n = P.n
    rows = list(lp["rows"])
    ncols = n + (1 if lp["has_alpha"] else 0)
    obj = np.zeros(ncols)
    obj[:len(lp["obj"])] = lp["obj"]
    col_lb = np.concatenate([P.xl, [-DMAX]])[:ncols]
    col_ub = np.concatenate([P.xu, [DMAX]])[:ncols]
    integ = np.zeros(ncols, int) if integer is None else integer
    out = dict(rounds=0, cuts=0, installed=0, cs=[], violation=0.0, obj=None)

    def lp_solve():
        st, x, val = _solve_master(ncols, rows, obj, lp["obj_offset"], col_lb, col_ub, integ, 1e100)
        return st, x, val

    st, x, val = lp_solve()
    if st != 0:
        out["obj"] = 1.7976931348623157e308
        return out
    alpha = (lambda xx: xx[n]) if ncols > n else (lambda xx: val)     # the LP's value of the objective (alpha column)
    orig = P.violation(x[:n], alpha(x))
    out["violation"] = orig
    viol = orig
    for i in range(rounds):
        if not (viol > abs_violation_tol and viol > rel_violation_tol * orig):
            break
        cuts = [c for c in P.oa(x[:n][None, :]) if not (c[0] == P.m and ncols == n)]
        out["cs"].extend(cuts)          # EcpCuts hands the cuts of EVERY round back in `cs` — the last round's too, which
        out["cuts"] += len(cuts)        # are generated but neither installed in the LP nor re-solved (:86)
        if not cuts or i + 1 >= rounds:
            break
        rows.extend(cuts)
        out["installed"] += len(cuts)
        out["rounds"] += 1
        st, x, val = lp_solve()
        if st != 0:
            out["obj"] = 1.7976931348623157e308
            out["rows"] = rows
            return out
        viol = P.violation(x[:n], alpha(x))
        out["violation"] = viol
    out["obj"] = val
    out["rows"] = rows
    out["x"] = x
    return out


def run_oa(provider, integer_tolerance=1e-6, cutoff_increment=1e-5, gap_tol=1e-9, max_passes=200, time_limit=1e10,
           add_master_point=True, log=None):
    """B-OA.  Returns dict(status 0 optimal / 1 infeasible / 2 limit, best_obj, best_x, lower_bound, passes, nlp_solves,
    cuts, oa_points, seconds)."""
    t0 = time.perf_counter()
    P = provider
    n = P.n
    ints = np.flatnonzero(np.asarray(P.vtype) != 0)
    out = dict(status=2, best_obj=1e100, best_x=None, lower_bound=-1e200, passes=0, nlp_solves=0, cuts=0, oa_points=0)
    # continuous relaxation and the initial LP at its optimum (BonOaDecBase / BonminSetup::initializeBHyb: the LP is
    # extracted at the solution of the first NLP)
    root = P.solve(np.asarray(P.xl)[None, :], np.asarray(P.xu)[None, :], P.x0)
    out["nlp_solves"] += 1
    if root["status"][0] not in (1, 2):
        out["status"] = 1 if root["status"][0] == 3 else 2
        out["seconds"] = time.perf_counter() - t0
        return out
    xr, dr = root["x"][0], root["duals"][0]
    lp = P.extract(xr, dr)
    rows = list(lp["rows"])
    ncols = n + (1 if lp["has_alpha"] else 0)
    obj = np.zeros(ncols)
    obj[:len(lp["obj"])] = lp["obj"]
    col_lb = np.concatenate([P.xl, [-DMAX]])[:ncols]
    col_ub = np.concatenate([P.xu, [DMAX]])[:ncols]
    integer = np.zeros(ncols, int)
    integer[ints] = 1
    cutoff, ub = 1e100, 1e100
    last_y = None
    while out["passes"] < max_passes and time.perf_counter() - t0 < time_limit:
        st, xm, bound = _solve_master(ncols, rows, obj, lp["obj_offset"], col_lb, col_ub, integer, cutoff)   # subMip_->solve
        if st != 0:                                       # master infeasible under the cutoff: the incumbent is optimal
            out["lower_bound"] = max(out["lower_bound"], min(ub, 1e200))
            out["status"] = 0 if ub < 1e99 else 1
            break
        out["lower_bound"] = max(out["lower_bound"], bound)
        out["passes"] += 1
        y = np.round(xm[ints])
        if last_y is not None and np.array_equal(y, last_y):   # isDifferentOnIntegers == false: fathomed (:196-199)
            out["status"] = 0 if ub < 1e99 else 1
            break
        last_y = y
        if ub < 1e99 and (ub - bound) / max(abs(ub), 1e-10) < gap_tol:
            out["status"] = 0
            break
        # NLP(y): integers fixed at the master's values (fixIntegers, BonSolverHelp.cpp:49-93)
        XL, XU = np.array(P.xl, float), np.array(P.xu, float)
        XL[ints] = XU[ints] = y
        r = P.solve(XL[None, :], XU[None, :], xm[:n])
        out["nlp_solves"] += 1
        xs = r["x"][0]
        if r["status"][0] in (1, 2):                      # post_nlp_solve: feasible -> incumbent and new cutoff (:146-151)
            if r["obj"][0] < ub:
                ub = float(r["obj"][0])
                out["best_obj"], out["best_x"] = ub, xs.copy()
            cutoff = ub * (1 - cutoff_increment) if ub > 0 else ub * (1 + cutoff_increment)
            if ub == 0:
                cutoff = -cutoff_increment
        # linearise at the NLP(y) point (for an infeasible NLP(y): at its minimum-violation point) and at the master point
        pts = [xs] + ([np.clip(xm[:n], P.xl, P.xu)] if add_master_point else [])
        cuts = P.oa(np.array(pts), X2=None)
        out["oa_points"] += len(pts)
        new = 0
        for c in cuts:
            if c[0] == P.m and ncols == n:               # objective cut of a linear objective: not generated (:2053)
                continue
            rows.append(c)
            new += 1
        out["cuts"] += new
        if log:
            log(f"pass {out['passes']}: master {bound:.6f} ub {ub:.6f} cuts +{new}")
    out["seconds"] = time.perf_counter() - t0
    return out


# ------------------------------------------------------------------------------------------ single-tree generators
def _lp_solve(ncols, rows, obj, obj_offset, lb, ub, cutoff):
    """LP relaxation at a node (what lp->resolve() is to the generators below); (feasible, x, value)."""
    st, x, val = _solve_master(ncols, rows, obj, obj_offset, lb, ub, np.zeros(ncols, int), 1e100)
    return (st == 0 and val < cutoff), x, val


def feasibility_check(provider, lp, col_lb, col_ub, x_lp, cutoff, integer_tolerance=1e-6, cutoff_increment=1e-5,
                      add_only_violated=False, max_passes=50):
    """OaFeasibilityChecker::performOa (src/Algorithms/OaGenerators/BonOaFeasChecker.cpp:38-135): called when the node LP
    has an integer solution.  Loop: fix the integers at the LP values (fixIntegers), solve NLP(y) on the engine, on
    success update incumbent and cutoff (post_nlp_solve), linearise at the NLP point (getOuterApproximation, getObj = 1;
    against the LP point only when add_only_violated), install the cuts in the node LP, re-solve it; go on while the LP
    stays feasible below the cutoff, integer, and different on the integers from the NLP point.
    lp: dict(rows, obj, obj_offset, ncols) — rows is extended in place (the cuts are global).  Returns dict(cutoff,
    incumbent (obj, x) or None, cuts, nlp_solves, fathomed = the node can be cut off, x_lp = last LP point or None)."""
    P = provider
    n = P.n
    ints = np.flatnonzero(np.asarray(P.vtype) != 0)
    out = dict(cutoff=cutoff, incumbent=None, cuts=0, nlp_solves=0, fathomed=False, x_lp=x_lp)
    is_integer, feasible = True, True
    passes = 0
    while is_integer and feasible and passes < max_passes:
        passes += 1
        XL, XU = np.array(P.xl, float), np.array(P.xu, float)
        XL[ints] = XU[ints] = np.round(x_lp[ints])
        r = P.solve(XL[None, :], XU[None, :], x_lp[:n])
        out["nlp_solves"] += 1
        xs = r["x"][0]
        if r["status"][0] in (1, 2) and r["obj"][0] < out["cutoff"]:
            ub = float(r["obj"][0])
            if out["incumbent"] is None or ub < out["incumbent"][0]:
                out["incumbent"] = (ub, xs.copy())
            out["cutoff"] = ub * (1 - cutoff_increment) if ub > 0 else ub * (1 + cutoff_increment)
        cuts = P.oa(xs[None, :], X2=(_lp_point(x_lp, n) if add_only_violated else None))
        cuts = [c for c in cuts if not (c[0] == P.m and lp["ncols"] == n)]
        lp["rows"].extend(cuts)
        out["cuts"] += len(cuts)
        feasible, x_new, val = _lp_solve(lp["ncols"], lp["rows"], lp["obj"], lp["obj_offset"], col_lb, col_ub, out["cutoff"])
        is_integer = False
        if feasible:
            changed = np.abs(xs[ints] - x_new[ints]).max(initial=0.0) > 0.1          # isDifferentOnIntegers(.., 0.1, ..)
            if changed:
                is_integer = np.abs(x_new[ints] - np.round(x_new[ints])).max(initial=0.0) <= integer_tolerance
            else:
                out["fathomed"] = True           # same integers as the NLP point: milpBound = 1e200 (:115-119)
            x_lp = x_new
            out["x_lp"] = x_new
        else:
            out["fathomed"] = True
            out["x_lp"] = None
    return out


def nlp_optim_cuts(provider, node_lb, node_ub, x_lp, add_only_violated=False):
    """OaNlpOptim::generateCuts (src/Algorithms/OaGenerators/BonOaNlpOptim.cpp:63-173): the NLP relaxation of the NODE
    (integer columns take the node LP's bounds), OA cuts at its solution (getObj = 1), and the solution itself when it
    happens to be integer (1e-7).  An abandoned / iteration-limited NLP is an error, as in the reference.  Returns
    dict(cuts, solution (obj, x) or None, status, obj)."""
    P = provider
    n = P.n
    ints = np.flatnonzero(np.asarray(P.vtype) != 0)
    XL, XU = np.array(P.xl, float), np.array(P.xu, float)
    XL[ints], XU[ints] = np.asarray(node_lb)[ints], np.asarray(node_ub)[ints]
    r = P.solve(XL[None, :], XU[None, :], x_lp[:n])
    st, xs = int(r["status"][0]), r["x"][0]
    if st in (-3, -2):            # isAbandoned() || isIterationLimitReached(): "Unsolved NLP ... exit"
        raise RuntimeError("Unsolved NLP ... exit")
    cuts = provider.oa(xs[None, :], X2=(_lp_point(x_lp, n) if add_only_violated else None))
    sol = None
    if st in (1, 2) and np.abs(xs[ints] - np.floor(xs[ints] + 0.5)).max(initial=0.0) <= 1e-7:
        sol = (float(r["obj"][0]), xs.copy())
    return dict(cuts=cuts, solution=sol, status=st, obj=float(r["obj"][0]))


def nlp_optim_wanted(level, solves_per_level=1.0, max_depth=10, rng=None):
    """The depth test of OaNlpOptim::generateCuts (:84-92): never below max_depth, else with probability
    2^-level * solves_per_level (always when solves_per_level >= 1e10)."""
    if level > max_depth:
        return False
    if solves_per_level >= 1e10:
        return True
    rnd = (rng or np.random.default_rng()).random()
    return 2.0 ** (-level) * solves_per_level > rnd


def run_qg(provider, integer_tolerance=1e-6, cutoff_increment=1e-5, node_limit=100000, nlp_solves_per_level=0.0,
           nlp_max_depth=10, seed=0, time_limit=1e10, log=None, feas_check=None, nlp_optim=None):
    """A single-tree LP/NLP-based branch and bound in the manner of B-QG (nlp_solves_per_level = 0) and B-Hyb (> 0):
    LP relaxation from extractLinearRelaxation at the root NLP optimum, best-bound search branching on the most
    fractional integer, OaFeasibilityChecker at integer LP points, OaNlpOptim at nodes by the depth rule.  Host control
    flow (the reference's is Cbc + Cgl generators, BonBonminSetup.cpp:478-620); the LPs are solved by HiGHS; NLPs, the
    initial LP and every linearisation come from the provider (the engine in the product).  Returns dict(status 0 optimal /
    1 infeasible / 2 limit, best_obj, best_x, nodes, nlp_solves, cuts, seconds)."""
    import heapq
    t0 = time.perf_counter()
    P = provider
    n = P.n
    rng = np.random.default_rng(seed)
    ints = np.flatnonzero(np.asarray(P.vtype) != 0)
    out = dict(status=2, best_obj=1e100, best_x=None, nodes=0, nlp_solves=1, cuts=0)
    root = P.solve(np.asarray(P.xl)[None, :], np.asarray(P.xu)[None, :], P.x0)
    if root["status"][0] not in (1, 2):
        out["status"] = 1 if root["status"][0] == 3 else 2
        out["seconds"] = time.perf_counter() - t0
        return out
    ext = P.extract(root["x"][0], root["duals"][0])
    ncols = n + (1 if ext["has_alpha"] else 0)
    obj = np.zeros(ncols)
    obj[:len(ext["obj"])] = ext["obj"]
    lp = dict(rows=list(ext["rows"]), obj=obj, obj_offset=ext["obj_offset"], ncols=ncols)
    lb0 = np.concatenate([P.xl, [-DMAX]])[:ncols]
    ub0 = np.concatenate([P.xu, [DMAX]])[:ncols]
    cutoff = 1e100
    heap, tick = [(float(root["obj"][0]), 0, 0, lb0, ub0)], 1
    while heap:
        if out["nodes"] >= node_limit or time.perf_counter() - t0 > time_limit:
            break
        bound, _, level, lb, ub = heapq.heappop(heap)
        if bound >= cutoff:
            continue
        out["nodes"] += 1
        feas, x, val = _lp_solve(ncols, lp["rows"], lp["obj"], lp["obj_offset"], lb, ub, cutoff)
        if not feas:
            continue
        if nlp_solves_per_level > 0 and nlp_optim_wanted(level, nlp_solves_per_level, nlp_max_depth, rng):
            g = (nlp_optim or nlp_optim_cuts)(P, lb, ub, x)      # or the C++ generator (nlp_optim_cuts_native)
            out["nlp_solves"] += 1
            new = [c for c in g["cuts"] if not (c[0] == P.m and ncols == n)]
            lp["rows"].extend(new)
            out["cuts"] += len(new)
            if g["status"] == 3:
                continue                                   # the node's NLP relaxation is infeasible
            if g["solution"] and g["solution"][0] < out["best_obj"]:
                out["best_obj"], out["best_x"] = g["solution"]
                cutoff = min(cutoff, out["best_obj"] * (1 - cutoff_increment) if out["best_obj"] > 0 else out["best_obj"] * (1 + cutoff_increment))
            feas, x, val = _lp_solve(ncols, lp["rows"], lp["obj"], lp["obj_offset"], lb, ub, cutoff)
            if not feas:
                continue
        frac = np.abs(x[ints] - np.round(x[ints]))
        if frac.max(initial=0.0) <= integer_tolerance:
            # feas_check: feasibility_check (default) or the C++ loop behind the same signature (feasibility_check_native)
            f = (feas_check or feasibility_check)(P, lp, lb, ub, x, cutoff, integer_tolerance, cutoff_increment)
            out["nlp_solves"] += f["nlp_solves"]
            out["cuts"] += f["cuts"]
            cutoff = min(cutoff, f["cutoff"])
            if f["incumbent"] and f["incumbent"][0] < out["best_obj"]:
                out["best_obj"], out["best_x"] = f["incumbent"]
            if f["fathomed"] or f["x_lp"] is None:
                continue
            x = f["x_lp"]                                   # the LP moved to a fractional point: branch on it
            feas, x, val = _lp_solve(ncols, lp["rows"], lp["obj"], lp["obj_offset"], lb, ub, cutoff)
            if not feas:
                continue
            frac = np.abs(x[ints] - np.round(x[ints]))
            if frac.max(initial=0.0) <= integer_tolerance:
                heapq.heappush(heap, (val, tick, level, lb, ub)); tick += 1      # integer again: re-examine the node
                continue
        j = ints[int(np.argmax(frac))]
        for way in (0, 1):
            l2, u2 = lb.copy(), ub.copy()
            if way == 0:
                u2[j] = np.floor(x[j])
            else:
                l2[j] = np.ceil(x[j])
            heapq.heappush(heap, (val, tick, level + 1, l2, u2)); tick += 1
        if log:
            log(f"node {out['nodes']}: lp {val:.6f} cutoff {cutoff:.6f} rows {len(lp['rows'])}")
    else:
        out["status"] = 0 if out["best_obj"] < 1e99 else 1
    if not heap and out["status"] == 2 and out["nodes"] < node_limit:
        out["status"] = 0 if out["best_obj"] < 1e99 else 1
    out["seconds"] = time.perf_counter() - t0
    return out


# ------------------------------------------------------------------------------------------ the C++ B-OA loop
# bonmin_b200/host/B200Oa.cpp (bnlp_oa_run_engine / bnlp_oa_run_cb, include/bnlp.h): the same decomposition loop as run_oa
# in C++, engine calls made directly from the driver; only the master MILP (the reference's SubMipSolver — Cbc there,
# HiGHS here) is a callback.  run_oa above stays as the readable statement of the loop and as its cross-check.
import ctypes as _ct

_c_int, _c_dbl, _c_void = _ct.c_int, _ct.c_double, _ct.c_void_p
_P_int, _P_dbl = _ct.POINTER(_c_int), _ct.POINTER(_c_dbl)
MASTER_CB = _ct.CFUNCTYPE(_c_int, _c_void, _c_int, _c_int, _P_int, _P_int, _P_dbl, _P_dbl, _P_dbl, _P_dbl, _P_dbl, _P_dbl, _c_dbl,
                          _P_int, _c_dbl, _c_dbl, _P_dbl, _P_dbl)
CUTS_CB = _ct.CFUNCTYPE(_c_int, _c_void, _c_int, _P_dbl, _P_dbl, _P_int, _P_int, _P_int, _P_dbl, _P_dbl, _P_dbl, _P_int, _c_int, _c_int)
EXTRACT_CB = _ct.CFUNCTYPE(_c_int, _c_void, _P_dbl, _P_dbl, _P_int, _P_int, _P_dbl, _P_int, _P_int, _P_dbl, _P_dbl, _P_dbl, _P_dbl,
                           _c_int, _c_int)


class OaRunOptions(_ct.Structure):
    _fields_ = [("cutoff_increment", _c_dbl), ("gap_tol", _c_dbl), ("time_limit", _c_dbl), ("max_passes", _c_int),
                ("no_master_point", _c_int)]


class OaRunResult(_ct.Structure):
    _fields_ = [("status", _c_int), ("passes", _c_int), ("nlp_solves", _c_int), ("cuts", _c_int), ("oa_points", _c_int),
                ("best_obj", _c_dbl), ("lower_bound", _c_dbl), ("seconds", _c_dbl), ("master_seconds", _c_dbl)]


def _as(p, shape):
    return np.ctypeslib.as_array(p, shape)


def highs_master(failure=None):
    """The master MILP as a bnlp_master_mip_cb: HiGHS through _solve_master (the same call run_oa makes)."""
    def cb(ctx, ncols, nrows, RS, COL, VAL, RLB, RUB, CLB, CUB, OBJ, off, ISINT, cutoff, max_s, X, BOUND):
        try:
            rs = _as(RS, (nrows + 1,))
            nnz = int(rs[nrows])
            col, val = _as(COL, (max(nnz, 1),)), _as(VAL, (max(nnz, 1),))
            lb, ub = _as(RLB, (max(nrows, 1),)), _as(RUB, (max(nrows, 1),))
            rows = [(-1, col[rs[k]:rs[k + 1]].tolist(), val[rs[k]:rs[k + 1]].tolist(), float(lb[k]), float(ub[k]))
                    for k in range(nrows)]
            st, x, bound = _solve_master(ncols, rows, _as(OBJ, (ncols,)).copy(), off, _as(CLB, (ncols,)).copy(),
                                         _as(CUB, (ncols,)).copy(), _as(ISINT, (ncols,)).copy(), cutoff)
            if st != 0:
                return 1
            _as(X, (ncols,))[:] = x
            BOUND[0] = bound
            return 0
        except Exception as e:      # a master failure reaches the driver as an error code and the caller as an exception
            if failure is not None:
                failure.append(e)
            return -1
    return MASTER_CB(cb)


def _run_opts(cutoff_increment=1e-5, gap_tol=1e-9, max_passes=200, time_limit=1e10, add_master_point=True):
    return OaRunOptions(cutoff_increment, gap_tol, time_limit, max_passes, 0 if add_master_point else 1)


def _run_result(res, x, n, failure):
    if failure:
        raise failure[0]
    out = {k: getattr(res, k) for k, _ in res._fields_}
    out["best_x"] = x.copy() if res.best_obj < 1e99 else None
    return out


def run_oa_engine(eng, **kw):
    """B-OA with the C++ loop calling the CUDA engine directly (bnlp_oa_run_engine).  Same result keys as run_oa, plus
    master_seconds (time inside the master callback)."""
    from . import engine as _e
    L = _e.lib()
    L.bnlp_oa_run_engine.argtypes = [_c_void, _c_void, MASTER_CB, _c_void, _ct.POINTER(OaRunOptions), _ct.POINTER(OaRunResult), _c_void]
    failure = []
    master = highs_master(failure)
    res, x, o = OaRunResult(), np.zeros(eng.n), _run_opts(**kw)
    rc = L.bnlp_oa_run_engine(eng.h, eng.problem.h, master, None, _ct.byref(o), _ct.byref(res), x.ctypes.data_as(_c_void))
    if rc and not failure:
        raise _e.BnlpError(f"bnlp_oa_run_engine failed ({rc}): " + L.bnlp_last_error().decode())
    return _run_result(res, x, eng.n, failure)


def run_oa_callback(provider, **kw):
    """The C++ loop (bnlp_oa_run_cb, libbnlp_host.so: no CUDA needed) over any arithmetic provider with solve / oa / extract
    (tests: the CPU oracle)."""
    from . import bb as _bb
    P = provider
    n, m = P.n, P.m
    failure = []

    def solve(ctx, B, XL, XU, X0, OBJ, ST, IT, X, G, D):
        try:
            r = P.solve(_as(XL, (B, n)).copy(), _as(XU, (B, n)).copy(), _as(X0, (n,)).copy())
            _as(OBJ, (B,))[:] = r["obj"]
            _as(ST, (B,))[:] = r["status"]
            _as(IT, (B,))[:] = r["iters"]
            if X:
                _as(X, (B, n))[:] = r["x"]
            if D:
                _as(D, (B, 2 * n + m))[:] = r["duals"]
            return 0
        except Exception as e:
            failure.append(e)
            return -1

    def cuts(ctx, B, X, DU, NC, RS, COL, VAL, LB, UB, SRC, cap_cuts, cap_nnz):
        try:
            P.duals = _as(DU, (2 * n + m,)).copy()
            cs = P.oa(_as(X, (B, n)).copy())
            nnz = sum(len(c[1]) for c in cs)
            if len(cs) > cap_cuts or nnz > cap_nnz:
                return -2
            rs = _as(RS, (cap_cuts + 1,))
            rs[0] = 0
            e = 0
            for k, (src, cols, vals, lb, ub) in enumerate(cs):
                _as(COL, (cap_nnz,))[e:e + len(cols)] = cols
                _as(VAL, (cap_nnz,))[e:e + len(cols)] = vals
                e += len(cols)
                rs[k + 1] = e
                LB[k], UB[k], SRC[k] = lb, ub, src
            NC[0] = len(cs)
            return 0
        except Exception as ex:
            failure.append(ex)
            return -1

    def extract(ctx, X, DU, NR, HA, OFF, RS, COL, VAL, RLB, RUB, OBJ, cap_rows, cap_nnz):
        try:
            lp = P.extract(_as(X, (n,)).copy(), _as(DU, (2 * n + m,)).copy())
            rows = lp["rows"]
            nnz = sum(len(r[1]) for r in rows)
            if len(rows) > cap_rows or nnz > cap_nnz:
                return -2
            RS[0] = 0
            e = 0
            for k, (_, cols, vals, lb, ub) in enumerate(rows):
                for c, v in zip(cols, vals):
                    COL[e], VAL[e] = c, v
                    e += 1
                RS[k + 1] = e
                RLB[k], RUB[k] = lb, ub
            ob = np.zeros(n + 1)
            ob[:len(lp["obj"])] = lp["obj"]
            _as(OBJ, (n + 1,))[:] = ob
            NR[0], HA[0], OFF[0] = len(rows), int(bool(lp["has_alpha"])), float(lp["obj_offset"])
            return 0
        except Exception as ex:
            failure.append(ex)
            return -1

    L = _bb.host_lib()
    L.bnlp_oa_run_cb.argtypes = [_c_int, _c_int, _c_void, _c_void, _c_void, _c_void, _bb.CB, CUTS_CB, EXTRACT_CB, MASTER_CB, _c_void,
                                 _ct.POINTER(OaRunOptions), _ct.POINTER(OaRunResult), _c_void]
    keep = (_bb.CB(solve), CUTS_CB(cuts), EXTRACT_CB(extract), highs_master(failure))
    vt = np.ascontiguousarray(P.vtype, np.int32)
    xl, xu, x0 = (np.ascontiguousarray(a, float) for a in (P.xl, P.xu, P.x0))
    res, x, o = OaRunResult(), np.zeros(n), _run_opts(**kw)
    rc = L.bnlp_oa_run_cb(n, m, vt.ctypes.data_as(_c_void), xl.ctypes.data_as(_c_void), xu.ctypes.data_as(_c_void),
                          x0.ctypes.data_as(_c_void), *keep, None, _ct.byref(o), _ct.byref(res), x.ctypes.data_as(_c_void))
    if rc and not failure:
        raise RuntimeError(f"bnlp_oa_run_cb failed ({rc})")
    return _run_result(res, x, n, failure)


class OaFeasOptions(_ct.Structure):
    _fields_ = [("integer_tolerance", _c_dbl), ("cutoff_increment", _c_dbl), ("max_passes", _c_int)]


class OaFeasResult(_ct.Structure):
    _fields_ = [("status", _c_int), ("cuts", _c_int), ("nlp_solves", _c_int), ("fathomed", _c_int), ("has_incumbent", _c_int),
                ("has_x_lp", _c_int), ("new_nnz", _c_int), ("cutoff", _c_dbl), ("incumbent_obj", _c_dbl)]


def _lp_csr(lp, col_lb, col_ub):
    rows = lp["rows"]
    rs = np.zeros(len(rows) + 1, np.int32)
    rs[1:] = np.cumsum([len(r[1]) for r in rows])
    col = np.ascontiguousarray(np.concatenate([np.asarray(r[1], np.int32) for r in rows]) if rows else np.zeros(1), np.int32)
    val = np.ascontiguousarray(np.concatenate([np.asarray(r[2], float) for r in rows]) if rows else np.zeros(1), float)
    lb = np.ascontiguousarray([r[3] for r in rows] or [0.0], float)
    ub = np.ascontiguousarray([r[4] for r in rows] or [0.0], float)
    nc = lp["ncols"]
    obj = np.zeros(nc)
    obj[:min(nc, len(lp["obj"]))] = np.asarray(lp["obj"], float)[:nc]
    return rs, col, val, lb, ub, np.ascontiguousarray(col_lb, float), np.ascontiguousarray(col_ub, float), obj


def feasibility_check_native(target, lp, col_lb, col_ub, x_lp, cutoff, integer_tolerance=1e-6, cutoff_increment=1e-5, max_passes=50):
    """feasibility_check through the C++ loop (B200Oa.cpp: oa_feas_check).  target: an Engine (bnlp_oa_feas_check_engine, the
    product path) or an arithmetic provider (bnlp_oa_feas_check_cb; tests pass the CPU oracle).  Same arguments and result
    as feasibility_check (add_only_violated is not offered); lp["rows"] is extended in place."""
    from . import bb as _bb, engine as _e
    is_engine = isinstance(target, _e.Engine)
    n, m = target.n, target.m
    nc = lp["ncols"]
    rs, col, val, lb, ub, clb, cub, obj = _lp_csr(lp, col_lb, col_ub)
    xlp = np.ascontiguousarray(np.asarray(x_lp, float)[:nc])
    cap_cuts, cap_nnz = max_passes * (m + 1), max_passes * (m + 1) * (n + 1)
    # (np.empty: the capacity is the worst case of max_passes full rounds; untouched pages are never committed)
    nrs, ncol, nval = np.zeros(cap_cuts + 1, np.int32), np.empty(cap_nnz, np.int32), np.empty(cap_nnz)
    nlb, nub, nsrc = np.zeros(cap_cuts), np.zeros(cap_cuts), np.zeros(cap_cuts, np.int32)
    inc_x, x_out = np.zeros(n), np.zeros(nc)
    res, o = OaFeasResult(), OaFeasOptions(integer_tolerance, cutoff_increment, max_passes)
    failure = []
    master = highs_master(failure)
    vp = lambda a: a.ctypes.data_as(_c_void)
    lp_args = [nc, len(lp["rows"]), vp(rs), vp(col), vp(val), vp(lb), vp(ub), vp(clb), vp(cub), vp(obj), float(lp["obj_offset"]),
               vp(xlp), float(cutoff), _ct.byref(o)]
    out_args = [_ct.byref(res), vp(inc_x), vp(x_out), vp(nrs), vp(ncol), vp(nval), vp(nlb), vp(nub), vp(nsrc), cap_cuts, cap_nnz]
    lp_types = [_c_int, _c_int] + [_c_void] * 8 + [_c_dbl, _c_void, _c_dbl, _ct.POINTER(OaFeasOptions)]
    out_types = [_ct.POINTER(OaFeasResult)] + [_c_void] * 8 + [_c_int, _c_int]
    if is_engine:
        L = _e.lib()
        L.bnlp_oa_feas_check_engine.argtypes = [_c_void, _c_void] + lp_types + [MASTER_CB, _c_void] + out_types
        rc = L.bnlp_oa_feas_check_engine(target.h, target.problem.h, *lp_args, master, None, *out_args)
    else:
        P = target

        def solve(ctx, B, XL, XU, X0, OBJ, ST, IT, X, G, D):
            try:
                r = P.solve(_as(XL, (B, n)).copy(), _as(XU, (B, n)).copy(), _as(X0, (n,)).copy())
                _as(OBJ, (B,))[:] = r["obj"]
                _as(ST, (B,))[:] = r["status"]
                _as(IT, (B,))[:] = r["iters"]
                if X:
                    _as(X, (B, n))[:] = r["x"]
                if D:
                    _as(D, (B, 2 * n + m))[:] = r["duals"]
                return 0
            except Exception as e:
                failure.append(e)
                return -1

        def cuts(ctx, B, X, DU, NC, RS, COL, VAL, LB, UB, SRC, cc, cn):
            try:
                P.duals = _as(DU, (2 * n + m,)).copy()
                cs = P.oa(_as(X, (B, n)).copy())
                if len(cs) > cc or sum(len(c[1]) for c in cs) > cn:
                    return -2
                RS[0] = 0
                e = 0
                for k, (src, cols, vals, l, u) in enumerate(cs):
                    _as(COL, (cn,))[e:e + len(cols)] = cols
                    _as(VAL, (cn,))[e:e + len(cols)] = vals
                    e += len(cols)
                    RS[k + 1] = e
                    LB[k], UB[k], SRC[k] = l, u, src
                NC[0] = len(cs)
                return 0
            except Exception as ex:
                failure.append(ex)
                return -1

        L = _bb.host_lib()
        L.bnlp_oa_feas_check_cb.argtypes = ([_c_int, _c_int, _c_void, _c_void, _c_void] + lp_types +
                                            [_bb.CB, CUTS_CB, MASTER_CB, _c_void] + out_types)
        vt = np.ascontiguousarray(P.vtype, np.int32)
        xl, xu = np.ascontiguousarray(P.xl, float), np.ascontiguousarray(P.xu, float)
        keep = (_bb.CB(solve), CUTS_CB(cuts))
        rc = L.bnlp_oa_feas_check_cb(n, m, vp(vt), vp(xl), vp(xu), *lp_args, keep[0], keep[1], master, None, *out_args)
    if failure:
        raise failure[0]
    if rc:
        raise RuntimeError(f"bnlp_oa_feas_check failed ({rc})")
    for k in range(res.cuts):
        a, b = nrs[k], nrs[k + 1]
        lp["rows"].append((int(nsrc[k]), ncol[a:b].tolist(), nval[a:b].tolist(), float(nlb[k]), float(nub[k])))
    return dict(cutoff=res.cutoff, incumbent=(res.incumbent_obj, inc_x.copy()) if res.has_incumbent else None, cuts=res.cuts,
                nlp_solves=res.nlp_solves, fathomed=bool(res.fathomed), x_lp=x_out.copy() if res.has_x_lp else None)


class NlpOptimResult(_ct.Structure):
    _fields_ = [("nlp_status", _c_int), ("cuts", _c_int), ("new_nnz", _c_int), ("has_solution", _c_int), ("obj", _c_dbl)]


class EcpResult(_ct.Structure):
    _fields_ = [("rounds", _c_int), ("cuts", _c_int), ("installed", _c_int), ("new_nnz", _c_int), ("lp_infeasible", _c_int),
                ("violation", _c_dbl), ("obj", _c_dbl)]


VIOL_CB = _ct.CFUNCTYPE(_c_int, _c_void, _P_dbl, _c_dbl, _P_dbl)


def _provider_callbacks(P, failure):
    """ctypes callbacks (solve, cuts, violation) over an arithmetic provider."""
    n, m = P.n, P.m

    def solve(ctx, B, XL, XU, X0, OBJ, ST, IT, X, G, D):
        try:
            r = P.solve(_as(XL, (B, n)).copy(), _as(XU, (B, n)).copy(), _as(X0, (n,)).copy())
            _as(OBJ, (B,))[:] = r["obj"]
            _as(ST, (B,))[:] = r["status"]
            _as(IT, (B,))[:] = r["iters"]
            if X:
                _as(X, (B, n))[:] = r["x"]
            if D:
                _as(D, (B, 2 * n + m))[:] = r["duals"]
            return 0
        except Exception as e:
            failure.append(e)
            return -1

    def cuts(ctx, B, X, DU, NC, RS, COL, VAL, LB, UB, SRC, cc, cn):
        try:
            P.duals = _as(DU, (2 * n + m,)).copy()
            cs = P.oa(_as(X, (B, n)).copy())
            if len(cs) > cc or sum(len(c[1]) for c in cs) > cn:
                return -2
            RS[0] = 0
            e = 0
            for k, (src, cols, vals, l, u) in enumerate(cs):
                _as(COL, (cn,))[e:e + len(cols)] = cols
                _as(VAL, (cn,))[e:e + len(cols)] = vals
                e += len(cols)
                RS[k + 1] = e
                LB[k], UB[k], SRC[k] = l, u, src
            NC[0] = len(cs)
            return 0
        except Exception as ex:
            failure.append(ex)
            return -1

    def viol(ctx, X, obj, V):
        try:
            V[0] = P.violation(_as(X, (n,)).copy(), obj)
            return 0
        except Exception as ex:
            failure.append(ex)
            return -1

    from . import bb as _bb
    return _bb.CB(solve), CUTS_CB(cuts), VIOL_CB(viol)


def _cut_outs(cap_cuts, cap_nnz):
    return (np.zeros(cap_cuts + 1, np.int32), np.empty(max(cap_nnz, 1), np.int32), np.empty(max(cap_nnz, 1)), np.zeros(max(cap_cuts, 1)),
            np.zeros(max(cap_cuts, 1)), np.zeros(max(cap_cuts, 1), np.int32))


def _cuts_list(k_cuts, outs):
    nrs, ncol, nval, nlb, nub, nsrc = outs
    return [(int(nsrc[k]), ncol[nrs[k]:nrs[k + 1]].tolist(), nval[nrs[k]:nrs[k + 1]].tolist(), float(nlb[k]), float(nub[k]))
            for k in range(k_cuts)]


def nlp_optim_cuts_native(target, node_lb, node_ub, x_lp):
    """nlp_optim_cuts through the C++ generator (B200Oa.cpp: oa_nlp_optim); target: an Engine (bnlp_oa_nlp_optim_engine) or a
    provider (bnlp_oa_nlp_optim_cb).  Same result dict as nlp_optim_cuts, plus duals."""
    from . import bb as _bb, engine as _e
    n, m = target.n, target.m
    vp = lambda a: a.ctypes.data_as(_c_void)
    lb, ub = np.ascontiguousarray(np.asarray(node_lb, float)[:n]), np.ascontiguousarray(np.asarray(node_ub, float)[:n])
    xlp = np.ascontiguousarray(np.asarray(x_lp, float)[:n])
    outs = _cut_outs(m + 1, (m + 1) * (n + 1))
    res, x, du = NlpOptimResult(), np.zeros(n), np.zeros(2 * n + m)
    tail = [_ct.byref(res), vp(x), vp(du)] + [vp(a) for a in outs] + [m + 1, (m + 1) * (n + 1)]
    tail_t = [_ct.POINTER(NlpOptimResult)] + [_c_void] * 8 + [_c_int, _c_int]
    failure = []
    if isinstance(target, _e.Engine):
        L = _e.lib()
        L.bnlp_oa_nlp_optim_engine.argtypes = [_c_void] * 5 + tail_t
        rc = L.bnlp_oa_nlp_optim_engine(target.h, target.problem.h, vp(lb), vp(ub), vp(xlp), *tail)
    else:
        P = target
        cbs = _provider_callbacks(P, failure)
        L = _bb.host_lib()
        L.bnlp_oa_nlp_optim_cb.argtypes = [_c_int, _c_int] + [_c_void] * 6 + [_bb.CB, CUTS_CB, _c_void] + tail_t
        vt = np.ascontiguousarray(P.vtype, np.int32)
        xl, xu = np.ascontiguousarray(P.xl, float), np.ascontiguousarray(P.xu, float)
        rc = L.bnlp_oa_nlp_optim_cb(n, m, vp(vt), vp(xl), vp(xu), vp(lb), vp(ub), vp(xlp), cbs[0], cbs[1], None, *tail)
    if failure:
        raise failure[0]
    if rc == -3:
        raise RuntimeError("Unsolved NLP ... exit")
    if rc:
        raise RuntimeError(f"bnlp_oa_nlp_optim failed ({rc})")
    return dict(cuts=_cuts_list(res.cuts, outs), solution=(res.obj, x.copy()) if res.has_solution else None, status=res.nlp_status,
                obj=res.obj, duals=du)


def ecp_rounds_native(target, lp, rounds=5, abs_violation_tol=1e-6, rel_violation_tol=1e-4, integer=None, duals=None):
    """ecp_rounds through the C++ generator (B200Oa.cpp: oa_ecp_rounds); target: an Engine (bnlp_ecp_rounds_engine) or a provider
    (bnlp_ecp_rounds_cb).  lp: what provider.extract returns.  Same result keys as ecp_rounds."""
    from . import bb as _bb, engine as _e
    is_engine = isinstance(target, _e.Engine)
    n, m = target.n, target.m
    info = target.problem.info() if is_engine else None
    xl, xu = (info["xl"], info["xu"]) if is_engine else (target.xl, target.xu)
    ncols = n + (1 if lp["has_alpha"] else 0)
    lpd = dict(rows=list(lp["rows"]), obj=lp["obj"], obj_offset=lp["obj_offset"], ncols=ncols)
    rs, col, val, lb, ub, clb, cub, obj = _lp_csr(lpd, np.concatenate([xl, [-DMAX]])[:ncols], np.concatenate([xu, [DMAX]])[:ncols])
    vp = lambda a: None if a is None else a.ctypes.data_as(_c_void)
    integ = None if integer is None else np.ascontiguousarray(integer, np.int32)
    du = None if duals is None else np.ascontiguousarray(duals, float)
    cap_cuts, cap_nnz = rounds * (m + 1), rounds * (m + 1) * (n + 1)
    outs = _cut_outs(cap_cuts, cap_nnz)
    res, x = EcpResult(), np.zeros(ncols)
    failure = []
    master = highs_master(failure)
    head = [ncols, len(lpd["rows"]), vp(rs), vp(col), vp(val), vp(lb), vp(ub), vp(clb), vp(cub), vp(obj), float(lp["obj_offset"]),
            vp(integ), vp(du), int(rounds), float(abs_violation_tol), float(rel_violation_tol)]
    head_t = [_c_int, _c_int] + [_c_void] * 8 + [_c_dbl, _c_void, _c_void, _c_int, _c_dbl, _c_dbl]
    tail = [_ct.byref(res), vp(x)] + [vp(a) for a in outs] + [cap_cuts, cap_nnz]
    tail_t = [_ct.POINTER(EcpResult)] + [_c_void] * 7 + [_c_int, _c_int]
    if is_engine:
        L = _e.lib()
        L.bnlp_ecp_rounds_engine.argtypes = [_c_void, _c_void] + head_t + [MASTER_CB, _c_void] + tail_t
        rc = L.bnlp_ecp_rounds_engine(target.h, target.problem.h, *head, master, None, *tail)
    else:
        cbs = _provider_callbacks(target, failure)
        if du is not None:
            target.duals = du
        L = _bb.host_lib()
        L.bnlp_ecp_rounds_cb.argtypes = [_c_int, _c_int] + head_t + [CUTS_CB, VIOL_CB, MASTER_CB, _c_void] + tail_t
        if du is None:
            head[12] = vp(np.ascontiguousarray(target.duals, float))
        rc = L.bnlp_ecp_rounds_cb(n, m, *head, cbs[1], cbs[2], master, None, *tail)
    if failure:
        raise failure[0]
    if rc:
        raise RuntimeError(f"bnlp_ecp_rounds failed ({rc})")
    cs = _cuts_list(res.cuts, outs)
    out = dict(rounds=res.rounds, cuts=res.cuts, installed=res.installed, cs=cs, violation=res.violation,
               obj=1.7976931348623157e308 if res.lp_infeasible else res.obj)
    if not res.lp_infeasible or res.installed:
        out["rows"] = lpd["rows"] + cs[:res.installed]
    if not res.lp_infeasible:
        out["x"] = x.copy()
    return out
