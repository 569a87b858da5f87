"""Host cut strengthener: CutStrengthener::ComputeCuts / HandleOneCut / StrengthenCut
(src/Interfaces/BonCutStrengthener.cpp:36-304) around the batched engine.  The reference solves one small NLP per OA cut
(two with the most-fractional disjunction), one after the other; here every NLP of every cut of a row is one child of
ONE batch on the row's strengthening variant (bnlp_problem_cut_strengthening).  Host control flow only; the arithmetic
provider is `variant(ci)` -> object with solve(XL, XU, X0) -> dict(obj, status) (EngineVariants in the product; tests
also run the CPU oracle here)."""
from __future__ import annotations

import numpy as np

INFTY = 1e19                    # the NLP engine's infinity (nlp_upper_bound_inf)
DMAX = np.finfo(float).max      # COIN_DBL_MAX: an absent side of a cut


class EngineVariants:
    """variant(ci) = a CUDA engine on base_problem.cs_variant(ci), created on first use and kept."""

    def __init__(self, base_problem, **engine_kw):
        self.base, self.kw, self.cache = base_problem, engine_kw, {}

    def __call__(self, ci):
        from .engine import Engine
        if ci not in self.cache:
            eng = Engine(self.base.cs_variant(ci), **self.kw)
            self.cache[ci] = type("V", (), {"solve": staticmethod(
                lambda XL, XU, X0, eng=eng: eng.solve_batch(XL, XU, X0, want_g=False, want_duals=False))})
        return self.cache[ci]


def strengthen_cuts(variant, n, var_types, ci, row_cols, cuts, X, xl, xu, g_val=None, g_lb=None, g_ub=None,
                    disjunctive_cut_type="none"):
    """Strengthen the cuts of row `ci` (-1: objective cuts, whose rows carry the alpha column as index n).
    cuts: list of (cols, vals, lb, ub) generated at the points X[k]; xl/xu: the box to strengthen over (the original
    bounds for global cuts, the node's for local ones); row_cols: the columns of row ci (eval_grad_gi structure; for
    ci = -1 the structural nonzeros of grad f — z is added here); g_val[k]: g_i(X[k]) and g_lb/g_ub the row's bounds,
    for the tightness test of ComputeCuts (:146-160).  Returns a list of dict(cols, vals, lb, ub, changed, ok)."""
    nv = n + (1 if ci == -1 else 0)
    part = np.zeros(nv, bool)
    part[np.asarray(row_cols, int)] = True
    if ci == -1:
        part[n] = True
    xl = np.asarray(xl, float)
    xu = np.asarray(xu, float)
    jobs = []        # (cut index, tag, XL row, XU row, X0 row)
    plan = []
    for k, (cols, vals, lb, ub) in enumerate(cuts):
        cols = np.asarray(cols, int)
        vals = np.asarray(vals, float)
        if np.any(~part[cols]):
            raise ValueError(f"cut {k}: index {cols[~part[cols]][0]} does not appear in the constraint")   # exit(-99), :353-356
        lower_bound = lb > -DMAX            # :277-285
        if ci == -1:
            is_tight = True                                                          # :148-150
        elif g_val is None:
            is_tight = False
        else:
            is_tight = (lb <= -INFTY and g_ub - g_val[k] <= 1e-8) or (ub >= INFTY and g_val[k] - g_lb <= 1e-8)
        p = np.zeros(nv)
        p[cols] = vals if lower_bound else -vals
        x = np.asarray(X[k], float)

        def job(tag, lo, hi):
            XL = np.empty(2 * nv)
            XU = np.empty(2 * nv)
            XL[:n] = np.where(part[:n], lo, x)
            XU[:n] = np.where(part[:n], hi, x)
            if ci == -1:
                XL[n], XU[n] = -INFTY, INFTY                                         # :373-376
            XL[nv:] = XU[nv:] = p
            x0 = np.concatenate([x, [0.0] if ci == -1 else [], p])
            jobs.append((k, tag, XL, XU, x0))
        imost = -1
        if disjunctive_cut_type == "most-fractional":                               # :53-71
            viol = 1e-6
            for i, idx in enumerate(cols):
                if idx < n and var_types[idx] != 0:
                    v = min(x[idx] - np.floor(x[idx]), np.ceil(x[idx]) - x[idx])
                    if v > viol:
                        imost, viol = i, v
        if imost >= 0:
            idx = cols[imost]
            down = np.floor(x[idx])
            hi = xu.copy(); hi[idx] = down
            lo = xl.copy(); lo[idx] = down + 1.0
            job("down", xl, hi)
            job("up", lo, xu)
            plan.append(("disj", lower_bound, imost, down))
        elif not is_tight:
            job("one", xl, xu)
            plan.append(("one", lower_bound, -1, 0.0))
        else:
            plan.append(("skip", lower_bound, -1, 0.0))
    res = {}
    if jobs:
        r = variant(ci).solve(np.array([j[2] for j in jobs]), np.array([j[3] for j in jobs]), np.array([j[4] for j in jobs]))
        for (k, tag, *_), obj, st in zip(jobs, r["obj"], r["status"]):
            res[(k, tag)] = (float(obj), int(st) in (1, 2))
    out = []
    for k, ((cols, vals, lb, ub), (kind, lower_bound, imost, down)) in enumerate(zip(cuts, plan)):
        vals = np.array(vals, float)
        new = dict(cols=list(cols), vals=vals, lb=lb, ub=ub, changed=False, ok=True)

        def bound(tag):      # StrengthenedBound (:593-602): the optimal value, sign restored
            obj, ok = res[(k, tag)]
            return (obj if lower_bound else -obj), ok
        if kind == "one":
            b, ok = bound("one")
            new["ok"] = ok
            if ok:
                new["lb" if lower_bound else "ub"] = b
                new["changed"] = True
        elif kind == "disj":                                                       # :87-123
            bd, okd = bound("down")
            bu, oku = bound("up")
            new["ok"] = okd and oku
            if okd and oku:
                vals[imost] += bd - bu
                new["lb" if lower_bound else "ub"] = bu + (down + 1.0) * (bd - bu)
                new["changed"] = True
        out.append(new)
    return out
