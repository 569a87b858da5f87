"""CPU oracle (TEST INFRASTRUCTURE ONLY) for model evaluation and outer-approximation cut assembly.

* `Evaluator`  restates what the reference obtains from TMINLP::eval_f/eval_grad_f/eval_g/eval_jac_g/eval_h
  (src/Interfaces/BonTMINLP.hpp:268-315; for .nl models forwarded to ASL, src/Interfaces/Ampl/BonAmplTMINLP.cpp:629-661)
  on a StructuredModel, in plain numpy / Python.
* `outer_approximation` restates OsiTMINLPInterface::getOuterApproximation line by line
  (src/Interfaces/BonOsiTMINLPInterface.cpp:1924-2107) including cleanNnz (:1881-1919) and its quirks
  (dead lb relaxation, two infinities, objective cut with hard-coded 1e-15).
Pinned against the reference's known answers for test/mytoy.nl (test/InterfaceTest.cpp:191-264: OA row
(2/sqrt5, 1/sqrt5 | 1/2 + 3/(2 sqrt5)))."""
from __future__ import annotations

import math
import sys

import numpy as np

DBL_MAX = sys.float_info.max

# opcodes: include/bnlp.h
(OP_CONST, OP_VAR, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_POWC, OP_POW, OP_EXP, OP_LOG, OP_SQRT, OP_SIN, OP_COS,
 OP_SQR, OP_ABS, OP_ATAN, OP_ASIN, OP_ACOS) = range(19)


class _HD:
    """value / gradient / Hessian triple w.r.t. the local variables of a term (second-order forward mode)."""
    __slots__ = ("v", "g", "h")

    def __init__(self, v, g, h):
        self.v, self.g, self.h = v, g, h

    def unary(self, f, f1, f2):
        return _HD(f, f1 * self.g, f2 * np.outer(self.g, self.g) + f1 * self.h)


def _mul(a, b):
    return _HD(a.v * b.v, a.v * b.g + b.v * a.g, a.v * b.h + b.v * a.h + np.outer(a.g, b.g) + np.outer(b.g, a.g))


def eval_term(vars_, prog, x):
    nv = len(vars_)
    st = []
    for op, ia, da in prog:
        if op == OP_CONST:
            st.append(_HD(float(da), np.zeros(nv), np.zeros((nv, nv))))
        elif op == OP_VAR:
            g = np.zeros(nv)
            g[ia] = 1.0
            st.append(_HD(float(x[vars_[ia]]), g, np.zeros((nv, nv))))
        elif op in (OP_ADD, OP_SUB):
            b = st.pop()
            a = st.pop()
            s = 1.0 if op == OP_ADD else -1.0
            st.append(_HD(a.v + s * b.v, a.g + s * b.g, a.h + s * b.h))
        elif op == OP_MUL:
            b = st.pop()
            a = st.pop()
            st.append(_mul(a, b))
        elif op == OP_DIV:
            b = st.pop()
            a = st.pop()
            r = 1.0 / b.v
            st.append(_mul(a, b.unary(r, -r * r, 2 * r ** 3)))
        elif op == OP_POW:
            b = st.pop()
            a = st.pop()
            t = _mul(a.unary(math.log(a.v), 1 / a.v, -1 / a.v ** 2), b)
            e = math.exp(t.v)
            st.append(t.unary(e, e, e))
        else:
            a = st.pop()
            v = a.v
            if op == OP_NEG:
                st.append(a.unary(-v, -1.0, 0.0))
            elif op == OP_SQR:
                st.append(a.unary(v * v, 2 * v, 2.0))
            elif op == OP_POWC:
                p = da
                f1 = 0.0 if p == 0 else p * v ** (p - 1)
                f2 = 0.0 if p in (0.0, 1.0) else p * (p - 1) * v ** (p - 2)
                st.append(a.unary(v ** p, f1, f2))
            elif op == OP_EXP:
                e = math.exp(v)
                st.append(a.unary(e, e, e))
            elif op == OP_LOG:
                st.append(a.unary(math.log(v), 1 / v, -1 / v ** 2))
            elif op == OP_SQRT:
                s = math.sqrt(v)
                st.append(a.unary(s, 0.5 / s, -0.25 / (s * v)))
            elif op == OP_SIN:
                st.append(a.unary(math.sin(v), math.cos(v), -math.sin(v)))
            elif op == OP_COS:
                st.append(a.unary(math.cos(v), -math.sin(v), -math.cos(v)))
            elif op == OP_ABS:
                st.append(a.unary(abs(v), 1.0 if v >= 0 else -1.0, 0.0))
            elif op == OP_ATAN:
                q = 1.0 / (1.0 + v * v)
                st.append(a.unary(math.atan(v), q, -2.0 * v * q * q))
            elif op == OP_ASIN:
                q = 1.0 / math.sqrt(1.0 - v * v)
                st.append(a.unary(math.asin(v), q, v * q ** 3))
            elif op == OP_ACOS:
                q = 1.0 / math.sqrt(1.0 - v * v)
                st.append(a.unary(math.acos(v), -q, -v * q ** 3))
            else:
                raise ValueError(op)
    assert len(st) == 1
    return st[0]


class Evaluator:
    """Dense reference evaluation of a StructuredModel at one point."""

    def __init__(self, model):
        self.m_ = model
        n, m = model.n, model.m
        self.Q = {}
        il = np.tril_indices(n)
        for k, fn in enumerate(model.qd_fn):
            Q = np.zeros((n, n))
            Q[il] = model.qd[k]
            self.Q[int(fn)] = Q + Q.T - np.diag(np.diag(Q))
        for fn, i, j, v in zip(model.qs_fn, model.qs_i, model.qs_j, model.qs_v):
            Q = self.Q.setdefault(int(fn), np.zeros((n, n)))
            Q[i, j] += v
            if i != j:
                Q[j, i] += v
        self.terms = model.get("terms", [])
        # sparsity structures exactly as the product derives them (HostModel::finalize)
        cols = [set(np.flatnonzero(model.A[k]).tolist()) for k in range(m + 1)]
        hes = set()
        nonlin = np.zeros(m + 1, bool)
        for fn in model.qd_fn:
            cols[fn] |= set(range(n))
            nonlin[fn] = True
        if len(model.qd_fn):
            hes |= {(i, j) for i in range(n) for j in range(i + 1)}
        for fn, i, j in zip(model.qs_fn, model.qs_i, model.qs_j):
            cols[fn] |= {int(i), int(j)}
            hes.add((max(i, j), min(i, j)))
            nonlin[fn] = True
        for fn, vars_, _ in self.terms:
            nonlin[fn] = True
            cols[fn] |= set(vars_)
            hes |= {(max(a, b), min(a, b)) for a in vars_ for b in vars_}
        self.jac = [(i, j) for i in range(m) for j in sorted(cols[i + 1])]
        self.hes = sorted(hes)
        self.con_nonlinear = nonlin[1:]
        self.obj_nonlinear = bool(nonlin[0])

    def dense(self, x, lam=None, sigma=1.0):
        """values (m+1), dense gradients (m+1, n), dense Lagrangian Hessian (n, n)."""
        M = self.m_
        n, m = M.n, M.m
        x = np.asarray(x, float)
        vals = M.c + M.A @ x
        grads = np.array(M.A, float).copy()
        w = np.zeros(m + 1)
        w[0] = sigma
        if lam is not None:
            w[1:] = lam
        H = np.zeros((n, n))
        for fn, Q in self.Q.items():
            u = Q @ x
            vals[fn] += 0.5 * x @ u
            grads[fn] += u
            H += w[fn] * Q
        for fn, vars_, prog in self.terms:
            r = eval_term(vars_, prog, x)
            vals[fn] += r.v
            for a, va in enumerate(vars_):
                grads[fn, va] += r.g[a]
                for b, vb in enumerate(vars_):
                    H[va, vb] += w[fn] * r.h[a, b]
        return vals, grads, H

    def sparse(self, x, lam=None, sigma=1.0):
        vals, grads, H = self.dense(x, lam, sigma)
        jv = np.array([grads[i + 1, j] for i, j in self.jac])
        hv = np.array([H[i, j] for i, j in self.hes])
        return dict(f=vals[0], g=vals[1:], gradf=grads[0], jac=jv, hess=hv)


def clean_nnz(value, colLower, colUpper, rowLower, rowUpper, colsol, lbub, tiny, veryTiny, infty):
    """BonOsiTMINLPInterface.cpp:1881-1919; lbub = [lb, ub] is updated in place; returns keep flag."""
    if abs(value) >= tiny:
        return True
    if abs(value) < veryTiny:
        return False
    colUpBounded = colUpper < infty
    colLoBounded = colLower > -infty
    rowNotLoBounded = rowLower <= -infty
    rowNotUpBounded = rowUpper >= infty
    pos = value > 0
    if colUpBounded and pos and rowNotUpBounded:
        lbub[0] += value * (colsol - colUpper)
        return False
    if colUpBounded and not pos and rowNotLoBounded:
        lbub[1] += value * (colsol - colUpper)
        return False
    if colLoBounded and not pos and rowNotUpBounded:
        lbub[0] += value * (colsol - colLower)
        return False
    if colLoBounded and pos and rowNotLoBounded:
        lbub[1] += value * (colsol - colLower)
        return False
    return True


def outer_approximation(ev: Evaluator, x, duals=None, xl=None, xu=None, x2=None, theta=0.0, get_obj=0,
                        tiny=1e-8, very_tiny=1e-17, rhs_relax=1e-8, infty_=1e19, log_level=0):
    """Returns a list of cuts (src_row, cols, vals, lb, ub); src_row == m is the objective cut.
    `duals` is the full [zL|zU|lambda] vector or None (then the ranged-row side selection is skipped)."""
    M = ev.m_
    n, m = M.n, M.m
    infty = DBL_MAX                       # getInfinity() = COIN_DBL_MAX (:1256-1260)
    xl = M.xl if xl is None else xl
    xu = M.xu if xu is None else xu
    r = ev.sparse(x)
    g, jv = r["g"], r["jac"]
    cuts = []
    rows = {}
    for i in range(m):
        if not ev.con_nonlinear[i]:
            continue
        lb = M.gl[i] - g[i] if M.gl[i] > -infty_ else -infty
        ub = M.gu[i] - g[i] if M.gu[i] < infty_ else infty
        if duals is not None and M.gl[i] > -infty and M.gu[i] < infty:      # :1968-1974
            d = duals[2 * n + i]
            if d >= 0:
                lb = -infty
            if d <= 0:
                ub = infty
        rows[i] = dict(cols=[], vals=[], lbub=[lb, ub])
    for (i, j), v in zip(ev.jac, jv):                                         # :1981-1999
        if i not in rows:
            continue
        c = rows[i]
        if clean_nnz(v, xl[j], xu[j], M.gl[i], M.gu[i], x[j], c["lbub"], tiny, very_tiny, infty_):
            c["cols"].append(j)
            c["vals"].append(v)
            if c["lbub"][0] > -infty:
                c["lbub"][0] += v * x[j]
            if c["lbub"][1] < infty:
                c["lbub"][1] += v * x[j]
    for i, c in rows.items():
        lb, ub = c["lbub"]
        if x2 is not None:                                                    # :2011-2023
            rhs = float(np.dot(c["vals"], np.asarray(x2)[c["cols"]])) if c["cols"] else 0.0
            violation = max(0.0, rhs - ub, lb - rhs)
            if violation < theta and log_level > 0:
                continue
        if lb > infty:                                                        # dead code in the reference (:2043)
            lb -= rhs_relax * max(abs(lb), 1.0)
        if ub < infty:
            ub += rhs_relax * max(abs(ub), 1.0)
        cuts.append((i, list(c["cols"]), list(c["vals"]), lb, ub))
    if get_obj == 2 or (get_obj and ev.obj_nonlinear):                        # :2053-2105
        f, obj = r["f"], r["gradf"]
        lbub = [-f, -f]
        cols, vals = [], []
        for i in range(n):
            if clean_nnz(obj[i], xl[i], xu[i], -infty, 0.0, x[i], lbub, tiny, 1e-15, infty_):
                cols.append(i)
                vals.append(obj[i])
                lbub[0] += obj[i] * x[i]
                lbub[1] += obj[i] * x[i]
        cols.append(n)
        vals.append(-1.0)
        gen = True
        if x2 is not None:
            rhs = float(np.dot(vals, np.asarray(x2)[cols]))
            if max(0.0, rhs - lbub[1]) < theta:
                gen = False
        if gen:
            cuts.append((m, cols, vals, -DBL_MAX, lbub[1]))
    return cuts


def constraints_violation(ev: Evaluator, x, obj=None):
    """getConstraintsViolation / getNonLinearitiesViolation (BonOsiTMINLPInterface.cpp:1839-1876)."""
    M = ev.m_
    vals = ev.dense(x)[0]
    norm = 0.0
    for i in range(M.m):
        if not ev.con_nonlinear[i]:
            continue
        rv = 0.0
        if M.gl[i] > -1e10:
            rv = max(0.0, M.gl[i] - vals[i + 1])
        if M.gu[i] < 1e10:
            rv = max(rv, vals[i + 1] - M.gu[i])
        norm = max(norm, rv)
    if obj is not None and vals[0] - obj > norm:
        norm = vals[0] - obj
    return norm, vals[0]


def benders_cut(ev: Evaluator, x, duals, xl=None, xu=None, with_objective=True, tiny=1e-8, very_tiny=1e-17,
                infty_=1e19):
    """getBendersCut (BonOsiTMINLPInterface.cpp:2111-2202) -> (cols, vals, ub); lb is -COIN_DBL_MAX."""
    M = ev.m_
    n, m = M.n, M.m
    xl = M.xl if xl is None else xl
    xu = M.xu if xu is None else xu
    r = ev.sparse(x)
    g, jv = r["g"], r["jac"]
    lam = duals[2 * n:]
    cut = np.zeros(n + 1)
    lbub = [0.0, 0.0]
    keep = np.zeros(m, bool)
    for i in range(m):
        if ev.con_nonlinear[i] and abs(lam[i]) > 1e-6:
            keep[i] = True
            lbub[1] += lam[i] * ((M.gl[i] if lam[i] < 0 else M.gu[i]) - g[i])
    for (i, j), v in zip(ev.jac, jv):
        if not keep[i]:
            continue
        coeff = lam[i] * v
        if clean_nnz(coeff, xl[j], xu[j], M.gl[i], M.gu[i], x[j], lbub, tiny, very_tiny, infty_):
            cut[j] += coeff
            lbub[1] += coeff * x[j]
    cols, vals = [], []
    if ev.obj_nonlinear and with_objective:
        lbub[1] = -r["f"]
        for i in range(n):
            if clean_nnz(r["gradf"][i], xl[i], xu[i], -DBL_MAX, 0.0, x[i], lbub, tiny, 1e-15, infty_):
                cut[i] += r["gradf"][i]
                lbub[1] += r["gradf"][i] * x[i]
        cols.append(n)
        vals.append(-1.0)
    for i in range(n):
        if abs(cut[i]) > 1e-20:
            cols.append(i)
            vals.append(cut[i])
    return cols, np.array(vals), lbub[1]


def branching_tqp_model(model, ev: Evaluator, x_sol, duals_sol):
    """CPU restatement of Bonmin::BranchingTQP (src/Interfaces/BonBranchingTQP.cpp:22-108,136-153,192-243): the node NLP
    frozen at (x_sol, duals_sol) as a QP in d = x - x_sol.  Returns (StructuredModel of the QP, f(x_sol)).
    Test infrastructure only."""
    from bonmin_b200.instances import StructuredModel, pack_lower, _empty_sparse
    n, m = model.n, model.m
    x = np.asarray(x_sol, float)
    lam = np.asarray(duals_sol, float)[2 * n:]
    vals, grads, W = ev.dense(x, lam, 1.0)
    g = vals[1:]
    INF = 1e19
    xl = np.where(model.xl > -INF, model.xl - x, model.xl)
    xu = np.where(model.xu < INF, model.xu - x, model.xu)
    gl = np.where(model.gl > -INF, model.gl - g, model.gl)
    gu = np.where(model.gu < INF, model.gu - g, model.gu)
    A = np.array(grads, float)
    q = StructuredModel(n=n, m=m, xl=xl, xu=xu, gl=gl, gu=gu, vtype=np.asarray(model.vtype, np.int32).copy(),
                        x0=np.zeros(n), c=np.zeros(m + 1), A=A, qd_fn=np.array([0], np.int32),
                        qd=pack_lower(W).reshape(1, n * (n + 1) // 2), terms=[], name="TQP(" + str(model.get("name", "")) + ")")
    return _empty_sparse(q), float(vals[0])


def extract_linear_relaxation(ev: Evaluator, x, duals=None, xl=None, xu=None, get_obj=True, tiny=1e-8, very_tiny=1e-17,
                              rhs_relax=1e-8, infty_=1e19):
    """CPU restatement of OsiTMINLPInterface::extractLinearRelaxation + addObjectiveFunction
    (src/Interfaces/BonOsiTMINLPInterface.cpp:2378-2590).  Rows as (src_row, cols, vals, lb, ub).  Deviation shared with
    the product: bounds of removed rows are removed with them (the reference leaves its bound arrays misaligned after
    CoinPackedMatrix::deleteRows, :2494-2511).  Test infrastructure only."""
    M = ev.m_
    n, m = M.n, M.m
    DMAX = np.finfo(float).max
    x = np.asarray(x, float)
    xl = M.xl if xl is None else xl
    xu = M.xu if xu is None else xu
    vals, grads, _ = ev.dense(x)
    g = vals[1:]
    rows = []
    jac_by_row = {}
    for (i, j) in ev.jac:
        jac_by_row.setdefault(i, []).append(j)
    for i in range(m):
        rl, ru = M.gl[i], M.gu[i]
        du = 0.0 if duals is None else duals[2 * n + i]
        cols = jac_by_row.get(i, [])
        v = [grads[i + 1, j] for j in cols]
        if ev.con_nonlinear[i]:
            if rl > -infty_ and ru < infty_ and abs(du) == 0.0:
                continue
            lb = (rl - g[i]) - 1e-07 if rl > -infty_ else -DMAX
            ub = (ru - g[i]) + 1e-07 if ru < infty_ else DMAX
            if rl > -infty_ and ru < infty_:
                if du >= 0:
                    lb = -DMAX
                if du <= 0:
                    ub = DMAX
            lbub = [lb, ub]
            for j, vj in zip(cols, v):
                if clean_nnz(vj, xl[j], xu[j], rl, ru, x[j], lbub, tiny, very_tiny, infty_):
                    if lbub[0] > -DMAX:
                        lbub[0] += vj * x[j]
                    if lbub[1] < DMAX:
                        lbub[1] += vj * x[j]
            lb, ub = lbub
        else:
            lb = rl if rl > -infty_ else -DMAX
            ub = ru if ru < infty_ else DMAX
        if ub < DMAX:
            ub += rhs_relax * max(abs(ub), 1.0)
        rows.append((i, list(cols), v, lb, ub))
    out = dict(rows=rows, has_alpha=False, obj=np.zeros(n), obj_offset=0.0)
    if get_obj:
        if not ev.obj_nonlinear:
            out["obj"] = grads[0].copy()
            out["obj_offset"] = -float(ev.dense(np.zeros(n))[0][0])
        else:
            lbub = [-1e300, -vals[0]]
            cols, v = [], []
            has_rows = len(rows) > 0
            for j in range(n):
                if clean_nnz(grads[0, j], xl[j], xu[j], -DMAX, 0.0, x[j], lbub, tiny if has_rows else 1e-3,
                             very_tiny if has_rows else 1e-8, infty_):
                    cols.append(j)
                    v.append(grads[0, j])
                    lbub[0] += grads[0, j] * x[j]
                    lbub[1] += grads[0, j] * x[j]
            cols.append(n)
            v.append(-1.0)
            rows.append((m, cols, v, lbub[0], lbub[1]))
            out["has_alpha"] = True
            out["obj"] = np.concatenate([np.zeros(n), [1.0]])
    return out


def constraint_outer_approximation(ev: Evaluator, row, x, duals=None, xl=None, xu=None, tiny=1e-8, very_tiny=1e-17,
                                   infty_=1e19):
    """CPU restatement of OsiTMINLPInterface::getConstraintOuterApproximation (BonOsiTMINLPInterface.cpp:2207-2275).
    Returns (cols, vals, lb, ub).  Test infrastructure only."""
    M = ev.m_
    n = M.n
    xl = M.xl if xl is None else xl
    xu = M.xu if xu is None else xu
    vals, grads, _ = ev.dense(x)
    g = vals[1 + row]
    rl, ru = M.gl[row], M.gu[row]
    lbub = [rl - g if rl > -infty_ else -DBL_MAX, ru - g if ru < infty_ else DBL_MAX]
    if rl > -DBL_MAX and ru < DBL_MAX:
        du = 0.0 if duals is None else duals[2 * n + row]
        if du >= 0:
            lbub[0] = -DBL_MAX
        if du <= 0:
            lbub[1] = DBL_MAX
    cols, out = [], []
    for (i, j) in ev.jac:
        if i != row:
            continue
        v = grads[1 + row, j]
        if clean_nnz(v, xl[j], xu[j], rl, ru, x[j], lbub, tiny, very_tiny, infty_):
            cols.append(j)
            out.append(v)
            if lbub[0] > -DBL_MAX:
                lbub[0] += v * x[j]
            if lbub[1] < DBL_MAX:
                lbub[1] += v * x[j]
    return cols, out, lbub[0], lbub[1]


def fp_model(model, inds, norm=2, lam=1.0, sigma=1.0, scaling=1.0, use_fp_objective=True, cutoff=None, rhs_local_branching=None):
    """CPU restatement of Bonmin::TNLP2FPNLP (src/Interfaces/BonTNLP2FPNLP.cpp:75-330) as a StructuredModel with the
    rounding target as len(inds) extra variables (fixed per child through bounds).  Test infrastructure only."""
    from bonmin_b200.instances import StructuredModel, pack_lower, unpack_lower, _empty_sparse, OP_CONST, OP_MUL
    n, m, k = model.n, model.m, len(inds)
    n2 = n + k
    madd = (cutoff is not None) + (rhs_local_branching is not None)
    m2 = m + madd
    fs = (1.0 - lam) * sigma if use_fp_objective else 1.0
    ds = scaling * lam if use_fp_objective else 0.0
    Q = {}
    for kk, fn in enumerate(model.qd_fn):
        Q[int(fn)] = unpack_lower(model.qd[kk], n)
    for fn, i, j, v in zip(model.qs_fn, model.qs_i, model.qs_j, model.qs_v):
        q = Q.setdefault(int(fn), np.zeros((n, n)))
        q[i, j] += v
        if i != j:
            q[j, i] += v
    A = np.zeros((m2 + 1, n2))
    A[:m + 1, :n] = model.A
    c = np.zeros(m2 + 1)
    c[:m + 1] = model.c
    Q2 = {}
    for fn, q in Q.items():
        z = np.zeros((n2, n2))
        z[:n, :n] = q
        Q2[fn] = z
    A[0] *= fs
    c[0] *= fs
    if 0 in Q2:
        Q2[0] *= fs

    def add_dist(fn, w, nrm):
        q = Q2.setdefault(fn, np.zeros((n2, n2)))
        for i, xi in enumerate(inds):
            pi = n + i
            if nrm == 2:
                q[xi, xi] += 2 * w
                q[pi, pi] += 2 * w
                q[xi, pi] -= 2 * w
                q[pi, xi] -= 2 * w
            else:
                A[fn, xi] += w
                A[fn, pi] += w
                q[xi, pi] -= 2 * w
                q[pi, xi] -= 2 * w
    if use_fp_objective:
        add_dist(0, ds, norm)
    terms = []
    for fn, vars_, prog in model.get("terms", []):
        sc = fs if fn == 0 else 1.0
        p2 = list(prog) + ([(OP_CONST, 0, sc), (OP_MUL, 0, 0.0)] if sc != 1.0 else [])
        terms.append((fn, list(vars_), p2))
    gl = list(model.gl)
    gu = list(model.gu)
    row = m
    if cutoff is not None:
        row += 1
        A[row] = A[0]
        c[row] = c[0]
        if 0 in Q2:
            Q2[row] = Q2[0].copy()
        for fn, vars_, prog in list(terms):
            if fn == 0:
                terms.append((row, list(vars_), list(prog)))
        gl.append(-1e19)
        gu.append(cutoff)
    if rhs_local_branching is not None:
        row += 1
        add_dist(row, 1.0, 1)
        gl.append(-1e19)
        gu.append(rhs_local_branching)
    fns = sorted(Q2)
    qd = np.array([pack_lower(Q2[f]) for f in fns]).reshape(len(fns), n2 * (n2 + 1) // 2)
    INF = 1e19
    return _empty_sparse(StructuredModel(
        n=n2, m=m2, xl=np.concatenate([model.xl, np.full(k, -INF)]), xu=np.concatenate([model.xu, np.full(k, INF)]),
        gl=np.array(gl, float), gu=np.array(gu, float), vtype=np.concatenate([model.vtype, np.zeros(k, np.int32)]).astype(np.int32),
        x0=np.concatenate([model.x0, np.zeros(k)]), c=c, A=A, qd_fn=np.array(fns, np.int32), qd=qd, terms=terms,
        name="FP(" + str(model.get("name", "")) + ")"))


def cs_model(model, constr_index):
    """CPU restatement of CutStrengthener::StrengtheningTNLP (src/Interfaces/BonCutStrengthener.cpp:306-560) as a
    StructuredModel:  min p.x  s.t.  g_l <= g_i(x) <= g_u  (constr_index = -1:  min p.(x, z)  s.t.  f(x) - z <= 0), the cut
    coefficients p as nv extra variables fixed per child through bounds (nv = n, or n + 1 with z).  Test infrastructure
    only; mirrors bnlp_problem_cut_strengthening."""
    from bonmin_b200.instances import StructuredModel, pack_lower, unpack_lower, _empty_sparse
    n = model.n
    ci = int(constr_index)
    if not -1 <= ci < model.m:
        raise ValueError("constraint index out of range")
    nv = n + (1 if ci == -1 else 0)
    n2 = 2 * nv
    src = ci + 1
    Q0 = np.zeros((n2, n2))
    for j in range(nv):
        Q0[nv + j, j] = Q0[j, nv + j] = 1.0
    Q1 = np.zeros((n2, n2))
    for kk, fn in enumerate(model.qd_fn):
        if int(fn) == src:
            Q1[:n, :n] += unpack_lower(model.qd[kk], n)
    for fn, i, j, v in zip(model.qs_fn, model.qs_i, model.qs_j, model.qs_v):
        if int(fn) == src:
            Q1[i, j] += v
            if i != j:
                Q1[j, i] += v
    A = np.zeros((2, n2))
    A[1, :n] = model.A[src]
    if ci == -1:
        A[1, n] = -1.0
    INF = 1e19
    xl = np.full(n2, -INF)
    xu = np.full(n2, INF)
    xl[:n], xu[:n] = model.xl, model.xu
    x0 = np.zeros(n2)
    x0[:n] = model.x0
    vt = np.zeros(n2, np.int32)
    vt[:n] = model.vtype
    fns, blocks = [0], [pack_lower(Q0)]
    if np.any(Q1):
        fns.append(1)
        blocks.append(pack_lower(Q1))
    terms = [(1, list(v), list(p)) for fn, v, p in model.get("terms", []) if fn == src]
    return _empty_sparse(StructuredModel(
        n=n2, m=1, xl=xl, xu=xu, gl=np.array([-INF if ci == -1 else model.gl[ci]], float),
        gu=np.array([0.0 if ci == -1 else model.gu[ci]], float), vtype=vt, x0=x0, c=np.array([0.0, model.c[src]]), A=A,
        qd_fn=np.array(fns, np.int32), qd=np.array(blocks).reshape(len(fns), n2 * (n2 + 1) // 2), terms=terms,
        name=f"CS{ci}(" + str(model.get("name", "")) + ")"))
