"""Write a StructuredModel (bonmin_b200.instances) as an AMPL .nl file in the ASCII `g` format of the reference's
fixture (test/mytoy.nl), so that the same instance can be given to a real `bonmin` wherever one exists
(SURVEY.md Appendix B).  Variables keep the model's order; integrality is only recoverable from the header when the
model already follows AMPL's ordering (nonlinear first, then linear continuous | binary | integer) — the writer checks
this and refuses otherwise."""
from __future__ import annotations

import numpy as np

from .instances import BINARY, CONTINUOUS, INF, INTEGER, unpack_lower

_OPS = {2: "o0", 3: "o1", 4: "o2", 5: "o3", 8: "o5", 6: "o16", 9: "o44", 10: "o43", 11: "o39", 12: "o41", 13: "o46",
        14: "o76", 15: "o15", 16: "o49", 17: "o51", 18: "o53"}


def _f(v):
    return repr(float(v))


def _prefix(vars_, prog):
    """postfix element program -> list of prefix .nl lines"""
    st = []
    for op, ia, da in prog:
        if op == 0:
            st.append([f"n{_f(da)}"])
        elif op == 1:
            st.append([f"v{vars_[ia]}"])
        elif op in (2, 3, 4, 5, 8):
            b = st.pop()
            a = st.pop()
            st.append([_OPS[op]] + a + b)
        elif op == 7:
            a = st.pop()
            st.append(["o5"] + a + [f"n{_f(da)}"])
        else:
            a = st.pop()
            st.append([_OPS[op]] + a)
    assert len(st) == 1
    return st[0]


def write_nl(model) -> str:
    n, m = model.n, model.m
    Q = {}
    for k, fn in enumerate(model.qd_fn):
        Q[int(fn)] = unpack_lower(model.qd[k], n)
    for fn, i, j, v in zip(model.qs_fn, model.qs_i, model.qs_j, model.qs_v):
        M = Q.setdefault(int(fn), np.zeros((n, n)))
        M[i, j] += v
        if i != j:
            M[j, i] += v
    terms = {}
    for fn, vars_, prog in model.get("terms", []):
        terms.setdefault(fn, []).append((vars_, prog))
    nonlin_fn = set(Q) | set(terms)
    nl_cons = [i for i in range(m) if (i + 1) in nonlin_fn]
    # AMPL puts nonlinear constraints first; refuse models that are not ordered that way
    if nl_cons != list(range(len(nl_cons))):
        raise ValueError("nonlinear rows must come first to be written as .nl")
    in_con, in_obj = np.zeros(n, bool), np.zeros(n, bool)
    for fn in nonlin_fn:
        used = np.zeros(n, bool)
        if fn in Q:
            used |= np.abs(Q[fn]).sum(0) > 0
        for vars_, _ in terms.get(fn, []):
            used[list(vars_)] = True
        if fn == 0:
            in_obj |= used
        else:
            in_con |= used
    vt = np.asarray(model.vtype)
    both, conly, oonly = in_con & in_obj, in_con & ~in_obj, in_obj & ~in_con
    nlvb, nlvc, nlvo = int(both.sum()), int(in_con.sum()), int(both.sum() + oonly.sum())
    linear = ~(in_con | in_obj)
    # the header can only describe AMPL's canonical order
    order_key = np.where(both, 0, np.where(conly, 1, np.where(oonly, 2, 3))) * 10 + \
        np.where(linear, np.where(vt == CONTINUOUS, 0, np.where(vt == BINARY, 1, 2)), (vt != CONTINUOUS).astype(int))
    if np.any(np.diff(order_key) < 0):
        raise ValueError("variables are not in AMPL's canonical order; cannot encode integrality in the .nl header")
    nlvbi = int((both & (vt != CONTINUOUS)).sum())
    nlvci = int((conly & (vt != CONTINUOUS)).sum())
    nlvoi = int((oonly & (vt != CONTINUOUS)).sum())
    nbv = int((linear & (vt == BINARY)).sum())
    niv = int((linear & (vt == INTEGER)).sum())
    A = np.asarray(model.A, float)
    # structural Jacobian / gradient entries: linear coefficients plus variables of the nonlinear part
    struct = [set(np.flatnonzero(A[k]).tolist()) for k in range(m + 1)]
    for fn in nonlin_fn:
        if fn in Q:
            struct[fn] |= set(np.flatnonzero(np.abs(Q[fn]).sum(0) > 0).tolist())
        for vars_, _ in terms.get(fn, []):
            struct[fn] |= set(vars_)
    nzc = sum(len(struct[k]) for k in range(1, m + 1))
    nzo = len(struct[0])
    n_ranges = int(((model.gl > -INF) & (model.gu < INF) & (model.gl != model.gu)).sum())
    n_eqns = int(((model.gl > -INF) & (model.gl == model.gu)).sum())
    out = ["g3 1 1 0", f" {n} {m} 1 {n_ranges} {n_eqns}", f" {len(nl_cons)} {1 if 0 in nonlin_fn else 0}", " 0 0",
           f" {nlvc} {nlvo} {nlvb}", " 0 0 0 1", f" {nbv} {niv} {nlvbi} {nlvci} {nlvoi}", f" {nzc} {nzo}", " 0 0",
           " 0 0 0 0 0"]

    def body(fn):
        parts = []
        if model.c[fn] != 0.0:
            parts.append([f"n{_f(model.c[fn])}"])
        if fn in Q:
            M = Q[fn]
            for i in range(n):
                for j in range(i + 1):
                    if M[i, j] != 0.0:
                        coef = 0.5 * M[i, j] if i == j else M[i, j]
                        parts.append(["o2", f"n{_f(coef)}", "o2", f"v{i}", f"v{j}"])
        for vars_, prog in terms.get(fn, []):
            parts.append(_prefix(vars_, prog))
        if not parts:
            return ["n0"]
        if len(parts) == 1:
            return parts[0]
        if len(parts) == 2:
            return ["o0"] + parts[0] + parts[1]
        return ["o54", str(len(parts))] + [ln for p in parts for ln in p]

    for i in range(m):
        out.append(f"C{i}")
        out += body(i + 1)
    out.append("O0 0")
    out += body(0)
    out.append("x" + str(n))
    out += [f"{j} {_f(model.x0[j])}" for j in range(n)]
    out.append("r")
    for i in range(m):
        lo, up = model.gl[i], model.gu[i]
        if lo > -INF and up < INF:
            out.append(f"4 {_f(lo)}" if lo == up else f"0 {_f(lo)} {_f(up)}")
        elif up < INF:
            out.append(f"1 {_f(up)}")
        elif lo > -INF:
            out.append(f"2 {_f(lo)}")
        else:
            out.append("3")
    out.append("b")
    for j in range(n):
        lo, up = model.xl[j], model.xu[j]
        if lo > -INF and up < INF:
            out.append(f"4 {_f(lo)}" if lo == up else f"0 {_f(lo)} {_f(up)}")
        elif up < INF:
            out.append(f"1 {_f(up)}")
        elif lo > -INF:
            out.append(f"2 {_f(lo)}")
        else:
            out.append("3")
    colcount = np.zeros(n, int)
    for k in range(1, m + 1):
        for j in struct[k]:
            colcount[j] += 1
    out.append(f"k{n - 1}")
    out += [str(int(c)) for c in np.cumsum(colcount)[:-1]]
    for i in range(m):
        cols = sorted(struct[i + 1])
        if cols:
            out.append(f"J{i} {len(cols)}")
            out += [f"{j} {_f(A[i + 1, j])}" for j in cols]
    cols = sorted(struct[0])
    if cols:
        out.append(f"G0 {len(cols)}")
        out += [f"{j} {_f(A[0, j])}" for j in cols]
    return "\n".join(out) + "\n"
