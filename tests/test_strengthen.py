"""Cut strengthener: CutStrengthener::ComputeCuts / HandleOneCut / StrengthenCut
(src/Interfaces/BonCutStrengthener.cpp:36-304) on the strengthening variant of a row
(StrengtheningTNLP, :306-560 = bnlp_problem_cut_strengthening / oracle cs_model).  The reference has no test for it
(off by default, `cut_strengthening_type none`), so the oracle is pinned here by what the NLP means: the strengthened
bound is valid for every feasible point, attained at one, and equals an independent SLSQP solve of the same NLP."""
import numpy as np
import pytest

from bonmin_b200 import instances as I, strengthen


def _row(model, i):
    """dense (c, a, Q) of constraint row i (function i + 1)."""
    from bonmin_b200.instances import unpack_lower
    n = model.n
    Q = np.zeros((n, n))
    for kk, fn in enumerate(model.qd_fn):
        if int(fn) == i + 1:
            Q += unpack_lower(model.qd[kk], n)
    return model.c[i + 1], model.A[i + 1], Q


def _oa_cut(model, i, x):
    c, a, Q = _row(model, i)
    g = c + a @ x + 0.5 * x @ Q @ x
    grad = a + Q @ x
    cols = np.flatnonzero(grad)
    return (cols, grad[cols], -strengthen.DMAX, model.gu[i] - g + grad @ x), g


def _oracle_variants(model, oracle_ipm, struct_eval):
    cache = {}

    def variant(ci):
        if ci not in cache:
            vm = struct_eval.cs_model(model, ci)
            cache[ci] = type("V", (), {"solve": staticmethod(lambda XL, XU, X0, vm=vm: oracle_ipm.solve_batch(vm, XL, XU, X0, threads=4))})
        return cache[ci]
    return variant


def _feasible_samples(model, i, rng, k=20000):
    c, a, Q = _row(model, i)
    X = rng.uniform(model.xl, model.xu, size=(k, model.n))
    g = c + X @ a + 0.5 * np.einsum("bi,ij,bj->b", X, Q, X)
    return X[g <= model.gu[i]]


@pytest.fixture(scope="module")
def struct_eval():
    from oracle import struct_eval as se
    return se


def check_strengthening(model, variant, rng, slsqp=True):
    n = model.n
    i = 0                                             # a convex quadratic row  1/2 x'Qx + a.x <= b
    c, a, Q = _row(model, i)
    # cut points strictly INSIDE the row (at a violated point g_ub - g <= tol counts as tight for the reference, :153-159)
    F = _feasible_samples(model, i, rng)
    gF = c + F @ a + 0.5 * np.einsum("bi,ij,bj->b", F, Q, F)
    pts = list(F[gF <= model.gu[i] - 0.5][:4])
    assert len(pts) == 4
    cuts, gv = zip(*[_oa_cut(model, i, x) for x in pts])
    out = strengthen.strengthen_cuts(variant, n, model.vtype, i, np.arange(n), list(cuts), pts, model.xl, model.xu,
                                     g_val=gv, g_lb=model.gl[i], g_ub=model.gu[i])
    S = _feasible_samples(model, i, rng)
    from scipy.optimize import minimize
    for (cols, vals, lb, ub), new, x in zip(cuts, out, pts):
        assert new["ok"] and new["changed"] and new["lb"] == lb
        # an OA cut of a convex row is valid, so the tightest bound is not above it (equal when the row is linear along the
        # way to the maximiser — the blocks have rank 2 — up to the IPM's 1e-8 bound relaxation)
        assert new["ub"] <= ub + 1e-6 * max(1.0, abs(ub))
        row = np.zeros(n); row[cols] = vals
        if len(S):
            assert (S @ row).max() <= new["ub"] + 1e-6     # still valid
        if not slsqp:      # (SLSQP stalls on the larger instance of the GPU test; there the engine is checked against the oracle)
            continue
        ref = minimize(lambda y: -row @ y, np.clip(x, model.xl, model.xu), jac=lambda y: -row, method="SLSQP",
                       bounds=list(zip(model.xl, model.xu)),
                       constraints=[dict(type="ineq", fun=lambda y: model.gu[i] - (c + a @ y + 0.5 * y @ Q @ y),
                                         jac=lambda y: -(a + Q @ y))], options=dict(ftol=1e-12, maxiter=500))
        assert ref.success
        assert new["ub"] == pytest.approx(-ref.fun, rel=1e-5, abs=1e-5)
    return out, cuts, pts


def test_strengthen_cut_oracle(oracle_ipm, struct_eval):
    m = I.syn_miqcp(12, 6, 3, 3, 1)
    rng = np.random.default_rng(3)
    variant = _oracle_variants(m, oracle_ipm, struct_eval)
    out, cuts, pts = check_strengthening(m, variant, rng)
    assert any(o["ub"] < c[3] - 1e-3 for o, c in zip(out, cuts))   # cuts taken away from the boundary do get tighter
    # a cut generated ON the boundary is tight: ComputeCuts skips it (:146-160)
    x = pts[0]
    cut, g = _oa_cut(m, 0, x)
    r = strengthen.strengthen_cuts(variant, m.n, m.vtype, 0, np.arange(m.n), [cut], [x], m.xl, m.xu, g_val=[m.gu[0] - 1e-9],
                                   g_lb=m.gl[0], g_ub=m.gu[0])
    assert not r[0]["changed"] and r[0]["ub"] == cut[3]
    # an index outside the row is the reference's exit(-99)
    with pytest.raises(ValueError):
        strengthen.strengthen_cuts(variant, m.n, m.vtype, 0, np.arange(m.n - 1), [cut], [x], m.xl, m.xu)


def test_disjunctive_strengthening_oracle(oracle_ipm, struct_eval):
    """DC_MostFractional (:53-123): strengthen on both sides of the most fractional integer of the cut and lift its
    coefficient; the lifted cut must hold for feasible points on either side of the disjunction."""
    m = I.syn_miqcp(12, 6, 3, 3, 1)
    rng = np.random.default_rng(4)
    variant = _oracle_variants(m, oracle_ipm, struct_eval)
    x = np.clip(m.x0 + 0.37 + 0.3 * rng.standard_normal(m.n), m.xl, m.xu)
    cut, g = _oa_cut(m, 0, x)
    r = strengthen.strengthen_cuts(variant, m.n, m.vtype, 0, np.arange(m.n), [cut], [x], m.xl, m.xu, g_val=[g], g_lb=m.gl[0],
                                   g_ub=m.gu[0], disjunctive_cut_type="most-fractional")[0]
    assert r["ok"] and r["changed"]
    row = np.zeros(m.n); row[r["cols"]] = r["vals"]
    changed = np.flatnonzero(np.abs(np.asarray(r["vals"]) - cut[1]) > 1e-12)
    assert len(changed) <= 1 and (len(changed) == 0 or m.vtype[cut[0][changed[0]]] != 0)
    S = _feasible_samples(m, 0, rng, 40000)
    S[:, :6] = np.round(S[:, :6])                                   # integer points (then re-check feasibility)
    c, a, Q = _row(m, 0)
    S = S[c + S @ a + 0.5 * np.einsum("bi,ij,bj->b", S, Q, S) <= m.gu[0]]
    assert len(S) > 10 and (S @ row).max() <= r["ub"] + 1e-6
    # objective cut (ci = -1, alpha column = index n): only the disjunction strengthens it (is_tight is forced, :148-150)
    c0, a0 = m.c[0], m.A[0]
    from bonmin_b200.instances import unpack_lower
    Q0 = sum(unpack_lower(m.qd[k], m.n) for k, fn in enumerate(m.qd_fn) if int(fn) == 0)
    f = lambda y: c0 + a0 @ y + 0.5 * y @ Q0 @ y                    # noqa: E731
    gf = a0 + Q0 @ x
    ocut = (np.append(np.flatnonzero(gf), m.n), np.append(gf[np.flatnonzero(gf)], -1.0), -strengthen.DMAX, gf @ x - f(x))
    r0 = strengthen.strengthen_cuts(variant, m.n, m.vtype, -1, np.flatnonzero(gf), [ocut], [x], m.xl, m.xu)[0]
    assert not r0["changed"]
    r1 = strengthen.strengthen_cuts(variant, m.n, m.vtype, -1, np.flatnonzero(gf), [ocut], [x], m.xl, m.xu,
                                    disjunctive_cut_type="most-fractional")[0]
    assert r1["ok"] and r1["changed"] and r1["ub"] <= 1e300
    row = np.zeros(m.n + 1); row[r1["cols"]] = r1["vals"]
    Y = rng.uniform(m.xl, m.xu, size=(5000, m.n)); Y[:, :6] = np.round(Y[:, :6])
    lhs = Y @ row[:m.n] + row[m.n] * np.array([f(y) for y in Y])   # z = f(y) is the tightest feasible z
    assert lhs.max() <= r1["ub"] + 1e-6


@pytest.mark.gpu
def test_strengthen_cut_engine_matches_oracle(bnlp_lib, oracle_ipm, struct_eval):
    from bonmin_b200.engine import Problem
    m = I.syn_miqcp(40, 20, 8, 8, 3)
    ev = strengthen.EngineVariants(Problem.from_model(m))
    ov = _oracle_variants(m, oracle_ipm, struct_eval)
    out_e, cuts, pts = check_strengthening(m, ev, np.random.default_rng(3), slsqp=False)
    out_o, _, _ = check_strengthening(m, ov, np.random.default_rng(3), slsqp=False)
    for a, b in zip(out_e, out_o):
        assert a["ok"] == b["ok"] and a["ub"] == pytest.approx(b["ub"], rel=1e-6, abs=1e-6)
    x = np.clip(m.x0 + 0.37, m.xl, m.xu)
    cut, g = _oa_cut(m, 1, x)
    kw = dict(g_val=[g], g_lb=m.gl[1], g_ub=m.gu[1], disjunctive_cut_type="most-fractional")
    a = strengthen.strengthen_cuts(ev, m.n, m.vtype, 1, np.arange(m.n), [cut], [x], m.xl, m.xu, **kw)[0]
    b = strengthen.strengthen_cuts(ov, m.n, m.vtype, 1, np.arange(m.n), [cut], [x], m.xl, m.xu, **kw)[0]
    assert a["ok"] and b["ok"] and a["ub"] == pytest.approx(b["ub"], rel=1e-6, abs=1e-6)
    np.testing.assert_allclose(a["vals"], b["vals"], rtol=1e-6, atol=1e-6)
