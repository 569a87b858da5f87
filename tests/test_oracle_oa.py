"""CPU suite: the evaluation / OA oracle against the reference's known answers and against finite differences."""
import numpy as np
import pytest

from bonmin_b200 import instances as I
from oracle.struct_eval import DBL_MAX, Evaluator, outer_approximation

SQ5 = 5 ** 0.5


def test_mytoy_derivatives_match_hand_coded_example():
    # examples/CppExample/MyTMINLP.cpp:103-202 (variable order x,y1,y2,z there; y1,y2,x,z here):
    # g0 = (y1-1/2)^2+(y2-1/2)^2, dg0 = (2y1-1, 2y2-1), Hessian of the Lagrangian = 2*lambda0 on (y1,y1),(y2,y2)
    m = I.mytoy()
    ev = Evaluator(m)
    x = np.array([0.3, 0.9, 0.4, 2.0])
    r = ev.sparse(x, lam=np.array([1.5, -2.0, 0.25]), sigma=3.0)
    assert r["f"] == pytest.approx(-(0.3 + 0.9 + 0.4))
    assert r["g"] == pytest.approx([(0.3 - .5) ** 2 + (0.9 - .5) ** 2, 0.4 - 0.3, 0.4 + 0.9 + 2.0])
    assert list(r["gradf"]) == [-1, -1, -1, 0]
    assert r["jac"] == pytest.approx([2 * 0.3 - 1, 2 * 0.9 - 1, -1, 1, 1, 1, 1])
    assert r["hess"] == pytest.approx([3.0, 3.0])      # 2*lambda0
    assert len(ev.jac) == 7 and len(ev.hes) == 2         # MyTMINLP.cpp:60-66


def test_mytoy_oa_row_golden():
    # test/InterfaceTest.cpp:191-264 (+ test/toy.mps.gz): row 0 = (2/sqrt5) y1 + (1/sqrt5) y2 <= 1/2 + 3/(2 sqrt5)
    m = I.mytoy()
    ev = Evaluator(m)
    x = np.array([.5 + 1 / SQ5, .5 + 1 / (2 * SQ5), .5 + 1 / SQ5, 0.])
    d = np.zeros(2 * 4 + 3)
    d[8], d[9] = SQ5, 1.0
    cuts = outer_approximation(ev, x, duals=d)
    assert len(cuts) == 1
    row, cols, vals, lb, ub = cuts[0]
    assert row == 0 and cols == [0, 1]
    assert vals == pytest.approx([2 / SQ5, 1 / SQ5], rel=1e-12)
    assert lb == -DBL_MAX
    rhs = .5 + 3 / (2 * SQ5)
    assert ub == pytest.approx(rhs + 1e-8 * max(rhs, 1.0), rel=1e-12)      # oa_rhs_relax (:2044)


def test_objective_cut_and_filters():
    m = I.syn_cvx_minlp(12, 4, 3, 4, 1)
    ev = Evaluator(m)
    x = m.x0 + 0.05
    cuts = outer_approximation(ev, x, get_obj=2)
    assert cuts[-1][0] == m.m and cuts[-1][1][-1] == m.n and cuts[-1][2][-1] == -1.0
    f, gradf = ev.sparse(x)["f"], ev.sparse(x)["gradf"]
    kept = np.abs(gradf) >= 1e-8
    assert cuts[-1][4] == pytest.approx(float(gradf[kept] @ x[kept] - f))
    # linear objective + getObj=1 -> no objective cut (:2053)
    assert all(c[0] < m.m for c in outer_approximation(ev, x, get_obj=1))
    # violation filter: with log level 0 non-violated constraint cuts are still generated
    x2 = np.append(m.x0, 0.0)
    assert len(outer_approximation(ev, x, x2=x2, theta=1e9)) == 3
    assert len(outer_approximation(ev, x, x2=x2, theta=1e9, log_level=1)) == 0


def test_term_derivatives_finite_differences():
    m = I.syn_cvx_minlp(20, 6, 4, 6, 7)
    ev = Evaluator(m)
    x = m.x0 + 0.1
    lam = np.linspace(0.5, 1.5, m.m)
    v0, g0, H0 = ev.dense(x, lam, 1.0)
    e = 1e-6
    for k in (1, 2):
        gn = np.array([(ev.dense(x + e * np.eye(m.n)[j])[0][k] - ev.dense(x - e * np.eye(m.n)[j])[0][k]) / (2 * e)
                       for j in range(m.n)])
        assert np.abs(gn - g0[k]).max() < 1e-6
    w = np.concatenate([[1.0], lam])
    Hn = np.array([((ev.dense(x + e * np.eye(m.n)[j])[1] - ev.dense(x - e * np.eye(m.n)[j])[1]) / (2 * e)).T @ w
                   for j in range(m.n)])
    assert np.abs(Hn - H0).max() < 1e-5


def test_mytoy_extract_linear_relaxation_golden():
    """test/InterfaceTest.cpp:191-264 (testOa): the OA LP at the root optimum has 4 columns, 3 rows, 7 nonzeros; row 0 =
    (2/sqrt5) y1 + (1/sqrt5) y2 <= 1/2 + 3/(2 sqrt5) (+1e-7 nonlinear-row slack, +1e-8 relative oa_rhs_relax), row 1
    <= 0 (+1e-8), row 2 <= 2 (1 + 1e-8); the linear objective is copied (no alpha column)."""
    from oracle.struct_eval import extract_linear_relaxation
    m = I.mytoy()
    ev = Evaluator(m)
    s5 = 5 ** 0.5
    x = np.array([0.5 + 1 / s5, 0.5 + 1 / (2 * s5), 0.5 + 1 / s5, 0.0])   # root optimum: y1, y2, x, z
    # the product orders variables as the model does; find y1, y2 through the nonlinear row's structure
    lp = extract_linear_relaxation(ev, _toy_point(m, s5), None, get_obj=True)
    rows = lp["rows"]
    assert len(rows) == 3 and sum(len(r[1]) for r in rows) == 7 and not lp["has_alpha"]
    nl = [r for r in rows if ev.con_nonlinear[r[0]]]
    assert len(nl) == 1
    coefs = sorted(abs(v) for v in nl[0][2])
    assert coefs == pytest.approx([1 / s5, 2 / s5], rel=1e-12)
    ub0 = 0.5 + 3 / (2 * s5)
    assert nl[0][4] == pytest.approx((ub0 + 1e-7) * (1 + 1e-8), abs=1e-12)
    assert nl[0][3] == -np.finfo(float).max
    lin = sorted(r[4] for r in rows if not ev.con_nonlinear[r[0]])
    assert lin == pytest.approx([1e-8, 2 * (1 + 1e-8)], abs=1e-15)
    assert np.count_nonzero(lp["obj"]) == 3 and lp["obj_offset"] == 0.0


def _toy_point(m, s5):
    """root optimum of mytoy in the model's own variable order (from the oracle IPM, then snapped to the analytic values)."""
    from oracle import ipm
    r = ipm.solve_batch(m, m.xl[None, :], m.xu[None, :], m.x0)
    x = r["x"][0].copy()
    for j in range(m.n):
        for target in (0.5 + 1 / s5, 0.5 + 1 / (2 * s5), 0.0):
            if abs(x[j] - target) < 1e-5:
                x[j] = target
    return x


def test_cpp_outer_approximation_matches_the_numpy_restatement(oracle_ipm):
    """oracle_oa_batch (C++, one point per host thread: the CPU baseline of the OA kernels in bench.py) against
    struct_eval.outer_approximation (numpy), both restated from BonOsiTMINLPInterface.cpp:1881-2107: same kept columns,
    coefficients, sides, with and without multipliers, for every getObj mode."""
    from bonmin_b200 import instances as I
    from oracle import struct_eval as SE
    for m in (I.mytoy(), I.syn_cvx_minlp(14, 4, 3, 4, 2), I.syn_miqcp(20, 8, 4, 4, 3), I.wide_terms_minlp()):
        ev = SE.Evaluator(m)
        rng = np.random.default_rng(1)
        X = np.clip(m.x0 + rng.normal(0, 0.3, (5, m.n)), np.maximum(m.xl, -20), np.minimum(m.xu, 20))
        X[0, : min(3, m.n)] = np.asarray(m.xl)[: min(3, m.n)] + 1e-12           # tiny gradients exercise cleanNnz
        D = rng.normal(0, 1, (5, 2 * m.n + m.m))
        for du in (None, D):
            for go in (0, 1, 2):
                got = oracle_ipm.oa_batch(m, ev, X, du, get_obj=go, threads=3)
                for b in range(5):
                    want = SE.outer_approximation(ev, X[b], None if du is None else du[b], m.xl, m.xu, get_obj=go)
                    assert [g[0] for g in got[b]] == [w[0] for w in want]
                    for g, w in zip(got[b], want):
                        assert g[1] == list(w[1])
                        np.testing.assert_allclose(g[2], w[2], rtol=1e-13, atol=1e-14)
                        sc = 1 + 20 * np.abs(np.array(w[2])).sum()
                        assert all(a == b_ or abs(a - b_) <= 1e-12 * sc for a, b_ in ((g[3], w[3]), (g[4], w[4])))
