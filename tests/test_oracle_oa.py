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
