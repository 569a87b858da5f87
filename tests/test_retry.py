"""Retry policies of OsiTMINLPInterface: resolveForRobustness / resolveForCost with the restarts as one
batch of per-child start points."""
import numpy as np
import pytest

from bonmin_b200 import bb, instances as I


def _starts_solver(model, oracle_ipm, fail_first=0):
    dense = oracle_ipm.dense_arrays(model)
    calls = []

    def solve(XL, XU, X0):
        r = oracle_ipm.solve_batch(model, XL, XU, X0, threads=4, dense=dense)
        calls.append(X0.copy())
        r["status"][:fail_first] = -3          # pretend the first attempts hit the iteration limit (isAbandoned)
        return r
    solve.calls = calls
    return solve


def test_resolve_for_robustness_ladder(oracle_ipm):
    m = I.syn_miqcp(12, 6, 3, 3, 4)
    base = oracle_ipm.solve_batch(m, m.xl, m.xu, m.x0)
    # nothing fails: the user's start point wins, all 1 + 3 attempts were one batch
    s = _starts_solver(m, oracle_ipm)
    r = bb.resolve_callback(m, s, "robustness", m.x0, 3)
    assert r["used"] == 0 and r["nlp_solves"] == 4 and len(s.calls) == 1 and s.calls[0].shape == (4, m.n)
    np.testing.assert_allclose(s.calls[0][0], m.x0)
    assert r["status"] == 1 and r["obj"] == pytest.approx(base["obj"][0], rel=1e-9)
    # random points respect the bounds (randomStartingPoint :1768-1777)
    assert (s.calls[0][1:] >= m.xl - 1e-12).all() and (s.calls[0][1:] <= m.xu + 1e-12).all()
    # the first two attempts abandoned: the second random point is kept; convex problem, same optimum
    s = _starts_solver(m, oracle_ipm, fail_first=2)
    r = bb.resolve_callback(m, s, "robustness", m.x0, 3)
    assert r["used"] == 2 and r["num_failed"] == 2 and r["status"] == 1
    assert r["obj"] == pytest.approx(base["obj"][0], rel=1e-7)
    # everything abandoned: used = -1 and the status stays abandoned (the caller throws / pretends infeasible)
    s = _starts_solver(m, oracle_ipm, fail_first=4)
    r = bb.resolve_callback(m, s, "robustness", m.x0, 3)
    assert r["used"] == -1 and r["status"] == -3


def test_resolve_for_cost(oracle_ipm):
    m = I.syn_miqcp(12, 6, 3, 3, 4)
    base = oracle_ipm.solve_batch(m, m.xl, m.xu, m.x0)
    s = _starts_solver(m, oracle_ipm)
    # convex: no restart improves on the current optimum by more than round-off; statistics over the restarts
    r = bb.resolve_callback(m, s, "cost", m.x0, 5, cur_obj=base["obj"][0] - 1e-6, cur_status=1)
    assert r["used"] == -1 and r["nlp_solves"] == 5 and r["num_failed"] == 0
    assert r["mean"] == pytest.approx(base["obj"][0], rel=1e-7) and r["std_dev"] <= 1e-6 * abs(base["obj"][0])
    # no current solution: the best restart is installed
    r = bb.resolve_callback(m, s, "cost", m.x0, 5)
    assert r["used"] >= 1 and r["status"] == 1 and r["obj"] == pytest.approx(base["obj"][0], rel=1e-7)
    np.testing.assert_allclose(r["x"], base["x"][0], atol=1e-4)


@pytest.mark.gpu
def test_resolve_engine(bnlp_lib, oracle_ipm):
    from bonmin_b200.engine import Engine, Problem
    for m in (I.mytoy(), I.syn_miqcp(40, 20, 8, 8, 3)):
        eng = Engine(Problem.from_model(m))
        base = eng.solve_batch(m.xl, m.xu, m.x0)
        r = bb.resolve_engine(eng, "robustness", m.x0, 7)
        assert r["used"] == 0 and r["nlp_solves"] == 8 and r["status"] == base["status"][0]
        assert r["obj"] == pytest.approx(base["obj"][0], rel=1e-9)
        c = bb.resolve_engine(eng, "cost", m.x0, 16)
        assert c["num_failed"] == 0 and c["status"] == 1
        # gradient-based scaling is computed at the start point (nlp_scaling_max_gradient 100): from a far random start of
        # an unbounded variable (mytoy, radius 1e5) the scaled 1e-8 tolerance is an unscaled ~2e-5 (< constr_viol_tol 1e-4),
        # so restarts agree with the base optimum to ~1e-4 only
        assert c["obj"] == pytest.approx(base["obj"][0], rel=1e-4) and c["mean"] == pytest.approx(base["obj"][0], rel=1e-4)
        assert c["obj"] <= base["obj"][0] + 1e-9
        # same restarts on the oracle: same statistics
        s = _starts_solver(m, oracle_ipm)
        o = bb.resolve_callback(m, s, "cost", m.x0, 16)
        # (from the far starts the accepted accuracy is the loose one described above, and the two arithmetic providers
        # stop on different last bits: they agree to that accuracy, not to 1e-6)
        assert o["mean"] == pytest.approx(c["mean"], rel=1e-4 if m.n < 10 else 1e-6)
        eng.close()
