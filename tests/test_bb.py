"""Host B-BB driver (bonmin_b200/host): strong branching with batched children behind Bonmin's plugin interfaces.
CPU tests run the driver with the CPU oracle as arithmetic provider; the GPU test checks that the CUDA engine behind
the same driver reaches the same MINLP optimum (north_star: same final optimum within 1e-6)."""
import numpy as np
import pytest

from bonmin_b200 import bb, instances as I


def _oracle_solver(model, oracle_ipm):
    dense = oracle_ipm.dense_arrays(model)
    return lambda XL, XU, x0: oracle_ipm.solve_batch(model, XL, XU, x0, threads=4, dense=dense)


def test_bb_mytoy_optimum(oracle_ipm):
    # config C1: B-BB on the toy; true optimum -2.5 at x=1, y=(1,1/2), z=0 (not asserted by the reference)
    m = I.mytoy()
    r = bb.run_callback(m, _oracle_solver(m, oracle_ipm))
    assert r["status"] == 0
    # the incumbent is the x = 1 node: -2.5 moved by the 1e-8 relaxations to -(2.5 - 1e-8 + sqrt(2e-8 - 1e-16)) (test_oracle_golden.py)
    assert r["best_obj"] == pytest.approx(-(2.5 - 1e-8 + (2e-8 - 1e-16) ** 0.5), abs=1e-6)
    assert r["best_x"][2] == pytest.approx(1.0, abs=1e-6) and r["best_x"][3] == pytest.approx(0.0, abs=1e-6)
    assert r["sb_children"] > 0 and r["nodes"] >= 1


def test_bb_strong_branching_settings_agree(oracle_ipm):
    m = I.syn_miqcp(12, 6, 3, 3, 4)
    s = _oracle_solver(m, oracle_ipm)
    a = bb.run_callback(m, s, number_strong_branch=6)
    b = bb.run_callback(m, s, number_strong_branch=2, number_before_trust=1)
    assert a["status"] == b["status"] == 0
    assert a["best_obj"] == pytest.approx(b["best_obj"], rel=1e-6)
    xi = a["best_x"][:6]
    assert np.abs(xi - np.round(xi)).max() <= 1e-6
    # every child of a strong-branched node that is later explored is served from the batch (Appendix C note 2)
    assert a["served_from_sb_cache"] > 0 and a["nlp_solves"] < a["nodes"]


def test_bb_optimum_is_a_valid_lower_bound_of_enumeration(oracle_ipm):
    """Tiny instance: enumerate all integer assignments of 3 integer variables in [0,10] and solve the continuous
    NLP for each with the oracle; the B-BB optimum must equal the best of them."""
    m = I.syn_miqcp(6, 3, 2, 2, 8)
    s = _oracle_solver(m, oracle_ipm)
    r = bb.run_callback(m, s, number_strong_branch=3)
    pts = np.array([(a, b, c) for a in range(11) for b in range(11) for c in range(11)], float)
    XL = np.tile(m.xl, (len(pts), 1))
    XU = np.tile(m.xu, (len(pts), 1))
    XL[:, :3] = pts
    XU[:, :3] = pts
    o = s(XL, XU, m.x0)
    best = o["obj"][o["status"] == 1].min()
    assert r["best_obj"] == pytest.approx(best, rel=1e-6)


def test_bb_parallel_nodes_same_optimum(oracle_ipm):
    """Expanding several best-bound nodes per round (one batched launch for all their children) changes the node
    order, not the optimum."""
    m = I.syn_miqcp(24, 12, 5, 5, 6)
    s = _oracle_solver(m, oracle_ipm)
    a = bb.run_callback(m, s, number_strong_branch=12, number_before_trust=1000000)
    b = bb.run_callback(m, s, number_strong_branch=12, number_before_trust=1000000, parallel_nodes=8)
    assert a["status"] == b["status"] == 0
    assert a["best_obj"] == pytest.approx(b["best_obj"], rel=1e-6)
    assert b["sb_batches"] < a["sb_batches"]


@pytest.mark.gpu
def test_bb_engine_matches_oracle(bnlp_lib, oracle_ipm):
    from bonmin_b200.engine import Engine, Problem
    for m, kw in [(I.mytoy(), {}), (I.syn_miqcp(20, 10, 4, 4, 9), dict(number_strong_branch=10)),
                  (I.syn_miqcp(40, 20, 8, 8, 3), dict(number_strong_branch=16, node_limit=400))]:
        eng = Engine(Problem.from_model(m))
        g = bb.run_engine(eng, **kw)
        c = bb.run_callback(m, _oracle_solver(m, oracle_ipm), **kw)
        assert g["status"] == c["status"]
        assert g["best_obj"] == pytest.approx(c["best_obj"], rel=1e-6)
        assert g["nodes"] == c["nodes"] and g["sb_children"] == c["sb_children"]


def _oracle_tqp(model, oracle_ipm):
    """tqp_solve provider built from the CPU restatement of BranchingTQP + the oracle IPM."""
    from oracle import struct_eval as SE
    ev = SE.Evaluator(model)
    INF = 1e19

    def solve(xs, du, XL, XU):
        qm, f0 = SE.branching_tqp_model(model, ev, xs, du)
        dl = np.where(XL > -INF, XL - xs, XL)
        dh = np.where(XU < INF, XU - xs, XU)
        r = oracle_ipm.solve_batch(qm, dl, dh, np.zeros(model.n), threads=4)
        r["obj"] = np.where(r["status"] == 3, r["obj"], r["obj"] + f0)
        r["x"] = r["x"] + xs
        return r
    return solve


def test_bb_qp_strong_branching_same_optimum(oracle_ipm):
    """variable_selection = qp-strong-branching (a8): the candidates are ranked on the BranchingTQP surrogate; the tree
    differs, the MINLP optimum does not, and QP children are never reused as node relaxations."""
    for m, kw in [(I.mytoy(), {}), (I.syn_miqcp(12, 6, 3, 3, 4), dict(number_strong_branch=6))]:
        s = _oracle_solver(m, oracle_ipm)
        a = bb.run_callback(m, s, **kw)
        q = bb.run_callback(m, s, tqp_solve=_oracle_tqp(m, oracle_ipm), variable_selection="qp-strong-branching", **kw)
        assert a["status"] == q["status"] == 0
        assert q["best_obj"] == pytest.approx(a["best_obj"], rel=1e-6)
        assert q["served_from_sb_cache"] == 0 and q["sb_children"] > 0
        assert q["nlp_solves"] == q["nodes"]


@pytest.mark.gpu
def test_bb_engine_qp_strong_branching(bnlp_lib, oracle_ipm):
    from bonmin_b200.engine import Engine, Problem
    for m, kw in [(I.mytoy(), {}), (I.syn_miqcp(20, 10, 4, 4, 9), dict(number_strong_branch=10))]:
        eng = Engine(Problem.from_model(m))
        g = bb.run_engine(eng, variable_selection="qp-strong-branching", **kw)
        c = bb.run_callback(m, _oracle_solver(m, oracle_ipm), tqp_solve=_oracle_tqp(m, oracle_ipm),
                            variable_selection="qp-strong-branching", **kw)
        n = bb.run_engine(eng, **kw)
        assert g["status"] == c["status"] == 0
        assert g["best_obj"] == pytest.approx(c["best_obj"], rel=1e-6)
        assert g["best_obj"] == pytest.approx(n["best_obj"], rel=1e-6)
        assert g["nodes"] == c["nodes"] and g["sb_children"] == c["sb_children"]


def test_provider_error_is_not_a_fathomed_node():
    """A batch-solve provider that fails must abort the search with an error (status 3 / RuntimeError), never be read as
    'all children infeasible'."""
    from bonmin_b200 import bb, instances as I
    m = I.syn_miqcp(12, 6, 3, 3, 1)
    calls = {"n": 0}

    def broken(XL, XU, x0):
        calls["n"] += 1
        raise RuntimeError("device lost")
    with pytest.raises(RuntimeError):
        bb.run_callback(m, broken, number_strong_branch=4)
    assert calls["n"] >= 1


def test_priorities_restrict_the_candidates(oracle_ipm):
    """With branching priorities only the fractional integers of the best (lowest) level are strong-branching candidates
    (setupList, BonChooseVariable.cpp:330-352): giving one variable the best level makes the first batch its two children."""
    m = I.syn_miqcp(12, 6, 3, 3, 1)
    solve = _oracle_solver(m, oracle_ipm)
    batches = []

    def rec(XL, XU, x0):
        batches.append((XL.copy(), XU.copy()))
        return solve(XL, XU, x0)
    ref = bb.run_callback(m, solve, number_strong_branch=6)
    root = solve(np.atleast_2d(m.xl), np.atleast_2d(m.xu), m.x0)["x"][0]
    frac = [j for j in range(6) if abs(root[j] - round(root[j])) > 1e-6]
    assert len(frac) >= 2
    pri = np.full(12, 1000, np.int32)
    pri[frac[-1]] = 1
    r = bb.run_callback(m, rec, number_strong_branch=6, priorities=pri)
    first = next(b for b in batches if b[0].shape[0] > 1)
    assert first[0].shape[0] == 2                                  # one candidate, two children
    changed = np.flatnonzero((first[0] != m.xl).any(0) | (first[1] != m.xu).any(0))
    assert list(changed) == [frac[-1]]
    assert r["status"] == 0 and r["best_obj"] == pytest.approx(ref["best_obj"], rel=1e-6)   # same optimum, other tree


def test_tnlp_solver_contract_of_the_adapter():
    """B200Solver implements the whole TNLPSolver interface (BonTNLPSolver.hpp:95-205): Initialize from an options stream,
    the IpoptSolver warm-start semantics (BonIpoptSolver.cpp:345-430), error code + UnsolvedError, output levels, clone.
    The checks are C++ (bnlp_host_selftest, bonmin_b200/host/host_cabi.cpp) on a stub provider."""
    assert bb.host_lib().bnlp_host_selftest() == 0


def test_incumbent_from_a_heuristic_seeds_the_cutoff(oracle_ipm):
    """B-Hyb runs its primal heuristics (HeuristicDive*, BonHeuristicDive.cpp:67-302) before the tree: the search then starts
    from cutoff = incumbent - cutoff_decr.  Seeded with the optimum itself the tree proves it (same value, no more nodes than
    the unseeded search) and hands the seeded point back; seeded with a worse feasible point it still finds the optimum."""
    m = I.syn_miqcp(12, 6, 3, 3, 4)
    s = _oracle_solver(m, oracle_ipm)
    a = bb.run_callback(m, s, number_strong_branch=6)
    assert a["status"] == 0
    b = bb.run_callback(m, s, number_strong_branch=6, incumbent=(a["best_obj"], a["best_x"]))
    assert b["status"] == 0 and b["best_obj"] == a["best_obj"] and b["nodes"] <= a["nodes"]
    assert np.array_equal(b["best_x"], a["best_x"])
    c = bb.run_callback(m, s, number_strong_branch=6, incumbent=(a["best_obj"] + 5.0, a["best_x"]))
    assert c["status"] == 0 and c["best_obj"] == pytest.approx(a["best_obj"], rel=1e-9)
