"""Host B-OA driver (bonmin_b200/oa.py): OACutGenerator2::performOa around the batched engine, master
MILP by HiGHS.  Two different algorithms (B-OA here, B-BB in test_bb.py) must reach the same MINLP optimum."""
import numpy as np
import pytest

from bonmin_b200 import bb, oa, instances as I
from oracle import struct_eval as SE

scipy_milp = pytest.importorskip("scipy.optimize", reason="HiGHS master needs scipy").milp


class OracleProvider:
    """Arithmetic provider = the CPU oracle (checker only)."""

    def __init__(self, model, oracle_ipm):
        self.model, self.ipm = model, oracle_ipm
        self.dense = oracle_ipm.dense_arrays(model)
        self.ev = SE.Evaluator(model)
        self.n, self.m = model.n, model.m
        self.xl, self.xu, self.vtype, self.x0 = model.xl, model.xu, model.vtype, model.x0
        self.duals = np.zeros(2 * self.n + self.m)

    def solve(self, XL, XU, x0):
        r = self.ipm.solve_batch(self.model, XL, XU, x0, threads=4, dense=self.dense)
        self.duals = np.array(r["duals"][-1])       # getRowPrice() of the last solve picks the side of ranged rows
        return r

    def extract(self, x, duals):
        return SE.extract_linear_relaxation(self.ev, x, duals, get_obj=True)

    def violation(self, x, obj):
        return float(SE.constraints_violation(self.ev, x, obj)[0])

    def oa(self, X, X2=None, theta=0.0):
        out = []
        for x in np.atleast_2d(X):
            out.extend(SE.outer_approximation(self.ev, x, self.duals, self.xl, self.xu, get_obj=1))
        return out


def _bb_opt(m, oracle_ipm, **kw):
    dense = oracle_ipm.dense_arrays(m)
    return bb.run_callback(m, lambda XL, XU, x0: oracle_ipm.solve_batch(m, XL, XU, x0, threads=4, dense=dense), **kw)


@pytest.mark.parametrize("mk", [lambda: I.mytoy(), lambda: I.syn_miqcp(12, 6, 3, 3, 4),
                                lambda: I.syn_cvx_minlp(16, 6, 3, 5, 2)])
def test_oa_reaches_the_bb_optimum(oracle_ipm, mk):
    m = mk()
    r = oa.run_oa(OracleProvider(m, oracle_ipm))
    b = _bb_opt(m, oracle_ipm, number_strong_branch=6)
    assert r["status"] == b["status"] == 0
    # both stop on tolerances (B-BB cutoff_decr 1e-5, B-OA cutoff increment 1e-5 relative, 1e-8 bound relaxation)
    assert r["best_obj"] == pytest.approx(b["best_obj"], rel=2e-5, abs=2e-5)
    xi = r["best_x"][np.asarray(m.vtype) != 0]
    assert np.abs(xi - np.round(xi)).max() <= 1e-6
    assert r["lower_bound"] <= r["best_obj"] + 1e-6 * max(1.0, abs(r["best_obj"])) or r["lower_bound"] >= 1e49
    assert r["cuts"] > 0 and r["oa_points"] >= r["passes"]


@pytest.mark.gpu
def test_oa_engine_matches_oracle(bnlp_lib, oracle_ipm):
    from bonmin_b200.engine import Engine, Problem
    for m in (I.mytoy(), I.syn_miqcp(12, 6, 3, 3, 4), I.syn_cvx_minlp(16, 6, 3, 5, 2)):
        eng = Engine(Problem.from_model(m))
        g = oa.run_oa(oa.EngineProvider(eng))
        c = oa.run_oa(OracleProvider(m, oracle_ipm))
        assert g["status"] == c["status"] == 0
        assert g["best_obj"] == pytest.approx(c["best_obj"], rel=1e-6, abs=1e-8)
        # the master (HiGHS) may pick another of several optimal integer points when cut coefficients differ in the last
        # bit, so pass counts can differ by a few; the optimum cannot
        assert abs(g["passes"] - c["passes"]) <= max(3, c["passes"] // 5)
        eng.close()


def test_ecp_rounds_tighten_the_lp_bound(oracle_ipm):
    """EcpCuts::generateCuts: cuts at successive LP optima raise the LP bound monotonically towards the NLP relaxation
    value and shrink the violation of the nonlinear rows; the LP bound never exceeds the relaxation optimum (convexity)."""
    m = I.syn_miqcp(12, 6, 3, 3, 4)
    P = OracleProvider(m, oracle_ipm)
    root = P.solve(m.xl[None, :], m.xu[None, :], m.x0)
    # start from the OA at a point far from the optimum so that the first LP is loose
    lp = P.extract(np.clip(m.x0, m.xl, m.xu), None)
    prev, prev_viol = -1e300, None
    first = oa.ecp_rounds(P, lp, rounds=1)["obj"]
    for rounds in (1, 3, 8, 30):
        r = oa.ecp_rounds(P, lp, rounds=rounds)
        assert r["obj"] >= prev - 1e-9 and r["obj"] <= root["obj"][0] + 1e-6 * abs(root["obj"][0])
        prev = r["obj"]
        if rounds == 1:
            # numRounds 1: the cuts are generated and handed back, but not installed and the LP is not re-solved (:86)
            assert r["installed"] == 0 and r["cuts"] == len(r["cs"]) > 0
        else:
            assert r["cuts"] > 0 and r["violation"] <= (prev_viol if prev_viol is not None else 1e300) + 1e-9
        prev_viol = r["violation"]
    # Kelley-type cutting planes converge slowly; 30 rounds must have closed most of the gap of the first LP
    assert root["obj"][0] - prev <= 0.25 * (root["obj"][0] - first)


@pytest.mark.gpu
def test_ecp_engine_matches_oracle(bnlp_lib, oracle_ipm):
    from bonmin_b200.engine import Engine, Problem
    m = I.syn_miqcp(12, 6, 3, 3, 4)
    eng = Engine(Problem.from_model(m))
    G, C = oa.EngineProvider(eng), OracleProvider(m, oracle_ipm)
    x0 = np.clip(m.x0, m.xl, m.xu)
    g = oa.ecp_rounds(G, G.extract(x0, None), rounds=8)
    c = oa.ecp_rounds(C, C.extract(x0, None), rounds=8)
    assert g["rounds"] == c["rounds"] and g["cuts"] == c["cuts"]
    assert g["obj"] == pytest.approx(c["obj"], rel=1e-8) and g["violation"] == pytest.approx(c["violation"], rel=1e-6, abs=1e-9)
    eng.close()


@pytest.mark.parametrize("mk", [lambda: I.mytoy(), lambda: I.syn_miqcp(12, 6, 3, 3, 4),
                                lambda: I.syn_cvx_minlp(16, 6, 3, 5, 2)])
@pytest.mark.parametrize("hyb", [0.0, 1.0])
def test_single_tree_qg_and_hyb_reach_the_bb_optimum(oracle_ipm, mk, hyb):
    """OaFeasibilityChecker (B-QG) and OaNlpOptim (B-Hyb, nlp_solves_per_level > 0) inside a single LP tree: the same MINLP
    optimum as B-BB."""
    m = mk()
    r = oa.run_qg(OracleProvider(m, oracle_ipm), nlp_solves_per_level=hyb)
    b = _bb_opt(m, oracle_ipm, number_strong_branch=6)
    assert r["status"] == b["status"] == 0
    assert r["best_obj"] == pytest.approx(b["best_obj"], rel=2e-5, abs=2e-5)
    xi = r["best_x"][np.asarray(m.vtype) != 0]
    assert np.abs(xi - np.round(xi)).max() <= 1e-6
    assert r["cuts"] > 0 and r["nlp_solves"] >= 2


def test_nlp_optim_depth_rule():
    rng = np.random.default_rng(0)
    assert not oa.nlp_optim_wanted(11, 1.0, 10, rng)                 # below max depth: never
    assert oa.nlp_optim_wanted(30, 1e10, 100, rng)                    # "always"
    assert oa.nlp_optim_wanted(0, 1.0, 10, rng)                       # 2^0 * 1 > rand in [0, 1)
    hits = sum(oa.nlp_optim_wanted(3, 1.0, 10, rng) for _ in range(4000))
    assert 400 < hits < 600                                           # probability 1/8


@pytest.mark.gpu
def test_single_tree_engine_matches_oracle(bnlp_lib, oracle_ipm):
    from bonmin_b200.engine import Engine, Problem
    for m in (I.mytoy(), I.syn_miqcp(12, 6, 3, 3, 4), I.syn_cvx_minlp(16, 6, 3, 5, 2)):
        eng = Engine(Problem.from_model(m))
        for hyb in (0.0, 1.0):
            g = oa.run_qg(oa.EngineProvider(eng), nlp_solves_per_level=hyb)
            c = oa.run_qg(OracleProvider(m, oracle_ipm), nlp_solves_per_level=hyb)
            assert g["status"] == c["status"] == 0
            assert g["best_obj"] == pytest.approx(c["best_obj"], rel=1e-6, abs=1e-8)
        eng.close()


def test_ranged_nonlinear_row_is_cut_on_its_binding_side_only(oracle_ipm):
    """A ranged nonlinear row (g_l <= g(x) <= g_u, g convex): only the side that binds at the NLP optimum is a valid outer
    approximation; the reference picks it by the sign of the row's multiplier (BonOsiTMINLPInterface.cpp:1968-1974, duals =
    getRowPrice() of the last NLP solve).  Without multipliers both sides are linearised, and the lower-side cut
    `grad g(xbar).(x - xbar) + g(xbar) >= g_l` is not an outer approximation of a convex g."""
    m = I.syn_miqcp(10, 5, 3, 2, 11)
    gl = np.array(m.gl, float)
    gl[:3] = np.asarray(m.gu)[:3] - 400.0            # the three quadratic rows become ranged; the upper side is the near one
    m["gl"] = gl
    P = OracleProvider(m, oracle_ipm)
    r = oa.run_oa(P)
    b = _bb_opt(m, oracle_ipm, number_strong_branch=5)
    assert r["status"] == b["status"] == 0
    assert r["best_obj"] == pytest.approx(b["best_obj"], rel=2e-5, abs=2e-5)
    root = P.solve(np.asarray(m.xl)[None, :], np.asarray(m.xu)[None, :], m.x0)
    one_sided = [c for c in P.oa(root["x"][0][None, :]) if c[0] < 3]
    assert len(one_sided) == 3 and all(lb <= -1e100 and ub < 1e100 for _, _, _, lb, ub in one_sided)
    two_sided = [c for c in SE.outer_approximation(P.ev, root["x"][0], None, m.xl, m.xu, get_obj=1) if c[0] < 3]
    assert all(lb > -1e100 and ub < 1e100 for _, _, _, lb, ub in two_sided)      # what a provider without duals would add


_KEYS = ("status", "best_obj", "lower_bound", "passes", "nlp_solves", "cuts", "oa_points")


@pytest.mark.parametrize("mk", [lambda: I.mytoy(), lambda: I.syn_miqcp(12, 6, 3, 3, 4),
                                lambda: I.syn_cvx_minlp(16, 6, 3, 5, 2)])
def test_cpp_oa_loop_reproduces_the_python_loop(oracle_ipm, mk):
    # bonmin_b200/host/B200Oa.cpp (bnlp_oa_run_cb) against oa.run_oa on the same provider and the same master: the two
    # statements of OACutGenerator2::performOa must take the same passes, install the same cuts and end at the same point
    m = mk()
    p = oa.run_oa(OracleProvider(m, oracle_ipm))
    c = oa.run_oa_callback(OracleProvider(m, oracle_ipm))
    assert {k: c[k] for k in _KEYS} == {k: p[k] for k in _KEYS}
    assert np.array_equal(c["best_x"], p["best_x"])
    # without the master point (the reference's plain performOa): still the same optimum
    c1 = oa.run_oa_callback(OracleProvider(m, oracle_ipm), add_master_point=False)
    p1 = oa.run_oa(OracleProvider(m, oracle_ipm), add_master_point=False)
    assert {k: c1[k] for k in _KEYS} == {k: p1[k] for k in _KEYS}
    assert c1["best_obj"] == pytest.approx(p["best_obj"], rel=2e-5, abs=2e-5)


def test_cpp_oa_loop_limits_and_provider_failure(oracle_ipm):
    m = I.syn_miqcp(12, 6, 3, 3, 4)
    r = oa.run_oa_callback(OracleProvider(m, oracle_ipm), max_passes=3)
    assert r["status"] == 2 and r["passes"] == 3 and r["lower_bound"] <= r["best_obj"]

    class Broken(OracleProvider):
        def oa(self, X, X2=None, theta=0.0):
            raise RuntimeError("provider down")
    with pytest.raises(RuntimeError, match="provider down"):
        oa.run_oa_callback(Broken(m, oracle_ipm))


@pytest.mark.gpu
def test_cpp_oa_loop_on_the_engine(bnlp_lib, oracle_ipm):
    # product path: bnlp_oa_run_engine (C++ loop -> bnlp_solve_batch / bnlp_extract_linear_relaxation / bnlp_oa_batch,
    # HiGHS master behind the callback) against the Python loop on the same engine and against the CPU oracle
    from bonmin_b200.engine import Engine, Problem
    for m in (I.mytoy(), I.syn_miqcp(12, 6, 3, 3, 4), I.syn_cvx_minlp(16, 6, 3, 5, 2), I.syn_cvx_minlp(60, 24, 10, 10, 5)):
        eng = Engine(Problem.from_model(m))
        l0 = eng.launches()
        g = oa.run_oa_engine(eng)
        assert eng.launches() > l0
        p = oa.run_oa(oa.EngineProvider(eng))
        assert {k: g[k] for k in _KEYS} == {k: p[k] for k in _KEYS}
        assert np.array_equal(g["best_x"], p["best_x"])
        c = oa.run_oa(OracleProvider(m, oracle_ipm))
        assert g["status"] == c["status"] == 0
        # both loops stop on the relative cutoff increment (1e-5): two arithmetic providers may end on different incumbents
        # inside that band
        assert g["best_obj"] == pytest.approx(c["best_obj"], rel=2e-5, abs=2e-5)
        eng.close()


_QG_KEYS = ("status", "best_obj", "nodes", "nlp_solves", "cuts")


@pytest.mark.parametrize("mk", [lambda: I.mytoy(), lambda: I.syn_miqcp(12, 6, 3, 3, 4),
                                lambda: I.syn_cvx_minlp(16, 6, 3, 5, 2)])
def test_cpp_feasibility_checker_inside_the_single_tree(oracle_ipm, mk):
    # OaFeasibilityChecker::performOa in C++ (B200Oa.cpp: oa_feas_check, bnlp_oa_feas_check_cb) in the place of the Python
    # feasibility_check inside B-QG: the tree must be the same tree (nodes, NLP solves, cuts, optimum)
    m = mk()
    p = oa.run_qg(OracleProvider(m, oracle_ipm))
    P = OracleProvider(m, oracle_ipm)
    c = oa.run_qg(P, feas_check=lambda _, lp, lb, ub, x, cutoff, it, ci: oa.feasibility_check_native(P, lp, lb, ub, x, cutoff, it, ci))
    assert {k: c[k] for k in _QG_KEYS} == {k: p[k] for k in _QG_KEYS}
    assert np.array_equal(c["best_x"], p["best_x"])


@pytest.mark.gpu
def test_cpp_feasibility_checker_on_the_engine(bnlp_lib, oracle_ipm):
    from bonmin_b200.engine import Engine, Problem
    for m in (I.mytoy(), I.syn_miqcp(12, 6, 3, 3, 4), I.syn_cvx_minlp(16, 6, 3, 5, 2)):
        eng = Engine(Problem.from_model(m))
        p = oa.run_qg(oa.EngineProvider(eng))
        g = oa.run_qg(oa.EngineProvider(eng),
                      feas_check=lambda _, lp, lb, ub, x, cutoff, it, ci: oa.feasibility_check_native(eng, lp, lb, ub, x, cutoff, it, ci))
        assert {k: g[k] for k in _QG_KEYS} == {k: p[k] for k in _QG_KEYS}
        c = oa.run_qg(OracleProvider(m, oracle_ipm))
        assert g["status"] == c["status"] == 0
        assert g["best_obj"] == pytest.approx(c["best_obj"], rel=1e-6, abs=1e-8)
        eng.close()


def _native(target):
    return dict(feas_check=lambda _, lp, lb, ub, x, cutoff, it, ci: oa.feasibility_check_native(target, lp, lb, ub, x, cutoff, it, ci),
                nlp_optim=lambda _, lb, ub, x: oa.nlp_optim_cuts_native(target, lb, ub, x))


@pytest.mark.parametrize("mk", [lambda: I.syn_miqcp(12, 6, 3, 3, 4), lambda: I.syn_cvx_minlp(16, 6, 3, 5, 2)])
def test_cpp_generators_inside_b_hyb(oracle_ipm, mk):
    # OaNlpOptim::generateCuts and OaFeasibilityChecker::performOa in C++ (B200Oa.cpp) in the places of the Python ones inside
    # the single-tree B-Hyb: same tree, same cuts, same optimum
    m = mk()
    p = oa.run_qg(OracleProvider(m, oracle_ipm), nlp_solves_per_level=1.0, seed=3)
    P = OracleProvider(m, oracle_ipm)
    c = oa.run_qg(P, nlp_solves_per_level=1.0, seed=3, **_native(P))
    assert {k: c[k] for k in _QG_KEYS} == {k: p[k] for k in _QG_KEYS}
    assert np.array_equal(c["best_x"], p["best_x"])


def test_cpp_ecp_rounds_reproduce_the_python_rounds(oracle_ipm):
    m = I.syn_miqcp(12, 6, 3, 3, 4)
    P = OracleProvider(m, oracle_ipm)
    P.solve(m.xl[None, :], m.xu[None, :], m.x0)
    lp = P.extract(np.clip(m.x0, m.xl, m.xu), None)
    for rounds in (1, 3, 8):
        p = oa.ecp_rounds(P, lp, rounds=rounds)
        c = oa.ecp_rounds_native(P, lp, rounds=rounds)
        assert (c["rounds"], c["cuts"], c["installed"]) == (p["rounds"], p["cuts"], p["installed"])
        assert c["obj"] == p["obj"] and c["violation"] == p["violation"]
        assert c["cs"] == [(int(a), list(b), list(v), float(l), float(u)) for a, b, v, l, u in p["cs"]]


@pytest.mark.gpu
def test_cpp_generators_on_the_engine(bnlp_lib, oracle_ipm):
    from bonmin_b200.engine import Engine, Problem
    for m in (I.syn_miqcp(12, 6, 3, 3, 4), I.syn_cvx_minlp(16, 6, 3, 5, 2)):
        eng = Engine(Problem.from_model(m))
        p = oa.run_qg(oa.EngineProvider(eng), nlp_solves_per_level=1.0, seed=3)
        g = oa.run_qg(oa.EngineProvider(eng), nlp_solves_per_level=1.0, seed=3, **_native(eng))
        assert {k: g[k] for k in _QG_KEYS} == {k: p[k] for k in _QG_KEYS}
        c = oa.run_qg(OracleProvider(m, oracle_ipm), nlp_solves_per_level=1.0, seed=3)
        assert g["status"] == c["status"] == 0 and g["best_obj"] == pytest.approx(c["best_obj"], rel=2e-5, abs=2e-5)
        # ECP rounds: the C++ generator on the engine against the Python rounds on the engine
        P = oa.EngineProvider(eng)
        root = P.solve(m.xl[None, :], m.xu[None, :], m.x0)
        lp = P.extract(np.clip(m.x0, m.xl, m.xu), None)
        for rounds in (1, 4):
            pe = oa.ecp_rounds(P, lp, rounds=rounds)
            ce = oa.ecp_rounds_native(eng, lp, rounds=rounds, duals=root["duals"][0])
            assert (ce["rounds"], ce["cuts"], ce["installed"]) == (pe["rounds"], pe["cuts"], pe["installed"])
            assert ce["obj"] == pe["obj"] and ce["violation"] == pe["violation"]
        eng.close()
