"""Wide element terms (round 2): a non-quadratic expression that couples more than BNLP_MAX_TERM_VARS = 6 variables — the
reference forwards any ASL expression DAG (src/Interfaces/Ampl/BonAmplTMINLP.cpp:629-710).  A model with such a term runs on
the wide build of the term machine (include/bnlp.h: up to 24 variables per term, bnlp_launch_wide.cu).  The instance has a
log-sum-exp over 20 variables, a norm over 12 and a narrow term; it is written as `.nl`, read back by the product's parser and
by the independent pure-Python reader, evaluated on the device against both, and its relaxations are solved against the
oracle."""
import numpy as np
import pytest

from bonmin_b200 import instances as I, nlwrite
from oracle.nl_reader import NlModel
from oracle.struct_eval import Evaluator


def test_wide_model_loads_and_limits(bnlp_lib):
    from bonmin_b200.engine import BnlpError, Problem
    m = I.wide_terms_minlp()
    ev = Evaluator(m)
    txt = nlwrite.write_nl(m)
    for p in (Problem.from_nl(txt), Problem.from_model(m)):
        assert (p.n, p.m, p.nnz_jac, p.nnz_hess) == (30, 6, len(ev.jac), len(ev.hes))
        info = p.info()
        assert list(zip(info["jac_irow"], info["jac_jcol"])) == ev.jac
    # the independent reader agrees with the structured evaluator on the written file
    x = np.random.default_rng(0).uniform(-1, 2, m.n)
    np.testing.assert_allclose(NlModel(txt).values(x), ev.dense(x)[0], rtol=1e-14)
    # 25 variables in one expression: beyond the wide machine too — refused, with the limit in the message
    big = I.wide_terms_minlp()
    prog = [(I.OP_VAR, 0, 0.0), (I.OP_EXP, 0, 0.0)]
    for i in range(1, 25):
        prog += [(I.OP_VAR, i, 0.0), (I.OP_EXP, 0, 0.0), (I.OP_ADD, 0, 0.0)]
    prog.append((I.OP_LOG, 0, 0.0))
    big["terms"] = [(1, list(range(25)), prog)]
    with pytest.raises(BnlpError, match="24"):
        Problem.from_model(big)
    # a wide term next to a term that needs a deeper stack than the wide machine has: refused as well
    deep = I.wide_terms_minlp()
    dprog = [(I.OP_VAR, 0, 0.0)] * 10 + [(I.OP_ADD, 0, 0.0)] * 9
    deep["terms"] = list(deep["terms"]) + [(3, [20], dprog)]
    with pytest.raises(BnlpError, match="stack"):
        Problem.from_model(deep)
    # ... while the same deep term alone (narrow model, stack of 12) is fine
    narrow = I.wide_terms_minlp()
    narrow["terms"] = [(3, [20], dprog)]
    assert Problem.from_model(narrow).n == 30


@pytest.mark.gpu
def test_wide_terms_eval_matches_both_readers(bnlp_lib):
    from bonmin_b200.engine import Engine, Problem
    m = I.wide_terms_minlp()
    ev = Evaluator(m)
    txt = nlwrite.write_nl(m)
    nm = NlModel(txt)
    rng = np.random.default_rng(3)
    B = 41
    X = np.clip(m.x0 + rng.normal(0, 0.5, (B, m.n)), m.xl, m.xu)
    LAM = rng.normal(0, 1, (B, m.m))
    for prob in (Problem.from_nl(txt), Problem.from_model(m)):
        eng = Engine(prob)
        r = eng.eval_batch(X, LAM, sigma=0.7)
        for b in range(B):
            o = ev.sparse(X[b], LAM[b], 0.7)
            for k in ("f", "g", "gradf", "jac", "hess"):
                np.testing.assert_allclose(r[k][b], o[k], rtol=1e-12, atol=1e-12 * (1 + np.abs(o[k]).max(initial=0)))
            np.testing.assert_allclose(np.concatenate([[r["f"][b]], r["g"][b]]), nm.values(X[b]), rtol=1e-12, atol=1e-12)
        eng.close()


@pytest.mark.gpu
def test_wide_terms_relaxations_match_the_oracle(bnlp_lib, oracle_ipm):
    from bonmin_b200.engine import Engine, Problem
    m = I.wide_terms_minlp()
    eng = Engine(Problem.from_nl(nlwrite.write_nl(m)))
    root = eng.solve_batch(m.xl, m.xu, m.x0)
    oroot = oracle_ipm.solve_batch(m, m.xl, m.xu, m.x0)
    assert root["status"][0] == oroot["status"][0] == 1
    assert root["obj"][0] == pytest.approx(oroot["obj"][0], rel=1e-6)
    xs = root["x"][0]
    XL, XU = I.c3_children(m, xs, 96)
    r = eng.solve_batch(XL, XU, xs)
    o = oracle_ipm.solve_batch(m, XL, XU, xs, threads=8)
    assert (r["status"] == o["status"]).all(), np.flatnonzero(r["status"] != o["status"])
    ok = o["status"] == 1
    assert ok.sum() >= 40 and (o["status"] == 3).sum() >= 5
    rel = np.abs(r["obj"] - o["obj"]) / np.maximum(1.0, np.abs(o["obj"]))
    assert rel[ok].max() <= 1e-6
    # the rows hold at the returned points, evaluated independently
    ev = Evaluator(m)
    for b in np.flatnonzero(ok)[:10]:
        g = ev.dense(r["x"][b])[0][1:]
        assert (g <= np.asarray(m.gu) + 1e-6).all()
    # linearisation of the wide rows at the root (getOuterApproximation): a valid cut at random feasible-ish points
    cuts = eng.oa_batch(xs[None, :], root["duals"], m.xl, m.xu, get_obj=0)
    assert cuts["rows_per_point"] >= 3
    eng.close()


def _inverse_trig_model():
    """atan / asin / acos (AMPL o49 / o51 / o53): evaluation only (the rows are not convex)."""
    m = I.wide_terms_minlp()
    m["terms"] = [(1, [0, 1], [(I.OP_VAR, 0, 0.0), (I.OP_VAR, 1, 0.0), (I.OP_MUL, 0, 0.0), (16, 0, 0.0)]),                    # atan(x0 x1)
                  (2, [2], [(I.OP_CONST, 0, 0.3), (I.OP_VAR, 0, 0.0), (I.OP_MUL, 0, 0.0), (17, 0, 0.0)]),                     # asin(0.3 x2)
                  (3, [3, 4], [(I.OP_CONST, 0, 0.1), (I.OP_VAR, 0, 0.0), (I.OP_MUL, 0, 0.0), (I.OP_CONST, 0, 0.2), (I.OP_VAR, 1, 0.0),
                               (I.OP_MUL, 0, 0.0), (I.OP_ADD, 0, 0.0), (18, 0, 0.0)])]                                        # acos(0.1 x3 + 0.2 x4)
    m["name"] = "inverse trigonometric terms"
    m["vtype"] = np.zeros(m.n, np.int32)         # (all continuous: only x_0..x_4 are nonlinear now, AMPL's order must still hold)
    return m


def test_inverse_trig_opcodes_are_read(bnlp_lib):
    from bonmin_b200.engine import Problem
    m = _inverse_trig_model()
    txt = nlwrite.write_nl(m)
    assert "o49" in txt and "o51" in txt and "o53" in txt
    p = Problem.from_nl(txt)
    ev = Evaluator(m)
    assert (p.nnz_jac, p.nnz_hess) == (len(ev.jac), len(ev.hes))
    x = np.random.default_rng(1).uniform(-1, 1, m.n)
    np.testing.assert_allclose(NlModel(txt).values(x), ev.dense(x)[0], rtol=1e-14)
    # derivatives of the evaluator against central differences
    v0, g0, _ = ev.dense(x)
    for j in range(5):
        e = np.zeros(m.n)
        e[j] = 1e-6
        fd = (ev.dense(x + e)[0] - ev.dense(x - e)[0]) / 2e-6
        np.testing.assert_allclose(g0[:, j], fd, atol=1e-7)


@pytest.mark.gpu
def test_inverse_trig_eval_on_device(bnlp_lib):
    from bonmin_b200.engine import Engine, Problem
    m = _inverse_trig_model()
    ev = Evaluator(m)
    eng = Engine(Problem.from_nl(nlwrite.write_nl(m)))
    rng = np.random.default_rng(2)
    X = rng.uniform(-1, 1, (9, m.n))
    LAM = rng.normal(0, 1, (9, m.m))
    r = eng.eval_batch(X, LAM, sigma=1.0)
    for b in range(9):
        o = ev.sparse(X[b], LAM[b], 1.0)
        for k in ("f", "g", "gradf", "jac", "hess"):
            np.testing.assert_allclose(r[k][b], o[k], rtol=1e-12, atol=1e-12 * (1 + np.abs(o[k]).max(initial=0)))
    eng.close()


def test_oracle_on_wide_terms_against_scipy(oracle_ipm):
    """The oracle's wide-term path (24-variable second-order forward mode in ipm_oracle.cpp) pinned independently: root and
    children of the wide model against scipy SLSQP on the numpy evaluator (objective 1e-6 relative, same feasibility verdict)."""
    from scipy.optimize import minimize
    m = I.wide_terms_minlp()
    ev = Evaluator(m)
    root = oracle_ipm.solve_batch(m, m.xl, m.xu, m.x0)
    assert root["status"][0] == 1
    xs = root["x"][0]
    XL, XU = I.c3_children(m, xs, 12)
    XL = np.vstack([m.xl, XL])
    XU = np.vstack([m.xu, XU])
    r = oracle_ipm.solve_batch(m, XL, XU, xs, threads=4)
    gu = np.asarray(m.gu)

    def fg(x):
        v, g, _ = ev.dense(x)
        return v, g

    n_feas = 0
    for b in range(len(XL)):
        fixed = XL[b] >= XU[b]
        if (XL[b] > XU[b]).any():
            assert r["status"][b] == 3
            continue
        cons = [dict(type="ineq", fun=lambda x: gu - fg(x)[0][1:], jac=lambda x: -fg(x)[1][1:])]
        res = minimize(lambda x: fg(x)[0][0], np.clip(xs, XL[b], XU[b]), jac=lambda x: fg(x)[1][0], method="SLSQP",
                       bounds=list(zip(XL[b], np.where(fixed, XL[b], XU[b]))), constraints=cons, options=dict(maxiter=500, ftol=1e-12))
        viol = max(0.0, float((fg(res.x)[0][1:] - gu).max()))
        if r["status"][b] == 1:
            assert viol < 1e-6 and abs(res.fun - r["obj"][b]) <= 1e-6 * max(1.0, abs(res.fun)), (b, res.fun, r["obj"][b], viol)
            n_feas += 1
        else:
            assert r["status"][b] == 3 and (viol > 1e-7 or not res.success), (b, viol, res.message)
    assert n_feas >= 6
