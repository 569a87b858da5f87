"""GPU suite (-m gpu): batched tape evaluation and OA cut kernels through the C ABI vs the numpy oracle.
Structure (indices, row counts) must be identical; values within 1e-12 relative."""
import numpy as np
import pytest

from bonmin_b200 import instances as I
from oracle.struct_eval import Evaluator, outer_approximation

pytestmark = pytest.mark.gpu


def _engine(model):
    from bonmin_b200.engine import Engine, Problem
    return Engine(Problem.from_model(model))


MODELS = [lambda: I.mytoy(), lambda: I.syn_miqcp(40, 20, 8, 8, 3), lambda: I.syn_cvx_minlp(30, 10, 5, 8, 3),
          lambda: I.syn_cvx_minlp(120, 40, 20, 25, 5)]


@pytest.mark.parametrize("mk", MODELS)
def test_eval_batch_parity(bnlp_lib, mk):
    m = mk()
    ev = Evaluator(m)
    eng = _engine(m)
    info = eng.problem.info()
    assert list(zip(info["jac_irow"], info["jac_jcol"])) == ev.jac
    assert list(zip(info["hes_irow"], info["hes_jcol"])) == [(int(a), int(b)) for a, b in ev.hes]
    rng = np.random.default_rng(0)
    B = 37
    X = np.clip(m.x0 + rng.normal(0, 0.3, (B, m.n)), np.maximum(m.xl, -20), np.minimum(m.xu, 20))
    LAM = rng.normal(0, 1, (B, m.m))
    r = eng.eval_batch(X, LAM, sigma=0.7)
    for b in range(B):
        o = ev.sparse(X[b], LAM[b], 0.7)
        for k in ("f", "g", "gradf", "jac", "hess"):
            np.testing.assert_allclose(r[k][b], o[k], rtol=1e-12, atol=1e-12 * (1 + np.abs(o[k]).max(initial=0)))
    # LAM = None means zero multipliers; outputs may be skipped
    r2 = eng.eval_batch(X, None, sigma=1.0, want=("f", "hess"))
    o = ev.sparse(X[0], np.zeros(m.m), 1.0)
    np.testing.assert_allclose(r2["hess"][0], o["hess"], rtol=1e-12, atol=1e-13)
    assert r2["g"] is None and r2["jac"] is None


def _cuts_from_csr(out, b):
    rpp = out["rows_per_point"]
    cuts = []
    for s in range(rpp):
        o = b * rpp + s
        if out["src_row"][o] < 0:
            continue
        a, e = out["row_start"][o], out["row_start"][o + 1]
        cuts.append((int(out["src_row"][o]), out["col"][a:e].tolist(), out["val"][a:e], out["lb"][o], out["ub"][o]))
    return cuts


@pytest.mark.parametrize("mk", MODELS)
@pytest.mark.parametrize("get_obj", [0, 2])
def test_oa_batch_parity(bnlp_lib, mk, get_obj):
    m = mk()
    ev = Evaluator(m)
    eng = _engine(m)
    rng = np.random.default_rng(1)
    B = 19
    X = np.clip(m.x0 + rng.normal(0, 0.2, (B, m.n)), np.maximum(m.xl, -20), np.minimum(m.xu, 20))
    X[:, ::7] = np.where(np.isfinite(m.xl[::7]) & (m.xl[::7] > -1e18), m.xl[::7], X[:, ::7])  # some columns at a bound
    D = rng.normal(0, 1, (B, 2 * m.n + m.m))
    D[:, 2 * m.n::3] = 0.0
    out = eng.oa_batch(X, DUALS=D, get_obj=get_obj)
    for b in range(B):
        ref = outer_approximation(ev, X[b], duals=D[b], get_obj=get_obj)
        got = _cuts_from_csr(out, b)
        assert [c[0] for c in got] == [c[0] for c in ref]
        for g, r in zip(got, ref):
            assert g[1] == r[1]                                   # identical structure after cleanNnz
            np.testing.assert_allclose(g[2], r[2], rtol=1e-12, atol=1e-14)
            assert g[3] == pytest.approx(r[3], rel=1e-12, abs=1e-12)
            assert g[4] == pytest.approx(r[4], rel=1e-12, abs=1e-12)


def test_oa_mytoy_golden_and_tiny_coefficients(bnlp_lib):
    sq5 = 5 ** 0.5
    m = I.mytoy()
    eng = _engine(m)
    x = np.array([[.5 + 1 / sq5, .5 + 1 / (2 * sq5), .5 + 1 / sq5, 0.]])
    d = np.zeros((1, 11))
    d[0, 8], d[0, 9] = sq5, 1.0
    out = eng.oa_batch(x, DUALS=d)
    (row, cols, vals, lb, ub), = _cuts_from_csr(out, 0)
    assert row == 0 and cols == [0, 1]
    np.testing.assert_allclose(vals, [2 / sq5, 1 / sq5], rtol=1e-12)       # InterfaceTest.cpp:255-260
    rhs = .5 + 3 / (2 * sq5)
    assert ub == pytest.approx(rhs * (1 + 1e-8), rel=1e-12)                 # :230-231 (+ oa_rhs_relax)
    # a point where 2*y1-1 is "tiny": the coefficient is absorbed into the rhs through the column bound (cleanNnz)
    ev = Evaluator(m)
    x = np.array([[0.5 + 2e-9, 0.9, 0.4, 1.0]])
    out = eng.oa_batch(x, DUALS=d)
    ref = outer_approximation(ev, x[0], duals=d[0])
    got = _cuts_from_csr(out, 0)
    assert got[0][1] == ref[0][1] == [1]
    assert got[0][4] == pytest.approx(ref[0][4], rel=1e-12)


def test_oa_violation_filter_and_no_duals(bnlp_lib):
    from bonmin_b200.engine import OaTols
    m = I.syn_cvx_minlp(30, 10, 5, 8, 3)
    ev = Evaluator(m)
    eng = _engine(m)
    X = np.tile(m.x0 + 0.05, (3, 1))
    X2 = np.tile(np.append(m.x0, 0.0), (3, 1))
    out = eng.oa_batch(X, X2=X2, theta=1e9, get_obj=2)
    ref = outer_approximation(ev, X[0], x2=X2[0], theta=1e9, get_obj=2)
    assert [c[0] for c in _cuts_from_csr(out, 0)] == [c[0] for c in ref]   # constraint cuts kept, objective cut filtered
    out = eng.oa_batch(X, X2=X2, theta=1e9, get_obj=2, tols=OaTols(1e-8, 1e-17, 1e-8, 1e19, 1))
    assert _cuts_from_csr(out, 0) == []


@pytest.mark.parametrize("mk", MODELS[1:])
def test_violation_and_benders_parity(bnlp_lib, mk):
    from oracle.struct_eval import benders_cut, constraints_violation
    m = mk()
    ev = Evaluator(m)
    eng = _engine(m)
    rng = np.random.default_rng(3)
    B = 11
    X = np.clip(m.x0 + rng.normal(0, 0.4, (B, m.n)), np.maximum(m.xl, -20), np.minimum(m.xu, 20))
    V, F = eng.violation_batch(X)
    OBJ = F - rng.uniform(0, 2, B)
    V2, _ = eng.violation_batch(X, OBJ)
    D = rng.normal(0, 1, (B, 2 * m.n + m.m))
    D[:, 2 * m.n::2] *= 1e-8            # some multipliers below the 1e-6 threshold
    cuts = eng.benders_batch(X, D)
    for b in range(B):
        v, f = constraints_violation(ev, X[b])
        assert V[b] == pytest.approx(v, rel=1e-12, abs=1e-12) and F[b] == pytest.approx(f, rel=1e-12)
        assert V2[b] == pytest.approx(constraints_violation(ev, X[b], OBJ[b])[0], rel=1e-12, abs=1e-12)
        cols, vals, ub = benders_cut(ev, X[b], D[b])
        assert cuts[b][0] == cols
        np.testing.assert_allclose(cuts[b][1], vals, rtol=1e-11, atol=1e-13)
        assert cuts[b][2] == pytest.approx(ub, rel=1e-11, abs=1e-11)


@pytest.mark.gpu
def test_branching_tqp_parity(bnlp_lib):
    """QP strong branching: the QP surrogate built on the device at a node optimum equals the CPU
    restatement of BranchingTQP, and its children (bnlp_tqp_solve_batch) match the oracle IPM on that QP."""
    from bonmin_b200.engine import Engine, Problem
    from oracle import ipm as oipm
    from oracle import struct_eval as SE
    for m in (I.mytoy(), I.syn_miqcp(40, 20, 8, 8, 3)):
        eng = Engine(Problem.from_model(m))
        root = eng.solve_batch(m.xl, m.xu, m.x0)
        assert root["status"][0] == 1
        xs, du = root["x"][0], root["duals"][0]
        ev = SE.Evaluator(m)
        qm, f0 = SE.branching_tqp_model(m, ev, xs, du)
        P, f0_dev = eng.branching_tqp(xs, du)
        assert abs(f0 - f0_dev) <= 1e-9 * max(1.0, abs(f0))
        info = P.info()
        np.testing.assert_allclose(info["xl"], qm.xl, atol=1e-12)
        np.testing.assert_allclose(info["gu"], qm.gu, rtol=1e-9, atol=1e-9)
        # the device-built QP evaluates like the oracle-built one at a random displacement
        qe = Engine(P)
        rng = np.random.default_rng(1)
        d = rng.uniform(-0.1, 0.1, m.n)
        r = qe.eval_batch(d[None, :], want=("f", "g", "gradf"))
        vals, grads, _ = SE.Evaluator(qm).dense(d)
        assert abs(r["f"][0] - vals[0]) <= 1e-9 * max(1.0, abs(vals[0]))
        np.testing.assert_allclose(r["g"][0], vals[1:], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(r["gradf"][0], grads[0], rtol=1e-9, atol=1e-9)
        qe.close()
        # children: every integer variable both ways, on the QP surrogate
        ints = np.flatnonzero(m.vtype != 0)
        frac = [j for j in ints if abs(xs[j] - round(xs[j])) > 1e-6][:8]
        XL, XU = I.strong_branching_children(m, m.xl, m.xu, xs, frac)
        got = eng.tqp_solve_batch(xs, du, XL, XU)
        INF = 1e19
        dl = np.where(XL > -INF, XL - xs, XL)
        du_ = np.where(XU < INF, XU - xs, XU)
        ref = oipm.solve_batch(qm, dl, du_, np.zeros(m.n), threads=4)
        assert (got["status"] == ref["status"]).all()
        ok = ref["status"] == 1
        np.testing.assert_allclose(got["obj"][ok], ref["obj"][ok] + f0, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(got["x"][ok], ref["x"][ok] + xs, atol=1e-5)
        # a QP child is a lower-order model of the NLP child: close to it for these quadratic instances
        nlp = eng.solve_batch(XL, XU, xs)
        both = ok & (nlp["status"] == 1)
        if m.n > 10:
            assert both.sum() >= 4
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mk", MODELS)
def test_extract_linear_relaxation_parity(bnlp_lib, mk):
    """a10: the initial LP (extractLinearRelaxation + addObjectiveFunction) from the engine equals the CPU restatement:
    same rows, same columns, coefficients and bounds to 1e-12."""
    from oracle.struct_eval import extract_linear_relaxation
    m = mk()
    eng = _engine(m)
    ev = Evaluator(m)
    rng = np.random.default_rng(3)
    x = np.clip(m.x0 + rng.uniform(-0.2, 0.2, m.n), m.xl, m.xu)
    duals = np.concatenate([np.zeros(2 * m.n), rng.uniform(0.0, 1.0, m.m) * (rng.random(m.m) < 0.7)])
    for get_obj in (True, False):
        got = eng.extract_linear_relaxation(x, duals, get_obj=get_obj)
        ref = extract_linear_relaxation(ev, x, duals, get_obj=get_obj)
        assert got["n_rows"] == len(ref["rows"]) and got["has_alpha"] == ref["has_alpha"]
        assert got["n_cols"] == m.n + (1 if ref["has_alpha"] else 0)
        for r, (src, cols, vals, lb, ub) in enumerate(ref["rows"]):
            a, b = got["row_start"][r], got["row_start"][r + 1]
            assert got["src_row"][r] == src
            assert got["col_idx"][a:b].tolist() == cols
            np.testing.assert_allclose(got["val"][a:b], vals, rtol=1e-12, atol=1e-14)
            assert got["row_lb"][r] == pytest.approx(lb, rel=1e-12, abs=1e-12)
            assert got["row_ub"][r] == pytest.approx(ub, rel=1e-12, abs=1e-12)
        np.testing.assert_allclose(got["obj"][:len(ref["obj"])], ref["obj"], rtol=1e-12, atol=1e-14)
        assert got["obj_offset"] == pytest.approx(ref["obj_offset"], abs=1e-12)
        assert (got["is_integer"][:m.n] == (np.asarray(m.vtype) != 0)).all()
    eng.close()


@pytest.mark.gpu
def test_constraint_oa_batch_parity(bnlp_lib):
    """getConstraintOuterApproximation for single rows (linear and nonlinear) at a batch of points."""
    from oracle.struct_eval import constraint_outer_approximation
    m = I.syn_cvx_minlp(30, 10, 5, 8, 3)
    eng = _engine(m)
    ev = Evaluator(m)
    rng = np.random.default_rng(11)
    X = np.clip(m.x0 + rng.uniform(-0.2, 0.2, (5, m.n)), m.xl, m.xu)
    D = np.concatenate([np.zeros((5, 2 * m.n)), rng.uniform(0, 1, (5, m.m))], axis=1)
    for row in (0, 3, m.m - 1):
        got = eng.constraint_oa_batch(row, X, D)
        for b in range(5):
            cols, vals, lb, ub = constraint_outer_approximation(ev, row, X[b], D[b])
            assert got[b][0] == cols
            np.testing.assert_allclose(got[b][1], vals, rtol=1e-12, atol=1e-14)
            assert got[b][2] == pytest.approx(lb, rel=1e-12, abs=1e-12) and got[b][3] == pytest.approx(ub, rel=1e-12, abs=1e-12)
    eng.close()


@pytest.mark.gpu
def test_fp_variant_parity(bnlp_lib):
    """TNLP2FPNLP (§8f rank 3): the feasibility-pump / local-branching variant built by the product evaluates like the
    CPU restatement, and a batch of rounding targets (children that differ only in the bounds of the target variables)
    solves to the oracle's optima; with lambda = 1 and norm 2 an unconstrained-in-effect pump step is a projection."""
    from bonmin_b200.engine import Engine, Problem
    from oracle import ipm as oipm
    from oracle.struct_eval import fp_model
    rng = np.random.default_rng(4)
    for m, kw in [(I.syn_miqcp(20, 10, 4, 4, 9), dict(norm=2, lam=0.7, sigma=0.5, scaling=2.0)),
                  (I.syn_cvx_minlp(16, 6, 3, 5, 2), dict(norm=1, lam=1.0, sigma=1.0, scaling=1.0, rhs_local_branching=3.0)),
                  (I.syn_miqcp(12, 6, 3, 3, 4), dict(norm=2, lam=0.5, sigma=1.0, scaling=1.0, cutoff=-60.0)),
                  (I.syn_miqcp(12, 6, 3, 3, 4), dict(norm=2, lam=0.5, sigma=1.0, scaling=1.0, cutoff=-100.0))]:   # last: cut off
        inds = np.flatnonzero(np.asarray(m.vtype) != 0).astype(np.int32)
        base = Problem.from_model(m)
        P = base.fp_variant(inds, norm=kw["norm"], lam=kw["lam"], sigma=kw["sigma"], objective_scaling_factor=kw["scaling"],
                            cutoff=kw.get("cutoff"), rhs_local_branching=kw.get("rhs_local_branching"))
        ref = fp_model(m, inds, **kw)
        assert (P.n, P.m) == (ref.n, ref.m)
        eng = Engine(P)
        ev = Evaluator(ref)
        X = np.concatenate([np.clip(m.x0 + rng.uniform(-0.2, 0.2, (3, m.n)), m.xl, m.xu),
                            rng.integers(0, 2, (3, len(inds))).astype(float)], axis=1)
        r = eng.eval_batch(X, want=("f", "g", "gradf"))
        for b in range(3):
            vals, grads, _ = ev.dense(X[b])
            assert r["f"][b] == pytest.approx(vals[0], rel=1e-11, abs=1e-11)
            np.testing.assert_allclose(r["g"][b], vals[1:], rtol=1e-11, atol=1e-11)
            np.testing.assert_allclose(r["gradf"][b], grads[0], rtol=1e-10, atol=1e-10)
        # a batch of rounding targets: children differ only in the bounds of the target variables
        B = 6
        root = Engine(base).solve_batch(m.xl, m.xu, m.x0)["x"][0]
        T = np.clip(np.round(root[inds] + rng.uniform(-0.6, 0.6, (B, len(inds)))), m.xl[inds], m.xu[inds])
        if kw["norm"] == 1:
            T = np.clip(T, 0, 1)
        XL = np.tile(np.concatenate([m.xl, np.zeros(len(inds))]), (B, 1))
        XU = np.tile(np.concatenate([m.xu, np.zeros(len(inds))]), (B, 1))
        XL[:, m.n:] = T
        XU[:, m.n:] = T
        x0 = np.concatenate([root, T[0]])
        g = eng.solve_batch(XL, XU, x0)
        o = oipm.solve_batch(ref, XL, XU, x0, threads=4)
        assert (g["status"] == o["status"]).all()
        ok = o["status"] == 1
        assert ok.all() or kw.get("cutoff") == -100.0       # the last case is infeasible under its cutoff row
        np.testing.assert_allclose(g["obj"][ok], o["obj"][ok], rtol=1e-6, atol=1e-7)
        eng.close()


@pytest.mark.gpu
def test_linearizer_adapter_matches_the_batch_calls(bnlp_lib):
    """B200Linearizer (TMINLP2OsiLP: get_oas / extract) with the engine behind it returns exactly the rows of bnlp_oa_batch
    and the LP of bnlp_extract_linear_relaxation."""
    import ctypes
    from bonmin_b200 import engine as E_, instances as I
    m = I.syn_cvx_minlp(16, 6, 3, 5, 2)
    eng = E_.Engine(E_.Problem.from_model(m))
    rng = np.random.default_rng(1)
    X = np.clip(m.x0 + rng.uniform(-0.3, 0.3, size=(5, m.n)), m.xl, m.xu)
    ref = eng.oa_batch(X, None, m.xl, m.xu, get_obj=1)
    keep = np.flatnonzero(ref["src_row"] >= 0)
    cap_c, cap_z = len(ref["src_row"]), len(ref["val"])
    rs = np.zeros(cap_c + 1, np.int32); col = np.zeros(cap_z, np.int32); val = np.zeros(cap_z)
    lb = np.zeros(cap_c); ub = np.zeros(cap_c); gl = np.zeros(cap_c, np.int32)
    nc, lc, lr, lz, li = (ctypes.c_int() for _ in range(5))
    L = E_.lib()
    L.bnlp_linearizer_engine.argtypes = ([ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
                                         + [ctypes.c_void_p] * 7 + [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    rc = L.bnlp_linearizer_engine(eng.h, eng.problem.h, 5, vp(np.ascontiguousarray(X)), 1, 1, ctypes.byref(nc), vp(rs), vp(col), vp(val),
                                  vp(lb), vp(ub), vp(gl), cap_c, cap_z, ctypes.byref(lc), ctypes.byref(lr), ctypes.byref(lz), ctypes.byref(li))
    assert rc == 0 and nc.value == len(keep) and gl[:nc.value].all()
    for k, slot in enumerate(keep):
        a, b = ref["row_start"][slot], ref["row_start"][slot + 1]
        assert np.array_equal(col[rs[k]:rs[k + 1]], ref["col"][a:b]) and np.array_equal(val[rs[k]:rs[k + 1]], ref["val"][a:b])
        assert lb[k] == ref["lb"][slot] and ub[k] == ref["ub"][slot]
    lp = eng.extract_linear_relaxation(X[0], None, get_obj=True)
    assert lc.value == m.n + (1 if lp["has_alpha"] else 0) and lr.value == lp["n_rows"] and lz.value == lp["row_start"][lp["n_rows"]]
    assert li.value == int((np.asarray(m.vtype) != 0).sum())
    assert L.bnlp_linearizer_engine(eng.h, eng.problem.h, 5, vp(np.ascontiguousarray(X)), 1, 1, ctypes.byref(nc), vp(rs), vp(col), vp(val),
                                    vp(lb), vp(ub), vp(gl), 2, cap_z, None, None, None, None) == -2


def test_hessian_values_are_bit_reproducible(bnlp_lib):
    """eval_h with element terms: entries of different functions that meet in one position of the Hessian are summed in a
    fixed order (per-position lists), not with floating-point atomics: two evaluations give the same bits."""
    m = I.syn_cvx_minlp(40, 10, 12, 6, 9)
    eng = _engine(m)
    rng = np.random.default_rng(4)
    X = np.clip(m.x0 + rng.normal(0, 0.3, (64, m.n)), m.xl, m.xu)
    LAM = rng.normal(0, 1, (64, m.m))
    a = eng.eval_batch(X, LAM, sigma=0.9, want=("hess",))["hess"]
    for _ in range(3):
        assert np.array_equal(a, eng.eval_batch(X, LAM, sigma=0.9, want=("hess",))["hess"])
    eng.close()
