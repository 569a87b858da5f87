"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.
Bar (BASELINE.json north_star): same feasibility status per child, objective within 1e-6 relative."""
import numpy as np
import pytest

from bonmin_b200 import instances as I

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _engine(model, **kw):
    from bonmin_b200.engine import Engine, Problem
    return Engine(Problem.from_model(model), **kw)


def _compare(r, o, want_iters=False):
    assert (r["status"] == o["status"]).all(), np.flatnonzero(r["status"] != o["status"])
    ok = o["status"] == 1
    rel = np.abs(r["obj"] - o["obj"]) / np.maximum(1.0, np.abs(o["obj"]))
    assert rel[ok].max(initial=0.0) <= RTOL
    return ok


def test_mytoy_goldens(bnlp_lib, oracle_ipm):
    m = I.mytoy()
    XL = np.tile(m.xl, (3, 1))
    XU = np.tile(m.xu, (3, 1))
    XL[1, 2] = 1.0
    XU[2, 2] = 0.0
    eng = _engine(m)
    r = eng.solve_batch(XL, XU, m.x0)
    o = oracle_ipm.solve_batch(m, XL, XU, m.x0)
    _compare(r, o)
    sq5 = 5 ** 0.5
    assert r["obj"][0] == pytest.approx(-(1.5 + sq5 / 2), rel=1e-7)       # InterfaceTest.cpp:134
    assert r["duals"][0][8] == pytest.approx(sq5, rel=1e-6)               # :155-159
    assert r["duals"][0][9] == pytest.approx(1.0, rel=1e-6)
    assert r["x"][1][2] == 1.0 and r["x"][2][2] == 0.0                    # :165-189
    # the x >= 1 child at what Ipopt's 1e-8 relaxations predict analytically (tests/test_oracle_golden.py::test_mytoy_children)
    assert r["obj"][1] == pytest.approx(-(2.5 - 1e-8 + (2e-8 - 1e-16) ** 0.5), abs=2e-9)
    assert r["obj"][2] == pytest.approx(-(1 + 2 ** -0.5), rel=1e-6)
    assert list(r["status"]) == [1, 1, 1]
    np.testing.assert_allclose(r["x"], o["x"], atol=1e-6)
    np.testing.assert_allclose(r["duals"], o["duals"], atol=1e-5)


@pytest.mark.parametrize("shape,B", [((100, 50, 20, 20, 20261019), 256), ((200, 100, 40, 40, 20261020), 300),
                                     ((100, 50, 20, 20, 20261019), 640)])   # 640 >= 2 x SMs: the 256-thread, 2-CTA/SM kernel
def test_children_parity(bnlp_lib, oracle_ipm, shape, B):
    m = I.syn_miqcp(*shape)
    eng = _engine(m)
    root = eng.solve_batch(m.xl, m.xu, m.x0)
    oroot = oracle_ipm.solve_batch(m, m.xl, m.xu, m.x0)
    _compare(root, oroot)
    xs = root["x"][0]
    XL, XU = I.c3_children(m, xs, B)
    r = eng.solve_batch(XL, XU, xs)
    o = oracle_ipm.solve_batch(m, XL, XU, xs, threads=8)
    ok = _compare(r, o)
    assert ok.sum() > B // 2
    np.testing.assert_allclose(r["x"][ok], o["x"][ok], atol=1e-5)
    np.testing.assert_allclose(r["g"][ok], o["g"][ok], atol=1e-5)
    st = eng.stats()
    assert st["iterations"] == int(r["iters"].sum())


def test_infeasible_children_status(bnlp_lib, oracle_ipm):
    """Constraint-infeasible children (not just crossed bounds): status parity, and obj = max-norm violation
    (BonTMINLP2TNLP.cpp:503-505) within 1e-4 relative."""
    m = I.syn_miqcp(30, 15, 6, 6, 5)
    eng = _engine(m)
    xs = eng.solve_batch(m.xl, m.xu, m.x0)["x"][0]
    rng = np.random.default_rng(1)
    B = 200
    XL = np.tile(m.xl, (B, 1))
    XU = np.tile(m.xu, (B, 1))
    for b in range(B):
        for j in rng.choice(15, rng.integers(1, 12), replace=False):
            if rng.random() < 0.5:
                XL[b, j] = XU[b, j] = rng.integers(0, 11)
            else:
                v = rng.integers(0, 10)
                if rng.random() < 0.5:
                    XL[b, j] = v + 1
                else:
                    XU[b, j] = v
    r = eng.solve_batch(XL, XU, xs)
    o = oracle_ipm.solve_batch(m, XL, XU, xs, threads=8)
    _compare(r, o)
    inf = o["status"] == 3
    assert inf.sum() >= 10
    np.testing.assert_allclose(r["obj"][inf], o["obj"][inf], rtol=1e-4)


def test_edge_cases(bnlp_lib, oracle_ipm):
    m = I.syn_miqcp(20, 10, 4, 4, 9)
    eng = _engine(m)
    xs = eng.solve_batch(m.xl, m.xu, m.x0)["x"][0]
    XL = np.tile(m.xl, (4, 1))
    XU = np.tile(m.xu, (4, 1))
    XL[0] = XU[0] = np.round(m.x0)            # all fixed -> zeroDimension short cut (BonTNLPSolver.cpp:78-161)
    XL[1, 0], XU[1, 0] = 5.0, 4.0             # crossed bounds
    XL[2, :10] = XU[2, :10] = np.round(xs[:10])  # all integers fixed (the NLP(y) of OA)
    r = eng.solve_batch(XL, XU, xs)           # child 3: the root itself
    o = oracle_ipm.solve_batch(m, XL, XU, xs)
    _compare(r, o)
    assert r["iters"][0] == 0 and r["status"][1] == 3
    # per-child start points and B = 1
    r1 = eng.solve_batch(XL[3], XU[3], np.tile(xs, (1, 1)))
    assert r1["status"][0] == 1 and abs(r1["obj"][0] - r["obj"][3]) <= RTOL * abs(r["obj"][3])
    # outputs may be skipped
    r2 = eng.solve_batch(XL, XU, xs, want_x=False, want_g=False, want_duals=False)
    assert np.array_equal(r2["obj"], r["obj"])


def test_determinism_and_queue_order(bnlp_lib):
    """The work queue hands children to CTAs in a timing-dependent order; results must not depend on it."""
    m = I.syn_miqcp(40, 20, 8, 8, 3)
    eng = _engine(m)
    xs = eng.solve_batch(m.xl, m.xu, m.x0)["x"][0]
    XL, XU = I.c3_children(m, xs, 400)
    a = eng.solve_batch(XL, XU, xs)
    b = eng.solve_batch(XL, XU, xs)
    eng2 = _engine(m, ctas_per_device=7)
    c = eng2.solve_batch(XL, XU, xs)
    for k in ("obj", "status", "iters", "x"):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k])


def test_two_ctas_per_sm_variant_is_bit_identical(bnlp_lib):
    """n + 3 <= 128 and B >= 2 x SMs selects the 256-thread build of the cooperative kernel (two children per SM); a fixed
    CTA count keeps the 512-thread one.  Same arithmetic in the same order: identical bits."""
    m = I.syn_miqcp(100, 50, 20, 20, 20261019)
    xs = _engine(m).solve_batch(m.xl, m.xu, m.x0)["x"][0]
    XL, XU = I.c3_children(m, xs, 1000)
    a = _engine(m).solve_batch(XL, XU, xs)
    b = _engine(m, ctas_per_device=148).solve_batch(XL, XU, xs)
    for k in ("obj", "status", "iters", "x", "duals"):
        assert np.array_equal(a[k], b[k]), k


def test_solve_children_equals_explicit_bounds(bnlp_lib):
    """bnlp_solve_children (parents + one changed bound per child, expanded on the device) against bnlp_solve_batch on the
    host-expanded bound vectors: the same children, identical bits.  Includes per-parent start points and bad indices."""
    m = I.syn_miqcp(60, 30, 12, 12, 3)
    eng = _engine(m)
    xs = eng.solve_batch(m.xl, m.xu, m.x0)["x"][0]
    rng = np.random.default_rng(5)
    ints = np.flatnonzero(np.asarray(m.vtype) != 0)
    P, K = 5, 12
    PXL, PXU = np.tile(m.xl, (P, 1)), np.tile(m.xu, (P, 1))
    for q in range(1, P):                         # parents = the root with a few integers already branched on
        for j in rng.choice(ints, q, replace=False):
            if rng.random() < 0.5:
                PXU[q, j] = np.floor(xs[j])
            else:
                PXL[q, j] = np.ceil(xs[j])
    cp, cc, cw, cb = [], [], [], []
    for q in range(P):
        for j in rng.choice(ints, K, replace=False):
            for w in (0, 1):
                cp.append(q); cc.append(j); cw.append(w); cb.append(np.floor(xs[j]) if w == 0 else np.ceil(xs[j]))
    cp, cc, cw, cb = map(np.asarray, (cp, cc, cw, cb))
    XL, XU = PXL[cp].copy(), PXU[cp].copy()
    for b in range(len(cp)):
        (XU if cw[b] == 0 else XL)[b, cc[b]] = cb[b]
    a = eng.solve_children(PXL, PXU, cp, cc, cw, cb, xs, want_g=True, want_duals=True)
    b_ = eng.solve_batch(XL, XU, xs)
    for k in ("obj", "status", "iters", "x", "g", "duals"):
        assert np.array_equal(a[k], b_[k]), k
    assert (a["status"] == 1).sum() > len(cp) // 2
    X0p = np.tile(xs, (P, 1)) + 1e-3 * rng.standard_normal((P, m.n))
    a = eng.solve_children(PXL, PXU, cp, cc, cw, cb, X0p)
    b_ = eng.solve_batch(XL, XU, X0p[cp])
    assert np.array_equal(a["obj"], b_["obj"]) and np.array_equal(a["status"], b_["status"])
    from bonmin_b200.engine import BnlpError
    bad = cp.copy(); bad[3] = P
    with pytest.raises(BnlpError):
        eng.solve_children(PXL, PXU, bad, cc, cw, cb, xs)


def test_batch_replay_file_reproduces_the_batch(bnlp_lib, tmp_path):
    """A dumped batch (bonmin_b200.files.write_batch_replay) read back and solved again gives the same bits."""
    from bonmin_b200 import files
    m = I.syn_miqcp(40, 20, 8, 8, 3)
    eng = _engine(m)
    xs = eng.solve_batch(m.xl, m.xu, m.x0)["x"][0]
    XL, XU = I.c3_children(m, xs, 50)
    a = eng.solve_batch(XL, XU, xs)
    files.write_batch_replay(str(tmp_path / "b"), XL, XU, m.xl, m.xu, xs, m=m.m)
    rl, ru, x0 = files.read_batch_replay(str(tmp_path / "b"), m.xl, m.xu)
    b = eng.solve_batch(rl, ru, x0)
    for k in ("obj", "status", "iters", "x"):
        assert np.array_equal(a[k], b[k]), k


def test_size_independent_properties_full_batch(bnlp_lib):
    """BASELINE-size batch (C3, 4096 children) checked through properties that need no oracle:
    (i) optimal children are feasible for the original rows and bounds, (ii) a child's optimum is never below its
    parent's (relaxation monotonicity), (iii) KKT stationarity holds with the returned multipliers."""
    m = I.config_c3()
    eng = _engine(m)
    root = eng.solve_batch(m.xl, m.xu, m.x0)
    xs = root["x"][0]
    B = 4096
    XL, XU = I.c3_children(m, xs, B)
    r = eng.solve_batch(XL, XU, xs)
    ok = r["status"] == 1
    assert ok.sum() > B // 2 and set(np.unique(r["status"])) <= {1, 2, 3}
    assert (r["obj"][ok] >= root["obj"][0] - 1e-6 * abs(root["obj"][0])).all()
    assert (r["g"][ok] <= m.gu + 1e-6 * (1 + np.abs(m.gu))).all()
    assert (r["x"][ok] >= XL[ok]).all() and (r["x"][ok] <= XU[ok]).all()
    # stationarity on a sample: grad f + J' lambda - zL + zU = 0
    from bonmin_b200.instances import unpack_lower
    n = m.n
    Q = {int(fn): unpack_lower(m.qd[k], n) for k, fn in enumerate(m.qd_fn)}
    for b in np.flatnonzero(ok)[:16]:
        x = r["x"][b]
        lam = r["duals"][b][2 * n:]
        g = m.A[0] + Q[0] @ x
        for i in range(m.m):
            gi = m.A[i + 1] + (Q[i + 1] @ x if (i + 1) in Q else 0.0)
            g = g + lam[i] * gi
        g = g - r["duals"][b][:n] + r["duals"][b][n:2 * n]
        free = XU[b] > XL[b]
        assert np.abs(g[free]).max() <= 1e-5 * (1 + np.abs(m.A[0]).max())


def test_equality_rows_portfolio(bnlp_lib, oracle_ipm):
    """C5 shape (SYN-PORTFOLIO): a budget EQUALITY row handled through the Schur-complement path of the general kernel;
    strong-branching children on the binaries z_j."""
    for N, B in [(20, 60), (200, 48)]:
        m = I.syn_portfolio(N, 20261022)
        eng = _engine(m)
        root = eng.solve_batch(m.xl, m.xu, m.x0)
        oroot = oracle_ipm.solve_batch(m, m.xl, m.xu, m.x0)
        _compare(root, oroot)
        assert abs(root["x"][0][:N].sum() - 1.0) < 1e-6
        xs = root["x"][0]
        XL, XU = I.c3_children(m, xs, B)
        r = eng.solve_batch(XL, XU, xs)
        o = oracle_ipm.solve_batch(m, XL, XU, xs, threads=8)
        ok = _compare(r, o)
        np.testing.assert_allclose(r["duals"][ok][:, 2 * m.n + 1], o["duals"][ok][:, 2 * m.n + 1], rtol=1e-4, atol=1e-6)


def test_element_terms_cvx_minlp(bnlp_lib, oracle_ipm):
    """C4 shape (SYN-CVX-MINLP): exp() element terms + dense rank-4 quadratics + big-M rows; the NLP(y) solves of
    OA / node NLPs go through the general kernel (element terms evaluated by the second-order stack machine)."""
    for args, B in [((30, 10, 5, 8, 3), 40), ((120, 40, 20, 25, 5), 64)]:
        m = I.syn_cvx_minlp(*args)
        eng = _engine(m)
        root = eng.solve_batch(m.xl, m.xu, m.x0)
        oroot = oracle_ipm.solve_batch(m, m.xl, m.xu, m.x0)
        _compare(root, oroot)
        assert root["status"][0] == 1
        xs = root["x"][0]
        XL, XU = I.c3_children(m, xs, B)
        # a few children with all binaries fixed (the NLP(y) of OACutGenerator2::performOa)
        nb = args[1]
        XL[:4, :nb] = XU[:4, :nb] = np.round(xs[:nb])
        r = eng.solve_batch(XL, XU, xs)
        o = oracle_ipm.solve_batch(m, XL, XU, xs, threads=8)
        ok = _compare(r, o)
        assert ok.sum() >= B // 3
        # element terms are accumulated in a fixed order on the device (no floating-point atomics) but not in the
        # oracle's order: weakly determined components of x may differ by ~2e-5 at tol 1e-8
        np.testing.assert_allclose(r["x"][ok], o["x"][ok], atol=1e-4)
        # run-to-run: bitwise identical, whatever CTA picks up which child
        r2 = eng.solve_batch(XL[::-1].copy(), XU[::-1].copy(), xs)
        assert np.array_equal(r["obj"], r2["obj"][::-1]) and np.array_equal(r["x"], r2["x"][::-1])
        assert np.array_equal(r["iters"], r2["iters"][::-1])


def test_randomised_shapes(bnlp_lib, oracle_ipm):
    """Random shapes, seeds and batch sizes (1 .. 300 children, groups of 32 not filled, more/fewer children than SMs):
    same status for every child, objectives within 1e-6 (scripts/stress_parity.py is the long version)."""
    rng = np.random.default_rng(20261020)
    for _ in range(10):
        n = int(rng.choice([30, 48, 64, 100, 128]))
        m = I.syn_miqcp(n, n // 2, int(rng.integers(1, n // 5)), int(rng.integers(0, n // 5)), int(rng.integers(1, 10 ** 6)))
        B = int(rng.choice([1, 7, 33, 64, 149, 300]))
        eng = _engine(m)
        root = eng.solve_batch(m.xl, m.xu, m.x0)
        if root["status"][0] != 1:
            eng.close()
            continue
        xs = root["x"][0]
        XL, XU = I.c3_children(m, xs, B)
        for b in range(0, B, 5):
            j = int(rng.integers(0, n // 2))
            XL[b, j] = XU[b, j] = np.round(rng.uniform(m.xl[j], m.xu[j]))
        _compare(eng.solve_batch(XL, XU, xs), oracle_ipm.solve_batch(m, XL, XU, xs, threads=8))
        eng.close()


def _custom(n, m, Q0=None, c0=None, A=None, gl=None, gu=None, xl=None, xu=None, x0=None, qd_rows=()):
    """hand-built StructuredModel: objective 1/2 x'Q0 x + c0.x (Q0 may be None), rows A x (+ 1/2 x'Qi x for qd_rows)."""
    INF = 1e19
    Afull = np.zeros((m + 1, n))
    if c0 is not None:
        Afull[0] = c0
    if A is not None:
        Afull[1:] = A
    qd_fn, qd = [], []
    if Q0 is not None:
        qd_fn.append(0)
        qd.append(I.pack_lower(Q0))
    for i, Qi in qd_rows:
        qd_fn.append(i)
        qd.append(I.pack_lower(Qi))
    return I._empty_sparse(I.StructuredModel(
        n=n, m=m, xl=np.full(n, -INF) if xl is None else np.asarray(xl, float), xu=np.full(n, INF) if xu is None else np.asarray(xu, float),
        gl=np.full(m, -INF) if gl is None else np.asarray(gl, float), gu=np.full(m, INF) if gu is None else np.asarray(gu, float),
        vtype=np.zeros(n, np.int32), x0=np.zeros(n) if x0 is None else np.asarray(x0, float), c=np.zeros(m + 1), A=Afull,
        qd_fn=np.array(qd_fn, np.int32), qd=np.array(qd).reshape(len(qd_fn), n * (n + 1) // 2), name="custom"))


def test_unusual_model_shapes(bnlp_lib, oracle_ipm):
    """Shapes at the edge of the interface: no rows at all, one variable, free variables, a linear program (no
    quadratic block anywhere), an equality-only model, a ranged row; each with a small batch of bound changes."""
    rng = np.random.default_rng(12)
    G = rng.standard_normal((6, 6))
    Q = G @ G.T + np.eye(6)
    models = {
        "box QP without rows": _custom(6, 0, Q0=Q, c0=rng.standard_normal(6), xl=-np.ones(6), xu=np.ones(6)),
        "one variable": _custom(1, 1, Q0=np.array([[2.0]]), c0=[-3.0], A=[[1.0]], gu=[1.0], xl=[-5.0], xu=[5.0]),
        "free variables": _custom(6, 2, Q0=Q, c0=rng.standard_normal(6), A=rng.standard_normal((2, 6)), gu=[1.0, 2.0]),
        "linear program": _custom(5, 3, c0=[1, 2, -1, 0.5, -2], A=rng.standard_normal((3, 5)), gu=[1, 1, 1], gl=[-1, -1, -1],
                                  xl=-np.ones(5), xu=np.ones(5)),
        "equality only": _custom(6, 2, Q0=Q, A=rng.standard_normal((2, 6)), gl=[0.5, -0.25], gu=[0.5, -0.25], xl=-3 * np.ones(6),
                                 xu=3 * np.ones(6)),
        "ranged quadratic row": _custom(6, 1, c0=rng.standard_normal(6), qd_rows=[(1, Q)], gl=[0.5], gu=[4.0], xl=-3 * np.ones(6),
                                        xu=3 * np.ones(6)),
    }
    for name, m in models.items():
        eng = _engine(m)
        B = 5
        XL = np.tile(m.xl, (B, 1))
        XU = np.tile(m.xu, (B, 1))
        XU[1, 0] = min(m.xu[0], 0.25)
        XL[2, 0] = max(m.xl[0], -0.25)
        XL[3, -1] = XU[3, -1] = 0.1           # one variable fixed
        XL[4, 0], XU[4, 0] = 0.3, 0.2         # crossed bounds
        r = eng.solve_batch(XL, XU, m.x0)
        o = oracle_ipm.solve_batch(m, XL, XU, m.x0, threads=4)
        assert (r["status"] == o["status"]).all(), (name, r["status"], o["status"])
        ok = o["status"] == 1
        assert ok[:4].any() or name == "linear program", name
        np.testing.assert_allclose(r["obj"][ok], o["obj"][ok], rtol=1e-6, atol=1e-8, err_msg=name)
        eng.close()


def test_largest_supported_model(bnlp_lib, oracle_ipm):
    """n = 512: the largest model the batched IPM accepts (condensed matrix in global memory); n = 513 is refused loudly."""
    from bonmin_b200.engine import BnlpError
    m = I.syn_miqcp(512, 8, 2, 3, 77)
    eng = _engine(m)
    XL = np.tile(m.xl, (2, 1))
    XU = np.tile(m.xu, (2, 1))
    XU[1, 0] = 3.0
    r = eng.solve_batch(XL, XU, m.x0)
    o = oracle_ipm.solve_batch(m, XL, XU, m.x0, threads=2)
    _compare(r, o)
    eng.close()
    big = _engine(I.syn_miqcp(513, 4, 1, 1, 78))
    with pytest.raises(BnlpError, match="512"):
        big.solve_batch(XL[:, :1].repeat(513, 1), XU[:, :1].repeat(513, 1), np.zeros(513))
    big.close()


def test_reused_and_pinned_result_buffers(bnlp_lib):
    """solve_batch(out=, pin=) and oa_batch(out=, pin=): reused, page-locked result arrays hold the same results."""
    m = I.syn_miqcp(40, 20, 8, 8, 3)
    eng = _engine(m)
    xs = eng.solve_batch(m.xl, m.xu, m.x0)["x"][0]
    XL, XU = I.c3_children(m, xs, 64)
    a = eng.solve_batch(XL, XU, xs)
    r = eng.solve_batch(XL, XU, xs, pin=True)
    ids = {k: id(v) for k, v in r.items()}
    r["obj"][:] = 0
    r2 = eng.solve_batch(XL, XU, xs, out=r)
    assert all(id(r2[k]) == ids[k] for k in ids)
    for k in ("obj", "status", "iters", "x", "g", "duals"):
        assert np.array_equal(a[k], r2[k]), k
    X = np.clip(xs + 0.1, m.xl, m.xu)[None, :].repeat(8, 0)
    c0 = eng.oa_batch(X, None, m.xl, m.xu, get_obj=1)
    c1 = eng.oa_batch(X, None, m.xl, m.xu, get_obj=1, pin=True)
    c2 = eng.oa_batch(X, None, m.xl, m.xu, get_obj=1, out=c1)
    for k in ("row_start", "src_row", "lb", "ub"):
        assert np.array_equal(c0[k], c2[k]), k
    nnz = c0["row_start"][-1]
    assert np.array_equal(c0["col"][:nnz], c2["col"][:nnz]) and np.array_equal(c0["val"][:nnz], c2["val"][:nnz])
    del r, r2, c1, c2        # finalizers unregister the pages


def test_two_devices_in_one_process_match_one_device(bnlp_lib):
    """Engine over two devices (contiguous child shards, no data-path collective): same bits as one device, for both
    batch forms and for a batch large enough to select the 256-thread kernel on every device (its shared-memory opt-in is
    per device).  Skipped on a single-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from bonmin_b200.engine import Engine, Problem
    m = I.syn_miqcp(100, 50, 20, 20, 20261019)
    p = Problem.from_model(m)
    e1, e2 = Engine(p), Engine(p, devices=[0, 1])
    xs = e1.solve_batch(m.xl, m.xu, m.x0)["x"][0]
    for B in (64, 801):
        XL, XU = I.c3_children(m, xs, B)
        a, b = e1.solve_batch(XL, XU, xs), e2.solve_batch(XL, XU, xs)
        for k in ("obj", "status", "iters", "x", "g", "duals"):
            assert np.array_equal(a[k], b[k]), (B, k)
    ints = np.flatnonzero(np.asarray(m.vtype) != 0)[:30]
    cp = np.zeros(60, np.int32); cc = np.repeat(ints, 2).astype(np.int32); cw = np.tile([0, 1], 30).astype(np.int32)
    cb = np.where(cw == 0, np.floor(xs[cc]), np.ceil(xs[cc]))
    a = e1.solve_children(m.xl[None, :], m.xu[None, :], cp, cc, cw, cb, xs)
    b = e2.solve_children(m.xl[None, :], m.xu[None, :], cp, cc, cw, cb, xs)
    assert np.array_equal(a["obj"], b["obj"]) and np.array_equal(a["status"], b["status"]) and np.array_equal(a["x"], b["x"])


@pytest.mark.gpu
def test_warm_started_resolve_parity(bnlp_lib, oracle_ipm):
    """bnlp_solve_batch_warm = IpoptSolver::ReOptimizeTNLP after setWarmStart (BonIpoptSolver.cpp:180-220, 369-419; the
    sequence of test/InterfaceTest.cpp:165-189: solve, getWarmStart, change a bound, setWarmStart, resolve): primal AND
    dual start, warm_start_init_point = yes pushes (1e-3).  Same statuses and objectives as the CPU oracle, and as the cold
    solves of the same children (1e-6); re-solving the parent itself from its own optimum takes a few iterations."""
    from bonmin_b200.engine import Engine, Problem
    for m, B in ((I.syn_miqcp(40, 20, 8, 8, 3), 64), (I.config_c2(), 96), (I.mytoy(), 0)):
        eng = Engine(Problem.from_model(m))
        root = eng.solve_batch(m.xl, m.xu, m.x0)
        xs, du = root["x"][0], root["duals"][0]
        again = eng.solve_batch(m.xl, m.xu, xs, dual0=du)
        assert again["status"][0] == 1 and again["iters"][0] < root["iters"][0]
        assert again["obj"][0] == pytest.approx(root["obj"][0], rel=1e-7)
        if B:
            XL, XU = I.c3_children(m, xs, B)
        else:                                   # mytoy: the two children of test/InterfaceTest.cpp:165-189 (x >= 1, x <= 0)
            XL, XU = np.tile(m.xl, (2, 1)), np.tile(m.xu, (2, 1))
            XL[0, 2] = 1.0
            XU[1, 2] = 0.0
        cold = eng.solve_batch(XL, XU, xs)
        warm = eng.solve_batch(XL, XU, xs, dual0=du)
        per_child = eng.solve_batch(XL, XU, xs, dual0=np.tile(du, (len(XL), 1)))       # the B-row form of DUAL0
        o = oracle_ipm.solve_batch(m, XL, XU, xs, threads=4, dual0=du)
        assert np.array_equal(warm["status"], o["status"]) and np.array_equal(warm["status"], cold["status"])
        assert np.array_equal(per_child["status"], warm["status"]) and np.array_equal(per_child["iters"], warm["iters"])
        ok = o["status"] == 1
        assert np.allclose(warm["obj"][ok], o["obj"][ok], rtol=1e-9, atol=1e-9)
        assert np.allclose(warm["obj"][ok], cold["obj"][ok], rtol=1e-6, atol=1e-8)
        assert np.abs(warm["iters"] - o["iters"]).max() <= 1
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("B", [4096, 65536])
def test_c3_baseline_batches_against_the_oracle_on_a_strided_sample(bnlp_lib, oracle_ipm, B):
    """BASELINE config 3 at its own sizes ("1k-64k independent child NLPs", n = 200): the whole batch is solved on the GPU
    and a strided sample of 512 children — the same relaxations, solved alone — is checked against the CPU oracle: same
    status, objective within 1e-6 relative (north_star), same iteration count (the two run the same algorithm)."""
    m = I.config_c3()
    eng = _engine(m)
    root = eng.solve_batch(m.xl, m.xu, m.x0)
    xs = root["x"][0]
    XL, XU = I.c3_children(m, xs, B)
    r = eng.solve_batch(XL, XU, xs, want_g=False, want_duals=False)
    assert set(np.unique(r["status"])) <= {1, 2, 3}
    idx = np.unique(np.linspace(0, B - 1, 512).astype(int))
    o = oracle_ipm.solve_batch(m, XL[idx], XU[idx], xs, threads=16)
    assert np.array_equal(r["status"][idx], o["status"])
    ok = o["status"] == 1
    rel = np.abs(r["obj"][idx] - o["obj"]) / np.maximum(1.0, np.abs(o["obj"]))
    assert rel[ok].max() <= 1e-6
    assert np.abs(r["iters"][idx] - o["iters"]).max() <= 1
    np.testing.assert_allclose(r["x"][idx][ok], o["x"][ok], rtol=0, atol=1e-6 * 10.0)
    # the same child gives the same bits wherever it sits in a batch (a sample re-solved as a small batch)
    again = eng.solve_batch(XL[idx[:64]], XU[idx[:64]], xs, want_g=False, want_duals=False)
    assert np.array_equal(again["obj"], r["obj"][idx[:64]]) and np.array_equal(again["status"], r["status"][idx[:64]])
    eng.close()


@pytest.mark.gpu
def test_c4_full_size_children_and_oa_rows(bnlp_lib, oracle_ipm):
    """BASELINE config 4 at full size (SYN-CVX-MINLP n = 500: 100 dense quadratic rows + 1000 exp element terms): the NLP(y)
    and node relaxations that B-OA / B-QG solve, and the outer-approximation rows of all nonlinear constraints at their
    solutions, against the CPU oracle."""
    from oracle.struct_eval import Evaluator, outer_approximation
    m = I.config_c4()
    assert m.n == 500
    eng = _engine(m)
    root = eng.solve_batch(m.xl, m.xu, m.x0)
    oroot = oracle_ipm.solve_batch(m, m.xl, m.xu, m.x0)
    _compare(root, oroot)
    assert root["status"][0] == 1
    xs = root["x"][0]
    B = 24
    XL, XU = I.c3_children(m, xs, B)
    nb = int((np.asarray(m.vtype) != 0).sum())
    ints = np.flatnonzero(np.asarray(m.vtype) != 0)
    XL[:6, ints] = XU[:6, ints] = np.round(xs[ints])            # NLP(y): every integer fixed (OACutGenerator2::performOa)
    r = eng.solve_batch(XL, XU, xs)
    o = oracle_ipm.solve_batch(m, XL, XU, xs, threads=16)
    ok = _compare(r, o)
    assert ok.sum() >= 6 and nb > 0
    # OA rows at the optimal points, cut sides from the NLP multipliers (getOuterApproximation, getObj = 1)
    ev = Evaluator(m)
    pts = np.flatnonzero(ok)[:4]
    out = eng.oa_batch(r["x"][pts], DUALS=r["duals"][pts], get_obj=1)
    rpp = out["rows_per_point"]
    for q, b in enumerate(pts):
        ref = outer_approximation(ev, r["x"][b], duals=r["duals"][b], get_obj=1)
        got = []
        for srow in range(rpp):
            k = q * rpp + srow
            if out["src_row"][k] < 0:
                continue
            a, e = out["row_start"][k], out["row_start"][k + 1]
            got.append((int(out["src_row"][k]), out["col"][a:e].tolist(), out["val"][a:e], out["lb"][k], out["ub"][k]))
        assert [c[0] for c in got] == [c[0] for c in ref]
        for gc, rc in zip(got, ref):
            assert gc[1] == rc[1]
            np.testing.assert_allclose(gc[2], rc[2], rtol=1e-11, atol=1e-13)
            assert gc[4] == pytest.approx(rc[4], rel=1e-11, abs=1e-11)
    eng.close()


def test_sparse_linear_rows_in_the_cooperative_kernel(bnlp_lib, oracle_ipm, monkeypatch):
    """Round 2: linear rows with few coefficients leave the staged rank-16 updates of the cooperative kernel (entries of
    J' Sigma J scattered from slot-major lists, bnlp_api.cu build_ipm_view).  The model puts them where the lists are
    stressed: sparse linear rows BEFORE the quadratic ones, ranged (two one-sided rows each), some that share columns, one
    empty of structure but for a single coefficient; half of the children are infeasible, so the phase-1 elastic's row
    and column are live.  Against the oracle, and against the same engine with every row staged (BNLP_SPARSE_ROW_MAX=0)."""
    n, m_l, m_q = 120, 24, 10
    rng = np.random.default_rng(6)
    base = I.syn_miqcp(n, 40, m_q, 0, 99)
    A = np.zeros((1 + m_l + m_q, n))
    A[0] = base.A[0]
    gl = np.full(m_l + m_q, -I.INF)
    gu = np.zeros(m_l + m_q)
    x0 = np.asarray(base.x0)
    for i in range(m_l):                      # sparse linear rows first: 1 .. 14 coefficients, clustered columns
        k = 1 if i == 0 else int(rng.integers(2, 15))
        cols = rng.choice(30 if i % 3 == 0 else n, k, replace=False)
        A[1 + i, cols] = rng.standard_normal(k)
        act = A[1 + i] @ x0
        gu[i] = act + rng.uniform(0.5, 1.5)
        if i % 2 == 0:
            gl[i] = act - rng.uniform(2.0, 4.0)                  # ranged
    A[1 + m_l:] = base.A[1:]
    gu[m_l:] = base.gu
    mdl = I.StructuredModel(dict(base))
    mdl.update(m=m_l + m_q, A=A, c=np.zeros(1 + m_l + m_q), gl=gl, gu=gu,
               qd_fn=np.array([0] + [1 + m_l + i for i in range(m_q)], np.int32), name="sparse rows first, ranged")
    eng = _engine(mdl)
    root = eng.solve_batch(mdl.xl, mdl.xu, mdl.x0)
    oroot = oracle_ipm.solve_batch(mdl, mdl.xl, mdl.xu, mdl.x0)
    _compare(root, oroot)
    xs = root["x"][0]
    XL, XU = I.c3_children(mdl, xs, 320)
    XL[::2, 40:60] = np.minimum(XL[::2, 40:60] + 3.0, XU[::2, 40:60])       # squeeze every other child: many infeasible
    r = eng.solve_batch(XL, XU, xs)
    o = oracle_ipm.solve_batch(mdl, XL, XU, xs, threads=8)
    ok = _compare(r, o)
    assert ok.sum() >= 40 and (o["status"] == 3).sum() >= 10
    assert np.abs(r["iters"] - o["iters"]).max() <= 2
    eng.close()
    monkeypatch.setenv("BNLP_SPARSE_ROW_MAX", "0")
    dense = _engine(mdl)
    d = dense.solve_batch(XL, XU, xs)
    monkeypatch.delenv("BNLP_SPARSE_ROW_MAX")
    _compare(d, o)
    rel = np.abs(d["obj"] - r["obj"]) / np.maximum(1.0, np.abs(r["obj"]))
    assert rel[ok].max() <= 1e-9 and (d["status"] == r["status"]).all()
    dense.close()
