"""Seedable synthetic MINLP instance generators and the
child-batch generator of BASELINE config C3.

These produce *structured models*: every function k (k=0 objective, k=1..m rows) is

    F_k(x) = c_k + a_k.x + 1/2 x'Q_k x + sum_t phi_t(x[S_t])

stored as a python dict of numpy arrays (see `StructuredModel`).  The same dict feeds
the product (through `bnlp_problem_from_tape`) and the CPU oracle, and can be written
as an AMPL `.nl` file (`bonmin_b200.nlwrite`).  This module is input generation only:
no solver arithmetic lives here.
"""
from __future__ import annotations

import numpy as np

INF = 1e19  # Bonmin's infinity: min(|nlp_lower_bound_inf|,|nlp_upper_bound_inf|), BonOsiTMINLPInterface.cpp:3002-3005

# variable types as in Bonmin::TMINLP::VariableType (BonTMINLP.hpp:192-196)
CONTINUOUS, BINARY, INTEGER = 0, 1, 2


def pack_lower(Q: np.ndarray) -> np.ndarray:
    """Row-major packed lower triangle (r>=c) of a symmetric matrix."""
    n = Q.shape[0]
    il = np.tril_indices(n)
    return np.ascontiguousarray(Q[il])


def unpack_lower(p: np.ndarray, n: int) -> np.ndarray:
    Q = np.zeros((n, n))
    il = np.tril_indices(n)
    Q[il] = p
    Q = Q + Q.T - np.diag(np.diag(Q))
    return Q


class StructuredModel(dict):
    """dict with keys:
    n, m                       ints
    xl, xu (n) gl, gu (m)      bounds (|.|>=1e19 = infinite)
    vtype (n) int32            0 continuous, 1 binary, 2 integer
    x0 (n)                     starting point
    c (m+1)                    constants, index 0 = objective
    A (m+1, n)                 dense linear coefficients (converted to CSR at the C-ABI)
    qd_fn (nqd) int32          function index of each dense quadratic block
    qd (nqd, n(n+1)/2)         packed lower triangles, convention 1/2 x'Qx
    qs_fn,qs_i,qs_j (nqs) int32, qs_v (nqs)   sparse quadratic triplets, i>=j, Q_ij=Q_ji=v
    terms                      list of (fn, vars[int], prog[(op, iarg, darg)]) general nonlinear element terms
    name
    """

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


def _empty_sparse(model):
    model.setdefault("qs_fn", np.zeros(0, np.int32))
    model.setdefault("qs_i", np.zeros(0, np.int32))
    model.setdefault("qs_j", np.zeros(0, np.int32))
    model.setdefault("qs_v", np.zeros(0))
    model.setdefault("terms", [])
    return model


def syn_miqcp(n: int, n_int: int, m_q: int, m_l: int, seed: int) -> StructuredModel:
    """SYN-MIQCP(n, n_int, m_q, m_l, seed) of SURVEY.md Appendix B (C2/C3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = m_q + m_l
    xl = np.where(np.arange(n) < n_int, 0.0, -10.0)
    xu = np.full(n, 10.0)
    vtype = np.where(np.arange(n) < n_int, INTEGER, CONTINUOUS).astype(np.int32)
    x0 = np.empty(n)
    x0[:n_int] = rng.integers(2, 9, size=n_int)
    x0[n_int:] = rng.uniform(-5, 5, size=n - n_int)
    B0 = rng.standard_normal((n, n)) / np.sqrt(n)
    Q0 = B0.T @ B0 + 0.1 * np.eye(n)
    xhat = x0 + 3.0 * rng.standard_normal(n)
    c0 = -Q0 @ xhat
    A = np.zeros((m + 1, n))
    c = np.zeros(m + 1)
    gl = np.full(m, -INF)
    gu = np.zeros(m)
    A[0] = c0
    qd = [pack_lower(Q0)]
    qd_fn = [0]
    for i in range(1, m_q + 1):
        Bi = rng.standard_normal((8, n)) / np.sqrt(8.0)
        Qi = Bi.T @ Bi
        ai = rng.standard_normal(n)
        q0 = 0.5 * x0 @ Qi @ x0 + ai @ x0
        bi = q0 + 0.05 * (1 + abs(q0)) * rng.uniform(0.5, 1.5)
        A[i] = ai
        gu[i - 1] = bi
        qd.append(pack_lower(Qi))
        qd_fn.append(i)
    for i in range(m_q + 1, m + 1):
        li = rng.standard_normal(n)
        li[rng.random(n) < 0.9] = 0.0
        di = li @ x0 + rng.uniform(0.5, 1.5)
        A[i] = li
        gu[i - 1] = di
    return _empty_sparse(StructuredModel(
        n=n, m=m, xl=xl, xu=xu, gl=gl, gu=gu, vtype=vtype, x0=x0, c=c, A=A,
        qd_fn=np.array(qd_fn, np.int32), qd=np.array(qd),
        name=f"SYN-MIQCP(n={n},n_int={n_int},m_q={m_q},m_l={m_l},seed={seed})"))


def config_c2() -> StructuredModel:
    return syn_miqcp(100, 50, 20, 20, 20261019)


def config_c3() -> StructuredModel:
    return syn_miqcp(200, 100, 40, 40, 20261020)


def c3_children(model: StructuredModel, xstar: np.ndarray, B: int, first: int = 0):
    """Child-batch generator for C3: child k = root bounds +
    d_k ~ U{0..20} random prior integer branchings + final branch on j = k mod n_int
    (down if k even else up) at floor/ceil of the root optimum x*.
    Returns XL, XU of shape (B, n).  Children may be infeasible (intended)."""
    n = model.n
    ints = np.flatnonzero(model.vtype != CONTINUOUS)
    n_int = len(ints)
    XL = np.tile(model.xl, (B, 1))
    XU = np.tile(model.xu, (B, 1))
    for b in range(B):
        k = first + b
        r = np.random.Generator(np.random.PCG64(7 + k))
        d = min(int(r.integers(0, 21)), n_int)
        if d:
            for jj in r.choice(n_int, d, replace=False):
                j = ints[jj]
                up = r.random() < 0.5
                if abs(xstar[j] - round(xstar[j])) <= 1e-6:
                    continue
                if up:
                    XL[b, j] = max(XL[b, j], np.ceil(xstar[j]))
                else:
                    XU[b, j] = min(XU[b, j], np.floor(xstar[j]))
        j = ints[k % n_int]
        if k % 2 == 0:
            XU[b, j] = min(XU[b, j], np.floor(xstar[j]))
        else:
            XL[b, j] = max(XL[b, j], np.ceil(xstar[j]))
    return XL, XU


def strong_branching_children(model, xl, xu, xsol, cand):
    """2K children of one B-BB node (BonChooseVariable.cpp:807-905): for each candidate
    column j: down child xu_j := floor(x_j), up child xl_j := ceil(x_j).
    Order: [cand0 down, cand0 up, cand1 down, ...]."""
    K = len(cand)
    XL = np.tile(xl, (2 * K, 1))
    XU = np.tile(xu, (2 * K, 1))
    for i, j in enumerate(cand):
        XU[2 * i, j] = np.floor(xsol[j])
        XL[2 * i + 1, j] = np.ceil(xsol[j])
    return XL, XU


def mytoy() -> StructuredModel:
    """The model of the reference's only fixture (examples/amplExamples/toy.mod:5-14,
    variable order of test/mytoy.nl:10: y1, y2, x, z), written out structurally:
      min -y1 - y2 - x
      (y1-1/2)^2 + (y2-1/2)^2 <= 1/4 ;  x - y1 <= 0 ;  x + y2 + z <= 2."""
    n, m = 4, 3
    A = np.zeros((m + 1, n))
    c = np.zeros(m + 1)
    A[0, :3] = -1.0
    # row 1: y1^2 - y1 + y2^2 - y2 + 1/2
    A[1, 0] = A[1, 1] = -1.0
    c[1] = 0.5
    A[2, 0], A[2, 2] = -1.0, 1.0
    A[3, 1] = A[3, 2] = A[3, 3] = 1.0
    return _empty_sparse(StructuredModel(
        n=n, m=m, xl=np.zeros(n), xu=np.array([INF, INF, 1.0, 5.0]),
        gl=np.full(m, -INF), gu=np.array([0.25, 0.0, 2.0]),
        vtype=np.array([CONTINUOUS, CONTINUOUS, BINARY, INTEGER], np.int32),
        x0=np.zeros(n), c=c, A=A,
        qd_fn=np.zeros(0, np.int32), qd=np.zeros((0, n * (n + 1) // 2)),
        qs_fn=np.array([1, 1], np.int32), qs_i=np.array([0, 1], np.int32),
        qs_j=np.array([0, 1], np.int32), qs_v=np.array([2.0, 2.0]),
        name="mytoy"))


# opcodes of include/bnlp.h (BNLP_OP_*)
OP_CONST, OP_VAR, OP_MUL, OP_EXP = 0, 1, 4, 9


def syn_cvx_minlp(n: int, n_bin: int, m_nl: int, m_l: int, seed: int, dense_quadratic: bool = True) -> StructuredModel:
    """SYN-CVX-MINLP(n, n_bin, m_nl, m_l, seed) of SURVEY.md Appendix B (C4):
    g_i(x) = sum_{j in S_i} exp(alpha_ij x_j) + 1/2 x'(B_i'B_i)x + a_i'x <= b_i, |S_i| = 10, rank-4 B_i; linear
    objective; big-M linking rows x_c <= 5 z_b are part of m_l."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = m_nl + m_l
    xl = np.zeros(n)
    xu = np.where(np.arange(n) < n_bin, 1.0, 5.0)
    vtype = np.where(np.arange(n) < n_bin, BINARY, CONTINUOUS).astype(np.int32)
    x0 = np.empty(n)
    x0[:n_bin] = rng.integers(0, 2, size=n_bin)
    x0[n_bin:] = rng.uniform(0.5, 2.0, size=n - n_bin)
    npair = min(n_bin, n - n_bin, m_l)
    for p in range(npair):          # keep x0 feasible for the linking rows
        if x0[p] < 0.5:
            x0[n_bin + p] = 0.0
    cobj = np.where(np.arange(n) < n_bin, rng.uniform(0, 2, n), rng.uniform(-1, 1, n))
    A = np.zeros((m + 1, n))
    c = np.zeros(m + 1)
    gl = np.full(m, -INF)
    gu = np.zeros(m)
    A[0] = cobj
    qd, qd_fn, terms = [], [], []
    for i in range(1, m_nl + 1):
        S = rng.choice(n, 10, replace=False)
        alpha = rng.uniform(0.1, 0.5, 10)
        Bi = rng.standard_normal((4, n)) / 2.0
        Qi = Bi.T @ Bi
        ai = rng.standard_normal(n)
        ai[rng.random(n) < 0.9] = 0.0
        gi = float(np.exp(alpha * x0[S]).sum() + 0.5 * x0 @ Qi @ x0 + ai @ x0)
        A[i] = ai
        gu[i - 1] = gi + 0.1 * (1 + abs(gi))
        qd.append(pack_lower(Qi))
        qd_fn.append(i)
        for j, al in zip(S, alpha):
            terms.append((i, [int(j)], [(OP_CONST, 0, float(al)), (OP_VAR, 0, 0.0), (OP_MUL, 0, 0.0),
                                          (OP_EXP, 0, 0.0)]))
    for r in range(m_l):
        i = m_nl + 1 + r
        if r < npair:
            A[i, n_bin + r] = 1.0
            A[i, r] = -5.0
            gu[i - 1] = 0.0
        else:
            li = rng.standard_normal(n)
            li[rng.random(n) < 0.9] = 0.0
            A[i] = li
            gu[i - 1] = li @ x0 + rng.uniform(0.5, 1.5)
    return _empty_sparse(StructuredModel(
        n=n, m=m, xl=xl, xu=xu, gl=gl, gu=gu, vtype=vtype, x0=x0, c=c, A=A,
        qd_fn=np.array(qd_fn, np.int32), qd=np.array(qd).reshape(len(qd_fn), n * (n + 1) // 2), terms=terms,
        name=f"SYN-CVX-MINLP(n={n},n_bin={n_bin},m_nl={m_nl},m_l={m_l},seed={seed})"))


OP_ADD, OP_LOG, OP_SQRT, OP_SQR = 2, 10, 11, 14


def wide_terms_minlp(seed: int = 1) -> StructuredModel:
    """A small convex MINLP whose nonlinear rows couple MANY variables in one non-quadratic expression — what the narrow
    element-term machine (<= 6 variables) cannot hold and the wide build (include/bnlp.h, BNLP_MAX_WIDE_TERM_VARS) is for:
      row 1: log(sum_{i<20} exp(0.3 x_i)) <= b1          (log-sum-exp over 20 variables)
      row 2: sqrt(1 + sum_{8<=i<20} x_i^2) <= b2          (a norm over 12 variables)
      row 3: exp(0.5 x_12 + 0.5 x_13) + x_28 <= b3        (a narrow term in the same model)
      rows 4..6: sparse linear; linear objective.
    n = 30 in AMPL's canonical order so that nlwrite can write it: x_0..x_13 continuous and x_14..x_19 integer (nonlinear in
    constraints), x_20..x_27 continuous and x_28, x_29 integer (linear)."""
    n, m = 30, 6
    rng = np.random.Generator(np.random.PCG64(seed))
    is_int = np.zeros(n, bool)
    is_int[14:20] = True
    is_int[28:] = True
    xl = np.where(is_int, 0.0, -3.0)
    xu = np.where(is_int, 5.0, 3.0)
    vtype = np.where(is_int, INTEGER, CONTINUOUS).astype(np.int32)
    x0 = np.where(is_int, rng.integers(1, 4, n).astype(float), rng.uniform(-1, 1, n))
    A = np.zeros((m + 1, n))
    A[0] = -rng.uniform(0.2, 1.0, n)                    # push every variable up against the rows
    A[3, 28] = 1.0
    gl = np.full(m, -INF)
    gu = np.zeros(m)
    lse = [(OP_VAR, 0, 0.0), (OP_CONST, 0, 0.3), (OP_MUL, 0, 0.0), (OP_EXP, 0, 0.0)]
    for i in range(1, 20):
        lse += [(OP_VAR, i, 0.0), (OP_CONST, 0, 0.3), (OP_MUL, 0, 0.0), (OP_EXP, 0, 0.0), (OP_ADD, 0, 0.0)]
    lse.append((OP_LOG, 0, 0.0))
    nrm = [(OP_CONST, 0, 1.0)]
    for i in range(12):
        nrm += [(OP_VAR, i, 0.0), (OP_SQR, 0, 0.0), (OP_ADD, 0, 0.0)]
    nrm.append((OP_SQRT, 0, 0.0))
    ex2 = [(OP_CONST, 0, 0.5), (OP_VAR, 0, 0.0), (OP_MUL, 0, 0.0), (OP_CONST, 0, 0.5), (OP_VAR, 1, 0.0), (OP_MUL, 0, 0.0),
           (OP_ADD, 0, 0.0), (OP_EXP, 0, 0.0)]
    terms = [(1, list(range(20)), lse), (2, list(range(8, 20)), nrm), (3, [12, 13], ex2)]
    g0 = np.array([np.log(np.exp(0.3 * x0[:20]).sum()), np.sqrt(1.0 + (x0[8:20] ** 2).sum()),
                   np.exp(0.5 * x0[12] + 0.5 * x0[13]) + x0[28]])
    gu[:3] = g0 + np.array([0.15, 0.2, 0.3])
    for r in range(3, m):
        li = np.zeros(n)
        cols = rng.choice(n, 6, replace=False)
        li[cols] = rng.uniform(0.2, 1.0, 6)
        A[r + 1] = li
        gu[r] = li @ x0 + rng.uniform(0.5, 1.5)
    return _empty_sparse(StructuredModel(
        n=n, m=m, xl=xl, xu=xu, gl=gl, gu=gu, vtype=vtype, x0=x0, c=np.zeros(m + 1), A=A,
        qd_fn=np.zeros(0, np.int32), qd=np.zeros((0, n * (n + 1) // 2)), terms=terms,
        name=f"WIDE-TERMS-MINLP(seed={seed})"))


def config_c4() -> StructuredModel:
    return syn_cvx_minlp(500, 250, 100, 100, 20261021)


def syn_portfolio(N: int, seed: int) -> StructuredModel:
    """SYN-PORTFOLIO(N, seed) of SURVEY.md Appendix B (C5): variables w (N, [0,1]) then z (N, binary);
    min -mu'w s.t. w'Sigma w <= sigma2_max, 1'w = 1, w_j <= z_j, 1'z <= 10."""
    rng = np.random.Generator(np.random.PCG64(seed))
    F = rng.standard_normal((N, 20)) * 0.1
    d = rng.uniform(0.01, 0.05, N)
    Sigma = F @ F.T + np.diag(d)
    mu = rng.uniform(0.02, 0.15, N)
    top = np.argsort(-mu)[:10]
    weq = np.zeros(N)
    weq[top] = 0.1
    s2 = 1.5 * float(weq @ Sigma @ weq)
    n, m = 2 * N, N + 3
    A = np.zeros((m + 1, n))
    c = np.zeros(m + 1)
    gl = np.full(m, -INF)
    gu = np.zeros(m)
    A[0, :N] = -mu
    Qfull = np.zeros((n, n))
    Qfull[:N, :N] = 2.0 * Sigma          # 1/2 x'Qx = w'Sigma w
    gu[0] = s2
    A[2, :N] = 1.0
    gl[1] = gu[1] = 1.0
    for j in range(N):
        A[3 + j, j] = 1.0
        A[3 + j, N + j] = -1.0
    A[3 + N, N:] = 1.0
    gu[2 + N] = 10.0
    x0 = np.concatenate([weq, (weq > 0).astype(float)])
    return _empty_sparse(StructuredModel(
        n=n, m=m, xl=np.zeros(n), xu=np.ones(n), gl=gl, gu=gu,
        vtype=np.concatenate([np.zeros(N, np.int32), np.full(N, BINARY, np.int32)]), x0=x0, c=c, A=A,
        qd_fn=np.array([1], np.int32), qd=pack_lower(Qfull)[None, :], name=f"SYN-PORTFOLIO(N={N},seed={seed})"))


def config_c5() -> StructuredModel:
    return syn_portfolio(200, 20261022)
