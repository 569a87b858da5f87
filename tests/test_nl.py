"""The .nl reader/tape compiler of the product (host C++, no GPU needed for structure) against the reference's
fixture goldens (test/InterfaceTest.cpp:59-121) and against an independent pure-Python reader."""
import json
import os

import numpy as np
import pytest

from bonmin_b200 import instances as I, nlwrite
from oracle.nl_reader import NlModel
from oracle.struct_eval import Evaluator

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _problem(txt):
    from bonmin_b200.engine import Problem
    return Problem.from_nl(txt)


def test_mytoy_get_methods(bnlp_lib):
    # testGetMethods (InterfaceTest.cpp:59-121): n=4, m=3, col bounds [0,inf),[0,inf),[0,1],[0,5], row uppers 1/4,0,2,
    # types continuous, continuous, binary, integer
    gold = json.load(open(os.path.join(G, "mytoy_golden.json")))
    p = _problem(open(os.path.join(G, "mytoy.nl")).read())
    assert (p.n, p.m, p.nnz_jac, p.nnz_hess) == (gold["n"], gold["m"], gold["nnz_jac"], gold["nnz_hess"])
    info = p.info()
    assert list(info["var_type"]) == gold["var_type"]
    assert list(info["xl"]) == gold["xl"] and list(info["xu"]) == gold["xu"] and list(info["gu"]) == gold["gu"]
    assert (info["gl"] <= -1e19).all()
    assert [list(map(int, e)) for e in zip(info["jac_irow"], info["jac_jcol"])] == gold["jac_structure"]
    assert list(info["con_nonlinear"]) == [1, 0, 0]


def test_python_reader_agrees_on_fixture():
    nm = NlModel(open(os.path.join(G, "mytoy.nl")).read())
    assert list(nm.var_types()) == [0, 0, 1, 2]
    ev = Evaluator(I.mytoy())
    x = np.array([0.3, 0.9, 0.4, 2.0])
    np.testing.assert_allclose(nm.values(x), ev.dense(x)[0], rtol=1e-14)


@pytest.mark.parametrize("name", ["miqcp_n8", "cvx_n10"])
def test_written_samples_structure(bnlp_lib, name):
    s = json.load(open(os.path.join(G, "nl_samples.json")))[name]
    txt = open(os.path.join(G, "nl_samples", name + ".nl")).read()
    p = _problem(txt)
    assert (p.n, p.m, p.nnz_jac, p.nnz_hess) == (s["n"], s["m"], s["nnz_jac"], s["nnz_hess"])
    info = p.info()
    assert [list(map(int, e)) for e in zip(info["jac_irow"], info["jac_jcol"])] == s["jac"]
    np.testing.assert_allclose(info["gu"], s["gu"])
    nm = NlModel(txt)
    for x, v in zip(s["points"], s["values"]):
        np.testing.assert_allclose(nm.values(np.array(x)), v, rtol=1e-12, atol=1e-12)


def test_writer_roundtrip_and_errors(bnlp_lib):
    m = I.syn_cvx_minlp(14, 0, 3, 4, 2)
    p = _problem(nlwrite.write_nl(m))
    ev = Evaluator(m)
    assert (p.nnz_jac, p.nnz_hess) == (len(ev.jac), len(ev.hes))
    from bonmin_b200.engine import BnlpError
    with pytest.raises(BnlpError):
        _problem("b3 0 1 0\n")                      # binary .nl
    with pytest.raises(BnlpError):
        _problem(open(os.path.join(G, "mytoy.nl")).read().replace("o5", "o38", 1))   # unsupported opcode (tan)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["miqcp_n8", "cvx_n10"])
def test_nl_eval_matches_python_reader(bnlp_lib, name):
    from bonmin_b200.engine import Engine
    s = json.load(open(os.path.join(G, "nl_samples.json")))[name]
    txt = open(os.path.join(G, "nl_samples", name + ".nl")).read()
    eng = Engine(_problem(txt))
    X = np.array(s["points"])
    r = eng.eval_batch(X, None, 1.0, want=("f", "g"))
    for b in range(len(X)):
        np.testing.assert_allclose(np.concatenate([[r["f"][b]], r["g"][b]]), s["values"][b], rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
def test_mytoy_nl_solve_goldens(bnlp_lib, oracle_ipm):
    """Config C1 plumbing: the fixture .nl through reader -> tape -> batched IPM gives the reference's known answers
    (InterfaceTest.cpp:122-189)."""
    from bonmin_b200.engine import Engine
    gold = json.load(open(os.path.join(G, "mytoy_golden.json")))
    eng = Engine(_problem(open(os.path.join(G, "mytoy.nl")).read()))
    info = eng.problem.info()
    XL = np.tile(info["xl"], (3, 1))
    XU = np.tile(info["xu"], (3, 1))
    XL[1, 2] = 1.0
    XU[2, 2] = 0.0
    r = eng.solve_batch(XL, XU, info["x0"])
    assert list(r["status"]) == [1, 1, 1]
    assert r["obj"][0] == pytest.approx(gold["root"]["obj"], rel=gold["root"]["rel_tol"])
    np.testing.assert_allclose(r["x"][0][:3], gold["root"]["x"], rtol=1e-6)
    np.testing.assert_allclose(r["duals"][0][8:], gold["root"]["lambda"], atol=1e-6)
    assert r["x"][1][2] == 1.0 and r["x"][2][2] == 0.0
    assert r["obj"][2] == pytest.approx(gold["children"]["x<=0"]["obj"], rel=1e-6)
    o = oracle_ipm.solve_batch(I.mytoy(), XL, XU, info["x0"])
    np.testing.assert_allclose(r["obj"], o["obj"], rtol=1e-6)


def _defvar_values(x):
    v3 = 2 * x[0] + x[1] * x[2]
    return np.array([np.sinh(x[0]) + v3 * v3, np.tanh(v3) + np.log10(x[2]), np.tan(x[0]) + np.cosh(x[1])])


def test_defined_variables_and_lowered_functions(bnlp_lib):
    """V segments (defined variables) are substituted; tanh/tan/sinh/cosh/log10 are lowered to the tape's own operations.
    Structure on the CPU; the independent Python reader evaluates the same file against the closed form."""
    txt = open(os.path.join(G, "nl_samples", "defvar_funcs.nl")).read()
    p = _problem(txt)
    assert (p.n, p.m) == (3, 2)
    info = p.info()
    assert list(info["con_nonlinear"]) == [1, 1]
    assert [list(map(int, e)) for e in zip(info["jac_irow"], info["jac_jcol"])] == [[0, 0], [0, 1], [0, 2], [1, 0], [1, 1]]
    np.testing.assert_allclose(info["xu"], [1.2, 2, 2])
    nm = NlModel(txt)
    for x in ([0.3, 0.7, 1.1], [1.0, 0.2, 0.9]):
        np.testing.assert_allclose(nm.values(np.array(x)), _defvar_values(x), rtol=1e-13)


@pytest.mark.gpu
def test_defined_variables_eval_on_device(bnlp_lib):
    from bonmin_b200.engine import Engine
    txt = open(os.path.join(G, "nl_samples", "defvar_funcs.nl")).read()
    eng = Engine(_problem(txt))
    X = np.array([[0.3, 0.7, 1.1], [1.0, 0.2, 0.9], [0.5, 1.5, 0.6]])
    r = eng.eval_batch(X, want=("f", "g", "gradf", "jac"))
    h = 1e-6
    for b, x in enumerate(X):
        v = _defvar_values(x)
        assert r["f"][b] == pytest.approx(v[0], rel=1e-12)
        np.testing.assert_allclose(r["g"][b], v[1:], rtol=1e-12)
        for j in range(3):                     # gradient of the objective by central differences
            e = np.zeros(3); e[j] = h
            fd = (_defvar_values(x + e)[0] - _defvar_values(x - e)[0]) / (2 * h)
            assert r["gradf"][b, j] == pytest.approx(fd, rel=1e-6, abs=1e-7)
    eng.close()


def test_variable_suffixes_branching_info(bnlp_lib):
    """AmplTMINLP::read_priorities (BonAmplTMINLP.cpp:215-252): priority -> 9999 - value, direction, one pseudo-cost suffix
    serves both ways; constraint / SOS suffixes are skipped.  tests/golden/nl_samples/mytoy_suffixes.nl = the reference's
    mytoy.nl with hand-written S segments (the reference ships no .nl with suffixes)."""
    txt = open(os.path.join(G, "nl_samples", "mytoy_suffixes.nl")).read()
    p = _problem(txt)
    b = p.branching_info()
    assert list(b["priorities"]) == [9998, 9999, 9992, 9994]
    assert list(b["directions"]) == [0, 0, 1, -1]
    assert list(b["up"]) == [0, 0, 0.5, 1.25] and list(b["down"]) == [0, 0, 0.5, 1.25]
    o = NlModel(txt).branching_info()
    assert list(o["priorities"]) == list(b["priorities"]) and list(o["directions"]) == list(b["directions"])
    assert list(o["up"]) == list(b["up"]) and list(o["down"]) == list(b["down"])
    # the model itself is unchanged by the suffixes
    q = _problem(open(os.path.join(G, "mytoy.nl")).read())
    assert (p.n, p.m, p.nnz_jac, p.nnz_hess) == (q.n, q.m, q.nnz_jac, q.nnz_hess)
    assert q.branching_info() == dict(priorities=None, directions=None, up=None, down=None)


def test_sos_suffixes(bnlp_lib):
    """AmplTMINLP::read_sos (BonAmplTMINLP.cpp:254-296): the .sosno / .ref variable suffixes become TMINLP::SosInfo — sets by
    set number, type from its sign, members by reference weight (made strictly increasing, sos_kludge.cpp:29-45), set priority
    from the members' .priority.  The fixture has one set {y1, y2} (sosno 1); a second file adds a type-2 set with equal
    reference weights."""
    txt = open(os.path.join(G, "nl_samples", "mytoy_suffixes.nl")).read()
    p = _problem(txt)
    s = p.sos_info()
    assert list(s["types"]) == [1] and list(s["starts"]) == [0, 2] and list(s["indices"]) == [0, 1]
    assert list(s["priorities"]) == [9998]                     # y1 has priority 1, y2 none
    assert s["weights"][1] > s["weights"][0]
    o = NlModel(txt).sos_info()
    for k in ("types", "priorities", "starts", "indices"):
        assert list(o[k]) == list(s[k])
    np.testing.assert_allclose(o["weights"], s["weights"])
    txt2 = txt.replace("S0 2 sosno\n0 1\n1 1", "S0 4 sosno\n0 1\n1 1\n2 -3\n3 -3\nS4 4 ref\n0 2.5\n1 1.5\n2 4\n3 4")
    s2 = _problem(txt2).sos_info()
    assert list(s2["types"]) == [2, 1] and list(s2["starts"]) == [0, 2, 4]          # set -3 first, then set 1
    assert list(s2["indices"]) == [2, 3, 1, 0]                                        # set 1 ordered by its weights 1.5 < 2.5
    assert s2["weights"][1] == pytest.approx(4.0 + 2e-10, abs=1e-12) and list(s2["weights"][2:]) == [1.5, 2.5]
    assert list(s2["priorities"]) == [9992, 9998]                                      # x has priority 7, z 5; y1 1
    o2 = NlModel(txt2).sos_info()
    for k in ("types", "priorities", "starts", "indices"):
        assert list(o2[k]) == list(s2[k])
    np.testing.assert_allclose(o2["weights"], s2["weights"], rtol=0, atol=1e-15)
    assert _problem(open(os.path.join(G, "mytoy.nl")).read()).sos_info() is None
