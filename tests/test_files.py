"""On-disk formats either side of the path (bonmin_b200/host/node_files.cpp): failed-node dump and its readers
(BonTMINLP2TNLP.cpp:551-615, BonBoundsReader.cpp:52-119, BonStartPointReader.cpp:15-37), the batch replay file, and the
.sol writer (BonAmplTMINLP.cpp:712-794).  Host code: runs without a GPU."""
import os

import numpy as np
import pytest

from bonmin_b200 import files, instances as I

HERE = os.path.dirname(os.path.abspath(__file__))


def test_node_dump_round_trip(tmp_path):
    m = I.syn_miqcp(12, 6, 3, 3, 1)
    xl, xu = np.array(m.xl, float), np.array(m.xu, float)
    xl[2] = 1.0; xu[2] = 1.0; xu[5] = 0.0; xl[7] = xl[7] + 1e-13       # the last one is "equal" for CoinRelFltEq
    x0 = np.linspace(0.1, 1.3, 12) / 3.0
    du = np.arange(2 * 12 + m.m, dtype=float) * 0.1 - 1
    name = str(tmp_path / "node")
    files.write_node_files(name, xl, xu, m.xl, m.xu, x0, du, m=m.m)
    lines = open(name + ".bounds").read().split("\n")
    assert lines[:3] == ["LO\t2\t1", "UP\t2\t1", "UP\t5\t0"] and lines[3] == ""
    head = open(name + ".start").readline()
    assert head == f"12\t{2 * 12 + m.m}\n"
    rl, ru = files.read_bounds_file(name + ".bounds", m.xl, m.xu)
    assert np.array_equal(rl, np.where(np.arange(12) == 7, m.xl, xl)) and np.array_equal(ru, xu)
    rx, rd = files.read_start_file(name + ".start")
    assert np.array_equal(rx, x0) and np.array_equal(rd, du)           # 17 digits: exact
    files.write_node_files(name, xl, xu, m.xl, m.xu, x0, None, m=m.m)   # no duals: the dump stops after the primals
    rx, rd = files.read_start_file(name + ".start")
    assert np.array_equal(rx, x0) and rd is None
    open(name + ".bounds", "w").write("LO\t1\t0\nXX\t2\t1\n")
    with pytest.raises(ValueError):
        files.read_bounds_file(name + ".bounds", m.xl, m.xu)


def test_batch_replay_round_trip(tmp_path):
    m = I.syn_miqcp(20, 10, 4, 4, 2)
    rng = np.random.default_rng(0)
    xs = np.clip(rng.uniform(-1, 3, 20), m.xl, m.xu)
    XL, XU = I.c3_children(m, xs, 37)
    XL[5, 3] = 0.1234567890123456789
    name = str(tmp_path / "batch")
    files.write_batch_replay(name, XL, XU, m.xl, m.xu, xs, m=m.m)
    rl, ru, x0 = files.read_batch_replay(name, m.xl, m.xu)
    assert np.array_equal(rl, XL) and np.array_equal(ru, XU) and np.array_equal(x0, xs)
    X0 = np.tile(xs, (37, 1)) + rng.standard_normal((37, 20))
    files.write_batch_replay(name, XL, XU, m.xl, m.xu, X0, m=m.m)
    rl, ru, x0 = files.read_batch_replay(name, m.xl, m.xu)
    assert x0.shape == (37, 20) and np.array_equal(x0, X0)
    with pytest.raises(ValueError):
        files.read_batch_replay(name, m.xl[:-1], m.xu[:-1])


def test_sol_status_table():
    # finalize_solution's table (BonAmplTMINLP.cpp:716-753)
    assert files.sol_status("SUCCESS") == ("\nbonmin: Optimal", 3, 1)
    assert files.sol_status("INFEASIBLE", has_x=False) == ("\nbonmin: Infeasible problem", 220, 1)
    assert files.sol_status("CONTINUOUS_UNBOUNDED")[1:] == (300, 1)
    assert files.sol_status("LIMIT_EXCEEDED", has_x=True)[1:] == (421, 0)
    assert files.sol_status("LIMIT_EXCEEDED", has_x=False)[1:] == (410, 0)
    assert files.sol_status("USER_INTERRUPT", has_x=True)[1:] == (422, 0)
    assert files.sol_status("USER_INTERRUPT", has_x=False)[1:] == (411, 0)
    assert files.sol_status("MINLP_ERROR") == ("\nbonmin: Error encountered in optimization.", 500, -1)
    with pytest.raises(ValueError):
        files.sol_status(9)


def test_sol_file_for_mytoy(tmp_path):
    nl = open(os.path.join(HERE, "golden", "mytoy.nl")).read()
    opts, vb = files.nl_header_options(nl)
    assert opts == [3, 0, 1, 0] and vb == 0.0                          # test/mytoy.nl line 1: "g3 0 1 0"
    msg, num, _ = files.sol_status("SUCCESS")
    x = np.array([1.0, 1.0, 0.5, 0.0])
    path = str(tmp_path / "mytoy.sol")
    files.write_sol(path, msg, opts, 3, x=x, solve_result_num=num)
    assert open(path).read() == "\nbonmin: Optimal\n\nOptions\n3\n0\n1\n0\n3\n0\n4\n4\n1\n1\n0.5\n0\nobjno 0 3\n"
    files.write_sol(path, files.sol_status("INFEASIBLE", False)[0], opts, 3, x=None, n_var=4, solve_result_num=220)
    assert open(path).read().endswith("Options\n3\n0\n1\n0\n3\n0\n4\n0\nobjno 0 220\n")
    # values are written as the shortest decimals that read back exactly
    v = np.array([0.1, 1 / 3, -2.5e-300, 123456789.125])
    files.write_sol(path, msg, opts, 0, x=v, y=None, solve_result_num=3)
    body = open(path).read().split("\n")
    back = np.array([float(s) for s in body[body.index("Options") + 9:-2]])
    assert np.array_equal(back, v) and body[body.index("Options") + 9] == "0.1"
    files.write_plain_sol(path, x)
    assert open(path).read() == "0\t1\n1\t1\n2\t0.5\n3\t0\n-1\n"
