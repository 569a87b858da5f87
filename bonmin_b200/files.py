"""ctypes bindings of the on-disk formats either side of the path (bonmin_b200/host/node_files.cpp; include/bnlp.h):
failed-node dump / replay (TMINLP2TNLP::outputDiffs, BoundsReader, StartPointReader), the batch replay file, and the
AMPL .sol writer (AmplTMINLP::finalize_solution / write_solution).  Host code only — lives in libbnlp_host.so."""
from __future__ import annotations

import ctypes

import numpy as np

from .bb import host_lib

c_int, c_dbl, c_void = ctypes.c_int, ctypes.c_double, ctypes.c_void_p
SOLVER_RETURN = {"SUCCESS": 0, "INFEASIBLE": 1, "CONTINUOUS_UNBOUNDED": 2, "LIMIT_EXCEEDED": 3, "USER_INTERRUPT": 4,
                 "MINLP_ERROR": 5}


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void)


def _lib():
    L = host_lib()
    if not getattr(L, "_files_ready", False):
        L.bnlp_write_node_files.argtypes = [ctypes.c_char_p, c_int, c_int] + [c_void] * 6
        L.bnlp_read_bounds_file.argtypes = [ctypes.c_char_p, c_int, c_void, c_void]
        L.bnlp_read_start_file.argtypes = [ctypes.c_char_p, c_void, c_void, c_void, c_int, c_void, c_int]
        L.bnlp_write_batch_replay.argtypes = [ctypes.c_char_p, c_int, c_int, c_int] + [c_void] * 5 + [c_int]
        L.bnlp_read_batch_replay_header.argtypes = [ctypes.c_char_p] + [c_void] * 4
        L.bnlp_read_batch_replay.argtypes = [ctypes.c_char_p, c_int, c_int] + [c_void] * 4
        L.bnlp_sol_status.argtypes = [c_int, c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, c_void, c_void]
        L.bnlp_nl_header_options.argtypes = [ctypes.c_char_p, ctypes.c_size_t, c_void, c_int, c_void]
        L.bnlp_write_sol.argtypes = [ctypes.c_char_p, ctypes.c_char_p, c_void, c_dbl, c_int, c_int, c_void, c_void, c_int]
        L.bnlp_write_plain_sol.argtypes = [ctypes.c_char_p, c_int, c_void]
        L._files_ready = True
    return L


def write_node_files(prob_name, xl, xu, orig_xl, orig_xu, x_init, duals_init=None, m=0):
    xl, xu, oxl, oxu, x0 = map(_d, (xl, xu, orig_xl, orig_xu, x_init))
    du = None if duals_init is None else _d(duals_init)
    if _lib().bnlp_write_node_files(prob_name.encode(), len(xl), m, _p(xl), _p(xu), _p(oxl), _p(oxu), _p(x0), _p(du)):
        raise OSError(f"cannot write {prob_name}.bounds/.start")


def read_bounds_file(path, xl, xu):
    """Returns (xl, xu) = the given bounds with the file's LO/UP lines applied (BoundsReader::readAndApply)."""
    xl, xu = _d(xl).copy(), _d(xu).copy()
    if _lib().bnlp_read_bounds_file(path.encode(), len(xl), _p(xl), _p(xu)) < 0:
        raise ValueError(f"malformed bounds file {path}")
    return xl, xu


def read_start_file(path):
    np_, nd = c_int(), c_int()
    L = _lib()
    if L.bnlp_read_start_file(path.encode(), ctypes.byref(np_), ctypes.byref(nd), None, 0, None, 0):
        raise ValueError(f"malformed start file {path}")
    x, du = np.zeros(np_.value), np.zeros(nd.value)
    L.bnlp_read_start_file(path.encode(), ctypes.byref(np_), ctypes.byref(nd), _p(x), len(x), _p(du), len(du))
    return x, (du if nd.value else None)


def write_batch_replay(prob_name, XL, XU, orig_xl, orig_xu, X0, m=0):
    XL, XU, oxl, oxu, X0 = map(_d, (np.atleast_2d(XL), np.atleast_2d(XU), orig_xl, orig_xu, X0))
    B, n = XL.shape
    if _lib().bnlp_write_batch_replay(prob_name.encode(), n, m, B, _p(XL), _p(XU), _p(oxl), _p(oxu), _p(X0), int(X0.ndim == 2)):
        raise OSError(f"cannot write {prob_name}.children/.start")


def read_batch_replay(prob_name, orig_xl, orig_xu):
    """-> (XL, XU, X0) of the dumped batch (X0: n for a shared start point, B x n per child)."""
    L = _lib()
    n, m, B, per = c_int(), c_int(), c_int(), c_int()
    f = (prob_name + ".children").encode()
    if L.bnlp_read_batch_replay_header(f, ctypes.byref(n), ctypes.byref(m), ctypes.byref(B), ctypes.byref(per)):
        raise ValueError(f"malformed replay file {prob_name}.children")
    oxl, oxu = _d(orig_xl), _d(orig_xu)
    if len(oxl) != n.value:
        raise ValueError("replay file is for another model")
    XL, XU = np.zeros((B.value, n.value)), np.zeros((B.value, n.value))
    if L.bnlp_read_batch_replay(f, n.value, B.value, _p(oxl), _p(oxu), _p(XL), _p(XU)):
        raise ValueError(f"malformed replay file {prob_name}.children")
    x0, _ = read_start_file(prob_name + ".start")
    return XL, XU, (x0.reshape(B.value, n.value) if per.value else x0)


def sol_status(solver_return, has_x=True, app_name="bonmin"):
    """AmplTMINLP::finalize_solution: -> (message, solve_result_num, finished)."""
    code = SOLVER_RETURN[solver_return] if isinstance(solver_return, str) else int(solver_return)
    buf = ctypes.create_string_buffer(256)
    num, fin = c_int(), c_int()
    if _lib().bnlp_sol_status(code, int(has_x), app_name.encode(), buf, len(buf), ctypes.byref(num), ctypes.byref(fin)):
        raise ValueError(f"unknown TMINLP::SolverReturn {solver_return}")
    return buf.value.decode(), num.value, fin.value


def nl_header_options(nl_text):
    data = nl_text if isinstance(nl_text, bytes) else nl_text.encode()
    opts = (c_int * 10)()
    vb = c_dbl()
    cnt = _lib().bnlp_nl_header_options(data, len(data), opts, 10, ctypes.byref(vb))
    if cnt < 0:
        raise ValueError("not a text/binary .nl header")
    return list(opts[:cnt + 1]), vb.value


def write_sol(path, message, ampl_options, n_con, x=None, y=None, solve_result_num=0, n_var=None, vbtol=0.0):
    opts = (c_int * max(1, len(ampl_options)))(*ampl_options)
    xa = None if x is None else _d(x)
    ya = None if y is None else _d(y)
    nv = len(xa) if xa is not None else int(n_var or 0)
    if _lib().bnlp_write_sol(path.encode(), message.encode(), opts, float(vbtol), n_con, nv, _p(xa), _p(ya), solve_result_num):
        raise OSError(f"cannot write {path}")


def write_plain_sol(path, x):
    xa = _d(x)
    if _lib().bnlp_write_plain_sol(path.encode(), len(xa), _p(xa)):
        raise OSError(f"cannot write {path}")
