"""ctypes binding of the C ABI (include/bnlp.h).  This is plumbing: every number is produced by
libbnlp.so's CUDA kernels.  There is no CPU fallback — a missing library or GPU raises."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libbnlp.so")
if os.environ.get("BNLP_LIB"):      # development aid: A/B runs of two builds of the same source on one GPU box (scripts/dev_variants.py)
    _LIBPATH = os.path.abspath(os.environ["BNLP_LIB"])
_lib = None

c_int, c_dbl, c_void = ctypes.c_int, ctypes.c_double, ctypes.c_void_p
P_int, P_dbl = ctypes.POINTER(c_int), ctypes.POINTER(c_dbl)


class BnlpError(RuntimeError):
    pass


class TapeDesc(ctypes.Structure):
    _fields_ = [("n", c_int), ("m", c_int), ("xl", P_dbl), ("xu", P_dbl), ("gl", P_dbl), ("gu", P_dbl),
                ("var_type", P_int), ("x0", P_dbl), ("c", P_dbl), ("lin_start", P_int), ("lin_col", P_int),
                ("lin_val", P_dbl), ("n_qdense", c_int), ("qd_fn", P_int), ("qd", P_dbl), ("n_qsparse", c_int),
                ("qs_fn", P_int), ("qs_i", P_int), ("qs_j", P_int), ("qs_v", P_dbl), ("n_terms", c_int),
                ("term_fn", P_int), ("term_var_start", P_int), ("term_var", P_int), ("term_op_start", P_int),
                ("term_op", P_int), ("term_iarg", P_int), ("term_darg", P_dbl)]


class Options(ctypes.Structure):
    _fields_ = [("tol", c_dbl), ("max_iter", c_int), ("ctas_per_device", c_int)]


class FpDesc(ctypes.Structure):
    _fields_ = [("n_inds", ctypes.c_int), ("inds", ctypes.POINTER(ctypes.c_int)), ("norm", ctypes.c_int),
                ("lambda_", ctypes.c_double), ("sigma", ctypes.c_double), ("objective_scaling_factor", ctypes.c_double),
                ("use_fp_objective", ctypes.c_int), ("use_cutoff_constraint", ctypes.c_int),
                ("use_local_branching_constraint", ctypes.c_int), ("cutoff", ctypes.c_double),
                ("rhs_local_branching", ctypes.c_double)]


class LpDims(ctypes.Structure):
    _fields_ = [("n_cols", ctypes.c_int), ("n_rows", ctypes.c_int), ("nnz", ctypes.c_int), ("has_alpha", ctypes.c_int),
                ("obj_offset", ctypes.c_double)]


class OaTols(ctypes.Structure):
    _fields_ = [("tiny", c_dbl), ("very_tiny", c_dbl), ("rhs_relax", c_dbl), ("infty", c_dbl),
                ("skip_non_violated", c_int)]


EXPORTS = ["bnlp_last_error", "bnlp_problem_from_nl", "bnlp_problem_from_tape", "bnlp_problem_free",
           "bnlp_problem_sizes", "bnlp_problem_info", "bnlp_engine_create", "bnlp_engine_free", "bnlp_eval_batch",
           "bnlp_solve_batch", "bnlp_solve_batch_warm", "bnlp_solve_batch_device", "bnlp_oa_capacity", "bnlp_oa_x2_width", "bnlp_oa_batch", "bnlp_engine_stats",
           "bnlp_engine_launches", "bnlp_measure_fp64_peak", "bnlp_bb_run_cb", "bnlp_bb_run_cb2", "bnlp_bb_run_engine", "bnlp_dive_run_cb", "bnlp_dive_run_engine", "bnlp_violation_batch",
           "bnlp_benders_batch", "bnlp_branching_tqp", "bnlp_tqp_solve_batch",
           "bnlp_extract_linear_relaxation", "bnlp_constraint_oa_batch", "bnlp_engine_sizes",
           "bnlp_resolve_cb", "bnlp_resolve_engine", "bnlp_problem_fp", "bnlp_solve_children", "bnlp_bb_run_cb3",
           "bnlp_write_node_files", "bnlp_read_bounds_file", "bnlp_read_start_file", "bnlp_write_batch_replay",
           "bnlp_read_batch_replay_header", "bnlp_read_batch_replay", "bnlp_sol_status", "bnlp_nl_header_options",
           "bnlp_write_sol", "bnlp_write_plain_sol", "bnlp_problem_cut_strengthening", "bnlp_problem_branching_info", "bnlp_problem_sos_info", "bnlp_host_register", "bnlp_host_unregister", "bnlp_host_selftest", "bnlp_linearizer_engine",
           "bnlp_comm_unique_id", "bnlp_comm_create", "bnlp_comm_free", "bnlp_comm_stats", "bnlp_comm_allreduce_min",
           "bnlp_solve_batch_sharded", "bnlp_bb_run_engine_sharded", "bnlp_oa_run_cb", "bnlp_oa_run_engine", "bnlp_oa_feas_check_cb", "bnlp_oa_feas_check_engine", "bnlp_oa_nlp_optim_cb", "bnlp_oa_nlp_optim_engine",
           "bnlp_ecp_rounds_cb", "bnlp_ecp_rounds_engine"]


def pin_array(a):
    """Page-lock a numpy array for its lifetime (bnlp_host_register now, bnlp_host_unregister when it is collected)."""
    import weakref
    ptr = a.ctypes.data
    if a.nbytes and lib().bnlp_host_register(c_void(ptr), a.nbytes) == 0:
        weakref.finalize(a, lambda p=ptr: lib().bnlp_host_unregister(c_void(p)))
        return True
    return False


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            raise BnlpError(f"{_LIBPATH} is missing: run `python -m bonmin_b200.build` (no CPU fallback exists)")
        L = ctypes.CDLL(_LIBPATH)
        L.bnlp_last_error.restype = ctypes.c_char_p
        L.bnlp_problem_from_tape.restype = c_void
        L.bnlp_problem_from_tape.argtypes = [ctypes.POINTER(TapeDesc), ctypes.c_char_p, ctypes.c_size_t]
        L.bnlp_problem_from_nl.restype = c_void
        L.bnlp_problem_from_nl.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
        L.bnlp_problem_fp.restype = c_void
        L.bnlp_problem_fp.argtypes = [c_void, ctypes.POINTER(FpDesc), ctypes.c_char_p, ctypes.c_size_t]
        L.bnlp_problem_cut_strengthening.restype = c_void
        L.bnlp_problem_cut_strengthening.argtypes = [c_void, c_int, ctypes.c_char_p, ctypes.c_size_t]
        L.bnlp_problem_branching_info.argtypes = [c_void] * 5
        L.bnlp_host_register.argtypes = [c_void, ctypes.c_size_t]
        L.bnlp_host_unregister.argtypes = [c_void]
        L.bnlp_problem_free.argtypes = [c_void]
        L.bnlp_problem_sizes.argtypes = [c_void, P_int, P_int, P_int, P_int]
        L.bnlp_problem_info.argtypes = [c_void] + [c_void] * 11
        L.bnlp_engine_create.restype = c_void
        L.bnlp_engine_create.argtypes = [c_void, P_int, c_int, ctypes.POINTER(Options)]
        L.bnlp_engine_free.argtypes = [c_void]
        L.bnlp_eval_batch.argtypes = [c_void, c_int, c_void, c_void, c_dbl, c_void, c_void, c_void, c_void, c_void]
        L.bnlp_solve_batch.argtypes = [c_void, c_int, c_void, c_void, c_void, c_int] + [c_void] * 6
        L.bnlp_solve_batch_warm.argtypes = [c_void, c_int, c_void, c_void, c_void, c_int, c_void, c_int] + [c_void] * 6
        L.bnlp_solve_children.argtypes = [c_void, c_int, c_void, c_void, c_int] + [c_void] * 5 + [c_int] + [c_void] * 6
        L.bnlp_solve_batch_device.argtypes = [c_void, c_int, c_int, c_void, c_void, c_void, c_int] + [c_void] * 7
        L.bnlp_oa_capacity.argtypes = [c_void, c_int, P_int, ctypes.POINTER(ctypes.c_size_t)]
        L.bnlp_oa_batch.argtypes = [c_void, c_int, c_void, c_void, c_void, c_void, c_void, c_dbl, c_int,
                                    ctypes.POINTER(OaTols), c_void, c_void, c_void, c_void, c_void, c_void]
        L.bnlp_violation_batch.argtypes = [c_void, c_int, c_void, c_void, c_void, c_void]
        L.bnlp_benders_batch.argtypes = [c_void, c_int, c_void, c_void, c_void, c_void, c_int, ctypes.POINTER(OaTols),
                                         c_void, c_void, c_void, c_void]
        L.bnlp_branching_tqp.restype = c_void
        L.bnlp_branching_tqp.argtypes = [c_void, c_void, c_void, P_dbl, ctypes.c_char_p, ctypes.c_size_t]
        L.bnlp_tqp_solve_batch.argtypes = [c_void, c_void, c_void, c_int] + [c_void] * 6
        L.bnlp_extract_linear_relaxation.argtypes = [c_void, c_void, c_void, c_void, c_void, c_int, ctypes.POINTER(OaTols),
                                                     ctypes.POINTER(LpDims)] + [c_void] * 10
        L.bnlp_constraint_oa_batch.argtypes = [c_void, c_int, c_int, c_void, c_void, c_void, c_void, ctypes.POINTER(OaTols)] + \
            [c_void] * 5
        L.bnlp_engine_stats.argtypes = [c_void, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong),
                                        P_dbl]
        L.bnlp_engine_launches.restype = ctypes.c_longlong
        L.bnlp_engine_launches.argtypes = [c_void]
        L.bnlp_measure_fp64_peak.restype = c_dbl
        L.bnlp_measure_fp64_peak.argtypes = [c_void, c_int]
        _lib = L
    return _lib


def _err():
    return lib().bnlp_last_error().decode()


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _pd(a):
    return a.ctypes.data_as(P_dbl)


def _pi(a):
    return a.ctypes.data_as(P_int)


def _vp(a):
    return None if a is None else a.ctypes.data_as(c_void)


class Problem:
    """A compiled model (`bnlp_problem`)."""

    def __init__(self, handle):
        if not handle:
            raise BnlpError(_err())
        self.h = handle
        n, m, nj, nh = c_int(), c_int(), c_int(), c_int()
        lib().bnlp_problem_sizes(self.h, n, m, nj, nh)
        self.n, self.m, self.nnz_jac, self.nnz_hess = n.value, m.value, nj.value, nh.value

    @classmethod
    def from_model(cls, model):
        """StructuredModel (bonmin_b200.instances) -> bnlp_problem_from_tape."""
        n, m = int(model.n), int(model.m)
        A = _d(model.A)
        keep = []  # arrays must stay alive during the call
        lin_start, lin_col, lin_val = [0], [], []
        for k in range(m + 1):
            nzc = np.flatnonzero(A[k])
            lin_col.extend(nzc.tolist())
            lin_val.extend(A[k, nzc].tolist())
            lin_start.append(len(lin_col))
        terms = model.get("terms", [])
        tv_start, tv, to_start, top, tia, tda, tfn = [0], [], [0], [], [], [], []
        for fn, vars_, prog in terms:
            tfn.append(fn)
            tv.extend(vars_)
            tv_start.append(len(tv))
            for op, ia, da in prog:
                top.append(op)
                tia.append(ia)
                tda.append(da)
            to_start.append(len(top))
        arrs = dict(xl=_d(model.xl), xu=_d(model.xu), gl=_d(model.gl), gu=_d(model.gu), vt=_i(model.vtype),
                    x0=_d(model.x0), c=_d(model.c), ls=_i(lin_start), lc=_i(lin_col), lv=_d(lin_val),
                    qdf=_i(model.qd_fn), qd=_d(model.qd), qsf=_i(model.qs_fn), qsi=_i(model.qs_i),
                    qsj=_i(model.qs_j), qsv=_d(model.qs_v), tfn=_i(tfn), tvs=_i(tv_start), tv=_i(tv),
                    tos=_i(to_start), top=_i(top), tia=_i(tia), tda=_d(tda))
        d = TapeDesc(n, m, _pd(arrs["xl"]), _pd(arrs["xu"]), _pd(arrs["gl"]), _pd(arrs["gu"]), _pi(arrs["vt"]),
                     _pd(arrs["x0"]), _pd(arrs["c"]), _pi(arrs["ls"]), _pi(arrs["lc"]), _pd(arrs["lv"]),
                     len(arrs["qdf"]), _pi(arrs["qdf"]), _pd(arrs["qd"]), len(arrs["qsf"]), _pi(arrs["qsf"]),
                     _pi(arrs["qsi"]), _pi(arrs["qsj"]), _pd(arrs["qsv"]), len(tfn), _pi(arrs["tfn"]),
                     _pi(arrs["tvs"]), _pi(arrs["tv"]), _pi(arrs["tos"]), _pi(arrs["top"]), _pi(arrs["tia"]),
                     _pd(arrs["tda"]))
        err = ctypes.create_string_buffer(512)
        h = lib().bnlp_problem_from_tape(ctypes.byref(d), err, 512)
        if not h:
            raise BnlpError(err.value.decode())
        del keep
        return cls(h)

    @classmethod
    def from_nl(cls, text):
        if isinstance(text, str):
            text = text.encode()
        err = ctypes.create_string_buffer(512)
        h = lib().bnlp_problem_from_nl(text, len(text), err, 512)
        if not h:
            raise BnlpError(err.value.decode())
        return cls(h)

    def fp_variant(self, inds, norm=2, lam=1.0, sigma=1.0, objective_scaling_factor=1.0, use_fp_objective=True,
                   cutoff=None, rhs_local_branching=None):
        """TNLP2FPNLP as a new Problem with len(inds) extra target variables (see include/bnlp.h: bnlp_problem_fp)."""
        ia = _i(inds)
        d = FpDesc(len(ia), _pi(ia), int(norm), float(lam), float(sigma), float(objective_scaling_factor),
                   int(use_fp_objective), int(cutoff is not None), int(rhs_local_branching is not None),
                   float(cutoff or 0.0), float(rhs_local_branching or 0.0))
        err = ctypes.create_string_buffer(512)
        h = lib().bnlp_problem_fp(self.h, ctypes.byref(d), err, 512)
        if not h:
            raise BnlpError(err.value.decode())
        return Problem(h)

    def branching_info(self):
        """TMINLP::branchingInfo(): dict(priorities, directions, up, down), None for what the model does not have."""
        n = self.n
        pri, dr = np.zeros(n, np.int32), np.zeros(n, np.int32)
        up, dn = np.zeros(n), np.zeros(n)
        mask = lib().bnlp_problem_branching_info(self.h, _vp(pri), _vp(dr), _vp(up), _vp(dn))
        if mask < 0:
            raise BnlpError("bnlp_problem_branching_info failed")
        return dict(priorities=pri if mask & 1 else None, directions=dr if mask & 2 else None,
                    up=up if mask & 4 else None, down=dn if mask & 8 else None)

    def sos_info(self):
        """TMINLP::sosConstraints(): None, or dict(types, priorities, starts, indices, weights) from the .sosno / .ref suffixes."""
        L = lib()
        L.bnlp_problem_sos_info.argtypes = [c_void] * 8
        num, nz = c_int(), c_int()
        if L.bnlp_problem_sos_info(self.h, ctypes.byref(num), ctypes.byref(nz), None, None, None, None, None):
            raise BnlpError("bnlp_problem_sos_info failed")
        if num.value == 0:
            return None
        types = np.zeros(num.value, np.int8)
        pri, starts = np.zeros(num.value, np.int32), np.zeros(num.value + 1, np.int32)
        idx, wts = np.zeros(nz.value, np.int32), np.zeros(nz.value)
        L.bnlp_problem_sos_info(self.h, ctypes.byref(num), ctypes.byref(nz), _vp(types), _vp(pri), _vp(starts), _vp(idx), _vp(wts))
        return dict(types=types, priorities=pri, starts=starts, indices=idx, weights=wts)

    def cs_variant(self, constr_index):
        """CutStrengthener::StrengtheningTNLP for row `constr_index` (-1: the objective cut) as a new Problem whose last
        nv variables are the cut coefficients (see include/bnlp.h: bnlp_problem_cut_strengthening)."""
        err = ctypes.create_string_buffer(512)
        h = lib().bnlp_problem_cut_strengthening(self.h, int(constr_index), err, 512)
        if not h:
            raise BnlpError(err.value.decode())
        return Problem(h)

    def info(self):
        n, m = self.n, self.m
        out = dict(jac_irow=np.zeros(self.nnz_jac, np.int32), jac_jcol=np.zeros(self.nnz_jac, np.int32),
                   hes_irow=np.zeros(self.nnz_hess, np.int32), hes_jcol=np.zeros(self.nnz_hess, np.int32),
                   var_type=np.zeros(n, np.int32), con_nonlinear=np.zeros(m, np.int32), xl=np.zeros(n),
                   xu=np.zeros(n), gl=np.zeros(m), gu=np.zeros(m), x0=np.zeros(n))
        lib().bnlp_problem_info(self.h, *[_vp(out[k]) for k in
                                          ("jac_irow", "jac_jcol", "hes_irow", "hes_jcol", "var_type",
                                           "con_nonlinear", "xl", "xu", "gl", "gu", "x0")])
        return out

    def __del__(self):
        try:
            if self.h:
                lib().bnlp_problem_free(self.h)
                self.h = None
        except Exception:
            pass


class Engine:
    """`bnlp_engine`: the tape replicated on the listed devices."""

    def __init__(self, problem: Problem, devices=(0,), tol=0.0, max_iter=0, ctas_per_device=0):
        self.problem = problem
        ids = _i(list(devices))
        opt = Options(tol, max_iter, ctas_per_device)
        self.h = lib().bnlp_engine_create(problem.h, _pi(ids), len(ids), ctypes.byref(opt))
        if not self.h:
            raise BnlpError(_err())
        self.n, self.m = problem.n, problem.m

    def _ck(self, rc):
        if rc != 0:
            raise BnlpError(f"engine error {rc}: {_err()}")

    def solve_batch(self, XL, XU, X0, want_x=True, want_g=True, want_duals=True, out=None, pin=False, dual0=None):
        """B x TNLPSolver::OptimizeTNLP with host buffers (copies inside the call).  `out`: the dict of an earlier call
        with the same B and outputs, whose arrays are reused; `pin`: page-lock freshly allocated result arrays (for a
        caller that keeps them with `out=`).  dual0: multipliers [z_L | z_U | lambda] of a warm start (2n+m values for
        every child, or B rows): B x ReOptimizeTNLP after setWarmStart (bnlp_solve_batch_warm)."""
        n, m = self.n, self.m
        XL = _d(np.atleast_2d(XL))
        XU = _d(np.atleast_2d(XU))
        B = XL.shape[0]
        X0 = _d(X0)
        per_child = int(X0.ndim == 2)
        if (out is not None and len(out["obj"]) == B and (out["x"] is not None) == want_x and (out["g"] is not None) == want_g
                and (out["duals"] is not None) == want_duals):
            OBJ, ST, IT, X, G, D = (out[k] for k in ("obj", "status", "iters", "x", "g", "duals"))
        else:
            OBJ = np.zeros(B)
            ST = np.zeros(B, np.int32)
            IT = np.zeros(B, np.int32)
            X = np.zeros((B, n)) if want_x else None
            G = np.zeros((B, m)) if want_g else None
            D = np.zeros((B, 2 * n + m)) if want_duals else None
            if pin:
                for a in (OBJ, ST, IT, X, G, D):
                    if a is not None:
                        pin_array(a)
        D0 = None
        if dual0 is not None:
            D0 = _d(dual0)
            if D0.shape[-1] != 2 * n + m or (D0.ndim == 2 and D0.shape[0] != B):
                raise ValueError(f"solve_batch: dual0 must be (2n+m,) or ({B}, 2n+m), got {D0.shape}")
        self._ck(lib().bnlp_solve_batch_warm(self.h, B, _vp(XL), _vp(XU), _vp(X0), per_child, _vp(D0),
                                             int(D0 is not None and D0.ndim == 2), _vp(OBJ), _vp(ST), _vp(IT),
                                             _vp(X), _vp(G), _vp(D)))
        return dict(obj=OBJ, status=ST, iters=IT, x=X, g=G, duals=D)

    def solve_children(self, parent_XL, parent_XU, child_parent, child_col, child_way, child_bound, X0, want_x=True,
                       want_g=False, want_duals=False):
        """The strong-branching form of solve_batch (bnlp_solve_children): child b = parent child_parent[b] with the upper
        (way 0) or lower (way 1) bound of column child_col[b] replaced by child_bound[b]; bound vectors are expanded on
        the device.  X0: n (shared) or P x n (per parent)."""
        n, m = self.n, self.m
        PXL = _d(np.atleast_2d(parent_XL))
        PXU = _d(np.atleast_2d(parent_XU))
        P = PXL.shape[0]
        cp = np.ascontiguousarray(child_parent, np.int32)
        cc = np.ascontiguousarray(child_col, np.int32)
        cw = np.ascontiguousarray(child_way, np.int32)
        cb = _d(child_bound)
        B = len(cp)
        X0 = _d(X0)
        per_parent = int(X0.ndim == 2)
        OBJ = np.zeros(B)
        ST = np.zeros(B, np.int32)
        IT = np.zeros(B, np.int32)
        X = np.zeros((B, n)) if want_x else None
        G = np.zeros((B, m)) if want_g else None
        D = np.zeros((B, 2 * n + m)) if want_duals else None
        self._ck(lib().bnlp_solve_children(self.h, P, _vp(PXL), _vp(PXU), B, _vp(cp), _vp(cc), _vp(cw), _vp(cb), _vp(X0),
                                           per_parent, _vp(OBJ), _vp(ST), _vp(IT), _vp(X), _vp(G), _vp(D)))
        return dict(obj=OBJ, status=ST, iters=IT, x=X, g=G, duals=D)

    def solve_batch_device(self, dXL, dXU, dX0, dOBJ, dST, dIT, dX=None, dG=None, dD=None, x0_per_child=False,
                           stream=0, dev_index=0):
        """Device-pointer variant (torch tensors or raw ints); asynchronous on `stream`."""
        def ptr(t):
            if t is None:
                return None
            return c_void(t if isinstance(t, int) else t.data_ptr())
        B = dOBJ.shape[0]
        self._ck(lib().bnlp_solve_batch_device(self.h, dev_index, B, ptr(dXL), ptr(dXU), ptr(dX0), int(x0_per_child),
                                               ptr(dOBJ), ptr(dST), ptr(dIT), ptr(dX), ptr(dG), ptr(dD),
                                               c_void(stream)))

    def eval_batch(self, X, LAM=None, sigma=1.0, want=("f", "g", "gradf", "jac", "hess")):
        n, m = self.n, self.m
        X = _d(np.atleast_2d(X))
        B = X.shape[0]
        LAMa = None if LAM is None else _d(np.atleast_2d(LAM))
        F = np.zeros(B) if "f" in want else None
        G = np.zeros((B, m)) if "g" in want else None
        GF = np.zeros((B, n)) if "gradf" in want else None
        J = np.zeros((B, self.problem.nnz_jac)) if "jac" in want else None
        H = np.zeros((B, self.problem.nnz_hess)) if "hess" in want else None
        self._ck(lib().bnlp_eval_batch(self.h, B, _vp(X), _vp(LAMa), float(sigma), _vp(F), _vp(G), _vp(GF), _vp(J),
                                       _vp(H)))
        return dict(f=F, g=G, gradf=GF, jac=J, hess=H)

    def oa_batch(self, X, DUALS=None, XL=None, XU=None, X2=None, theta=0.0, get_obj=0, tols=None, out=None, pin=False):
        """B x OsiTMINLPInterface::getOuterApproximation -> CSR rows (see include/bnlp.h).  `out`: a dict returned by an
        earlier call with the same B, whose arrays are reused (a caller in a loop keeps its cut buffers: allocating and
        first-touching ~300 MB per call costs more than the whole device side).  `pin`: page-lock freshly allocated cut
        buffers (bnlp_host_register; released when the arrays are garbage-collected) — for buffers that are reused."""
        n = self.n
        X = _d(np.atleast_2d(X))
        B = X.shape[0]
        rpp = c_int()
        cap = ctypes.c_size_t()
        self._ck(lib().bnlp_oa_capacity(self.h, B, ctypes.byref(rpp), ctypes.byref(cap)))
        nrows = B * rpp.value
        if out is not None and len(out["src_row"]) == nrows and len(out["val"]) == cap.value:
            row_start, col, val, LB, UB, src = (out[k] for k in ("row_start", "col", "val", "lb", "ub", "src_row"))
        else:
            row_start = np.zeros(nrows + 1, np.int32)
            col = np.zeros(cap.value, np.int32)
            val = np.zeros(cap.value)
            LB = np.zeros(nrows)
            UB = np.zeros(nrows)
            src = np.zeros(nrows, np.int32)
            if pin:
                for a in (col, val):
                    pin_array(a)
        if XL is None or XU is None:
            info = self.problem.info()
        XL = _d(info["xl"] if XL is None else XL)
        XU = _d(info["xu"] if XU is None else XU)
        Du = None
        if DUALS is not None:
            Du = np.atleast_2d(np.asarray(DUALS, float))
            if Du.shape[0] == 1 and B > 1:       # one multiplier vector for every point (the reference's getRowPrice())
                Du = np.repeat(Du, B, axis=0)
            if Du.shape != (B, 2 * n + self.m):
                raise ValueError(f"oa_batch: DUALS must be {B} x {2 * n + self.m}, got {Du.shape}")
            Du = _d(Du)
        X2a = None
        if X2 is not None:
            # the width of an X2 row depends on whether the objective cut is generated (it is tested against the LP's alpha
            # column): a short row would be over-read by the library
            w = lib().bnlp_oa_x2_width(self.h, int(get_obj))
            X2a = np.atleast_2d(np.asarray(X2, float))
            if X2a.shape[0] != B or X2a.shape[1] < w:
                raise ValueError(f"oa_batch: X2 must be {B} x {w} for get_obj={get_obj} on this model "
                                 f"(n = {n}; the objective cut needs the LP's alpha column), got {X2a.shape}")
            X2a = _d(X2a[:, :w])
        t = tols or OaTols(1e-8, 1e-17, 1e-8, 1e19, 0)
        self._ck(lib().bnlp_oa_batch(self.h, B, _vp(X), _vp(Du), _vp(XL), _vp(XU), _vp(X2a), float(theta),
                                     int(get_obj), ctypes.byref(t), _vp(row_start), _vp(col), _vp(val), _vp(LB),
                                     _vp(UB), _vp(src)))
        return dict(rows_per_point=rpp.value, row_start=row_start, col=col, val=val, lb=LB, ub=UB, src_row=src)

    def violation_batch(self, X, OBJ=None):
        """getConstraintsViolation (OBJ None) / getNonLinearitiesViolation per point."""
        X = _d(np.atleast_2d(X))
        B = X.shape[0]
        V = np.zeros(B)
        F = np.zeros(B)
        O = None if OBJ is None else _d(OBJ)
        self._ck(lib().bnlp_violation_batch(self.h, B, _vp(X), _vp(O), _vp(V), _vp(F)))
        return V, F

    def benders_batch(self, X, DUALS, XL=None, XU=None, with_objective=True, tols=None):
        n = self.n
        X = _d(np.atleast_2d(X))
        B = X.shape[0]
        info = self.problem.info()
        XL = _d(info["xl"] if XL is None else XL)
        XU = _d(info["xu"] if XU is None else XU)
        Du = _d(np.atleast_2d(DUALS))
        rl = np.zeros(B, np.int32)
        ci = np.zeros((B, n + 1), np.int32)
        cv = np.zeros((B, n + 1))
        ub = np.zeros(B)
        t = tols or OaTols(1e-8, 1e-17, 1e-8, 1e19, 0)
        self._ck(lib().bnlp_benders_batch(self.h, B, _vp(X), _vp(Du), _vp(XL), _vp(XU), int(with_objective),
                                          ctypes.byref(t), _vp(rl), _vp(ci), _vp(cv), _vp(ub)))
        return [(ci[b, :rl[b]].tolist(), cv[b, :rl[b]].copy(), ub[b]) for b in range(B)]

    def branching_tqp(self, x_sol, duals_sol):
        """BranchingTQP at (x_sol, duals_sol): returns (Problem of the QP in displacement variables, f(x_sol))."""
        x = _d(x_sol)
        du = _d(duals_sol)
        f0 = c_dbl()
        err = ctypes.create_string_buffer(512)
        h = lib().bnlp_branching_tqp(self.h, _vp(x), _vp(du), ctypes.byref(f0), err, 512)
        if not h:
            raise BnlpError(err.value.decode() or _err())
        return Problem(h), f0.value

    def tqp_solve_batch(self, x_sol, duals_sol, XL, XU, want_x=True):
        """B x QpBranchingSolver::solveFromHotStart: children (original-variable bounds) on the QP surrogate."""
        n = self.n
        XL = _d(np.atleast_2d(XL))
        XU = _d(np.atleast_2d(XU))
        B = XL.shape[0]
        x = _d(x_sol)
        du = _d(duals_sol)
        out = dict(obj=np.zeros(B), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32))
        X = np.zeros((B, n)) if want_x else None
        self._ck(lib().bnlp_tqp_solve_batch(self.h, _vp(x), _vp(du), B, _vp(XL), _vp(XU), _vp(out["obj"]),
                                            _vp(out["status"]), _vp(out["iters"]), _vp(X)))
        if want_x:
            out["x"] = X
        return out

    def extract_linear_relaxation(self, x, duals=None, xl=None, xu=None, get_obj=True, tols=None):
        """OsiTMINLPInterface::extractLinearRelaxation at x: the initial LP as a dict (CSR rows, bounds, objective)."""
        n, m = self.n, self.m
        cap = self.problem.nnz_jac + n + 1
        rs = np.zeros(m + 2, np.int32)
        ci = np.zeros(cap, np.int32)
        cv = np.zeros(cap)
        rl, ru, sr = np.zeros(m + 1), np.zeros(m + 1), np.zeros(m + 1, np.int32)
        cl, cu, ob, ii = np.zeros(n + 1), np.zeros(n + 1), np.zeros(n + 1), np.zeros(n + 1, np.int32)
        d = LpDims()
        t = tols or OaTols(1e-8, 1e-17, 1e-8, 1e19, 0)
        xx = _d(x)
        du = None if duals is None else _d(duals)
        xl_ = None if xl is None else _d(xl)
        xu_ = None if xu is None else _d(xu)
        self._ck(lib().bnlp_extract_linear_relaxation(self.h, _vp(xx), _vp(du), _vp(xl_), _vp(xu_), int(get_obj),
                                                      ctypes.byref(t), ctypes.byref(d), _vp(rs), _vp(ci), _vp(cv), _vp(rl),
                                                      _vp(ru), _vp(sr), _vp(cl), _vp(cu), _vp(ob), _vp(ii)))
        R, C, Z = d.n_rows, d.n_cols, d.nnz
        return dict(n_cols=C, n_rows=R, has_alpha=bool(d.has_alpha), obj_offset=d.obj_offset, row_start=rs[:R + 1].copy(),
                    col_idx=ci[:Z].copy(), val=cv[:Z].copy(), row_lb=rl[:R].copy(), row_ub=ru[:R].copy(),
                    src_row=sr[:R].copy(), col_lb=cl[:C].copy(), col_ub=cu[:C].copy(), obj=ob[:C].copy(),
                    is_integer=ii[:C].copy())

    def constraint_oa_batch(self, row, X, DUALS=None, XL=None, XU=None, tols=None):
        """B x getConstraintOuterApproximation of row `row`: list of (cols, vals, lb, ub)."""
        n = self.n
        X = _d(np.atleast_2d(X))
        B = X.shape[0]
        Du = None
        if DUALS is not None:
            Du = np.atleast_2d(np.asarray(DUALS, float))
            if Du.shape[0] == 1 and B > 1:       # one multiplier vector for every point (the reference's getRowPrice())
                Du = np.repeat(Du, B, axis=0)
            if Du.shape != (B, 2 * n + self.m):
                raise ValueError(f"oa_batch: DUALS must be {B} x {2 * n + self.m}, got {Du.shape}")
            Du = _d(Du)
        xl = None if XL is None else _d(XL)
        xu = None if XU is None else _d(XU)
        rl, ci, cv = np.zeros(B, np.int32), np.zeros((B, n), np.int32), np.zeros((B, n))
        lb, ub = np.zeros(B), np.zeros(B)
        t = tols or OaTols(1e-8, 1e-17, 1e-8, 1e19, 0)
        self._ck(lib().bnlp_constraint_oa_batch(self.h, B, int(row), _vp(X), _vp(Du), _vp(xl), _vp(xu), ctypes.byref(t),
                                                _vp(rl), _vp(ci), _vp(cv), _vp(lb), _vp(ub)))
        return [(ci[b, :rl[b]].tolist(), cv[b, :rl[b]].copy(), lb[b], ub[b]) for b in range(B)]

    def stats(self):
        it, fa, ms = ctypes.c_longlong(), ctypes.c_longlong(), c_dbl()
        lib().bnlp_engine_stats(self.h, ctypes.byref(it), ctypes.byref(fa), ctypes.byref(ms))
        return dict(iterations=it.value, factorisations=fa.value, device_ms=ms.value)

    def launches(self):
        return int(lib().bnlp_engine_launches(self.h))

    def fp64_peak_tflops(self, dev_index=0):
        return float(lib().bnlp_measure_fp64_peak(self.h, dev_index))

    def close(self):
        if getattr(self, "h", None):
            lib().bnlp_engine_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Comm:
    """Multi-rank exchange of the engine (`bnlp_comm`, NCCL inside libbnlp): one process per GPU.  `broadcast_id(buf)` must
    hand rank 0's 128-byte id to every rank (e.g. torch.distributed.broadcast of a uint8 tensor); `Comm.from_torch` does that."""

    def __init__(self, engine: "Engine", rank: int, world: int, uid: bytes):
        L = lib()
        L.bnlp_comm_create.restype = c_void
        L.bnlp_comm_create.argtypes = [c_void, c_int, c_int, ctypes.c_char_p]
        self.engine, self.rank, self.world = engine, rank, world
        self.h = L.bnlp_comm_create(engine.h, rank, world, uid)
        if not self.h:
            raise BnlpError(L.bnlp_last_error().decode())

    @staticmethod
    def unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        if lib().bnlp_comm_unique_id(buf):
            raise BnlpError(lib().bnlp_last_error().decode())
        return buf.raw

    @classmethod
    def from_torch(cls, engine, device):
        """The id travels through the existing torch.distributed process group (NCCL or gloo)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        t = torch.zeros(128, dtype=torch.uint8, device=device if dist.get_backend() == "nccl" else "cpu")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return cls(engine, rank, world, bytes(t.cpu().numpy().tobytes()))

    def solve_batch(self, XL, XU, X0, want_x=True):
        """B children, the same arrays on every rank; rank r solves its block and every rank gets every result."""
        e = self.engine
        XL = _d(np.atleast_2d(XL)); XU = _d(np.atleast_2d(XU)); X0 = _d(X0)
        B = XL.shape[0]
        OBJ = np.zeros(B); ST = np.zeros(B, np.int32); IT = np.zeros(B, np.int32)
        X = np.zeros((B, e.n)) if want_x else None
        L = lib()
        L.bnlp_solve_batch_sharded.argtypes = [c_void, c_int, c_void, c_void, c_void, c_int] + [c_void] * 4
        if L.bnlp_solve_batch_sharded(self.h, B, _vp(XL), _vp(XU), _vp(X0), int(X0.ndim == 2), _vp(OBJ), _vp(ST), _vp(IT), _vp(X)):
            raise BnlpError(L.bnlp_last_error().decode())
        return dict(obj=OBJ, status=ST, iters=IT, x=X)

    def allreduce_min(self, value: float) -> float:
        v = ctypes.c_double(value)
        L = lib()
        L.bnlp_comm_allreduce_min.argtypes = [c_void, ctypes.POINTER(ctypes.c_double)]
        if L.bnlp_comm_allreduce_min(self.h, ctypes.byref(v)):
            raise BnlpError(L.bnlp_last_error().decode())
        return v.value

    def stats(self):
        n = ctypes.c_longlong(); ms = ctypes.c_double()
        L = lib()
        L.bnlp_comm_stats.argtypes = [c_void, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_double)]
        L.bnlp_comm_stats(self.h, ctypes.byref(n), ctypes.byref(ms))
        return dict(collectives=n.value, exchange_ms=ms.value)

    def close(self):
        if self.h:
            lib().bnlp_comm_free.argtypes = [c_void]
            lib().bnlp_comm_free(self.h)
            self.h = None
