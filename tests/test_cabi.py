"""CPU suite: the C-ABI library builds, loads and exports every symbol include/bnlp.h declares.
No compute call is made here (there is no GPU in this container and no CPU fallback in the product)."""
import os
import re

from bonmin_b200 import engine, instances as I

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header(bnlp_lib):
    hdr = open(os.path.join(ROOT, "include", "bnlp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bnlp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(bnlp_lib, name), f"libbnlp.so does not export {name}"
    assert declared == set(engine.EXPORTS)


def test_problem_from_tape_host_only(bnlp_lib):
    # model compilation is host code: sizes and structure must match TMINLP::get_nlp_info of the toy
    # (test/InterfaceTest.cpp:59-121: n=4, m=3; examples/CppExample/MyTMINLP.cpp:60-66: 7 Jacobian nonzeros, 2 Hessian)
    p = engine.Problem.from_model(I.mytoy())
    assert (p.n, p.m, p.nnz_jac, p.nnz_hess) == (4, 3, 7, 2)
    info = p.info()
    assert list(info["var_type"]) == [0, 0, 1, 2]
    assert list(info["con_nonlinear"]) == [1, 0, 0]
    assert list(info["gu"]) == [0.25, 0.0, 2.0]
    assert list(zip(info["jac_irow"], info["jac_jcol"])) == [(0, 0), (0, 1), (1, 0), (1, 2), (2, 1), (2, 2), (2, 3)]
    assert list(zip(info["hes_irow"], info["hes_jcol"])) == [(0, 0), (1, 1)]


def test_engine_needs_gpu(bnlp_lib):
    import torch
    if torch.cuda.is_available():
        return
    p = engine.Problem.from_model(I.mytoy())
    try:
        engine.Engine(p)
    except engine.BnlpError as e:
        assert "no CPU fallback" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("engine creation must fail loudly without a GPU")
