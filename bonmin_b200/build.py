"""In-tree build of libbnlp.so (sm_100a only).  `python -m bonmin_b200.build [--force] [-v]`.

Every source is compiled to an object under csrc/_obj/ (the three big kernels are separate translation units, so they
compile in parallel) and linked into bonmin_b200/libbnlp.so; objects are rebuilt when their source or any header is
newer."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libbnlp.so")
SOURCES = ["bnlp_launch_ipm.cu", "bnlp_launch_wide.cu", "bnlp_launch_coop.cu", "bnlp_launch_coop256.cu", "bnlp_api.cu", "bnlp_model.cpp", "bnlp_nl.cpp",
           "../host/B200Solver.cpp", "../host/host_cabi.cpp", "../host/node_files.cpp", "../host/B200Linearizer.cpp", "../host/B200Oa.cpp"]
HOST_LIB = os.path.join(HERE, "libbnlp_host.so")
HEADERS = ["../host/B200Solver.hpp", "../host/B200Linearizer.hpp", "../host/B200Oa.hpp", "../host/bonmin_shim.hpp", "bnlp_internal.h", "bnlp_ipm.cuh", "bnlp_ipm_coop.cuh",
           "bnlp_dense.cuh", "bnlp_dense_tm.cuh", "bnlp_eval.cuh", "bnlp_api_eval.inc", "bnlp_api_comm.inc", "../../include/bnlp.h"]
# headers each source depends on (None = all)
DEPS = {
    "bnlp_launch_ipm.cu": ["bnlp_internal.h", "bnlp_ipm.cuh", "bnlp_dense_tm.cuh", "bnlp_eval.cuh", "../../include/bnlp.h"],
    "bnlp_launch_wide.cu": ["bnlp_internal.h", "bnlp_ipm.cuh", "bnlp_dense_tm.cuh", "bnlp_eval.cuh", "../../include/bnlp.h"],
    "bnlp_launch_coop.cu": ["bnlp_internal.h", "bnlp_ipm.cuh", "bnlp_ipm_coop.cuh", "bnlp_dense.cuh", "bnlp_dense_tm.cuh", "bnlp_eval.cuh",
                            "../../include/bnlp.h"],
    "bnlp_launch_coop256.cu": ["bnlp_internal.h", "bnlp_ipm.cuh", "bnlp_ipm_coop.cuh", "bnlp_dense.cuh", "bnlp_dense_tm.cuh", "bnlp_eval.cuh",
                               "../../include/bnlp.h"],
}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-DBNLP_WITH_ENGINE"]


def _mtime(p: str) -> float:
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _obj_of(src: str) -> str:
    return os.path.join(OBJ, os.path.basename(src).rsplit(".", 1)[0] + ".o")


def _obj_stale(src: str) -> bool:
    o = _obj_of(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    deps = [src] + (DEPS.get(src) or HEADERS)
    return any(_mtime(os.path.join(CSRC, d)) > t for d in deps)


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(_mtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    procs = []
    for s in SOURCES:
        if not force and not _obj_stale(s):
            continue
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", _obj_of(s)]
        procs.append((s, subprocess.Popen(cmd)))
    bad = [s for s, p in procs if p.wait() != 0]
    if bad:
        raise RuntimeError(f"nvcc failed for {bad}")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] +
                          [_obj_of(s) for s in SOURCES] + ["-ldl"])
    build_host()
    return LIB


def build_host() -> str:
    """Host-only library (adapters + B-BB driver with a callback solver): needs no CUDA, used by the CPU tests."""
    host = os.path.join(HERE, "host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", HOST_LIB,
                           os.path.join(host, "B200Solver.cpp"), os.path.join(host, "host_cabi.cpp"),
                           os.path.join(host, "node_files.cpp"), os.path.join(host, "B200Linearizer.cpp"),
                           os.path.join(host, "B200Oa.cpp")])
    return HOST_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
