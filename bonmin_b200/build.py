"""In-tree build of libbnlp.so (sm_100a only).  `python -m bonmin_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbnlp.so")
SOURCES = ["bnlp_api.cu", "bnlp_model.cpp", "bnlp_nl.cpp", "../host/B200Solver.cpp", "../host/host_cabi.cpp"]
HOST_LIB = os.path.join(HERE, "libbnlp_host.so")
HEADERS = ["../host/B200Solver.hpp", "../host/bonmin_shim.hpp", "bnlp_internal.h", "bnlp_ipm.cuh", "bnlp_ipm_coop.cuh", "bnlp_eval.cuh", "bnlp_api_eval.inc", "../../include/bnlp.h"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-DBNLP_WITH_ENGINE", "-o", LIB] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    build_host()
    return LIB


def build_host() -> str:
    """Host-only library (adapters + B-BB driver with a callback solver): needs no CUDA, used by the CPU tests."""
    host = os.path.join(HERE, "host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", HOST_LIB,
                           os.path.join(host, "B200Solver.cpp"), os.path.join(host, "host_cabi.cpp")])
    return HOST_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
