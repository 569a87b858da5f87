"""Development aid: build variants of libbnlp.so that differ only in -D flags of the two cooperative-kernel translation
units (and of the general kernel's), for A/B measurements inside ONE gpurun call (box-to-box noise is ~4 %, larger than most single changes).
  python scripts/dev_variants.py name1:-DFOO=1,-DBAR name2: ...      -> bonmin_b200/_ab/libbnlp_<name>.so
Run a variant with BNLP_LIB=bonmin_b200/_ab/libbnlp_<name>.so."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import build as b

VAR_SRC = ("bnlp_launch_coop.cu", "bnlp_launch_coop256.cu", "bnlp_launch_ipm.cu")   # the kernels' translation units


def main():
    b.build()
    out = os.path.join(b.HERE, "_ab")
    os.makedirs(out, exist_ok=True)
    procs = []
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition(":")
        flags = [f for f in flags.split(",") if f]
        objs = []
        for src in VAR_SRC:
            o = os.path.join(out, f"{name}_{src[:-3]}.o")
            objs.append(o)
            procs.append(subprocess.Popen(["nvcc"] + b.NVCC_FLAGS + flags + ["-c", os.path.join(b.CSRC, src), "-o", o]))
        others = [b._obj_of(s) for s in b.SOURCES if s not in VAR_SRC]
        procs.append((name, objs + others))
    for p in procs:
        if isinstance(p, subprocess.Popen) and p.wait() != 0:
            raise SystemExit("nvcc failed")
    for p in procs:
        if isinstance(p, tuple):
            name, objs = p
            lib = os.path.join(out, f"libbnlp_{name}.so")
            subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs + ["-ldl"])
            print(lib)

if __name__ == "__main__":
    main()
