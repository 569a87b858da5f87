"""Regenerates the committed fixtures of tests/golden (run from the repo root: python tests/golden/make_golden.py).
Nothing here reads /root/reference: the toy's known answers are the analytic values quoted by
test/InterfaceTest.cpp:122-264, restated below; the .nl samples come from the seeded generators."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from bonmin_b200 import instances as I, nlwrite  # noqa: E402
from oracle.struct_eval import Evaluator  # noqa: E402

SQ5 = 5 ** 0.5
golden = {
    "source": "test/InterfaceTest.cpp:59-264 (analytic values), variable order y1,y2,x,z of test/mytoy.nl",
    "n": 4, "m": 3, "nnz_jac": 7, "nnz_hess": 2,
    "var_type": [0, 0, 1, 2], "xl": [0, 0, 0, 0], "xu": [1e19, 1e19, 1, 5], "gu": [0.25, 0.0, 2.0],
    "root": {"obj": -(1.5 + SQ5 / 2), "x": [0.5 + 1 / SQ5, 0.5 + 1 / (2 * SQ5), 0.5 + 1 / SQ5],
             "g": [0.25, 0.0], "lambda": [SQ5, 1.0, 0.0], "rel_tol": 1e-7},
    "children": {"x>=1": {"x": 1.0, "obj": -2.5, "status": 1}, "x<=0": {"x": 0.0, "obj": -(1 + 2 ** -0.5), "status": 1}},
    "oa_row0": {"cols": [0, 1], "vals": [2 / SQ5, 1 / SQ5], "ub": 0.5 + 3 / (2 * SQ5), "rhs_relax": 1e-8},
    "jac_structure": [[0, 0], [0, 1], [1, 0], [1, 2], [2, 1], [2, 2], [2, 3]],
}
json.dump(golden, open(os.path.join(HERE, "mytoy_golden.json"), "w"), indent=1)

samples = {}
for name, model in [("miqcp_n8", I.syn_miqcp(8, 0, 2, 2, 11)), ("cvx_n10", I.syn_cvx_minlp(10, 0, 2, 3, 12))]:
    txt = nlwrite.write_nl(model)
    open(os.path.join(HERE, "nl_samples", name + ".nl"), "w").write(txt)
    ev = Evaluator(model)
    rng = np.random.default_rng(5)
    pts = (model.x0 + rng.normal(0, 0.2, (3, model.n))).tolist()
    samples[name] = {"n": model.n, "m": model.m, "nnz_jac": len(ev.jac), "nnz_hess": len(ev.hes),
                     "jac": [list(map(int, e)) for e in ev.jac], "points": pts,
                     "values": [ev.dense(np.array(p))[0].tolist() for p in pts],
                     "gu": model.gu.tolist(), "xl": model.xl.tolist(), "xu": model.xu.tolist()}
json.dump(samples, open(os.path.join(HERE, "nl_samples.json"), "w"), indent=1)
print("wrote fixtures")
