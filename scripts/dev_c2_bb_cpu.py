"""C2 B-BB to optimum with the CPU oracle as arithmetic provider (same host driver, same options as bench.py's c2_to_optimum)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import bb, instances as I
from oracle import ipm
m = I.config_c2()
dense = ipm.dense_arrays(m)
cores = os.cpu_count()
t0 = time.time()
r = bb.run_callback(m, lambda XL, XU, x0: ipm.solve_batch(m, XL, XU, x0, threads=cores, dense=dense), number_strong_branch=32,
                    number_before_trust=1000000, node_limit=5000000, time_limit=7200.0, parallel_nodes=128)
r.pop("best_x")
r["cores"] = cores
r["wall_s"] = time.time() - t0
print(json.dumps(r))
