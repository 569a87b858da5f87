"""C2 (BASELINE config 2) B-BB to optimum on the engine with multi-node rounds."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import bb, instances as I
from bonmin_b200.engine import Problem, Engine
m = I.config_c2()
import torch
nd = int(os.environ.get('BNLP_NDEV', '1'))
eng = Engine(Problem.from_model(m), devices=list(range(min(nd, torch.cuda.device_count()))))
print('devices', min(nd, torch.cuda.device_count()))
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
bb.run_engine(eng, number_strong_branch=32, number_before_trust=1000000, node_limit=20)
r = bb.run_engine(eng, number_strong_branch=32, number_before_trust=1000000, node_limit=5000000, time_limit=T, parallel_nodes=P)
print({k: v for k, v in r.items() if k != "best_x"})
print("children/s", r["sb_children"] / r["seconds"])
