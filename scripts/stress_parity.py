"""Randomised parity sweep (not part of the test suite): many shapes / seeds / batch sizes, CUDA engine vs CPU oracle."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import instances as I
from bonmin_b200.engine import Problem, Engine
from oracle import ipm
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
t0 = time.time()
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    n = int(rng.choice([30, 36, 48, 64, 90, 100, 128, 160, 200]))
    n_int = n // 2
    m_q = int(rng.integers(1, max(2, n // 5)))
    m_l = int(rng.integers(0, max(1, n // 5)))
    seed = int(rng.integers(1, 10 ** 6))
    B = int(rng.choice([1, 7, 31, 33, 64, 147, 149, 300, 700]))
    fam = sys.argv[3] if len(sys.argv) > 3 else "miqcp"
    if fam == "cvx":
        n = int(rng.choice([24, 40, 60, 90])); n_int = n // 3
        m = I.syn_cvx_minlp(n, n_int, max(2, m_q // 2), max(n_int + 2, m_l), seed)
        B = min(B, 149)
    elif fam == "portfolio":
        m = I.syn_portfolio(int(rng.choice([10, 20, 40])), seed); n = m.n; n_int = n // 2
        B = min(B, 149)
    else:
        m = I.syn_miqcp(n, n_int, m_q, m_l, seed)
    eng = Engine(Problem.from_model(m))
    root = eng.solve_batch(m.xl, m.xu, m.x0)
    if root["status"][0] != 1:
        print("root not optimal", n, m_q, m_l, seed, root["status"]); eng.close(); continue
    xs = root["x"][0]
    XL, XU = I.c3_children(m, xs, B)
    # a few harder children: random extra fixings, crossed/tight bounds
    ints = np.flatnonzero(np.asarray(m.vtype) != 0)
    for b in range(0, B, 5):
        j = int(rng.choice(ints)); XL[b, j] = XU[b, j] = np.round(rng.uniform(m.xl[j], m.xu[j]))
    r = eng.solve_batch(XL, XU, xs)
    o = ipm.solve_batch(m, XL, XU, xs, threads=16)
    same = r["status"] == o["status"]
    ok = (o["status"] == 1) & same
    rel = np.abs(r["obj"] - o["obj"]) / np.maximum(1.0, np.abs(o["obj"]))
    worst = rel[ok].max(initial=0.0)
    flag = "" if same.all() and worst <= 1e-6 else "  <-- MISMATCH"
    bad += bool(flag)
    print(f"n={n:3d} m_q={m_q:2d} m_l={m_l:2d} seed={seed:6d} B={B:3d}: statuses equal {int(same.sum())}/{B}, optimal {int((o['status']==1).sum())}, "
          f"infeasible {int((o['status']==3).sum())}, other {int(((o['status']!=1)&(o['status']!=3)).sum())}, max rel obj diff {worst:.2e}, iters {r['iters'].mean():.1f}{flag}")
    if flag:
        idx = np.flatnonzero(~same)[:5]
        print("   mismatching children", idx, r["status"][idx], o["status"][idx], r["obj"][idx], o["obj"][idx])
        from bonmin_b200 import files   # two-file repro of the whole batch (gpurun_out/ travels back from the GPU box)
        os.makedirs("gpurun_out", exist_ok=True)
        files.write_batch_replay(f"gpurun_out/mismatch_{fam}_{n}_{m_q}_{m_l}_{seed}", XL, XU, m.xl, m.xu, xs, m=m.m)
    eng.close()
print("trials with mismatches:", bad, "time", time.time() - t0)
