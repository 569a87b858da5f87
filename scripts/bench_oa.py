"""Outer-approximation throughput on C4: B x getOuterApproximation through the C ABI (host buffers).
Prints one JSON line: points/s, cuts/s, algorithmic GB/s against the measured HBM peak, CPU oracle beside it."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bonmin_b200 import instances as I
from bonmin_b200.engine import Engine, Problem


def run(B=1024, reps=5, cpu_points=8, model=None):
    m = model or I.config_c4()
    eng = Engine(Problem.from_model(m))
    rng = np.random.default_rng(5)
    X = np.clip(m.x0 + rng.uniform(-0.3, 0.3, size=(B, m.n)), m.xl, m.xu)
    X[:, m.vtype != 0] = np.round(X[:, m.vtype != 0])
    r = eng.oa_batch(X, None, m.xl, m.xu, get_obj=1, pin=True)   # warm-up (allocations; the reused cut buffers are page-locked)
    r = eng.oa_batch(X, None, m.xl, m.xu, get_obj=1, out=r)
    ts = []
    for _ in range(reps):            # the caller keeps its cut buffers (out=r), as an LP-side consumer would
        t0 = time.perf_counter()
        r = eng.oa_batch(X, None, m.xl, m.xu, get_obj=1, out=r)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    kms = eng.stats()["device_ms"]                                   # kernels only (quad + eval + rows + scan + compact)
    ncuts = int((r["src_row"] >= 0).sum())
    nnz = int(r["row_start"][-1])
    alg_bytes = B * (8 * m.n + 8 * m.m) + nnz * 12 + ncuts * 16     # x, duals in; rows out
    out = {"workload": m.name, "points": B, "cuts": ncuts, "nnz": nnz, "e2e_ms": t * 1e3, "points_per_s": B / t,
           "cuts_per_s": ncuts / t, "algorithmic_GB_per_s_e2e": alg_bytes / t / 1e9,
           "kernels_ms": kms, "points_per_s_device": B / (kms * 1e-3), "cuts_per_s_device": ncuts / (kms * 1e-3),
           "algorithmic_bytes": alg_bytes, "algorithmic_GB_per_s_device": alg_bytes / (kms * 1e-3) / 1e9}
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        hb = peak.get("hbm_gbs") or peak.get("hbm_gbs_sustained")
        if hb:
            out["hbm_peak_GB_per_s"] = hb
            out["hbm_frac_device"] = out["algorithmic_GB_per_s_device"] / hb
    except Exception:
        pass
    if cpu_points:
        # CPU baseline: the C++ restatement of getOuterApproximation (oracle/ipm_oracle.cpp: oracle_oa_batch), one point per
        # host thread at a time on every core of the box; its rows are checked against the engine's on the way
        import psutil
        from oracle import ipm as OI, struct_eval as SE
        ev = SE.Evaluator(m)
        dense = OI.dense_arrays(m)
        cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
        nb = min(B, max(cpu_points, 256))
        OI.oa_batch(m, ev, X[:min(nb, 2 * cores)], None, get_obj=1, threads=cores, dense=dense, as_cuts=False)     # warm-up
        t0 = time.perf_counter()
        c = OI.oa_batch(m, ev, X[:nb], None, get_obj=1, threads=cores, dense=dense, as_cuts=False)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": nb / dt, "unit": "points/s", "cores": cores, "kind": "port",
                               "sample": f"first {nb} of the {B} points, C++ restatement of getOuterApproximation (oracle_oa_batch), one point per thread"}
        # parity of the first points: same kept coefficients and sides as the engine
        rpp = r["rows_per_point"]
        bad = 0
        for b in range(min(nb, 16)):
            for sl in range(c["slots"]):
                k = int(c["cnt"][b * c["slots"] + sl])
                o = b * rpp + sl
                a, e = int(r["row_start"][o]), int(r["row_start"][o + 1])
                if k < 0:
                    bad += int(r["src_row"][o] >= 0)
                    continue
                oc = b * c["cap"] + int(c["cap_start"][sl])
                if e - a != k or not np.array_equal(r["col"][a:e], c["cols"][oc:oc + k]) or \
                        not np.allclose(r["val"][a:e], c["vals"][oc:oc + k], rtol=1e-10, atol=1e-12):
                    bad += 1
        out["cpu_baseline"]["rows_differing_from_the_engine_on_16_points"] = bad
    out["pipeline"] = "chunks of ~48 MB of cuts; kernels of chunk k+1 overlap the device-to-host copy of chunk k (two streams)"
    eng.close()
    return out


class _quiet_stdout:
    """HiGHS prints debug lines with printf: keep fd 1 clean while the master runs (bench.py prints ONE JSON line)."""
    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def run_b_oa(model=None):
    """B-OA to optimality on a C4-SHAPED instance through the C++ loop (bnlp_oa_run_engine: OACutGenerator2::performOa,
    bonmin_b200/host/B200Oa.cpp) with the HiGHS master behind the SubMipSolver callback; the Python loop beside it as the
    cross-check.  C4 itself (n = 500) is bound by the master (minutes per MILP solve), not by this engine."""
    from bonmin_b200 import oa
    m = model or I.syn_cvx_minlp(60, 24, 10, 10, 5)
    eng = Engine(Problem.from_model(m))
    eng.solve_batch(m.xl, m.xu, m.x0)      # warm-up
    keys = ("status", "best_obj", "lower_bound", "passes", "nlp_solves", "cuts", "oa_points")
    try:
        with _quiet_stdout():
            t0 = time.perf_counter(); g = oa.run_oa_engine(eng); tg = time.perf_counter() - t0
            p = oa.run_oa(oa.EngineProvider(eng))
        return {"workload": m.name, "algorithm": "B-OA, C++ loop (bnlp_oa_run_engine), HiGHS master behind bnlp_master_mip_cb",
                **{k: g[k] for k in keys}, "wall_s": tg, "master_s": g["master_seconds"],
                "engine_and_loop_s": g["seconds"] - g["master_seconds"],
                "same_search_as_the_python_loop": all(g[k] == p[k] for k in keys)}
    except ImportError as e:               # no scipy on the box: the master is missing, not the engine
        return {"workload": m.name, "unavailable": str(e)}
    finally:
        eng.close()


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    print(json.dumps(run(B)))
