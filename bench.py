#!/usr/bin/env python
"""bench.py — child NLP relaxations solved per second (BASELINE.json metric) on N B200s of one node.

Workload (BASELINE config C3): SYN-MIQCP(n=200, n_int=100, m_q=40, m_l=40, seed=20261020);
one STEP = one batch of B child relaxations PER GPU (strong-branching / node children: root bounds + random
prior branchings + one final branch, shared start point = root optimum), solved to Bonmin's Ipopt tolerances.
Ranks solve disjoint child index ranges (weak scaling, no data-path collective; NCCL only gathers the
per-child {objective,status} and all-reduces the incumbent bound, outside the kernel).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

`value`   : device-timed throughput, inputs resident in HBM (CUDA events on the launching stream, L2 flushed
            between steps, max over ranks).
`e2e`     : the same metric through the C ABI call bnlp_solve_batch with HOST buffers (H2D of the bounds and
            D2H of objective/status/iters/x inside the timed region).
`roofline`: FP64 pipe.  achieved = algorithmic FLOPs (SURVEY.md §8d F_iter x iterations actually taken) / kernel
            time; peak = FP64 FMA peak measured live by a micro-kernel (MEASURED_PEAKS.json has no FP64 entry).
`cpu_baseline`: the CPU oracle (restatement of the Bonmin/Ipopt path; Ipopt itself is unavailable offline) on all
            host cores, on a bounded sample of the same children.  --impl reference times the same thing per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bonmin_b200 import instances as I  # noqa: E402

C3 = dict(n=200, n_int=100, m_q=40, m_l=40, seed=20261020)
WORKLOAD = "C3 SYN-MIQCP(n=200,n_int=100,m_q=40,m_l=40,seed=20261020) child relaxations, shared start point x*"


def workload_config(batch, world):
    """The workload both arms report (identical dicts: the driver compares them)."""
    return {"workload": WORKLOAD, "children_per_gpu_per_step": batch, "children_per_step": batch * world,
            "l2": "256 MiB flush write between timed steps (untimed)"}


def measured_traffic():
    """DRAM bytes per IPM iteration of ipm_coop_kernel from the committed ncu capture of this round
    (profiles/r2_traffic.json, written by scripts/ncu_traffic.py from `ncu --set full`: dram__bytes_read.sum +
    dram__bytes_write.sum of one launch / its iterations).  None when the file is absent: the bench itself cannot read
    DRAM counters."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def f_iter(n, m_q, m_l, r=2, t=1):
    """Algorithmic FLOPs of one IPM iteration."""
    m = m_q + m_l
    return (t * (2 * n * n * (m_q + 1) + 2 * m_l * n) + m_q * n * (n + 1) + m * n * (n + 1) + n ** 3 / 3.0
            + r * (2 * n * n + 4 * m * n))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_children(model, xs, B, first):
    return I.c3_children(model, xs, B, first=first)


def _all_gather(t, world):
    import torch
    import torch.distributed as dist
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return out


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


BB_INSTANCE = dict(n=60, n_int=30, m_q=12, m_l=12, seed=3)


def bb_section(args, local_rank):
    """Second half of the BASELINE metric: B-BB wall time to optimum.  Host B-BB driver (bonmin_b200/host, restating
    Bonmin's B-BB node loop) with 32 strong-branching candidates per node, every node's 2K children solved as one batch;
    arithmetic provider = CUDA engine vs CPU oracle (all cores).  Same optimum is asserted.  C2 (n=100, 50 integers) does
    not close within minutes on either arm, so it is reported as a bounded run (nodes processed per second)."""
    from bonmin_b200 import bb
    from bonmin_b200.engine import Engine, Problem
    from oracle import ipm
    cores = cpu_cores()
    out = {}
    m = I.syn_miqcp(**BB_INSTANCE)
    eng = Engine(Problem.from_model(m), devices=[local_rank])
    kw = dict(number_strong_branch=32, number_before_trust=1000000, node_limit=200000, time_limit=120.0)
    bb.run_engine(eng, **dict(kw, node_limit=20))            # warm-up (buffers, clocks)
    g = bb.run_engine(eng, **kw)
    dense = ipm.dense_arrays(m)
    c = bb.run_callback(m, lambda XL, XU, x0: ipm.solve_batch(m, XL, XU, x0, threads=cores, dense=dense), **kw)
    out["to_optimum"] = {
        "instance": "SYN-MIQCP(n=60,n_int=30,m_q=12,m_l=12,seed=3), B-BB, number_strong_branch 32, all candidates strong-branched",
        "gpu_wall_s": g["seconds"], "cpu_wall_s": c["seconds"], "cpu_cores": cores, "optimum_gpu": g["best_obj"],
        "optimum_cpu": c["best_obj"], "same_optimum_1e-6": bool(abs(g["best_obj"] - c["best_obj"]) <= 1e-6 * max(1.0, abs(c["best_obj"]))),
        "status": g["status"], "nodes": g["nodes"], "child_nlps": g["sb_children"], "ipm_iterations": g["ipm_iterations"],
        "served_from_sb_cache": g["served_from_sb_cache"]}
    # the same search with 32 best-bound nodes expanded per round: all their strong-branching children in ONE launch
    g32 = bb.run_engine(eng, **dict(kw, parallel_nodes=32))
    out["to_optimum_32_nodes_per_round"] = {
        "gpu_wall_s": g32["seconds"], "optimum_gpu": g32["best_obj"], "nodes": g32["nodes"], "launch_rounds": g32["sb_batches"],
        "child_nlps": g32["sb_children"],
        "same_optimum_1e-6": bool(abs(g32["best_obj"] - c["best_obj"]) <= 1e-6 * max(1.0, abs(c["best_obj"])))}
    # variable_selection qp-strong-branching: the children are solved on the BranchingTQP surrogate (ranking only)
    gq = bb.run_engine(eng, **dict(kw, variable_selection="qp-strong-branching"))
    out["to_optimum_qp_strong_branching"] = {
        "gpu_wall_s": gq["seconds"], "optimum_gpu": gq["best_obj"], "nodes": gq["nodes"], "child_qps": gq["sb_children"],
        "same_optimum_1e-6": bool(abs(gq["best_obj"] - c["best_obj"]) <= 1e-6 * max(1.0, abs(c["best_obj"])))}
    eng.close()
    m = I.config_c2()
    eng = Engine(Problem.from_model(m), devices=[local_rank])
    kw = dict(number_strong_branch=32, number_before_trust=1000000, node_limit=300, time_limit=120.0)
    bb.run_engine(eng, **dict(kw, node_limit=10))
    g = bb.run_engine(eng, **kw)
    g32 = bb.run_engine(eng, **dict(kw, parallel_nodes=32))
    out["c2_bounded"] = {"instance": m.name + ", B-BB, 32 candidates/node (64 children/node), first 300 nodes",
                         "gpu_wall_s": g["seconds"], "nodes": g["nodes"], "child_nlps": g["sb_children"],
                         "children_per_s": g["sb_children"] / g["seconds"], "best_obj": g["best_obj"], "status": g["status"],
                         "with_32_nodes_per_round": {"gpu_wall_s": g32["seconds"], "nodes": g32["nodes"],
                                                     "child_nlps": g32["sb_children"],
                                                     "children_per_s": g32["sb_children"] / max(g32["seconds"], 1e-9)}}
    # BASELINE config 2 to proven optimality: 128 best-bound nodes per round (their ~8 k strong-branching children are
    # one launch).  The CPU arm is not run to the end (tens of minutes); its rate on a sample of the same children is given.
    gf = bb.run_engine(eng, **dict(kw, node_limit=5000000, time_limit=150.0, parallel_nodes=128))
    root = eng.solve_batch(m.xl, m.xu, m.x0)
    XLs, XUs = I.c3_children(m, root["x"][0], 1024)
    dense = ipm.dense_arrays(m)
    ipm.solve_batch(m, XLs[:cores], XUs[:cores], root["x"][0], threads=cores, dense=dense)
    t0 = time.perf_counter()
    ipm.solve_batch(m, XLs, XUs, root["x"][0], threads=cores, dense=dense)
    cpu_rate = 1024 / (time.perf_counter() - t0)
    out["c2_to_optimum"] = {"instance": m.name + ", B-BB, 32 candidates/node, 128 nodes per round", "gpu_wall_s": gf["seconds"],
                            "status": gf["status"], "optimum_gpu": gf["best_obj"], "nodes": gf["nodes"],
                            "child_nlps": gf["sb_children"], "ipm_iterations": gf["ipm_iterations"],
                            "launch_rounds": gf["sb_batches"], "children_per_s": gf["sb_children"] / max(gf["seconds"], 1e-9),
                            "cpu_children_per_s_on_a_sample": cpu_rate, "cpu_cores": cores,
                            "cpu_wall_s_estimate": gf["sb_children"] / cpu_rate}
    eng.close()
    return out


def sweep_section(eng, model, xs, dev, stream):
    """BASELINE config 3 ("1k-64k independent child NLPs") and config 2's unit of work (one node = 64 children) as
    driver-visible numbers: device-resident batches of 1k .. 64k C3 children (relaxations/s) and the LATENCY of a
    64-child strong-branching batch at C3 and C2 (ms per batch: what one B-BB node waits for)."""
    import torch
    from bonmin_b200.engine import Engine, Problem
    out = {"c3_batch_sweep": [], "node_batch_latency": []}
    n = model.n

    def timed(e, XL, XU, x0, reps):
        B = len(XL)
        d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (XL, XU, x0)]
        O = torch.zeros(B, dtype=torch.float64, device=dev); S = torch.zeros(B, dtype=torch.int32, device=dev)
        It = torch.zeros(B, dtype=torch.int32, device=dev); X = torch.zeros(B, XL.shape[1], dtype=torch.float64, device=dev)
        e.solve_batch_device(d[0], d[1], d[2], O, S, It, X, None, None, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            e.solve_batch_device(d[0], d[1], d[2], O, S, It, X, None, None, stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, int(It.sum().item())

    for B in (1024, 4096, 16384, 65536):
        XL, XU = make_children(model, xs, B, 0)
        ms, its = timed(eng, XL, XU, xs, 2 if B < 65536 else 1)
        out["c3_batch_sweep"].append({"children": B, "ms": ms, "relaxations_per_s": B / (ms * 1e-3), "mean_iterations": its / B})
    XL, XU = make_children(model, xs, 64, 0)
    ms, _ = timed(eng, XL, XU, xs, 10)
    out["node_batch_latency"].append({"workload": "C3 (n=200), 32 candidates = 64 children", "ms_per_batch": ms})
    m2 = I.config_c2()
    e2 = Engine(Problem.from_model(m2), devices=[dev.index or 0])
    x2 = e2.solve_batch(m2.xl, m2.xu, m2.x0)["x"][0]
    XL, XU = make_children(m2, x2, 64, 0)
    ms, _ = timed(e2, XL, XU, x2, 10)
    out["node_batch_latency"].append({"workload": "C2 (n=100, m=40), 32 candidates = 64 children", "ms_per_batch": ms})
    XL, XU = make_children(m2, x2, 8192, 0)
    ms, its = timed(e2, XL, XU, x2, 2)
    out["c2_throughput"] = {"children": 8192, "ms": ms, "relaxations_per_s": 8192 / (ms * 1e-3), "mean_iterations": its / 8192}
    e2.close()
    return out


def large_models_section(dev, stream, peak):
    """BASELINE configs 4 and 5 at full size on the large-model path (H in global memory, tile-major, DMMA): child / NLP(y)
    relaxations per second, device-resident, with the FP64 roofline from the same SURVEY.md §8d formula (dense counts: the
    algorithmic FLOPs are a property of the problem, the sparse / low-rank short cuts the engine takes are not credited)."""
    import torch
    from bonmin_b200.engine import Engine, Problem
    out = []
    for name, mk, B, (mq, ml) in (("C5 SYN-PORTFOLIO(N=200): n=400, m=203 (1 quadratic row, 1 equality)", I.config_c5, 592, (1, 202)),
                                  ("C4 SYN-CVX-MINLP(n=500, m=200: 100 quadratic+exp rows)", I.config_c4, 296, (100, 100))):
        m = mk()
        eng = Engine(Problem.from_model(m), devices=[dev.index or 0])
        root = eng.solve_batch(m.xl, m.xu, m.x0)
        xs = root["x"][0]
        XL, XU = make_children(m, xs, B, 0)
        d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (XL, XU, xs)]
        O = torch.zeros(B, dtype=torch.float64, device=dev); S = torch.zeros(B, dtype=torch.int32, device=dev)
        It = torch.zeros(B, dtype=torch.int32, device=dev); X = torch.zeros(B, m.n, dtype=torch.float64, device=dev)
        for _ in range(2):
            eng.solve_batch_device(d[0], d[1], d[2], O, S, It, X, None, None, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            eng.solve_batch_device(d[0], d[1], d[2], O, S, It, X, None, None, stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        its = int(It.sum().item())
        fl = f_iter(m.n, mq, ml) * its
        ach = fl / (ms * 1e-3) / 1e12
        out.append({"workload": name, "children": B, "ms": ms, "relaxations_per_s": B / (ms * 1e-3), "mean_iterations": its / B,
                    "optimal": int((S == 1).sum().item()), "infeasible": int((S == 3).sum().item()),
                    "roofline": {"bound": "fp64", "kernel": "ipm_kernel (tile-major large path)", "achieved": ach, "peak": peak,
                                 "unit": "TFLOP/s", "frac": ach / peak if peak > 0 else None,
                                 "flops_per_iteration": f_iter(m.n, mq, ml)}})
        eng.close()
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for this workload.  The real Bonmin/Ipopt cannot be built or
    installed offline (no Ipopt/ASL/Osi/Cbc), so this times the CPU oracle restatement — all host
    threads, each step a bounded sample of the same children."""
    if rank != 0:
        return
    from oracle import ipm
    ipm.build()
    model = I.syn_miqcp(**C3)
    dense = ipm.dense_arrays(model)
    root = ipm.solve_batch(model, model.xl, model.xu, model.x0, dense=dense)
    xs = root["x"][0]
    cores = cpu_cores()
    # the arm's step = the same children_per_step as the GPU arm (same config); bounded so that the run ends within
    # minutes on the host cores: at most --ref-cap children of the step's sequence are solved per step
    step_children = args.batch * world
    sample = min(step_children, args.ref_cap)
    XL, XU = make_children(model, xs, sample, 0)
    for _ in range(args.warmup):
        ipm.solve_batch(model, XL[: max(cores, 32)], XU[: max(cores, 32)], xs, threads=cores, dense=dense)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ipm.solve_batch(model, XL, XU, xs, threads=cores, dense=dense)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {"impl": "reference", "metric": "child NLP relaxations solved/sec", "value": val, "unit": "relaxations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.batch, world),
            "cpu_baseline": {"value": val, "unit": "relaxations/s", "cores": cores, "kind": "port",
                             "sample": f"first {sample} children of the step's {step_children}-child sequence per step "
                                       "(a rate: children/s does not depend on how many are solved), one relaxation per "
                                       "thread; CPU restatement of the Bonmin/Ipopt path — NOT Ipopt (Ipopt 3.13 cannot be "
                                       "built offline)"},
            "e2e": {"value": val, "unit": "relaxations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="children per GPU per step")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ref-cap", type=int, default=4096, help="--impl reference: children solved per step (cap)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the batch-size sweep / node-batch latency section")
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--parity-sample", type=int, default=512, help="children per rank checked against the CPU oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bb", action="store_true", help="skip the B-BB wall-time-to-optimum section")
    ap.add_argument("--no-oa", action="store_true", help="skip the C4 outer-approximation section")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from bonmin_b200 import build
    from bonmin_b200.engine import Engine, Problem

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()

    model = I.syn_miqcp(**C3)
    n, m = model.n, model.m
    B = args.batch
    eng = Engine(Problem.from_model(model), devices=[local_rank])
    root = eng.solve_batch(model.xl, model.xu, model.x0)
    assert root["status"][0] == 1
    xs = root["x"][0]
    XL, XU = make_children(model, xs, B, first=rank * B)

    # ---------------- device-resident arm
    stream = torch.cuda.current_stream()
    dXL = torch.from_numpy(XL).to(dev)
    dXU = torch.from_numpy(XU).to(dev)
    dX0 = torch.from_numpy(xs).to(dev)
    dOBJ = torch.zeros(B, dtype=torch.float64, device=dev)
    dST = torch.zeros(B, dtype=torch.int32, device=dev)
    dIT = torch.zeros(B, dtype=torch.int32, device=dev)
    dX = torch.zeros(B, n, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device():
        eng.solve_batch_device(dXL, dXU, dX0, dOBJ, dST, dIT, dX, None, None, stream=stream.cuda_stream)

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.fill_(k)              # evict L2 between timed steps (untimed)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = eng.launches() - launches0
    iters_total = int(dIT.sum().item())          # iterations of ONE step (identical inputs every step)
    status = dST.cpu().numpy()
    obj_dev = dOBJ.cpu().numpy()

    # ---------------- end-to-end arm through the C ABI with host buffers (pinned); the exchange step of SURVEY §8e
    # (all-gather of every rank's per-child {objective, status}, all-reduce(min) of the incumbent bound: what a multi-rank
    # B-BB node needs before it can branch) is INSIDE every timed step and also timed on its own with CUDA events
    from bonmin_b200 import shard
    hXL = torch.from_numpy(XL).pin_memory().numpy()
    hXU = torch.from_numpy(XU).pin_memory().numpy()
    r = eng.solve_batch(hXL, hXU, xs, want_g=False, want_duals=False, pin=True)   # result buffers: page-locked, kept
    r = eng.solve_batch(hXL, hXU, xs, want_g=False, want_duals=False, out=r)
    ex_obj = torch.empty(B, dtype=torch.float64, device=dev)
    ex_st = torch.empty(B, dtype=torch.int32, device=dev)

    def exchange(res):
        """per-child results of this rank -> every rank; best feasible child objective over all ranks."""
        ex_obj.copy_(torch.from_numpy(res["obj"]), non_blocking=True)
        ex_st.copy_(torch.from_numpy(res["status"]), non_blocking=True)
        all_obj, all_st = shard.gather_children(ex_obj, ex_st, B * world) if world == 1 else (
            torch.cat(_all_gather(ex_obj, world)), torch.cat(_all_gather(ex_st, world)))
        mine = torch.where(ex_st == 1, ex_obj, torch.full_like(ex_obj, float("inf"))).min()
        if world > 1:
            dist.all_reduce(mine, op=dist.ReduceOp.MIN)
        return all_obj, all_st, mine

    exchange(r)                       # warm-up (NCCL communicators)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    xev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        r = eng.solve_batch(hXL, hXU, xs, want_g=False, want_duals=False, out=r)   # every step copies bounds in, results out
        xev[k][0].record()
        all_obj, all_st, best = exchange(r)
        xev[k][1].record()
    best_bound = float(best.item())   # device -> host read of the step's result
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    exchange_us = 1e3 * sum(a.elapsed_time(b) for a, b in xev) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    assert np.array_equal(r["status"], status)
    assert all_obj.numel() == B * world
    h2d = 2 * B * n * 8 + n * 8
    d2h = B * (8 + 4 + 4) + B * n * 8

    # ---------------- parity of THIS rank's children against the CPU oracle (every rank checks a strided sample of its own
    # shard; rank 0 reports the worst)
    par = torch.zeros(3, dtype=torch.float64, device=dev)        # checked, status mismatches, max rel obj diff
    if not args.no_cpu_baseline:
        from oracle import ipm as _ipm
        if rank == 0:
            _ipm.build()
        if world > 1:
            dist.barrier()
        S = min(args.parity_sample, B)
        idx = np.unique(np.linspace(0, B - 1, S).astype(int))
        dense_p = _ipm.dense_arrays(model)
        o = _ipm.solve_batch(model, XL[idx], XU[idx], xs, threads=max(1, cpu_cores() // max(1, world)), dense=dense_p)
        same = o["status"] == status[idx]
        okk = (o["status"] == 1) & same
        rel = np.abs(o["obj"] - obj_dev[idx]) / np.maximum(1.0, np.abs(o["obj"]))
        par = torch.tensor([float(len(idx)), float((~same).sum()), float(rel[okk].max(initial=0.0))], dtype=torch.float64, device=dev)

    t_ms = torch.tensor([dev_ms, e2e_s * 1e3, exchange_us], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(iters_total), float((status == 1).sum()), float((status == 3).sum())],
                       dtype=torch.float64, device=dev)
    par_sum = par[:2].clone()
    par_max = par[2:].clone()
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(par_sum, op=dist.ReduceOp.SUM)
        dist.all_reduce(par_max, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, exchange_us_max = t_ms.tolist()
    iters_all, n_opt, n_inf = tot.tolist()

    # ---------------- the same step through the ENGINE's own multi-rank call (N > 1): every rank hands the whole B x N batch
    # to bnlp_solve_batch_sharded (host buffers), solves its block and receives every child's {objective, status, iterations}
    # through ncclAllGather issued by libbnlp; then the incumbent bound through ncclAllReduce(min) (bnlp_comm_allreduce_min)
    sharded = None
    if world > 1:
        from bonmin_b200.engine import Comm
        comm = Comm.from_torch(eng, dev)
        XLa, XUa = make_children(model, xs, B * world, first=0)
        comm.solve_batch(XLa, XUa, xs, want_x=False)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rs = comm.solve_batch(XLa, XUa, xs, want_x=False)
            inc = comm.allreduce_min(float(np.where(rs["status"] == 1, rs["obj"], np.inf).min()))
        torch.cuda.synchronize()
        ts = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        mine_ok = bool(np.array_equal(rs["status"][rank * B:(rank + 1) * B], status) and np.array_equal(rs["obj"][rank * B:(rank + 1) * B], obj_dev))
        agree = torch.tensor([1.0 if mine_ok else 0.0], device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        sharded = {"value": B * world * args.steps / float(ts.item()), "unit": "relaxations/s", "children_per_step": B * world,
                   "exchange_us_per_step": comm.stats()["exchange_ms"] * 1e3, "nccl_collectives": comm.stats()["collectives"],
                   "incumbent": inc, "every_rank_sees_its_own_block_bit_for_bit": bool(agree.item() == 1.0),
                   "what": "bnlp_solve_batch_sharded + bnlp_comm_allreduce_min (NCCL inside libbnlp), host buffers, no primal points"}
        # config 2's tree search with the children of every round sharded over the ranks (bnlp_bb_run_engine_sharded)
        from bonmin_b200 import bb
        from bonmin_b200.engine import Engine, Problem
        m2 = I.config_c2()
        e2 = Engine(Problem.from_model(m2), devices=[local_rank])
        c2 = Comm.from_torch(e2, dev)
        kw = dict(number_strong_branch=32, number_before_trust=1000000, node_limit=5000000, time_limit=60.0, parallel_nodes=128)
        bb.run_engine_sharded(c2, **dict(kw, node_limit=200))
        dist.barrier()
        g = bb.run_engine_sharded(c2, **kw)
        tb = torch.tensor([g["seconds"]], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        sharded["c2_bb_sharded"] = {"instance": m2.name + ", B-BB, 32 candidates/node, 128 nodes per round, children sharded over the ranks",
                                    "wall_s": float(tb.item()), "status": g["status"], "optimum": g["best_obj"], "nodes": g["nodes"],
                                    "child_nlps": g["sb_children"], "nccl_collectives": c2.stats()["collectives"]}
        c2.close(); e2.close(); comm.close()

    # ---------------- strong-scaling companion (N > 1): ONE 4096-child batch split over the ranks (a single B-BB round)
    strong = None
    if world > 1:
        Bs = max(2, 4096 // world)
        sXL = dXL[:Bs].contiguous(); sXU = dXU[:Bs].contiguous()
        sO = torch.zeros(Bs, dtype=torch.float64, device=dev); sS = torch.zeros(Bs, dtype=torch.int32, device=dev)
        sI = torch.zeros(Bs, dtype=torch.int32, device=dev); sX = torch.zeros(Bs, n, dtype=torch.float64, device=dev)
        for _ in range(2):
            eng.solve_batch_device(sXL, sXU, dX0, sO, sS, sI, sX, None, None, stream=stream.cuda_stream)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            eng.solve_batch_device(sXL, sXU, dX0, sO, sS, sI, sX, None, None, stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        sms = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        strong = {"children_total": Bs * world, "children_per_gpu": Bs, "ms": float(sms.item()),
                  "relaxations_per_s": Bs * world / (float(sms.item()) * 1e-3)}

    if rank == 0:
        total = B * world
        value = total * args.steps / (dev_ms_max * 1e-3)
        e2e = total * args.steps / (e2e_ms_max * 1e-3)
        peak = eng.fp64_peak_tflops(0)
        fl_step = f_iter(C3["n"], C3["m_q"], C3["m_l"]) * iters_total      # this rank's kernel
        kern_ms = dev_ms / args.steps
        achieved = fl_step / (kern_ms * 1e-3) / 1e12
        traffic = measured_traffic()
        line = {
            "metric": "child NLP relaxations solved/sec", "value": value, "unit": "relaxations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(B, world),
            "results": {"mean_ipm_iterations": iters_all / total,
                        "ipm_iterations_p50_p99_max_rank0": [float(v) for v in torch.quantile(
                            dIT.double().flatten(), torch.tensor([0.5, 0.99, 1.0], dtype=torch.float64, device=dIT.device)).tolist()],
                        "optimal": int(n_opt), "infeasible": int(n_inf),
                        "best_child_objective": best_bound},
            "e2e": {"value": e2e, "unit": "relaxations/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world,
                    "exchange_us_per_step": exchange_us_max,
                    "exchange": "all-gather of per-child {objective, status} over all ranks + all-reduce(min) of the "
                                "incumbent bound, inside every timed step (torch.distributed / NCCL; device copies at N=1)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": "ipm_coop_kernel", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak if peak > 0 else None,
                         "traffic": (traffic["dram_bytes_per_iteration"] * iters_total) if traffic else None,
                         "traffic_note": (traffic or {}).get("source", "no ncu capture of this build is committed (profiles/r2_traffic.json)"),
                         "peak_source": "FP64 FMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                         "flops_per_iteration": f_iter(C3["n"], C3["m_q"], C3["m_l"]),
                         "iterations_per_launch": iters_total, "kernel_ms": kern_ms},
            "parity": {"checked": int(par_sum[0].item()), "status_mismatches": int(par_sum[1].item()),
                       "max_rel_obj_diff": float(par_max[0].item()), "tolerance": 1e-6,
                       "note": "every rank checks a strided sample of its own children against the CPU oracle"},
        }
        if strong:
            line["strong_scaling_4096"] = strong
        if sharded:
            # N > 1: the headline end-to-end number is the ENGINE's multi-rank call (host buffers in, every child's result on
            # every rank out, exchange by NCCL inside libbnlp); the per-rank variant with a torch.distributed exchange — whose
            # all-gather also absorbs the skew between the ranks' kernels — is kept beside it
            line["e2e_per_rank_with_torch_exchange"] = line["e2e"]
            line["e2e"] = {"value": sharded["value"], "unit": "relaxations/s",
                           "h2d_bytes_per_step": world * (2 * B * n * 8 + n * 8), "d2h_bytes_per_step": world * (B * world * 16),
                           "exchange_us_per_step": sharded["exchange_us_per_step"],
                           "exchange": "ncclAllGather of per-child {objective, status, iterations} + ncclAllReduce(min) of the incumbent, "
                                       "issued by libbnlp (bnlp_solve_batch_sharded, bnlp_comm_allreduce_min) inside every timed step"}
            line["e2e_engine_sharded"] = sharded
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ipm
            cores = cpu_cores()
            S = min(args.cpu_sample, B)
            dense = ipm.dense_arrays(model)
            ipm.solve_batch(model, XL[:cores], XU[:cores], xs, threads=cores, dense=dense)
            t0 = time.perf_counter()
            ipm.solve_batch(model, XL[:S], XU[:S], xs, threads=cores, dense=dense)
            dt = time.perf_counter() - t0
            S1 = min(48, S)                     # the same port on ONE thread beside the all-cores figure
            t1 = time.perf_counter()
            ipm.solve_batch(model, XL[:S1], XU[:S1], xs, threads=1, dense=dense)
            dt1 = time.perf_counter() - t1
            line["cpu_baseline"] = {"value": S / dt, "unit": "relaxations/s", "cores": cores, "kind": "port",
                                    "one_thread": {"value": S1 / dt1, "unit": "relaxations/s", "sample": f"first {S1} children"},
                                    "sample": f"first {S} children of the step's batch, one relaxation per thread; "
                                              "CPU restatement of the Bonmin/Ipopt path — NOT Ipopt (Ipopt 3.13 cannot be built offline)"}
        if world == 1 and not args.no_sweep:
            line["sweep"] = sweep_section(eng, model, xs, dev, stream)
            line["large_models"] = large_models_section(dev, stream, peak)
        if world == 1 and not args.no_bb:
            line["bb"] = bb_section(args, local_rank)
        if world == 1 and not args.no_oa:
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
            import bench_oa
            eng.close()
            line["oa"] = bench_oa.run(B=512, reps=3, cpu_points=4)
            line["oa"]["b_oa"] = bench_oa.run_b_oa()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
