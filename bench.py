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
    sample = args.ref_sample
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
            "config": {"workload": WORKLOAD, "children_per_step": sample},
            "cpu_baseline": {"value": val, "unit": "relaxations/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} children per step (first {sample} of the C3 child sequence), "
                                       "CPU restatement of the Bonmin/Ipopt path (Ipopt 3.13 unavailable offline)"},
            "e2e": {"value": val, "unit": "relaxations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="children per GPU per step")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ref-sample", type=int, default=1024)
    ap.add_argument("--cpu-sample", type=int, default=2048)
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

    # ---------------- end-to-end arm through the C ABI with host buffers (pinned)
    hXL = torch.from_numpy(XL).pin_memory().numpy()
    hXU = torch.from_numpy(XU).pin_memory().numpy()
    r = eng.solve_batch(hXL, hXU, xs, want_g=False, want_duals=False, pin=True)   # result buffers: page-locked, kept
    r = eng.solve_batch(hXL, hXU, xs, want_g=False, want_duals=False, out=r)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = eng.solve_batch(hXL, hXU, xs, want_g=False, want_duals=False, out=r)   # every step copies bounds in, results out
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    assert np.array_equal(r["status"], status)
    h2d = 2 * B * n * 8 + n * 8
    d2h = B * (8 + 4 + 4) + B * n * 8

    # ---------------- exchange step: gather per-child results, all-reduce the incumbent bound
    t_ms = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(iters_total), float((status == 1).sum()), float((status == 3).sum())],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    # every rank gets every child's {objective, status} and the best bound (what a multi-GPU B-BB node needs)
    from bonmin_b200 import shard
    all_obj, all_st = shard.gather_children(dOBJ, dST, B * world) if world == 1 else (
        torch.cat([t for t in _all_gather(dOBJ, world)]), torch.cat([t for t in _all_gather(dST, world)]))
    mine = torch.where(dST == 1, dOBJ, torch.full_like(dOBJ, float("inf"))).min().item()
    best_bound = shard.allreduce_incumbent(mine, dev)
    dev_ms_max, e2e_ms_max = t_ms.tolist()
    iters_all, n_opt, n_inf = tot.tolist()

    if rank == 0:
        total = B * world
        value = total * args.steps / (dev_ms_max * 1e-3)
        e2e = total * args.steps / (e2e_ms_max * 1e-3)
        peak = eng.fp64_peak_tflops(0)
        fl_step = f_iter(C3["n"], C3["m_q"], C3["m_l"]) * iters_total      # this rank's kernel
        kern_ms = dev_ms / args.steps
        achieved = fl_step / (kern_ms * 1e-3) / 1e12
        line = {
            "metric": "child NLP relaxations solved/sec", "value": value, "unit": "relaxations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "children_per_gpu_per_step": B, "children_per_step": total,
                       "l2": "256 MiB flush write between timed steps (untimed)",
                       "mean_ipm_iterations": iters_all / total, "optimal": int(n_opt), "infeasible": int(n_inf),
                       "best_child_objective": best_bound},
            "e2e": {"value": e2e, "unit": "relaxations/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": "ipm_coop_kernel", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak if peak > 0 else None,
                         # DRAM bytes per launch: ncu dram__bytes_read + write of this kernel on this workload
                         # (profiles/r1g_traffic.csv: 12.05 GB per 52 232-iteration launch), scaled by iterations
                         "traffic": 12.05e9 / 52232 * iters_total,
                         "traffic_note": "exchange buffers of the cooperative passes (W, Q.dx) written through to HBM; "
                                         "2 % of HBM bandwidth, the kernel is FP64-pipe / latency bound",
                         "peak_source": "FP64 FMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                         "flops_per_iteration": f_iter(C3["n"], C3["m_q"], C3["m_l"]),
                         "iterations_per_launch": iters_total, "kernel_ms": kern_ms},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ipm
            ipm.build()
            cores = cpu_cores()
            S = min(args.cpu_sample, B)
            dense = ipm.dense_arrays(model)
            ipm.solve_batch(model, XL[:cores], XU[:cores], xs, threads=cores, dense=dense)
            t0 = time.perf_counter()
            o = ipm.solve_batch(model, XL[:S], XU[:S], xs, threads=cores, dense=dense)
            dt = time.perf_counter() - t0
            same = o["status"] == status[:S]
            okk = o["status"] == 1
            rel = np.abs(o["obj"] - obj_dev[:S]) / np.maximum(1.0, np.abs(o["obj"]))
            line["cpu_baseline"] = {"value": S / dt, "unit": "relaxations/s", "cores": cores, "kind": "port",
                                    "sample": f"first {S} children of the step's batch, one relaxation per thread; "
                                              "CPU restatement of the Bonmin/Ipopt path (Ipopt 3.13 unavailable offline)"}
            line["parity"] = {"checked": int(S), "status_mismatches": int((~same).sum()),
                              "max_rel_obj_diff": float(rel[okk & same].max(initial=0.0)), "tolerance": 1e-6}
        if world == 1 and not args.no_bb:
            line["bb"] = bb_section(args, local_rank)
        if world == 1 and not args.no_oa:
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
            import bench_oa
            eng.close()
            line["oa"] = bench_oa.run(B=512, reps=3, cpu_points=4)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
