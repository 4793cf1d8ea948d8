#!/usr/bin/env python
"""bench.py -- MMA iterations/sec of the deferred min-marginal-averaging sweep on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (N>1 under torchrun) prints
ONE JSON line from rank 0.  A step is one `iteration()` = forward pass + backward pass over the
whole BDD collection (bdd_cuda_parallel_mma.cu:142-153).

Workload.  N=1: BASELINE.json configs[1], the synthetic set-cover ILP with 25 000 rows x 50 000
columns x 20 columns per row = 1 025 000 BDD nodes, float.  N>1 (weak scaling): the
same generator with N x rows and N x columns (N x 1.025 M nodes) sharded by constraint, one
shard per GPU, un-normalised per-variable deltas all-reduced after every pass.
`value` is in iterations/s of 1.025 M-node shards: N_shards * K / time, i.e. plain iterations/s
at N=1.

Timed quantities
  value      device-resident state, per-step CUDA events on the solver's stream, L2 flushed
             between timed steps (the 1 M-node working set would otherwise live in the 126 MB L2);
             `back_to_back` in the same line is the un-flushed steady state of a real solve.
  e2e        the same step driven through the reference-facing API with HOST buffers every step:
             update_costs(host lo, host hi) [H2D of 2V REALs, the perturbation step of the
             rounding loop, bdd_solver.cpp:318-380] -> iteration() -> lower_bound() [D2H of the
             bound, what run_solver does each iteration, run_solver_util.h:37-49].
  roofline   forward / backward sweep kernel: algorithmic bytes per pass (SURVEY 8d formula)
             divided by the kernel's mean duration from CUDA events around each pass launch.
  cpu_baseline / --impl reference: the reference's own CPU `parallel mma` solver
             (oracle/_ref/libbdd_ref.so, built from /root/reference sources) or, where that
             library is absent, the plain-C port (oracle/liboracle_mma.so), all host threads.
"""
from __future__ import annotations

import argparse
import json
import os

import numpy as np
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NODES_PER_SHARD = 1025000


# --------------------------------------------------------------------------- workload ---
def make_instance(n_shards: int, workload: str):
    from bdd_b200 import instances
    if workload == "set_cover_1m":
        col, costs = instances.set_cover(m=25000 * n_shards, n=50000 * n_shards, k=20, seed=1)
        return col, costs, "float"
    if workload == "qap_5m":
        col, costs = instances.qap(n=40, seed=2)       # 5.06 M nodes
        return col, costs, "double"
    if workload == "grid_mrf_20m":
        col, costs = instances.grid_mrf(283, 283, 4, seed=4)   # 20.0 M nodes
        return col, costs, "float"
    if workload == "assignment_5m":
        col, costs = instances.assignment(1118, seed=3)
        return col, costs, "float"
    if workload == "assignment_5m_split_auto":   # split length chosen by bdd_b200.split.compute_split_length (fills 16 warps per SM)
        from bdd_b200.split import compute_split_length, split_long_bdds
        col, costs = instances.assignment(1118, seed=3)
        col, n_all = split_long_bdds(col, compute_split_length(col))
        return col, np.concatenate([costs, np.zeros(n_all - len(costs))]), "float"
    if workload == "assignment_5m_split64":      # the same instance after split_qbdd with chunk length 64 (the reference's answer to long BDDs)
        from bdd_b200.split import split_long_bdds
        col, costs = instances.assignment(1118, seed=3)
        col, n_all = split_long_bdds(col, 64)
        return col, np.concatenate([costs, np.zeros(n_all - len(costs))]), "float"
    raise SystemExit(f"unknown workload {workload}")


def shape_numbers(col, n_vars_total: int):
    """N_nt (non-terminal nodes), L_v (inner layers), V, B, H of a collection."""
    import numpy as np
    from bdd_b200.instances import BOTSINK
    idx = col.instrs[:, 2]
    inner = idx < BOTSINK
    n_nt = int(inner.sum())
    bdd_of = np.repeat(np.arange(col.nr_bdds), np.diff(col.delims.astype(np.int64)))
    var = idx[inner]
    b = bdd_of[inner]
    head = np.ones(n_nt, dtype=bool)
    head[1:] = (var[1:] != var[:-1]) | (b[1:] != b[:-1])
    l_v = int(head.sum())
    hops = int(np.bincount(b[head]).max())
    return {"N": int(col.nr_nodes), "N_nt": n_nt, "L_v": l_v, "V": int(n_vars_total), "B": int(col.nr_bdds), "H": hops}


def algorithmic_bytes_per_pass(sh, R: int) -> float:
    """per pass  N_nt*(8+2R) + L_v*(8+8R) + 6R*V."""
    return sh["N_nt"] * (8 + 2 * R) + sh["L_v"] * (8 + 8 * R) + 6 * R * sh["V"]


# ----------------------------------------------------------------------------- clocks ---
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------ CPU baseline ---
def cpu_solver(col, costs, precision):
    """(solver with .iteration()/.lower_bound(), kind, threads).  The ONLY place bench.py touches oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bindings as B
    # torchrun exports OMP_NUM_THREADS=1 to its workers: count the cores this process may run on instead
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    if B.ref_available():
        n = max(B.ref_max_threads(), avail)
        B.ref_set_num_threads(n)
        rc = B.RefCollection.from_arrays(col.instrs, col.delims)
        return B.RefSolver(rc, costs, precision), "reference", n
    n = max(B.oracle_max_threads(), avail)
    B.oracle_set_num_threads(n)
    return B.Oracle(col.instrs, col.delims, costs, precision), "port", n


class _StdoutToStderr:
    """The reference logs to stdout (bdd_log); keep stdout for the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def time_cpu(col, costs, precision, warmup, max_steps, budget_s):
    with _StdoutToStderr():
        return _time_cpu(col, costs, precision, warmup, max_steps, budget_s)


def _time_cpu(col, costs, precision, warmup, max_steps, budget_s):
    s, kind, threads = cpu_solver(col, costs, precision)
    for _ in range(warmup):
        s.iteration()
    t0 = time.perf_counter()
    done = 0
    while done < max_steps:
        s.iteration()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"iters": done, "seconds": dt, "kind": kind, "threads": threads, "lb": s.lower_bound()}


def run_reference(args):
    """--impl reference: the reference CPU `parallel mma` on this box's host cores, same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.gpus
    col, costs, precision = make_instance(n, args.workload)
    sh = shape_numbers(col, len(costs))
    r = time_cpu(col, costs, precision, max(args.warmup, 1), args.steps, 240.0)
    shards = col.nr_nodes / NODES_PER_SHARD if args.workload == "set_cover_1m" else 1.0
    value = shards * r["iters"] / r["seconds"]
    unit = "iterations/s (1.025M-node shard equivalents)" if args.workload == "set_cover_1m" else "iterations/s"
    line = {
        "impl": "reference", "metric": "mma_iterations_per_sec", "value": value, "unit": unit,
        "n_gpus": n, "steps": r["iters"], "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * r["seconds"] / r["iters"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if precision == "float" else "f64", "data": "synthetic",
        "config": {"workload": args.workload, "precision": precision, **sh, "shards": n},
        "cpu_baseline": {"value": value, "unit": unit, "cores": r["threads"], "kind": r["kind"],
                         "sample": f"{r['iters']} full iterations of the {sh['N']}-node instance, {r['threads']} OpenMP threads"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lower_bound": r["lb"],
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------- ours --
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from bdd_b200 import dist as bdist
    from bdd_b200.solver import bdd_cuda_parallel_mma

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    col, costs, precision = make_instance(world, args.workload)
    R = 4 if precision == "float" else 8
    sh = shape_numbers(col, len(costs))
    V = sh["V"]

    t_c0 = time.perf_counter()
    if world > 1:
        solver = bdist.sharded_mma(col, costs, rank, world, bdist.make_cuda_local(precision, local_rank))
        local = solver.local
        local_sh = shape_numbers(solver.local_col, V)
    else:
        solver = local = bdd_cuda_parallel_mma(col, costs, precision=precision, device=local_rank)
        local_sh = sh
    local.synchronize()
    construct_ms = 1e3 * (time.perf_counter() - t_c0)
    st = local.stream
    pass_bytes = algorithmic_bytes_per_pass(local_sh, R)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        with torch.cuda.stream(st):
            flush_buf.zero_()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    lb0 = solver.lower_bound()
    # one launch per iteration (the on-chip kernel, bdd_b200/csrc/resident.cuh) when the collection is eligible
    if world == 1:
        local.iteration()
    l0 = local.kernel_launches()
    if world == 1:
        local.iteration()
    fused = world == 1 and local.kernel_launches() - l0 == 1

    def one_iteration():
        if world > 1:
            solver.iteration()
        elif fused:
            local.iteration()
        else:
            local.forward_pass(0.5)
            local.backward_pass(0.5)

    for _ in range(max(args.warmup, 3)):
        flush_l2()
        one_iteration()
    barrier()

    # ---- value: per-step events, L2 flushed between steps ---------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    launches0 = local.kernel_launches()
    barrier()
    wall0 = time.perf_counter()
    for k in range(K):
        flush_l2()
        ev[k][0].record(st)
        if world > 1:
            local.forward_pass(0.5)
            ev[k][1].record(st)
            solver.exchange_sums()
            local.backward_pass(0.5)
            solver.exchange_sums()
        elif fused:
            local.iteration()
            ev[k][1].record(st)
        else:
            local.forward_pass(0.5)
            ev[k][1].record(st)
            local.backward_pass(0.5)
        ev[k][2].record(st)
    barrier()
    wall = time.perf_counter() - wall0
    launches = local.kernel_launches() - launches0
    step_ms = [e[0].elapsed_time(e[2]) for e in ev]
    fwd_ms = [e[0].elapsed_time(e[1]) for e in ev]
    total_ms = sum(step_ms)

    # ---- back-to-back steady state (no flush; graph replay at N=1) --------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nb = max(K, 300)
    if world == 1:
        local.iterations(9)
    e0.record(st)
    if world == 1:
        local.iterations(nb)
    else:
        for _ in range(nb):
            solver.iteration()
    e1.record(st)
    barrier()
    b2b_ms = e0.elapsed_time(e1) / nb

    # ---- roofline pass timing at N>1 needs the backward kernel alone --------------------------
    if fused:
        # the launch IS the iteration: forward and backward pass in one kernel
        bwd_ms = fwd_ms
        pass_bytes *= 2
    elif world == 1:
        bwd_ms = [e[1].elapsed_time(e[2]) for e in ev]
    else:
        bwd_ms = fwd_ms
    kern_ms = (sum(fwd_ms) + sum(bwd_ms)) / (len(fwd_ms) + len(bwd_ms))

    # ---- e2e: host buffers every step -----------------------------------------------------------
    # pinned host vectors of the solver's REAL type (std::vector<REAL> overload of update_costs); at N=1 the three calls
    # update_costs / iteration / lower_bound go through the fused entry point bddb200_step_host (one upload, one graph launch,
    # one read-back), and the same step made with the three separate calls is reported next to it
    rng = np.random.default_rng(123)
    tdt = torch.float64 if precision == "double" else torch.float32
    # two pinned buffers [lo | hi] (lo = zeros): adjacent vectors go up in one copy
    buf_a = torch.zeros(2 * V, dtype=tdt, pin_memory=True); buf_b = torch.zeros(2 * V, dtype=tdt, pin_memory=True)
    pert, neg = buf_a.numpy()[V:], buf_b.numpy()[V:]
    pert[:] = rng.integers(-1, 2, size=V); neg[:] = -pert

    def e2e_step(k, fused_call):
        hi = pert if k % 2 == 0 else neg
        zeros = (buf_a if k % 2 == 0 else buf_b).numpy()[:V]
        if fused_call:
            return local.step(zeros, hi)
        local.update_costs(zeros, hi)
        one_iteration()
        return solver.lower_bound()

    def time_e2e(fused_call, steps):
        for k in range(4):           # even count: the net perturbation is zero again
            e2e_step(k, fused_call)
        barrier()
        total, lbv = 0.0, None
        for k in range(steps):
            flush_l2()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            lbv = e2e_step(k, fused_call)
            total += time.perf_counter() - t0
        barrier()
        return total, lbv

    e2e_steps = K + (K % 2)
    e2e_sep_total, lb = time_e2e(False, e2e_steps)
    if world == 1:
        e2e_total, lb = time_e2e(True, e2e_steps)
    else:
        e2e_total = e2e_sep_total
    clocks = sampler.stop()

    # ---- run_solver-style loop (iteration + LB read-back, no cost upload) --------------------------
    t0 = time.perf_counter()
    for _ in range(K):
        one_iteration()
        lb = solver.lower_bound()
    torch.cuda.synchronize(dev)
    rs_s = (time.perf_counter() - t0) / K

    # ---- max over ranks ------------------------------------------------------------------------------
    red = torch.tensor([total_ms, e2e_total, b2b_ms, kern_ms, rs_s, e2e_sep_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    total_ms, e2e_total, b2b_ms, kern_ms_max, rs_s, e2e_sep_total = (float(x) for x in red.cpu())

    shards = col.nr_nodes / NODES_PER_SHARD if args.workload == "set_cover_1m" else 1.0
    unit = "iterations/s (1.025M-node shard equivalents)" if args.workload == "set_cover_1m" else "iterations/s"
    value = shards * K / (total_ms * 1e-3)
    e2e_value = shards * e2e_steps / e2e_total

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        achieved = pass_bytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu:
            r = time_cpu(col, costs, precision, 2, 40, 12.0)
            cpu = {"value": r["iters"] / r["seconds"], "unit": "iterations/s", "cores": r["threads"], "kind": r["kind"],
                   "sample": f"{r['iters']} full iterations of the same {sh['N']}-node instance after 2 warm-ups, {r['threads']} OpenMP threads, {r['seconds']:.1f} s"}
        line = {
            "metric": "mma_iterations_per_sec", "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if precision == "float" else "f64", "data": "synthetic",
            "config": {"workload": args.workload, "precision": precision, **sh, "shards": world,
                       "l2": "flushed (256 MiB memset) between timed steps; back_to_back = steady state without flush",
                       "parallelism": (f"constraint-sharded x{world}, per-pass exchange of {solver.n_shared} shared variables via {solver.exchange}"
                                       if world > 1 else "single GPU")},
            "back_to_back": {"value": shards / (b2b_ms * 1e-3), "unit": unit, "ms_per_step": b2b_ms},
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": 2 * V * R, "d2h_bytes_per_step": 8,
                    "step": ("bddb200_step_host(pinned host lo, hi) = update_costs + iteration + lower_bound in one C-ABI call, wall clock" if world == 1
                             else "update_costs(host lo, host hi) + iteration() + lower_bound() through the C ABI, wall clock"),
                    "separate_calls": {"value": shards * e2e_steps / e2e_sep_total, "unit": unit, "step": "update_costs(host lo, host hi); iteration(); lower_bound() as three calls"},
                    "run_solver_loop": {"value": shards / rs_s, "unit": unit, "step": "iteration() + lower_bound()"}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "resident_kernel<REAL> (one launch = forward + backward pass)" if fused else "sweep_lane_kernel<REAL, MODE_MMA, fwd|bwd>", "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": pass_bytes,
                         "peak_source": peak_src},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "construct_ms": construct_ms, "wall_s_timed_region": wall,
            "lower_bound": {"initial": lb0, "final": lb},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("BENCH_WORKLOAD", "set_cover_1m"))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    # the reference's static timers print to stdout at process exit: keep stdout to the JSON line
    sys.stdout.flush()
    os.dup2(2, 1)


if __name__ == "__main__":
    main()
