#!/usr/bin/env python
"""bench.py -- MMA iterations/sec of the deferred min-marginal-averaging sweep on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (N>1 under torchrun) prints
ONE JSON line from rank 0.  A step is one `iteration()` = forward pass + backward pass over the
whole BDD collection (bdd_cuda_parallel_mma.cu:142-153).

Workload.  N=1: BASELINE.json configs[1], the synthetic set-cover ILP with 25 000 rows x 50 000
columns x 20 columns per row = 1 025 000 BDD nodes, float.  N>1 (weak scaling): the
same generator with N x rows and N x columns (N x 1.025 M nodes) sharded by constraint, one
shard per GPU, the per-variable sums of the shared variables exchanged after every pass.
`value` is in iterations/s of 1.025 M-node shards: N_shards * K / time, i.e. plain iterations/s
at N=1.

Timed quantities of the top-level line
  value      device-resident state, one iteration() call per step, CUDA events on the solver's stream
             around every step, L2 flushed between timed steps (the 1 M-node working set would otherwise
             live in the 126 MB L2); `back_to_back` in the same line is the un-flushed steady state of a
             real solve (CUDA-graph replay inside the library).
  e2e        the same step driven with HOST buffers every step: update_costs(host lo, host hi) [H2D of 2V
             REALs from pinned memory, the perturbation step of the rounding loop, bdd_solver.cpp:318-380]
             -> iteration() -> lower_bound() [D2H of the bound, what run_solver does each iteration,
             run_solver_util.h:37-49].  At N=1 through the fused C-ABI call bddb200_step_host; the same step as
             three separate calls is `e2e.separate_calls`.
  roofline   forward / backward sweep kernel: algorithmic bytes per pass (SURVEY 8d formula) divided by the
             kernel's mean duration = flushed step time of the value loop / sweep launches per step (CUDA
             events around every step); a second loop of K flushed steps with an event between the two
             launches gives the per-pass split (kernel_ms_event_per_launch, _fwd_cold, _bwd).
  cpu_baseline / --impl reference: the reference's own CPU `parallel mma` solver
             (oracle/_ref/libbdd_ref.so, built from /root/reference sources) or, where that
             library is absent, the plain-C port (oracle/liboracle_mma.so), all host threads.

Additional keys (same run, same process)
  workloads  (N=1) BASELINE configs 3, 4, 5: qap_5m double, grid_mrf_20m float, `lbfgs cuda parallel mma` (history 5) on
             qap_5m -- value / back_to_back / roofline / cpu_baseline each, measured like the top-level line.
  lb_vs_time (N=1) lower bound vs wall clock at 1, 10, 100, 1000 iterations: GPU mma, GPU lbfgs, CPU reference.
  strong_scaling, parity_ok (N>1) config 4 (grid_mrf_20m sharded by constraint over the N GPUs) and the in-process
             sharded == whole check of tools/gpu_dist_check.py on a small instance.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

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
    """SM clock and throttle reasons sampled in-process through NVML while the timed regions run."""

    def __init__(self, gpu_index: int, period_s: float = 0.02):
        self.idx, self.period = gpu_index, period_s
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.error = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4))}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(self.period)
        except Exception as e:       # no NVML on this box: the line says so
            self.error = repr(e)

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(self.sm), "source": "NVML, sampled in-process every 20 ms during the timed regions"}
        if self.error:
            out["error"] = self.error
        return out


# ------------------------------------------------------------------------ CPU baseline ---
def host_threads() -> int:
    # torchrun exports OMP_NUM_THREADS=1 to its workers: count the cores this process may run on instead
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_solver(col, costs, precision):
    """(solver with .iteration()/.lower_bound(), kind, threads).  The ONLY place bench.py touches oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bindings as B
    avail = host_threads()
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


def time_cpu(col, costs, precision, warmup, max_steps, budget_s, checkpoints=None):
    with _StdoutToStderr():
        s, kind, threads = cpu_solver(col, costs, precision)
        for _ in range(warmup):
            s.iteration()
        pts = []
        t0 = time.perf_counter()
        done = 0
        while done < max_steps:
            s.iteration()
            done += 1
            if checkpoints and done in checkpoints:
                pts.append([done, time.perf_counter() - t0, s.lower_bound()])
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        return {"iters": done, "seconds": dt, "kind": kind, "threads": threads, "lb": s.lower_bound(), "points": pts}


def run_reference_cuda(args):
    """--impl reference_cuda: the reference's OWN `cuda parallel mma` (src/bdd_solver/bdd_cuda_parallel_mma.cu, compiled for sm_100a by
    oracle/Makefile from the sources under /root/reference) on this box's GPU 0, same config: a same-box GPU baseline."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    line = time_reference_cuda(args.workload, args.gpus, args.steps, max(args.warmup, 3))
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def time_reference_cuda(workload, n_shards, steps, warmup, instance=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bindings as B
    if not B.ref_cuda_available():
        return {"impl": "reference_cuda", "unavailable": "oracle/_ref/libbdd_ref_cuda.so not built (needs /root/reference at build time)"}
    col, costs, precision = instance if instance is not None else make_instance(n_shards, workload)
    sh = shape_numbers(col, len(costs))
    s = B.RefCudaSolver(col.instrs, col.delims, costs, precision)
    s.iterations(warmup)
    seconds = s.iterations(steps)
    shards = col.nr_nodes / NODES_PER_SHARD if workload == "set_cover_1m" else 1.0
    unit = "iterations/s (1.025M-node shard equivalents)" if workload == "set_cover_1m" else "iterations/s"
    value = shards * steps / seconds
    return {"impl": "reference_cuda", "metric": "mma_iterations_per_sec", "value": value, "unit": unit, "n_gpus": 1, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * seconds / steps, "higher_is_better": True, "dtype": "f32" if precision == "float" else "f64", "data": "synthetic",
            "config": {"workload": workload, "precision": precision, **sh, "l2": "not flushed (back to back)"},
            "what": "reference bdd_cuda_parallel_mma<REAL>::iteration, unmodified sources built for sm_100a, default stream, back to back, wall clock around a synchronised loop",
            "lower_bound": s.lower_bound()}


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
class Env:
    """torch / distributed context of this rank."""

    def __init__(self, gpus):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != gpus and self.world == 1 and gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.peak = float(peaks.get("hbm_gbs", 6650.0))
        self.peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]


def measure(env: Env, workload: str, K: int, W: int, with_cpu: bool, with_e2e: bool = True, lbfgs: bool = False):
    """value / back_to_back / e2e / roofline / cpu_baseline of one workload on env.world GPUs (the dict of one bench line)."""
    torch = env.torch
    from bdd_b200 import dist as bdist
    from bdd_b200.solver import bdd_cuda_parallel_mma, lbfgs_cuda_mma
    world, rank, dev = env.world, env.rank, env.dev
    col, costs, precision = make_instance(world if workload == "set_cover_1m" else 1, workload)
    R = 4 if precision == "float" else 8
    sh = shape_numbers(col, len(costs))
    V = sh["V"]

    t_c0 = time.perf_counter()
    if world > 1:
        solver = bdist.sharded_mma(col, costs, rank, world, bdist.make_cuda_native(precision, env.local_rank))      # the library plans and builds the shard
        local = solver.local
        local_sh = shape_numbers(solver.local_col, V)
    elif lbfgs:
        solver = local = lbfgs_cuda_mma(col, costs, precision=precision, device=env.local_rank, history_size=5)
        local_sh = sh
    else:
        solver = local = bdd_cuda_parallel_mma(col, costs, precision=precision, device=env.local_rank)
        local_sh = sh
    local.synchronize()
    construct_ms = 1e3 * (time.perf_counter() - t_c0)
    st = local.stream
    pass_bytes = algorithmic_bytes_per_pass(local_sh, R)

    def flush_l2():
        with torch.cuda.stream(st):
            env.flush_buf.zero_()

    lb0 = solver.lower_bound()
    for _ in range(max(W, 3)):
        flush_l2()
        solver.iteration()
    env.barrier()

    # ---- value: one iteration() call per step, events around every step, L2 flushed between steps ---------------
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    launches0 = local.kernel_launches()
    env.barrier()
    wall0 = time.perf_counter()
    for k in range(K):
        flush_l2()
        ev[k][0].record(st)
        solver.iteration()                   # forward pass + backward pass (at N > 1 each ends with the exchange of the shared variables' sums)
        ev[k][1].record(st)
    env.barrier()
    wall = time.perf_counter() - wall0
    launches = local.kernel_launches() - launches0
    total_ms = sum(e[0].elapsed_time(e[1]) for e in ev)

    # ---- roofline: the same flushed step with an event between the two sweep launches (the event keeps the second launch from
    # overlapping the first one's tail, so these steps are a little slower than the ones `value` is computed from)
    kern_ms = float("nan")
    fwd_ms = bwd_ms = [float("nan")]
    if not lbfgs:
        evp = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
        env.barrier()
        for k in range(K):
            flush_l2()
            evp[k][0].record(st)
            local.forward_pass(0.5)
            evp[k][1].record(st)
            if world > 1:
                solver.exchange_sums()       # (a no-op when the library issues the exchange itself)
            local.backward_pass(0.5)
            if world > 1:
                solver.exchange_sums()
            evp[k][2].record(st)
        env.barrier()
        fwd_ms = [e[0].elapsed_time(e[1]) for e in evp]
        bwd_ms = [e[1].elapsed_time(e[2]) for e in evp]
        kern_ms = (sum(fwd_ms) + sum(bwd_ms)) / (2 * K)          # at N > 1: pass + exchange

    # ---- back-to-back steady state (no flush; CUDA-graph replay inside the library) ------------
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nb = max(K, 300)
    if lbfgs:
        e0.record(st)
        for _ in range(nb):
            solver.iteration()
        e1.record(st)
    else:
        solver.iterations(9)
        e0.record(st)
        solver.iterations(nb)
        e1.record(st)
    env.barrier()
    b2b_ms = e0.elapsed_time(e1) / nb

    # ---- e2e: host buffers every step -----------------------------------------------------------
    e2e_total = e2e_sep_total = float("nan")
    e2e_steps = K + (K % 2)
    lb = None
    if with_e2e and not lbfgs:
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
            solver.iteration()
            return solver.lower_bound()

        def time_e2e(fused_call):
            for k in range(4):           # even count: the net perturbation is zero again
                e2e_step(k, fused_call)
            env.barrier()
            total, lbv = 0.0, None
            for k in range(e2e_steps):
                flush_l2()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                lbv = e2e_step(k, fused_call)
                total += time.perf_counter() - t0
            env.barrier()
            return total, lbv

        e2e_sep_total, lb = time_e2e(False)
        e2e_total, lb = time_e2e(True) if world == 1 else (e2e_sep_total, lb)

    # ---- run_solver-style loop (iteration + LB read-back, no cost upload) --------------------------
    t0 = time.perf_counter()
    for _ in range(K):
        solver.iteration()
        lb = solver.lower_bound()
    torch.cuda.synchronize(dev)
    rs_s = (time.perf_counter() - t0) / K

    total_ms, e2e_total, b2b_ms, kern_ms_max, rs_s, e2e_sep_total = env.max_over_ranks([total_ms, e2e_total, b2b_ms, kern_ms, rs_s, e2e_sep_total])

    shards = col.nr_nodes / NODES_PER_SHARD if workload == "set_cover_1m" else 1.0
    unit = "iterations/s (1.025M-node shard equivalents)" if workload == "set_cover_1m" else "iterations/s"
    out = None
    if rank == 0:
        # the sweep kernel's mean duration: the flushed steps of the `value` loop are nothing but sweep launches (2 per step on one
        # GPU), so step time / launches per step is the per-launch time without the cost of an event between two small kernels;
        # the per-pass figures of the second loop are kept next to it
        kern_ms_evt = kern_ms
        if not lbfgs:
            kern_ms = total_ms / (2 * K)          # two passes per step (at N > 1: pass + exchange)
        achieved = pass_bytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
        except Exception:
            pass
        cpu = None
        if with_cpu:
            budget = 12.0 if sh["N"] < 3e6 else 8.0
            r = time_cpu(col, costs, precision, 1 if sh["N"] > 3e6 else 2, 40, budget)
            cpu = {"value": r["iters"] / r["seconds"], "unit": "iterations/s", "cores": r["threads"], "kind": r["kind"],
                   "sample": f"{r['iters']} full iterations of the same {sh['N']}-node instance after warm-up, {r['threads']} OpenMP threads, {r['seconds']:.1f} s"}
        out = {
            "metric": "mma_iterations_per_sec", "value": shards * K / (total_ms * 1e-3), "unit": unit, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak" if workload == "set_cover_1m" else "strong", "vs_baseline": None,
            "dtype": "f32" if precision == "float" else "f64", "data": "synthetic",
            "config": {"workload": workload + (" + lbfgs wrapper (history 5)" if lbfgs else ""), "precision": precision, **sh, "shards": world,
                       "l2": "flushed (256 MiB memset) between timed steps; back_to_back = steady state without flush",
                       "parallelism": (f"constraint-sharded x{world}, per-pass exchange of {solver.n_shared} shared variables via {solver.exchange}"
                                       if world > 1 else "single GPU")},
            "back_to_back": {"value": shards / (b2b_ms * 1e-3), "unit": unit, "ms_per_step": b2b_ms},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": env.peak, "unit": "GB/s", "frac": achieved / env.peak,
                         "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full dram__bytes of an earlier capture, not measured in this run)" if traffic else None,
                         "kernel": "sweep_lane_kernel<REAL, MODE_MMA, fwd|bwd>" + (" + exchange" if world > 1 else ""), "kernel_ms": kern_ms,
                         "kernel_ms_source": "CUDA events around each flushed step of the value loop / 2 sweep launches per step",
                         "kernel_ms_event_per_launch": kern_ms_evt, "kernel_ms_fwd_cold": sum(fwd_ms) / K, "kernel_ms_bwd": sum(bwd_ms) / K,
                         "event_per_launch_note": "second loop of K flushed steps with an event between the two sweep launches (forward: state from HBM; backward: state the forward pass left in L2); the event itself costs ~3 us per launch",
                         "algorithmic_bytes_per_launch": pass_bytes, "peak_source": env.peak_src},
            "cpu_baseline": cpu,
            "construct_ms": construct_ms, "wall_s_timed_region": wall,
            "lower_bound": {"initial": lb0, "final": lb},
        }
        if lbfgs:
            it_l, it_m, step = solver.lbfgs_stats()
            out["lbfgs"] = {"lbfgs_iterations": it_l, "mma_iterations": it_m, "step_size": step}
            out["roofline"] = None
        if with_e2e and not lbfgs:
            out["e2e"] = {"value": shards * e2e_steps / e2e_total, "unit": unit, "h2d_bytes_per_step": 2 * V * R, "d2h_bytes_per_step": 8,
                          "step": ("bddb200_step_host(pinned host lo, hi) = update_costs + iteration + lower_bound in one C-ABI call, wall clock" if world == 1
                                   else "update_costs(host lo, host hi) + iteration() + lower_bound() through the C ABI, wall clock"),
                          "separate_calls": {"value": shards * e2e_steps / e2e_sep_total, "unit": unit, "step": "update_costs(host lo, host hi); iteration(); lower_bound() as three calls"},
                          "run_solver_loop": {"value": shards / rs_s, "unit": unit, "step": "iteration() + lower_bound()"}}
    del solver, local
    torch.cuda.empty_cache()
    return out, (col, costs, precision)


def lb_vs_time(env: Env, workload: str, instance, with_cpu: bool, cpu_budget: float):
    """Lower bound vs wall clock at 1, 10, 100, 1000 iterations: GPU mma, GPU lbfgs, CPU reference (BASELINE metric, second half;
    the reference prints this per iteration, run_solver_util.h:44-49).  GPU iterations run back to back between the checkpoints;
    every checkpoint ends with a lower_bound() read-back."""
    torch = env.torch
    from bdd_b200.solver import bdd_cuda_parallel_mma, lbfgs_cuda_mma
    col, costs, precision = instance
    checkpoints = [1, 10, 100, 1000]

    def run_gpu(s, step):
        pts = [[0, 0.0, s.lower_bound()]]
        done = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for c in checkpoints:
            step(s, c - done)
            done = c
            lb = s.lower_bound()             # waits for the device: the clock is read after it
            pts.append([c, time.perf_counter() - t0, lb])
        return pts

    out = {"workload": workload, "precision": precision, "columns": ["iterations", "seconds", "lower_bound"]}
    s = bdd_cuda_parallel_mma(col, costs, precision=precision, device=env.local_rank)
    out["gpu_mma"] = run_gpu(s, lambda s, n: s.iterations(n))
    del s
    l = lbfgs_cuda_mma(col, costs, precision=precision, device=env.local_rank, history_size=5)

    def lstep(s, n):
        for _ in range(n):
            s.iteration()
    out["gpu_lbfgs"] = run_gpu(l, lstep)
    it_l, it_m, step = l.lbfgs_stats()
    out["gpu_lbfgs_stats"] = {"lbfgs_iterations": it_l, "mma_iterations": it_m, "step_size": step}
    del l
    if with_cpu:
        r = time_cpu(col, costs, precision, 0, 1000, cpu_budget, checkpoints=set(checkpoints))
        out["cpu_reference"] = r["points"]
        out["cpu_reference_note"] = f"{r['kind']}, {r['threads']} OpenMP threads; stopped after {r['iters']} iterations ({cpu_budget:.0f} s budget)"
    torch.cuda.empty_cache()
    return out


def dist_parity(env: Env) -> bool:
    """sharded == whole on a small instance, pass by pass (the check of tools/gpu_dist_check.py, in-process)."""
    torch = env.torch
    from bdd_b200 import dist as bdist, instances
    from bdd_b200.instances import BOTSINK
    from bdd_b200.solver import bdd_cuda_parallel_mma
    ok = True
    col, costs = instances.set_cover(m=3000, n=5000, k=9, seed=5)
    for precision, tol in (("double", 1e-9), ("float", 2e-4)):
        sh = bdist.sharded_mma(col, costs, env.rank, env.world, bdist.make_cuda_native(precision, env.local_rank))
        whole = bdd_cuda_parallel_mma(col, costs, precision=precision, device=env.local_rank)
        local_new = np.unique(sh.local_col.instrs[sh.local_col.instrs[:, 2] < BOTSINK, 2].astype(np.int64))
        local_vars = np.argsort(sh.new_of_old)[local_new]
        sel = np.stack([2 * local_vars, 2 * local_vars + 1], axis=1).reshape(-1)
        cnt = np.maximum(bdist.global_nr_bdds_per_var(col, sh.nr_vars), 1).astype(np.float64)
        for it in range(4):
            if it < 2:
                sh.iteration(); whole.iteration()
            else:
                sh.iterations(3); whole.iterations(3)          # the graph-replayed form (pass + exchange inside the library)
            d_s = sh.delta_sums().astype(np.float64)
            d_s[0::2] /= cnt; d_s[1::2] /= cnt
            d_w = whole.get_delta().double().cpu().numpy()
            err = float(np.abs(d_s[sel] - d_w[sel]).max()) / max(1.0, float(np.abs(d_w).max()))
            lb_s, lb_w = sh.lower_bound(), whole.lower_bound()
            ok = ok and err <= tol and abs(lb_s - lb_w) <= tol * max(1.0, abs(lb_w))
        del sh, whole
    t = torch.tensor([1 if ok else 0], device=env.dev)
    if env.world > 1:
        env.dist.all_reduce(t, op=env.dist.ReduceOp.MIN)
    return int(t.item()) == 1


def run_ours(args):
    # libraries (NCCL's version banner, the reference's logger) write to stdout: keep the real stdout for the ONE JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    env = Env(args.gpus)
    args.gpus = env.world
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    line, instance = measure(env, args.workload, args.steps, args.warmup, with_cpu=(env.world == 1 and not args.no_cpu))
    extras = not args.no_extras and args.workload == "set_cover_1m"
    workloads, lbt, strong, parity, ref_cuda = {}, [], None, None, None
    if extras and env.world == 1:
        Kx = max(10, min(args.steps, 50))
        lbt.append(lb_vs_time(env, "set_cover_1m", instance, not args.no_cpu, 6.0))
        try:
            with _StdoutToStderr():
                ref_cuda = {"set_cover_1m": time_reference_cuda("set_cover_1m", 1, 50, 5, instance)}
        except Exception as e:
            ref_cuda = {"error": repr(e)}
        del instance
        q, q_inst = measure(env, "qap_5m", Kx, 3, with_cpu=not args.no_cpu)
        workloads["qap_5m"] = q
        try:
            with _StdoutToStderr():
                ref_cuda["qap_5m"] = time_reference_cuda("qap_5m", 1, 20, 3, q_inst)
        except Exception as e:
            ref_cuda["qap_5m_error"] = repr(e)
        l, _ = measure(env, "qap_5m", Kx, 3, with_cpu=False, lbfgs=True)
        workloads["lbfgs_qap_5m"] = l
        lbt.append(lb_vs_time(env, "qap_5m", q_inst, not args.no_cpu, 8.0))
        del q_inst
        g, _ = measure(env, "grid_mrf_20m", Kx, 3, with_cpu=not args.no_cpu)
        workloads["grid_mrf_20m"] = g
    if extras and env.world > 1:
        del instance
        parity = dist_parity(env)
        strong, _ = measure(env, "grid_mrf_20m", max(10, min(args.steps, 50)), 3, with_cpu=False, with_e2e=False)
    clocks = sampler.stop()
    if env.rank == 0:
        line["clocks"] = clocks
        if workloads:
            line["workloads"] = workloads
        if lbt:
            line["lb_vs_time"] = lbt
        if ref_cuda:
            line["reference_cuda"] = ref_cuda       # the reference's own CUDA solver on this GPU, back to back (compare with back_to_back)
        if strong is not None:
            line["strong_scaling"] = strong
        if parity is not None:
            line["parity_ok"] = parity
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if env.world > 1:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_cuda"])
    ap.add_argument("--workload", default=os.environ.get("BENCH_WORKLOAD", "set_cover_1m"))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-extras", action="store_true", help="only the top-level workload (no workloads / lb_vs_time / strong_scaling keys)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference_cuda":
        run_reference_cuda(args)
    else:
        run_ours(args)
    # the reference's static timers print to stdout at process exit: keep stdout to the JSON line
    sys.stdout.flush()
    os.dup2(2, 1)


if __name__ == "__main__":
    main()
