"""Lower bound vs wall clock (BASELINE.json metric, second half): GPU `cuda parallel mma`, GPU `lbfgs cuda parallel mma`
(history 5) and the reference's CPU `parallel mma` (oracle/_ref, all host threads) on the same instance.
    python tools/lb_vs_time.py [workload] [cpu_seconds]
Prints one JSON line per solver: {"solver", "workload", "points": [[iterations, seconds, lower_bound], ...]}.
GPU iterations run back to back between checkpoints; every checkpoint ends with a lower_bound() read-back."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "set_cover_1m"
    cpu_budget = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
    import torch
    from bdd_b200.solver import bdd_cuda_parallel_mma, lbfgs_cuda_mma
    col, costs, precision = bench.make_instance(1, workload)
    checkpoints = [1, 2, 5, 10, 20, 50, 100, 200, 500, 1000]

    def run_gpu(name, s, step):
        pts = [[0, 0.0, s.lower_bound()]]
        done = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for c in checkpoints:
            step(s, c - done)
            done = c
            lb = s.lower_bound()
            pts.append([c, time.perf_counter() - t0, lb])
        print(json.dumps({"solver": name, "workload": workload, "precision": precision, "points": pts}), flush=True)

    s = bdd_cuda_parallel_mma(col, costs, precision=precision)
    s.iterations(3); s = bdd_cuda_parallel_mma(col, costs, precision=precision)     # first solver warmed the context up
    run_gpu("cuda parallel mma", s, lambda s, n: s.iterations(n))
    del s
    l = lbfgs_cuda_mma(col, costs, precision=precision, history_size=5, init_step_size=1e-3 if precision == "double" else 1e-3)
    def lstep(s, n):
        for _ in range(n):
            s.iteration()
    run_gpu("lbfgs cuda parallel mma (history 5)", l, lstep)
    stats = l.lbfgs_stats()
    print(json.dumps({"lbfgs_stats": {"lbfgs_iterations": stats[0], "mma_iterations": stats[1], "step_size": stats[2]}}), flush=True)
    del l
    with bench._StdoutToStderr():
        cpu, kind, threads = bench.cpu_solver(col, costs, precision)
        pts = [[0, 0.0, cpu.lower_bound()]]
        done = 0
        t0 = time.perf_counter()
        for c in checkpoints:
            for _ in range(c - done):
                cpu.iteration()
            done = c
            pts.append([c, time.perf_counter() - t0, cpu.lower_bound()])
            if time.perf_counter() - t0 > cpu_budget:
                break
    print(json.dumps({"solver": f"parallel mma (CPU {kind}, {threads} threads)", "workload": workload, "precision": precision, "points": pts}), flush=True)
    sys.stdout.flush()
    os.dup2(2, 1)


if __name__ == "__main__":
    main()
