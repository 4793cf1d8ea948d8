"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):
the constraint-sharded solve (bdd_b200/dist.py) must reproduce the single-GPU solve pass by pass
(port of test/test_hybrid_parallel_mma_base.cu:14-167 with GPU shards instead of CPU + GPU).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/gpu_dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bdd_b200 import dist as bdist, instances
from bdd_b200.solver import bdd_cuda_parallel_mma


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = {
        "set_cover": lambda: instances.set_cover(m=3000, n=5000, k=9, seed=5),
        "grid_mrf": lambda: instances.grid_mrf(12, 11, 3, seed=6),
        "qap": lambda: instances.qap(n=7, seed=7),
    }
    for name, gen in cases.items():
        col, costs = gen()
        for precision, det, tol in (("double", True, 1e-9), ("double", False, 1e-9), ("float", False, 2e-4)):
            # BDDB200_SHARD_NATIVE=0: shard planned in numpy (dist.py); default: by the library (bddb200_create_shard)
            mk = bdist.make_cuda_local if os.environ.get("BDDB200_SHARD_NATIVE", "1") == "0" else bdist.make_cuda_native
            sh = bdist.sharded_mma(col, costs, rank, world, mk(precision, local, deterministic=det),
                                   exchange=os.environ.get("BDDB200_EXCHANGE", "auto"))
            whole = bdd_cuda_parallel_mma(col, costs, precision=precision, device=local, deterministic=det)
            lb_s, lb_w = sh.lower_bound(), whole.lower_bound()
            scale = max(1.0, abs(lb_w))
            good = abs(lb_s - lb_w) <= tol * scale
            worst = 0.0
            from bdd_b200.instances import BOTSINK
            local_new = np.unique(sh.local_col.instrs[sh.local_col.instrs[:, 2] < BOTSINK, 2].astype(np.int64))
            local_vars = np.argsort(sh.new_of_old)[local_new]
            sel = np.stack([2 * local_vars, 2 * local_vars + 1], axis=1).reshape(-1)
            cnt = np.maximum(bdist.global_nr_bdds_per_var(col, sh.nr_vars), 1).astype(np.float64)
            for it in range(6):
                sh.iteration(); whole.iteration()
                d_s = sh.delta_sums().astype(np.float64)               # raw sums, original variable order
                d_s[0::2] /= cnt; d_s[1::2] /= cnt
                d_w = whole.get_delta().double().cpu().numpy()
                err = float(np.abs(d_s[sel] - d_w[sel]).max()) / max(1.0, float(np.abs(d_w).max()))
                lb_s, lb_w = sh.lower_bound(), whole.lower_bound()
                worst = max(worst, err, abs(lb_s - lb_w) / scale)
            # the graph-replayed form (pass + exchange inside the library)
            sh.iterations(7); whole.iterations(7)
            d_s = sh.delta_sums().astype(np.float64)
            d_s[0::2] /= cnt; d_s[1::2] /= cnt
            d_w = whole.get_delta().double().cpu().numpy()
            err = float(np.abs(d_s[sel] - d_w[sel]).max()) / max(1.0, float(np.abs(d_w).max()))
            lb_s, lb_w = sh.lower_bound(), whole.lower_bound()
            worst = max(worst, err, abs(lb_s - lb_w) / scale)
            good = good and worst <= tol
            ok &= good
            if rank == 0:
                print(f"{name:10s} {precision:6s} det={int(det)} world={world} exchange={sh.exchange} shared={sh.n_shared}/{sh.nr_vars}: lb sharded {lb_s:.9f} whole {lb_w:.9f} worst rel err {worst:.2e} {'OK' if good else 'MISMATCH'}", flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST PARITY OK" if int(t.item()) == 1 else "DIST PARITY FAILED", flush=True)
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
