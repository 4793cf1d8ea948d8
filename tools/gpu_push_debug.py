"""Debug: one pass with the push exchange against one pass with the one-shot pull exchange (2+ ranks under torchrun)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bdd_b200 import dist as bdist, instances

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    col, costs = instances.set_cover(m=3000, n=5000, k=9, seed=5)
    os.environ["BDDB200_EXCHANGE_SHOTS"] = "1"
    A = bdist.sharded_mma(col, costs, rank, world, bdist.make_cuda_local("double", local))
    os.environ["BDDB200_EXCHANGE_SHOTS"] = "push"
    B = bdist.sharded_mma(col, costs, rank, world, bdist.make_cuda_local("double", local))
    print(rank, "modes", A.exchange, "|", B.exchange, "n_shared", A.n_shared, flush=True)
    for p in range(4):
        for s in (A, B):
            if p % 2 == 0: s.local.forward_pass(0.5)
            else: s.local.backward_pass(0.5)
        a = A.delta_sums(); b = B.delta_sums()
        torch.cuda.synchronize(); dist.barrier()
        a_own = A.local.delta_sum_view().cpu().numpy()
        n_ex = A.n_exchange
        b_raw = B.local.delta_sum_view().cpu().numpy()
        a_raw = np.concatenate([A.symm.sums()[:n_ex].cpu().numpy(), a_own[n_ex:]])
        d = np.abs(a_raw - b_raw)
        nz = a_raw != 0
        ratio = b_raw[nz] / a_raw[nz]
        print(f"rank {rank} pass {p}: max diff shared {d[:n_ex].max():.3e} nonshared {d[n_ex:].max():.3e}  sum a {a_raw.sum():.6f} b {b_raw.sum():.6f} own-only {a_own[:n_ex].sum():.6f} "
              f"ratio median {np.median(ratio):.4f} min {ratio.min():.4f} max {ratio.max():.4f}  lbA {A.local.lower_bound():.6f} lbB {B.local.lower_bound():.6f}", flush=True)
    dist.destroy_process_group()

main()
