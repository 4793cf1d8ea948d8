"""Per-rank timing of a sharded solve (torchrun): local passes without any exchange, with the push exchange, with the one-shot pull exchange."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bdd_b200 import dist as bdist, instances
from bdd_b200._lib import check

def timed(local, n):
    st = local.stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    local.iterations(9)
    torch.cuda.synchronize(); dist.barrier()
    e0.record(st); local.iterations(n); e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    which = sys.argv[1] if len(sys.argv) > 1 else "grid_mrf_20m"
    if which == "grid_mrf_20m":
        col, costs = instances.grid_mrf(283, 283, 4, seed=4); prec = "float"
    else:
        col, costs = instances.set_cover(m=25000 * world, n=50000 * world, k=20, seed=1); prec = "float"
    out = {}
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["push", "1", "none"]
    for mode in modes:
        if mode.startswith("push") and len(mode) > 4:      # push<bits>: BDDB200_PUSH_DEBUG
            os.environ["BDDB200_PUSH_DEBUG"] = mode[4:]
            mode_name, mode = mode, "push"
        else:
            os.environ.pop("BDDB200_PUSH_DEBUG", None)
            mode_name = mode
        os.environ["BDDB200_EXCHANGE_SHOTS"] = "1" if mode == "none" else mode
        s = bdist.sharded_mma(col, costs, rank, world, bdist.make_cuda_local(prec, local))
        if mode == "none":
            check(s.local.lib.bddb200_set_exchange(s.local.h, 0, 0, None, None, None, None, None, None, 0, 0))
        out[mode_name] = timed(s.local, 300)
        nodes = s.local_col.nr_nodes; bdds = s.local_col.nr_bdds
        del s
        torch.cuda.empty_cache()
    print(f"rank {rank}/{world} {which}: local nodes {nodes} bdds {bdds}  us per iteration: " + "  ".join(f"{k} {v:.1f}" for k, v in out.items()), flush=True)
    dist.destroy_process_group()
main()
