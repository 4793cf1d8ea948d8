"""Per-warp timeline of one launch of the on-chip iteration kernel (clock64 stamps), set_cover_1m.
Stamps: 0 start, 1 state on chip, 2 owned sums published, 3 gather done, 4 forward hops done, 5 owned sums published,
6 gather done, 7 backward hops done, 8 written back."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bdd_b200 import instances
from bdd_b200.solver import bdd_cuda_parallel_mma

col, costs = instances.set_cover()
s = bdd_cuda_parallel_mma(col, costs, precision=os.environ.get("PRECISION", "float"))
s.iterations(5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = ["loaded", "reduce1", "gather1", "fwd hops", "reduce2", "gather2", "bwd hops", "written back"]
for label, do_flush in (("L2 warm", False), ("L2 flushed", True)):
    if do_flush:
        flush.zero_(); torch.cuda.synchronize()
    tr = s.trace_pass(2, max_bundles=4096).astype(np.int64)
    valid = tr[:, 0] > 0
    tr = tr[valid]
    t0 = tr[:, 0].min()
    rel = tr[:, :9] - t0
    d = np.diff(tr[:, :9], axis=1)
    print(f"== {label}: {valid.sum()} bundles; cycles since the first warp start (median / p90 / max), then per-phase durations")
    print("   at    ", ["start"] + names)
    print("   median", np.median(rel, axis=0).astype(int).tolist())
    print("   p90   ", np.percentile(rel, 90, axis=0).astype(int).tolist())
    print("   max   ", rel.max(axis=0).tolist())
    print("   phase median", dict(zip(names, np.median(d, axis=0).astype(int).tolist())))
    print("   phase p90   ", dict(zip(names, np.percentile(d, 90, axis=0).astype(int).tolist())))
    x = tr[:, 9:14]
    print("   exchange (lane 0 of each warp, whole launch): median", dict(zip(["reduce retry rounds", "gather retry rounds", "reduce first-attempt cycles", "gather first-attempt cycles", "list staging cycles"], np.median(x, axis=0).astype(int).tolist())),
          "max", x.max(axis=0).tolist())
