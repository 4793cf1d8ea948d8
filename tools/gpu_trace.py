"""Per-warp timeline of one forward and one backward pass (clock64 stamps), set_cover_1m."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bdd_b200 import instances
from bdd_b200.solver import bdd_cuda_parallel_mma

col, costs = instances.set_cover()
s = bdd_cuda_parallel_mma(col, costs, precision="float")
for _ in range(5):
    s.iteration()
import torch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fwd in (("forward", True), ("backward", False), ("forward, L2 flushed", True), ("backward, L2 flushed", False)):
    if "flushed" in name:
        flush.zero_(); torch.cuda.synchronize()
    tr = s.trace_pass(fwd, max_bundles=4096).astype(np.int64)
    t0 = tr[:, 0:1]
    rel = tr[:, 1:14] - t0
    valid = tr[:, 1] > 0
    rel = rel[valid]
    print(f"== {name}: {valid.sum()} bundles; median cycles since warp start at each stamp (lane kernel: 1 desc, 2 static copies issued, 3 variable loads issued, 4 previous kernel complete, 5 dynamic copies issued, 6 gathers issued, 7 zeroing issued, 8 first chunk landed, then ready/computed per chunk)")
    print("   median", np.median(rel, axis=0).astype(int).tolist())
    print("   p90   ", np.percentile(rel, 90, axis=0).astype(int).tolist())
    print("   max   ", rel.max(axis=0).tolist())
    # start skew: per SM earliest start vs each warp start
    sm = tr[valid, 15]
    starts = tr[valid, 0]
    sk = []
    for m in np.unique(sm):
        x = starts[sm == m]
        sk.append(x.max() - x.min())
    print("   start skew within an SM: median", int(np.median(sk)), "max", int(np.max(sk)), "cycles; bundles per SM median", int(np.median(np.bincount(sm.astype(int))[np.unique(sm).astype(int)])))
    if fwd:
        s.backward_pass(0.5) if False else None
