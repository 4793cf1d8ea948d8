"""Multi-GPU deferred MMA: shard the BDD collection by constraint, one process per GPU.

Within a pass BDDs are independent; the only coupling is the per-variable sum of min-marginal
differences.  Each rank sweeps its own BDDs with the single-GPU kernels in
shard mode (global ``nr_bdds_per_var``, full 2V delta vector), and after every pass the
un-normalised delta sums are all-reduced over ``torch.distributed`` (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  The next pass divides by the GLOBAL BDD count while reading.
This is the hybrid CPU+GPU solver's exchange (bdd_multi_parallel_mma_base.cu:266-354:
accumulate_delta / split_delta / normalize_delta around forward_mm / backward_mm) with a
collective in place of the host copy.

The local solver only has to provide ``forward_pass / backward_pass / delta_sum_view /
lower_bound`` (``bdd_b200.solver.bdd_cuda_parallel_mma`` does), so the sharding logic can be
exercised on CPU with a stand-in local solver (tests/test_dist_cpu.py).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .instances import BOTSINK, BddCollection


def partition_bdds(col: BddCollection, world: int) -> List[np.ndarray]:
    """Contiguous blocks of BDDs with balanced node counts (constraint order is kept, so a
    grid-tile ordered instance gets spatially compact shards)."""
    sizes = np.diff(col.delims.astype(np.int64))
    csum = np.cumsum(sizes)
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(csum, total * r / world, side="left")) + 1)
    bounds.append(col.nr_bdds)
    bounds = np.maximum.accumulate(np.minimum(bounds, col.nr_bdds))
    return [np.arange(bounds[r], bounds[r + 1]) for r in range(world)]


def global_nr_bdds_per_var(col: BddCollection, nr_variables: Optional[int] = None) -> np.ndarray:
    """In how many BDDs each variable occurs (bdd_cuda_base.cu:66-78), for the whole collection."""
    idx = col.instrs[:, 2]
    inner = idx < BOTSINK
    var = idx[inner].astype(np.int64)
    # one count per (BDD, variable): count layer heads = positions where the variable changes
    bdd_of = np.repeat(np.arange(col.nr_bdds), np.diff(col.delims.astype(np.int64)))[inner]
    head = np.ones(var.shape[0], dtype=bool)
    head[1:] = (var[1:] != var[:-1]) | (bdd_of[1:] != bdd_of[:-1])
    n = int(var.max()) + 1 if nr_variables is None else nr_variables
    return np.bincount(var[head], minlength=n).astype(np.int32)


class sharded_mma:
    """One rank's view of a sharded ``cuda parallel mma`` solve."""

    def __init__(self, col: BddCollection, costs: Sequence[float], rank: int, world: int,
                 make_local: Callable[[BddCollection, np.ndarray, int, np.ndarray], object],
                 group=None, shard_ids: Optional[np.ndarray] = None):
        self.rank, self.world, self.group = rank, world, group
        self.nr_vars = col.nr_variables()
        costs = np.asarray(costs, dtype=np.float64)
        if costs.shape[0] < self.nr_vars:
            costs = np.concatenate([costs, np.zeros(self.nr_vars - costs.shape[0])])
        self.counts = global_nr_bdds_per_var(col, self.nr_vars)
        self.ids = partition_bdds(col, world)[rank] if shard_ids is None else shard_ids
        self.local_col = col.select(self.ids)
        self.local = make_local(self.local_col, costs, self.nr_vars, self.counts)

    def _allreduce(self, t: torch.Tensor):
        if self.world <= 1:
            return
        st = getattr(self.local, "stream", None)
        if st is not None and t.is_cuda:
            with torch.cuda.stream(st):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def iteration(self, omega: float = 0.5):
        """bdd_multi_parallel_mma_base.cu:320-354 with all-reduce as the exchange step."""
        self.local.forward_pass(omega)
        self._allreduce(self.local.delta_sum_view())
        self.local.backward_pass(omega)
        self._allreduce(self.local.delta_sum_view())

    def lower_bound(self) -> float:
        lb = torch.tensor([self.local.lower_bound()], dtype=torch.float64)
        if self.world > 1:
            dev = getattr(self.local, "device", None)
            if dev is not None and dist.get_backend(self.group) == "nccl":
                lb = lb.to(dev)
            self._allreduce(lb)
        return float(lb.item())


def make_cuda_local(precision: str, device: int, deterministic: bool = False):
    """Factory for the GPU local solver: bdd_cuda_parallel_mma in shard mode.  The collective is
    issued under the solver's own stream (NCCL enqueues on torch's current stream), so pass and
    exchange stay ordered without host synchronisation."""
    from .solver import bdd_cuda_parallel_mma

    def make(col: BddCollection, costs: np.ndarray, nr_vars: int, counts: np.ndarray):
        return bdd_cuda_parallel_mma(col, costs, precision=precision, device=device, deterministic=deterministic,
                                     nr_variables=nr_vars, nr_bdds_per_var=counts)

    return make
