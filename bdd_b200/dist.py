"""Multi-GPU deferred MMA: shard the BDD collection by constraint, one process per GPU.

Within a pass BDDs are independent; the only coupling is the per-variable sum of min-marginal
differences.  Each rank sweeps its own BDDs with the single-GPU kernels in
shard mode (global ``nr_bdds_per_var``, full 2V delta vector) and after every pass the
un-normalised delta sums of the variables that occur in more than one shard are summed over
all ranks; the next pass divides by the GLOBAL BDD count while reading.  This is the hybrid
CPU+GPU solver's exchange (bdd_multi_parallel_mma_base.cu:266-354: accumulate_delta /
split_delta / normalize_delta around forward_mm / backward_mm) between GPUs.

Variables are relabelled per solve so that the shared ones come first: only that prefix of
the delta vector crosses NVLink (for a grid-tile sharded MRF a thin boundary set).

Exchange back ends
  "symm"  the sum buffers live in symmetric memory (torch.distributed._symmetric_memory:
          every peer maps them); ``bddb200_delta_exchange`` -- one kernel of this library on
          the solver's stream -- signals the peers, waits for them and reads their buffers
          directly over NVLink / NVSwitch (one-shot all-reduce, fixed summation order).
  "nccl"  ``torch.distributed.all_reduce`` on the prefix (also the gloo path of the CPU tests).

The local solver only has to provide ``forward_pass / backward_pass / delta_sum_view /
lower_bound`` (``bdd_b200.solver.bdd_cuda_parallel_mma`` does), so the sharding logic can be
exercised on CPU with a stand-in local solver (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import os
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


def _layer_heads(col: BddCollection) -> Tuple[np.ndarray, np.ndarray]:
    """(variable, BDD) of every inner layer of the collection."""
    idx = col.instrs[:, 2]
    inner = idx < BOTSINK
    var = idx[inner].astype(np.int64)
    bdd_of = np.repeat(np.arange(col.nr_bdds), np.diff(col.delims.astype(np.int64)))[inner]
    head = np.ones(var.shape[0], dtype=bool)
    head[1:] = (var[1:] != var[:-1]) | (bdd_of[1:] != bdd_of[:-1])
    return var[head], bdd_of[head]


def global_nr_bdds_per_var(col: BddCollection, nr_variables: Optional[int] = None) -> np.ndarray:
    """In how many BDDs each variable occurs (bdd_cuda_base.cu:66-78), for the whole collection."""
    var, _ = _layer_heads(col)
    n = int(var.max()) + 1 if nr_variables is None else nr_variables
    return np.bincount(var, minlength=n).astype(np.int32)


def plan_shard_native(col: BddCollection, world: int, rank: int, nr_variables_min: int = 0):
    """The library's own shard planning (``bddb200_plan_shard``, bdd_b200/csrc/shard.hpp; no GPU needed): (info dict, new_of_old,
    global BDD count per NEW variable index).  Same rules as partition_bdds / shared_first_relabeling / global_nr_bdds_per_var."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    instrs = np.ascontiguousarray(col.instrs, dtype=np.uint64)
    delims = np.ascontiguousarray(col.delims, dtype=np.uint64)
    n_vars = max(col.nr_variables(), nr_variables_min)
    new_of_old = np.empty(n_vars, dtype=np.int32)
    counts_new = np.empty(n_vars, dtype=np.int32)
    masks = np.zeros(n_vars, dtype=np.uint16)
    info = _lib.ShardInfo()
    _lib.check(lib.bddb200_plan_shard(instrs.ctypes.data, instrs.shape[0], delims.ctypes.data, delims.shape[0] - 1, nr_variables_min, world, rank,
                                      C.byref(info), new_of_old.ctypes.data, counts_new.ctypes.data, masks.ctypes.data))
    d = {k: int(getattr(info, k)) for k, _ in _lib.ShardInfo._fields_}
    d["share_mask"] = masks[: d["n_shared"]].copy()       # per shared variable (new index): the shards that contain it, bit r = rank r
    return d, new_of_old.astype(np.int64), counts_new


def share_masks(col: BddCollection, parts: List[np.ndarray], new_of_old: np.ndarray, n_shared: int) -> np.ndarray:
    """Per shared variable (new index) the shards that contain it, bit r = rank r (what the push exchange sends a variable's
    differences to)."""
    var, bdd = _layer_heads(col)
    shard_of_bdd = np.empty(col.nr_bdds, dtype=np.int64)
    for r, ids in enumerate(parts):
        shard_of_bdd[ids] = r
    m = np.zeros(new_of_old.shape[0], dtype=np.uint16)
    np.bitwise_or.at(m, new_of_old[var], (1 << shard_of_bdd[bdd]).astype(np.uint16))
    return m[:n_shared].copy()


def shared_first_relabeling(col: BddCollection, parts: List[np.ndarray], nr_vars: int) -> Tuple[np.ndarray, int]:
    """new_of_old[v] = index of variable v after moving the variables that occur in more than one
    shard to the front (both groups keep their relative order), and the number of shared ones."""
    var, bdd = _layer_heads(col)
    shard_of_bdd = np.empty(col.nr_bdds, dtype=np.int64)
    for r, ids in enumerate(parts):
        shard_of_bdd[ids] = r
    sh = shard_of_bdd[bdd]
    lo = np.full(nr_vars, np.iinfo(np.int64).max, dtype=np.int64)
    hi = np.full(nr_vars, -1, dtype=np.int64)
    np.minimum.at(lo, var, sh)
    np.maximum.at(hi, var, sh)
    shared = hi > lo
    order = np.concatenate([np.nonzero(shared)[0], np.nonzero(~shared)[0]])
    new_of_old = np.empty(nr_vars, dtype=np.int64)
    new_of_old[order] = np.arange(nr_vars)
    return new_of_old, int(shared.sum())


def relabel_variables(col: BddCollection, new_of_old: np.ndarray) -> BddCollection:
    instrs = col.instrs.copy()
    inner = instrs[:, 2] < BOTSINK
    instrs[inner, 2] = new_of_old[instrs[inner, 2].astype(np.int64)].astype(np.uint64)
    return BddCollection(instrs, col.delims)


class SymmExchange:
    """Peer-memory exchange: sum buffers in symmetric memory, exchange kernels of the library issued BY the library after every
    pass (``bddb200_set_exchange``; the flag epochs live on the device, so ``iterations(n)`` replays pass + exchange as one CUDA
    graph).  Modes: "push" no exchange kernel at all -- every pass adds the shared variables' differences to its own buffer and, with
    peer-memory reductions over NVLink, to the buffers of the other ranks that hold the variable; the bundles with shared variables run
    first and carry a flag barrier (right when few layer entries belong to shared variables: each is one small packet over NVLink);
    "mc" in-switch reduction of the finished sums
    (multimem.ld_reduce / multimem.st), "1" one-shot reads of every peer, "2" two-shot.  "auto" = push when at most PUSH_MAX_ENTRIES
    layer entries of any shard belong to shared variables, else one-shot below 6 ranks and mc / two-shot from 6."""

    PUSH_MAX_ENTRIES = 1 << 16

    def __init__(self, local, n_vars: int, n_exchange: int, rank: int, world: int, group=None, shared_entries: int = -1):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        from ._lib import check
        self.lib, self.local = _lib.load(), local
        self.rank, self.world, self.n_exchange = rank, world, n_exchange
        self.precision = _lib.DOUBLE if local.precision == "double" else _lib.FLOAT
        grp = group if group is not None else dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(grp.group_name)
        except Exception:
            pass
        dev = local.device
        # every sum buffer starts on a 16-byte boundary (the in-switch reduction moves 16-byte units)
        self.stride = (2 * n_vars + 3) // 4 * 4
        self.n_total = self.stride
        n_out = max((n_exchange + 3) // 4 * 4, 4)
        self.block = symm_mem.empty(3 * self.stride, dtype=local.value_type, device=dev)
        self.block.zero_()
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=dev)
        self.flags.zero_()
        self.h_block = symm_mem.rendezvous(self.block, grp)
        self.h_flags = symm_mem.rendezvous(self.flags, grp)
        mode = os.environ.get("BDDB200_EXCHANGE_SHOTS", "auto")
        mc_in = int(getattr(self.h_block, "multicast_ptr", 0) or 0)
        has_mc = mc_in != 0
        if mode == "mc" and not has_mc:
            raise RuntimeError("BDDB200_EXCHANGE_SHOTS=mc: this box offers no multicast mapping of symmetric memory")
        if mode == "auto":
            # measured on 2 and 4 x B200 (profiles/r02_multi_gpu.md): one flag barrier + direct reads beat the two barriers of the
            # in-switch and two-shot forms while (world - 1) x prefix stays small; from 6 ranks on the slice-wise forms read less
            many = world >= 6 and n_exchange >= (1 << 17)
            mode = ("mc" if has_mc else "2") if many else "1"
            if shared_entries >= 0:
                worst = torch.tensor([shared_entries], dtype=torch.int64, device=dev)
                dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
                if int(worst.item()) <= self.PUSH_MAX_ENTRIES:
                    mode = "push"
        if mode == "push":
            # needs lane-class bundles only and the atomic sums on every rank: agree before anything is set up
            okt = torch.tensor([1 if local.push_exchange_supported() else 0], device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN, group=group)
            if int(okt.item()) == 0:        # deterministic sums or wide BDDs on some shard
                mode = "1"
        # the result buffer: plain device memory for the one-shot form (every pass gathers from it), symmetric for the others
        mc_out, self.h_out = 0, None
        if mode == "push":
            self.out = None
        elif mode == "1":
            self.out = torch.zeros(n_out, dtype=local.value_type, device=dev)
        else:
            self.out = symm_mem.empty(n_out, dtype=local.value_type, device=dev)
            self.out.zero_()
            self.h_out = symm_mem.rendezvous(self.out, grp)
            mc_out = int(getattr(self.h_out, "multicast_ptr", 0) or 0)
            if mode == "mc" and mc_out == 0:
                raise RuntimeError("no multicast mapping for the result buffer")
        self.mode = {"1": 1, "2": 2, "mc": 3, "push": 4}[mode]
        self.two_shot = self.mode == 2
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        item = self.block.element_size()
        check(self.lib.bddb200_set_delta_buffers(local.h, self.block.data_ptr(), self.block.data_ptr() + self.stride * item,
                                                 self.block.data_ptr() + 2 * self.stride * item))
        local._delta_block = self.block
        if self.out is not None:
            local.set_delta_input(self.out, n_exchange // 2)
        check(self.lib.bddb200_set_exchange(local.h, world, rank, self.h_block.buffer_ptrs_dev, self.h_flags.buffer_ptrs_dev,
                                            self.out.data_ptr() if self.out is not None else None,
                                            self.h_out.buffer_ptrs_dev if self.h_out is not None else None,
                                            mc_in if self.mode >= 3 else None, mc_out if self.mode == 3 else None, n_exchange, self.mode))
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)

    name = property(lambda self: {1: "symm one-shot", 2: "symm two-shot", 3: "symm in-switch (multimem)", 4: "peer-memory push inside the pass"}[self.mode])

    def __call__(self):
        """Nothing to do: forward_pass / backward_pass of the local solver end with the exchange."""

    def sums(self) -> Optional[torch.Tensor]:
        """The exchanged sums of the shared prefix (None in push mode: the rank's own buffer holds them)."""
        return self.out


class sharded_mma:
    """One rank's view of a sharded ``cuda parallel mma`` solve."""

    def __init__(self, col: BddCollection, costs: Sequence[float], rank: int, world: int,
                 make_local: Callable[[BddCollection, np.ndarray, int, np.ndarray], object],
                 group=None, shard_ids: Optional[np.ndarray] = None, exchange: Optional[str] = None):
        self.rank, self.world, self.group = rank, world, group
        self.nr_vars = col.nr_variables()
        costs = np.asarray(costs, dtype=np.float64)
        if costs.shape[0] < self.nr_vars:
            costs = np.concatenate([costs, np.zeros(self.nr_vars - costs.shape[0])])
        self._col = col
        if isinstance(make_local, dict):
            # native path: partition, relabelling, global counts and the shard-mode solver all come from the library
            # (bddb200_create_shard); make_local = {"precision": ..., "device": ..., "deterministic": ...}
            from .solver import bdd_cuda_parallel_mma
            self.local, info, self.new_of_old = bdd_cuda_parallel_mma.create_shard(col, costs, rank, world, **make_local)
            self.nr_vars = info["nr_variables"]
            self.n_shared = info["n_shared"]
            self.ids = np.arange(info["first_bdd"], info["first_bdd"] + info["n_bdds"])
            self._local_col = None
            self._finish_setup(exchange, info["shared_entries"])
            return
        parts = partition_bdds(col, world)
        if shard_ids is not None:
            # explicit shard of this rank: the relabelling needs every rank's shard, so all variables count as shared
            self.ids = shard_ids
            self.new_of_old, self.n_shared = np.arange(self.nr_vars), self.nr_vars
        else:
            self.ids = parts[rank]
            self.new_of_old, self.n_shared = shared_first_relabeling(col, parts, self.nr_vars)
        counts = global_nr_bdds_per_var(col, self.nr_vars)
        self.counts = np.empty_like(counts)
        self.counts[self.new_of_old] = counts
        costs_new = np.empty_like(costs)
        costs_new[self.new_of_old] = costs
        self._local_col = relabel_variables(col.select(self.ids), self.new_of_old)
        try:
            self.local = make_local(self._local_col, costs_new, self.nr_vars, self.counts, self.n_shared if shard_ids is None else 0)
        except TypeError:          # a stand-in local solver (tests) that does not take the hint
            self.local = make_local(self._local_col, costs_new, self.nr_vars, self.counts)
        if shard_ids is None and hasattr(self.local, "set_push_masks") and world <= 16:
            self.local.set_push_masks(share_masks(col, parts, self.new_of_old, self.n_shared))
        lv, _ = _layer_heads(self._local_col)
        self._finish_setup(exchange, int((lv < self.n_shared).sum()))        # layer entries of this shard that push across NVLink

    @property
    def local_col(self) -> BddCollection:
        """This rank's sub-collection with relabelled variables (built on demand on the native path)."""
        if self._local_col is None:
            self._local_col = relabel_variables(self._col.select(self.ids), self.new_of_old)
        return self._local_col

    def _finish_setup(self, exchange: Optional[str], shared_entries: int):
        rank, world, group = self.rank, self.world, self.group
        self.n_exchange = 2 * self.n_shared
        self.symm = None
        want = exchange or os.environ.get("BDDB200_EXCHANGE", "auto")
        is_cuda_local = hasattr(self.local, "set_delta_buffers") and world > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl"
        if want in ("auto", "symm") and is_cuda_local:
            ok = torch.ones(1, device=self.local.device)
            try:
                self.symm = SymmExchange(self.local, self.nr_vars, self.n_exchange, rank, world, group, shared_entries)
            except Exception as e:          # no symmetric-memory support on this box: every rank must fall back together
                if want == "symm":
                    raise
                self.symm_error = repr(e)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if ok.item() == 0 and self.symm is not None:
                from ._lib import check
                check(self.local.lib.bddb200_set_exchange(self.local.h, 0, 0, None, None, None, None, None, None, 0, 0))
                check(self.local.lib.bddb200_set_delta_buffers(self.local.h, None, None, None))
                self.local.set_delta_input(None, 0)
                self.symm = None
        self.exchange = self.symm.name if self.symm is not None else ("all_reduce" if world > 1 else "none")

    def _allreduce(self, t: torch.Tensor):
        if self.world <= 1:
            return
        st = getattr(self.local, "stream", None)
        if st is not None and t.is_cuda:
            with torch.cuda.stream(st):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def exchange_sums(self):
        """Sum the last pass's per-variable sums of the shared variables over all ranks."""
        if self.world <= 1:
            return
        if self.symm is not None:
            self.symm()
        elif self.n_exchange > 0:
            self._allreduce(self.local.delta_sum_view()[: self.n_exchange])

    def iteration(self, omega: float = 0.5):
        """bdd_multi_parallel_mma_base.cu:320-354 with the GPU-to-GPU exchange in place of the host copy."""
        self.local.forward_pass(omega)
        self.exchange_sums()
        self.local.backward_pass(omega)
        self.exchange_sums()

    def iterations(self, n: int, omega: float = 0.5):
        """n iterations; with the peer-memory exchange they run inside the library (pass, exchange, pass, exchange replayed as a CUDA graph)."""
        if self.symm is not None:
            self.local.iterations(n, omega)
        else:
            for _ in range(n):
                self.iteration(omega)

    def delta_sums(self) -> np.ndarray:
        """Un-normalised per-variable sums after the last exchange, in the ORIGINAL variable order (2V, lo/hi interleaved).
        Entries of variables that occur neither in this rank's shard nor in several shards are not meaningful."""
        t = self.local.delta_sum_view()
        st = getattr(self.local, "stream", None)
        if st is not None and t.is_cuda:
            torch.cuda.current_stream(t.device).wait_stream(st)       # the passes and the exchange run on the solver's stream
        if self.symm is not None and self.symm.sums() is not None:           # shared prefix from the exchanged buffer, the rest from the local sums
            t = torch.cat([self.symm.sums()[: self.n_exchange], t[self.n_exchange:]])
        a = t.detach().cpu().numpy().reshape(-1, 2)
        return a[self.new_of_old].reshape(-1).copy()

    def lower_bound(self) -> float:
        """Sum of the shards' lower bounds (the local value is already on the host: lower_bound() synchronises)."""
        lb = torch.tensor([self.local.lower_bound()], dtype=torch.float64)
        if self.world > 1:
            dev = getattr(self.local, "device", None)
            if dev is not None and dist.get_backend(self.group) == "nccl":
                lb = lb.to(dev)
            dist.all_reduce(lb, op=dist.ReduceOp.SUM, group=self.group)     # on the current stream; .item() waits for it
        return float(lb.item())


def make_cuda_native(precision: str, device: int, deterministic: bool = False) -> dict:
    """``make_local`` argument that selects the native path of ``sharded_mma``: the library plans the shard and builds its solver
    (``bddb200_create_shard``); nothing but the collection itself is prepared in Python."""
    return {"precision": precision, "device": device, "deterministic": deterministic}


def make_cuda_local(precision: str, device: int, deterministic: bool = False):
    """Factory for the GPU local solver: bdd_cuda_parallel_mma in shard mode.  Collectives and the
    exchange kernel are issued on the solver's own stream, so pass and exchange stay ordered without
    host synchronisation."""
    from .solver import bdd_cuda_parallel_mma

    def make(col: BddCollection, costs: np.ndarray, nr_vars: int, counts: np.ndarray, n_shared: int = 0):
        return bdd_cuda_parallel_mma(col, costs, precision=precision, device=device, deterministic=deterministic,
                                     nr_variables=nr_vars, nr_bdds_per_var=counts, n_shared_vars=n_shared)

    return make
