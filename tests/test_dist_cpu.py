"""World-size-2 (and 3) gloo tests of the constraint-sharding logic (bdd_b200/dist.py) on CPU.
The local solver is a stand-in built on the CPU oracle in shard mode (test infrastructure);
on the GPU box the same sharded_mma class drives bdd_cuda_parallel_mma.  Mirrors
test/test_hybrid_parallel_mma_base.cu:14-167 (split solver == whole solver, pass by pass)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleLocal:
    """forward_pass / backward_pass / delta_sum_view / lower_bound on top of the oracle."""

    def __init__(self, col, costs, nr_vars, counts, precision):
        import bindings as B
        self.o = B.Oracle(col.instrs, col.delims, None, precision)
        self.o.set_shard(nr_vars, counts)
        self.o.update_costs(None, costs)
        self.counts = np.maximum(counts, 1).astype(self.o.dtype)
        self.sums = torch.zeros(2 * nr_vars, dtype=torch.float64 if precision == "double" else torch.float32)

    def _normalized(self):
        d = self.sums.numpy().copy()
        d[0::2] /= self.counts
        d[1::2] /= self.counts
        return d

    def forward_pass(self, omega):
        d = self._normalized()
        self.o.forward_mm(omega, d)
        self.sums.copy_(torch.from_numpy(d))

    def backward_pass(self, omega):
        d = self._normalized()
        self.o.backward_mm(omega, d)
        self.sums.copy_(torch.from_numpy(d))

    def delta_sum_view(self):
        return self.sums

    def lower_bound(self):
        return self.o.lower_bound()


def _worker(rank, world, port, gen, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bindings as B
    from bdd_b200 import instances
    from bdd_b200.dist import sharded_mma
    B.oracle_set_num_threads(1)
    col, costs = {"cover": lambda: instances.set_cover(m=200, n=300, k=7, seed=2),
                  "mrf": lambda: instances.grid_mrf(5, 4, 3, seed=1)}[gen]()
    s = sharded_mma(col, costs, rank, world, lambda c, cs, nv, cnt: OracleLocal(c, cs, nv, cnt, "double"))
    lbs = [s.lower_bound()]
    for _ in range(12):
        s.iteration()
        lbs.append(s.lower_bound())
    if rank == 0:
        from bdd_b200.instances import BOTSINK
        local_new = np.unique(s.local_col.instrs[s.local_col.instrs[:, 2] < BOTSINK, 2].astype(np.int64))
        old_of_new = np.argsort(s.new_of_old)
        q.put((lbs, s.delta_sums(), s.n_shared, old_of_new[local_new]))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("gen", ["cover", "mrf"])
def test_sharded_equals_single(gen, world):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bindings as B
    from bdd_b200 import instances
    col, costs = {"cover": lambda: instances.set_cover(m=200, n=300, k=7, seed=2),
                  "mrf": lambda: instances.grid_mrf(5, 4, 3, seed=1)}[gen]()
    B.oracle_set_num_threads(1)
    o = B.Oracle(col.instrs, col.delims, costs, "double")
    want = [o.lower_bound()]
    for _ in range(12):
        o.iteration()
        want.append(o.lower_bound())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, gen, q)) for r in range(world)]
    for p in procs:
        p.start()
    lbs, sums, n_shared, local_vars = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.allclose(lbs, want, rtol=0, atol=1e-9)
    # the all-reduced sums, normalised, equal the single solver's delta vector
    from bdd_b200.dist import global_nr_bdds_per_var
    cnt = np.maximum(global_nr_bdds_per_var(col), 1)
    d = sums.copy(); d[0::2] /= cnt; d[1::2] /= cnt
    # (rank 0 knows the sums of the variables of its own shard; the shared ones among them were exchanged)
    sel = np.stack([2 * local_vars, 2 * local_vars + 1], axis=1).reshape(-1)
    assert np.allclose(d[sel], o.get_delta()[sel], rtol=0, atol=1e-9)
    # only variables that occur in more than one shard are exchanged: a thin boundary set for the tiled grid
    assert 0 < n_shared <= col.nr_variables()
    if gen == "mrf" and world == 2:
        assert n_shared < col.nr_variables() // 2


def test_partition_and_counts():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bindings as B
    from bdd_b200 import instances
    from bdd_b200.dist import global_nr_bdds_per_var, partition_bdds
    col, costs = instances.random_inequalities(90, 50, seed=4)
    parts = partition_bdds(col, 4)
    assert np.array_equal(np.concatenate(parts), np.arange(col.nr_bdds))
    sizes = [int(np.diff(col.delims.astype(np.int64))[p].sum()) for p in parts]
    assert max(sizes) - min(sizes) <= 2 * int(np.diff(col.delims.astype(np.int64)).max())
    o = B.Oracle(col.instrs, col.delims, costs, "double")
    cnt = global_nr_bdds_per_var(col)
    assert np.array_equal(cnt, [o.nr_bdds_of_var(v) for v in range(o.n_vars)])
    # shards see the same structure as select() of the reference's remove()
    sub = col.select(parts[1])
    so = B.Oracle(sub.instrs, sub.delims, None, "double")
    assert so.n_bdds == len(parts[1])


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_native_shard_plan_equals_the_numpy_one(world):
    """bddb200_plan_shard (bdd_b200/csrc/shard.hpp, what bddb200_create_shard builds a rank's solver from) against partition_bdds /
    shared_first_relabeling / global_nr_bdds_per_var on four instance shapes; host only."""
    from bdd_b200 import dist as bdist, instances
    from bdd_b200.dist import _layer_heads
    cases = [instances.set_cover(m=300, n=500, k=7, seed=3), instances.grid_mrf(9, 7, 3, seed=6), instances.qap(n=5, seed=7), instances.assignment(12, seed=2)]
    for col, costs in cases:
        n_vars = col.nr_variables()
        parts = bdist.partition_bdds(col, world)
        new_of_old, n_shared = bdist.shared_first_relabeling(col, parts, n_vars)
        counts = bdist.global_nr_bdds_per_var(col, n_vars)
        for rank in range(world):
            info, noo, counts_new = bdist.plan_shard_native(col, world, rank)
            assert info["nr_variables"] == n_vars and info["n_shared"] == n_shared
            assert (info["first_bdd"], info["n_bdds"]) == ((int(parts[rank][0]) if len(parts[rank]) else info["first_bdd"]), len(parts[rank]))
            assert np.array_equal(noo, new_of_old)
            want = np.empty_like(counts); want[new_of_old] = counts
            assert np.array_equal(counts_new, want)
            local = bdist.relabel_variables(col.select(parts[rank]), new_of_old) if len(parts[rank]) else None
            entries = int((_layer_heads(local)[0] < n_shared).sum()) if local is not None else 0
            assert info["shared_entries"] == entries
            assert np.array_equal(info["share_mask"], bdist.share_masks(col, parts, new_of_old, n_shared))
            assert all(bin(int(m)).count("1") >= 2 for m in info["share_mask"])
    with pytest.raises(Exception):
        bdist.plan_shard_native(cases[0][0], 2, 2)
