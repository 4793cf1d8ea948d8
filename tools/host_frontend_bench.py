"""Host-side front end, timed on the CPU (no GPU): .lp -> ILP -> BDD collection, long-BDD splitting with and without implication BDDs
(this build's collection class against the reference's own bdd_collection compiled into oracle/_ref, same inputs, outputs compared),
and the layout builder's phases on the three big benchmark shapes.

    python tools/host_frontend_bench.py            # prints a markdown report (profiles/r02_host_frontend.md was made with it)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))

import bindings as B  # noqa: E402
from bdd_b200 import _lib, instances  # noqa: E402
from bdd_b200.collection import bdd_collection  # noqa: E402
from bdd_b200.split import split_long_bdds  # noqa: E402


def timed(f, repeat=1):
    best = float("inf")
    for _ in range(repeat):
        t = time.perf_counter()
        out = f()
        best = min(best, time.perf_counter() - t)
    return out, best


def same(a, b_instrs, b_delims):
    inner = a.instrs[:, 2] < instances.BOTSINK
    return np.array_equal(a.delims, b_delims) and np.array_equal(a.instrs[inner], b_instrs[inner]) and np.array_equal(a.instrs[:, 2], b_instrs[:, 2])


def main():
    print("# Host-side front end on the CPU of the build container (%d cores), no GPU involved\n" % os.cpu_count())
    # ---- splitting: the 5 M-node assignment instance (BASELINE config 3b), chunk length 64
    col, costs = instances.assignment(1118, seed=3)
    print("## Long-BDD splitting, assignment 1118 x 1118 (%d BDDs, %.2f M nodes), chunk length 64\n" % (col.nr_bdds, col.nr_nodes / 1e6))
    print("| implementation | whole collection | per BDD |\n|---|---|---|")
    _, t = timed(lambda: split_long_bdds(col, 64))
    print("| numpy (`bdd_b200/split.py`), no implication BDD | %.2f s | %.2f ms |" % (t, 1e3 * t / col.nr_bdds))

    def native(implication):
        c = bdd_collection(col)
        c.split_long_bdds(64, len(costs), implication)
        return c
    for implication in (False, True):
        c, t = timed(lambda: native(implication))
        print("| C++ collection behind the C ABI (ctypes copy in included), implication BDD %s | %.2f s | %.2f ms | <!-- %d BDDs -->" % ("on" if implication else "off", t, 1e3 * t / col.nr_bdds, c.nr_bdds()))
    if B.ref_available():
        sub = col.select(list(range(40)))
        for implication in (False, True):
            rc = B.RefCollection.from_arrays(sub.instrs, sub.delims)
            mine = bdd_collection(sub)
            aux = [len(costs), len(costs)]

            def ref_run():
                for b in range(40):
                    _, aux[0] = (rc.split_qbdd_implication if implication else rc.split_qbdd)(b, 64, aux[0])

            def my_run():
                for b in range(40):
                    _, aux[1] = mine.split_qbdd(b, 64, aux[1], implication)
            _, tr = timed(ref_run)
            _, tm = timed(my_run)
            r_instrs, r_delims = rc.export()
            print("| 40 of these BDDs, implication BDD %s: reference `split_qbdd` (its own object code) / this build, outputs identical: %s | | %.2f ms / %.2f ms |"
                  % ("on" if implication else "off", same(mine.export(), r_instrs, r_delims), 1e3 * tr / 40, 1e3 * tm / 40))
    # ---- layout builder
    print("\n## Layout builder (`bdd_b200/csrc/layout.hpp`, OpenMP, %d threads here): `bddb200_layout_stats`, best of 3\n" % os.cpu_count())
    print("| instance | nodes | BDDs | build time |\n|---|---|---|---|")
    lib = _lib.load()
    for name, make in (("set_cover_1m", instances.set_cover), ("qap_5m", lambda: instances.qap(n=40, seed=2)), ("grid_mrf_20m", lambda: instances.grid_mrf(283, 283, 4, seed=4))):     # bench.py's workloads
        col, _ = make()
        instrs, delims = np.ascontiguousarray(col.instrs), np.ascontiguousarray(col.delims)
        out = np.zeros(13, dtype=np.uint64)
        _, t = timed(lambda: lib.bddb200_layout_stats(instrs.ctypes.data, instrs.shape[0], delims.ctypes.data, delims.shape[0] - 1, 0, out.ctypes.data, 13), repeat=3)
        print("| %s | %.2f M | %d | %.0f ms |" % (name, col.nr_nodes / 1e6, col.nr_bdds, 1e3 * t))


if __name__ == "__main__":
    main()
