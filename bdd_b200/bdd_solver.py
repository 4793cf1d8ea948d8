"""``bdd_solver`` driver (include/bdd_solver/bdd_solver.h:45-103, src/bdd_solver/bdd_solver.cpp) for the GPU paths:
the reference's JSON configuration drives read_ILP -> transform_to_BDDs -> construct_solver -> solve_dual ->
perturbation_rounding on the B200-native solver.  Same keys and defaults as the reference:

    {"input": "<file.lp or LP text>",                       bdd_solver.cpp:44-66
     "precision": "float" | "double",                       :145-147 (default double)
     "relaxation solver": "cuda parallel mma" | "lbfgs cuda mma" | "cuda lbfgs parallel mma" | "lbfgs cuda parallel mma",
     "split bdds": {"split length": n},                                                        :105-123, bdd_preprocessor.cpp:372-415
     "lbfgs": {"history size", "initial step size", "required relative lb increase",
               "step size decrease factor", "step size increase factor"},                       :177-199
     "termination criteria": {"maximum iterations": 1000, "minimum improvement": 1e-6,
                              "improvement slope": 1e-9, "time limit": 3600},                   :277-309
     "perturbation rounding": {"initial perturbation": 0.1, "perturbation growth rate": 1.1,
                               "inner iterations": 100, "outer iterations": 100}}                :318-380

Deviations from the reference, all documented in SURVEY 3.1: ``precision`` means what it says (the reference builds the
float solver for "double" and vice versa, bdd_solver.cpp:167-174); the README spelling "lbfgs cuda parallel mma"
(README.md:56), which matches none of the reference's strings and throws there (:237), is accepted; GPU rounding works
for every GPU solver (the reference's type list omits cuda parallel mma double, :353-356).  CPU solvers
("sequential mma", "parallel mma", ...), variable reordering, constraint normalisation and the key "export bdd graph" belong to subsystems
outside this build's scope and raise.  ``"split bdds": {"implication bdd": true}`` (also spelled ``"implication"``, the key
the reference reads, bdd_solver.cpp:119) adds the splitter's implication BDDs through the library's host-side collection.
"""
from __future__ import annotations

import json
import os
import sys
import time
from typing import Optional

import numpy as np

from . import instances, lp

CUDA_MMA = ("cuda parallel mma",)
CUDA_LBFGS = ("lbfgs cuda mma", "cuda lbfgs parallel mma", "lbfgs cuda parallel mma")


def read_config(c: str) -> dict:
    """bdd_solver::read_config (bdd_solver.cpp:468-475): a file name or an inline JSON string."""
    if os.path.exists(c):
        with open(c) as f:
            return json.load(f)
    return json.loads(c)


class bdd_solver:
    def __init__(self, config=None, log=None):
        self.log = log if log is not None else (lambda *a: print(*a, file=sys.stderr))
        self.ilp: Optional[lp.ILP] = None
        self.bdd_col = None
        self.costs = None
        self.solver = None
        self.solution = None
        if config is not None:
            self.solve(config)

    # ---- read_ILP, bdd_solver.cpp:44-66 ----------------------------------------------------------------------------------
    def read_ILP(self, config: dict) -> lp.ILP:
        if "input" not in config:
            raise RuntimeError("no input specified")
        inp = config["input"]
        if os.path.exists(inp):
            self.log(f"[bdd_solver] Read input file {inp}")
            with open(inp) as f:
                return lp.parse_lp(f.read())
        self.log("[bdd_solver] Read input string")
        return lp.parse_lp(inp)

    def process_ILP(self, config: dict):
        if config.get("variable order", "input") != "input":
            raise RuntimeError("variable reordering is outside this build's scope (SURVEY 2, row 23)")
        if config.get("normalize constraints", False):
            raise RuntimeError("constraint normalisation is outside this build's scope")

    # ---- export_lp, :412-430 ---------------------------------------------------------------------------------------------------
    def export_lp(self, config: dict):
        if "export lp" not in config:
            return
        path = config["export lp"]
        extension = os.path.splitext(path)[1]
        if extension != ".lp":          # the reference also writes .opb and .mps
            raise RuntimeError(f"Cannot recognize file extension {extension} for exporting problem file")
        with open(path, "w") as f:
            f.write(lp.write_lp(self.ilp))

    # ---- transform_to_BDDs, :112-123 ------------------------------------------------------------------------------------------
    def transform_to_BDDs(self, config: dict):
        self.log("[bdd solver] Compute BDDs")
        col, costs = instances.from_ilp(self.ilp)
        if "split bdds" in config:
            sb = config["split bdds"] or {}
            # the reference tests the key "implication bdd" and then reads "implication" (bdd_solver.cpp:119); both spellings are taken here
            implication = bool(sb.get("implication bdd", False) or sb.get("implication", False))
            from .split import compute_split_length, split_long_bdds
            # no length given: a value that fills the GPU (the reference's rule, bdd_preprocessor.cpp:32-121, targets its hop-synchronous kernels)
            length = int(sb["split length"]) if "split length" in sb else compute_split_length(col)
            n_before = col.nr_bdds
            try:
                if implication:      # the library's collection builds the implication BDD over the auxiliary variables (bdd_collection.cpp:805-940)
                    from .collection import bdd_collection
                    c = bdd_collection(col)
                    c.split_long_bdds(length, len(costs), True)
                    col = c.export()
                else:
                    col, _ = split_long_bdds(col, length, nr_variables=len(costs))      # auxiliary variables carry no cost
            except ValueError as e:
                raise RuntimeError(f"split bdds: {e}") from e
            self.log(f"[bdd preprocessor] split BDDs longer than {length}: {n_before} -> {col.nr_bdds} BDDs")
        return col, costs

    # ---- construct_solver, :130-267 --------------------------------------------------------------------------------------------
    def construct_solver(self, config: dict):
        from .solver import bdd_cuda_parallel_mma, lbfgs_cuda_mma
        precision = config.get("precision", "double")
        if precision not in ("float", "double"):
            raise RuntimeError("precision must be float or double")
        kind = config.get("relaxation solver", "cuda parallel mma")
        kw = dict(precision=precision, device=int(config.get("device", 0)))
        if kind in CUDA_MMA:
            self.log(f"[bdd solver] construct cuda parallel mma solver with {precision} precision")
            return bdd_cuda_parallel_mma(self.bdd_col, self.costs, **kw)
        if kind in CUDA_LBFGS:
            l = config.get("lbfgs", {}) if kind != "cuda lbfgs parallel mma" else {}      # :251-264 takes the defaults only
            self.log(f"[bdd solver] construct lbfgs cuda parallel mma solver with {precision} precision")
            return lbfgs_cuda_mma(self.bdd_col, self.costs,
                                  history_size=int(l.get("history size", 5)), init_step_size=float(l.get("initial step size", 1e-6)),
                                  req_rel_lb_increase=float(l.get("required relative lb increase", 1e-6)),
                                  step_size_decrease_factor=float(l.get("step size decrease factor", 0.8)),
                                  step_size_increase_factor=float(l.get("step size increase factor", 1.1)), **kw)
        raise RuntimeError(f"solver {kind} unknown (this build provides the GPU solvers: {CUDA_MMA + CUDA_LBFGS})")

    # ---- solve_dual, :277-309 --------------------------------------------------------------------------------------------------
    def solve_dual(self, config: dict) -> float:
        tc = config.get("termination criteria", {})
        max_iter = int(tc.get("maximum iterations", 1000))
        min_improvement = float(tc.get("minimum improvement", 1e-6))
        improvement_slope = float(tc.get("improvement slope", 1e-9))
        time_limit = float(tc.get("time limit", 3600))
        t0 = time.monotonic()
        lb = self.solver.run_solver(max_iter, min_improvement, improvement_slope, time_limit)
        self.log(f"[bdd solver] Terminated dual optimization: lower bound = {lb}, time = {time.monotonic() - t0:.3f} s")
        return lb

    # ---- perturbation_rounding, :318-380 ----------------------------------------------------------------------------------------
    def perturbation_rounding(self, config: dict):
        if "perturbation rounding" not in config:
            return None
        pr = config["perturbation rounding"] or {}
        sol, rounds = self.solver.incremental_mm_agreement_rounding(
            init_delta=float(pr.get("initial perturbation", 0.1)), delta_growth_rate=float(pr.get("perturbation growth rate", 1.1)),
            num_itr_lb=int(pr.get("inner iterations", 100)), num_rounds=int(pr.get("outer iterations", 100)))
        obj = float("inf")
        if sol is not None:
            n = self.ilp.nr_variables()
            obj = float(np.dot(self.costs[:n], sol[:n])) + getattr(self.ilp, "constant", 0.0)
        self.log(f"[incremental primal rounding] solution objective = {obj} after {rounds} rounds")
        return sol

    # ---- solve, :477-495 -----------------------------------------------------------------------------------------------------------
    def solve(self, config):
        if isinstance(config, str):
            config = read_config(config)
        if "export bdd graph" in config:
            raise RuntimeError("'export bdd graph' is outside this build's scope")
        if self.ilp is None:
            self.ilp = self.read_ILP(config)
            self.process_ILP(config)
            self.export_lp(config)
            self.bdd_col, self.costs = self.transform_to_BDDs(config)
            if "export bdd lp" in config:        # export_bdd_lp, :400-410: the relaxation as a linear programme over arc-flow variables
                from .collection import bdd_collection
                bdd_collection(self.bdd_col).write_bdd_lp(config["export bdd lp"], self.costs)
            if "print statistics" in config:
                self.log(f"[print_statistics] #variables = {self.ilp.nr_variables()}, #constraints = {len(self.ilp.constraints)}, #BDDs = {self.bdd_col.nr_bdds}")
            self.solver = self.construct_solver(config)
        self.solve_dual(config)
        self.solution = self.perturbation_rounding(config)
        return self

    def lower_bound(self) -> float:
        return self.solver.lower_bound()

    def min_marginals(self):
        """bdd_solver::min_marginals (:497-513): per variable the (mm_0, mm_1) pairs of its BDDs."""
        return self.solver.min_marginals()


def main(argv=None):
    """bdd_solver_cl (src/bdd_solver/bdd_solver_cl.cpp:3-10): ``python -m bdd_b200.bdd_solver config.json``"""
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("usage: python -m bdd_b200.bdd_solver <config.json | inline json>", file=sys.stderr)
        return 2
    s = bdd_solver(read_config(argv[0]))
    out = {"lower_bound": s.lower_bound()}
    if s.solution is not None:
        n = s.ilp.nr_variables()
        out["solution"] = {s.ilp.var_names[i]: int(s.solution[i]) for i in range(n)}
        out["objective"] = float(np.dot(s.costs[:n], s.solution[:n])) + getattr(s.ilp, "constant", 0.0)
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
