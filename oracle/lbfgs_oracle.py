"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the L-BFGS wrapper around the CPU oracle solver.

Follows src/bdd_solver/lbfgs_impl.h of the reference (iteration :138-157, store_iterate :46-135,
compute_update_direction :226-316, search_step_size_and_apply :159-224, lbfgs_update_possible :335-341)
on top of oracle/bindings.Oracle (net_solver_costs, bdds_solution, make_dual_feasible, gradient_step).
PARITY UNPINNED: the reference template is broken at this commit on both back ends (every CUDA branch is
compiled out by `#ifdef CUDACC`, and the first loop's alpha is shadowed, :251-263), and no
reference test covers it.  This restatement and bdd_b200/csrc/lbfgs.cuh implement the algorithm that
code spells out with those two defects removed; the GPU path is checked against this file, and both
against solver-independent properties (monotone bound, dual feasibility)."""
from collections import deque

import numpy as np


class LbfgsOracle:
    def __init__(self, oracle, history_size=5, init_step_size=1e-6, req_rel_lb_increase=1e-6,
                 step_size_decrease_factor=0.8, step_size_increase_factor=1.1):
        assert history_size > 1
        self.o = oracle
        self.m = history_size
        self.step_size = init_step_size
        self.req = req_rel_lb_increase
        self.dec, self.inc = step_size_decrease_factor, step_size_increase_factor
        self.history = deque()            # (s, y, rho_inv)
        self.prev_x = self.prev_g = None
        self.prev_stored = False
        self.unsuccessful = 0
        self.lb_history = []
        self.lbfgs_iterations = self.mma_iterations = 0

    def lower_bound(self):
        return self.o.lower_bound()

    def store_iterate(self, g):
        x = self.o.net_solver_costs().astype(np.float64)
        if not self.prev_stored:
            self.prev_x, self.prev_g, self.prev_stored = x, g.copy(), True
            return
        s = x - self.prev_x
        y = (self.prev_g.astype(np.int64) - g.astype(np.int64)).astype(np.float64)
        rho_inv = float(np.dot(s, y))
        if rho_inv > 1e-8:
            self.history.append((s, y, rho_inv))
            if len(self.history) > self.m:
                self.history.popleft()
        else:
            self.prev_stored = False
        self.prev_x, self.prev_g = x, g.copy()

    def compute_update_direction(self, g):
        d = g.astype(np.float64)
        alphas = []
        for s, y, rho_inv in reversed(self.history):
            alpha = float(np.dot(s, d)) / rho_inv
            alphas.append(alpha)
            d = d - alpha * y
        alphas.reverse()
        s_l, y_l, rho_l = self.history[-1]
        h0 = rho_l / (1e-8 + float(np.dot(y_l, y_l)))
        for i, (s, y, rho_inv) in enumerate(self.history):
            rho = 1.0 / rho_inv
            if i == 0:
                rho *= h0
            beta = rho * float(np.dot(y, d))
            d = d + (alphas[i] - beta) * s
        return d

    def search_step_size_and_apply(self, d):
        lb_pre = self.o.lower_bound()
        past = self.lb_history[-(self.m - 1)] - self.lb_history[-self.m]
        prev_step = [0.0]

        def apply(new_step):
            net = new_step - prev_step[0]
            if net != 0.0:
                self.o.gradient_step(d.astype(self.o.dtype), net)
            prev_step[0] = new_step

        num_updates, cur, best_step, best_rel = 0, 0.0, 0.0, 0.0
        while True:
            apply(self.step_size)
            cur = (self.o.lower_bound() - lb_pre) / (1e-9 + past)
            if best_rel < cur:
                best_rel, best_step = cur, self.step_size
            if cur <= 0.0:
                self.step_size *= self.dec
            elif cur < self.req:
                self.step_size *= self.inc
            if num_updates > 5:
                if best_rel > self.req / 10.0:
                    apply(best_step)
                else:
                    apply(0.0)
                    self.unsuccessful += 1
                return
            num_updates += 1
            if not cur < self.req:
                break
        if num_updates == 1 and self.unsuccessful == 0:
            self.step_size *= self.inc
        self.unsuccessful = 0

    def iteration(self):
        if not self.lb_history:
            self.lb_history.append(self.o.lower_bound())
        g = self.o.bdds_solution()
        self.store_iterate(g)
        if len(self.history) >= self.m and self.unsuccessful <= 5:
            d = self.compute_update_direction(g)
            d = self.o.make_dual_feasible(d.astype(self.o.dtype)).astype(np.float64)
            self.search_step_size_and_apply(d)
            self.o.iteration()
            self.lbfgs_iterations += 1
        else:
            self.o.iteration()
            self.mma_iterations += 1
        self.lb_history.append(self.o.lower_bound())
