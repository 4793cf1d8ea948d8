"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* :class:`Oracle`     -- oracle/liboracle_mma.so, the plain-C restatement (mma_oracle.c).
* :class:`RefCollection`, :class:`RefSolver` -- oracle/_ref/libbdd_ref.so, the reference's own
  CPU sources compiled here (ref_wrap.cpp).  Available only where that library was built.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module;
nothing under bdd_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_PATH = os.path.join(_HERE, "liboracle_mma.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libbdd_ref.so")

_szp = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")


def oracle_available() -> bool:
    return os.path.exists(_ORACLE_PATH)


def ref_available() -> bool:
    return os.path.exists(_REF_PATH)


_oracle_lib = None


def _oracle():
    global _oracle_lib
    if _oracle_lib is None:
        if not oracle_available():
            raise RuntimeError("oracle/liboracle_mma.so missing: run `make -C oracle` (or __graft_entry__.build())")
        _oracle_lib = C.CDLL(_ORACLE_PATH)
    return _oracle_lib


class Oracle:
    """Plain-C CPU restatement of the deferred MMA sweep (one object = one solver)."""

    def __init__(self, instrs: np.ndarray, delims: np.ndarray, costs: Optional[np.ndarray] = None,
                 precision: str = "double", inf_mode: int = 0):
        lib = _oracle()
        self.suf = "_f64" if precision == "double" else "_f32"
        self.dtype = np.float64 if precision == "double" else np.float32
        self._rp = _f64p if precision == "double" else _f32p
        self._real = C.c_double if precision == "double" else C.c_float
        instrs = np.ascontiguousarray(instrs, dtype=np.uint64)
        delims = np.ascontiguousarray(delims, dtype=np.uint64)
        f = self._fn("oracle_create", C.c_void_p, [_szp, C.c_size_t, _szp, C.c_size_t])
        self.h = f(instrs.reshape(-1), instrs.shape[0], delims, delims.shape[0] - 1)
        if not self.h:
            raise RuntimeError("oracle_create failed")
        self._fn("oracle_set_inf_mode", None, [C.c_void_p, C.c_int])(self.h, inf_mode)
        self.n_vars = self._fn("oracle_nr_variables", C.c_size_t, [C.c_void_p])(self.h)
        self.n_bdds = self._fn("oracle_nr_bdds", C.c_size_t, [C.c_void_p])(self.h)
        self.n_layers = self._fn("oracle_nr_layers", C.c_size_t, [C.c_void_p])(self.h)
        self.n_nodes = self._fn("oracle_nr_nodes", C.c_size_t, [C.c_void_p])(self.h)
        if costs is not None:
            self.update_costs(None, costs)

    def _fn(self, name, restype, argtypes):
        f = getattr(_oracle(), name + self.suf)
        f.restype = restype
        f.argtypes = argtypes
        return f

    def __del__(self):
        if getattr(self, "h", None):
            self._fn("oracle_destroy", None, [C.c_void_p])(self.h)
            self.h = None

    def nr_bdds_of_var(self, v: int) -> int:
        return self._fn("oracle_nr_bdds_of_var", C.c_size_t, [C.c_void_p, C.c_size_t])(self.h, v)

    def set_shard(self, n_vars: int, counts: np.ndarray):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        assert counts.shape[0] == n_vars
        self._fn("oracle_set_shard", None, [C.c_void_p, C.c_size_t, _szp])(self.h, n_vars, counts)
        self.n_vars = n_vars

    def layer_vars(self) -> np.ndarray:
        out = np.empty(self.n_layers, dtype=np.uint64)
        self._fn("oracle_layer_vars", None, [C.c_void_p, _szp])(self.h, out)
        return out

    def bdd_layer_begin(self) -> np.ndarray:
        out = np.empty(self.n_bdds + 1, dtype=np.uint64)
        self._fn("oracle_bdd_layer_begin", None, [C.c_void_p, _szp])(self.h, out)
        return out

    def update_costs(self, lo, hi):
        lo = np.zeros(0) if lo is None else np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.zeros(0) if hi is None else np.ascontiguousarray(hi, dtype=np.float64)
        self._fn("oracle_update_costs", None, [C.c_void_p, _f64p, C.c_size_t, _f64p, C.c_size_t])(
            self.h, lo, lo.shape[0], hi, hi.shape[0])

    def set_cost(self, c: float, var: int):
        self._fn("oracle_set_cost", None, [C.c_void_p, C.c_double, C.c_size_t])(self.h, c, var)

    def lower_bound(self) -> float:
        return self._fn("oracle_lower_bound", C.c_double, [C.c_void_p])(self.h)

    def backward_run(self):
        self._fn("oracle_backward_run", None, [C.c_void_p])(self.h)

    def forward_run(self):
        self._fn("oracle_forward_run", None, [C.c_void_p])(self.h)

    def forward_mm(self, omega: float, delta: np.ndarray) -> np.ndarray:
        assert delta.dtype == self.dtype and delta.shape[0] == 2 * self.n_vars
        self._fn("oracle_forward_mm", None, [C.c_void_p, self._real, self._rp])(self.h, omega, delta)
        return delta

    def backward_mm(self, omega: float, delta: np.ndarray) -> float:
        assert delta.dtype == self.dtype and delta.shape[0] == 2 * self.n_vars
        return self._fn("oracle_backward_mm", C.c_double, [C.c_void_p, self._real, self._rp])(self.h, omega, delta)

    def normalize_delta(self, delta: np.ndarray):
        self._fn("oracle_normalize_delta", None, [C.c_void_p, self._rp])(self.h, delta)

    def iteration(self, omega: float = 0.5):
        self._fn("oracle_iteration", None, [C.c_void_p, self._real])(self.h, omega)

    def get_delta(self) -> np.ndarray:
        out = np.empty(2 * self.n_vars, dtype=self.dtype)
        self._fn("oracle_get_delta", None, [C.c_void_p, self._rp])(self.h, out)
        return out

    def distribute_delta(self):
        self._fn("oracle_distribute_delta", None, [C.c_void_p])(self.h)

    def distribute_delta_cpu(self):
        self._fn("oracle_distribute_delta_cpu", None, [C.c_void_p])(self.h)

    def min_marginals(self) -> np.ndarray:
        out = np.empty((self.n_layers, 2), dtype=self.dtype)
        self._fn("oracle_min_marginals", None, [C.c_void_p, self._rp])(self.h, out.reshape(-1))
        return out

    def get_costs(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        lo = np.empty(self.n_layers, dtype=self.dtype)
        hi = np.empty(self.n_layers, dtype=self.dtype)
        mm = np.empty(self.n_layers, dtype=self.dtype)
        self._fn("oracle_get_costs", None, [C.c_void_p, self._rp, self._rp, self._rp])(self.h, lo, hi, mm)
        return lo, hi, mm

    def set_costs(self, lo, hi, mm):
        self._fn("oracle_set_costs", None, [C.c_void_p, self._rp, self._rp, self._rp])(
            self.h, np.ascontiguousarray(lo, self.dtype), np.ascontiguousarray(hi, self.dtype), np.ascontiguousarray(mm, self.dtype))

    def primal_objective(self) -> np.ndarray:
        out = np.empty(self.n_vars, dtype=self.dtype)
        self._fn("oracle_primal_objective", None, [C.c_void_p, self._rp])(self.h, out)
        return out

    def net_solver_costs(self) -> np.ndarray:
        out = np.empty(self.n_layers, dtype=self.dtype)
        self._fn("oracle_net_solver_costs", None, [C.c_void_p, self._rp])(self.h, out)
        return out

    def bdds_solution(self) -> np.ndarray:
        out = np.empty(self.n_layers, dtype=np.int8)
        self._fn("oracle_bdds_solution", None, [C.c_void_p, _i8p])(self.h, out)
        return out

    def make_dual_feasible(self, d: np.ndarray) -> np.ndarray:
        d = np.ascontiguousarray(d, dtype=self.dtype)
        self._fn("oracle_make_dual_feasible", None, [C.c_void_p, self._rp])(self.h, d)
        return d

    def gradient_step(self, g: np.ndarray, step: float):
        self._fn("oracle_gradient_step", None, [C.c_void_p, self._rp, C.c_double])(
            self.h, np.ascontiguousarray(g, dtype=self.dtype), step)


def oracle_set_num_threads(n: int):
    _oracle().oracle_set_num_threads(int(n))


def oracle_max_threads() -> int:
    return int(_oracle().oracle_max_threads())


# ------------------------------------------------------------------ reference (oracle/_ref)

_ref_lib = None


def _ref():
    global _ref_lib
    if _ref_lib is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref/libbdd_ref.so missing (needs /root/reference: `make -C oracle ref`)")
        lib = C.CDLL(_REF_PATH)
        lib.refw_collection_new.restype = C.c_void_p
        lib.refw_collection_free.argtypes = [C.c_void_p]
        lib.refw_collection_error.restype = C.c_char_p
        lib.refw_collection_error.argtypes = [C.c_void_p]
        lib.refw_add_constraint.restype = C.c_long
        lib.refw_add_constraint.argtypes = [C.c_void_p, _i32p, _szp, C.c_size_t, C.c_int, C.c_int]
        lib.refw_nr_bdds.restype = C.c_size_t
        lib.refw_nr_bdds.argtypes = [C.c_void_p]
        lib.refw_nr_instructions.restype = C.c_size_t
        lib.refw_nr_instructions.argtypes = [C.c_void_p]
        lib.refw_export.argtypes = [C.c_void_p, _szp, _szp]
        lib.refw_collection_from_arrays.restype = C.c_void_p
        lib.refw_collection_from_arrays.argtypes = [_szp, C.c_size_t, _szp, C.c_size_t]
        lib.refw_set_num_threads.argtypes = [C.c_int]
        lib.refw_max_threads.restype = C.c_int
        lib.refw_solver_new.restype = C.c_void_p
        lib.refw_solver_new.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        lib.refw_solver_free.argtypes = [C.c_void_p]
        for name in ("nr_variables", "nr_bdds", "nr_layers"):
            f = getattr(lib, "refw_solver_" + name)
            f.restype = C.c_size_t
            f.argtypes = [C.c_void_p]
        lib.refw_solver_nr_bdds_of_var.restype = C.c_size_t
        lib.refw_solver_nr_bdds_of_var.argtypes = [C.c_void_p, C.c_size_t]
        lib.refw_solver_lower_bound.restype = C.c_double
        lib.refw_solver_lower_bound.argtypes = [C.c_void_p]
        lib.refw_solver_iteration.argtypes = [C.c_void_p]
        lib.refw_solver_distribute_delta.argtypes = [C.c_void_p]
        lib.refw_solver_forward_mm.argtypes = [C.c_void_p, C.c_double, _f64p]
        lib.refw_solver_backward_mm.argtypes = [C.c_void_p, C.c_double, _f64p]
        lib.refw_solver_min_marginals.restype = C.c_size_t
        lib.refw_solver_min_marginals.argtypes = [C.c_void_p, C.c_void_p]
        lib.refw_solver_bdds_solution.argtypes = [C.c_void_p, _i8p]
        lib.refw_solver_net_solver_costs.argtypes = [C.c_void_p, _f64p]
        _ref_lib = lib
    return _ref_lib


class RefCollection:
    """The reference's BDD::bdd_collection, filled either through the reference's own
    inequality -> BDD converter or from flat arrays."""

    def __init__(self, handle=None):
        self.lib = _ref()
        self.h = handle if handle is not None else self.lib.refw_collection_new()

    @classmethod
    def from_arrays(cls, instrs: np.ndarray, delims: np.ndarray) -> "RefCollection":
        lib = _ref()
        instrs = np.ascontiguousarray(instrs, dtype=np.uint64)
        delims = np.ascontiguousarray(delims, dtype=np.uint64)
        h = lib.refw_collection_from_arrays(instrs.reshape(-1), instrs.shape[0], delims, delims.shape[0] - 1)
        col = cls(h)
        err = lib.refw_collection_error(h)
        if err:
            raise RuntimeError(err.decode())
        return col

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.refw_collection_free(self.h)
            self.h = None

    def add_constraint(self, coeffs, variables, ineq: int, rhs: int) -> int:
        co = np.ascontiguousarray(coeffs, dtype=np.int32)
        va = np.ascontiguousarray(variables, dtype=np.uint64)
        r = self.lib.refw_add_constraint(self.h, co, va, co.shape[0], ineq, rhs)
        if r == -2:
            raise RuntimeError(self.lib.refw_collection_error(self.h).decode())
        return r

    def split_qbdd(self, bdd_nr: int, chunk_size: int, aux_var_start: int) -> Tuple[int, int]:
        """bdd_collection::split_qbdd without implication BDD; appends the chunks; returns (number of new BDDs, next aux variable)."""
        f = self.lib.refw_split_qbdd
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]
        n = C.c_size_t()
        nxt = f(self.h, bdd_nr, chunk_size, aux_var_start, C.byref(n))
        if nxt == 2 ** 64 - 1:
            raise RuntimeError(self.lib.refw_collection_error(self.h).decode())
        return n.value, nxt

    # the reference's direct generators and structural operations, one to one (ref_wrap.cpp)
    def _call(self, name: str, *args, restype=C.c_size_t):
        f = getattr(self.lib, name)
        f.restype = restype
        f.argtypes = [C.c_void_p] + [_szp if isinstance(a, np.ndarray) else C.c_size_t for a in args]
        return f(self.h, *args)

    def simplex_constraint(self, n: int) -> int:
        return self._call("refw_simplex_constraint", n)

    def not_all_false_constraint(self, n: int) -> int:
        return self._call("refw_not_all_false_constraint", n)

    def all_equal_constraint(self, n: int) -> int:
        return self._call("refw_all_equal_constraint", n)

    def cardinality_constraint(self, n: int, k: int) -> int:
        return self._call("refw_cardinality_constraint", n, k)

    def rebase(self, bdd_nr: int, variables) -> None:
        v = np.ascontiguousarray(variables, dtype=np.uint64)
        self._call("refw_rebase", bdd_nr, v, v.shape[0], restype=None)

    def negate(self, bdd_nr: int) -> None:
        self._call("refw_negate", bdd_nr, restype=None)

    def invert(self, bdd_nr: int, var: int) -> None:
        self._call("refw_invert", bdd_nr, var, restype=None)

    def reorder(self, bdd_nr: int) -> None:
        self._call("refw_reorder", bdd_nr, restype=None)

    def make_qbdd(self, bdd_nr: int) -> int:
        return self._call("refw_make_qbdd", bdd_nr)

    def bdd_and(self, bdd_nrs) -> int:
        v = np.ascontiguousarray(bdd_nrs, dtype=np.uint64)
        return self._call("refw_bdd_and", v, v.shape[0])

    def remove(self, bdd_nrs) -> None:
        v = np.ascontiguousarray(bdd_nrs, dtype=np.uint64)
        self._call("refw_remove", v, v.shape[0], restype=None)

    def is_qbdd(self, bdd_nr: int) -> bool:
        return bool(self._call("refw_is_qbdd", bdd_nr, restype=C.c_int))

    def is_reordered(self, bdd_nr: int) -> bool:
        return bool(self._call("refw_is_reordered", bdd_nr, restype=C.c_int))

    def variables(self, bdd_nr: int) -> np.ndarray:
        f = self.lib.refw_variables
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        out = np.empty(f(self.h, bdd_nr, None), dtype=np.uint64)
        f(self.h, bdd_nr, out.ctypes.data)
        return out

    def write_bdd_lp(self, costs) -> str:
        f = self.lib.refw_write_bdd_lp
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, _f64p, C.c_size_t, C.c_char_p, C.c_size_t]
        costs = np.ascontiguousarray(costs, dtype=np.float64)
        n = f(self.h, costs, costs.shape[0], None, 0)
        buf = C.create_string_buffer(n)
        f(self.h, costs, costs.shape[0], buf, n)
        return buf.raw[:n].decode()

    def split_qbdd_implication(self, bdd_nr: int, chunk_size: int, aux_var_start: int) -> Tuple[int, int]:
        """bdd_collection::split_qbdd WITH the implication BDD; returns (number of new BDDs, next aux variable)."""
        f = self.lib.refw_split_qbdd_implication
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]
        n = C.c_size_t()
        nxt = f(self.h, bdd_nr, chunk_size, aux_var_start, C.byref(n))
        if nxt == 2 ** 64 - 1:
            raise RuntimeError(self.lib.refw_collection_error(self.h).decode())
        return n.value, nxt

    def export(self) -> Tuple[np.ndarray, np.ndarray]:
        n = self.lib.refw_nr_instructions(self.h)
        b = self.lib.refw_nr_bdds(self.h)
        instrs = np.empty(3 * n, dtype=np.uint64)
        delims = np.empty(b + 1, dtype=np.uint64)
        self.lib.refw_export(self.h, instrs, delims)
        return instrs.reshape(n, 3), delims


class RefSolver:
    """The reference's CPU `parallel mma` solver (bdd_parallel_mma_base<bdd_branch_instruction<REAL,uint16_t>>)."""

    def __init__(self, col: RefCollection, costs: Optional[np.ndarray], precision: str = "double"):
        self.lib = _ref()
        self.col = col
        if costs is not None:
            costs = np.ascontiguousarray(costs, dtype=np.float64)
            self.h = self.lib.refw_solver_new(col.h, costs.ctypes.data_as(C.c_void_p), costs.shape[0], int(precision == "double"))
        else:
            self.h = self.lib.refw_solver_new(col.h, None, 0, int(precision == "double"))
        self.n_vars = self.lib.refw_solver_nr_variables(self.h)
        self.n_bdds = self.lib.refw_solver_nr_bdds(self.h)
        self.n_layers = self.lib.refw_solver_nr_layers(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.refw_solver_free(self.h)
            self.h = None

    def nr_bdds_of_var(self, v: int) -> int:
        return self.lib.refw_solver_nr_bdds_of_var(self.h, v)

    def lower_bound(self) -> float:
        return self.lib.refw_solver_lower_bound(self.h)

    def iteration(self):
        self.lib.refw_solver_iteration(self.h)

    def distribute_delta(self):
        self.lib.refw_solver_distribute_delta(self.h)

    def forward_mm(self, omega: float, delta: np.ndarray):
        self.lib.refw_solver_forward_mm(self.h, omega, delta)

    def backward_mm(self, omega: float, delta: np.ndarray):
        self.lib.refw_solver_backward_mm(self.h, omega, delta)

    def min_marginals(self) -> np.ndarray:
        n = self.lib.refw_solver_min_marginals(self.h, None)
        out = np.empty((n, 2), dtype=np.float64)
        self.lib.refw_solver_min_marginals(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def bdds_solution(self) -> np.ndarray:
        out = np.empty(self.n_layers, dtype=np.int8)
        self.lib.refw_solver_bdds_solution(self.h, out)
        return out

    def net_solver_costs(self) -> np.ndarray:
        out = np.empty(self.n_layers, dtype=np.float64)
        self.lib.refw_solver_net_solver_costs(self.h, out)
        return out


def ref_set_num_threads(n: int):
    _ref().refw_set_num_threads(int(n))


def ref_max_threads() -> int:
    return int(_ref().refw_max_threads())


def ref_mm_decode(mm_per_var):
    """The reference's CPU rounding decoder (mm_primal_decoder) on min-marginals given as a list of [k_v, 2] arrays:
    (types per variable: 0 zero / 1 one / 2 equal / 3 inconsistent, sums [V, 2], statistics {one, zero, equal, inconsistent}, solution or None)."""
    lib = _ref()
    lib.refw_mm_decode.restype = C.c_int
    lib.refw_mm_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    n = len(mm_per_var)
    counts = np.array([m.shape[0] for m in mm_per_var], dtype=np.uint64)
    flat = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.float64).reshape(-1, 2) for m in mm_per_var] + [np.zeros((0, 2))]), dtype=np.float64)
    types = np.zeros(n, dtype=np.int8); sums = np.zeros((n, 2), dtype=np.float64); stats = np.zeros(4, dtype=np.uint64); sol = np.zeros(n, dtype=np.int8)
    ok = lib.refw_mm_decode(flat.ctypes.data, counts.ctypes.data, n, types.ctypes.data, sums.ctypes.data, stats.ctypes.data, sol.ctypes.data)
    return types, sums, {"one": int(stats[0]), "zero": int(stats[1]), "equal": int(stats[2]), "inconsistent": int(stats[3])}, (sol if ok else None)


# ------------------------------------------------------------------ reference CUDA solver ---
_REF_CUDA_PATH = os.path.join(_HERE, "_ref", "libbdd_ref_cuda.so")
_ref_cuda_lib = None


def ref_cuda_available() -> bool:
    return os.path.exists(_REF_CUDA_PATH)


def _ref_cuda():
    global _ref_cuda_lib
    if _ref_cuda_lib is None:
        lib = C.CDLL(_REF_CUDA_PATH)
        lib.refcu_solver_new.restype = C.c_void_p
        lib.refcu_solver_new.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        lib.refcu_solver_free.argtypes = [C.c_void_p]
        lib.refcu_lower_bound.restype = C.c_double
        lib.refcu_lower_bound.argtypes = [C.c_void_p]
        lib.refcu_iterations.restype = C.c_double
        lib.refcu_iterations.argtypes = [C.c_void_p, C.c_size_t]
        lib.refcu_nr_hops.restype = C.c_size_t
        lib.refcu_nr_hops.argtypes = [C.c_void_p]
        lib.refcu_nr_bdd_nodes.restype = C.c_size_t
        lib.refcu_nr_bdd_nodes.argtypes = [C.c_void_p]
        lib.refcu_nr_layers.restype = C.c_size_t
        lib.refcu_nr_layers.argtypes = [C.c_void_p]
        lib.refcu_layer_indices.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.refcu_get_solver_costs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _ref_cuda_lib = lib
    return _ref_cuda_lib


class RefCudaSolver:
    """The reference's own `cuda parallel mma` (bdd_cuda_parallel_mma<REAL>, src/bdd_solver/bdd_cuda_parallel_mma.cu) compiled for
    sm_100a from the sources under /root/reference (oracle/Makefile, target ref_cuda).  A same-box GPU baseline and checker."""

    def __init__(self, instrs: np.ndarray, delims: np.ndarray, costs: np.ndarray, precision: str = "double"):
        self.lib = _ref_cuda()
        instrs = np.ascontiguousarray(instrs, dtype=np.uint64)
        delims = np.ascontiguousarray(delims, dtype=np.uint64)
        costs = np.ascontiguousarray(costs, dtype=np.float64)
        self.h = self.lib.refcu_solver_new(instrs.ctypes.data, instrs.shape[0], delims.ctypes.data, delims.shape[0] - 1,
                                           costs.ctypes.data, costs.shape[0], int(precision == "double"))
        if not self.h:
            raise RuntimeError("the reference CUDA solver could not be constructed")

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.refcu_solver_free(self.h)
            self.h = None

    def lower_bound(self) -> float:
        return self.lib.refcu_lower_bound(self.h)

    def iterations(self, n: int) -> float:
        """n iterations back to back; returns the seconds they took (device synchronised on both sides)."""
        return self.lib.refcu_iterations(self.h, int(n))

    def iteration(self):
        self.lib.refcu_iterations(self.h, 1)

    def nr_hops(self) -> int:
        return self.lib.refcu_nr_hops(self.h)

    def nr_layers(self) -> int:
        return self.lib.refcu_nr_layers(self.h)

    def layer_indices(self) -> Tuple[np.ndarray, np.ndarray]:
        """(get_primal_variable_index(), get_bdd_index()) in the reference's own layer order."""
        primal, bdd = np.empty(self.nr_layers(), dtype=np.int32), np.empty(self.nr_layers(), dtype=np.int32)
        self.lib.refcu_layer_indices(self.h, primal.ctypes.data, bdd.ctypes.data)
        return primal, bdd

    def get_solver_costs(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        out = [np.empty(self.nr_layers(), dtype=np.float64) for _ in range(3)]
        self.lib.refcu_get_solver_costs(self.h, *(o.ctypes.data for o in out))
        return tuple(out)
