"""ctypes mirror of the host-side BDD collection in libbdd_b200.so (include/bdd_b200_collection.h; the C++ class is
``bddb200_host::bdd_collection``, bdd_b200/csrc/host/bdd_collection.hpp).

Method names, argument meaning and results follow the reference's ``BDD::bdd_collection`` (include/bdd_collection/bdd_collection.h):
the direct generators ``simplex_constraint`` / ``not_all_false_constraint`` / ``all_equal_constraint`` / ``cardinality_constraint``
(src/bdd_collection/bdd_collection.cpp:2039-2263), ``rebase`` / ``negate`` / ``invert``, ``reorder`` (:1429), ``make_qbdd`` (:1670),
``bdd_and`` (:31-315), ``remove`` and ``split_qbdd`` with its optional implication BDD (:507-949).  Host code only: no GPU is needed.
Errors of the library surface as :class:`bdd_b200._lib.BddB200Error` (the reference asserts or throws ``std::runtime_error``).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .instances import BddCollection


class bdd_collection:
    def __init__(self, col: Optional[BddCollection] = None, _handle: Optional[C.c_void_p] = None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        elif col is None:
            _lib.check(self._lib.bddb200_collection_create(None, 0, None, 0, C.byref(self._h)))
        else:
            instrs = np.ascontiguousarray(col.instrs, dtype=np.uint64)
            delims = np.ascontiguousarray(col.delims, dtype=np.uint64)
            _lib.check(self._lib.bddb200_collection_create(instrs.ctypes.data, instrs.shape[0], delims.ctypes.data, delims.shape[0] - 1, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.bddb200_collection_destroy(self._h)
            self._h = C.c_void_p()

    # ------------------------------------------------------------------ sizes / export
    def _size(self, f, *args) -> int:
        out = C.c_size_t()
        _lib.check(f(self._h, *args, C.byref(out)))
        return out.value

    def nr_bdds(self) -> int:
        return self._size(self._lib.bddb200_collection_nr_bdds)

    def nr_bdd_nodes(self) -> int:
        return self._size(self._lib.bddb200_collection_nr_instructions)

    def export(self) -> BddCollection:
        """The instruction and delimiter arrays, the form every solver constructor of this package takes."""
        instrs = np.empty((self.nr_bdd_nodes(), 3), dtype=np.uint64)
        delims = np.empty(self.nr_bdds() + 1, dtype=np.uint64)
        _lib.check(self._lib.bddb200_collection_export(self._h, instrs.ctypes.data, delims.ctypes.data))
        return BddCollection(instrs, delims)

    # ------------------------------------------------------------------ generators (variables 0 .. n-1; rebase afterwards)
    def simplex_constraint(self, n: int) -> int:
        return self._size(self._lib.bddb200_collection_simplex_constraint, n)

    def not_all_false_constraint(self, n: int) -> int:
        return self._size(self._lib.bddb200_collection_not_all_false_constraint, n)

    def all_equal_constraint(self, n: int) -> int:
        return self._size(self._lib.bddb200_collection_all_equal_constraint, n)

    def cardinality_constraint(self, n: int, k: int) -> int:
        return self._size(self._lib.bddb200_collection_cardinality_constraint, n, k)

    def add_linear_constraint(self, coefficients: Sequence[int], variables: Sequence[int], ineq: int, rhs: int) -> Optional[int]:
        """The quasi-reduced BDD of ``sum coefficients[k] * x[variables[k]]  {0 '<=', 1 '>=', 2 '='}  rhs`` (variables ascending), built
        directly (bdd_preprocessor.cpp:175-228 without the BDD manager).  ``None`` when the constraint is always satisfied."""
        co = np.ascontiguousarray(coefficients, dtype=np.int64)
        va = np.ascontiguousarray(variables, dtype=np.uint64)
        if co.shape != va.shape:
            raise ValueError("one coefficient per variable")
        nr = self._size(self._lib.bddb200_collection_add_linear_constraint, co.ctypes.data, va.ctypes.data, co.shape[0], int(ineq), int(rhs))
        return None if nr == 2 ** 64 - 1 else nr

    # ------------------------------------------------------------------ relabelling
    def rebase(self, bdd_nr: int, variables: Sequence[int]) -> None:
        v = np.ascontiguousarray(variables, dtype=np.uint64)
        _lib.check(self._lib.bddb200_collection_rebase(self._h, bdd_nr, v.ctypes.data, v.shape[0]))

    def negate(self, bdd_nr: int) -> None:
        _lib.check(self._lib.bddb200_collection_negate(self._h, bdd_nr))

    def invert(self, bdd_nr: int, var: int) -> None:
        _lib.check(self._lib.bddb200_collection_invert(self._h, bdd_nr, var))

    # ------------------------------------------------------------------ structure
    def variables(self, bdd_nr: int) -> np.ndarray:
        n = C.c_size_t()
        _lib.check(self._lib.bddb200_collection_variables(self._h, bdd_nr, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.uint64)
        _lib.check(self._lib.bddb200_collection_variables(self._h, bdd_nr, out.ctypes.data, out.shape[0], C.byref(n)))
        return out

    def _flag(self, f, *args) -> bool:
        out = C.c_int()
        _lib.check(f(self._h, *args, C.byref(out)))
        return bool(out.value)

    def is_qbdd(self, bdd_nr: int) -> bool:
        return self._flag(self._lib.bddb200_collection_is_qbdd, bdd_nr)

    def is_reordered(self, bdd_nr: int) -> bool:
        return self._flag(self._lib.bddb200_collection_is_reordered, bdd_nr)

    def evaluate(self, bdd_nr: int, labeling: Sequence[int]) -> bool:
        x = np.ascontiguousarray(labeling, dtype=np.int8)
        return self._flag(self._lib.bddb200_collection_evaluate, bdd_nr, x.ctypes.data, x.shape[0])

    def reorder(self, bdd_nr: int) -> None:
        _lib.check(self._lib.bddb200_collection_reorder(self._h, bdd_nr))

    def make_qbdd(self, bdd_nr: int) -> int:
        return self._size(self._lib.bddb200_collection_make_qbdd, bdd_nr)

    def bdd_and(self, bdd_nrs: Iterable[int]) -> int:
        v = np.ascontiguousarray(list(bdd_nrs), dtype=np.uint64)
        return self._size(self._lib.bddb200_collection_bdd_and, v.ctypes.data, v.shape[0])

    def remove(self, bdd_nrs) -> None:
        v = np.ascontiguousarray([bdd_nrs] if np.isscalar(bdd_nrs) else list(bdd_nrs), dtype=np.uint64)
        _lib.check(self._lib.bddb200_collection_remove(self._h, v.ctypes.data, v.shape[0]))

    def write_bdd_lp(self, path: str, costs: Sequence[float]) -> None:
        """``bdd_collection::write_bdd_lp``: the relaxation the dual solvers work on as an ``.lp`` file (arc-flow variables per BDD linked
        through the shared variables ``x_<var>``); its optimum is the best bound the BDD decomposition can give."""
        c = np.ascontiguousarray(costs, dtype=np.float64)
        _lib.check(self._lib.bddb200_collection_write_bdd_lp(self._h, c.ctypes.data, c.shape[0], str(path).encode()))

    # ------------------------------------------------------------------ splitting
    def split_qbdd(self, bdd_nr: int, chunk_size: int, aux_var_start: int, with_implication_bdd: bool = False) -> Tuple[List[int], int]:
        """``bdd_collection::split_qbdd``: returns (numbers of the new BDDs -- ``[bdd_nr]`` when nothing was cut --, next free auxiliary variable)."""
        before = self.nr_bdds()
        n_new, nxt = C.c_size_t(), C.c_size_t()
        _lib.check(self._lib.bddb200_collection_split_qbdd(self._h, bdd_nr, chunk_size, aux_var_start, int(with_implication_bdd), C.byref(n_new), C.byref(nxt)))
        return ([bdd_nr] if n_new.value == 0 else list(range(before, before + n_new.value))), nxt.value

    def split_long_bdds(self, split_length: int, nr_variables: int = 0, with_implication_bdd: bool = False) -> Tuple[int, int]:
        """The preprocessor's loop with a forced split length (bdd_preprocessor.cpp:372-415), in place.
        Returns (number of BDDs that were cut, number of variables including the auxiliary ones)."""
        n_split, n_vars = C.c_size_t(), C.c_size_t()
        _lib.check(self._lib.bddb200_collection_split_long_bdds(self._h, split_length, nr_variables, int(with_implication_bdd), C.byref(n_split), C.byref(n_vars)))
        return n_split.value, n_vars.value


class ilp_input:
    """The library's .lp reader (bdd_b200/csrc/host/lp_reader.hpp behind ``bddb200_ilp_*``): ``ILP_input`` as far as the drivers need
    it (src/ILP/ILP_parser.cpp:25-160; ``bdd_solver::read_ILP``, src/bdd_solver/bdd_solver.cpp:44-66).  ``file_or_text`` is the name of a
    readable ``.lp`` file, else the LP text itself."""

    def __init__(self, file_or_text: str):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.bddb200_ilp_read(file_or_text.encode(), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.bddb200_ilp_destroy(self._h)
            self._h = C.c_void_p()

    def _size(self, f) -> int:
        out = C.c_size_t()
        _lib.check(f(self._h, C.byref(out)))
        return out.value

    def nr_variables(self) -> int:
        return self._size(self._lib.bddb200_ilp_nr_variables)

    def nr_constraints(self) -> int:
        return self._size(self._lib.bddb200_ilp_nr_constraints)

    def objective(self) -> Tuple[np.ndarray, float]:
        """(coefficient of every variable, constant term)"""
        out = np.empty(self.nr_variables(), dtype=np.float64)
        const = C.c_double()
        _lib.check(self._lib.bddb200_ilp_objective(self._h, out.ctypes.data, C.byref(const)))
        return out, const.value

    def variable_names(self) -> List[str]:
        names = []
        for v in range(self.nr_variables()):
            p = C.c_char_p()
            _lib.check(self._lib.bddb200_ilp_variable_name(self._h, v, C.byref(p)))
            names.append(p.value.decode())
        return names

    def constraint(self, c: int) -> Tuple[List[int], List[int], int, int]:
        """(variables, coefficients, relation 0 '<=' / 1 '>=' / 2 '=', right-hand side)"""
        n, rel, rhs = C.c_size_t(), C.c_int(), C.c_longlong()
        _lib.check(self._lib.bddb200_ilp_constraint(self._h, c, C.byref(n), None, None, 0, None, None))
        va, co = np.empty(n.value, dtype=np.uint64), np.empty(n.value, dtype=np.int64)
        _lib.check(self._lib.bddb200_ilp_constraint(self._h, c, C.byref(n), va.ctypes.data, co.ctypes.data, n.value, C.byref(rel), C.byref(rhs)))
        return va.astype(np.int64).tolist(), co.tolist(), rel.value, rhs.value

    def to_bdds(self) -> bdd_collection:
        """One quasi-reduced BDD per constraint that is not always satisfied (bdd_preprocessor.cpp:123-228)."""
        h = C.c_void_p()
        _lib.check(self._lib.bddb200_ilp_to_bdds(self._h, C.byref(h)))
        return bdd_collection(_handle=h)
