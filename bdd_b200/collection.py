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
    def __init__(self, col: Optional[BddCollection] = None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        if col is None:
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
