"""Host-side mirror of the reference's GPU solver class over the C ABI.

``bdd_cuda_parallel_mma`` has the method set of ``LPMP::bdd_cuda_parallel_mma<REAL>`` and its
base ``LPMP::bdd_cuda_base<REAL>`` (include/bdd_solver/bdd_cuda_parallel_mma.h:7-52,
include/bdd_solver/bdd_cuda_base.h:57-226): same names, argument meaning and error behaviour
(exceptions), with ``torch`` CUDA tensors standing in for ``thrust::device_vector``.  Every
method is one call into libbdd_b200.so (include/bdd_b200.h); PyTorch only owns device
buffers and streams.  There is no CPU fallback.

``run_solver`` is the termination loop of include/run_solver_util.h:10-77.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import Options, check
from .instances import BddCollection

INT_MAX = 2 ** 31 - 1


def reference_layer_order(primal_index: np.ndarray, bdd_index: np.ndarray) -> np.ndarray:
    """The reference sorts its BDD nodes by (hop distance from the root, primal variable, BDD) and compresses equal keys to layers
    (bdd_cuda_base.cu:146-188, :240-285); the two sinks of a BDD share the hop after its last variable and the variable INT_MAX
    (:113-129).  Its per-layer vectors are therefore ordered by (position of the layer in its BDD, variable, BDD), terminal layers
    last within their hop.  Given the per-layer variable and BDD of a BDD-major order (layers of BDD 0, its terminal layer, BDD 1, ...),
    returns the permutation that lists the BDD-major positions in the reference's order."""
    primal_index = np.asarray(primal_index, dtype=np.int64)
    bdd_index = np.asarray(bdd_index, dtype=np.int64)
    n = bdd_index.shape[0]
    first = np.flatnonzero(np.concatenate([[True], bdd_index[1:] != bdd_index[:-1]])) if n else np.zeros(0, dtype=np.int64)
    run = np.cumsum(np.concatenate([[True], bdd_index[1:] != bdd_index[:-1]])) - 1 if n else np.zeros(0, dtype=np.int64)
    hop = np.arange(n, dtype=np.int64) - first[run]
    return np.lexsort((bdd_index, primal_index, hop)).astype(np.int64)


class bdd_cuda_parallel_mma:
    """``bdd_cuda_parallel_mma<REAL>(bdd_col, costs)`` (bdd_cuda_parallel_mma.cu:7-17).

    precision: "float" or "double" (REAL).  ``deterministic`` selects fixed-order per-variable
    sums; ``nr_variables`` / ``nr_bdds_per_var`` put the solver in shard mode.
    """

    def __init__(self, bdd_col: BddCollection, costs: Optional[Sequence[float]] = None, precision: str = "float",
                 device: int = 0, deterministic: bool = False, lanes_per_bdd: int = 0, nr_variables: int = 0,
                 nr_bdds_per_var: Optional[np.ndarray] = None, stream: Optional[torch.cuda.Stream] = None, n_shared_vars: int = 0):
        self.lib = _lib.load()
        if precision not in ("float", "double"):
            raise ValueError("precision must be 'float' or 'double'")
        if not torch.cuda.is_available():
            raise RuntimeError("bdd_cuda_parallel_mma needs a CUDA device (no CPU fallback)")
        self.precision = precision
        self.value_type = torch.float64 if precision == "double" else torch.float32
        self.np_type = np.float64 if precision == "double" else np.float32
        self.device = torch.device("cuda", device)
        self.stream = stream if stream is not None else torch.cuda.Stream(self.device)
        instrs = np.ascontiguousarray(bdd_col.instrs, dtype=np.uint64)
        delims = np.ascontiguousarray(bdd_col.delims, dtype=np.uint64)
        opts = Options()
        self.lib.bddb200_default_options(C.byref(opts))
        opts.device = device
        opts.stream = self.stream.cuda_stream
        opts.deterministic = int(deterministic)
        opts.lanes_per_bdd = int(lanes_per_bdd)
        opts.nr_variables = int(nr_variables)
        opts.n_shared_vars = int(n_shared_vars)
        self._nbpv = None
        if nr_bdds_per_var is not None:
            self._nbpv = np.ascontiguousarray(nr_bdds_per_var, dtype=np.int32)
            opts.nr_bdds_per_var_host = self._nbpv.ctypes.data
        cptr, ncost = None, 0
        if costs is not None:
            self._costs = np.ascontiguousarray(costs, dtype=np.float64)
            cptr, ncost = self._costs.ctypes.data, self._costs.shape[0]
        h = C.c_void_p()
        check(self.lib.bddb200_create(instrs.ctypes.data, instrs.shape[0], delims.ctypes.data, delims.shape[0] - 1,
                                      cptr, ncost, _lib.DOUBLE if precision == "double" else _lib.FLOAT,
                                      C.byref(opts), C.byref(h)))
        self.h = h

    def __del__(self):
        h = getattr(self, "h", None)
        if h:
            self.lib.bddb200_destroy(h)
            self.h = None

    # ------------------------------------------------------------------ save / load ----
    def save(self) -> bytes:
        """The whole solver state as one blob (cereal save of the reference class, bdd_cuda_base.cu:1486-1544; what the
        reference's pybind module pickles, bdd_cuda_parallel_mma_py.cu:29-38)."""
        n = C.c_size_t()
        check(self.lib.bddb200_save_size(self.h, C.byref(n)))
        buf = (C.c_ubyte * n.value)()
        check(self.lib.bddb200_save(self.h, buf, n.value, C.byref(n)))
        return bytes(buf)[: n.value]

    @classmethod
    def load(cls, blob: bytes, device: int = 0) -> "bdd_cuda_parallel_mma":
        """A new solver from ``save()``'s blob, on ``device`` (same GPU model), running on a stream of its own."""
        self = cls.__new__(cls)
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("bdd_cuda_parallel_mma needs a CUDA device (no CPU fallback)")
        h = C.c_void_p()
        raw = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        check(self.lib.bddb200_load(raw, len(blob), device, C.byref(h)))
        self.h = h
        self.precision = "double" if self.lib.bddb200_precision_of(h) == _lib.DOUBLE else "float"
        self.value_type = torch.float64 if self.precision == "double" else torch.float32
        self.np_type = np.float64 if self.precision == "double" else np.float32
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.ExternalStream(self.lib.bddb200_stream(h), device=self.device)
        self._nbpv = None
        return self

    @classmethod
    def create_shard(cls, bdd_col: BddCollection, costs: Optional[Sequence[float]], rank: int, world: int, precision: str = "float",
                     device: int = 0, deterministic: bool = False):
        """This rank's solver of a ``world``-way constraint-sharded solve, planned and built inside the library
        (``bddb200_create_shard``): returns (solver, info dict, new_of_old).  Solver vectors indexed by variable use the NEW numbering."""
        self = cls.__new__(cls)
        self.lib = _lib.load()
        if precision not in ("float", "double"):
            raise ValueError("precision must be 'float' or 'double'")
        if not torch.cuda.is_available():
            raise RuntimeError("bdd_cuda_parallel_mma needs a CUDA device (no CPU fallback)")
        self.precision = precision
        self.value_type = torch.float64 if precision == "double" else torch.float32
        self.np_type = np.float64 if precision == "double" else np.float32
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(self.device)
        instrs = np.ascontiguousarray(bdd_col.instrs, dtype=np.uint64)
        delims = np.ascontiguousarray(bdd_col.delims, dtype=np.uint64)
        opts = Options()
        self.lib.bddb200_default_options(C.byref(opts))
        opts.device = device
        opts.stream = self.stream.cuda_stream
        opts.deterministic = int(deterministic)
        self._nbpv = None
        cptr, ncost = None, 0
        if costs is not None:
            self._costs = np.ascontiguousarray(costs, dtype=np.float64)
            cptr, ncost = self._costs.ctypes.data, self._costs.shape[0]
        n_vars = max(bdd_col.nr_variables(), ncost)
        new_of_old = np.empty(n_vars, dtype=np.int32)
        info = _lib.ShardInfo()
        h = C.c_void_p()
        check(self.lib.bddb200_create_shard(instrs.ctypes.data, instrs.shape[0], delims.ctypes.data, delims.shape[0] - 1, cptr, ncost,
                                            _lib.DOUBLE if precision == "double" else _lib.FLOAT, C.byref(opts), world, rank,
                                            C.byref(info), new_of_old.ctypes.data, C.byref(h)))
        self.h = h
        d = {k: int(getattr(info, k)) for k, _ in _lib.ShardInfo._fields_}
        return self, d, new_of_old.astype(np.int64)

    def __getstate__(self):
        return {"blob": self.save(), "device": self.device.index}

    def __setstate__(self, state):
        other = type(self).load(state["blob"], state["device"])
        self.__dict__.update(other.__dict__)
        other.h = None

    # ------------------------------------------------------------------ helpers --------
    def _empty(self, n: int, dtype=None) -> torch.Tensor:
        with torch.cuda.stream(self.stream):
            return torch.empty(n, dtype=dtype or self.value_type, device=self.device)

    def _in(self, t: torch.Tensor, n: Optional[int] = None, dtype=None) -> torch.Tensor:
        dtype = dtype or self.value_type
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
            raise TypeError(f"expected a contiguous CUDA tensor of dtype {dtype}")
        if n is not None and t.numel() != n:
            raise ValueError(f"expected {n} elements, got {t.numel()}")
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        return t

    def _out(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    # ------------------------------------------------------------------ sizes ----------
    def nr_variables(self) -> int:
        return self.lib.bddb200_nr_variables(self.h)

    def nr_bdds(self, var: Optional[int] = None) -> int:
        if var is None:
            return self.lib.bddb200_nr_bdds(self.h)
        return int(self.get_num_bdds_per_var()[var])

    def nr_layers(self) -> int:
        return self.lib.bddb200_nr_layers(self.h)

    def nr_bdd_nodes(self) -> int:
        return self.lib.bddb200_nr_bdd_nodes(self.h)

    def nr_hops(self) -> int:
        return self.lib.bddb200_nr_hops(self.h)

    def get_num_bdds_per_var(self) -> np.ndarray:
        out = np.empty(self.nr_variables(), dtype=np.int32)
        check(self.lib.bddb200_nr_bdds_per_var(self.h, out.ctypes.data))
        return out

    def get_primal_variable_index(self) -> np.ndarray:
        out = np.empty(self.nr_layers(), dtype=np.int32)
        check(self.lib.bddb200_layer_primal_indices(self.h, out.ctypes.data))
        return out

    def get_bdd_index(self) -> np.ndarray:
        out = np.empty(self.nr_layers(), dtype=np.int32)
        check(self.lib.bddb200_layer_bdd_indices(self.h, out.ctypes.data))
        return out

    def reference_layer_order(self) -> np.ndarray:
        """``perm`` with ``perm[k]`` = position, in this class's BDD-major layer order, of the layer the reference's
        ``bdd_cuda_base`` keeps at position ``k`` (see :func:`reference_layer_order`).  ``v[perm]`` turns a per-layer vector of
        this solver (``get_solver_costs``, ``net_solver_costs``, ``bdds_solution_vec``, ...) into the reference's order;
        ``w = empty_like(v); w[perm] = v_ref`` brings one saved from the reference back."""
        return reference_layer_order(self.get_primal_variable_index(), self.get_bdd_index())

    # ------------------------------------------------------------------ hot path -------
    def iteration(self, omega: float = 0.5):
        check(self.lib.bddb200_iteration(self.h, omega))

    def iterations(self, n: int, omega: float = 0.5):
        check(self.lib.bddb200_iterations(self.h, omega, n))

    def forward_pass(self, omega: float = 0.5):
        check(self.lib.bddb200_forward_pass(self.h, omega))

    def backward_pass(self, omega: float = 0.5):
        check(self.lib.bddb200_backward_pass(self.h, omega))

    def forward_mm(self, omega: float, delta_lo_hi: torch.Tensor):
        check(self.lib.bddb200_forward_mm(self.h, omega, self._in(delta_lo_hi, 2 * self.nr_variables()).data_ptr()))
        self._out()

    def backward_mm(self, omega: float, delta_lo_hi: torch.Tensor):
        check(self.lib.bddb200_backward_mm(self.h, omega, self._in(delta_lo_hi, 2 * self.nr_variables()).data_ptr()))
        self._out()

    def normalize_delta(self, delta_lo_hi: torch.Tensor):
        check(self.lib.bddb200_normalize_delta(self.h, self._in(delta_lo_hi, 2 * self.nr_variables()).data_ptr()))
        self._out()

    def get_delta(self) -> torch.Tensor:
        out = self._empty(2 * self.nr_variables())
        check(self.lib.bddb200_get_delta(self.h, out.data_ptr(), 0))
        self._out()
        return out

    def delta_sum_view(self) -> torch.Tensor:
        """The solver's current un-normalised delta sums as a tensor aliasing solver memory
        (for the multi-GPU all-reduce; valid until the next pass)."""
        p = C.c_void_p()
        check(self.lib.bddb200_delta_sum_buffer(self.h, C.byref(p)))
        n = 2 * self.nr_variables()
        itemsize = 8 if self.precision == "double" else 4
        iface = {"shape": (n,), "typestr": "<f8" if self.precision == "double" else "<f4",
                 "data": (p.value, False), "version": 3, "strides": (itemsize,)}
        holder = type("_DevView", (), {"__cuda_array_interface__": iface})()
        with torch.cuda.stream(self.stream):
            return torch.as_tensor(holder, device=self.device)

    def delta_sum_index(self) -> int:
        """Which of the three rotating sum buffers holds the sums of the last pass."""
        out = C.c_int()
        check(self.lib.bddb200_delta_sum_index(self.h, C.byref(out)))
        return out.value

    def push_exchange_supported(self) -> bool:
        """Whether this solver can run the multi-GPU push exchange (bddb200_set_exchange mode 4)."""
        out = C.c_int()
        check(self.lib.bddb200_push_exchange_supported(self.h, C.byref(out)))
        return out.value != 0

    def set_push_masks(self, masks: np.ndarray):
        """Per shared variable the ranks whose shards contain it (bit r = rank r): where the push exchange sends its differences."""
        m = np.ascontiguousarray(masks, dtype=np.uint16)
        check(self.lib.bddb200_set_push_masks(self.h, m.ctypes.data, m.shape[0]))

    def set_delta_buffers(self, block: torch.Tensor):
        """Use ``block`` (3 x 2V zero-filled REALs, e.g. symmetric memory mapped by the peer GPUs) as the rotating sum buffers."""
        n = 2 * self.nr_variables()
        self._in(block, 3 * n)
        self._delta_block = block
        item = block.element_size()
        check(self.lib.bddb200_set_delta_buffers(self.h, block.data_ptr(), block.data_ptr() + n * item, block.data_ptr() + 2 * n * item))

    def set_delta_input(self, t: Optional[torch.Tensor], n_shared_vars: int = 0):
        """Passes read the (exchanged, un-normalised) sums of variables [0, n_shared_vars) from ``t`` and the rest from
        the rotating buffers."""
        if t is not None:
            self._in(t)
        self._delta_input = t
        check(self.lib.bddb200_set_delta_input(self.h, t.data_ptr() if t is not None else None, int(n_shared_vars)))

    def lower_bound(self) -> float:
        out = C.c_double()
        check(self.lib.bddb200_lower_bound(self.h, C.byref(out)))
        return out.value

    def lower_bound_per_bdd(self) -> torch.Tensor:
        out = self._empty(self.nr_bdds())
        check(self.lib.bddb200_lower_bound_per_bdd(self.h, out.data_ptr()))
        self._out()
        return out

    def forward_run(self):
        check(self.lib.bddb200_forward_run(self.h))

    def backward_run(self):
        check(self.lib.bddb200_backward_run(self.h))

    def flush_forward_states(self):
        self.lib.bddb200_flush_forward_states(self.h)

    def flush_backward_states(self):
        self.lib.bddb200_flush_backward_states(self.h)

    # ------------------------------------------------------------------ costs ----------
    def update_costs(self, cost_delta_0, cost_delta_1):
        """update_costs(lo, hi): host sequences (std::vector overload, bdd_cuda_base.cu:519-523)
        or CUDA tensors of REAL (device_vector overload, :525-558)."""
        if isinstance(cost_delta_0, torch.Tensor) or isinstance(cost_delta_1, torch.Tensor):
            lo = cost_delta_0 if isinstance(cost_delta_0, torch.Tensor) else None
            hi = cost_delta_1 if isinstance(cost_delta_1, torch.Tensor) else None
            nlo = lo.numel() if lo is not None else 0
            nhi = hi.numel() if hi is not None else 0
            check(self.lib.bddb200_update_costs_dev(self.h, self._in(lo).data_ptr() if nlo else None, nlo,
                                                    self._in(hi).data_ptr() if nhi else None, nhi))
            return
        def as_host(x):
            if x is None:
                return None
            if isinstance(x, np.ndarray) and x.dtype == self.np_type and x.flags.c_contiguous:
                return x
            return np.ascontiguousarray(x, dtype=np.float64)
        lo, hi = as_host(cost_delta_0), as_host(cost_delta_1)
        real = [a.dtype == self.np_type for a in (lo, hi) if a is not None and a.size]
        if real and all(real) and self.np_type != np.float64:
            fn = self.lib.bddb200_update_costs_host_real      # std::vector<REAL> overload
        else:
            lo = None if lo is None else np.ascontiguousarray(lo, dtype=np.float64)
            hi = None if hi is None else np.ascontiguousarray(hi, dtype=np.float64)
            fn = self.lib.bddb200_update_costs_host
        nlo = 0 if lo is None else lo.size
        nhi = 0 if hi is None else hi.size
        check(fn(self.h, lo.ctypes.data if nlo else None, nlo, hi.ctypes.data if nhi else None, nhi))

    def step(self, cost_delta_0, cost_delta_1, omega: float = 0.5) -> float:
        """update_costs(lo, hi) from host arrays + iteration(omega) + lower_bound() in one library call (bddb200_step_host): one
        upload straight from the arrays (DMA when they are pinned, e.g. views of ``torch.empty(..., pin_memory=True)``), one CUDA
        graph launch, one read-back.  Arrays of the solver's REAL type go as they are, anything else is converted to double."""
        def as_host(x):
            if x is None:
                return None
            if isinstance(x, np.ndarray) and x.dtype == self.np_type and x.flags.c_contiguous:
                return x
            return np.ascontiguousarray(x, dtype=np.float64)
        lo, hi = as_host(cost_delta_0), as_host(cost_delta_1)
        real = [a.dtype == self.np_type for a in (lo, hi) if a is not None and a.size]
        is_real = bool(real) and all(real)
        if not is_real:
            lo = None if lo is None else np.ascontiguousarray(lo, dtype=np.float64)
            hi = None if hi is None else np.ascontiguousarray(hi, dtype=np.float64)
        nlo = 0 if lo is None else lo.size
        nhi = 0 if hi is None else hi.size
        out = C.c_double()
        check(self.lib.bddb200_step_host(self.h, lo.ctypes.data if nlo else None, nlo, hi.ctypes.data if nhi else None, nhi,
                                         int(is_real), omega, C.byref(out)))
        return out.value

    def set_cost(self, c: float, var: int):
        check(self.lib.bddb200_set_cost(self.h, c, var))

    def distribute_delta(self):
        check(self.lib.bddb200_distribute_delta(self.h))

    def get_solver_costs(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        n = self.nr_layers()
        lo, hi, mm = self._empty(n), self._empty(n), self._empty(n)
        check(self.lib.bddb200_get_solver_costs(self.h, lo.data_ptr(), hi.data_ptr(), mm.data_ptr()))
        self._out()
        return lo, hi, mm

    def set_solver_costs(self, costs: Tuple[torch.Tensor, torch.Tensor, torch.Tensor]):
        n = self.nr_layers()
        lo, hi, mm = (self._in(t, n) for t in costs)
        check(self.lib.bddb200_set_solver_costs(self.h, lo.data_ptr(), hi.data_ptr(), mm.data_ptr()))

    def get_primal_objective_vector_host(self) -> np.ndarray:
        out = np.empty(self.nr_variables(), dtype=np.float64)
        check(self.lib.bddb200_primal_objective_host(self.h, out.ctypes.data))
        return out

    # ------------------------------------------------------------------ min-marginals --
    def min_marginals_cuda(self, get_sorted: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(primal index, mm_lo, mm_hi) per layer; sorted by (variable, BDD) with the
        nr_bdds() terminal entries last when get_sorted (bdd_cuda_base.cu:716-751)."""
        n = self.nr_layers()
        idx, lo, hi = self._empty(n, torch.int32), self._empty(n), self._empty(n)
        check(self.lib.bddb200_min_marginals(self.h, int(get_sorted), idx.data_ptr(), lo.data_ptr(), hi.data_ptr()))
        self._out()
        return idx, lo, hi

    def min_marginals(self) -> List[np.ndarray]:
        """two_dim_variable_array<array<double,2>>: per variable an [nr_bdds(var), 2] array
        (bdd_cuda_base.cu:753-786)."""
        _, lo, hi = self.min_marginals_cuda(True)
        self.synchronize()
        lo = lo.cpu().numpy().astype(np.float64)
        hi = hi.cpu().numpy().astype(np.float64)
        counts = self.get_num_bdds_per_var()
        local = np.bincount(self.get_primal_variable_index()[self.get_primal_variable_index() != INT_MAX], minlength=self.nr_variables())
        out, pos = [], 0
        for v in range(self.nr_variables()):
            k = int(local[v])
            out.append(np.stack([lo[pos:pos + k], hi[pos:pos + k]], axis=1))
            pos += k
        del counts
        return out

    # ------------------------------------------------------------------ L-BFGS surface -
    def bdds_solution_vec(self) -> torch.Tensor:
        out = self._empty(self.nr_layers(), torch.int8)
        check(self.lib.bddb200_bdds_solution(self.h, out.data_ptr()))
        self._out()
        return out

    def net_solver_costs(self) -> torch.Tensor:
        out = self._empty(self.nr_layers())
        check(self.lib.bddb200_net_solver_costs(self.h, out.data_ptr()))
        self._out()
        return out

    def make_dual_feasible(self, d: torch.Tensor):
        check(self.lib.bddb200_make_dual_feasible(self.h, self._in(d, self.nr_layers()).data_ptr()))
        self._out()

    def gradient_step(self, g: torch.Tensor, step_size: float):
        check(self.lib.bddb200_gradient_step(self.h, self._in(g, self.nr_layers()).data_ptr(), step_size))

    # ------------------------------------------------------------------ plumbing -------
    # ------------------------------------------------------------------ primal rounding --
    def rounding_perturb(self, delta: float, round_index: int = 0):
        """perturb_primal_costs (incremental_mm_agreement_rounding_cuda.cu:264-335).  Returns (solution or None, counts, types):
        counts = variables of type [zero, one, equal, inconsistent], types = per-variable type (CUDA char tensor)."""
        n = self.nr_variables()
        counts = np.zeros(4, dtype=np.uint64)
        types = self._empty(n, dtype=torch.int8)
        sol = np.zeros(n, dtype=np.int8)
        solved = C.c_int()
        check(self.lib.bddb200_rounding_perturb(self.h, float(delta), int(round_index), counts.ctypes.data, types.data_ptr(), sol.ctypes.data, C.byref(solved)))
        self._out()
        return (sol if solved.value else None), counts, types

    def incremental_mm_agreement_rounding(self, init_delta: float = 1.0, delta_growth_rate: float = 1.2, num_itr_lb: int = 100,
                                          num_rounds: int = 500):
        """incremental_mm_agreement_rounding_cuda (:338-375; defaults of bdd_solver.cpp:318-335).  Returns (solution or None, rounds used)."""
        sol = np.zeros(self.nr_variables(), dtype=np.int8)
        solved, rounds = C.c_int(), C.c_int()
        check(self.lib.bddb200_incremental_mm_agreement_rounding(self.h, getattr(self, "lh", None), float(init_delta), float(delta_growth_rate),
                                                                 int(num_itr_lb), int(num_rounds), sol.ctypes.data, C.byref(solved), C.byref(rounds)))
        return (sol if solved.value else None), rounds.value

    def run_solver(self, max_iter: int = 1000, tolerance: float = 1e-6, improvement_slope: float = 1e-9, time_limit: float = 3600.0) -> float:
        """run_solver (include/run_solver_util.h:10-77) inside the library; returns the final lower bound."""
        lb = C.c_double()
        check(self.lib.bddb200_run_solver(self.h, getattr(self, "lh", None), int(max_iter), float(tolerance), float(improvement_slope), float(time_limit), C.byref(lb)))
        return lb.value

    def synchronize(self):
        check(self.lib.bddb200_synchronize(self.h))

    def trace_pass(self, forward: bool, omega: float = 0.5, max_bundles: int = 1 << 20) -> np.ndarray:
        """Diagnostics: per-bundle clock64() stamps of one MMA pass, [n_bundles, 16] (include/bdd_b200.h)."""
        out = np.zeros((max_bundles, 16), dtype=np.uint64)
        n = C.c_size_t()
        check(self.lib.bddb200_trace_pass(self.h, int(forward), omega, out.ctypes.data, max_bundles, C.byref(n)))
        return out[: n.value]

    def kernel_launches(self) -> int:
        return self.lib.bddb200_kernel_launches(self.h)


class lbfgs_cuda_mma(bdd_cuda_parallel_mma):
    """``lbfgs<bdd_cuda_parallel_mma<REAL>, device_vector<REAL>, REAL, device_vector<char>, true>`` (include/bdd_solver/lbfgs.h:35-110;
    config strings "lbfgs cuda mma", "cuda lbfgs parallel mma" and the README spelling "lbfgs cuda parallel mma").  ``iteration()`` is the
    wrapper's: history bookkeeping, L-BFGS step when the history is full, then one MMA iteration -- all on the device
    (bdd_b200/csrc/lbfgs.cuh)."""

    def __init__(self, bdd_col: BddCollection, costs: Optional[Sequence[float]] = None, precision: str = "float",
                 history_size: int = 5, init_step_size: float = 1e-6, req_rel_lb_increase: float = 1e-6,
                 step_size_decrease_factor: float = 0.8, step_size_increase_factor: float = 1.1, **kw):
        super().__init__(bdd_col, costs, precision=precision, **kw)
        h = C.c_void_p()
        check(self.lib.bddb200_lbfgs_create(self.h, int(history_size), float(init_step_size), float(req_rel_lb_increase),
                                            float(step_size_decrease_factor), float(step_size_increase_factor), C.byref(h)))
        self.lh = h

    def __del__(self):
        lh = getattr(self, "lh", None)
        if lh:
            self.lib.bddb200_lbfgs_destroy(lh)
            self.lh = None
        super().__del__()

    def iteration(self, omega: float = 0.5):
        check(self.lib.bddb200_lbfgs_iteration(self.lh))

    def mma_iteration(self, omega: float = 0.5):
        super().iteration(omega)

    def update_costs(self, cost_delta_0, cost_delta_1):
        check(self.lib.bddb200_lbfgs_flush(self.lh))          # lbfgs<>::update_costs, lbfgs_impl.h:343-348
        super().update_costs(cost_delta_0, cost_delta_1)

    def lbfgs_stats(self) -> Tuple[int, int, float]:
        """(L-BFGS iterations, plain MMA iterations, current step size)"""
        a, b, st = C.c_size_t(), C.c_size_t(), C.c_double()
        check(self.lib.bddb200_lbfgs_stats(self.lh, C.byref(a), C.byref(b), C.byref(st)))
        return a.value, b.value, st.value


def run_solver(s, max_iter: int = 1000, tolerance: float = 1e-6, improvement_slope: float = 1e-9,
               time_limit: float = 3600.0, verbose: bool = False, log=print):
    """include/run_solver_util.h:10-77: iterate until one of the four stop rules fires.
    Returns the list of (iteration, lower bound, seconds)."""
    start = time.monotonic()
    lb_initial = s.lower_bound()
    lb_first_iter = float("inf")
    lb_prev = lb_post = lb_initial
    trace = [(-1, lb_initial, 0.0)]
    if verbose:
        log(f"[bdd solver] initial lower bound = {lb_prev}, time = 0 s")
    for it in range(max_iter):
        s.iteration()
        lb_prev = lb_post
        lb_post = s.lower_bound()
        if it == 0:
            lb_first_iter = lb_post
        spent = time.monotonic() - start
        trace.append((it, lb_post, spent))
        if verbose:
            log(f"[bdd solver] iteration {it}, lower bound = {lb_post}, time = {spent:.3f} s")
        if spent > time_limit:
            break
        if abs(lb_prev - lb_post) < abs(tolerance * lb_prev):
            break
        if abs(lb_prev - lb_post) < improvement_slope * abs(lb_initial - lb_first_iter):
            break
        if lb_post == float("inf"):
            break
    return trace
