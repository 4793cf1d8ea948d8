"""Loader for libbdd_b200.so (the C-ABI library declared in include/bdd_b200.h).

Fails loudly when the library is missing: there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BDDB200_LIB", os.path.join(_HERE, "libbdd_b200.so"))      # BDDB200_LIB: A/B builds of the library

OK = 0
FLOAT, DOUBLE = 0, 1

# every symbol include/bdd_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "bddb200_default_options", "bddb200_last_error", "bddb200_version", "bddb200_create", "bddb200_destroy", "bddb200_clone", "bddb200_plan_shard", "bddb200_create_shard",
    "bddb200_save_size", "bddb200_save", "bddb200_load",
    "bddb200_nr_variables", "bddb200_nr_bdds", "bddb200_nr_layers", "bddb200_nr_bdd_nodes", "bddb200_nr_hops",
    "bddb200_precision_of", "bddb200_device_of", "bddb200_nr_bdds_per_var", "bddb200_layer_primal_indices",
    "bddb200_layer_bdd_indices", "bddb200_iteration", "bddb200_iterations", "bddb200_forward_pass",
    "bddb200_backward_pass", "bddb200_forward_mm", "bddb200_backward_mm", "bddb200_normalize_delta",
    "bddb200_get_delta", "bddb200_lower_bound", "bddb200_lower_bound_per_bdd", "bddb200_forward_run",
    "bddb200_backward_run", "bddb200_flush_forward_states", "bddb200_flush_backward_states",
    "bddb200_update_costs_host", "bddb200_update_costs_host_real", "bddb200_update_costs_dev", "bddb200_step_host", "bddb200_set_cost", "bddb200_distribute_delta",
    "bddb200_get_solver_costs", "bddb200_set_solver_costs", "bddb200_primal_objective_host",
    "bddb200_min_marginals", "bddb200_min_marginals_host", "bddb200_bdds_solution", "bddb200_net_solver_costs", "bddb200_make_dual_feasible",
    "bddb200_gradient_step", "bddb200_synchronize", "bddb200_stream", "bddb200_kernel_launches",
    "bddb200_delta_sum_buffer", "bddb200_layout_stats", "bddb200_trace_pass",
    "bddb200_delta_sum_index", "bddb200_push_exchange_supported", "bddb200_set_push_masks", "bddb200_set_delta_buffers", "bddb200_set_delta_input", "bddb200_set_exchange", "bddb200_delta_exchange", "bddb200_delta_exchange_two_shot",
    "bddb200_run_solver", "bddb200_rounding_perturb", "bddb200_incremental_mm_agreement_rounding",
    "bddb200_lbfgs_create", "bddb200_lbfgs_destroy", "bddb200_lbfgs_iteration", "bddb200_lbfgs_flush", "bddb200_lbfgs_stats",
]

# every symbol include/bdd_b200_collection.h declares (host-side BDD collection; no GPU needed)
COLLECTION_SYMBOLS = [
    "bddb200_collection_create", "bddb200_collection_destroy", "bddb200_collection_nr_bdds", "bddb200_collection_nr_instructions", "bddb200_collection_export",
    "bddb200_collection_simplex_constraint", "bddb200_collection_not_all_false_constraint", "bddb200_collection_all_equal_constraint",
    "bddb200_collection_cardinality_constraint", "bddb200_collection_add_linear_constraint", "bddb200_collection_rebase", "bddb200_collection_negate", "bddb200_collection_invert",
    "bddb200_collection_variables", "bddb200_collection_is_qbdd", "bddb200_collection_is_reordered", "bddb200_collection_evaluate",
    "bddb200_collection_reorder", "bddb200_collection_make_qbdd", "bddb200_collection_bdd_and", "bddb200_collection_remove",
    "bddb200_create_from_collection", "bddb200_collection_write_bdd_lp", "bddb200_collection_split_qbdd", "bddb200_collection_split_long_bdds",
    "bddb200_ilp_read", "bddb200_ilp_destroy", "bddb200_ilp_nr_variables", "bddb200_ilp_nr_constraints", "bddb200_ilp_objective",
    "bddb200_ilp_variable_name", "bddb200_ilp_constraint", "bddb200_ilp_to_bdds",
]


class ShardInfo(C.Structure):
    """bddb200_shard_info"""
    _fields_ = [("nr_variables", C.c_size_t), ("n_shared", C.c_size_t), ("shared_entries", C.c_size_t), ("first_bdd", C.c_size_t), ("n_bdds", C.c_size_t)]


class Options(C.Structure):
    _fields_ = [
        ("device", C.c_int),
        ("stream", C.c_void_p),
        ("deterministic", C.c_int),
        ("lanes_per_bdd", C.c_int),
        ("nr_variables", C.c_size_t),
        ("nr_bdds_per_var_host", C.c_void_p),
        ("stage_bytes", C.c_int),
        ("n_stages", C.c_int),
        ("warps_per_cta", C.c_int),
        ("n_shared_vars", C.c_size_t),
    ]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C bdd_b200/csrc` or __graft_entry__.build(). "
            "bdd_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, sz, dbl, i = C.c_void_p, C.c_size_t, C.c_double, C.c_int
    sig = {
        "bddb200_default_options": (None, [C.POINTER(Options)]),
        "bddb200_last_error": (C.c_char_p, []),
        "bddb200_version": (C.c_char_p, []),
        "bddb200_create": (i, [vp, sz, vp, sz, vp, sz, i, C.POINTER(Options), C.POINTER(vp)]),
        "bddb200_destroy": (None, [vp]),
        "bddb200_clone": (i, [vp, C.POINTER(vp)]),
        "bddb200_plan_shard": (i, [vp, sz, vp, sz, sz, i, i, C.POINTER(ShardInfo), vp, vp, vp]),
        "bddb200_create_shard": (i, [vp, sz, vp, sz, vp, sz, i, C.POINTER(Options), i, i, C.POINTER(ShardInfo), vp, C.POINTER(vp)]),
        "bddb200_save_size": (i, [vp, C.POINTER(sz)]),
        "bddb200_save": (i, [vp, vp, sz, C.POINTER(sz)]),
        "bddb200_load": (i, [vp, sz, i, C.POINTER(vp)]),
        "bddb200_nr_variables": (sz, [vp]),
        "bddb200_nr_bdds": (sz, [vp]),
        "bddb200_nr_layers": (sz, [vp]),
        "bddb200_nr_bdd_nodes": (sz, [vp]),
        "bddb200_nr_hops": (sz, [vp]),
        "bddb200_precision_of": (i, [vp]),
        "bddb200_device_of": (i, [vp]),
        "bddb200_nr_bdds_per_var": (i, [vp, vp]),
        "bddb200_layer_primal_indices": (i, [vp, vp]),
        "bddb200_layer_bdd_indices": (i, [vp, vp]),
        "bddb200_iteration": (i, [vp, dbl]),
        "bddb200_iterations": (i, [vp, dbl, sz]),
        "bddb200_forward_pass": (i, [vp, dbl]),
        "bddb200_backward_pass": (i, [vp, dbl]),
        "bddb200_forward_mm": (i, [vp, dbl, vp]),
        "bddb200_backward_mm": (i, [vp, dbl, vp]),
        "bddb200_normalize_delta": (i, [vp, vp]),
        "bddb200_get_delta": (i, [vp, vp, i]),
        "bddb200_lower_bound": (i, [vp, C.POINTER(dbl)]),
        "bddb200_lower_bound_per_bdd": (i, [vp, vp]),
        "bddb200_forward_run": (i, [vp]),
        "bddb200_backward_run": (i, [vp]),
        "bddb200_flush_forward_states": (None, [vp]),
        "bddb200_flush_backward_states": (None, [vp]),
        "bddb200_update_costs_host": (i, [vp, vp, sz, vp, sz]),
        "bddb200_update_costs_dev": (i, [vp, vp, sz, vp, sz]),
        "bddb200_update_costs_host_real": (i, [vp, vp, sz, vp, sz]),
        "bddb200_step_host": (i, [vp, vp, sz, vp, sz, i, dbl, C.POINTER(dbl)]),
        "bddb200_set_cost": (i, [vp, dbl, sz]),
        "bddb200_distribute_delta": (i, [vp]),
        "bddb200_get_solver_costs": (i, [vp, vp, vp, vp]),
        "bddb200_set_solver_costs": (i, [vp, vp, vp, vp]),
        "bddb200_primal_objective_host": (i, [vp, vp]),
        "bddb200_min_marginals": (i, [vp, i, vp, vp, vp]),
        "bddb200_min_marginals_host": (i, [vp, i, vp, vp, vp]),
        "bddb200_bdds_solution": (i, [vp, vp]),
        "bddb200_net_solver_costs": (i, [vp, vp]),
        "bddb200_make_dual_feasible": (i, [vp, vp]),
        "bddb200_gradient_step": (i, [vp, vp, dbl]),
        "bddb200_synchronize": (i, [vp]),
        "bddb200_stream": (vp, [vp]),
        "bddb200_kernel_launches": (sz, [vp]),
        "bddb200_delta_sum_buffer": (i, [vp, C.POINTER(vp)]),
        "bddb200_layout_stats": (i, [vp, sz, vp, sz, i, vp, sz]),
        "bddb200_trace_pass": (i, [vp, i, dbl, vp, sz, C.POINTER(sz)]),
        "bddb200_run_solver": (i, [vp, vp, sz, dbl, dbl, dbl, C.POINTER(dbl)]),
        "bddb200_rounding_perturb": (i, [vp, dbl, i, vp, vp, vp, C.POINTER(i)]),
        "bddb200_incremental_mm_agreement_rounding": (i, [vp, vp, dbl, dbl, i, i, vp, C.POINTER(i), C.POINTER(i)]),
        "bddb200_lbfgs_create": (i, [vp, i, dbl, dbl, dbl, dbl, C.POINTER(vp)]),
        "bddb200_lbfgs_destroy": (None, [vp]),
        "bddb200_lbfgs_iteration": (i, [vp]),
        "bddb200_lbfgs_flush": (i, [vp]),
        "bddb200_lbfgs_stats": (i, [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(dbl)]),
        "bddb200_delta_sum_index": (i, [vp, C.POINTER(i)]),
        "bddb200_push_exchange_supported": (i, [vp, C.POINTER(i)]),
        "bddb200_set_push_masks": (i, [vp, vp, sz]),
        "bddb200_set_delta_buffers": (i, [vp, vp, vp, vp]),
        "bddb200_set_delta_input": (i, [vp, vp, sz]),
        "bddb200_set_exchange": (i, [vp, i, i, vp, vp, vp, vp, vp, vp, sz, i]),
        "bddb200_delta_exchange": (i, [vp, i, i, i, vp, vp, C.c_uint32, sz, vp, sz]),
        "bddb200_delta_exchange_two_shot": (i, [vp, i, i, i, vp, vp, vp, C.c_uint32, sz, sz]),
        "bddb200_collection_create": (i, [vp, sz, vp, sz, C.POINTER(vp)]),
        "bddb200_collection_destroy": (i, [vp]),
        "bddb200_collection_nr_bdds": (i, [vp, C.POINTER(sz)]),
        "bddb200_collection_nr_instructions": (i, [vp, C.POINTER(sz)]),
        "bddb200_collection_export": (i, [vp, vp, vp]),
        "bddb200_collection_simplex_constraint": (i, [vp, sz, C.POINTER(sz)]),
        "bddb200_collection_not_all_false_constraint": (i, [vp, sz, C.POINTER(sz)]),
        "bddb200_collection_all_equal_constraint": (i, [vp, sz, C.POINTER(sz)]),
        "bddb200_collection_cardinality_constraint": (i, [vp, sz, sz, C.POINTER(sz)]),
        "bddb200_collection_add_linear_constraint": (i, [vp, vp, vp, sz, i, C.c_longlong, C.POINTER(sz)]),
        "bddb200_collection_rebase": (i, [vp, sz, vp, sz]),
        "bddb200_collection_negate": (i, [vp, sz]),
        "bddb200_collection_invert": (i, [vp, sz, sz]),
        "bddb200_collection_variables": (i, [vp, sz, vp, sz, C.POINTER(sz)]),
        "bddb200_collection_is_qbdd": (i, [vp, sz, C.POINTER(i)]),
        "bddb200_collection_is_reordered": (i, [vp, sz, C.POINTER(i)]),
        "bddb200_collection_evaluate": (i, [vp, sz, vp, sz, C.POINTER(i)]),
        "bddb200_collection_reorder": (i, [vp, sz]),
        "bddb200_collection_make_qbdd": (i, [vp, sz, C.POINTER(sz)]),
        "bddb200_collection_bdd_and": (i, [vp, vp, sz, C.POINTER(sz)]),
        "bddb200_collection_remove": (i, [vp, vp, sz]),
        "bddb200_create_from_collection": (i, [vp, vp, sz, i, C.POINTER(Options), C.POINTER(vp)]),
        "bddb200_collection_write_bdd_lp": (i, [vp, vp, sz, C.c_char_p]),
        "bddb200_collection_split_qbdd": (i, [vp, sz, sz, sz, i, C.POINTER(sz), C.POINTER(sz)]),
        "bddb200_collection_split_long_bdds": (i, [vp, sz, sz, i, C.POINTER(sz), C.POINTER(sz)]),
        "bddb200_ilp_read": (i, [C.c_char_p, C.POINTER(vp)]),
        "bddb200_ilp_destroy": (i, [vp]),
        "bddb200_ilp_nr_variables": (i, [vp, C.POINTER(sz)]),
        "bddb200_ilp_nr_constraints": (i, [vp, C.POINTER(sz)]),
        "bddb200_ilp_objective": (i, [vp, vp, C.POINTER(dbl)]),
        "bddb200_ilp_variable_name": (i, [vp, sz, C.POINTER(C.c_char_p)]),
        "bddb200_ilp_constraint": (i, [vp, sz, C.POINTER(sz), vp, vp, sz, C.POINTER(i), C.POINTER(C.c_longlong)]),
        "bddb200_ilp_to_bdds": (i, [vp, C.POINTER(vp)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib


class BddB200Error(RuntimeError):
    """Raised for any non-zero bddb200_status (the reference throws std::runtime_error)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"bdd_b200 error {code}: {msg}")
        self.code = code


def check(code: int):
    if code != OK:
        raise BddB200Error(code, load().bddb200_last_error().decode())
