"""Splitting long BDDs into chunks linked by auxiliary variables -- the reference's answer to constraints over
thousands of variables, whose BDDs are long and thin and make every pass a chain of thousands of dependent hops
(SURVEY 5 "long BDDs", 8f row 4).

``split_qbdd`` follows the chunk construction of ``bdd_collection::split_qbdd``
(src/bdd_collection/bdd_collection.cpp:507-790; the optional implication BDD, :805-940, is built by the library's collection class,
``bdd_b200.collection.bdd_collection.split_qbdd(..., with_implication_bdd=True)``); ``split_long_bdds`` follows the driver loop of
``bdd_preprocessor`` (src/bdd_conversion/bdd_preprocessor.cpp:372-415).  Cutting a quasi-reduced BDD in front of a layer
of width w introduces w auxiliary 0/1 variables that one-hot encode which node of that layer the path goes through
(aux variable w-1-k is 1 iff node k is used): the chunk before the cut ends in a *tail* gadget that accepts exactly the
one-hot pattern of the node each path reached, the chunk after it starts with a *head* gadget that routes the pattern to
that node.  Auxiliary variables carry no cost.  The instruction arrays produced here equal the reference's bit for bit
(tests/test_split.py checks against oracle/_ref).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .instances import BOTSINK, TOPSINK, BddCollection


def _layers(instrs: np.ndarray, first: int, last: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """variables, absolute layer offsets and layer widths of the BDD occupying instrs[first:last] (last two = sinks)"""
    idx = instrs[first:last - 2, 2]
    head = np.ones(idx.shape[0], dtype=bool)
    head[1:] = idx[1:] != idx[:-1]
    offs = np.nonzero(head)[0] + first
    widths = np.diff(np.append(offs, last - 2))
    return idx[head].astype(np.int64), offs.astype(np.int64), widths.astype(np.int64)


def split_qbdd(instrs: np.ndarray, first: int, last: int, chunk_size: int, aux_var_start: int, base: int, sinks=None) -> Tuple[List[np.ndarray], int]:
    """Chunk BDDs of the QBDD instrs[first:last] as instruction arrays positioned at ``base`` (absolute indices), and the
    next free auxiliary variable.  A BDD of at most chunk_size variables is returned unsplit (empty list)."""
    assert chunk_size > 0
    variables, layer_offsets, layer_widths = _layers(instrs, first, last)
    n_layers = variables.shape[0]
    if n_layers <= chunk_size:
        return [], aux_var_start
    is_bot, is_top = sinks if sinks is not None else (instrs[:, 2] == BOTSINK, instrs[:, 2] == TOPSINK)
    nr_chunks = (n_layers + chunk_size - 1) // chunk_size
    aux_vars = [aux_var_start]
    for c in range(1, nr_chunks - 1):
        aux_vars.append(aux_vars[-1] + int(layer_widths[c * chunk_size]))

    def layer_offset(layer: int) -> int:
        return int(layer_offsets[layer]) if layer < n_layers else last - 2

    out: List[np.ndarray] = []
    for c in range(nr_chunks):
        first_layer = c * chunk_size
        last_layer = min((c + 1) * chunk_size - 1, n_layers - 1)
        w_head = int(layer_widths[first_layer]) if c > 0 else 0
        n_head = (w_head * (w_head + 1)) // 2
        n_body = (last - 2 if c + 1 == nr_chunks else int(layer_offsets[last_layer + 1])) - int(layer_offsets[first_layer])
        w_tail = int(layer_widths[last_layer + 1]) if c + 1 < nr_chunks else 0
        if c + 1 < nr_chunks and w_tail <= 1:
            raise ValueError("split_qbdd: cannot cut in front of a layer of width 1 (bdd_collection.cpp:598)")
        n_tail = (w_tail * (w_tail + 1)) // 2 + w_tail - 1 if c + 1 < nr_chunks else 0
        bottom = base + n_head + n_body + n_tail
        top = bottom + 1
        rows: List[Tuple[int, int, int]] = []        # (lo, hi, index)

        def H(i: int, j: int) -> int:
            return base + (i * (i + 1)) // 2 + j

        def T(i: int, j: int) -> int:
            if i == 0:
                return base + n_head + n_body + j
            return base + n_head + n_body + w_tail + w_tail * (i - 1) + j - ((i - 1) * (i - 2)) // 2

        # 1) head gadget: routes the one-hot pattern of the cut in front of this chunk to the node it stands for
        if c > 0:
            av = aux_vars[c - 1]
            for i in range(w_head - 1):
                for j in range(i + 1):
                    if j == 0:
                        rows.append((H(i + 1, 0), H(i + 1, 1), av + i))
                    else:
                        rows.append((H(i + 1, j + 1), bottom, av + i))
            i = w_head - 1
            for j in range(i + 1):
                if j == 0:
                    rows.append((bottom, base + n_head + j, av + i))
                else:
                    rows.append((base + n_head + j, bottom, av + i))
        assert len(rows) == n_head

        # 2) the chunk's own nodes keep their relative positions; arcs into the layer after the chunk land in the tail gadget
        b0, b1 = layer_offset(first_layer), layer_offset(last_layer + 1)
        body = instrs[b0:b1].astype(np.int64)
        pos = base + n_head + np.arange(b1 - b0, dtype=np.int64)
        orig = np.arange(b0, b1, dtype=np.int64)
        for col in (0, 1):
            child = body[:, col]
            new = pos + (child - orig)
            new = np.where(is_bot[child], bottom, np.where(is_top[child], top, new))
            body[:, col] = new
        body_rows = body

        # 3) tail gadget: accepts exactly the one-hot pattern of the node of the next layer that the path reached
        tail: List[Tuple[int, int, int]] = []
        if c + 1 < nr_chunks:
            av = aux_vars[c]
            for j in range(w_tail):
                if j + 1 == w_tail:
                    tail.append((bottom, T(1, w_tail - 1), av))
                else:
                    tail.append((T(1, j), bottom, av))
            for i in range(1, w_tail - 1):
                for j in range(w_tail - i + 1):
                    if j + 1 == w_tail - i + 1:
                        tail.append((T(i + 1, w_tail - i - 1), bottom, av + i))
                    elif j + 1 == w_tail - i:
                        tail.append((bottom, T(i + 1, j), av + i))
                    else:
                        tail.append((T(i + 1, j), bottom, av + i))
            tail.append((bottom, top, av + w_tail - 1))
            tail.append((top, bottom, av + w_tail - 1))
        assert len(tail) == n_tail

        parts = []
        if rows:
            parts.append(np.asarray(rows, dtype=np.int64).reshape(-1, 3))
        parts.append(body_rows)
        if tail:
            parts.append(np.asarray(tail, dtype=np.int64).reshape(-1, 3))
        sinks = np.array([[-2, -2, -2], [-1, -1, -1]], dtype=np.int64)      # bot sink first, then top sink (:768-771); lo = hi = index
        arr = np.concatenate(parts + [sinks], axis=0).astype(np.uint64)
        out.append(arr)
        base += arr.shape[0]
    next_aux = aux_vars[-1] + int(layer_widths[(nr_chunks - 1) * chunk_size])
    return out, next_aux


def split_long_bdds(col: BddCollection, split_length: int, nr_variables: int = 0) -> Tuple[BddCollection, int]:
    """bdd_preprocessor.cpp:372-415 with a forced split length: every BDD over more than ``split_length`` variables is replaced
    by its chunks (appended at the end, in BDD order), auxiliary variables are numbered from ``nr_variables`` (default: the
    collection's).  Returns the new collection and the total number of variables including the auxiliary ones."""
    instrs = np.ascontiguousarray(col.instrs, dtype=np.uint64)
    delims = col.delims.astype(np.int64)
    aux = max(int(nr_variables), col.nr_variables())
    base = int(delims[-1])
    new_arrays: List[np.ndarray] = []
    removed = np.zeros(col.nr_bdds, dtype=bool)
    sinks = (instrs[:, 2] == BOTSINK, instrs[:, 2] == TOPSINK)
    for b in range(col.nr_bdds):
        first, last = int(delims[b]), int(delims[b + 1])
        try:
            chunks, aux_next = split_qbdd(instrs, first, last, split_length, aux, base, sinks)
        except ValueError as e:
            # a cut would land in front of a layer of width 1 (the reference asserts, bdd_collection.cpp:598): this BDD stays whole
            import warnings
            warnings.warn(f"split_long_bdds: BDD {b} left unsplit ({e})")
            continue
        aux = aux_next
        if len(chunks) > 1:
            removed[b] = True
            new_arrays.extend(chunks)
            base += sum(c.shape[0] for c in chunks)
    if not removed.any():
        return col, aux
    all_instrs = np.concatenate([instrs] + new_arrays, axis=0)
    all_delims = np.concatenate([delims, delims[-1] + np.cumsum([c.shape[0] for c in new_arrays])]).astype(np.uint64)
    whole = BddCollection(all_instrs, all_delims)
    keep = np.concatenate([np.nonzero(~removed)[0], np.arange(col.nr_bdds, whole.nr_bdds)])
    return whole.select(keep), aux


def _device_sm_count(default: int = 148) -> int:
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count)
    except Exception:
        pass
    return default


def compute_split_length(col: BddCollection, n_sms: int = 0, warps_per_sm: int = 16, min_length: int = 16) -> int:
    """Split length when the configuration gives none.  The reference's rule (compute_split_length, bdd_preprocessor.cpp:32-121) targets
    >= 50 % occupancy of its hop-synchronous kernels (nodes per hop against SMs x threads); the sweep here gives every warp a bundle of 32
    BDDs and walks it alone, so what long BDDs cost is (a) too few bundles to put ``warps_per_sm`` warps on every SM and (b) a pass that is
    a chain of H dependent hops.  Rule: the largest length that yields at least ``n_sms * warps_per_sm`` bundles, never below ``min_length``
    (every cut adds a head and a tail gadget and width-many auxiliary variables); collections that already fill the GPU are not split.
    Returns sys.maxsize when nothing should be split."""
    import sys
    if n_sms <= 0:
        n_sms = _device_sm_count()          # the SM count of the current device (148 on B200 and where no device is visible)
    sizes = np.diff(col.delims.astype(np.int64)) - 2
    idx = col.instrs[:, 2]
    inner = idx < BOTSINK
    bdd_of = np.repeat(np.arange(col.nr_bdds), np.diff(col.delims.astype(np.int64)))[inner]
    var = idx[inner]
    head = np.ones(var.shape[0], dtype=bool)
    head[1:] = (var[1:] != var[:-1]) | (bdd_of[1:] != bdd_of[:-1])
    layers = np.bincount(bdd_of[head], minlength=col.nr_bdds)          # layers (variables) per BDD
    target_bdds = 32 * n_sms * warps_per_sm
    if col.nr_bdds >= target_bdds or layers.max() <= min_length:
        return sys.maxsize
    # number of chunk BDDs for length L: sum over BDDs of ceil(layers / L); find the largest L reaching the target
    lo, hi = min_length, int(layers.max())
    if int(np.ceil(layers / lo).sum()) < target_bdds:
        return lo
    while lo < hi:
        mid = (lo + hi + 1) // 2
        if int(np.ceil(layers / mid).sum()) >= target_bdds:
            lo = mid
        else:
            hi = mid - 1
    del sizes
    return lo
