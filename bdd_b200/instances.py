"""Host-side construction of BDD collections and the synthetic benchmark instances.

The hot path consumes a ``BDD::bdd_collection`` (include/bdd_collection/bdd_collection.h):
a flat array of ``bdd_instruction{size_t lo, hi, index}`` plus ``bdd_delimiters``.  This
module produces exactly that flat form (``BddCollection``) without the reference:

* :func:`qbdd_template` turns one linear 0/1 constraint into its quasi-reduced ordered BDD
  (the canonical form the reference reaches through lineq_bdd -> bdd_mgr -> reorder ->
  make_qbdd, src/bdd_conversion/bdd_preprocessor.cpp:199-215): every arc goes to the next
  variable's layer or to the bot sink, the top sink is entered from the last layer only.
* :func:`from_constraints` instantiates cached templates for many constraints at once
  (vectorised with numpy, so 20 M-node collections build in seconds).
* the generators give the BASELINE.json configs: set cover,
  QAP-shaped assignment (Adams-Johnson linearisation, the shape
  src/specialized_solvers/graph_matching_input.cpp emits), pure assignment and grid MRF
  (src/specialized_solvers/mrf_input.cpp:100-161).

Input plumbing only; the sweep itself lives in csrc/.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .lp import EQ, GE, LE, ILP

TOPSINK = np.uint64(0xFFFFFFFFFFFFFFFF)  # bdd_instruction::topsink_index (bdd_collection.h:23)
BOTSINK = np.uint64(0xFFFFFFFFFFFFFFFE)  # bdd_instruction::botsink_index (bdd_collection.h:19)


@dataclass
class BddCollection:
    """Flat ``bdd_collection``: ``instrs[i] = (lo, hi, index)`` with absolute lo/hi, the
    last two instructions of every BDD being its two sinks (this module emits bot, top like
    bdd_collection.cpp:1581-1586; the reference's add_bdd may emit them in the other order and
    every consumer here accepts both); ``delims`` has nr_bdds+1 entries."""

    instrs: np.ndarray  # uint64 [n, 3]
    delims: np.ndarray  # uint64 [B + 1]

    @property
    def nr_bdds(self) -> int:
        return int(self.delims.shape[0] - 1)

    @property
    def nr_nodes(self) -> int:
        return int(self.instrs.shape[0])

    def nr_variables(self) -> int:
        idx = self.instrs[:, 2]
        return int(idx[idx < BOTSINK].max()) + 1

    def select(self, bdd_ids: Sequence[int]) -> "BddCollection":
        """Sub-collection with the given BDDs (what bdd_collection::remove leaves behind,
        bdd_collection.h:370-419); used to shard by constraint."""
        ids = np.asarray(bdd_ids, dtype=np.int64)
        if ids.size > 0 and np.array_equal(ids, np.arange(ids[0], ids[0] + ids.size)):
            # contiguous block: one slice, one shift
            first, last = int(self.delims[ids[0]]), int(self.delims[ids[-1] + 1])
            blk = self.instrs[first:last].copy()
            inner = blk[:, 2] < BOTSINK
            blk[inner, 0] -= np.uint64(first)
            blk[inner, 1] -= np.uint64(first)
            return BddCollection(blk, self.delims[ids[0]: ids[-1] + 2] - np.uint64(first))
        # general case, vectorised: gather the instruction ranges and shift the child indices of inner nodes
        sizes = np.diff(self.delims.astype(np.int64))
        sz = sizes[ids]
        new_delims = np.concatenate([[0], np.cumsum(sz)])
        old_first = self.delims.astype(np.int64)[ids]
        src = np.repeat(old_first - new_delims[:-1], sz) + np.arange(int(new_delims[-1]), dtype=np.int64)
        blk = self.instrs[src].copy()
        shift = np.repeat(new_delims[:-1] - old_first, sz)
        inner = blk[:, 2] < BOTSINK
        blk[inner, 0] = (blk[inner, 0].astype(np.int64) + shift[inner]).astype(np.uint64)
        blk[inner, 1] = (blk[inner, 1].astype(np.int64) + shift[inner]).astype(np.uint64)
        return BddCollection(blk, new_delims.astype(np.uint64))


def bdds_accept(col: "BddCollection", sol: Sequence[int]) -> np.ndarray:
    """For every BDD: does the 0/1 assignment ``sol`` (indexed by variable) lead from the root to the top sink?
    (A primal solution is feasible iff every constraint's BDD accepts it.)  Plain loop: for tests and small instances."""
    instrs = col.instrs.astype(np.int64)      # sinks wrap to -1 (top) / -2 (bot)
    ok = np.zeros(col.nr_bdds, dtype=bool)
    for b in range(col.nr_bdds):
        i = int(col.delims[b])
        while True:
            idx = instrs[i, 2]
            if idx == -1:
                ok[b] = True
                break
            if idx == -2:
                break
            i = int(instrs[i, 1] if sol[idx] else instrs[i, 0])
    return ok


# --------------------------------------------------------------------------- templates --

@dataclass
class QbddTemplate:
    """One BDD with local numbering: node ``i`` branches on position ``layer[i]`` of the
    constraint's variable list; children are local node ids, ``-1`` = bot sink, ``-2`` =
    top sink.  Nodes are ordered layer by layer."""

    layer: np.ndarray  # int64 [n]
    lo: np.ndarray  # int64 [n]
    hi: np.ndarray  # int64 [n]


MAX_PARTIAL_SUMS = 1 << 20     # the direct builder enumerates the reachable partial sums of a layer; a solver layer holds at most 65 504 nodes anyway


def qbdd_template(coeffs: Sequence[int], ineq: int, rhs: int) -> Optional[QbddTemplate]:
    """Quasi-reduced BDD of ``sum_i coeffs[i] * x_i  (<=|>=|=)  rhs`` in the given variable
    order.  Returns ``None`` if the constraint is always satisfied; raises if infeasible.

    Top-down over partial sums, then bottom-up merging of states with identical
    sub-functions (same (lo, hi) pair in the same layer).  This is the canonical
    quasi-reduced form, which is unique for a fixed variable order.
    """
    a = [int(c) for c in coeffs]
    n = len(a)
    if n == 0:
        raise ValueError("empty constraint")
    # reachable partial sums per layer
    sums: List[List[int]] = [[0]]
    for k in range(n):
        nxt = set()
        for s in sums[k]:
            nxt.add(s)
            nxt.add(s + a[k])
        if len(nxt) > MAX_PARTIAL_SUMS:
            raise ValueError(f"constraint with more than {MAX_PARTIAL_SUMS} distinct partial sums in a layer: outside the direct BDD builder's scope "
                             "(the reference converts such knapsack rows through its BDD manager, test/hard_ineqs.h)")
        sums.append(sorted(nxt))

    def accept(s: int) -> bool:
        return s <= rhs if ineq == LE else (s >= rhs if ineq == GE else s == rhs)

    BOT, TOP = -1, -2
    # ids of the sub-function of every state, bottom-up
    ident: List[Dict[int, int]] = [dict() for _ in range(n + 1)]
    for s in sums[n]:
        ident[n][s] = TOP if accept(s) else BOT
    nodes_per_layer: List[List[Tuple[int, int]]] = [[] for _ in range(n)]
    for k in range(n - 1, -1, -1):
        table: Dict[Tuple[int, int], int] = {}
        for s in sums[k]:
            key = (ident[k + 1][s], ident[k + 1][s + a[k]])
            if key == (BOT, BOT):
                ident[k][s] = BOT
                continue
            if key not in table:
                table[key] = len(nodes_per_layer[k])
                nodes_per_layer[k].append(key)
            ident[k][s] = table[key]
    if ident[0][0] == BOT:
        raise ValueError("problem is infeasible")
    # every emitted node is reachable; the function is constant true iff no arc enters the bot sink
    if not any(BOT in key for nl in nodes_per_layer for key in nl):
        return None
    # A variable the function does not depend on (every node of its layer has lo == hi) gets
    # no layer at all: the reference's reduced BDD does not contain it and make_qbdd only
    # bridges between the variables that are present.  Splice such layers out bottom-up.
    keep = [not all(l == h for (l, h) in nl) for nl in nodes_per_layer]
    resolve: List[Dict[int, int]] = [dict() for _ in range(n + 1)]   # node id in layer k -> (layer, id) it stands for
    new_ids: List[Dict[int, int]] = [dict() for _ in range(n)]
    target: List[List[Tuple[int, int]]] = [[] for _ in range(n)]     # per kept layer: children as (code or (layer,id))
    def stands_for(k: int, i: int):
        """(layer, id) of the first kept node reached from node i of layer k, or a terminal code."""
        while True:
            if i < 0:
                return i
            if keep[k]:
                return (k, i)
            i = nodes_per_layer[k][i][0]
            k += 1
    kept_layers = [k for k in range(n) if keep[k]]
    offsets = {}
    off = 0
    for k in kept_layers:
        offsets[k] = off
        off += len(nodes_per_layer[k])
    layer, lo, hi = [], [], []
    for k in kept_layers:
        for (l, h) in nodes_per_layer[k]:
            cl = stands_for(k + 1, l) if l >= 0 else l
            ch = stands_for(k + 1, h) if h >= 0 else h
            layer.append(k)
            lo.append(cl if isinstance(cl, int) else offsets[cl[0]] + cl[1])
            hi.append(ch if isinstance(ch, int) else offsets[ch[0]] + ch[1])
    del resolve, new_ids, target
    return QbddTemplate(np.asarray(layer, np.int64), np.asarray(lo, np.int64), np.asarray(hi, np.int64))


def _instantiate(t: QbddTemplate, var_lists: np.ndarray, first_instr: int) -> Tuple[np.ndarray, np.ndarray]:
    """Instantiate template ``t`` for every row of ``var_lists`` [m, n_vars]; the group
    occupies instructions [first_instr, first_instr + m*(n+2))."""
    m = var_lists.shape[0]
    nn = t.layer.shape[0]
    stride = nn + 2
    base = first_instr + np.arange(m, dtype=np.int64)[:, None] * stride  # [m,1]
    out = np.empty((m, stride, 3), dtype=np.uint64)

    def child(c: np.ndarray) -> np.ndarray:
        c = np.broadcast_to(c[None, :], (m, nn))
        r = base + c
        r = np.where(c == -1, base + nn, r)  # bot sink is the second to last instruction
        r = np.where(c == -2, base + nn + 1, r)
        return r.astype(np.uint64)

    out[:, :nn, 0] = child(t.lo)
    out[:, :nn, 1] = child(t.hi)
    out[:, :nn, 2] = var_lists[:, t.layer].astype(np.uint64)
    out[:, nn, :] = BOTSINK
    out[:, nn + 1, :] = TOPSINK
    delims = (base[:, 0] + stride).astype(np.uint64)
    return out.reshape(m * stride, 3), delims


class ConstraintBatch:
    """Constraints with identical (coefficients, relation, rhs), differing only in their
    variables: ``variables`` is an [m, n] integer array."""

    def __init__(self, coeffs: Sequence[int], ineq: int, rhs: int, variables: np.ndarray):
        self.coeffs = tuple(int(c) for c in coeffs)
        self.ineq = int(ineq)
        self.rhs = int(rhs)
        self.variables = np.asarray(variables, dtype=np.int64).reshape(-1, len(self.coeffs))


def from_batches(batches: Sequence[ConstraintBatch]) -> BddCollection:
    cache: Dict[Tuple, Optional[QbddTemplate]] = {}
    parts, delims = [], [np.zeros(1, dtype=np.uint64)]
    first = 0
    for b in batches:
        key = (b.coeffs, b.ineq, b.rhs)
        if key not in cache:
            cache[key] = qbdd_template(b.coeffs, b.ineq, b.rhs)
        t = cache[key]
        if t is None or b.variables.shape[0] == 0:
            continue
        instrs, d = _instantiate(t, b.variables, first)
        parts.append(instrs)
        delims.append(d)
        first += instrs.shape[0]
    return BddCollection(np.concatenate(parts, axis=0), np.concatenate(delims))


def from_constraints(constraints) -> BddCollection:
    """One BDD per constraint, in constraint order (objects with ``variables``,
    ``coefficients``, ``ineq``, ``rhs``, e.g. :class:`bdd_b200.lp.Constraint`)."""
    return from_batches([ConstraintBatch(c.coefficients, c.ineq, c.rhs, np.asarray([c.variables])) for c in constraints])


def from_ilp(ilp: ILP) -> Tuple[BddCollection, np.ndarray]:
    return from_constraints(ilp.constraints), np.asarray(ilp.objective, dtype=np.float64)


# -------------------------------------------------------------------------- generators --

def set_cover(m: int = 25000, n: int = 50000, k: int = 20, seed: int = 1) -> Tuple[BddCollection, np.ndarray]:
    """BASELINE config 2: ``m`` rows ``sum_{j in row} x_j >= 1`` with ``k``
    distinct columns each; column j is first placed in row j mod m so every variable is
    covered; costs uniform integer in [1, 100].  k=20 gives 41 nodes per BDD."""
    rng = np.random.default_rng(seed)
    rows = np.empty((m, k), dtype=np.int64)
    fixed = np.full((m, k), -1, dtype=np.int64)
    cnt = np.zeros(m, dtype=np.int64)
    for j in range(n):
        r = j % m
        if cnt[r] < k:
            fixed[r, cnt[r]] = j
            cnt[r] += 1
    for r in range(m):
        have = fixed[r, : cnt[r]]
        need = k - cnt[r]
        if need > 0:
            extra = rng.choice(n, size=need + cnt[r], replace=False)
            extra = extra[~np.isin(extra, have)][:need]
            rows[r] = np.concatenate([have, extra])
        else:
            rows[r] = have
        rows[r].sort()
    costs = rng.integers(1, 101, size=n).astype(np.float64)
    col = from_batches([ConstraintBatch([1] * k, GE, 1, rows)])
    return col, costs


def assignment(n: int = 1118, seed: int = 3) -> Tuple[BddCollection, np.ndarray]:
    """Config 3b (stress): n x n assignment, 2n simplex BDDs of n variables (H = n)."""
    rng = np.random.default_rng(seed)
    x = np.arange(n * n, dtype=np.int64).reshape(n, n)
    costs = rng.integers(0, 100, size=n * n).astype(np.float64)
    col = from_batches([ConstraintBatch([1] * n, EQ, 1, x), ConstraintBatch([1] * n, EQ, 1, x.T.copy())])
    return col, costs


def qap(n: int = 36, seed: int = 2) -> Tuple[BddCollection, np.ndarray]:
    """Config 3: QAP-shaped assignment ILP (Adams-Johnson linearisation).
    Variables x_ip (n^2) then y_{ipjq} for i<j, p!=q.  Constraints: 2n simplex rows on x and,
    for every ordered (i, j != i, p):  -x_ip + sum_{q != p} y_{ipjq} = 0  (x first)."""
    rng = np.random.default_rng(seed)
    f = rng.integers(0, 10, size=(n, n)); f = np.triu(f, 1); f = f + f.T
    d = rng.integers(0, 10, size=(n, n)); d = np.triu(d, 1); d = d + d.T
    x = np.arange(n * n, dtype=np.int64).reshape(n, n)
    # y index for i<j, p, q!=p
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    pair_id = -np.ones((n, n), dtype=np.int64)
    for k, (i, j) in enumerate(pairs):
        pair_id[i, j] = k
    ybase = n * n
    # y[(i,j), p, q] stored densely n*n per pair including p==q slots (unused -> compacted below)
    pq = np.arange(n * n, dtype=np.int64).reshape(n, n)
    valid = ~np.eye(n, dtype=bool)
    compact = -np.ones(n * n, dtype=np.int64)
    compact[pq[valid]] = np.arange(n * (n - 1))
    per_pair = n * (n - 1)

    def yidx(i: int, j: int, p: np.ndarray, q: np.ndarray) -> np.ndarray:
        # y_{ipjq} with i<j
        return ybase + pair_id[i, j] * per_pair + compact[pq[p, q]]

    nvars = ybase + len(pairs) * per_pair
    costs = np.zeros(nvars, dtype=np.float64)
    pp, qq = np.nonzero(valid)
    for (i, j) in pairs:
        costs[yidx(i, j, pp, qq)] = f[i, j] * d[pp, qq]
    marg = np.empty((n * (n - 1) * n, n), dtype=np.int64)
    r = 0
    ar = np.arange(n)
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            for p in range(n):
                q = ar[ar != p]
                if i < j:
                    ys = yidx(i, j, np.full(n - 1, p), q)
                else:
                    ys = yidx(j, i, q, np.full(n - 1, p))  # y_{jq ip}: j takes q, i takes p
                marg[r, 0] = x[i, p]
                marg[r, 1:] = ys
                r += 1
    col = from_batches([
        ConstraintBatch([1] * n, EQ, 1, x),
        ConstraintBatch([1] * n, EQ, 1, x.T.copy()),
        ConstraintBatch([-1] + [1] * (n - 1), EQ, 0, marg),
    ])
    return col, costs


def grid_mrf(w: int = 260, h: int = 260, K: int = 4, seed: int = 4) -> Tuple[BddCollection, np.ndarray]:
    """Config 4: MRF on a w x h 4-connected grid with K labels in the local-polytope ILP form
    the reference's MRF front end emits (mrf_input.cpp:100-161): per vertex a simplex over K
    unaries; per edge a simplex over K^2 pairwise variables and 2K marginalisation equalities
    ``mu_u(a) - sum_b mu_uv(a,b) = 0``.  Constraints are emitted in grid-tile order (vertex,
    then its right and down edges) so that sharding by constraint index keeps shards local."""
    rng = np.random.default_rng(seed)
    nv = w * h
    vert = np.arange(nv, dtype=np.int64).reshape(h, w)
    unary = vert[:, :, None] * K + np.arange(K)[None, None, :]
    edges = []
    for dy, dx in ((0, 1), (1, 0)):
        u = vert[: h - dy, : w - dx].reshape(-1)
        v = vert[dy:, dx:].reshape(-1)
        edges.append(np.stack([u, v], axis=1))
    edges = np.concatenate(edges, axis=0)
    ne = edges.shape[0]
    pair = nv * K + np.arange(ne, dtype=np.int64)[:, None, None] * K * K + (np.arange(K)[None, :, None] * K + np.arange(K)[None, None, :])
    nvars = nv * K + ne * K * K
    costs = np.zeros(nvars, dtype=np.float64)
    costs[: nv * K] = rng.integers(-10, 11, size=nv * K)
    potts = rng.integers(0, 6, size=ne).astype(np.float64)
    pc = np.where(np.eye(K, dtype=bool)[None], 0.0, potts[:, None, None])
    costs[nv * K:] = pc.reshape(-1)
    u_un = unary.reshape(nv, K)
    m_u = np.concatenate([u_un[edges[:, 0]][:, :, None], pair], axis=2).reshape(ne * K, K + 1)
    m_v = np.concatenate([u_un[edges[:, 1]][:, :, None], pair.transpose(0, 2, 1)], axis=2).reshape(ne * K, K + 1)
    col = from_batches([
        ConstraintBatch([1] * K, EQ, 1, u_un),
        ConstraintBatch([1] * (K * K), EQ, 1, pair.reshape(ne, K * K)),
        ConstraintBatch([1] + [-1] * K, EQ, 0, m_u),
        ConstraintBatch([1] + [-1] * K, EQ, 0, m_v),
    ])
    # grid-tile order: every constraint is keyed by its (first) vertex, row-major
    key = np.concatenate([np.arange(nv), edges[:, 0], np.repeat(edges[:, 0], K), np.repeat(edges[:, 0], K)])
    return col.select(np.argsort(key, kind="stable")), costs


def _both_values_feasible(t: QbddTemplate) -> bool:
    """True iff no variable of the constraint is forced: every layer has a node whose lo arc
    and a node whose hi arc avoids the bot sink (all nodes of a reduced BDD reach the top sink)."""
    nl = int(t.layer.max()) + 1
    present = np.zeros(nl, dtype=bool)
    present[t.layer] = True
    lo_ok = np.zeros(nl, dtype=bool)
    hi_ok = np.zeros(nl, dtype=bool)
    np.logical_or.at(lo_ok, t.layer, t.lo != -1)
    np.logical_or.at(hi_ok, t.layer, t.hi != -1)
    return bool((lo_ok | ~present).all() and (hi_ok | ~present).all())


def random_inequalities(nr_constraints: int, nr_vars: int, max_len: int = 8, max_coeff: int = 4, seed: int = 0) -> Tuple[BddCollection, np.ndarray]:
    """Random small knapsack-like constraints with wider, irregular BDDs (parity stress; in
    the spirit of test/test_problem_generator.h:10-101).  Every variable is covered."""
    rng = np.random.default_rng(seed)
    batches = []
    covered = np.zeros(nr_vars, dtype=bool)
    c = 0
    while c < nr_constraints or not covered.all():
        ln = int(rng.integers(2, max_len + 1))
        missing = np.nonzero(~covered)[0]
        if missing.size >= ln:
            vs = rng.choice(missing, size=ln, replace=False)
        else:
            vs = rng.choice(nr_vars, size=ln, replace=False)
        vs.sort()
        coeffs = rng.integers(1, max_coeff + 1, size=ln) * rng.choice([-1, 1], size=ln)
        # rhs chosen so that both x=0...0 is not forced and the constraint is satisfiable
        lo_sum = int(coeffs[coeffs < 0].sum()); hi_sum = int(coeffs[coeffs > 0].sum())
        rhs = int(rng.integers(lo_sum, hi_sum + 1))
        ineq = int(rng.integers(0, 3))
        try:
            t = qbdd_template(coeffs.tolist(), ineq, rhs)
        except ValueError:
            continue
        if t is None or not _both_values_feasible(t):
            continue
        batches.append(ConstraintBatch(coeffs.tolist(), ineq, rhs, vs[None, :]))
        covered[vs] = True
        c += 1
    costs = rng.integers(-10, 11, size=nr_vars).astype(np.float64)
    return from_batches(batches), costs
