"""CPU restatement of the reference's random-walk path (TEST INFRASTRUCTURE).

Restates, function by function, what ``node2vec/randomwalk.py`` and the step loop of
``node2vec/fugue.py`` compute, in plain Python (small cases only; the C twin in
``oracle/csrc/n2v_oracle.c`` covers the big ones).  Each function cites the reference
lines it follows.  Pinned against the reference by ``tests/test_oracle_golden.py``
using fixtures that ``tests/golden/make_golden.py`` produced by importing the
unmodified reference -- its transformer functions directly (``walks.json``) and its entry
points ``node2vec.fugue.random_walk`` / ``trim_index`` verbatim on a pandas stand-in for
Fugue (``fugue_verbatim.json``, ``tests/golden/fugue_shim``).

A note on ``sum()``: the reference divides by ``sum(node_weights) / n``
(``randomwalk.py:172``).  On the interpreters the reference supports (3.6/3.7,
``setup.py:29-30``) ``sum`` is a plain left-to-right fp64 accumulation; CPython >= 3.12
switched float ``sum`` to Neumaier compensated summation, which changes the last bit
of ``probs`` in about half of all random weight vectors.  Both are restated
(``sum_mode="naive"`` is the reference-as-shipped and the product default,
``"neumaier"`` is the reference-as-run-under-3.12) and both are pinned by fixtures.
"""
import math
import random as _random
from typing import Any, Dict, Iterable, List, Optional, Sequence, Set, Tuple

SUM_MODES = ("naive", "neumaier")


# ----------------------------------------------------------------------------------
# builtins.sum over floats, as the two interpreter generations evaluate it
# ----------------------------------------------------------------------------------
def float_sum(values: Sequence[float], mode: str = "naive") -> float:
    """``sum(values)`` for a list of floats starting from int 0 (randomwalk.py:172).

    naive:    ((0 + v0) + v1) + ...            (CPython <= 3.11)
    neumaier: CPython >= 3.12 ``builtin_sum`` float fast path (Kahan-Babuska/Neumaier
              compensation, added to the total at the end when finite and non-zero).
    """
    if mode not in SUM_MODES:
        raise ValueError(f"unknown sum mode {mode!r}")
    it = iter(values)
    try:
        total = 0 + float(next(it))
    except StopIteration:
        return 0
    if mode == "naive":
        for v in it:
            total = total + float(v)
        return total
    comp = 0.0
    for v in it:
        v = float(v)
        t = total + v
        if abs(total) >= abs(v):
            comp += (total - t) + v
        else:
            comp += (v - t) + total
        total = t
    if comp and math.isfinite(comp):
        total += comp
    return total


# ----------------------------------------------------------------------------------
# alias tables
# ----------------------------------------------------------------------------------
def alias_tables(weights: Sequence[float], sum_mode: str = "naive") -> Tuple[List[int], List[float]]:
    """Walker/Vose alias tables, LIFO work-lists (randomwalk.py:157-190).

    Order-sensitive details that make the result bit-exact:
      * mean = sum / n, probs[i] = w[i] / mean                       (:172-173)
      * the two work-lists are filled in index order, `< 1.0` decides (:175-180)
      * both are popped from the END; the donor is pushed back       (:182-189)
      * donor update is (probs[over] + probs[under]) - 1.0 in fp64   (:185)
      * entries never popped keep alias 0 and their residue          (:171)
    Raises ZeroDivisionError for an empty or all-zero vector, as the reference does.
    """
    n = len(weights)
    mean = float_sum(weights, sum_mode) / n
    probs = [w / mean for w in weights]
    alias = [0] * n
    small: List[int] = []
    large: List[int] = []
    for i, pr in enumerate(probs):
        (small if pr < 1.0 else large).append(i)
    while small and large:
        lo = small.pop()
        hi = large.pop()
        alias[lo] = hi
        probs[hi] = probs[hi] + probs[lo] - 1.0
        (small if probs[hi] < 1.0 else large).append(hi)
    return alias, probs


def biased_weights(
    prev_id: int,
    prev_out: Set[int],
    nbr_ids: Sequence[int],
    nbr_wts: Sequence[float],
    return_param: float = 1.0,
    inout_param: float = 1.0,
) -> List[float]:
    """Second-order (p, q) re-weighting of the current vertex's out-arcs
    (randomwalk.py:219-231).  Tested in this order: back to prev -> w/p; into
    N_out(prev) -> w; anywhere else -> w/q."""
    if len(nbr_ids) != len(nbr_wts):
        raise ValueError(f"Invalid neighbors tuple '{(nbr_ids, nbr_wts)}'!")
    if return_param == 0 or inout_param == 0:
        raise ValueError(f"Zero return ({return_param}) or inout ({inout_param}) parameter!")
    out = []
    for x, w in zip(nbr_ids, nbr_wts):
        if x == prev_id:
            out.append(w / return_param)
        elif x in prev_out:
            out.append(w)
        else:
            out.append(w / inout_param)
    return out


def edge_alias_tables(
    prev_id: int,
    prev_out: Set[int],
    nbrs: Tuple[Sequence[int], Sequence[float]],
    return_param: float = 1.0,
    inout_param: float = 1.0,
    sum_mode: str = "naive",
) -> Tuple[List[int], List[float]]:
    """randomwalk.py:193-232: validate, bias, then build alias tables."""
    if len(nbrs) != 2 or len(nbrs[0]) != len(nbrs[1]):
        raise ValueError(f"Invalid neighbors tuple '{nbrs}'!")
    return alias_tables(
        biased_weights(prev_id, prev_out, nbrs[0], nbrs[1], return_param, inout_param), sum_mode
    )


# ----------------------------------------------------------------------------------
# samplers and path extension
# ----------------------------------------------------------------------------------
def draw_two_uniform(alias: Sequence[int], probs: Sequence[float], r1: float, r2: float) -> int:
    """AliasProb.sampling_from_alias (randomwalk.py:86-99) -- the production sampler."""
    k = int(r1 * len(alias))
    return k if r2 < probs[k] else alias[k]


def draw_one_uniform(alias: Sequence[int], probs: Sequence[float], r1: float) -> int:
    """AliasProb.sampling_from_alias_wiki (randomwalk.py:70-84) -- tests only."""
    n = len(alias)
    k = int(n * r1)
    frac = n * r1 - k
    return k if frac < probs[k] else alias[k]


def extend_path(
    path: Sequence[int],
    nbr_ids: Sequence[int],
    alias: Sequence[int],
    probs: Sequence[float],
    r1: float,
    r2: Optional[float] = None,
) -> List[int]:
    """RandomPath.append (randomwalk.py:123-153).  A fresh walker's path is
    ``[-i, v]``; its first step rewrites that to ``[v, x]`` (:148-149)."""
    k = draw_one_uniform(alias, probs, r1) if r2 is None else draw_two_uniform(alias, probs, r1, r2)
    nxt = nbr_ids[k]
    path = list(path)
    if len(path) == 2 and path[0] < 0:
        return [path[1], nxt]
    path.append(nxt)
    return path


# ----------------------------------------------------------------------------------
# the transformer functions, on already-decoded adjacency
# ----------------------------------------------------------------------------------
Adjacency = Dict[int, Tuple[List[int], List[float]]]


def build_adjacency(src: Sequence[int], dst: Sequence[int], wt: Sequence[float]) -> Adjacency:
    """``partition(by=["src"], presort="dst")`` + get_vertex_neighbors
    (fugue.py:130, randomwalk.py:266-275): one (ids, weights) pair per vertex with at
    least one out-arc, ordered by dst; equal (src, dst) arcs keep input order."""
    order = sorted(range(len(src)), key=lambda i: (src[i], dst[i]))
    adj: Adjacency = {}
    for i in order:
        ids, wts = adj.setdefault(int(src[i]), ([], []))
        ids.append(int(dst[i]))
        wts.append(float(wt[i]))
    return adj


def start_rows(start_ids: Iterable[int], num_walks: int) -> List[Dict[str, Any]]:
    """initiate_random_walk (randomwalk.py:279-296)."""
    rows = []
    for v in start_ids:
        for i in range(1, num_walks + 1):
            rows.append({"src": -i, "dst": v, "path": [-i, v]})
    return rows


def step_row(
    row: Dict[str, Any],
    adj: Adjacency,
    return_param: float,
    inout_param: float,
    r1: float,
    r2: float,
    sum_mode: str = "naive",
) -> Dict[str, Any]:
    """One row of next_step_random_walk (randomwalk.py:316-339), with the adjacency
    already looked up (the two joins of fugue.py:147)."""
    prev, cur = row["src"], row["dst"]
    ids, wts = adj[cur]
    if prev < 0:
        alias, probs = alias_tables(wts, sum_mode)
    else:
        prev_out = set(adj[prev][0]) if prev in adj else set()
        alias, probs = edge_alias_tables(prev, prev_out, (ids, wts), return_param, inout_param, sum_mode)
    path = extend_path(row["path"], ids, alias, probs, r1, r2)
    return {"src": path[-2], "dst": path[-1], "path": path}


def random_walk(
    src: Sequence[int],
    dst: Sequence[int],
    wt: Sequence[float],
    n2v_params: Dict[str, Any],
    walk_seed: Optional[Iterable[int]] = None,
    random_seed: Optional[int] = None,
    sum_mode: str = "naive",
    rng: Any = _random,
) -> List[List[int]]:
    """The whole of fugue.random_walk (fugue.py:119-155) on one partition.

    * only vertices with an out-arc start walks (:132), optionally ∩ walk_seed (:133-134)
    * every step drops walkers whose current vertex has no out-arcs (inner join, :147)
    * with ``random_seed`` the module RNG is re-seeded at EVERY step (randomwalk.py:314-315)
    * rows are visited in start order: vertex id ascending, walk number 1..num_walks;
      two uniforms per row, r1 then r2 (randomwalk.py:336-337)
    Returns the walks (each ``walk_length + 1`` vertices) in row order.
    """
    adj = build_adjacency(src, dst, wt)
    starts = sorted(adj)
    if walk_seed is not None:
        # inner join on id: left (ascending id) order, one row per matching seed row -- a seed id
        # listed k times starts k x num_walks walkers (tests/test_fugue.py:73-75 does exactly that)
        mult: Dict[int, int] = {}
        for s in walk_seed:
            mult[int(s)] = mult.get(int(s), 0) + 1
        starts = [v for v in starts for _ in range(mult.get(v, 0))]
    rows = start_rows(starts, int(n2v_params["num_walks"]))
    p, q = n2v_params["return_param"], n2v_params["inout_param"]
    for _ in range(int(n2v_params["walk_length"])):
        if random_seed is not None:
            rng.seed(random_seed)
        nxt = []
        for row in rows:
            if row["dst"] not in adj:
                continue
            nxt.append(step_row(row, adj, p, q, rng.random(), rng.random(), sum_mode))
        rows = nxt
    return [list(r["path"]) for r in rows]


# ----------------------------------------------------------------------------------
# exact transition law (what the alias tables encode), for the chi-square gates
# ----------------------------------------------------------------------------------
def transition_law(
    adj: Adjacency, prev: Optional[int], cur: int, return_param: float, inout_param: float
) -> Dict[int, float]:
    """P(next = x | prev, cur) of the reference's sampler: proportional to the
    (biased) weights fed to generate_alias_tables; multi-arcs to the same x add up.
    ``prev=None`` is a first step (unbiased, randomwalk.py:320-321)."""
    ids, wts = adj[cur]
    if prev is None:
        bw = list(wts)
    else:
        prev_out = set(adj[prev][0]) if prev in adj else set()
        bw = biased_weights(prev, prev_out, ids, wts, return_param, inout_param)
    tot = math.fsum(bw)
    law: Dict[int, float] = {}
    for x, w in zip(ids, bw):
        law[x] = law.get(x, 0.0) + w / tot
    return law
