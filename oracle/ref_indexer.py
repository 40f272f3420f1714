"""CPU restatement of the reference's pandas graph indexer and hotspot trimming
(TEST INFRASTRUCTURE).  Follows ``node2vec/indexer.py:9-49`` and
``node2vec/randomwalk.py:238-262``; pinned by ``tests/golden/indexer.json`` /
``trim.json``, which were produced by running the unmodified reference functions
(``tests/golden/make_golden.py``; the indexer under a ``pyspark`` import stub and a
``DataFrame.append`` shim because pandas >= 2 removed ``append``).
"""
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

MAX_OUT_DEGREES = 100000  # constants.py:6


def index_graph(
    src: Sequence[Any],
    dst: Sequence[Any],
    weight: Optional[Sequence[float]],
    directed: bool,
) -> Tuple[np.ndarray, np.ndarray, np.ndarray, List[Any], np.ndarray]:
    """indexer.py:20-49 without DataFrames.

    vertex_id of a name = position of its FIRST occurrence in the concatenation
    ``[src..., dst...]`` (append(ignore_index) -> drop_duplicates -> reset_index keeps
    the pre-dedup row label, :26-35), so ids are sparse and can reach 2E-1.
    Arcs keep input order (:38-43).  Undirected: append the reversed arcs, then drop
    exact duplicate (src, dst, weight) triples keeping the first (:45-48).

    Returns (src_id, dst_id, weight_f64, vertex_names, vertex_ids).
    """
    n = len(src)
    wt = np.ones(n, dtype=np.float64) if weight is None else np.asarray(weight, dtype=np.float64)
    first: Dict[Any, int] = {}
    for pos, name in enumerate(list(src) + list(dst)):
        if name not in first:
            first[name] = pos
    names = list(first.keys())
    ids = np.fromiter(first.values(), dtype=np.int64, count=len(first))
    s = np.fromiter((first[x] for x in src), dtype=np.int64, count=n)
    d = np.fromiter((first[x] for x in dst), dtype=np.int64, count=n)
    if not directed:
        s2 = np.concatenate([s, d])
        d2 = np.concatenate([d, s])
        w2 = np.concatenate([wt, wt])
        seen = set()
        keep = []
        for i, key in enumerate(zip(s2.tolist(), d2.tolist(), w2.tolist())):
            if key not in seen:
                seen.add(key)
                keep.append(i)
        keep = np.asarray(keep, dtype=np.int64)
        s, d, wt = s2[keep], d2[keep], w2[keep]
    return s, d, wt, names, ids


def trim_hotspot(
    src: Sequence[int],
    dst: Sequence[int],
    weight: Sequence[float],
    max_out_degree: int = 0,
    random_seed: Optional[int] = None,
) -> pd.DataFrame:
    """trim_hotspot_vertices applied per ``src`` partition (fugue.py:57-67,
    randomwalk.py:238-262).  ``max_out_degree <= 0`` means 100000, not "off"
    (:254-255).  A vertex with more arcs keeps ``pandas.DataFrame.sample(n=max)`` of
    them, i.e. ``RandomState(seed).permutation(len)[:max]`` rows in that order.
    Partitions are emitted in ascending ``src`` order; rows inside keep input order
    unless sampled."""
    if max_out_degree <= 0:
        max_out_degree = MAX_OUT_DEGREES
    df = pd.DataFrame({"src": list(src), "dst": list(dst), "weight": list(weight)})
    parts = []
    for _, part in df.groupby("src", sort=True):
        if len(part) > max_out_degree:
            rs = np.random.RandomState(random_seed) if random_seed is not None else np.random
            take = rs.permutation(len(part))[:max_out_degree]
            part = part.iloc[take]
        parts.append(part)
    if not parts:
        return df
    return pd.concat(parts, ignore_index=True)
