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


# ----------------------------------------------------------------------------------
# numpy's legacy RandomState(seed).permutation(n), restated (what K5 n2v_trim_sample runs)
# ----------------------------------------------------------------------------------
class _MT19937(object):
    """MT19937 as numpy seeds it for an int seed (mt19937_seed == init_genrand) and draws 32-bit
    words from it (mt19937_next)."""

    def __init__(self, seed: int):
        if not 0 <= seed <= 0xFFFFFFFF:
            raise ValueError("Seed must be between 0 and 2**32 - 1")
        self.key = [0] * 624
        for pos in range(624):
            self.key[pos] = seed
            seed = (1812433253 * (seed ^ (seed >> 30)) + pos + 1) & 0xFFFFFFFF
        self.pos = 624

    def _refill(self):
        k = self.key
        for i in range(624):
            y = (k[i] & 0x80000000) | (k[(i + 1) % 624] & 0x7FFFFFFF)
            k[i] = k[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.pos = 0

    def next32(self) -> int:
        if self.pos == 624:
            self._refill()
        y = self.key[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def numpy_legacy_permutation(n: int, seed: int) -> List[int]:
    """``np.random.RandomState(seed).permutation(n)``: shuffle(arange(n)) = Fisher-Yates from the top
    index down, the partner of index i drawn by masked rejection (``random_interval(i)``: smallest
    all-ones mask >= i, draw 32-bit words until ``word & mask <= i``).  pandas' ``DataFrame.sample(
    n=k, random_state=seed)`` -- the reference's hotspot trimming, randomwalk.py:256-260 -- keeps rows
    ``permutation(len)[:k]`` in that order."""
    rng = _MT19937(seed)
    arr = list(range(n))
    for i in range(n - 1, 0, -1):
        mask = i
        for sh in (1, 2, 4, 8, 16):
            mask |= mask >> sh
        while True:
            j = rng.next32() & mask
            if j <= i:
                break
        arr[i], arr[j] = arr[j], arr[i]
    return arr
