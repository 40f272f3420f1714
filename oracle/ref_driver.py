"""Drive the UNMODIFIED reference walk path on the host (TEST / BENCH INFRASTRUCTURE).

The reference's transformer functions (``node2vec/randomwalk.py``: ``get_vertex_neighbors``,
``initiate_random_walk``, ``next_step_random_walk``, ``to_path``) are imported from where the
reference lies -- ``/root/reference`` in the build container, ``baseline/_ref`` (the pip
``--target`` install made by ``scripts/install_reference.sh``; git-ignored, shipped to the GPU
box) elsewhere -- and chained exactly as ``node2vec/fugue.py:130-153`` chains them, with pandas
merges standing in for Fugue's two joins per step (``fugue.py:147``: left join on ``src``, inner
join on ``dst``; Fugue itself is not installable offline).  Nothing of the reference is copied.

Only ``tests/`` and ``bench.py`` (``cpu_baseline`` / ``--impl reference``) import this module.
"""
import importlib
import os
import sys
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ("/root/reference", os.path.join(ROOT, "baseline", "_ref"))


def load_reference() -> Tuple[Optional[object], Optional[str]]:
    """(the reference's ``node2vec.randomwalk`` module, the directory it came from), or (None, None)."""
    for root in CANDIDATES:
        if os.path.exists(os.path.join(root, "node2vec", "randomwalk.py")):
            if root not in sys.path:
                sys.path.insert(0, root)
            try:
                return importlib.import_module("node2vec.randomwalk"), root
            except Exception:  # noqa: BLE001  (a broken copy is the same as no copy)
                sys.path.remove(root)
    return None, None


class AdjacencyRows(object):
    """The reference's ``df_adj`` (``fugue.py:130``) built lazily: one ``get_vertex_neighbors`` call
    per vertex the sample touches (the full frame for a 16 M-edge graph would take minutes of
    pickling before the first step; the rows produced are the same)."""

    def __init__(self, ref_rw, row_ptr: np.ndarray, col: np.ndarray, weight: Optional[np.ndarray] = None):
        self.rw, self.row_ptr, self.col, self.weight = ref_rw, row_ptr, col, weight
        self.cache: Dict[int, str] = {}
        self.build_seconds = 0.0

    def rows(self, ids: Sequence[int]) -> pd.DataFrame:
        t0 = time.perf_counter()
        out_id, out_nb = [], []
        for v in ids:
            v = int(v)
            s = self.cache.get(v)
            if s is None:
                lo, hi = int(self.row_ptr[v]), int(self.row_ptr[v + 1])
                if lo == hi:
                    continue                       # no out-arcs: no row in df_adj (the inner join drops the walker)
                w = np.ones(hi - lo) if self.weight is None else self.weight[lo:hi]
                part = pd.DataFrame({"src": v, "dst": self.col[lo:hi], "weight": w})   # sorted by dst (presort)
                s = next(iter(self.rw.get_vertex_neighbors(part)))["neighbors"]
                self.cache[v] = s
            out_id.append(v)
            out_nb.append(s)
        self.build_seconds += time.perf_counter() - t0
        return pd.DataFrame({"id": np.asarray(out_id, dtype=np.int64), "neighbors": pd.Series(out_nb, dtype=object)})


def walk(ref_rw, adj: AdjacencyRows, starts: Sequence[int], num_walks: int, walk_length: int, p: float, q: float,
         seed: Optional[int] = None, hot_budget: Optional[float] = None) -> Tuple[List[List[int]], int, float]:
    """``fugue.random_walk`` for the given start vertices.  Returns (walks, walker-steps done,
    seconds spent in the step loop = the two joins + ``next_step_random_walk``; building the
    adjacency rows is outside, like the GPU side's graph build).  ``hot_budget`` stops the loop
    after the step that exceeds that many step-loop seconds (the walks are then incomplete: timing
    samples only)."""
    start_rows = adj.rows(starts)
    walks = pd.DataFrame([dict(r) for r in ref_rw.initiate_random_walk(
        ({"id": int(v)} for v in start_rows["id"].tolist()), num_walks)], columns=["dst", "src", "path"])
    steps, hot = 0, 0.0
    for _ in range(walk_length):
        if len(walks) == 0:
            break
        need = np.union1d(walks["dst"].to_numpy(dtype=np.int64),
                          walks["src"].to_numpy(dtype=np.int64)[walks["src"].to_numpy(dtype=np.int64) >= 0])
        rows = adj.rows(need.tolist())                                     # untimed (one-off in the reference)
        t0 = time.perf_counter()
        df_src = rows.rename(columns={"id": "src", "neighbors": "src_neighbors"})
        df_dst = rows.rename(columns={"id": "dst", "neighbors": "dst_neighbors"})
        nxt = walks.merge(df_src, on="src", how="left").merge(df_dst, on="dst", how="inner").drop(columns=["dst"])
        recs = nxt.to_dict("records")
        for r in recs:                                                     # first step: no src row -> None (:318)
            if not isinstance(r["src_neighbors"], str):
                r["src_neighbors"] = None
        out = [dict(r) for r in ref_rw.next_step_random_walk(recs, p, q, seed)]
        walks = pd.DataFrame(out, columns=["src", "dst", "path"])
        hot += time.perf_counter() - t0
        steps += len(out)
        if hot_budget is not None and hot > hot_budget:
            break
    paths = [r["walk"] for r in ref_rw.to_path(walks.to_dict("records"))] if len(walks) else []
    return paths, steps, hot


def timed_sample(ref_rw, row_ptr, col, weight, starts, num_walks, walk_length, p, q, budget_s: float,
                 batch: int = 64) -> Tuple[int, float, int]:
    """Walk batches of start vertices until ``budget_s`` seconds of step-loop time are spent.
    Returns (walker-steps, step-loop seconds, start vertices done)."""
    adj = AdjacencyRows(ref_rw, row_ptr, col, weight)
    steps, hot, done = 0, 0.0, 0
    wall0 = time.perf_counter()
    for lo in range(0, len(starts), batch):
        _, s, t = walk(ref_rw, adj, starts[lo:lo + batch], num_walks, walk_length, p, q, hot_budget=budget_s - hot)
        steps, hot, done = steps + s, hot + t, done + len(starts[lo:lo + batch])
        if hot > budget_s or time.perf_counter() - wall0 > 4 * budget_s:
            break
    return steps, hot, done
