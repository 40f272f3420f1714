"""Host mirror of the reference's ``node2vec/randomwalk.py`` interface.

Same names, argument meaning and error behaviour as the reference module, so its tests
read the same against this one.  The value types (Neighbors / AliasProb / RandomPath)
are wire formats and stay on the host; every piece of ARITHMETIC (alias tables, p/q
biasing, alias draws) runs in the CUDA library through the C ABI -- there is no CPU
implementation of it here.  The production walk does not go through these row-level
functions at all (see ``node2vec_b200.fugue.random_walk`` / ``graph.DeviceGraph.walk``);
they exist for drop-in compatibility at the transformer level and for parity tests.
"""
import base64
import ctypes as C
import pickle
import random
from typing import Any, Dict, Iterable, List, Optional, Sequence, Set, Tuple, Union

import numpy as np
import pandas as pd
import torch

from . import _lib
from .constants import MAX_OUT_DEGREES

# the reference's golden strings are pickle protocol 3 (Python 3.6/3.7 default)
_PICKLE_PROTOCOL = 3
SUM_MODE = "naive"  # see DESIGN.md "sum modes"; "neumaier" = the reference under CPython >= 3.12


class ZeroWeightError(ValueError, ZeroDivisionError):
    """Empty or all-zero weight vector.  The reference fails with ZeroDivisionError
    (randomwalk.py:172-173); callers that validate inputs expect ValueError.  Both work."""


def _encode(obj) -> str:
    return base64.b64encode(pickle.dumps(obj, protocol=_PICKLE_PROTOCOL)).decode()


def _decode(text: str):
    return pickle.loads(base64.b64decode(text.encode()))


class Neighbors(object):
    """Adjacency of one vertex: (neighbour ids, weights).  Mirrors randomwalk.py:17-41."""

    def __init__(self, obj: Union[str, pd.DataFrame, Tuple[List[int], List[float]]]):
        if isinstance(obj, str):
            self._data = _decode(obj)
        elif isinstance(obj, pd.DataFrame):
            self._data = (obj["dst"].tolist(), obj["weight"].tolist())
        else:
            self._data = obj

    @property
    def dst_id(self):
        return self._data[0]

    @property
    def dst_wt(self):
        return self._data[1]

    def items(self):
        return zip(self._data[0], self._data[1])

    def serialize(self) -> str:
        return _encode(self._data)

    def as_pandas(self) -> pd.DataFrame:
        return pd.DataFrame({"dst": self._data[0], "weight": self._data[1]})


class AliasProb(object):
    """An alias table (alias, probs).  Mirrors randomwalk.py:44-99; the draws run on the
    device (n2v_alias_draw)."""

    def __init__(self, obj: Union[str, pd.DataFrame, Tuple[List[int], List[float]]]):
        if isinstance(obj, str):
            self._data = _decode(obj)
        elif isinstance(obj, pd.DataFrame):
            self._data = (obj["alias"].tolist(), obj["probs"].tolist())
        else:
            self._data = obj

    @property
    def alias(self):
        return self._data[0]

    @property
    def probs(self):
        return self._data[1]

    def serialize(self) -> str:
        return _encode(self._data)

    def sampling_from_alias_wiki(self, first_random: float) -> int:
        return int(alias_draw([self.alias], [self.probs], [first_random], None)[0])

    def sampling_from_alias(self, first_random: float, second_random: float) -> int:
        return int(alias_draw([self.alias], [self.probs], [first_random], [second_random])[0])


class RandomPath(object):
    """A walk in progress.  Mirrors randomwalk.py:102-153."""

    def __init__(self, obj: Union[str, List[int]]):
        self._data = _decode(obj) if isinstance(obj, str) else obj

    @property
    def path(self):
        return self._data

    @property
    def last_edge(self):
        return self._data[-2], self._data[-1]

    def serialize(self) -> str:
        return _encode(self._data)

    def __str__(self):
        return self._data.__repr__()

    def append(self, dst_neighbors: List[int], alias_prob: AliasProb, first_random: float,
               second_random: Optional[float] = None) -> "RandomPath":
        if second_random is not None:
            k = alias_prob.sampling_from_alias(first_random, second_random)
        else:
            k = alias_prob.sampling_from_alias_wiki(first_random)
        nxt = dst_neighbors[k]
        path = list(self._data)
        if len(path) == 2 and path[0] < 0:     # a fresh walker [-i, v] becomes [v, x]
            return RandomPath([path[1], nxt])
        path.append(nxt)
        return RandomPath(path)


# --------------------------------------------------------------------------------------
# device-backed batch primitives
# --------------------------------------------------------------------------------------
def _dev():
    _lib.require_cuda()
    return torch.device(f"cuda:{torch.cuda.current_device()}")


def alias_draw(alias: Sequence[Sequence[int]], probs: Sequence[Sequence[float]], first: Sequence[float],
               second: Optional[Sequence[float]]) -> np.ndarray:
    """Draw i picks from table i.  second=None selects the one-uniform sampler."""
    dev = _dev()
    lib = _lib.load()
    sizes = np.fromiter((len(a) for a in alias), dtype=np.int64, count=len(alias))
    offset = np.zeros(len(alias) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offset[1:])
    a = torch.as_tensor(np.concatenate([np.asarray(x, dtype=np.int32) for x in alias]), device=dev)
    p = torch.as_tensor(np.concatenate([np.asarray(x, dtype=np.float64) for x in probs]), device=dev)
    off = torch.as_tensor(offset, device=dev)
    r1 = torch.as_tensor(np.asarray(first, dtype=np.float64), device=dev)
    r2 = None if second is None else torch.as_tensor(np.asarray(second, dtype=np.float64), device=dev)
    out = torch.empty(len(alias), dtype=torch.int32, device=dev)
    _lib.check(lib.n2v_alias_draw(_lib.ptr(a), _lib.ptr(p), _lib.ptr(off), _lib.ptr(r1), _lib.ptr(r2),
                                  len(alias), _lib.ptr(out), _lib.current_stream_ptr()), "n2v_alias_draw")
    return out.cpu().numpy()


def _mini_graph(prev_ids, prev_sets, cur_ids, cur_wts):
    """Pack B independent rows (cur adjacency in the caller's order, prev id, prev out-set)
    as one CSR.  Row i gets its own block of vertex ids: every id the row mentions is
    replaced by block_start + rank (order preserving, so prev's out-set stays ascending
    and equality / membership are unchanged); the row's current vertex is one extra
    vertex at the end of the block.  Returns host arrays + per-row (prev, cur, offset, n)."""
    vtx_base, vtx_deg, col, wt = [], [], [], []
    prev_v, cur_v, offs, sizes = [], [], [], []
    pos = 0
    for pid, ps, ids, w in zip(prev_ids, prev_sets, cur_ids, cur_wts):
        ids = [int(x) for x in ids]
        pset = sorted(int(x) for x in ps) if ps else []
        first = pid is None or pid < 0
        universe = sorted(set(ids) | set(pset) | (set() if first else {int(pid)}))
        rank = {x: k for k, x in enumerate(universe)}
        block = len(vtx_base)
        m = len(universe)
        base = [0] * (m + 1)
        deg = [0] * (m + 1)
        if not first:
            base[rank[int(pid)]] = pos
            deg[rank[int(pid)]] = len(pset)
            col.extend(block + rank[x] for x in pset)
            wt.extend([1.0] * len(pset))
            pos += len(pset)
        base[m] = pos
        deg[m] = len(ids)
        col.extend(block + rank[x] for x in ids)
        wt.extend(float(x) for x in w)
        offs.append(pos)
        sizes.append(len(ids))
        pos += len(ids)
        vtx_base.extend(base)
        vtx_deg.extend(deg)
        prev_v.append(-1 if first else block + rank[int(pid)])
        cur_v.append(block + m)
    vtx = np.zeros((len(vtx_base), 4), dtype=np.int32)
    vtx[:, 0] = np.asarray(vtx_base, dtype=np.uint32).view(np.int32)
    vtx[:, 1] = np.asarray(vtx_deg, dtype=np.uint32).view(np.int32)
    return (vtx, np.asarray(col, dtype=np.int32), np.asarray(wt, dtype=np.float64),
            np.asarray(prev_v, dtype=np.int32), np.asarray(cur_v, dtype=np.int32),
            np.asarray(offs, dtype=np.int64), sizes)


def edge_alias_tables_batch(prev_ids: Sequence[int], prev_sets: Sequence[Optional[Iterable[int]]],
                            cur_ids: Sequence[Sequence[int]], cur_wts: Sequence[Sequence[float]],
                            return_param: float, inout_param: float,
                            sum_mode: Optional[str] = None) -> List[Tuple[List[int], List[float]]]:
    """generate_edge_alias_tables for a batch of rows on the device (n2v_edge_alias_build).
    prev_ids[i] < 0 means an unbiased first-step table (generate_alias_tables)."""
    dev = _dev()
    lib = _lib.load()
    B = len(cur_ids)
    for ids, w in zip(cur_ids, cur_wts):
        if len(ids) == 0 or not (np.asarray(w, dtype=np.float64).sum() != 0):
            raise ZeroWeightError("float division by zero")
    vtx, col, wt, prev_v, cur_v, offs, sizes = _mini_graph(prev_ids, prev_sets, cur_ids, cur_wts)
    t_vtx = torch.as_tensor(vtx, device=dev)
    t_col = torch.as_tensor(col, device=dev)
    t_wt = torch.as_tensor(wt, device=dev)
    g = _lib.Graph()
    g.n_vertices, g.n_arcs, g.flags, g.n_parts, g.part_size = len(vtx), len(col), 0, 1, max(len(vtx), 1)
    g.parts[0].vtx, g.parts[0].col, g.parts[0].weight = t_vtx.data_ptr(), t_col.data_ptr(), t_wt.data_ptr()
    t_prev = torch.as_tensor(prev_v, device=dev)
    t_cur = torch.as_tensor(cur_v, device=dev)
    t_off = torch.as_tensor(offs, device=dev)
    alias_out = torch.zeros(len(col), dtype=torch.int32, device=dev)
    probs_out = torch.zeros(len(col), dtype=torch.float64, device=dev)
    scratch = torch.empty(len(col), dtype=torch.int32, device=dev)
    _lib.check(lib.n2v_edge_alias_build(C.byref(g), _lib.ptr(t_prev), _lib.ptr(t_cur), B, float(return_param),
                                        float(inout_param), _lib.SUM_MODE[sum_mode or SUM_MODE], _lib.ptr(t_off),
                                        _lib.ptr(alias_out), _lib.ptr(probs_out), _lib.ptr(scratch),
                                        _lib.current_stream_ptr()), "n2v_edge_alias_build")
    a, p = alias_out.cpu().numpy(), probs_out.cpu().numpy()
    return [(a[o:o + n].tolist(), p[o:o + n].tolist()) for o, n in zip(offs.tolist(), sizes)]


# --------------------------------------------------------------------------------------
# the reference's function names
# --------------------------------------------------------------------------------------
def generate_alias_tables(node_weights: List[float]) -> Tuple[List[int], List[float]]:
    """Alias-method tables of one weight vector (reference randomwalk.py:157-190),
    bit-exact, computed by the CUDA library."""
    n = len(node_weights)
    return edge_alias_tables_batch([-1], [None], [list(range(n))], [node_weights], 1.0, 1.0)[0]


def generate_edge_alias_tables(
    src_id: int,
    src_nbs_id: Set[int],
    dst_neighbors: Tuple[List[int], List[float]],
    return_param: float = 1.0,
    inout_param: float = 1.0,
) -> Tuple[List[int], List[float]]:
    """p/q-biased alias tables of one (prev, cur) pair (reference randomwalk.py:193-232)."""
    if len(dst_neighbors) != 2 or len(dst_neighbors[0]) != len(dst_neighbors[1]):
        raise ValueError(f"Invalid neighbors tuple '{dst_neighbors}'!")
    if return_param == 0 or inout_param == 0:
        raise ValueError(f"Zero return ({return_param}) or inout ({inout_param}) parameter!")
    if src_id < 0:
        raise ValueError("src_id must be a vertex id (>= 0)")
    return edge_alias_tables_batch([src_id], [src_nbs_id], [dst_neighbors[0]], [dst_neighbors[1]],
                                   return_param, inout_param)[0]


def trim_hotspot_vertices(df: pd.DataFrame, max_out_degree: int = 0,
                          random_seed: Optional[int] = None) -> Iterable[Dict[str, Any]]:
    """Cap one vertex's out-arcs by uniform sampling without replacement
    (reference randomwalk.py:238-262).  ``max_out_degree <= 0`` means 100000."""
    if max_out_degree <= 0:
        max_out_degree = MAX_OUT_DEGREES
    if len(df) > max_out_degree:
        df = df.sample(n=max_out_degree, random_state=random_seed) if random_seed is not None \
            else df.sample(n=max_out_degree)
    for _, row in df.iterrows():
        yield dict(row)


def get_vertex_neighbors(df: pd.DataFrame) -> Iterable[Dict[str, Any]]:
    """One vertex's partition -> {"id", "neighbors"} (reference randomwalk.py:266-275)."""
    yield {"id": df.loc[0, "src"], "neighbors": Neighbors(df).serialize()}


def initiate_random_walk(df: Iterable[Dict[str, Any]], num_walks: int) -> Iterable[Dict[str, Any]]:
    """num_walks fresh walkers per start vertex (reference randomwalk.py:279-296).  As in
    the reference, the SAME dict object is yielded repeatedly per vertex."""
    for arow in df:
        src = arow["id"]
        row = {"dst": src}
        for i in range(1, num_walks + 1):
            row.update({"src": -i, "path": [-i, src]})
            yield row


def next_step_random_walk(df: Iterable[Dict[str, Any]], return_param: float, inout_param: float,
                          random_seed: Optional[int] = None) -> Iterable[Dict[str, Any]]:
    """Row-level compatibility form of one walk step (reference randomwalk.py:300-339):
    consumes joined rows {src, path, src_neighbors, dst_neighbors}, draws (r1, r2) per row
    from Python's ``random`` exactly as the reference does, and extends each path.  Table
    construction and the draws run batched on the device."""
    if random_seed is not None:
        random.seed(random_seed)
    rows = list(df)
    if not rows:
        return
    prev_ids, prev_sets, cur_ids, cur_wts, r1, r2 = [], [], [], [], [], []
    for row in rows:
        src, src_nbs = row["src"], row["src_neighbors"]
        prev_sets.append(set() if src_nbs is None else set(Neighbors(src_nbs).dst_id))
        nbs = Neighbors(row["dst_neighbors"])
        cur_ids.append(nbs.dst_id)
        cur_wts.append(nbs.dst_wt)
        prev_ids.append(src)
        r1.append(random.random())
        r2.append(random.random())
    tables = edge_alias_tables_batch(prev_ids, prev_sets, cur_ids, cur_wts, return_param, inout_param)
    picks = alias_draw([t[0] for t in tables], [t[1] for t in tables], r1, r2)
    for row, ids, k in zip(rows, cur_ids, picks):
        path = list(row["path"])
        nxt = ids[int(k)]
        path = [path[1], nxt] if (len(path) == 2 and path[0] < 0) else path + [nxt]
        yield {"src": path[-2], "dst": path[-1], "path": path}


def to_path(df: Iterable[Dict[str, Any]]) -> Iterable[Dict[str, Any]]:
    """{path} -> {src, walk} (reference randomwalk.py:343-349)."""
    for row in df:
        path = RandomPath(row["path"]).path
        yield {"src": path[0], "walk": path}
