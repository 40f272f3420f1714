"""Drop-in for the reference's ``node2vec/fugue.py``: same two entry points, same
arguments, same errors -- the body is the B200 engine instead of a Fugue DAG.

``compute_engine`` is accepted for signature compatibility and ignored (Fugue is not a
dependency; a Spark engine is rejected because its data lives in a JVM).  Frames are
duck-typed: a pandas DataFrame, anything exposing ``.as_pandas()`` (Fugue
``ArrayDataFrame`` / ``PandasDataFrame``), or -- for graphs too big for pandas -- a tuple
of torch tensors / numpy arrays ``(src, dst[, weight])``.
"""
import logging
import os
from typing import Any, Dict, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from .constants import MAX_OUT_DEGREES, NODE2VEC_PARAMS
from .graph import DeviceGraph, walk_to_host
from .indexer import index_graph_pandas


class Frame(object):
    """Minimal stand-in for a Fugue DataFrame around a pandas frame."""

    def __init__(self, df: pd.DataFrame):
        self._df = df

    @property
    def native(self) -> pd.DataFrame:
        return self._df

    def as_pandas(self) -> pd.DataFrame:
        return self._df

    def as_array(self):
        return self._df.values.tolist()

    def count(self) -> int:
        return len(self._df)

    @property
    def schema(self):
        return list(self._df.columns)

    def __len__(self):
        return len(self._df)


class WalkFrame(Frame):
    """Result of ``random_walk``: the ``[src, walk]`` frame of the reference
    (fugue.py:109-117), materialised lazily from the walk matrix.

    ``walks``       int32 [W_alive, L+1] numpy matrix (host)
    ``walks_device`` the same rows as a torch tensor still in HBM (for SGNS)
    """

    def __init__(self, walks_device: torch.Tensor, walks_host: Optional[np.ndarray] = None, stats=None):
        self.walks_device = walks_device
        self._host = walks_host
        self.stats = stats
        self._df = None

    @property
    def walks(self) -> np.ndarray:
        if self._host is None:
            self._host = self.walks_device.cpu().numpy()
        return self._host

    def as_pandas(self) -> pd.DataFrame:
        if self._df is None:
            w = self.walks
            self._df = pd.DataFrame({"src": w[:, 0].astype(np.int64) if len(w) else np.zeros(0, dtype=np.int64),
                                     "walk": w.tolist()})
        return self._df

    native = property(as_pandas)

    def as_array(self):
        return self.as_pandas().values.tolist()

    def to_parquet(self, path: str) -> None:
        """Write the [src, walk] frame as parquet (the reference's examples persist walks this way,
        examples/fugue_spark.py:49-50) without building Python lists: `walk` is an Arrow
        fixed-size list column backed by the walk matrix."""
        import pyarrow as pa
        import pyarrow.parquet as pq
        w = np.ascontiguousarray(self.walks)
        n, length = w.shape if w.ndim == 2 else (0, 1)
        walk = pa.FixedSizeListArray.from_arrays(pa.array(w.reshape(-1), type=pa.int32()), length)
        src = pa.array(w[:, 0].astype(np.int64) if n else np.zeros(0, dtype=np.int64))
        pq.write_table(pa.table({"src": src, "walk": walk}), path)

    def count(self) -> int:
        return int(self.walks_device.shape[0])

    @property
    def schema(self):
        return ["src", "walk"]

    def __len__(self):
        return self.count()


def read_walks_parquet(path: str) -> pd.DataFrame:
    """Read a [src, walk] parquet file (written by WalkFrame.to_parquet or by the reference's
    pipeline) back into the pandas frame Node2VecGensim accepts."""
    import pyarrow.parquet as pq
    t = pq.read_table(path)
    return pd.DataFrame({"src": t.column("src").to_numpy(), "walk": t.column("walk").to_pylist()})


def _columns(df) -> list:
    if isinstance(df, pd.DataFrame):
        return list(df.columns)
    schema = getattr(df, "schema", None)
    if schema is None:
        raise ValueError(f"Unsupported frame type {type(df)}")
    names = getattr(schema, "names", None)
    return list(names) if names is not None else list(schema)


def _to_pandas(df) -> pd.DataFrame:
    if isinstance(df, pd.DataFrame):
        return df
    if hasattr(df, "as_pandas"):
        return df.as_pandas()
    raise ValueError(f"Unsupported frame type {type(df)}")


def _reject_spark(engine) -> None:
    if engine is not None and "spark" in type(engine).__name__.lower():
        raise NotImplementedError("SparkExecutionEngine is out of scope: pass a pandas / Fugue-native frame")


def trim_index(
    compute_engine: Any,
    df_graph: Any,
    indexed: bool = False,
    directed: bool = True,
    max_out_deg: int = 0,
    random_seed: Optional[int] = None,
    *,
    dense_ids: bool = False,
) -> Tuple[Frame, Optional[Frame]]:
    """Validate, trim hotspot vertices, index.  Same contract as the reference
    (fugue.py:24-77): ``max_out_deg <= 0`` means 100000 (randomwalk.py:254-255);
    ``indexed=True`` returns the trimmed frame untouched and ``None``.
    Frames whose ``src`` / ``dst`` are integer names are trimmed, indexed and (``directed=False``)
    mirrored ON THE DEVICE when a GPU is present -- K5 / K6, one H2D of the columns in, the frames the
    reference returns out, row for row; string names go to the device as raw UTF-8 bytes (one Arrow buffer
    pair), K6 matches them byte for byte and only the V distinct names come back to be sorted on the host
    (the reference's partition order), the row-level work is the same device path.
    Tuples of device tensors ``(src, dst[, weight])`` (integer names) never leave the GPU and return
    ``((src, dst, weight), (vertex_id, vertex_name))`` tensors.  Keyword-only ``dense_ids=True`` numbers
    vertices 0..V-1 in first-occurrence order instead of the reference's sparse positions."""
    logging.info("trim_index(): start validating, trimming, and indexing ...")
    if isinstance(df_graph, (tuple, list)) and len(df_graph) in (2, 3) and isinstance(df_graph[0], torch.Tensor):
        # integer vertex names / ids as device tensors (graphs too large for pandas): everything stays on the GPU.
        # The reference's order: partition by src + trim (fugue.py:57-67), then index (first-occurrence ids,
        # indexer.py:26-43), then expand to both directions (indexer.py:45-48).
        from .preprocess import index_graph_device, symmetrise_device, trim_partitioned
        src, dst = df_graph[0], df_graph[1]
        w = df_graph[2] if len(df_graph) == 3 else None
        src, dst, w = trim_partitioned(src, dst, w, max_out_deg, random_seed)
        if indexed is True:
            # frames return here (fugue.py:70-71); tensors are always indexed, so `directed` is honoured for
            # them: trimmed first, mirrored afterwards, a trimmed hub keeps its mirrored in-arcs and the graph
            # stays symmetric (INTEGRATION.md notes the difference)
            if directed is not True:
                src, dst, w = symmetrise_device(src, dst, w)
            return ((src, dst) if w is None else (src, dst, w)), None
        s, d, wt, vid, vname = index_graph_device(src, dst, w, directed, dense_ids=dense_ids)
        return (s, d, wt), (vid, vname)
    cols = _columns(df_graph)
    if "src" not in cols or "dst" not in cols:
        raise ValueError(f"Input graph NOT in the right format: {cols}")
    _reject_spark(compute_engine)
    df = _to_pandas(df_graph)
    if indexed is not True and torch.cuda.is_available() and set(df.columns) <= {"src", "dst", "weight"}:
        res = _index_graph_frame_on_device(df, directed, max_out_deg, random_seed)
        if res is not None:
            return res
    cap = max_out_deg if max_out_deg > 0 else MAX_OUT_DEGREES
    # host path (no GPU, or columns / name types the device path does not take):
    # partition(by=["src"]) + trim_hotspot_vertices: groups come out in key order, rows in
    # input order, an oversize group is replaced by DataFrame.sample(n=cap, random_state=seed)
    df = df.sort_values("src", kind="stable")
    sizes = df.groupby("src", sort=False)["src"].transform("size")
    if (sizes > cap).any():
        parts = []
        for _, part in df.groupby("src", sort=True):
            if len(part) > cap:
                part = part.sample(n=cap, random_state=random_seed) if random_seed is not None \
                    else part.sample(n=cap)
            parts.append(part)
        df = pd.concat(parts)
    df = df.reset_index(drop=True)
    if indexed is True:
        return Frame(df), None
    df_res, name_id = index_graph_pandas(df, directed)
    return Frame(df_res.reset_index(drop=True)), Frame(name_id)


def _frame_name_keys(df: pd.DataFrame, device=None):
    """int64 keys for the ``src`` / ``dst`` names of a frame: integer names as they are; string names by
    their rank among the sorted distinct names (equality- and order-preserving) -- found ON THE DEVICE
    from the raw UTF-8 bytes (``preprocess.string_name_ranks``: K6 on an Arrow buffer, only the V distinct
    names are sorted on the host); any other hashable names by a host ``factorize``.
    Returns (src_key, dst_key, uniques or None, name dtype), or None when the names cannot be ranked."""
    src, dst = df["src"].to_numpy(), df["dst"].to_numpy()
    if src.dtype.kind in "iu" and dst.dtype.kind in "iu" and src.dtype != np.uint64 and dst.dtype != np.uint64:
        return src.astype(np.int64), dst.astype(np.int64), None, (src.dtype if src.dtype == dst.dtype else object)
    if device is not None:
        from .preprocess import string_name_ranks
        got = string_name_ranks(df["src"], df["dst"], device)
        if got is not None:
            return got[0], got[1], got[2], object
    try:
        codes, uniques = pd.factorize(np.concatenate([src.astype(object), dst.astype(object)]), sort=True)
    except TypeError:
        return None
    if (codes < 0).any():
        return None                                   # missing names: leave them to pandas' semantics
    return codes[: len(src)].astype(np.int64), codes[len(src):].astype(np.int64), uniques, object


def _index_graph_frame_on_device(df: pd.DataFrame, directed, max_out_deg=None, random_seed=None):
    """``index_graph_pandas`` (and, with ``max_out_deg`` given, the partition + trim step before it)
    for a pandas frame with the row-level work on the GPU (K5 trimming, K6 first-occurrence ids and
    de-duplication): one H2D of the columns, the reference's two frames back.  String names are replaced by
    order-preserving ranks first (``_frame_name_keys``: found on the device from the raw bytes) and mapped
    back in the name table.
    Returns None when the names cannot be ranked (mixed types): the caller falls back to pandas."""
    from .preprocess import index_graph_device, trim_partitioned
    dev = torch.device("cuda", torch.cuda.current_device())
    keys = _frame_name_keys(df, dev)
    if keys is None:
        return None
    s_key, d_key, uniques, name_dtype = keys
    ts, td = torch.as_tensor(s_key, device=dev), torch.as_tensor(d_key, device=dev)
    tw = None
    if "weight" in df.columns:
        tw = torch.as_tensor(df["weight"].to_numpy().astype(np.float64), device=dev)
    else:
        df["weight"] = 1.0                                # the reference adds the column to the caller's frame too
    if max_out_deg is not None:
        ts, td, tw = trim_partitioned(ts, td, tw, max_out_deg, random_seed)
    s, d, w, vid, vname = index_graph_device(ts, td, tw, directed)
    vname = vname.cpu().numpy()
    names = uniques[vname] if uniques is not None else (vname.astype(name_dtype) if name_dtype is not object
                                                        else vname.astype(object))
    df_edge = pd.DataFrame({"src": s.cpu().numpy(), "dst": d.cpu().numpy(), "weight": w.cpu().numpy()})
    name_id = pd.DataFrame({"vertex_id": vid.cpu().numpy(), "vertex_name": names})
    logging.info(f"Num of indexed vertices: {len(name_id)}")
    logging.info(f"Num of indexed edges: {len(df_edge)}")
    return Frame(df_edge), Frame(name_id)


def _graph_arrays(df_graph):
    if isinstance(df_graph, (tuple, list)) and len(df_graph) in (2, 3) and not isinstance(df_graph[0], (int, float)):
        src, dst = df_graph[0], df_graph[1]
        weight = df_graph[2] if len(df_graph) == 3 else None
        return src, dst, weight
    cols = _columns(df_graph)
    if "src" not in cols or "dst" not in cols:
        raise ValueError(f"Input graph NOT in the right format: {cols}")
    df = _to_pandas(df_graph)
    weight = df["weight"].to_numpy(dtype=np.float64) if "weight" in df.columns else None
    return df["src"].to_numpy(), df["dst"].to_numpy(), weight


def _walk_rows(graph, start, num_walks, walk_length, p, q, random_seed, collect_stats, out):
    """One walk over ``start``; rows of dropped walkers removed.  Returns (device rows, host rows or
    None, stats).  With ``out`` (pinned host matrix) the copy is pipelined with the kernel."""
    if out is None or collect_stats:
        walks, alive, stats = graph.walk(start, num_walks, walk_length, p, q, random_seed, collect_stats)
        walks = walks[alive] if not bool(alive.all()) else walks
        if out is None:
            return walks, None, stats
        out[: walks.shape[0]].copy_(walks)
        return walks, out.numpy()[: walks.shape[0]], stats
    walks, alive = walk_to_host(graph, start, num_walks, walk_length, p, q, random_seed, out)
    host = out.numpy()[: walks.shape[0]]
    if not bool(alive.all()):
        keep = alive.cpu().numpy()
        n = int(keep.sum())
        host[:n] = host[keep]                     # compact in place (boolean indexing copies first)
        walks, host = walks[alive], host[:n]
    return walks, host, None


_SEED_LAYER_MIX = 0x9E3779B97F4A7C15   # odd 64-bit constant: layer k walks under seed + k * MIX (mod 2^64)


def _walk_from_seed_ids(graph, start, walk_seed, num_walks, walk_length, p, q, random_seed, collect_stats, out=None):
    """``walk_start.inner_join(walk_seed)`` (fugue.py:133-134): rows stay in ascending-id order and an
    id listed k times in ``walk_seed`` starts k x num_walks independent walkers (the reference's own
    test feeds duplicated ids, tests/test_fugue.py:73-75).  A walker's Philox stream is keyed by
    (seed, start vertex, walk number), so the k-th copy of an id runs as layer k under a derived seed;
    layers are merged back into the join's row order on the device."""
    ids, mult = np.unique(_to_pandas(walk_seed)["id"].to_numpy().astype(np.int64), return_counts=True)
    dev = start.device
    ids_t = torch.as_tensor(ids, device=dev)
    mult_t = torch.as_tensor(mult, device=dev)
    start64 = start.to(torch.int64)
    pos = torch.searchsorted(ids_t, start64).clamp_(max=max(int(ids_t.numel()) - 1, 0))
    hit = (ids_t[pos] == start64) if ids_t.numel() else torch.zeros_like(start64, dtype=torch.bool)
    start, m = start[hit], mult_t[pos[hit]]
    layers = int(m.max().item()) if m.numel() else 1
    if random_seed is None:
        random_seed = int.from_bytes(os.urandom(8), "little")
    if layers <= 1:
        return _walk_rows(graph, start, num_walks, walk_length, p, q, random_seed, collect_stats, out)
    rank = torch.arange(start.numel(), device=dev, dtype=torch.int64)
    r = torch.arange(num_walks, device=dev, dtype=torch.int64)
    parts, keys, stats = [], [], None
    for k in range(layers):
        sel = m > k
        w, alive, st = graph.walk(start[sel], num_walks, walk_length, p, q,
                                  (random_seed + k * _SEED_LAYER_MIX) & 0xFFFFFFFFFFFFFFFF, collect_stats)
        key = ((rank[sel] * layers + k) * num_walks).view(-1, 1) + r.view(1, -1)
        alive = alive.view(-1).bool()
        parts.append(w[alive])
        keys.append(key.view(-1)[alive])
        if st is not None:
            stats = st if stats is None else {n: stats[n] + st[n] for n in st}
    order = torch.argsort(torch.cat(keys))
    walks = torch.cat(parts)[order]
    if out is None:
        return walks, None, stats
    out[: walks.shape[0]].copy_(walks)
    return walks, out.numpy()[: walks.shape[0]], stats


def random_walk(
    compute_engine: Any,
    df_graph: Any,
    n2v_params: Dict[str, Any],
    walk_seed: Any = None,
    random_seed: Optional[int] = None,
    checkpoint_dir: Optional[str] = "/tmp",
    *,
    graph: Optional[DeviceGraph] = None,
    collect_stats: bool = False,
    out: Optional[torch.Tensor] = None,
    process_group: Any = None,
    n_vertices: Optional[int] = None,
    assume_symmetric: bool = False,
) -> WalkFrame:
    """Second-order biased random walks; same contract as the reference (fugue.py:81-155).

    * defaults are merged into ``n2v_params`` in place (:120-122)
    * ``walk_seed`` must have an ``id`` column, else ValueError (:123-124)
    * only vertices with an out-arc start walks (:132), ``num_walks`` each, optionally
      restricted to ``walk_seed`` ids (:133-134)
    * a walker that reaches a vertex with no out-arcs before its last step is dropped (:147)
    * returns ``[src, walk]`` with ``walk_length + 1`` vertices per walk (:153)
    ``random_seed`` keys the Philox stream (None = fresh entropy).  ``checkpoint_dir`` is
    accepted and unused: the walk matrix lives in HBM, there is no lineage to truncate.
    Keyword-only extras: ``graph`` reuses a prebuilt DeviceGraph; ``collect_stats``; ``out`` is a
    pinned host int32 matrix ``[>= walkers, walk_length + 1]`` to receive the rows -- the walk then
    runs in chunks of start vertices whose device->host copies overlap the next chunk's kernel
    (same rows as one launch), and ``WalkFrame.walks`` is a view of ``out``.  ``process_group``
    (torch.distributed, one process per GPU) selects the VERTEX-PARTITIONED graph: ``df_graph``
    then holds only the arcs whose source lies in this rank's vertex range
    ``[rank * ceil(n_vertices / G), ...)`` (global ids), every rank builds its part, peers read each
    other's parts over NVLink, and the call returns this rank's rows (walks from its own start
    vertices) -- the union over ranks is what one GPU holding the whole graph would return.
    """
    logging.info("random_walk(): start random walking ...")
    for param in NODE2VEC_PARAMS:
        if param not in n2v_params:
            n2v_params[param] = NODE2VEC_PARAMS[param]
    if walk_seed is not None and "id" not in _columns(walk_seed):
        raise ValueError(f"walk_seed has no column of 'id': {_columns(walk_seed)}!")
    _reject_spark(compute_engine)
    p, q = n2v_params["return_param"], n2v_params["inout_param"]
    if p == 0 or q == 0:
        raise ValueError(f"Zero return ({p}) or inout ({q}) parameter!")
    num_walks, walk_length = int(n2v_params["num_walks"]), int(n2v_params["walk_length"])
    if num_walks < 1 or walk_length < 1:
        raise ValueError("num_walks and walk_length must be >= 1")

    own_graph = None
    if graph is None:
        src, dst, weight = _graph_arrays(df_graph)
        if process_group is not None:
            from .graph import PartitionedGraph
            if n_vertices is None:
                raise ValueError("a vertex-partitioned walk needs n_vertices (the global vertex count)")
            graph = own_graph = PartitionedGraph.from_local_arcs(src, dst, weight, int(n_vertices), group=process_group,
                                                                 assume_symmetric=assume_symmetric, keep_weight=False)
        else:
            graph = DeviceGraph.from_arcs(src, dst, weight, n_vertices=n_vertices)
    start = graph.start_vertices()
    host = None
    if walk_seed is None:
        walks, host, stats = _walk_rows(graph, start, num_walks, walk_length, p, q, random_seed, collect_stats, out)
    else:
        walks, host, stats = _walk_from_seed_ids(graph, start, walk_seed, num_walks, walk_length, p, q, random_seed,
                                                 collect_stats, out)
    if own_graph is not None:
        own_graph.close()       # unmap the peers' parts, free the shareable buffers (collective)
    logging.info("random_walk(): random walking done ...")
    return WalkFrame(walks, walks_host=host, stats=stats)
