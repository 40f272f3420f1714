"""Device-side graph preparation for integer-id arc lists too large for pandas: the work of
``trim_index`` (reference fugue.py:24-77) that precedes the hot path -- hotspot trimming
(randomwalk.py:238-262) and the undirected expansion of the indexer (indexer.py:45-48).

Data preparation: the reference's partition step is one stable device sort by source vertex; the
two order-sensitive pandas steps of its indexer -- first-occurrence vertex ids and the
first-occurrence de-duplication of the undirected expansion -- run in K6 ``n2v_first_occurrence``
(a sort-free position hash table, ``csrc/index.cu``); trimming is K5 ``n2v_trim_sample``.  Not part
of the measured hot path.  With a ``seed`` the kept arcs and their
order are bit-identical to the reference's pandas path (``DataFrame.sample(n, random_state=seed)``
= numpy's legacy ``RandomState(seed).permutation(deg)[:n]``, re-seeded per vertex; the kernel runs
MT19937 and the Fisher-Yates shuffle on the device).  Without a seed the reference draws from the
unseeded global generator; here a torch generator draws the same law (uniform without
replacement).
"""
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from .constants import MAX_OUT_DEGREES


def trim_hotspots_device(src: torch.Tensor, dst: torch.Tensor, weight: Optional[torch.Tensor] = None,
                         max_out_deg: int = 0, seed: Optional[int] = None
                         ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """Keep at most ``max_out_deg`` (<= 0 means 100000) uniformly chosen out-arcs per vertex."""
    cap = max_out_deg if max_out_deg > 0 else MAX_OUT_DEGREES
    s = src.long()
    n = int(s.max()) + 1 if s.numel() else 0
    deg = torch.bincount(s, minlength=n)
    if s.numel() == 0 or int(deg.max()) <= cap:
        return src, dst, weight
    if seed is not None:
        return _trim_exact(src, dst, weight, s, deg, cap, int(seed))
    gen = torch.Generator(device=src.device)
    gen.seed()
    over = deg[s] > cap                                   # arcs of oversize vertices
    idx = torch.nonzero(over).view(-1)
    key = (s[idx] << 31) | torch.randint(0, 2 ** 31 - 1, (idx.numel(),), device=src.device, generator=gen)
    order = torch.argsort(key)                            # grouped by vertex, random order inside
    s_sorted = s[idx][order]
    first = torch.ones_like(s_sorted, dtype=torch.bool)
    first[1:] = s_sorted[1:] != s_sorted[:-1]
    seg_start = torch.cummax(torch.where(first, torch.arange(s_sorted.numel(), device=src.device), 0), 0).values
    rank = torch.arange(s_sorted.numel(), device=src.device) - seg_start
    keep = torch.ones(src.numel(), dtype=torch.bool, device=src.device)
    keep[idx[order[rank >= cap]]] = False
    return src[keep], dst[keep], (None if weight is None else weight[keep])


def _trim_exact(src, dst, weight, s, deg, cap: int, seed: int):
    """The reference's output, row for row: partitions in ascending ``src`` order (fugue.py:57-67),
    untouched partitions in input order, an oversize partition replaced by its rows at positions
    ``RandomState(seed).permutation(deg)[:cap]`` in that order (randomwalk.py:256-260)."""
    if not 0 <= seed <= 0xFFFFFFFF:
        raise ValueError("Seed must be between 0 and 2**32 - 1")          # numpy's own message
    dev = src.device
    n = int(deg.numel())
    order = torch.argsort(s, stable=True)                                 # ascending src, input order inside
    group_start = torch.cumsum(deg, 0) - deg
    hot = torch.nonzero(deg > cap).view(-1)
    n_hot = int(hot.numel())
    hot_deg = deg[hot].contiguous()
    scratch_off = (torch.cumsum(hot_deg, 0) - hot_deg).contiguous()
    scratch = torch.empty(int(hot_deg.sum()), dtype=torch.int32, device=dev)
    picked = torch.empty((n_hot, cap), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().n2v_trim_sample(_lib.ptr(hot_deg), _lib.ptr(scratch_off), n_hot, cap, C.c_uint32(seed),
                                               _lib.ptr(scratch), _lib.ptr(picked), _lib.current_stream_ptr()),
                   "n2v_trim_sample")
    new_deg = deg.clamp(max=cap)
    out_start = torch.cumsum(new_deg, 0) - new_deg
    v = torch.repeat_interleave(torch.arange(n, device=dev), new_deg)     # vertex of every output row
    r = torch.arange(int(v.numel()), device=dev) - out_start[v]           # its rank inside the vertex
    hot_index = torch.full((n,), -1, dtype=torch.int64, device=dev)
    hot_index[hot] = torch.arange(n_hot, device=dev)
    hi = hot_index[v]
    pos = torch.where(hi >= 0, picked[hi.clamp(min=0), r.clamp(max=cap - 1)].to(torch.int64), r)
    sel = order[group_start[v] + pos]
    return src[sel], dst[sel], (None if weight is None else weight[sel])


def symmetrise_device(src: torch.Tensor, dst: torch.Tensor, weight: Optional[torch.Tensor] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """indexer.py:45-48 on integer ids: append the reversed arcs, keep the first occurrence of every
    (src, dst, weight) triple in frame order (K6, no sort)."""
    s = torch.cat([src, dst]).to(torch.int64)
    d = torch.cat([dst, src]).to(torch.int64)
    n = int(s.numel())
    pos = torch.arange(n, device=s.device)
    if weight is None:
        keep = first_occurrence(s, d) == pos
        return s[keep].to(src.dtype), d[keep].to(dst.dtype), None
    w = torch.cat([weight, weight])
    keep = first_occurrence(s, d, _weight_key(w)) == pos
    return s[keep].to(src.dtype), d[keep].to(dst.dtype), w[keep]


# ======================================================================================
# K6: the reference's graph indexer on the device (indexer.py:9-49)
# ======================================================================================
def first_occurrence(*cols: torch.Tensor) -> torch.Tensor:
    """first[i] = the smallest j whose key (1-3 int64 columns) equals row i's.  No sort."""
    lib = _lib.load()
    if not 1 <= len(cols) <= 3:
        raise ValueError("1 to 3 key columns")
    cols = [c.contiguous() if c.dtype == torch.int64 else c.to(torch.int64).contiguous() for c in cols]
    n, dev = int(cols[0].numel()), cols[0].device
    out = torch.empty(n, dtype=torch.int64, device=dev)
    if n == 0:
        return out
    n_slots = int(lib.n2v_first_occurrence_slots(n))
    with torch.cuda.device(dev):
        table = torch.empty(n_slots, dtype=torch.int32, device=dev)
        ptrs = [_lib.ptr(c) for c in cols] + [None] * (3 - len(cols))
        _lib.check(lib.n2v_first_occurrence(ptrs[0], ptrs[1], ptrs[2], n, _lib.ptr(table), n_slots, _lib.ptr(out),
                                            _lib.current_stream_ptr()), "n2v_first_occurrence")
    return out


def first_occurrence_bytes(offsets: torch.Tensor, data: torch.Tensor) -> torch.Tensor:
    """``first_occurrence`` for byte-string keys in Arrow layout (int64 ``offsets[n + 1]`` into the
    uint8 ``data``): first[i] = the smallest j whose bytes equal row i's."""
    lib = _lib.load()
    n, dev = int(offsets.numel()) - 1, offsets.device
    out = torch.empty(max(n, 0), dtype=torch.int64, device=dev)
    if n <= 0:
        return out
    n_slots = int(lib.n2v_first_occurrence_slots(n))
    with torch.cuda.device(dev):
        table = torch.empty(n_slots, dtype=torch.int32, device=dev)
        _lib.check(lib.n2v_first_occurrence_bytes(_lib.ptr(offsets.contiguous()), _lib.ptr(data.contiguous()), n,
                                                  _lib.ptr(table), n_slots, _lib.ptr(out), _lib.current_stream_ptr()),
                   "n2v_first_occurrence_bytes")
    return out


def string_name_ranks(src_names, dst_names, device):
    """Order- and equality-preserving int64 codes for STRING vertex names without factorising the 2E
    names on the host: the names go to the device as one Arrow ``large_string`` buffer pair, K6 finds
    every name's first occurrence byte for byte, only the V DISTINCT names come back to be sorted
    (Arrow's binary order = Python's ``str`` order), and the ranks are broadcast to the rows on the
    device.  Returns (src_code, dst_code, sorted distinct names as an object array) or None when a
    name is not a string / is missing (the caller falls back)."""
    import numpy as np
    import pyarrow as pa
    import pyarrow.compute as pc
    try:
        arr = pa.concat_arrays([pa.array(src_names, type=pa.large_string()), pa.array(dst_names, type=pa.large_string())])
    except (pa.ArrowInvalid, pa.ArrowTypeError, TypeError):
        return None
    if arr.null_count:
        return None
    n, e = len(arr), len(src_names)
    bufs = arr.buffers()
    offsets = np.frombuffer(bufs[1], dtype=np.int64)[arr.offset: arr.offset + n + 1]
    data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None and bufs[2].size else np.zeros(1, dtype=np.uint8)
    t_off = torch.as_tensor(offsets.copy(), device=device)
    t_dat = torch.as_tensor(data.copy(), device=device)
    first = first_occurrence_bytes(t_off, t_dat)
    is_first = first == torch.arange(n, device=device)
    pos = torch.nonzero(is_first).view(-1)
    distinct = arr.take(pa.array(pos.cpu().numpy()))                  # V names, first-occurrence order
    order = pc.sort_indices(distinct).to_numpy()                       # binary (UTF-8) order = str order
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(len(order), dtype=np.int64)
    dense = torch.cumsum(is_first.to(torch.int64), 0) - 1              # index among the distinct names
    code = torch.as_tensor(rank, device=device)[dense[first]]
    uniques = distinct.take(pa.array(order)).to_numpy(zero_copy_only=False)
    return code[:e].contiguous(), code[e:].contiguous(), uniques


def _weight_key(w: torch.Tensor) -> torch.Tensor:
    """IEEE bit pattern with pandas' equality: -0.0 == 0.0 and every NaN equals every NaN."""
    w = w.double()
    w = torch.where(w == 0, torch.zeros_like(w), w)
    w = torch.where(torch.isnan(w), torch.full_like(w, float("nan")), w)
    return w.contiguous().view(torch.int64)


def index_graph_device(src_name: torch.Tensor, dst_name: torch.Tensor, weight: Optional[torch.Tensor] = None,
                       directed: bool = True, dense_ids: bool = False):
    """``index_graph_pandas`` (indexer.py:9-49) for integer vertex names, on the device, row for row:
    vertex id = position of the name's first occurrence in [all src..., all dst...] (:26-35; sparse,
    up to 2E - 1), arcs keep their order, weight defaults to 1.0 (fp64), and ``directed=False``
    appends the reversed arcs and keeps the first occurrence of every (src, dst, weight) triple
    (:45-48).  ``dense_ids=True`` numbers the vertices 0..V-1 in first-occurrence order instead
    (compact tables for the walk / SGNS kernels; the name table maps back).
    Returns (src_id, dst_id, weight, vertex_id, vertex_name), all device tensors."""
    dev = src_name.device
    e = int(src_name.numel())
    names = torch.cat([src_name.to(torch.int64), dst_name.to(torch.int64)])
    first = first_occurrence(names)
    is_first = first == torch.arange(2 * e, device=dev)
    vertex_pos = torch.nonzero(is_first).view(-1)            # ascending positions = first-occurrence order
    vertex_name = names[vertex_pos]
    if dense_ids:
        rank = torch.cumsum(is_first.to(torch.int64), 0) - 1
        ids = rank[first]
        vertex_id = torch.arange(int(vertex_pos.numel()), device=dev)
    else:
        ids, vertex_id = first, vertex_pos
    s, d = ids[:e], ids[e:]
    w = torch.ones(e, dtype=torch.float64, device=dev) if weight is None else weight.to(torch.float64)
    if directed is not True:
        s2, d2, w2 = torch.cat([s, d]), torch.cat([d, s]), torch.cat([w, w])
        keep = first_occurrence(s2, d2, _weight_key(w2)) == torch.arange(2 * e, device=dev)
        s, d, w = s2[keep], d2[keep], w2[keep]               # boolean selection preserves frame order
    return s, d, w, vertex_id, vertex_name


def partition_by_src(src: torch.Tensor, dst: torch.Tensor, weight: Optional[torch.Tensor] = None):
    """``partition(by=["src"])`` (fugue.py:57-60): rows grouped by source in ascending key order,
    input order inside a group -- one stable device sort."""
    order = torch.sort(src, stable=True).indices
    return src[order], dst[order], (None if weight is None else weight[order])


def trim_partitioned(src: torch.Tensor, dst: torch.Tensor, weight: Optional[torch.Tensor] = None,
                     max_out_deg: int = 0, seed: Optional[int] = None):
    """``partition(by=["src"]).transform(trim_hotspot_vertices)`` (fugue.py:57-67) for integer source
    names of any range: rows come out grouped by ascending source, input order inside a group, an
    oversize group replaced by its sample (seeded: the rows ``DataFrame.sample`` keeps, in its order)."""
    src, dst, weight = partition_by_src(src, dst, weight)
    if src.numel() == 0:
        return src, dst, weight
    head = torch.ones(src.numel(), dtype=torch.bool, device=src.device)
    head[1:] = src[1:] != src[:-1]
    group = torch.cumsum(head.to(torch.int64), 0) - 1          # dense rank of the source among the sorted sources
    cap = max_out_deg if max_out_deg > 0 else MAX_OUT_DEGREES
    deg = torch.bincount(group)
    if int(deg.max()) <= cap:
        return src, dst, weight
    if seed is not None:
        return _trim_exact(src, dst, weight, group, deg, cap, int(seed))
    keep = torch.ones(src.numel(), dtype=torch.bool, device=src.device)
    gen = torch.Generator(device=src.device)
    gen.seed()
    tag = torch.rand(src.numel(), device=src.device, generator=gen)
    order = torch.sort(group.double() + tag * 0.999999, stable=True).indices      # random order inside each group
    start = torch.cumsum(deg, 0) - deg
    rank = torch.arange(src.numel(), device=src.device) - start[group[order]]
    keep[order[rank >= cap]] = False
    return src[keep], dst[keep], (None if weight is None else weight[keep])
