"""Device-side graph preparation for integer-id arc lists too large for pandas: the work of
``trim_index`` (reference fugue.py:24-77) that precedes the hot path -- hotspot trimming
(randomwalk.py:238-262) and the undirected expansion of the indexer (indexer.py:45-48).

Data-preparation built from torch primitives (sort / unique / bincount on the GPU) plus K5
``n2v_trim_sample``; not part of the measured hot path.  With a ``seed`` the kept arcs and their
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
    """indexer.py:45-48: append the reversed arcs, drop exact duplicate (src, dst, weight) triples
    (first occurrence kept; output ordered by (src, dst), which the CSR build would do anyway)."""
    s = torch.cat([src, dst]).long()
    d = torch.cat([dst, src]).long()
    if weight is None:
        key = torch.unique((s << 32) | d)
        return (key >> 32).to(src.dtype), (key & 0xFFFFFFFF).to(dst.dtype), None
    w = torch.cat([weight, weight])
    wbits = w.double().view(torch.int64)
    order = torch.argsort(wbits, stable=True)
    order = order[torch.argsort(((s << 32) | d)[order], stable=True)]
    k, wb = ((s << 32) | d)[order], wbits[order]
    first = torch.ones_like(k, dtype=torch.bool)
    first[1:] = (k[1:] != k[:-1]) | (wb[1:] != wb[:-1])
    sel = order[first]
    return s[sel].to(src.dtype), d[sel].to(dst.dtype), w[sel]
