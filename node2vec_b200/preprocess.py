"""Device-side graph preparation for integer-id arc lists too large for pandas: the work of
``trim_index`` (reference fugue.py:24-77) that precedes the hot path -- hotspot trimming
(randomwalk.py:238-262) and the undirected expansion of the indexer (indexer.py:45-48).

Data-preparation built from torch primitives (sort / unique / bincount on the GPU); not
part of the measured hot path.  The pandas path in ``fugue.trim_index`` stays the bit-exact
twin of the reference (same numpy RandomState sampling); here the sample is drawn with a
seeded torch generator -- the same law (uniform without replacement, ``max_out_deg`` arcs per
oversize vertex), not the same bits.
"""
from typing import Optional, Tuple

import torch

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
    gen = torch.Generator(device=src.device)
    if seed is not None:
        gen.manual_seed(int(seed))
    else:
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
