"""Synthetic graphs of the shapes BASELINE.json names (there is no network for datasets).

All generators are seeded and return SYMMETRISED, de-duplicated, loop-free arc lists
(src, dst) so that no walker dies; weights are implicit 1.0.
"""
import os
from typing import Tuple

import numpy as np


def _symmetrise(a: np.ndarray, b: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    src = np.concatenate([a, b]).astype(np.int32)
    dst = np.concatenate([b, a]).astype(np.int32)
    return src, dst


def erdos_renyi(n: int = 10000, m: int = 100000, seed: int = 42) -> Tuple[np.ndarray, np.ndarray]:
    """BASELINE configs[0]: G(n, m) uniform -- exactly m distinct undirected edges."""
    rng = np.random.default_rng(seed)
    keys = np.zeros(0, dtype=np.int64)
    while len(keys) < m:
        a = rng.integers(0, n, 2 * (m - len(keys)) + 16)
        b = rng.integers(0, n, len(a))
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        k = (lo.astype(np.int64) << 32 | hi)[lo != hi]
        keys = np.unique(np.concatenate([keys, k]))
    keys = rng.permutation(keys)[:m]
    return _symmetrise((keys >> 32).astype(np.int64), (keys & 0xFFFFFFFF).astype(np.int64))


def blogcatalog_like(n: int = 10000, m: int = 334000, seed: int = 42) -> Tuple[np.ndarray, np.ndarray]:
    """BASELINE configs[1]: a 10k-vertex power-law graph with clustering, thinned to m
    undirected edges (Holme-Kim growth, 34 links per new vertex, triad probability 0.3;
    max degree ~1.7k).  Cached under /tmp because the generator takes a few seconds."""
    cache = f"/tmp/n2v_blogcatalog_like_{n}_{m}_{seed}.npz"
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            return z["src"], z["dst"]
        except (OSError, ValueError, KeyError):
            pass
    import networkx as nx
    g = nx.powerlaw_cluster_graph(n, 34, 0.3, seed=seed)
    e = np.array(g.edges(), dtype=np.int64)
    rng = np.random.default_rng(seed)
    e = e[rng.permutation(len(e))]
    if len(e) > m:
        deg = np.bincount(e.reshape(-1), minlength=n)
        keep = np.ones(len(e), dtype=bool)
        drop = len(e) - m
        for i in range(len(e)):
            if drop == 0:
                break
            a, b = e[i]
            if deg[a] > 1 and deg[b] > 1:
                keep[i] = False
                deg[a] -= 1
                deg[b] -= 1
                drop -= 1
        e = e[keep]
    src, dst = _symmetrise(e[:, 0], e[:, 1])
    try:   # atomic publish: several ranks may generate the same graph at the same time
        tmp = f"{cache}.{os.getpid()}.tmp.npz"
        np.savez(tmp, src=src, dst=dst)
        os.replace(tmp, cache)
    except OSError:
        pass
    return src, dst


def rmat_device(scale: int, edge_factor: int = 16, seed: int = 42, device=None,
                abcd=(0.57, 0.19, 0.19, 0.05), hotspots: int = 0, hotspot_degree: int = 0):
    """Graph500-style R-MAT generated ON THE DEVICE (torch ops; this is input synthesis, not
    the hot path): 2^scale vertices, edge_factor * 2^scale undirected edges before
    de-duplication; optional injected hotspot vertices.  Returns (src, dst) int32 CUDA
    tensors, symmetrised, unique, loop-free."""
    import torch
    device = torch.device(device if device is not None else "cuda")
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_edges = edge_factor << scale
    a, b, c, _ = abcd
    src = torch.zeros(n_edges, dtype=torch.int64, device=device)
    dst = torch.zeros(n_edges, dtype=torch.int64, device=device)
    for _ in range(scale):
        r = torch.rand(n_edges, device=device, generator=gen)
        src = (src << 1) | (r >= a + b).long()
        dst = (dst << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c)).long()
    # scramble ids so hubs are not clustered at small ids
    perm = torch.randperm(1 << scale, device=device, generator=gen)
    src, dst = perm[src], perm[dst]
    if hotspots > 0:
        n = 1 << scale
        hubs = torch.randint(0, n, (hotspots,), device=device, generator=gen)
        hs = hubs.repeat_interleave(hotspot_degree)
        hd = torch.randint(0, n, (hotspots * hotspot_degree,), device=device, generator=gen)
        src, dst = torch.cat([src, hs]), torch.cat([dst, hd])
    lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
    keys = torch.unique(((lo << 32) | hi)[lo != hi])
    lo, hi = keys >> 32, keys & 0xFFFFFFFF
    del keys
    return torch.cat([lo, hi]).to(torch.int32), torch.cat([hi, lo]).to(torch.int32)


def rmat_hotspot_edges_device(scale: int = 20, edge_factor: int = 16, seed: int = 42, device=None,
                              hotspots: int = 16, hotspot_degree: int = 1 << 18):
    """BASELINE configs[2] input: the UNDIRECTED edge list (every edge once, src < dst, unique,
    loop-free) of an R-MAT graph with injected hotspot vertices -- the shape a reference user
    hands to ``trim_index(..., directed=False, max_out_deg=...)``, which trims the listed
    direction first and symmetrises afterwards (fugue.py:57-77, indexer.py:45-48)."""
    import torch
    src, dst = rmat_device(scale, edge_factor, seed, device, hotspots=hotspots, hotspot_degree=hotspot_degree)
    half = src.numel() // 2          # rmat_device returns [lo..., hi...] / [hi..., lo...]
    return src[:half].contiguous(), dst[:half].contiguous()


RMAT_BLOCKS = 64


def _rmat_keys(scale, n_edges, seed, device, abcd=(0.57, 0.19, 0.19, 0.05), chunk=1 << 26):
    """n_edges R-MAT edges as int64 keys (src << 32 | dst), ids scrambled by an affine bijection."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    a, b, c, _ = abcd
    n = 1 << scale
    mult = 0x9E3779B1 | 1                               # odd => bijection mod 2^scale
    out = torch.empty(n_edges, dtype=torch.int64, device=device)
    for lo in range(0, n_edges, chunk):
        m = min(chunk, n_edges - lo)
        src = torch.zeros(m, dtype=torch.int64, device=device)
        dst = torch.zeros(m, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(m, device=device, generator=gen)
            src = (src << 1) | (r >= a + b).long()
            dst = (dst << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c)).long()
        src = (src * mult + 12345) & (n - 1)
        dst = (dst * mult + 12345) & (n - 1)
        out[lo:lo + m] = (src << 32) | dst
        del src, dst, r
    return out


def rmat_partition_device(scale: int, edge_factor: int, rank: int, world: int, device, seed: int = 1000,
                          group=None):
    """BASELINE configs[4] input, generated where it will live: every rank draws 1/world of the
    R-MAT edges, both directions are routed to the owner of their source vertex (vertex ranges of
    ceil(V / world)) with one all-to-all, duplicates and self loops are dropped locally.
    Returns this rank's arcs (src, dst) as int32 CUDA tensors sorted by (src, dst)."""
    import torch
    import torch.distributed as dist
    V = 1 << scale
    S = (V + world - 1) // world
    n_edges = edge_factor << scale
    # the edge multiset is drawn in RMAT_BLOCKS fixed blocks (block b under seed + b), dealt round-robin
    # to the ranks: the GRAPH is the same for every world size (strong scaling on one fixed graph)
    blocks = [b for b in range(RMAT_BLOCKS) if b % world == rank]
    sizes = [n_edges * (b + 1) // RMAT_BLOCKS - n_edges * b // RMAT_BLOCKS for b in blocks]
    keys = torch.empty(sum(sizes), dtype=torch.int64, device=device)
    at = 0
    for b, m in zip(blocks, sizes):
        keys[at:at + m] = _rmat_keys(scale, m, seed + b, device)
        at += m
    keys = keys[(keys >> 32) != (keys & 0xFFFFFFFF)]                        # no self loops
    keys = torch.cat([keys, ((keys & 0xFFFFFFFF) << 32) | (keys >> 32)])    # both directions
    if world > 1:
        owner = torch.div(keys >> 32, S, rounding_mode="floor")
        send_counts = torch.bincount(owner, minlength=world)
        order = torch.argsort(owner)
        del owner
        keys = keys[order]
        del order
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        recv = torch.empty(int(recv_counts.sum()), dtype=torch.int64, device=device)
        dist.all_to_all_single(recv, keys, output_split_sizes=recv_counts.tolist(),
                               input_split_sizes=send_counts.tolist(), group=group)
        keys = recv
        del recv
    keys = torch.unique(keys)                                               # simple graph; sorted by (src, dst)
    src = (keys >> 32).to(torch.int32)
    dst = (keys & 0xFFFFFFFF).to(torch.int32)
    return src, dst


def products_like_device(n: int = 2449029, n_edges: int = 61859140, seed: int = 42, device=None):
    """BASELINE configs[3]: an ogbn-products-shaped graph (2.4 M vertices, ~62 M undirected edges):
    R-MAT scale 22 folded onto n vertices.  Returns symmetrised int32 CUDA tensors."""
    import torch
    device = torch.device(device if device is not None else "cuda")
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    a, b, c = 0.57, 0.19, 0.19
    m = int(n_edges * 1.08)                      # head-room for the duplicates removed below
    src = torch.zeros(m, dtype=torch.int64, device=device)
    dst = torch.zeros(m, dtype=torch.int64, device=device)
    for _ in range(22):
        r = torch.rand(m, device=device, generator=gen)
        src = (src << 1) | (r >= a + b).long()
        dst = (dst << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c)).long()
    perm = torch.randperm(1 << 22, device=device, generator=gen)
    src, dst = perm[src] % n, perm[dst] % n
    lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
    keys = torch.unique(((lo << 32) | hi)[lo != hi])
    if keys.numel() > n_edges:
        keys = keys[torch.randperm(keys.numel(), device=device, generator=gen)[:n_edges]]
    lo, hi = keys >> 32, keys & 0xFFFFFFFF
    return torch.cat([lo, hi]).to(torch.int32), torch.cat([hi, lo]).to(torch.int32)


def rmat_host(scale: int, edge_factor: int = 16, seed: int = 42, abcd=(0.57, 0.19, 0.19, 0.05)):
    """numpy twin of rmat_device for boxes without a GPU (CPU baseline legs)."""
    rng = np.random.default_rng(seed)
    n_edges = edge_factor << scale
    a, b, c, _ = abcd
    src = np.zeros(n_edges, dtype=np.int64)
    dst = np.zeros(n_edges, dtype=np.int64)
    for _ in range(scale):
        r = rng.random(n_edges)
        src = (src << 1) | (r >= a + b)
        dst = (dst << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c))
    perm = rng.permutation(1 << scale)
    src, dst = perm[src], perm[dst]
    lo, hi = np.minimum(src, dst), np.maximum(src, dst)
    keys = np.unique(((lo << 32) | hi)[lo != hi])
    lo, hi = keys >> 32, keys & 0xFFFFFFFF
    return np.concatenate([lo, hi]).astype(np.int32), np.concatenate([hi, lo]).astype(np.int32)
