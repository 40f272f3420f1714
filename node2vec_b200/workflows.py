"""Workflows around the hot path: resumable walks, incremental re-walks and the
(p, q) link-prediction sweep the reference's docstring asks its users to do by hand.

What they replace in the reference:
  * ``checkpoint_dir`` / ``persist()`` inside the step loop (fugue.py:105,149) -- the reference's
    only resilience feature: lineage truncation so that a lost Spark stage does not recompute all
    steps.  Here a whole walk is one kernel launch, so the unit of resumption is a RANGE OF START
    VERTICES: ``random_walk_resumable`` writes one shard per range plus a manifest and skips the
    shards that are already on disk.
  * ``walk_seed`` (fugue.py:96-100,132-134) -- "restrict walks to these vertices", whose purpose is
    incremental refresh.  ``rewalk`` finds the walks that a graph change invalidated and re-walks
    only their start vertices.
  * "node2vec needs hyper-parameter search on p and q" (fugue.py:90-95) -- ``pq_sweep``.

Because a walker's Philox stream is keyed by (seed, start vertex, walk number) and nothing else
(include/n2v_b200.h, n2v_walk), both shortcuts are EXACT: a resumed/chunked run and an incremental
re-walk give the same rows, bit for bit, as one monolithic walk of the final graph under the same
seed (as long as the graph keeps its sampler class, DeviceGraph.flags).  The tests assert that.
"""
import hashlib
import json
import os
import time
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .constants import NODE2VEC_PARAMS
from .fugue import WalkFrame
from .graph import DeviceGraph

MANIFEST = "manifest.json"


def _merged(n2v_params: Dict[str, Any]) -> Dict[str, Any]:
    out = dict(NODE2VEC_PARAMS)
    out.update(n2v_params)
    if out["return_param"] == 0 or out["inout_param"] == 0:
        raise ValueError(f"Zero return ({out['return_param']}) or inout ({out['inout_param']}) parameter!")
    if int(out["num_walks"]) < 1 or int(out["walk_length"]) < 1:
        raise ValueError("num_walks and walk_length must be >= 1")
    return out


def _walk_alive(graph: DeviceGraph, start: torch.Tensor, prm: Dict[str, Any], seed: int) -> torch.Tensor:
    walks, alive, _ = graph.walk(start, int(prm["num_walks"]), int(prm["walk_length"]),
                                 prm["return_param"], prm["inout_param"], seed)
    return walks if bool(alive.all()) else walks[alive]


# ------------------------------------------------------------------------------------------------
# resumable, chunked walks
# ------------------------------------------------------------------------------------------------
def graph_fingerprint(graph: DeviceGraph) -> str:
    """Cheap identity of a built graph for the manifest: sizes, flags and a checksum of the degree
    sequence, neighbour ids and weight bits (device reductions; not cryptographic)."""
    def mix(t):                                   # position-weighted wrap-around int64 checksum
        t = t.reshape(-1).to(torch.int64)
        pos = torch.arange(t.numel(), device=t.device, dtype=torch.int64) % 65521 + 1
        return int((t * pos).sum().item())
    m = hashlib.sha256()
    sums = [graph.n_vertices, graph.n_arcs, graph.flags, mix(graph.vtx[:, 1]), mix(graph.col),
            mix(graph.weight.view(torch.int64))]
    m.update(np.asarray(sums, dtype=np.int64).tobytes())
    return m.hexdigest()[:32]


def random_walk_resumable(graph: DeviceGraph, n2v_params: Dict[str, Any], out_dir: str, random_seed: int,
                          chunk_starts: int = 1 << 20, walk_seed: Optional[Iterable[int]] = None,
                          max_chunks: Optional[int] = None) -> Dict[str, Any]:
    """Walk every start vertex in chunks of ``chunk_starts`` start vertices, spilling each chunk's
    rows to ``out_dir/walks_<k>.npy`` (int32 [rows, L+1]) and recording it in ``manifest.json``.
    A second call with the same arguments skips the chunks already recorded (and present), so an
    interrupted job resumes where it stopped.  ``max_chunks`` stops after that many NEW chunks (used
    by the tests to simulate an interruption).  A manifest written for another graph, seed or
    parameter set is refused with ValueError rather than silently mixed.
    Returns the manifest (``complete`` tells whether every chunk is done)."""
    prm = _merged(n2v_params)
    if random_seed is None:
        raise ValueError("a resumable walk needs an explicit random_seed (resumed chunks must share it)")
    os.makedirs(out_dir, exist_ok=True)
    start = graph.start_vertices()
    if walk_seed is not None:
        ids = torch.as_tensor(np.unique(np.asarray(list(walk_seed), dtype=np.int64)), device=start.device)
        start = start[torch.isin(start.to(torch.int64), ids)]
    ident = {"graph": graph_fingerprint(graph), "seed": int(random_seed), "chunk_starts": int(chunk_starts),
             "n_start": int(start.numel()),
             "params": {k: (float(prm[k]) if k.endswith("param") else int(prm[k])) for k in NODE2VEC_PARAMS}}
    path = os.path.join(out_dir, MANIFEST)
    manifest = {"identity": ident, "chunks": {}, "complete": False}
    if os.path.exists(path):
        with open(path) as f:
            old = json.load(f)
        if old.get("identity") != ident:
            raise ValueError(f"{path} belongs to another job (graph, seed, chunking or parameters differ)")
        manifest = old
    n_chunks = (int(start.numel()) + chunk_starts - 1) // chunk_starts
    done_now = 0
    for k in range(n_chunks):
        rec = manifest["chunks"].get(str(k))
        if rec is not None and os.path.exists(os.path.join(out_dir, rec["file"])):
            continue
        if max_chunks is not None and done_now >= max_chunks:
            break
        rows = _walk_alive(graph, start[k * chunk_starts:(k + 1) * chunk_starts], prm, random_seed)
        name = f"walks_{k:06d}.npy"
        tmp = os.path.join(out_dir, name + ".tmp")
        with open(tmp, "wb") as f:                                   # atomic: a torn shard is never listed
            np.save(f, rows.cpu().numpy())
        os.replace(tmp, os.path.join(out_dir, name))
        manifest["chunks"][str(k)] = {"file": name, "rows": int(rows.shape[0])}
        done_now += 1
        _write_manifest(path, manifest)
    manifest["complete"] = len(manifest["chunks"]) == n_chunks and all(
        os.path.exists(os.path.join(out_dir, c["file"])) for c in manifest["chunks"].values())
    manifest["n_chunks"] = n_chunks
    _write_manifest(path, manifest)
    return manifest


def _write_manifest(path: str, manifest: Dict[str, Any]) -> None:
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        json.dump(manifest, f, sort_keys=True)
    os.replace(tmp, path)


def load_walk_shards(out_dir: str, device: Optional[torch.device] = None) -> WalkFrame:
    """Concatenate the shards of a COMPLETE resumable run, in chunk order, into a WalkFrame."""
    with open(os.path.join(out_dir, MANIFEST)) as f:
        manifest = json.load(f)
    if not manifest.get("complete"):
        raise ValueError(f"{out_dir}: the walk job is not complete ({len(manifest['chunks'])} of "
                         f"{manifest.get('n_chunks', '?')} chunks)")
    parts = [np.load(os.path.join(out_dir, manifest["chunks"][str(k)]["file"])) for k in range(manifest["n_chunks"])]
    length = int(manifest["identity"]["params"]["walk_length"]) + 1
    host = np.concatenate(parts) if parts else np.zeros((0, length), dtype=np.int32)
    dev = torch.as_tensor(host, device=device) if device is not None else torch.as_tensor(host)
    return WalkFrame(dev, walks_host=host)


# ------------------------------------------------------------------------------------------------
# incremental re-walks
# ------------------------------------------------------------------------------------------------
def stale_start_vertices(walks: torch.Tensor, changed_ids: torch.Tensor) -> torch.Tensor:
    """Start vertices owning at least one walk that VISITS a changed vertex (a vertex whose out-arcs
    or their weights changed).  Those are the only walks whose law changed: a walker's next-hop law
    depends on the adjacency of its current and previous vertex only (randomwalk.py:316-333)."""
    if walks.numel() == 0 or changed_ids.numel() == 0:
        return torch.zeros(0, dtype=torch.int64, device=walks.device)
    hit = torch.isin(walks, changed_ids.to(walks.dtype)).any(dim=1)
    return torch.unique(walks[hit, 0].to(torch.int64))


def rewalk(prev: WalkFrame, graph: DeviceGraph, changed_ids: Sequence[int], n2v_params: Dict[str, Any],
           random_seed: int) -> Tuple[WalkFrame, Dict[str, int]]:
    """Refresh ``prev`` (walks of the OLD graph under ``random_seed``) for ``graph`` (the NEW graph),
    given the ids of the vertices whose out-arcs changed (added, removed or re-weighted arcs; new
    vertices included).  Re-walks only the start vertices that own an invalidated walk plus the
    changed vertices themselves, and splices the rows back in start-vertex order.  With the same
    seed the result is identical to walking the whole new graph (see the module docstring).
    Returns (walks, {"rewalked_starts", "rewalked_rows", "kept_rows"})."""
    prm = _merged(n2v_params)
    dev = graph.device
    old = prev.walks_device.to(dev)
    if old.numel() and old.shape[1] != int(prm["walk_length"]) + 1:
        raise ValueError("prev was walked with another walk_length")
    changed = torch.as_tensor(np.unique(np.asarray(list(changed_ids), dtype=np.int64)), device=dev)
    start = graph.start_vertices()
    # walkers the OLD run dropped at a sink (fugue.py:147) left no row in `prev`: a sink that gained
    # out-arcs would let them live now, so every start vertex holding fewer than num_walks rows --
    # or none at all -- is re-walked as well
    num_walks = int(prm["num_walks"])
    if old.numel():
        owners, rows = torch.unique(old[:, 0].to(torch.int64), return_counts=True)
        short = owners[rows < num_walks]
        absent = start.to(torch.int64)[~torch.isin(start.to(torch.int64), owners)]
    else:
        short, absent = torch.zeros(0, dtype=torch.int64, device=dev), start.to(torch.int64)
    stale = torch.unique(torch.cat([stale_start_vertices(old, changed), changed, short, absent]))
    redo = start[torch.isin(start.to(torch.int64), stale)]
    fresh = _walk_alive(graph, redo, prm, random_seed)
    keep = old[~torch.isin(old[:, 0].to(torch.int64), stale)] if old.numel() else old.reshape(0, fresh.shape[1])
    merged = torch.cat([keep, fresh.to(keep.dtype)])
    order = torch.sort(merged[:, 0], stable=True).indices          # rows of one start vertex stay in walk order
    info = {"rewalked_starts": int(redo.numel()), "rewalked_rows": int(fresh.shape[0]), "kept_rows": int(keep.shape[0])}
    return WalkFrame(merged[order]), info


# ------------------------------------------------------------------------------------------------
# link prediction and the (p, q) sweep
# ------------------------------------------------------------------------------------------------
def split_edges(src: torch.Tensor, dst: torch.Tensor, n_vertices: int, holdout: float = 0.1, seed: int = 0):
    """Hold out a seeded ``holdout`` fraction of the UNDIRECTED edges {a, b} (given once each, in any
    orientation) such that every vertex keeps at least one training edge: each vertex first protects
    one random incident edge, the hold-out is drawn from the unprotected rest.  Negatives: as many
    seeded vertex pairs that are not edges.  Everything on the device.
    Returns (train_a, train_b, pos [m, 2], neg [m, 2])."""
    dev = src.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    a, b = torch.minimum(src, dst).to(torch.int64), torch.maximum(src, dst).to(torch.int64)
    key = torch.unique(a * n_vertices + b)
    key = key[(key // n_vertices) != (key % n_vertices)]
    a, b = key // n_vertices, key % n_vertices
    m = int(key.numel())
    # protect, for every vertex, the incident edge with the smallest random tag
    tag = torch.rand(m, device=dev, generator=gen)
    best = torch.full((n_vertices,), 2.0, device=dev)
    best.scatter_reduce_(0, a, tag, "amin")
    best.scatter_reduce_(0, b, tag, "amin")
    protected = (tag == best[a]) | (tag == best[b])
    free = torch.nonzero(~protected).view(-1)
    want = min(int(m * holdout), int(free.numel()))
    held = free[torch.randperm(int(free.numel()), device=dev, generator=gen)[:want]]
    mask = torch.zeros(m, dtype=torch.bool, device=dev)
    mask[held] = True
    pos = torch.stack([a[mask], b[mask]], dim=1)
    neg = torch.zeros((0, 2), dtype=torch.int64, device=dev)
    while neg.shape[0] < want:                                   # rejection: sparse graphs accept almost all
        n_draw = int((want - neg.shape[0]) * 1.2) + 16
        x = torch.randint(0, n_vertices, (n_draw,), device=dev, generator=gen)
        y = torch.randint(0, n_vertices, (n_draw,), device=dev, generator=gen)
        lo, hi = torch.minimum(x, y), torch.maximum(x, y)
        k = lo * n_vertices + hi
        ok = (lo != hi) & ~torch.isin(k, key)
        k = torch.unique(k[ok])
        if neg.shape[0]:
            k = k[~torch.isin(k, neg[:, 0] * n_vertices + neg[:, 1])]
        k = k[torch.randperm(int(k.numel()), device=dev, generator=gen)]
        neg = torch.cat([neg, torch.stack([k // n_vertices, k % n_vertices], dim=1)])[:want]
    return a[~mask], b[~mask], pos, neg


def auc_from_scores(pos_scores: torch.Tensor, neg_scores: torch.Tensor) -> float:
    """Area under the ROC curve = P(score+ > score-) + P(tie)/2, by the rank-sum (Mann-Whitney)
    identity with average ranks for ties; fp64 on the device, no sklearn needed."""
    n_pos, n_neg = int(pos_scores.numel()), int(neg_scores.numel())
    if n_pos == 0 or n_neg == 0:
        raise ValueError("need at least one positive and one negative pair")
    s = torch.cat([pos_scores, neg_scores]).to(torch.float64)
    vals, inverse, counts = torch.unique(s, sorted=True, return_inverse=True, return_counts=True)
    upper = torch.cumsum(counts, 0).to(torch.float64)               # rank of the last member of each tie group
    avg_rank = upper - (counts.to(torch.float64) - 1.0) / 2.0
    rank_sum = avg_rank[inverse[:n_pos]].sum()
    return float((rank_sum - n_pos * (n_pos + 1) / 2.0) / (float(n_pos) * float(n_neg)))


def link_auc(vectors: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> float:
    """Dot-product link-prediction AUC; ``vectors`` is indexed by vertex id (Word2Vec.syn0)."""
    def score(pairs):
        return (vectors[pairs[:, 0]].double() * vectors[pairs[:, 1]].double()).sum(dim=1)
    return auc_from_scores(score(pos), score(neg))


def pq_sweep(src: torch.Tensor, dst: torch.Tensor, n_vertices: int, p_values: Sequence[float],
             q_values: Sequence[float], n2v_params: Optional[Dict[str, Any]] = None,
             w2v_params: Optional[Dict[str, Any]] = None, holdout: float = 0.1, seed: int = 0) -> List[Dict[str, Any]]:
    """Grid search over (p, q) by held-out link-prediction AUC on an undirected, unweighted graph
    given as one (src, dst) row per edge.  The training graph's CSR, hash sets and alias tables are
    built ONCE and shared by all grid points (they do not depend on p, q); every grid point is one
    walk launch plus one SGNS fit.  Returns one record per grid point, best AUC first."""
    from .sgns import Word2Vec
    prm = dict(n2v_params or {})
    w2v = {"size": 128, "window": 5, "negative": 5, "min_count": 1, "iter": 1, "sg": 1, "batch_words": 10000}
    w2v.update(w2v_params or {})
    ta, tb, pos, neg = split_edges(src, dst, n_vertices, holdout, seed)
    arcs_s, arcs_d = torch.cat([ta, tb]).to(torch.int32), torch.cat([tb, ta]).to(torch.int32)
    graph = DeviceGraph.from_arcs(arcs_s, arcs_d, None, n_vertices=n_vertices)
    start = graph.start_vertices()
    out = []
    for p in p_values:
        for q in q_values:
            pt = _merged({**prm, "return_param": p, "inout_param": q})
            t0 = time.perf_counter()
            walks = _walk_alive(graph, start, pt, seed)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            model = Word2Vec(walks, seed=seed + 1, **w2v)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            out.append({"p": float(p), "q": float(q), "auc": link_auc(model.syn0, pos, neg),
                        "walk_s": t1 - t0, "sgns_s": t2 - t1, "walks": int(walks.shape[0])})
    out.sort(key=lambda r: -r["auc"])
    return out
