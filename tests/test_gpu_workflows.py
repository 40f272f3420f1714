"""node2vec_b200.workflows on the GPU: a chunked/resumed walk and an incremental re-walk must be
bit-identical to one monolithic walk of the final graph under the same seed; the (p, q) sweep runs
end to end and ranks grid points by held-out link-prediction AUC."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wf():
    assert torch.cuda.is_available()
    from node2vec_b200 import _lib, workflows
    _lib.load()
    return workflows


def _sym_graph(n, m, seed, drop_deg0=True):
    rng = np.random.default_rng(seed)
    a, b = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = a != b
    a, b = a[keep], b[keep]
    key = np.unique(np.minimum(a, b) * n + np.maximum(a, b))
    return key // n, key % n


def _build(a, b, n, weight=None):
    from node2vec_b200.graph import DeviceGraph
    src, dst = np.r_[a, b].astype(np.int32), np.r_[b, a].astype(np.int32)
    w = None if weight is None else np.r_[weight, weight]
    return DeviceGraph.from_arcs(src, dst, w, n_vertices=n)


PRM = {"num_walks": 3, "walk_length": 12, "return_param": 0.5, "inout_param": 2.0}


def test_resumable_walk_equals_monolithic(wf, tmp_path):
    a, b = _sym_graph(3000, 20000, 1)
    g = _build(a, b, 3000)
    full, alive, _ = g.walk(g.start_vertices(), 3, 12, 0.5, 2.0, 77)
    assert bool(alive.all())
    out = str(tmp_path / "job")
    m = wf.random_walk_resumable(g, dict(PRM), out, random_seed=77, chunk_starts=700, max_chunks=2)   # "interrupted"
    assert not m["complete"] and len(m["chunks"]) == 2 and m["n_chunks"] == 5
    with pytest.raises(ValueError):
        wf.load_walk_shards(out)
    mtimes = {c["file"]: os.path.getmtime(os.path.join(out, c["file"])) for c in m["chunks"].values()}
    with pytest.raises(ValueError):                                                                   # another job
        wf.random_walk_resumable(g, dict(PRM), out, random_seed=78, chunk_starts=700)
    with pytest.raises(ValueError):
        wf.random_walk_resumable(g, {**PRM, "inout_param": 3.0}, out, random_seed=77, chunk_starts=700)
    m = wf.random_walk_resumable(g, dict(PRM), out, random_seed=77, chunk_starts=700)                 # resume
    assert m["complete"] and len(m["chunks"]) == 5
    for f, t in mtimes.items():
        assert os.path.getmtime(os.path.join(out, f)) == t                                          # not redone
    got = wf.load_walk_shards(out)
    assert got.walks.dtype == np.int32 and np.array_equal(got.walks, full.cpu().numpy())
    assert json.load(open(os.path.join(out, "manifest.json")))["identity"]["seed"] == 77
    # a lost shard is redone on the next call, nothing else
    os.remove(os.path.join(out, "walks_000003.npy"))
    m = wf.random_walk_resumable(g, dict(PRM), out, random_seed=77, chunk_starts=700)
    assert m["complete"] and np.array_equal(wf.load_walk_shards(out).walks, full.cpu().numpy())


def test_resumable_walk_with_sinks_and_seed_ids(wf, tmp_path):
    from node2vec_b200.graph import DeviceGraph
    rng = np.random.default_rng(4)
    src, dst = rng.integers(0, 500, 1500).astype(np.int32), rng.integers(0, 800, 1500).astype(np.int32)   # 500..799 are sinks
    g = DeviceGraph.from_arcs(src, dst, rng.uniform(0.1, 2.0, 1500), n_vertices=800)
    ids = list(range(0, 500, 3))
    start = g.start_vertices()
    start = start[torch.isin(start.long(), torch.as_tensor(ids, device=start.device))]
    walks, alive, _ = g.walk(start, 3, 12, 0.5, 2.0, 5)
    assert not bool(alive.all())
    m = wf.random_walk_resumable(g, dict(PRM), str(tmp_path / "j"), random_seed=5, chunk_starts=64, walk_seed=ids)
    assert m["complete"]
    assert np.array_equal(wf.load_walk_shards(str(tmp_path / "j")).walks, walks[alive].cpu().numpy())


@pytest.mark.parametrize("weighted", [False, True])
def test_incremental_rewalk_equals_full_rewalk(wf, weighted):
    from node2vec_b200.fugue import WalkFrame
    n = 4000
    a, b = _sym_graph(n, 16000, 2)
    rng = np.random.default_rng(9)
    w = rng.uniform(0.2, 3.0, len(a)) if weighted else None
    g_old = _build(a, b, n, w)
    old, alive, _ = g_old.walk(g_old.start_vertices(), 3, 12, 0.5, 2.0, 31)
    prev = WalkFrame(old[alive])
    # change: drop 5 edges, add 5 new ones (one touching a previously isolated vertex if there is one)
    drop = rng.choice(len(a), 5, replace=False)
    keep = np.ones(len(a), dtype=bool)
    keep[drop] = False
    present = set(zip(a.tolist(), b.tolist()))
    add = []
    while len(add) < 5:
        x, y = sorted(rng.integers(0, n, 2).tolist())
        if x != y and (x, y) not in present:
            add.append((x, y))
            present.add((x, y))
    a2 = np.r_[a[keep], [x for x, _ in add]]
    b2 = np.r_[b[keep], [y for _, y in add]]
    w2 = None if w is None else np.r_[w[keep], rng.uniform(0.2, 3.0, 5)]
    changed = sorted(set(a[drop].tolist()) | set(b[drop].tolist()) | {v for e in add for v in e})
    g_new = _build(a2, b2, n, w2)
    assert g_new.flags == g_old.flags                 # same sampler class, or the identity does not hold
    full, alive2, _ = g_new.walk(g_new.start_vertices(), 3, 12, 0.5, 2.0, 31)
    got, info = wf.rewalk(prev, g_new, changed, dict(PRM), random_seed=31)
    assert torch.equal(got.walks_device, full[alive2])
    assert info["rewalked_rows"] + info["kept_rows"] == int(alive2.sum())
    assert info["rewalked_starts"] < 0.5 * int(g_new.start_vertices().numel())       # it really was incremental
    assert info["kept_rows"] > 0


def test_pq_sweep_ranks_by_auc(wf):
    """Stochastic block model (10 communities): held-out links are predictable from the embeddings, so
    every grid point must clear AUC 0.7; the records come back best first."""
    rng = np.random.default_rng(7)
    n, blocks = 1000, 10
    iu, ju = np.triu_indices(n, 1)
    same = (iu // (n // blocks)) == (ju // (n // blocks))
    keep = rng.random(len(iu)) < np.where(same, 0.10, 0.003)
    src, dst = torch.as_tensor(iu[keep]).cuda(), torch.as_tensor(ju[keep]).cuda()
    res = wf.pq_sweep(src, dst, n, [0.5, 2.0], [0.5, 2.0], {"num_walks": 10, "walk_length": 40},
                      {"size": 64, "iter": 5}, seed=1)
    assert len(res) == 4 and {(r["p"], r["q"]) for r in res} == {(0.5, 0.5), (0.5, 2.0), (2.0, 0.5), (2.0, 2.0)}
    aucs = [r["auc"] for r in res]
    print("pq sweep:", [(r["p"], r["q"], round(r["auc"], 4)) for r in res])
    assert aucs == sorted(aucs, reverse=True) and aucs[-1] > 0.7 and aucs[0] <= 1.0
    assert all(r["walks"] == 10 * n and r["walk_s"] > 0 and r["sgns_s"] > 0 for r in res)


def test_link_auc_matches_sklearn_protocol(wf):
    """The product's device AUC against the oracle's sklearn scorer on the same vectors and pairs."""
    from oracle import linkpred
    gen = torch.Generator().manual_seed(2)
    vec = torch.randn(300, 16, generator=gen)
    pos = torch.randint(0, 300, (500, 2), generator=gen)
    neg = torch.randint(0, 300, (400, 2), generator=gen)
    got = wf.link_auc(vec.cuda(), pos.cuda(), neg.cuda())
    want = linkpred.auc_dot(vec.double().numpy(), pos.numpy(), neg.numpy())
    assert got == pytest.approx(want, abs=1e-9)
