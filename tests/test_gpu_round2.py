"""Round-2 GPU parity gates (all through the C ABI):

* chi-square of device transition frequencies against the reference's law
  (randomwalk.py:219-231) on BASELINE-sized graphs -- ER-10k and the BlogCatalog-shaped graph at
  (p, q) in {(1,1), (1,.5), (.25,4), (4,.25)} -- conditioned on the highest-traffic (t, v) pairs AND
  on hub v with degree > 1000 (the fp32 fold threshold and the two-threshold skip at large degree);
* giant (> 12288 arcs) unit-weight vertices through the parallel alias path, bit-exact;
* vertex ids outside the graph raise instead of faulting; out-of-vocabulary tokens are skipped;
* the reference's trim -> symmetrise order on arc tensors; rewalk after a sink gains out-arcs.
"""
import numpy as np
import pytest

from oracle import clib, ref_walk
from tests.helpers import chi_square_ok, graph_flags, pack_arcs

pytestmark = pytest.mark.gpu

ALPHA = 1e-4          # per-test significance of the chi-square gates (the stated tolerance)


@pytest.fixture(scope="module")
def n2v():
    import torch
    assert torch.cuda.is_available()
    from node2vec_b200 import _lib, fugue, graph, synth, workflows
    _lib.load()

    class NS:
        pass
    ns = NS()
    ns.torch, ns.lib, ns.graph, ns.fugue, ns.synth, ns.workflows = torch, _lib, graph, fugue, synth, workflows
    return ns


_GRAPHS = {}


def _bench_graph(n2v, name):
    """(DeviceGraph, row_ptr, col) of a BASELINE graph, cached per module."""
    if name not in _GRAPHS:
        if name == "er_10k":
            src, dst = n2v.synth.erdos_renyi(10000, 100000, seed=42)
        else:
            src, dst = n2v.synth.blogcatalog_like(10000, 334000, seed=42)
        g = n2v.graph.DeviceGraph.from_arcs(src, dst, None, n_vertices=10000)
        row_ptr, col, _, _ = clib.csr_from_arcs(src, dst, None, 10000)
        _GRAPHS[name] = (g, row_ptr, col)
    return _GRAPHS[name]


def _law(row_ptr, col, t, v, p, q):
    """The reference's next-vertex law at (t, v) on a unit-weight graph (randomwalk.py:219-231):
    weight / p back to t, weight into N_out(t), weight / q elsewhere, normalised."""
    nb = col[row_ptr[v]:row_ptr[v + 1]]
    nt = col[row_ptr[t]:row_ptr[t + 1]]
    wts = np.where(nb == t, 1.0 / p, np.where(np.isin(nb, nt), 1.0, 1.0 / q))
    return nb, wts / wts.sum()


@pytest.mark.parametrize("name", ["er_10k", "blogcatalog_like"])
@pytest.mark.parametrize("p,q", [(1.0, 1.0), (1.0, 0.5), (0.25, 4.0), (4.0, 0.25)])
def test_transition_frequencies_at_baseline_size(n2v, name, p, q):
    torch = n2v.torch
    g, row_ptr, col = _bench_graph(n2v, name)
    deg = np.diff(row_ptr)
    # cross-check the vectorised law against the oracle's restatement on one pair
    t0 = int(np.argmax(deg)); v0 = int(col[row_ptr[t0]])
    adj = {u: (col[row_ptr[u]:row_ptr[u + 1]].tolist(), [1.0] * int(deg[u])) for u in (t0, v0)}
    nb, pr = _law(row_ptr, col, t0, v0, p, q)
    want = ref_walk.transition_law(adj, t0, v0, p, q)
    assert np.allclose(pr, [want[int(x)] for x in nb], rtol=1e-12)

    # (a) the highest-traffic (t, v) pairs of a BASELINE-shaped run
    start = g.start_vertices()
    walks, alive, _ = g.walk(start, 40, 12, p, q, seed=1234)
    assert bool(alive.all())
    w = walks.long()
    trip = torch.stack([w[:, :-2].reshape(-1), w[:, 1:-1].reshape(-1), w[:, 2:].reshape(-1)], 1)
    key = trip[:, 0] * 10000 + trip[:, 1]
    uniq, cnt = torch.unique(key, return_counts=True)
    top = uniq[torch.argsort(cnt, descending=True)[:6]].cpu().tolist()
    pairs = [(k // 10000, k % 10000) for k in top]
    # (b) hubs: the largest-degree v, entered from its lowest-degree neighbour t (so a walker started at t
    # lands on v with probability 1 / deg(t)); includes deg(v) > 1000 on the BlogCatalog-shaped graph
    hubs = np.argsort(-deg)[:3]
    for v in hubs:
        nbv = col[row_ptr[v]:row_ptr[v + 1]]
        t = int(nbv[np.argmin(deg[nbv])])
        pairs.append((t, int(v)))
    if name == "blogcatalog_like":
        assert deg[hubs[0]] > 1000
    for t, v in pairs:
        n_walkers = int(min(3_000_000, max(200_000, 300 * deg[v] * deg[t])))
        ws, al, _ = g.walk(torch.tensor([t], dtype=torch.int32, device="cuda"), n_walkers, 2, p, q, seed=99 + t)
        ws = ws[al]
        x = ws[ws[:, 1] == v][:, 2].cpu().numpy()
        nb, pr = _law(row_ptr, col, t, v, p, q)
        assert np.isin(x, nb).all()
        counts = np.bincount(np.searchsorted(nb, x), minlength=len(nb))
        ok, pval = chi_square_ok(counts, pr, ALPHA)
        assert ok, (name, p, q, t, v, int(deg[v]), len(x), pval)
        # the return arc on its own (fold component): binomial z-score
        k = int(counts[np.searchsorted(nb, t)]); n = len(x); pt = float(pr[np.searchsorted(nb, t)])
        z = (k - n * pt) / max(np.sqrt(n * pt * (1 - pt)), 1e-9)
        assert abs(z) < 4.5, (name, p, q, t, v, k, n * pt, z)


def test_alias_giant_unit_and_weighted_hubs_bit_exact(n2v):
    """Vertices beyond the shared-memory path (deg > 12288): all-1.0 weights take the parallel
    trivial construction, anything else the sequential one -- both bit-exact vs the oracle."""
    rng = np.random.default_rng(3)
    n = 70000
    src = [rng.integers(100, n, 40000)]; dst = [rng.integers(0, n, 40000)]; w = [rng.uniform(0.5, 2.0, 40000)]
    for v, d, unit in ((5, 20000, True), (6, 65537, True), (7, 13000, False), (8, 12289, True)):
        src.append(np.full(d, v)); dst.append(rng.permutation(n)[:d]); w.append(np.ones(d) if unit else rng.uniform(0.5, 2, d))
    src, dst, w = np.concatenate(src), np.concatenate(dst), np.concatenate(w)
    w[-3] = 1.0000000000000002                      # vertex 8: one weight an ulp off 1.0 -> sequential path
    for mode in ("naive", "neumaier"):
        g = n2v.graph.DeviceGraph.from_arcs(src, dst, w, n_vertices=n, sum_mode=mode, keep_tables=True)
        row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, w, n)
        alias, probs, bad = clib.alias_tables_csr(row_ptr, ws, mode, threads=4)
        h = g.to_host()
        assert bad == 0 and np.array_equal(h["alias"], alias)
        assert np.array_equal(h["probs"].view(np.uint64), probs.view(np.uint64))
        thr, adst, aalias = pack_arcs(row_ptr, col, alias, probs)
        assert np.array_equal(h["thr"], thr) and np.array_equal(h["dst"], adst) and np.array_equal(h["alias_dst"], aalias)
        assert np.array_equal(h["dst_deg"], h["deg"][adst]) and np.array_equal(h["adst_deg"], h["deg"][aalias])
        assert h["wsum"][5] == np.float32(20000.0) and h["wsum"][6] == np.float32(65537.0)


def test_alias_build_on_an_all_unit_graph_is_the_arc_parallel_path_and_bit_exact(n2v):
    """Every weight exactly 1.0 (weight=None or an explicit column of ones -- all BASELINE configs): K1 writes
    the records one thread per arc.  Same records, tables and wsum as the oracle's sequential construction,
    on every degree class (thread / CTA / giant), multi-arcs and isolated vertices included; one weight an
    ulp off 1.0 anywhere sends the whole graph back through the per-vertex kernels."""
    rng = np.random.default_rng(12)
    n = 40000
    src = [rng.integers(100, n - 50, 60000)]; dst = [rng.integers(0, n, 60000)]
    for v, d in ((3, 255), (4, 256), (8, 5000), (9, 12288), (10, 12289), (7, 30000)):
        src.append(np.full(d, v)); dst.append(rng.integers(0, n, d))         # with repeated neighbours
    src, dst = np.concatenate(src), np.concatenate(dst)
    ones = np.ones(len(src))
    off = ones.copy(); off[len(off) // 2] = 0.9999999999999999
    for mode in ("naive", "neumaier"):
        for w in (None, ones, off):
            g = n2v.graph.DeviceGraph.from_arcs(src, dst, w, n_vertices=n, sum_mode=mode, keep_tables=True)
            row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, w, n)
            alias, probs, bad = clib.alias_tables_csr(row_ptr, ws, mode, threads=4)
            h = g.to_host()
            assert bad == 0 and np.array_equal(h["alias"], alias)
            assert np.array_equal(h["probs"].view(np.uint64), probs.view(np.uint64))
            thr, adst, aalias = pack_arcs(row_ptr, col, alias, probs)
            assert np.array_equal(h["thr"], thr) and np.array_equal(h["dst"], adst)
            assert np.array_equal(h["alias_dst"], aalias) and np.array_equal(h["alias_idx"], alias)
            assert np.array_equal(h["dst_base"], h["base"][adst].astype(np.uint32)) and np.array_equal(h["dst_deg"], h["deg"][adst])
            assert np.array_equal(h["adst_base"], h["base"][aalias].astype(np.uint32)) and np.array_equal(h["adst_deg"], h["deg"][aalias])
            deg = np.diff(row_ptr)
            wsum = np.array([np.float32(ref_walk.float_sum(ws[a:b].tolist())) for a, b in zip(row_ptr[:-1], row_ptr[1:])],
                            dtype=np.float32)
            assert np.array_equal(h["wsum"][deg > 0], wsum[deg > 0]) and not h["wsum"][deg == 0].any()
            if w is not off:
                assert (h["thr"] == 0xFFFFFFFF).all() and not h["alias"].any()


def test_symmetric_flag_under_single_arc_mutations(n2v):
    """K0's SYMMETRIC flag selects the walk's fast fold, so it has to be exact.  The check searches from one
    side of every mirrored pair only (the arc whose head has the smaller (degree, id)) and balances the
    counts; here a symmetric weighted graph with a hub, equal-degree neighbours and a self-loop is mutated
    one arc at a time and the flags are compared with a set-based check on the host."""
    rng = np.random.default_rng(21)
    n = 400
    a = rng.integers(0, n, 3000); b = rng.integers(0, n, 3000)
    hub = np.full(350, 7); leaves = rng.permutation(n)[:350]
    a, b = np.concatenate([a, hub, [5, 9]]), np.concatenate([b, leaves, [5, 10]])      # (5,5): a self-loop; (9,10)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key, idx = np.unique(lo.astype(np.int64) << 32 | hi, return_index=True)
    lo, hi = lo[idx], hi[idx]
    w = rng.uniform(0.5, 2.0, len(lo))
    loop = lo == hi
    src = np.concatenate([lo, hi[~loop]]); dst = np.concatenate([hi, lo[~loop]]); ws = np.concatenate([w, w[~loop]])

    def flags_of(s, d, x):
        g = n2v.graph.DeviceGraph.from_arcs(s, d, x, n_vertices=n)
        row_ptr, col, wsorted, _ = clib.csr_from_arcs(s, d, x, n)
        want = graph_flags(row_ptr, col, wsorted)
        assert g.flags == want, (g.flags, want)
        return g.flags

    SYM = n2v.lib.GRAPH_SYMMETRIC
    assert flags_of(src, dst, ws) & SYM
    deg = np.bincount(src, minlength=n)
    picks = list(rng.integers(0, len(src), 12))
    picks += [int(np.flatnonzero(src == 7)[0]), int(np.flatnonzero(dst == 7)[0])]              # hub-side and leaf-side arcs
    same = np.flatnonzero((deg[src] == deg[dst]) & (src != dst))
    picks += [int(same[0]), int(same[-1])] if len(same) else []                               # the (degree, id) tie-break
    for i in picks:
        if src[i] == dst[i]:
            continue
        keep = np.ones(len(src), dtype=bool); keep[i] = False
        assert not flags_of(src[keep], dst[keep], ws[keep]) & SYM                             # a mirror is missing
        w2 = ws.copy(); w2[i] = np.nextafter(w2[i], 3.0)
        assert not flags_of(src, dst, w2) & SYM                                               # a mirror weighs an ulp more
    free = [(u, v) for u in range(20) for v in range(20, 40) if u != v and not ((src == u) & (dst == v)).any()][:3]
    for u, v in free:                                                                         # a one-way arc
        assert not flags_of(np.append(src, u), np.append(dst, v), np.append(ws, 1.0)) & SYM
        assert flags_of(np.append(src, [u, v]), np.append(dst, [v, u]), np.append(ws, [1.5, 1.5])) & SYM
    assert flags_of(np.append(src, 11), np.append(dst, 11), np.append(ws, 1.0)) & SYM         # one more self-loop


def test_bad_vertex_ids_raise_and_do_not_poison_the_context(n2v):
    torch = n2v.torch
    src = np.array([0, 1, 2, -1, 3], dtype=np.int64); dst = np.array([1, 2, 3, 0, 0], dtype=np.int64)
    with pytest.raises(ValueError):
        n2v.graph.DeviceGraph.from_arcs(src, dst, None, n_vertices=4)
    with pytest.raises(ValueError):
        n2v.graph.DeviceGraph.from_arcs(np.array([0, 7]), np.array([1, 0]), None, n_vertices=4)
    with pytest.raises(ValueError):                      # would wrap to a valid int32 id
        n2v.graph.DeviceGraph.from_arcs(np.array([0, (1 << 32) + 1]), np.array([1, 0]), None, n_vertices=4)
    with pytest.raises(ValueError):
        n2v.graph.DeviceGraph.from_arcs(torch.tensor([0, 1 << 33], device="cuda"), torch.tensor([1, 0], device="cuda"),
                                        None, n_vertices=4)
    torch.cuda.synchronize()                             # no sticky fault: the context still works
    g = n2v.graph.DeviceGraph.from_arcs(np.array([0, 1]), np.array([1, 0]), None, n_vertices=2)
    walks, alive, _ = g.walk(g.start_vertices(), 2, 3, seed=1)
    assert walks.cpu().numpy().tolist() == [[0, 1, 0, 1], [0, 1, 0, 1], [1, 0, 1, 0], [1, 0, 1, 0]]


def test_sgns_tokens_beyond_the_table_are_out_of_vocabulary(n2v):
    """gensim skips tokens it has no vocabulary entry for; ids >= the table size fixed by
    build_vocab must be skipped too, never dereferenced."""
    torch = n2v.torch
    from node2vec_b200.sgns import Word2Vec
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    walks = torch.randint(0, 100, (512, 21), device="cuda", dtype=torch.int32, generator=gen)
    m = Word2Vec(size=32, sg=1, negative=5, min_count=1, iter=1, seed=3, sample=0.0)
    m.build_vocab(walks)
    dirty = walks.clone()
    dirty[::3, 5] = 1_000_000
    dirty[1::7, 0] = 2_000_000_000
    clean = dirty.clone()
    clean[dirty >= 100] = -1
    before = m.syn0.clone()
    pairs_dirty, kept_dirty = m.train(dirty, epochs=1)
    after_dirty = m.syn0.clone()
    m2 = Word2Vec(size=32, sg=1, negative=5, min_count=1, iter=1, seed=3, sample=0.0)
    m2.build_vocab(walks)
    assert torch.equal(m2.syn0, before)
    pairs_clean, kept_clean = m2.train(clean, epochs=1)
    assert (pairs_dirty, kept_dirty) == (pairs_clean, kept_clean) and kept_dirty < dirty.numel()
    assert bool(torch.isfinite(after_dirty).all())


def test_trim_index_on_tensors_follows_the_reference_order(n2v):
    """trim the arcs as listed, THEN expand to both directions (fugue.py:57-77 + indexer.py:45-48):
    the arc set equals the pandas path's (mapped back through its name table), and the graph is
    symmetric although a hub was trimmed."""
    import pandas as pd
    torch = n2v.torch
    rng = np.random.default_rng(4)
    a = rng.integers(0, 300, 4000); b = rng.integers(0, 300, 4000)
    hub = np.full(900, 7); other = rng.permutation(np.arange(300, 1300))[:900]
    lo = np.concatenate([np.minimum(a, b), hub]); hi = np.concatenate([np.maximum(a, b), other])
    key = np.unique((lo.astype(np.int64) << 32 | hi)[lo != hi])
    lo, hi = (key >> 32), (key & 0xFFFFFFFF)
    perm = rng.permutation(len(lo)); lo, hi = lo[perm], hi[perm]
    ts, td = torch.as_tensor(lo, device="cuda").int(), torch.as_tensor(hi, device="cuda").int()
    (s, d), none = n2v.fugue.trim_index(None, (ts, td), indexed=True, directed=False, max_out_deg=100, random_seed=3)
    assert none is None
    got = set(zip(s.cpu().tolist(), d.cpu().tolist()))
    frame, names = n2v.fugue.trim_index(None, pd.DataFrame({"src": lo, "dst": hi}), indexed=False, directed=False,
                                        max_out_deg=100, random_seed=3)
    nm = dict(zip(names.as_pandas()["vertex_id"].tolist(), names.as_pandas()["vertex_name"].tolist()))
    f = frame.as_pandas()
    want = set(zip((nm[i] for i in f["src"].tolist()), (nm[i] for i in f["dst"].tolist())))
    assert got == want and len(got) == len(s)
    g = n2v.graph.DeviceGraph.from_arcs(s, d, None, n_vertices=1300)
    assert g.flags & 7 == 7                                         # unit, SYMMETRIC, simple
    deg = g.degrees().cpu().numpy()
    assert deg[7] >= 100                                            # 100 kept out-arcs + mirrored in-arcs


def test_rewalk_revives_walkers_that_died_at_a_former_sink(n2v):
    """a -> z with z a sink drops the walkers started at a (fugue.py:147); once z gains z -> b a full
    walk of the new graph has those rows, so the incremental re-walk must produce them too."""
    torch = n2v.torch
    src = np.array([0, 1, 1, 2, 3], dtype=np.int64); dst = np.array([4, 2, 3, 1, 1], dtype=np.int64)   # 4 = sink
    prm = {"num_walks": 3, "walk_length": 4, "return_param": 1.0, "inout_param": 1.0}
    old = n2v.fugue.random_walk(None, (src, dst), dict(prm), None, random_seed=5)
    assert 0 not in old.walks[:, 0]                                 # every walker from 0 died at vertex 4
    src2, dst2 = np.append(src, 4), np.append(dst, 1)
    g2 = n2v.graph.DeviceGraph.from_arcs(src2, dst2, None, n_vertices=5)
    full = n2v.fugue.random_walk(None, None, dict(prm), None, random_seed=5, graph=g2)
    inc, info = n2v.workflows.rewalk(old, g2, [4], dict(prm), 5)
    assert np.array_equal(inc.walks, full.walks) and (full.walks[:, 0] == 0).sum() == 3


@pytest.mark.parametrize("name", ["er_10k", "blogcatalog_like"])
def test_auc_gate_on_baseline_configs(n2v, name):
    """north_star gate at BASELINE sizes, through the reference-shaped API: embeddings from
    ``Node2VecGensim(...).fit()`` (D = 128, window 5, 5 negatives) reach the link-prediction AUC of the
    gensim-3.8 restatement trained on the SAME walk matrix within +-0.01 (mean of 3 seeds).
    configs[0]: ER 10k/100k, p=1 q=.5, 10 x 20.  configs[1] shape: BlogCatalog-like 10k/334k,
    p=.25 q=4, 10 walks x 40 (the bench uses 80 walks; 10 keep the CPU side of the gate short)."""
    import os
    from node2vec_b200.embedding import Node2VecGensim
    from oracle import linkpred
    if name == "er_10k":
        src, dst = n2v.synth.erdos_renyi(10000, 100000, seed=42)
        prm = {"num_walks": 10, "walk_length": 20, "return_param": 1.0, "inout_param": 0.5}
    else:
        src, dst = n2v.synth.blogcatalog_like(10000, 334000, seed=42)
        prm = {"num_walks": 10, "walk_length": 40, "return_param": 0.25, "inout_param": 4.0}
    n, half = 10000, len(src) // 2
    edges = np.stack([src[:half], dst[:half]], axis=1).astype(np.int64)
    train, pos, neg = linkpred.split_edges(edges, n, 0.1, 0)
    s = np.concatenate([train[:, 0], train[:, 1]]); d = np.concatenate([train[:, 1], train[:, 0]])
    res = n2v.fugue.random_walk(None, (s, d), dict(prm), random_seed=5)
    walks = res.walks
    counts = np.bincount(walks.reshape(-1), minlength=n)
    hp = dict(window=5, negative=5, alpha=0.025, min_alpha=1e-4, min_count=1, sample=1e-3)
    epochs = 3
    threads = max(1, min(32, os.cpu_count() or 1))
    auc_ref, auc_gpu = [], []
    for seed in (1, 2, 3):
        syn0, syn1 = clib.sgns_init(n, 128, seed)
        clib.sgns_train(walks, counts, syn0, syn1, epochs=epochs, seed=seed, batch_words=1000, threads=threads, **hp)
        auc_ref.append(linkpred.auc_dot(syn0, pos, neg))
        model = Node2VecGensim(res, {"sg": 1, "negative": 5, "iter": epochs, "min_count": 1}, window_size=5,
                               vector_size=128, random_seed=seed).fit()
        emb = np.zeros((n, 128), dtype=np.float32)
        emb[[int(t) for t in model.wv.index2word]] = model.wv.vectors
        auc_gpu.append(linkpred.auc_dot(emb, pos, neg))
    print(f"[{name}] AUC restatement {auc_ref} device {auc_gpu}")
    assert abs(np.mean(auc_gpu) - np.mean(auc_ref)) <= 0.01, (name, auc_gpu, auc_ref)


# ------------------------------------------------------------------ K6: the graph indexer on the device
def _frames_equal(a, b):
    assert list(a.columns) == list(b.columns), (list(a.columns), list(b.columns))
    assert len(a) == len(b)
    for c in a.columns:
        x, y = a[c].to_numpy(), b[c].to_numpy()
        if x.dtype.kind == "f":
            assert x.astype(np.float64).tobytes() == y.astype(np.float64).tobytes(), c
        else:
            assert x.tolist() == y.tolist(), c


def test_first_occurrence_kernel(n2v):
    """n2v_first_occurrence: first[i] = min { j : key[j] == key[i] } for 1-3 int64 key columns."""
    torch = n2v.torch
    from node2vec_b200 import preprocess
    rng = np.random.default_rng(2)
    for n, hi in ((1, 3), (1000, 10), (200000, 5000), (300000, 1 << 62)):
        a = rng.integers(-hi, hi, n)
        b = rng.integers(0, 3, n)
        c = rng.integers(0, 2, n)
        for cols in ((a,), (a, b), (a, b, c)):
            got = preprocess.first_occurrence(*(torch.as_tensor(x, device="cuda") for x in cols)).cpu().numpy()
            seen, want = {}, np.empty(n, dtype=np.int64)
            for i, key in enumerate(zip(*cols)):
                want[i] = seen.setdefault(key, i)
            assert np.array_equal(got, want), (n, len(cols))


def test_device_indexer_matches_the_reference_goldens(n2v):
    """index_graph_pandas' outputs as pinned by tests/golden/indexer.json (generated by the unmodified
    reference): the device path (string names ranked on the host, row-level work in K6) reproduces
    edge ids, weight bits, the name table and the undirected expansion row for row."""
    import pandas as pd
    from tests.helpers import load_golden
    fx = load_golden("indexer.json")
    assert len(fx["cases"]) >= 3
    for c in fx["cases"]:
        cols = {"src": c["src"], "dst": c["dst"]}
        if c["weight"] is not None:
            cols["weight"] = c["weight"]
        frame, names = n2v.fugue._index_graph_frame_on_device(pd.DataFrame(cols), c["directed"])
        e, nm = frame.as_pandas(), names.as_pandas()
        assert list(e.columns) == c["edge_columns"] and list(nm.columns) == c["name_id_columns"]
        assert e["src"].tolist() == c["edge_src"] and e["dst"].tolist() == c["edge_dst"]
        assert [float(x).hex() for x in e["weight"]] == c["edge_weight"] and str(e["weight"].dtype) == c["edge_weight_dtype"]
        assert nm["vertex_id"].tolist() == c["vertex_id"] and nm["vertex_name"].tolist() == c["vertex_name"]


@pytest.mark.parametrize("kind", ["int", "str"])
@pytest.mark.parametrize("directed", [True, False])
def test_trim_index_frame_device_path_equals_pandas_path(n2v, kind, directed, monkeypatch):
    """fugue.trim_index on a frame with names: the GPU path (default when a device is present) and
    the pandas path return the same two frames, row for row -- hot vertices trimmed with the
    reference's seeded sample, first-occurrence ids, weights with duplicates, undirected expansion."""
    import pandas as pd
    torch = n2v.torch
    rng = np.random.default_rng(17)
    a = np.concatenate([rng.integers(0, 60, 900), np.full(400, 7), np.full(260, 33)])
    b = rng.integers(0, 500, len(a))
    perm = rng.permutation(len(a)); a, b = a[perm], b[perm]
    w = rng.choice([0.5, 1.0, 2.0], len(a))
    if kind == "str":
        a, b = np.array([f"user{x}" for x in a]), np.array([f"user{x}" for x in b])
    else:
        a, b = a * 1000003 - 5, b * 1000003 - 5                       # sparse, partly negative integer names
    for with_w in (True, False):
        df = pd.DataFrame({"src": a, "dst": b, **({"weight": w} if with_w else {})})
        got_e, got_n = n2v.fugue.trim_index(None, df.copy(), indexed=False, directed=directed, max_out_deg=100,
                                            random_seed=9)
        monkeypatch.setattr(torch.cuda, "is_available", lambda: False)   # force the pandas path
        want_e, want_n = n2v.fugue.trim_index(None, df.copy(), indexed=False, directed=directed, max_out_deg=100,
                                              random_seed=9)
        monkeypatch.undo()
        _frames_equal(got_e.as_pandas(), want_e.as_pandas())
        _frames_equal(got_n.as_pandas(), want_n.as_pandas())
    # tensors with integer names: same ids as the frames; dense_ids renumbers in first-occurrence order
    if kind == "int":
        ts, td = torch.as_tensor(a, device="cuda"), torch.as_tensor(b, device="cuda")
        (s, d, wt), (vid, vname) = n2v.fugue.trim_index(None, (ts, td), indexed=False, directed=directed,
                                                        max_out_deg=100, random_seed=9)
        assert s.cpu().tolist() == want_e.as_pandas()["src"].tolist() and d.cpu().tolist() == want_e.as_pandas()["dst"].tolist()
        assert vid.cpu().tolist() == want_n.as_pandas()["vertex_id"].tolist()
        assert vname.cpu().tolist() == want_n.as_pandas()["vertex_name"].tolist()
        (s2, d2, _), (vid2, vname2) = n2v.fugue.trim_index(None, (ts, td), indexed=False, directed=directed,
                                                           max_out_deg=100, random_seed=9, dense_ids=True)
        assert vid2.cpu().tolist() == list(range(len(vid))) and torch.equal(vname2, vname)
        assert torch.equal(vname2[s2], vname[torch.searchsorted(vid, s)])


def test_data_parallel_averaging_auc_band(n2v):
    """Data-parallel SGNS = G replicas on contiguous walk shards, tables averaged every epoch
    (emulated on one GPU: replicas are independent between averages).  Averaging G replicas that each
    saw 1/G of the epoch is NOT the G = 1 computation: the mean moves by (1/G) * sum of the replicas'
    updates, i.e. an effective per-sample step of alpha / G with G-fold variance reduction -- the same
    trade Spark ML's per-partition Word2Vec makes (constants.py:34-35, numPartitions).  On a graph
    with community structure this does not cost link-prediction quality: gate = AUC(G) >= AUC(1) -
    0.01 for G in {2, 4, 8} (measured drift is positive, +0.005 / +0.02 / +0.02, reported in
    profiles/); |AUC(G) - AUC(1)| <= 0.035 bounds it from above."""
    torch = n2v.torch
    from node2vec_b200 import workflows as wf
    from node2vec_b200.sgns import Word2Vec
    rng = np.random.default_rng(7)
    n, blocks, E, DIM = 3000, 20, 5, 64
    iu, ju = np.triu_indices(n, 1)
    same = (iu // (n // blocks)) == (ju // (n // blocks))
    keep = rng.random(len(iu)) < np.where(same, 0.06, 0.001)
    src, dst = torch.as_tensor(iu[keep]).cuda(), torch.as_tensor(ju[keep]).cuda()
    ta, tb, pos, neg = wf.split_edges(src, dst, n, 0.1, seed=0)
    g = n2v.graph.DeviceGraph.from_arcs(torch.cat([ta, tb]).int(), torch.cat([tb, ta]).int(), None, n_vertices=n)
    walks, alive, _ = g.walk(g.start_vertices(), 10, 40, 1.0, 1.0, seed=5)
    W = int(walks.shape[0])
    auc = {}
    for G in (1, 2, 4, 8):
        vals = []
        for seed in (1, 2):
            m = Word2Vec(size=DIM, sg=1, negative=5, window=5, min_count=1, iter=E, seed=seed, batch_words=10000)
            m.build_vocab(walks)
            tabs = [(m.syn0.clone(), m.syn1neg.clone()) for _ in range(G)]
            bounds = [(W * r // G, W * (r + 1) // G) for r in range(G)]
            for ep in range(E):
                for r, (lo, hi) in enumerate(bounds):
                    m.syn0, m.syn1neg = tabs[r]
                    m._walk_offset, m._total_walks = lo, W
                    m.train(walks[lo:hi], epochs=E, epoch_range=(ep, ep + 1))
                for k in (0, 1):
                    mean = torch.stack([t[k] for t in tabs]).mean(dim=0)
                    for t in tabs:
                        t[k].copy_(mean)
            vals.append(wf.link_auc(tabs[0][0], pos, neg))
        auc[G] = float(np.mean(vals))
    print("data-parallel AUC by G:", auc)
    assert auc[1] > 0.75
    for G in (2, 4, 8):
        assert auc[G] >= auc[1] - 0.01 and abs(auc[G] - auc[1]) <= 0.035, auc


@pytest.mark.parametrize("dim,K", [(128, 5), (100, 5), (256, 7), (32, 1), (64, 32)])
def test_sgns_latency_hiding_modes_are_the_same_computation(n2v, monkeypatch, dim, K):
    """sgns_kernel MODE 1-4 (negatives drawn up front + L2 prefetch of their rows; lane-parallel draws
    from the jumped-ahead PCG stream; double-buffered target rows; all K = 5 distinct target rows in
    flight at once; the next pair drawn one pair ahead) make the same draws and the same arithmetic as MODE 0: in the deterministic single-warp trace mode the pair trace and both tables
    are bit-identical, and the multi-warp kernels agree on the pair / token counts.  The corpus has
    a 50-token head (repeated negatives inside a pair: the double-buffer hazard) and a 90k-id tail
    (two-level negative table: 4 draws per negative)."""
    torch = n2v.torch
    from node2vec_b200.sgns import Word2Vec
    gen = torch.Generator(device="cuda"); gen.manual_seed(dim)
    walks = torch.randint(0, 90000, (300, 25), device="cuda", dtype=torch.int32, generator=gen)
    walks[:, ::3] = torch.randint(0, 50, (300, 9), device="cuda", dtype=torch.int32, generator=gen)
    small = torch.randint(0, 12, (200, 25), device="cuda", dtype=torch.int32, generator=gen)       # single-level table, many repeats
    out = {}
    for mode in ("0", "1", "2", "3", "4", "5"):
        monkeypatch.setenv("N2V_SGNS_MODE", mode)
        res = []
        for corpus in (walks, small):
            m = Word2Vec(size=dim, sg=1, negative=K, window=5, min_count=1, iter=2, seed=5, sample=1e-3)
            m.build_vocab(corpus)
            m.train(corpus, trace_cap=400000)
            trace, alphas = m.last_trace
            full = Word2Vec(size=dim, sg=1, negative=K, window=5, min_count=1, iter=1, seed=5, sample=1e-3)
            full.build_vocab(corpus)
            full.train(corpus)
            res.append((m.syn0.clone(), m.syn1neg.clone(), trace.copy(), alphas.copy(), dict(m.train_stats),
                        (full.train_stats["pairs"], full.train_stats["tokens_kept"]), bool(torch.isfinite(full.syn0).all())))
        out[mode] = res
    for mode in ("1", "2", "3", "4", "5"):
        for a, b in zip(out["0"], out[mode]):
            assert a[4] == b[4] and a[4]["pairs"] > 1000, mode
            assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), mode
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), mode
            assert a[5] == b[5] and a[6] and b[6], mode


def test_transition_categories_at_a_giant_hub(n2v):
    """BASELINE configs[2] graph (reference order trim -> mirror: hubs of degree > 100000): walkers that
    arrive at the largest hub v from a low-degree neighbour t choose x == t / x in N(t) / x elsewhere
    with the reference's masses 1/p : |N(t) & N(v)| : (deg v - 1 - |N(t) & N(v)|)/q
    (randomwalk.py:223-230) -- the mixture sampler's fp32 component thresholds at deg(v) ~ 2e5."""
    torch = n2v.torch
    src, dst = n2v.synth.rmat_hotspot_edges_device(20, 16, seed=42, hotspots=16, hotspot_degree=1 << 18)
    (src, dst), _ = n2v.fugue.trim_index(None, (src, dst), indexed=True, directed=False, max_out_deg=10000, random_seed=1)
    g = n2v.graph.DeviceGraph.from_arcs(src, dst, None, n_vertices=1 << 20)
    assert g.flags & 7 == 7
    deg = g.degrees().long()
    v = int(torch.argmax(deg))
    assert int(deg[v]) > 100000
    base = g.vtx[:, 0].long()
    nv = g.col[int(base[v]): int(base[v]) + int(deg[v])].long()
    # t: a neighbour of v with a small degree that shares at least one neighbour with v
    cand = nv[(deg[nv] >= 3) & (deg[nv] <= 12)][:200]
    best = None
    for t in cand.tolist():
        nt = g.col[int(base[t]): int(base[t]) + int(deg[t])].long()
        c = int(torch.isin(nt, nv).sum())
        if c >= 1:
            best = (t, nt, c)
            break
    assert best is not None
    t, nt, c = best
    for p, q in ((0.25, 4.0), (4.0, 2.0), (1.0, 8.0)):
        ws, al, _ = g.walk(torch.tensor([t], dtype=torch.int32, device="cuda"), 3_000_000, 2, p, q, seed=77)
        x = ws[al & (ws[:, 1] == v)][:, 2].long()
        n = int(x.numel())
        assert n > 100000
        got = np.array([int((x == t).sum()), int((torch.isin(x, nt) & (x != t)).sum()), 0], dtype=np.float64)
        got[2] = n - got[0] - got[1]
        mass = np.array([1.0 / p, float(c), (int(deg[v]) - 1 - c) / q])
        ok, pval = chi_square_ok(got, mass / mass.sum(), ALPHA)
        assert ok, (p, q, t, v, int(deg[v]), c, got.tolist(), (mass / mass.sum() * n).tolist(), pval)


def test_first_occurrence_on_string_names(n2v):
    """n2v_first_occurrence_bytes: first occurrence of UTF-8 byte strings in Arrow layout (empty names,
    shared prefixes, multi-byte characters, one very long name), and the rank codes built from it."""
    import pandas as pd
    import pyarrow as pa
    torch = n2v.torch
    from node2vec_b200 import preprocess
    rng = np.random.default_rng(4)
    pool = ["", "a", "aa", "aaa", "b", "ab", "ba", "é", "éa", "日本", "user_" + "x" * 300] + [f"v{i}" for i in range(400)]
    names = [pool[i] for i in rng.integers(0, len(pool), 30000)]
    arr = pa.array(names, type=pa.large_string())
    off = np.frombuffer(arr.buffers()[1], dtype=np.int64)[: len(arr) + 1].copy()
    dat = np.frombuffer(arr.buffers()[2], dtype=np.uint8).copy()
    got = preprocess.first_occurrence_bytes(torch.as_tensor(off, device="cuda"), torch.as_tensor(dat, device="cuda")).cpu().numpy()
    seen, want = {}, np.empty(len(names), dtype=np.int64)
    for i, s in enumerate(names):
        want[i] = seen.setdefault(s, i)
    assert np.array_equal(got, want)
    src, dst = pd.Series(names[:15000]), pd.Series(names[15000:])
    s_code, d_code, uniques = preprocess.string_name_ranks(src, dst, torch.device("cuda"))
    assert uniques.tolist() == sorted(set(names))
    assert [uniques[c] for c in s_code.cpu().tolist()] == names[:15000] and [uniques[c] for c in d_code.cpu().tolist()] == names[15000:]
    assert preprocess.string_name_ranks(pd.Series(["a", 1], dtype=object), pd.Series(["b", "c"]), torch.device("cuda")) is None
