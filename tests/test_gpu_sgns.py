"""GPU parity tests for the SGNS half (K4 vocab / tables, K3 sgns_train) through the C ABI.

The oracle here is the gensim-3.8 restatement (oracle/csrc/sgns_ref.c) -- PARITY UNPINNED
against gensim itself (not installable; see that file's header).  Gates:
  * integer work (counts, first positions, vocabulary order, keep thresholds): bit-exact
  * the negative-sampling table encodes count^0.75 exactly (reconstructed law, 1e-9)
  * arithmetic: the kernel's own sampled pairs re-applied by the oracle's per-pair gensim
    arithmetic give the same tables within 2e-5 absolute (fp32, single-warp trace mode)
  * sampling laws (window, negatives, sub-sampling): chi-square, alpha = 1e-4
  * embeddings: link-prediction AUC within +-0.01 of the restatement on the same walks
    (north_star tolerance), averaged over seeds
"""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import clib, linkpred
from tests.helpers import chi_square_ok

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    from node2vec_b200 import embedding, fugue, graph, sgns

    class NS:
        pass
    ns = NS()
    ns.torch, ns.sgns, ns.embedding, ns.graph, ns.fugue = torch, sgns, embedding, graph, fugue
    return ns


def _toy_walks(rng, n_ids=50, W=400, L=21, skew=True):
    p = 1.0 / np.arange(1, n_ids + 1) if skew else np.ones(n_ids)
    p = p / p.sum()
    return rng.choice(n_ids, size=(W, L), p=p).astype(np.int32)


def test_vocab_counts_and_order(env):
    rng = np.random.default_rng(1)
    walks = _toy_walks(rng, 60, 500, 17)
    walks[walks == 7] = 8                                    # id 7 never occurs
    m = env.sgns.Word2Vec(size=8, min_count=3, sg=1, negative=5, sample=1e-3, seed=5)
    m.build_vocab(walks)
    counts = np.bincount(walks.reshape(-1), minlength=walks.max() + 1)
    flat = walks.reshape(-1)
    first = {int(v): int(np.flatnonzero(flat == v)[0]) for v in np.unique(flat)}
    kept = [v for v in sorted(first, key=first.get) if counts[v] >= 3]
    assert list(m.wv.vocab) == [str(v) for v in kept]        # dict order = first appearance
    assert all(m.wv.vocab[str(v)].count == counts[v] for v in kept)
    by_count = sorted(kept, key=lambda v: -counts[v])        # stable: ties by first appearance
    assert m.wv.index2word == [str(v) for v in by_count]
    assert [m.wv.vocab[w].index for w in m.wv.index2word] == list(range(len(kept)))
    # keep thresholds = gensim sample_int
    keep, order, cum = clib.sgns_vocab(counts, 3, 1e-3, 0.75)
    assert [counts[v] for v in order] == [counts[v] for v in by_count]     # the oracle breaks ties by id
    assert sorted(order.tolist()) == sorted(by_count)
    for v in kept:
        assert m.wv.vocab[str(v)].sample_int == int(keep[v]), v
    assert "7" not in m.wv.vocab
    # init law: (U - 0.5) / size, zero output table
    s0 = m.syn0.cpu().numpy()
    assert np.abs(s0).max() <= 0.5 / 8 and abs(s0.mean()) < 8e-3 and s0.std() > 0.03
    assert float(m.syn1neg.abs().max()) == 0.0


def test_negative_table_encodes_unigram_power(env):
    rng = np.random.default_rng(2)
    walks = _toy_walks(rng, 300, 3000, 21)
    m = env.sgns.Word2Vec(size=8, min_count=5, sg=1, negative=5, seed=1)
    m.build_vocab(walks)
    counts = np.bincount(walks.reshape(-1), minlength=m._n_rows).astype(np.float64)
    w = np.where(counts >= 5, counts ** 0.75, 0.0)
    tab = m._neg.cpu().numpy()
    thr = tab[:, 0].view(np.uint32).astype(np.float64)
    thr = np.where(thr == 4294967295.0, 4294967296.0, thr)
    law = thr.copy()
    np.add.at(law, tab[:, 1], 4294967296.0 - thr)
    law /= law.sum()
    np.testing.assert_allclose(law, w / w.sum(), atol=1e-9)
    # and it agrees with gensim's cum_table law
    keep, order, cum = clib.sgns_vocab(counts.astype(np.int64), 5, 1e-3, 0.75)
    p_cum = np.diff(np.concatenate([[0], cum.astype(np.float64)])) / cum[-1]
    np.testing.assert_allclose(law[order], p_cum, atol=2e-9)


def _alias_law(tab):
    """Exact law encoded by an alias table of {thr u32, alias i32} entries (aliases index `tab`)."""
    thr = tab[:, 0].view(np.uint32).astype(np.float64)
    thr = np.where(thr == 4294967295.0, 4294967296.0, thr)
    law = thr.copy()
    np.add.at(law, tab[:, 1], 4294967296.0 - thr)
    return law / law.sum()


def test_two_level_negative_table_large_id_space(env):
    """n_vertices > 65536: per-chunk alias tables + a top-level table over chunk masses must encode
    count^0.75 exactly as the single table does (ids below min_count get probability 0)."""
    import ctypes as C
    from node2vec_b200 import _lib
    torch = env.torch
    lib = _lib.load()
    n = 200_000 + 37                                   # last chunk is ragged
    gen = torch.Generator(device="cuda").manual_seed(5)
    counts = (torch.rand(n, device="cuda", generator=gen).clamp_min(1e-4).pow(-0.6) * 2).long()
    counts[8000:25000] = 0                             # two whole chunks (and a bit) of absent ids
    from node2vec_b200.sgns import NEG_CHUNK, neg_top_entries
    n_top = neg_top_entries(n)
    assert n_top == (n + NEG_CHUNK - 1) // NEG_CHUNK
    keep = torch.empty(n, dtype=torch.int32, device="cuda")
    neg = torch.empty((n + n_top, 2), dtype=torch.int32, device="cuda")
    scratch = torch.empty((n + n_top) * 20 + 64, dtype=torch.uint8, device="cuda")
    tot = (C.c_int64 * 2)()
    _lib.check(lib.n2v_sgns_prepare(_lib.ptr(counts), n, 3, 1e-3, 0.75, _lib.ptr(keep), _lib.ptr(neg),
                                    _lib.ptr(scratch), tot, _lib.current_stream_ptr()))
    tab = neg.cpu().numpy()
    c = counts.cpu().numpy().astype(np.float64)
    w = np.where(c >= 3, c ** 0.75, 0.0)
    top = _alias_law(tab[n:])
    law = np.zeros(n)
    for k in range(n_top):
        lo, hi = k * NEG_CHUNK, min(n, (k + 1) * NEG_CHUNK)
        if w[lo:hi].sum() == 0:
            assert top[k] < 1e-12
            continue
        local = tab[lo:hi].copy()
        assert ((local[:, 1] >= lo) & (local[:, 1] < hi)).all()      # aliases stay inside the chunk
        local[:, 1] -= lo
        law[lo:hi] = top[k] * _alias_law(local)
    np.testing.assert_allclose(law, w / w.sum(), atol=1e-10)
    assert int(tot[1]) == int((c >= 3).sum())


@pytest.mark.parametrize("dim,atomic", [(32, True), (128, True), (128, False), (256, True), (100, True)])
def test_kernel_arithmetic_matches_gensim_per_pair(env, dim, atomic):
    """Single-warp trace mode: the pairs the kernel sampled, re-applied sequentially with
    gensim's arithmetic (EXP_TABLE sigmoid, |f| >= 6 clip, skip negatives == centre)."""
    rng = np.random.default_rng(3)
    walks = _toy_walks(rng, 40, 30, 21)
    m = env.sgns.Word2Vec(size=dim, window=5, min_count=1, sg=1, negative=5, iter=2, seed=9, sample=1e-2,
                          alpha=0.05, batch_words=100, atomic_updates=atomic)
    m.build_vocab(walks)
    syn0, syn1 = m.syn0.cpu().numpy().copy(), m.syn1neg.cpu().numpy().copy()
    m.train(walks, epochs=1, trace_cap=100000)
    trace, alphas = m.last_trace
    n = m.train_stats["pairs"]
    assert 0 < n < 100000 and (trace[:n, :2] >= 0).all() and (trace[n:] == -2).all()
    assert alphas[0] == np.float32(0.05) and alphas[n - 1] < alphas[0]
    clib.sgns_apply_trace(trace, alphas, n, 5, syn0, syn1)
    np.testing.assert_allclose(m.syn0.cpu().numpy(), syn0, atol=2e-5, rtol=0)
    np.testing.assert_allclose(m.syn1neg.cpu().numpy(), syn1, atol=2e-5, rtol=0)
    assert m.train_stats["negatives_skipped"] == int((trace[:n, 2:] == -1).sum())


def test_sampling_laws(env):
    rng = np.random.default_rng(4)
    n_ids, W, L = 80, 600, 31
    walks = _toy_walks(rng, n_ids, W, L)
    m = env.sgns.Word2Vec(size=8, window=5, min_count=1, sg=1, negative=5, iter=1, seed=21, sample=2e-3)
    m.build_vocab(walks)
    m.train(walks, epochs=1, trace_cap=400000)
    trace, _ = m.last_trace
    n = m.train_stats["pairs"]
    assert n < 400000
    trace = trace[:n]
    counts = np.bincount(walks.reshape(-1), minlength=n_ids).astype(np.float64)
    # negatives ~ count^0.75 (before the "== centre" skip: include only rows whose centre differs)
    negs = trace[:, 2:].reshape(-1)
    negs = negs[negs >= 0]
    law = counts ** 0.75
    centre_w = np.bincount(trace[:, 0], minlength=n_ids) * 5.0
    expect = law / law.sum() * len(trace) * 5 - centre_w * (law / law.sum())   # minus skipped mass
    expect = np.maximum(expect, 1e-9)
    ok, pval = chi_square_ok(np.bincount(negs, minlength=n_ids), expect / expect.sum())
    assert ok, pval
    # sub-sampling: kept fraction per id ~ keep probability
    keep, _, _ = clib.sgns_vocab(counts.astype(np.int64), 1, 2e-3, 0.75)
    p_keep = np.minimum(keep.astype(np.float64) / 2 ** 32, 1.0)
    kept = m.train_stats["tokens_kept"]
    assert abs(kept - (counts * p_keep).sum()) < 5 * np.sqrt((counts * p_keep * (1 - p_keep)).sum() + 1)
    assert kept < W * L                                        # something was sub-sampled
    # reduced window: the number of pairs per kept token averages window + 1 (interior tokens)
    m2 = env.sgns.Word2Vec(size=8, window=5, min_count=1, sg=1, negative=1, iter=1, seed=3, sample=0.0)
    long_walks = _toy_walks(rng, n_ids, 200, 201)
    m2.build_vocab(long_walks)
    m2.train(long_walks, epochs=1)
    per_token = m2.train_stats["pairs"] / m2.train_stats["tokens_kept"]
    assert abs(per_token - 6.0) < 0.2                          # 2 * E[1..5] = 6, minus edge effects


def test_reference_embedding_tests_against_facade(env, tmp_path):
    """tests/test_embedding.py:17-84 of the reference, against node2vec_b200.embedding."""
    E = env.embedding
    base = E.Node2VecBase()
    for fn, args in [(base.fit, ()), (base.embedding, ()), (base.get_vector, (0,)),
                     (base.save_model, ("file:///a", "b")), (base.load_model, ("file:///a", "b"))]:
        with pytest.raises(NotImplementedError):
            fn(*args)
    df = pd.DataFrame.from_dict({"walk": [[0, 1, 1, 0, 3, 4], [1, 2, 3, 2, 0, 4], [2, 3, 1, 0, 4, 4]]})
    assert isinstance(E.Node2VecGensim(df, {}), E.Node2VecGensim)
    params = {"iter": 3}
    n2v = E.Node2VecGensim(df, w2v_params=params, window_size=6, vector_size=64, random_seed=1000)
    assert params["window"] == 6 and params["size"] == 64 and params["seed"] == 1000 and params["min_count"] == 10
    pytest.raises(ValueError, E.Node2VecGensim, df, {}, window_size=3)
    pytest.raises(ValueError, E.Node2VecGensim, df, {}, vector_size=16)
    w2v_params = {"min_count": 0, "iter": 1, "seed": 1000, "batch_words": 1, "size": 4, "workers": 4}
    n2v = E.Node2VecGensim(df, w2v_params=w2v_params)
    model = n2v.fit()
    assert isinstance(model, env.sgns.Word2Vec)
    res = n2v.embedding()
    assert isinstance(res, pd.DataFrame) and len(res) > 0 and list(res.columns) == ["id", "vector"]
    assert res["id"].tolist() == [0, 1, 3, 4, 2]              # first-appearance order, like gensim's dict
    assert len(n2v.get_vector(vertex_id="0")) > 0 and len(n2v.get_vector(vertex_id=1)) == 4
    d = str(tmp_path)
    n2v.save_model(d, "tmp")
    assert os.path.exists(os.path.join(d, "tmp.model"))
    assert isinstance(n2v.load_model(d, "tmp"), env.sgns.Word2Vec)
    n2v.save_vectors(d, "tmp_vec")
    kv = n2v.load_vectors(d, "tmp_vec")
    assert isinstance(kv, env.sgns.KeyedVectors)
    np.testing.assert_allclose(kv["3"], n2v.get_vector(3), rtol=1e-6)
    name_id = pd.DataFrame.from_dict({"name": ["a", "b", "c", "d", "e"], "id": [0, 1, 2, 3, 4]})
    n2v = E.Node2VecGensim(df, w2v_params, name_id=name_id)
    with pytest.raises(ValueError):
        n2v.embedding()
    n2v.fit()
    res = n2v.embedding()
    assert list(res.columns) == ["name", "vector"] and len(res) == 5
    with pytest.raises(NotImplementedError):
        E.Node2VecSpark(df, {})
    with pytest.raises(NotImplementedError):
        env.sgns.Word2Vec(size=8, sg=0, negative=5)


def _sbm(n_blocks, size, p_in, p_out, seed):
    rng = np.random.default_rng(seed)
    n = n_blocks * size
    blk = np.arange(n) // size
    iu = np.triu_indices(n, 1)
    prob = np.where(blk[iu[0]] == blk[iu[1]], p_in, p_out)
    keep = rng.random(len(prob)) < prob
    return n, np.stack([iu[0][keep], iu[1][keep]], axis=1).astype(np.int64)


def test_link_prediction_auc_matches_gensim_restatement(env):
    """north_star gate: AUC of the device embeddings within +-0.01 of the gensim-3.8
    restatement trained on the SAME walk matrix with the same hyper-parameters."""
    n, edges = _sbm(10, 150, 0.08, 0.002, 7)
    train, pos, neg = linkpred.split_edges(edges, n, 0.1, 0)
    src = np.concatenate([train[:, 0], train[:, 1]])
    dst = np.concatenate([train[:, 1], train[:, 0]])
    res = env.fugue.random_walk(None, (src, dst), {"num_walks": 10, "walk_length": 40, "return_param": 1.0,
                                                   "inout_param": 1.0}, random_seed=5)
    walks = res.walks
    counts = np.bincount(walks.reshape(-1), minlength=n)
    hp = dict(window=5, negative=5, alpha=0.025, min_alpha=1e-4, min_count=1, sample=1e-3)
    threads = max(1, min(16, os.cpu_count() or 1))
    auc_ref, auc_gpu, auc_plain = [], [], []
    for seed in (1, 2, 3):
        syn0, syn1 = clib.sgns_init(n, 64, seed)
        clib.sgns_train(walks, counts, syn0, syn1, epochs=5, seed=seed, batch_words=10000, threads=threads, **hp)
        auc_ref.append(linkpred.auc_dot(syn0, pos, neg))
        for atomic, out in ((True, auc_gpu), (False, auc_plain)):
            m = env.sgns.Word2Vec(size=64, sg=1, iter=5, seed=seed, batch_words=10000, atomic_updates=atomic, **hp)
            m.build_vocab(res.walks_device)
            m.train(res.walks_device)
            emb = np.zeros((n, 64), dtype=np.float32)
            emb[[int(t) for t in m.wv.index2word]] = m.wv.vectors
            out.append(linkpred.auc_dot(emb, pos, neg))
    print("AUC restatement", auc_ref, "gpu(atomic)", auc_gpu, "gpu(plain)", auc_plain)
    assert np.mean(auc_ref) > 0.7
    assert abs(np.mean(auc_gpu) - np.mean(auc_ref)) <= 0.01, (auc_gpu, auc_ref)
    # plain (non-atomic) Hogwild stores lose updates under ~10^4 concurrent warps on a 1.5k-row
    # table: reported, not gated -- red.global.add is the production path (DESIGN.md "SGNS").
    assert np.mean(auc_plain) > 0.5


def test_epoch_range_resumes_the_schedule(env, tmp_path):
    """train(epochs=4) == train(epoch_range=(0, 2)) + save/load + train(epoch_range=(2, 4)), bit for
    bit in the deterministic single-warp trace mode: the schedule and the random streams depend on
    the epoch index only."""
    torch = env.torch
    rng = np.random.default_rng(3)
    walks = torch.as_tensor(rng.integers(0, 60, (40, 12)).astype(np.int32)).cuda()

    def fresh():
        m = env.sgns.Word2Vec(size=32, window=3, min_count=1, sg=1, negative=3, iter=4, seed=11, sample=0.0)
        m.build_vocab(walks)
        return m
    a = fresh()
    a.train(walks, epochs=4, trace_cap=8)
    again = fresh()                                       # everything is reproducible run to run:
    assert torch.equal(again._neg, a._neg) and torch.equal(again._keep, a._keep)   # the sampling tables
    again.train(walks, epochs=4, trace_cap=8)
    assert torch.equal(again.syn0, a.syn0) and torch.equal(again.syn1neg, a.syn1neg)   # and the single-warp fit
    big = env.sgns.Word2Vec(size=8, min_count=1, sg=1, negative=5, seed=1)
    big2 = env.sgns.Word2Vec(size=8, min_count=1, sg=1, negative=5, seed=1)
    wide = torch.as_tensor(rng.integers(0, 100000, (3000, 30)).astype(np.int32)).cuda()   # two-level negative table
    big.build_vocab(wide)
    big2.build_vocab(wide)
    assert torch.equal(big._neg, big2._neg)
    b = fresh()
    b.train(walks, epochs=4, trace_cap=8, epoch_range=(0, 2))
    b.save(str(tmp_path / "half.model"))
    c = env.sgns.Word2Vec.load(str(tmp_path / "half.model"))
    c.train(walks, epochs=4, trace_cap=8, epoch_range=(2, 4))
    assert torch.equal(a.syn0, c.syn0) and torch.equal(a.syn1neg, c.syn1neg)
    assert not torch.equal(a.syn0, b.syn0)
    with pytest.raises(ValueError):
        c.train(walks, epochs=4, epoch_range=(3, 5))


def test_example_pipeline_end_to_end(env, tmp_path):
    """examples/pipeline.py: the reference example's three stages, parquet in / parquet out."""
    import importlib.util
    import pandas as pd
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("n2v_example_pipeline", os.path.join(root, "examples", "pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(1)
    names = np.array([f"user{i}" for i in range(300)])
    a, b = rng.integers(0, 300, 3000), rng.integers(0, 300, 3000)
    pd.DataFrame({"src": names[a], "dst": names[b], "weight": rng.uniform(0.5, 2.0, 3000)}).to_parquet(
        tmp_path / "input_graph.parquet")
    mod.stage_index(str(tmp_path))
    mod.stage_walk(str(tmp_path), random_seed=3)
    mod.stage_embed(str(tmp_path), random_seed=3)
    walks = pd.read_parquet(tmp_path / "graph_walks.parquet")
    assert list(walks.columns) == ["src", "walk"] and all(len(w) == 11 for w in walks["walk"][:50])
    emb = pd.read_parquet(tmp_path / "graph_embedding.parquet")
    assert list(emb.columns) == ["name", "vector"] and len(emb) > 250 and len(emb["vector"][0]) == 128
    assert set(emb["name"]) <= set(names)
    assert os.path.exists(tmp_path / "graph_vectors.txt")
