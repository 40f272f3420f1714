"""CPU checks of the gensim-3.8 SGNS restatement (oracle/csrc/sgns_ref.c).  gensim itself is not
installable here and the reference's tests assert no embedding numerics, so this half of the oracle
is PARITY UNPINNED.  What can be checked without gensim: the C code against
independent numpy restatements of the published formulas (EXP_TABLE, keep probabilities, the
cumulative negative table, the per-pair update), and that it learns (link-prediction AUC)."""
import numpy as np
import pytest

from oracle import clib, linkpred


def test_exp_table_is_gensims():
    """word2vec_inner.pyx: EXP_TABLE[i] = exp((i / 1000 * 2 - 1) * 6); EXP_TABLE[i] /= EXP_TABLE[i] + 1 (REAL_t)."""
    t = clib.sgns_exp_table()
    x = np.exp((np.arange(1000) / 1000.0 * 2 - 1) * 6.0).astype(np.float32)
    want = (x / (x + np.float32(1))).astype(np.float32)
    assert t.dtype == np.float32 and np.array_equal(t, want)
    assert t[0] < 0.003 and abs(t[500] - 0.5) < 1e-6 and t[999] > 0.997


def test_vocab_keep_probability_and_cum_table():
    """prepare_vocab: threshold = sample * retain_total; p = (sqrt(c / threshold) + 1) * (threshold / c), capped
    at 1, stored as int(round(p * 2**32)); make_cum_table: cumulative count**0.75 scaled to 2**31 - 1, ids by
    descending count (ties: id order)."""
    counts = np.array([0, 500, 3, 40, 4, 1200, 40, 7, 90000, 5], dtype=np.int64)
    keep, order, cum = clib.sgns_vocab(counts, min_count=5, sample=1e-3)
    kept = [i for i in range(len(counts)) if counts[i] >= 5]
    assert sorted(order.tolist()) == kept
    assert order.tolist() == sorted(kept, key=lambda i: (-counts[i], i))
    total = counts[kept].sum()
    thr = 1e-3 * total
    for i in range(len(counts)):
        if counts[i] < 5:
            assert keep[i] == 0
            continue
        p = min(1.0, (np.sqrt(counts[i] / thr) + 1) * (thr / counts[i]))
        assert int(keep[i]) == int(round(p * 2 ** 32))
    pw = counts[order].astype(np.float64) ** 0.75
    want = np.round(np.cumsum(pw) / pw.sum() * (2 ** 31 - 1)).astype(np.uint32)
    assert np.array_equal(cum, want) and cum[-1] == 2 ** 31 - 1
    assert np.all(np.diff(cum.astype(np.int64)) > 0)


def _apply_trace_numpy(trace, alphas, K, syn0, syn1, table):
    """w2v_fast_sentence_sg_neg, one pair at a time, in numpy float32."""
    f32 = np.float32
    for (wi, wj, *negs), alpha in zip(trace.tolist(), alphas.tolist()):
        row = syn0[wj]
        work = np.zeros_like(row)
        for d, tgt in enumerate([wi] + negs):
            if tgt < 0 or (d > 0 and tgt == wi):               # gensim: a negative equal to the centre is skipped
                continue
            label = f32(1.0) if d == 0 else f32(0.0)
            t = syn1[tgt]
            f = f32(0.0)
            for a, b in zip(row.tolist(), t.tolist()):      # sdot in index order, fp32
                f = f32(f + f32(f32(a) * f32(b)))
            if f <= -6.0 or f >= 6.0:
                continue
            s = table[int((f + f32(6.0)) * f32(1000 / 6.0 / 2.0))]
            g = f32((label - s) * f32(alpha))
            work = (work + g * t).astype(f32)
            syn1[tgt] = (t + g * row).astype(f32)
        syn0[wj] = (row + work).astype(f32)


def test_apply_trace_matches_numpy_per_pair_arithmetic():
    rng = np.random.default_rng(5)
    V, D, K, n = 12, 8, 5, 300
    syn0 = ((rng.random((V, D)) - 0.5)).astype(np.float32)        # large values: some |f| >= 6 clips too
    syn1 = ((rng.random((V, D)) - 0.5) * 6).astype(np.float32)
    trace = np.concatenate([rng.integers(0, V, (n, 2)), rng.integers(-1, V, (n, K))], axis=1).astype(np.int32)
    alphas = rng.uniform(0.01, 0.05, n).astype(np.float32)
    a0, a1 = syn0.copy(), syn1.copy()
    clib.sgns_apply_trace(trace, alphas, n, K, a0, a1)
    b0, b1 = syn0.copy(), syn1.copy()
    _apply_trace_numpy(trace, alphas, K, b0, b1, clib.sgns_exp_table())
    # same arithmetic up to the summation order inside the dot product / fused multiply-adds
    np.testing.assert_allclose(a0, b0, atol=2e-5, rtol=0)
    np.testing.assert_allclose(a1, b1, atol=2e-5, rtol=0)
    assert not np.allclose(a0, syn0)


def test_training_is_deterministic_single_threaded_and_learns():
    """Stochastic block model: the restatement's embeddings must predict held-out links."""
    rng = np.random.default_rng(7)
    n, blocks = 400, 8
    iu, ju = np.triu_indices(n, 1)
    same = (iu // (n // blocks)) == (ju // (n // blocks))
    keep = rng.random(len(iu)) < np.where(same, 0.2, 0.004)
    edges = np.stack([iu[keep], ju[keep]], axis=1)
    train, pos, neg = linkpred.split_edges(edges, n, 0.1, 0)
    nbr = [[] for _ in range(n)]
    for a, b in train.tolist():
        nbr[a].append(b)
        nbr[b].append(a)
    walks = np.zeros((n * 6, 21), dtype=np.int32)                   # plain uniform walks are enough here
    for w in range(len(walks)):
        v = w % n
        for k in range(21):
            walks[w, k] = v
            v = nbr[v][rng.integers(len(nbr[v]))]
    counts = np.bincount(walks.reshape(-1), minlength=n)
    out = []
    for _ in range(2):
        syn0, syn1 = clib.sgns_init(n, 32, 3)
        pairs = clib.sgns_train(walks, counts, syn0, syn1, epochs=3, min_count=1, seed=3, threads=1)
        out.append((pairs, syn0.copy()))
    assert out[0][0] == out[1][0] > 100000 and np.array_equal(out[0][1], out[1][1])
    assert linkpred.auc_dot(out[0][1], pos, neg) > 0.75
