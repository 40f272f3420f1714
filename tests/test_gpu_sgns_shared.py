"""EXPERIMENTAL window-shared-negatives SGNS kernel (csrc/sgns_shared.cu, Word2Vec(share_negatives=True)).
Not the parity path.  The arithmetic gate runs in the default GPU suite (deterministic, a second per case);
the AUC + speed comparison runs with  N2V_EXPERIMENTAL=1 pytest -m gpu -s.
Gates it must pass before it may be offered outside experiments:
  * arithmetic: the single-warp trace re-applied sequentially with gensim's per-pair arithmetic (oracle)
    reproduces the tables to 2e-5 -- the register-resident rows keep gensim's in-sentence order;
  * quality: link-prediction AUC within +-0.01 of the per-pair-negatives kernel on the same walks;
  * speed: reported, both kernels on the same walk matrix.
"""
import os

import numpy as np
import pytest

from oracle import clib

pytestmark = pytest.mark.gpu
experimental = pytest.mark.skipif(os.environ.get("N2V_EXPERIMENTAL") != "1",
                                  reason="AUC / speed comparison of the experimental kernel: set N2V_EXPERIMENTAL=1")


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    from node2vec_b200 import _lib, graph, sgns, workflows
    _lib.load()

    class NS:
        pass
    ns = NS()
    ns.torch, ns.sgns, ns.graph, ns.wf = torch, sgns, graph, workflows
    return ns


@pytest.mark.parametrize("dim", [128, 32, 100])
def test_shared_kernel_arithmetic_is_sequential_gensim_arithmetic(env, dim):
    rng = np.random.default_rng(3)
    p = 1.0 / np.arange(1, 31) ** 0.8
    walks = rng.choice(30, size=(40, 21), p=p / p.sum()).astype(np.int32)
    m = env.sgns.Word2Vec(size=dim, window=5, min_count=1, sg=1, negative=5, iter=2, seed=9, sample=1e-2,
                          alpha=0.05, batch_words=100, share_negatives=True)
    m.build_vocab(walks)
    syn0, syn1 = m.syn0.cpu().numpy().copy(), m.syn1neg.cpu().numpy().copy()
    m.train(walks, epochs=1, trace_cap=100000)
    trace, alphas = m.last_trace
    n = m.train_stats["pairs"]
    assert 0 < n < 100000 and (trace[:n, :2] >= 0).all() and (trace[n:] == -2).all()
    # the K-set is constant over the pairs of one centre and has no duplicates
    same_centre = (trace[1:n, 0] == trace[:n - 1, 0])
    assert (trace[1:n, 2:][same_centre] == trace[:n - 1, 2:][same_centre]).mean() > 0.9
    for row in trace[:n, 2:]:
        live = row[row >= 0]
        assert len(set(live.tolist())) == len(live)
    clib.sgns_apply_trace(trace, alphas, n, 5, syn0, syn1)
    np.testing.assert_allclose(m.syn0.cpu().numpy(), syn0, atol=2e-5, rtol=0)
    np.testing.assert_allclose(m.syn1neg.cpu().numpy(), syn1, atol=2e-5, rtol=0)


@experimental
def test_shared_kernel_auc_and_speed(env):
    torch, wf = env.torch, env.wf
    rng = np.random.default_rng(7)
    n, blocks = 3000, 20
    iu, ju = np.triu_indices(n, 1)
    same = (iu // (n // blocks)) == (ju // (n // blocks))
    keep = rng.random(len(iu)) < np.where(same, 0.06, 0.001)
    src, dst = torch.as_tensor(iu[keep]).cuda(), torch.as_tensor(ju[keep]).cuda()
    ta, tb, pos, neg = wf.split_edges(src, dst, n, 0.1, seed=0)
    g = env.graph.DeviceGraph.from_arcs(torch.cat([ta, tb]).int(), torch.cat([tb, ta]).int(), None, n_vertices=n)
    walks, alive, _ = g.walk(g.start_vertices(), 10, 40, 1.0, 1.0, seed=5)
    aucs = {False: [], True: []}
    for share in (False, True):
        for seed in (1, 2, 3):
            m = env.sgns.Word2Vec(walks, size=128, sg=1, negative=5, window=5, min_count=1, iter=5, seed=seed,
                                  batch_words=10000, share_negatives=share)
            aucs[share].append(wf.link_auc(m.syn0, pos, neg))
    # speed on a config-2-sized matrix (800k walks x 41 tokens, 3000-word vocabulary => L2-resident tables)
    big, _, _ = g.walk(g.start_vertices(), 270, 40, 1.0, 1.0, seed=6)
    ms = {}
    for share in (False, True):
        m = env.sgns.Word2Vec(size=128, sg=1, negative=5, window=5, min_count=1, iter=1, seed=1, batch_words=10000,
                              share_negatives=share)
        m.build_vocab(big)
        m.train(big, epochs=1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.train(big, epochs=1)
        e1.record()
        torch.cuda.synchronize()
        ms[share] = (e0.elapsed_time(e1), m.train_stats["pairs"])
    print("\nAUC per-pair negatives", [round(a, 4) for a in aucs[False]], "shared", [round(a, 4) for a in aucs[True]])
    print("epoch ms per-pair %.2f (%.3e pairs/s)   shared %.2f (%.3e pairs/s)   walks %s" % (
        ms[False][0], ms[False][1] / ms[False][0] * 1e3, ms[True][0], ms[True][1] / ms[True][0] * 1e3, tuple(big.shape)))
    with open(os.path.join("gpurun_out", "sgns_shared.txt") if os.path.isdir("gpurun_out") else os.devnull, "w") as f:
        f.write(repr({"auc": aucs, "ms_pairs": ms}) + "\n")
    assert abs(np.mean(aucs[True]) - np.mean(aucs[False])) <= 0.01, aucs
