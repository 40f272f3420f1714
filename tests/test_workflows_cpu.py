"""Host-side logic of node2vec_b200.workflows that needs no GPU: the AUC estimator, the edge
split and the staleness test (CPU tensors)."""
import numpy as np
import pytest
import torch

from node2vec_b200 import workflows as wf


def test_auc_matches_sklearn_with_ties():
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(3)
    for n_pos, n_neg, levels in ((50, 70, None), (200, 100, 7), (1, 1, None), (30, 30, 2)):
        pos = rng.normal(0.5, 1.0, n_pos)
        neg = rng.normal(0.0, 1.0, n_neg)
        if levels:                                   # quantise: many exact ties
            pos, neg = np.round(pos * levels) / levels, np.round(neg * levels) / levels
        want = roc_auc_score(np.r_[np.ones(n_pos), np.zeros(n_neg)], np.r_[pos, neg])
        got = wf.auc_from_scores(torch.as_tensor(pos), torch.as_tensor(neg))
        assert got == pytest.approx(want, abs=1e-12)
    assert wf.auc_from_scores(torch.tensor([2.0, 3.0]), torch.tensor([0.0, 1.0])) == 1.0
    assert wf.auc_from_scores(torch.tensor([1.0]), torch.tensor([1.0])) == 0.5
    with pytest.raises(ValueError):
        wf.auc_from_scores(torch.tensor([]), torch.tensor([1.0]))


def test_split_edges_keeps_every_vertex_connected():
    rng = np.random.default_rng(5)
    n = 400
    a = rng.integers(0, n, 3000)
    b = rng.integers(0, n, 3000)
    a = np.r_[a, np.arange(n)]                       # every vertex has at least one edge
    b = np.r_[b, (np.arange(n) + 1) % n]
    ta, tb, pos, neg = wf.split_edges(torch.as_tensor(a), torch.as_tensor(b), n, holdout=0.1, seed=11)
    edges = {(min(x, y), max(x, y)) for x, y in zip(a.tolist(), b.tolist()) if x != y}
    train = set(zip(ta.tolist(), tb.tolist()))
    held = set(map(tuple, pos.tolist()))
    assert train | held == edges and not (train & held)
    assert len(held) == int(len(edges) * 0.1) == len(neg)
    deg = np.bincount(np.r_[ta.numpy(), tb.numpy()], minlength=n)
    assert deg.min() >= 1                             # nobody lost all training edges
    negs = set(map(tuple, neg.tolist()))
    assert len(negs) == len(neg) and not (negs & edges) and all(x < y for x, y in negs)
    again = wf.split_edges(torch.as_tensor(a), torch.as_tensor(b), n, holdout=0.1, seed=11)
    assert torch.equal(again[2], pos) and torch.equal(again[3], neg)
    other = wf.split_edges(torch.as_tensor(a), torch.as_tensor(b), n, holdout=0.1, seed=12)
    assert not torch.equal(other[2], pos)


def test_stale_start_vertices():
    walks = torch.tensor([[0, 1, 2], [0, 3, 4], [5, 6, 7], [8, 9, 1], [8, 8, 8]], dtype=torch.int32)
    assert wf.stale_start_vertices(walks, torch.tensor([1])).tolist() == [0, 8]
    assert wf.stale_start_vertices(walks, torch.tensor([7, 5])).tolist() == [5]
    assert wf.stale_start_vertices(walks, torch.tensor([42])).tolist() == []
    assert wf.stale_start_vertices(walks[:0], torch.tensor([1])).tolist() == []


def test_parameter_validation_matches_reference():
    with pytest.raises(ValueError):
        wf._merged({"return_param": 0})
    with pytest.raises(ValueError):
        wf._merged({"walk_length": 0})
    assert wf._merged({"num_walks": 3})["walk_length"] == 20        # constants.py:9-18 defaults
