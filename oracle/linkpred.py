"""Link-prediction AUC protocol for the embedding gate (TEST INFRASTRUCTURE).

The reference has no evaluation code; BASELINE.json's north_star fixes the gate as
"link-prediction AUC within +-0.01 of the gensim reference on the same walks".  Protocol
(SURVEY 8c): hold out a seeded 10 % of the undirected edges while keeping every vertex's
degree >= 1, walk + train on the rest, score a pair by the dot product of its two
embeddings, use an equal number of seeded non-edges as negatives, sklearn roc_auc_score.
"""
import numpy as np


def split_edges(edges: np.ndarray, n: int, frac: float = 0.1, seed: int = 0):
    """edges: [m, 2] undirected, unique, a < b.  Returns (train_edges, test_pos, test_neg)."""
    rng = np.random.default_rng(seed)
    m = len(edges)
    order = rng.permutation(m)
    deg = np.bincount(edges.reshape(-1), minlength=n)
    want = int(m * frac)
    held = np.zeros(m, dtype=bool)
    for i in order:
        if want == 0:
            break
        a, b = edges[i]
        if deg[a] > 1 and deg[b] > 1:
            held[i] = True
            deg[a] -= 1
            deg[b] -= 1
            want -= 1
    pos = edges[held]
    present = set(map(tuple, edges.tolist()))
    neg = []
    while len(neg) < len(pos):
        a, b = rng.integers(0, n, 2)
        if a == b:
            continue
        key = (min(a, b), max(a, b))
        if key in present:
            continue
        present.add(key)
        neg.append(key)
    return edges[~held], pos, np.asarray(neg, dtype=np.int64)


def auc_dot(emb: np.ndarray, pos: np.ndarray, neg: np.ndarray) -> float:
    from sklearn.metrics import roc_auc_score
    s_pos = np.einsum("ij,ij->i", emb[pos[:, 0]], emb[pos[:, 1]])
    s_neg = np.einsum("ij,ij->i", emb[neg[:, 0]], emb[neg[:, 1]])
    y = np.concatenate([np.ones(len(s_pos)), np.zeros(len(s_neg))])
    return float(roc_auc_score(y, np.concatenate([s_pos, s_neg])))
