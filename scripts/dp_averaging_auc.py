#!/usr/bin/env python
"""AUC impact of data-parallel SGNS with model averaging ("S and its AUC impact are a
tuned, reported parameter"), on ONE GPU: G replicas that start from the same tables, each trained on
its contiguous shard of the walk matrix, averaged every S epochs -- the arithmetic of
Word2Vec(process_group=...) without needing G devices (replicas are independent between averages,
so running them one after the other is the same computation).

    python scripts/dp_averaging_auc.py            # prints a table; run under gpurun
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from node2vec_b200 import workflows as wf
from node2vec_b200.graph import DeviceGraph
from node2vec_b200.sgns import Word2Vec

EPOCHS, DIM = 5, 64
rng = np.random.default_rng(7)
n, blocks = 3000, 20
iu, ju = np.triu_indices(n, 1)
same = (iu // (n // blocks)) == (ju // (n // blocks))
keep = rng.random(len(iu)) < np.where(same, 0.06, 0.001)
src, dst = torch.as_tensor(iu[keep]).cuda(), torch.as_tensor(ju[keep]).cuda()
ta, tb, pos, neg = wf.split_edges(src, dst, n, 0.1, seed=0)
g = DeviceGraph.from_arcs(torch.cat([ta, tb]).int(), torch.cat([tb, ta]).int(), None, n_vertices=n)
walks, alive, _ = g.walk(g.start_vertices(), 10, 40, 1.0, 1.0, seed=5)
assert bool(alive.all())
W = int(walks.shape[0])
print(f"SBM {n} vertices / {int(keep.sum())} edges, {W} walks x 41, dim {DIM}, {EPOCHS} epochs, window 5, 5 negatives")
print(f"{'G':>3} {'avg every':>10} {'AUC (3 seeds)':>28} {'mean':>8}")
for G in (1, 2, 4, 8):
    for S in ((1,) if G == 1 else (1, EPOCHS)):
        aucs = []
        for seed in (1, 2, 3):
            m = Word2Vec(size=DIM, sg=1, negative=5, window=5, min_count=1, iter=EPOCHS, seed=seed, batch_words=10000)
            m.build_vocab(walks)
            tabs = [(m.syn0.clone(), m.syn1neg.clone()) for _ in range(G)]
            bounds = [(W * r // G, W * (r + 1) // G) for r in range(G)]
            for ep in range(EPOCHS):
                for r, (lo, hi) in enumerate(bounds):
                    m.syn0, m.syn1neg = tabs[r]
                    m._walk_offset, m._total_walks = lo, W
                    m.train(walks[lo:hi], epochs=EPOCHS, epoch_range=(ep, ep + 1))
                if G > 1 and ((ep + 1) % S == 0 or ep + 1 == EPOCHS):
                    for k in (0, 1):
                        mean = torch.stack([t[k] for t in tabs]).mean(dim=0)
                        for t in tabs:
                            t[k].copy_(mean)
            aucs.append(wf.link_auc(tabs[0][0], pos, neg))
        print(f"{G:>3} {S:>10} {str([round(a, 4) for a in aucs]):>28} {np.mean(aucs):8.4f}", flush=True)
