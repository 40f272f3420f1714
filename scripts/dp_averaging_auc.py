#!/usr/bin/env python
"""AUC impact of data-parallel SGNS with model averaging ("S and its AUC impact are a
tuned, reported parameter"), on ONE GPU: G replicas that start from the same tables, each trained on
its contiguous shard of the walk matrix, combined every S epochs -- the arithmetic of
Word2Vec(process_group=...) without needing G devices (replicas are independent between syncs,
so running them one after the other is the same computation).

    python scripts/dp_averaging_auc.py --graph blog            # prints a table; run under gpurun

--graph  sbm | er10k | blog   (BASELINE configs[0] / configs[1] shapes for the last two)
Combine rules compared: `avg` (the product: allreduce-mean of the tables) and `sum` (base + sum of
the replicas' deltas = mean extrapolated by G, i.e. the total update one GPU would have applied).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from node2vec_b200 import synth, workflows as wf
from node2vec_b200.graph import DeviceGraph
from node2vec_b200.sgns import Word2Vec


def graph(name):
    if name == "sbm":
        rng = np.random.default_rng(7)
        n, blocks = 3000, 20
        iu, ju = np.triu_indices(n, 1)
        same = (iu // (n // blocks)) == (ju // (n // blocks))
        keep = rng.random(len(iu)) < np.where(same, 0.06, 0.001)
        return n, iu[keep], ju[keep], dict(p=1.0, q=1.0, num_walks=10, walk_length=40, dim=64)
    if name == "er10k":
        s, d = synth.erdos_renyi(10000, 100000, seed=42)
        h = len(s) // 2
        return 10000, s[:h], d[:h], dict(p=1.0, q=0.5, num_walks=10, walk_length=20, dim=128)
    s, d = synth.blogcatalog_like(10000, 334000, seed=42)
    h = len(s) // 2
    return 10000, s[:h], d[:h], dict(p=0.25, q=4.0, num_walks=10, walk_length=40, dim=128)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", default="blog", choices=["sbm", "er10k", "blog"])
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--gs", default="1,2,4,8")
    args = ap.parse_args()
    n, a, b, c = graph(args.graph)
    E, DIM = args.epochs, c["dim"]
    src, dst = torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda()
    ta, tb, pos, neg = wf.split_edges(src, dst, n, 0.1, seed=0)
    g = DeviceGraph.from_arcs(torch.cat([ta, tb]).int(), torch.cat([tb, ta]).int(), None, n_vertices=n)
    walks, alive, _ = g.walk(g.start_vertices(), c["num_walks"], c["walk_length"], c["p"], c["q"], seed=5)
    assert bool(alive.all())
    W = int(walks.shape[0])
    print(f"{args.graph}: {n} vertices / {len(a)} edges, {W} walks x {walks.shape[1]}, dim {DIM}, {E} epochs, "
          f"window 5, 5 negatives, p={c['p']} q={c['q']}")
    print(f"{'G':>3} {'rule':>5} {'sync every':>10} {'AUC (3 seeds)':>28} {'mean':>8}")
    for G in [int(x) for x in args.gs.split(",")]:
        for rule in (("avg",) if G == 1 else ("avg", "sum")):
            for S in ((1,) if G == 1 else (1, E)):
                aucs = []
                for seed in (1, 2, 3):
                    m = Word2Vec(size=DIM, sg=1, negative=5, window=5, min_count=1, iter=E, seed=seed, batch_words=10000)
                    m.build_vocab(walks)
                    base = (m.syn0.clone(), m.syn1neg.clone())
                    tabs = [(m.syn0.clone(), m.syn1neg.clone()) for _ in range(G)]
                    bounds = [(W * r // G, W * (r + 1) // G) for r in range(G)]
                    for ep in range(E):
                        for r, (lo, hi) in enumerate(bounds):
                            m.syn0, m.syn1neg = tabs[r]
                            m._walk_offset, m._total_walks = lo, W
                            m.train(walks[lo:hi], epochs=E, epoch_range=(ep, ep + 1))
                        if G > 1 and ((ep + 1) % S == 0 or ep + 1 == E):
                            for k in (0, 1):
                                mean = torch.stack([t[k] for t in tabs]).mean(dim=0)
                                if rule == "sum":
                                    mean = base[k] + G * (mean - base[k])
                                base[k].copy_(mean)
                                for t in tabs:
                                    t[k].copy_(mean)
                    aucs.append(wf.link_auc(tabs[0][0], pos, neg))
                print(f"{G:>3} {rule:>5} {S:>10} {str([round(x, 4) for x in aucs]):>28} {np.mean(aucs):8.4f}", flush=True)


if __name__ == "__main__":
    main()
