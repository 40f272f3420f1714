"""Kernel-tuning helper (run under gpurun): time one SGNS epoch on the bench workload under
different update modes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from node2vec_b200 import synth
from node2vec_b200.graph import DeviceGraph
from node2vec_b200.sgns import Word2Vec

src, dst = synth.blogcatalog_like(seed=42)
g = DeviceGraph.from_arcs(src, dst, None, n_vertices=10000)
walks, alive, _ = g.walk(g.start_vertices(), 80, 40, 0.25, 4.0, seed=42)
dim = int(os.environ.get("DIM", "128"))
for atomic in (True, False):
    m = Word2Vec(size=dim, sg=1, negative=5, min_count=1, iter=1, seed=1, atomic_updates=atomic)
    m.build_vocab(walks)
    for _ in range(2):
        m.train(walks, epochs=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        m.train(walks, epochs=1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"dim {dim} atomic={atomic}: {dt*1e3:.1f} ms/epoch, {m.train_stats['pairs']/dt:.3e} pairs/s")
