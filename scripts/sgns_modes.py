#!/usr/bin/env python
"""Kernel-tuning helper (run under gpurun): one SGNS epoch per latency-hiding mode (N2V_SGNS_MODE=0..3,
csrc/sgns.cu) on the walks of BASELINE configs[2] (tables 2 x 0.54 GB, hub rows L2-resident) and on the
walks of an R-MAT scale-24 graph (tables 2 x 8.6 GB at D = 128: rows mostly miss L2, the regime of
configs[4]); D = 128 and 256."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from node2vec_b200 import synth
from node2vec_b200.graph import DeviceGraph
from node2vec_b200.sgns import Word2Vec

dev = torch.device("cuda", 0)
MODES = tuple(int(x) for x in os.environ.get("N2V_MODES", "0,1,2,3,4").split(","))


def time_modes(name, walks, dims):
    for dim in dims:
        for mode in MODES:
            os.environ["N2V_SGNS_MODE"] = str(mode)
            m = Word2Vec(size=dim, sg=1, iter=3, seed=1, batch_words=10000, **bench.SGNS_HP)
            m.build_vocab(walks)
            m.train(walks, epochs=1)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            m.train(walks, epochs=1)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            print(f"{name:10s} D={dim:3d} mode {mode}: {ms:8.1f} ms  {m.train_stats['pairs'] / ms / 1e6:7.3f} G pairs/s", flush=True)
            del m
            torch.cuda.empty_cache()


w2 = bench.WORKLOADS["blogcatalog_like"]
s2, d2 = bench.make_graph_host("blogcatalog_like")
g = DeviceGraph.from_arcs(s2, d2, None, n_vertices=w2["n"])
walks, alive, _ = g.walk(g.start_vertices(), w2["num_walks"], w2["walk_length"], w2["p"], w2["q"], seed=42)
del g
time_modes("blogcat", walks, (128,))
del walks
w = bench.WORKLOADS["rmat20"]
src, dst = bench.config3_arcs_device(w, dev)
g = DeviceGraph.from_arcs(src, dst, None, n_vertices=w["n"])
walks, alive, _ = g.walk(g.start_vertices(), w["num_walks"], w["walk_length"], w["p"], w["q"], seed=42)
del g, src, dst
time_modes("rmat20", walks, (128, 256))
del walks
torch.cuda.empty_cache()
src, dst = synth.rmat_partition_device(24, 16, 0, 1, dev)
g = DeviceGraph.from_arcs(src, dst, None, n_vertices=1 << 24)
del src, dst
walks, alive, _ = g.walk(g.start_vertices(), 2, 40, 0.25, 4.0, seed=42)
del g
torch.cuda.empty_cache()
time_modes("rmat24", walks, (128,))
