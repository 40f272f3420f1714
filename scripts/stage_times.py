"""Per-stage device times of the end-to-end path (run under gpurun):
   python scripts/stage_times.py [workload]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from node2vec_b200.graph import DeviceGraph
from node2vec_b200.sgns import Word2Vec

name = sys.argv[1] if len(sys.argv) > 1 else "blogcatalog_like"
w = bench.WORKLOADS[name]
src, dst = bench.make_graph_host(name)
torch.cuda.synchronize()


def timed(label, fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    print(f"{label:34s} {(time.perf_counter() - t0) / n * 1e3:10.2f} ms")
    return out


src_t = torch.as_tensor(src).pin_memory(); dst_t = torch.as_tensor(dst).pin_memory()
timed("H2D arcs", lambda: (src_t.cuda(), dst_t.cuda()))
g = timed("DeviceGraph.from_arcs (K0+K0b+K1)", lambda: DeviceGraph.from_arcs(src_t, dst_t, None, n_vertices=w["n"]))
print("  arcs", g.n_arcs, "max degree", int(g.degrees().max()), "graph bytes %.1f MB" % (g.nbytes() / 1e6))
start = g.start_vertices()
res = timed("walk kernel", lambda: g.walk(start, w["num_walks"], w["walk_length"], w["p"], w["q"], seed=1))
walks = res[0]
host = torch.empty(walks.shape, dtype=torch.int32).pin_memory()
timed("D2H walk matrix (%.0f MB)" % (walks.numel() * 4 / 1e6), lambda: host.copy_(walks))
m = Word2Vec(size=w["dim"], sg=1, negative=5, min_count=1, iter=1, seed=1)
timed("build_vocab (K4)", lambda: m.build_vocab(walks))
timed("sgns epoch (K3)", lambda: m.train(walks, epochs=1), n=1)
