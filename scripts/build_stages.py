"""Per-library-call device times of the graph build on a bench workload (run under gpurun):
   python scripts/build_stages.py [workload]
Wraps every n2v_* entry the build goes through with a synchronise + wall clock, then times the whole
DeviceGraph.from_arcs and one end-to-end fugue.random_walk pass for comparison."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from node2vec_b200 import _lib, fugue
from node2vec_b200.graph import DeviceGraph

name = sys.argv[1] if len(sys.argv) > 1 else "rmat20"
w = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
if name == "rmat20":
    src, dst = bench.config3_arcs_device(w, dev)
else:
    s, d = bench.make_graph_host(name)
    src, dst = torch.as_tensor(s, device=dev), torch.as_tensor(d, device=dev)
lib = _lib.load()
times = {}


class Timed:
    def __init__(self, inner):
        self._inner = inner

    def __getattr__(self, fn_name):
        fn = getattr(self._inner, fn_name)
        if fn_name not in ("n2v_csr_build", "n2v_hash_build", "n2v_alias_build", "n2v_walk"):
            return fn

        def call(*a):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = fn(*a)
            torch.cuda.synchronize()
            times.setdefault(fn_name, []).append((time.perf_counter() - t0) * 1e3)
            return rc
        return call


def med(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


DeviceGraph.from_arcs(src, dst, None, n_vertices=w["n"])            # warm the allocator
_lib._lib = Timed(lib)
for _ in range(5):
    g = DeviceGraph.from_arcs(src, dst, None, n_vertices=w["n"])
    del g
_lib._lib = lib
for k, v in times.items():
    print(f"{k:24s} median {med(v):8.2f} ms   min {min(v):8.2f}")
tt = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g = DeviceGraph.from_arcs(src, dst, None, n_vertices=w["n"])
    torch.cuda.synchronize()
    tt.append((time.perf_counter() - t0) * 1e3)
    del g
print(f"{'from_arcs (device arcs)':24s} median {med(tt):8.2f} ms   min {min(tt):8.2f}")
sp, dp = src.cpu().pin_memory(), dst.cpu().pin_memory()
tt = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g = DeviceGraph.from_arcs(sp, dp, None, n_vertices=w["n"])
    torch.cuda.synchronize()
    tt.append((time.perf_counter() - t0) * 1e3)
    W = int(g.start_vertices().numel()) * w["num_walks"]
    del g
print(f"{'from_arcs (pinned arcs)':24s} median {med(tt):8.2f} ms   min {min(tt):8.2f}")
host_out = torch.empty((W, w["walk_length"] + 1), dtype=torch.int32, pin_memory=True)
params = {"num_walks": w["num_walks"], "walk_length": w["walk_length"], "return_param": w["p"], "inout_param": w["q"]}
tt = []
for _ in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fugue.random_walk(None, (sp, dp), dict(params), None, random_seed=1, out=host_out, n_vertices=w["n"])
    torch.cuda.synchronize()
    tt.append((time.perf_counter() - t0) * 1e3)
print(f"{'fugue.random_walk e2e':24s} median {med(tt[1:]):8.2f} ms   min {min(tt[1:]):8.2f}   all {[round(t, 1) for t in tt]}")
