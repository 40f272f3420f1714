#!/usr/bin/env python
"""Where the end-to-end walk pass spends its time (run under gpurun): host-timed stages of
fugue.random_walk(host arcs) -> pinned host walk matrix, sequential vs pipelined delivery."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from node2vec_b200 import fugue
from node2vec_b200.graph import DeviceGraph, walk_to_host

name = sys.argv[1] if len(sys.argv) > 1 else "blogcatalog_like"
w = bench.WORKLOADS[name]
src, dst = bench.make_graph_host(name)
src_pin, dst_pin = torch.as_tensor(src).pin_memory(), torch.as_tensor(dst).pin_memory()


def timed(label, fn, n=7):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{label:52s} median {np.median(ts):7.2f} ms   min {min(ts):7.2f}   max {max(ts):7.2f}", flush=True)


g = DeviceGraph.from_arcs(src_pin, dst_pin, None, n_vertices=w["n"])
start = g.start_vertices()
W, L1 = int(start.numel()) * w["num_walks"], w["walk_length"] + 1
host = torch.empty((W, L1), dtype=torch.int32).pin_memory()
args = (w["num_walks"], w["walk_length"], w["p"], w["q"], 1)
timed("H2D arcs", lambda: (src_pin.cuda(non_blocking=True), dst_pin.cuda(non_blocking=True)))
timed("DeviceGraph.from_arcs (host arcs)", lambda: DeviceGraph.from_arcs(src_pin, dst_pin, None, n_vertices=w["n"]))
timed("walk kernel", lambda: g.walk(start, *args))
walks = g.walk(start, *args)[0]
timed("D2H view -> pinned (compaction + memcpy)", lambda: host.copy_(walks, non_blocking=True))
comp = walks.contiguous()
timed("D2H contiguous -> pinned (memcpy only)", lambda: host.copy_(comp, non_blocking=True))
timed("walk + D2H sequential", lambda: host.copy_(g.walk(start, *args)[0], non_blocking=True))
for cw in (1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 30):
    timed(f"walk_to_host chunk_walkers={cw}", lambda: walk_to_host(g, start, *args, host, chunk_walkers=cw))
prm = {"num_walks": w["num_walks"], "walk_length": w["walk_length"], "return_param": w["p"], "inout_param": w["q"]}
timed("fugue.random_walk sequential (+ copy)", lambda: host.copy_(fugue.random_walk(None, (src_pin, dst_pin), dict(prm), random_seed=1).walks_device, non_blocking=True))
timed("fugue.random_walk out=host (pipelined)", lambda: fugue.random_walk(None, (src_pin, dst_pin), dict(prm), random_seed=1, out=host))
