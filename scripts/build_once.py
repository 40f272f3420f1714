"""One trim_index + one graph build on configs[2] (for an ncu launch list of the build kernels):
   ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:'...' python scripts/build_once.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from node2vec_b200.graph import DeviceGraph

w = bench.WORKLOADS["rmat20"]
dev = torch.device("cuda", 0)
src, dst = bench.config3_arcs_device(w, dev)          # K5 trim_sample + K6 first_occurrence inside trim_index
g = DeviceGraph.from_arcs(src, dst, None, n_vertices=w["n"])   # K0 + K0b + K1
torch.cuda.synchronize()
print("arcs", g.n_arcs, "flags", g.flags)
