#!/usr/bin/env python
"""bench.py -- the node2vec hot path on B200: walk-steps/s + SGNS pairs/s.

Contract (task prompt): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on
rank 0.  Workloads are BASELINE.json's configs:

  N = 1   `rmat20`  configs[2]: R-MAT scale 20 (1 M vertices, 16 M edges) + 16 injected hotspot
          vertices of degree 2^18, trimmed with max_out_deg = 10000 through `trim_index` and
          symmetrised (the reference's own order), 10 walks x 80 from every vertex, SGNS D = 128.
          The largest configuration quoted on ONE B200; graph 1.8 GB and tables 1.07 GB, far
          beyond L2.  A `secondary` block carries configs[1] (`blogcatalog_like`, L2-resident).
  N > 1   `rmat26`  configs[4]: R-MAT scale 26 (67 M vertices, ~1.07 B edges, 2.1 G arcs),
          VERTEX-PARTITIONED CSR with NVLink peer reads, 5 walks x 40, D = 128, data-parallel SGNS
          with NCCL allreduce-averaging of the replicated tables.  STRONG scaling: one fixed graph
          and one fixed amount of work at N = 2 / 4 / 8.

A "step" is one pass of the hot path over one batch of synthetic input:
  walk   every start vertex of the workload x num_walks walkers x walk_length steps;
  SGNS   N = 1: one epoch over those walks.  N > 1: one slice (1 / --sgns-slices) of the epoch on
         every rank, both tables allreduce-averaged every --sync-steps steps INSIDE the timed region.

  value     whole-job walk-steps/s with the graph already packed in HBM (kernel only)
  e2e       the same metric through the public API (node2vec_b200.fugue.random_walk) with HOST
            buffers: H2D of the arc list, CSR + hash + alias build, walk, D2H of the walk matrix
  roofline  HBM model of the dominant kernel (SURVEY 8d bytes per unit x units / kernel time)
  cpu_baseline  the reference's own functions (baseline/_ref, kind "reference") on the host cores
  sgns      the same set of keys for the SGNS half

`--impl reference` times the reference's CPU path instead (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2] -- the N = 1 default
    "rmat20": dict(graph="rmat(scale 20, edge factor 16) + 16 hotspots of degree 2^18, trim_index(max_out_deg=10000, "
                         "directed=False)", n=1 << 20, scale=20, edge_factor=16, hotspots=16, hotspot_degree=1 << 18,
                   max_out_deg=10000, p=0.25, q=4.0, num_walks=10, walk_length=80, dim=128),
    # BASELINE.json configs[1] -- L2-resident; reported as the `secondary` block of the N = 1 line
    "blogcatalog_like": dict(graph="blogcatalog_like(10k vertices, 334k edges)", n=10000, m=334000,
                             p=0.25, q=4.0, num_walks=80, walk_length=40, dim=128),
    # BASELINE.json configs[0] -- the reference's own CPU-runnable case (parity tests; selectable here)
    "er_10k": dict(graph="erdos_renyi(10k vertices, 100k edges)", n=10000, m=100000,
                   p=1.0, q=0.5, num_walks=10, walk_length=20, dim=128),
    # BASELINE.json configs[4] -- the N > 1 default (vertex-partitioned, strong scaling)
    "rmat26": dict(graph="rmat(scale 26, edge factor 16), vertex-partitioned", scale=26, edge_factor=16,
                   p=0.25, q=4.0, num_walks=5, walk_length=40, dim=128),
}
SGNS_HP = dict(window=5, negative=5, alpha=0.025, min_alpha=1e-4, min_count=1, sample=1e-3)
SEED = 42


# ======================================================================================
# graphs
# ======================================================================================
def config3_arcs_device(w, dev, prep=None):
    """configs[2] on the device: undirected edge list -> trim_index(trim, then symmetrise).
    `prep` (a dict) receives the wall time of that trim_index call (not part of `value`)."""
    import torch
    from node2vec_b200 import fugue, synth
    src, dst = synth.rmat_hotspot_edges_device(w["scale"], w["edge_factor"], seed=SEED, device=dev,
                                               hotspots=w["hotspots"], hotspot_degree=w["hotspot_degree"])
    n_in = int(src.numel())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    (src, dst), _ = fugue.trim_index(None, (src, dst), indexed=True, directed=False, max_out_deg=w["max_out_deg"],
                                     random_seed=1)
    torch.cuda.synchronize()
    if prep is not None:
        prep.update({"trim_index_ms": round(1e3 * (time.perf_counter() - t0), 2), "edges_in": n_in,
                     "arcs_out": int(src.numel()),
                     "what": "fugue.trim_index(device arc tensors, indexed=True, directed=False, max_out_deg, "
                             "random_seed): K5 seeded trimming (numpy-RandomState-exact) + K6 undirected expansion; "
                             "first call, allocator cold"})
    return src.to(dev).contiguous(), dst.to(dev).contiguous()


def config3_arcs_host(w):
    """configs[2] built with numpy only (the reference arm must not touch the GPU library): same
    generator parameters and pipeline, numpy's random streams."""
    from node2vec_b200 import synth
    src, dst = synth.rmat_host(w["scale"], w["edge_factor"], seed=SEED)
    half = len(src) // 2
    lo, hi = src[:half].astype(np.int64), dst[:half].astype(np.int64)
    rng = np.random.default_rng(SEED + 1)
    n = 1 << w["scale"]
    hubs = rng.integers(0, n, w["hotspots"])
    hs = np.repeat(hubs, w["hotspot_degree"])
    hd = rng.integers(0, n, len(hs))
    a, b = np.concatenate([lo, np.minimum(hs, hd)]), np.concatenate([hi, np.maximum(hs, hd)])
    keys = np.unique(((a << 32) | b)[a != b])
    a, b = keys >> 32, keys & 0xFFFFFFFF
    # trim_hotspot_vertices on the listed direction (randomwalk.py:238-262), then both directions
    deg = np.bincount(a, minlength=n)
    keep = np.ones(len(a), dtype=bool)
    start = np.concatenate([[0], np.cumsum(deg)])
    for v in np.flatnonzero(deg > w["max_out_deg"]):
        drop = np.random.RandomState(1).permutation(int(deg[v]))[w["max_out_deg"]:]
        keep[start[v] + drop] = False
    a, b = a[keep], b[keep]
    return np.concatenate([a, b]).astype(np.int32), np.concatenate([b, a]).astype(np.int32)


def make_graph_host(name):
    from node2vec_b200 import synth
    w = WORKLOADS[name]
    if name == "blogcatalog_like":
        return synth.blogcatalog_like(w["n"], w["m"], seed=SEED)
    if name == "rmat20":
        return config3_arcs_host(w)
    if name == "rmat26":     # the CPU legs sample an R-MAT of the same generator parameters at scale 20
        return synth.rmat_host(20, w["edge_factor"], seed=SEED)
    return synth.erdos_renyi(w["n"], w["m"], seed=SEED)


def csr_host(src, dst, n):
    order = np.lexsort((dst, src))
    col = np.ascontiguousarray(dst[order]).astype(np.int32)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=row_ptr[1:])
    return row_ptr, col


# ======================================================================================
# peaks, ceilings, models
# ======================================================================================
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def ncu_traffic(workload, kernel):
    """DRAM bytes per launch measured by ncu for this workload/kernel (profiles/ncu_traffic.json), or None."""
    return _profile_json("ncu_traffic.json").get(workload, {}).get(kernel)


def gather_ceilings():
    """Random 32-byte-sector gather ceilings measured with scripts/gather_peak.cu / peer_gather.cu
    (profiles/gather_ceilings.json, G sectors/s per GPU)."""
    d = _profile_json("gather_ceilings.json")
    return {"l2": d.get("l2", 235.0), "hbm": d.get("hbm", 40.0), "nvlink_peer": d.get("nvlink_peer", 10.9),
            "source": d.get("source", "profiles/r01_gather_peak.txt, r01_peer_gather.txt")}


def walk_bytes_per_step(stats):
    """B_step = 16 (vertex header) + T * (12 (alias record) + 4 * l) + 4 (store), T = trials
    per step and l = membership probes per trial, both from the kernel's own counters."""
    steps = max(stats["steps"], 1)
    T = stats["trials"] / steps
    probes_per_trial = stats["probes"] / max(stats["trials"], 1)
    return 16.0 + T * (12.0 + 4.0 * probes_per_trial) + 4.0, T, probes_per_trial


def walk_roofline(stats, steps_per_launch, kernel_ms, level, workload, note, remote_frac=None):
    peak, peak_src = measured_peaks()
    b_step, T, lp = walk_bytes_per_step(stats)
    achieved = steps_per_launch * b_step / (kernel_ms * 1e-3) / 1e9
    sectors = (stats["trials"] + stats["probes"]) / max(stats["steps"], 1)
    ceil = gather_ceilings()
    g_ach = sectors * steps_per_launch / (kernel_ms * 1e-3) / 1e9
    if level == "nvlink_peer" and remote_frac is not None:     # harmonic mix of local HBM and remote gathers
        c = 1.0 / (remote_frac / ceil["nvlink_peer"] + (1.0 - remote_frac) / ceil["hbm"])
    else:
        c = ceil[level]
    traffic = ncu_traffic(workload, "walk_kernel")
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": "walk_kernel", "bytes_per_step": b_step,
            # ncu DRAM bytes of one launch of this workload over THIS run's kernel time, as a fraction of the copy peak
            "traffic_frac_of_peak": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
            "bytes_per_step_formula": "16 + T*(12 + 4*l) + 4 (SURVEY 8d)",
            "algorithmic_bytes_per_launch": steps_per_launch * b_step, "steps_per_launch": steps_per_launch,
            "trials_per_step": T, "probes_per_trial": lp, "sectors_per_step": sectors,
            "sector_model_bytes_per_launch": (32.0 * sectors + 4.0) * steps_per_launch,
            "gather": {"achieved_gsectors_per_s": g_ach, "ceiling_gsectors_per_s": c, "level": level,
                       "frac": g_ach / c, "source": ceil["source"]},
            "kernel_ms": kernel_ms, "peak_source": peak_src, "note": note}


def sgns_bytes_per_pair(dim, negative):
    """(K+2) rows read + (K+2) rows written, D fp32 each, no reuse assumed."""
    return 8.0 * dim * (negative + 2)


# ======================================================================================
# clocks
# ======================================================================================
class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def start(self):
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


# ======================================================================================
# CPU legs (the ONLY place that touches oracle/): reference functions / oracle port on the host cores
# ======================================================================================
_CSR = None


def _ref_walk_chunk(a):
    """One worker process: the unmodified reference functions (oracle/ref_driver.py) when the
    reference is importable, else the oracle's Python port of the same per-row algorithm."""
    starts, num_walks, L, p, q, budget = a
    from oracle import ref_driver
    rw, _ = ref_driver.load_reference()
    row_ptr, col = _CSR
    if rw is not None:
        return ref_driver.timed_sample(rw, row_ptr, col, None, starts, num_walks, L, p, q, budget, batch=8)
    import random
    from oracle import ref_walk

    class Lazy(dict):
        def __missing__(self, v):
            lo, hi = int(row_ptr[v]), int(row_ptr[v + 1])
            if lo == hi:
                raise KeyError(v)
            self[v] = (col[lo:hi].tolist(), [1.0] * (hi - lo))
            return self[v]

        def __contains__(self, v):
            return row_ptr[v + 1] > row_ptr[v]

    adj, rng = Lazy(), random.Random(1000 + int(starts[0]))
    steps = done = 0
    t0 = time.perf_counter()
    for v in starts:
        rows = ref_walk.start_rows([v], num_walks)
        for _ in range(L):
            rows = [ref_walk.step_row(r, adj, p, q, rng.random(), rng.random()) for r in rows if r["dst"] in adj]
        steps += len(rows) * L
        done += 1
        if time.perf_counter() - t0 > budget:
            break
    return steps, time.perf_counter() - t0, done


def cpu_baseline_walk(name, graph=None, budget_s=10.0, procs=None):
    """The reference's walk path on the host cores, a bounded sample of the same workload: every
    process walks seeded start vertices of the same graph (num_walks walkers each, full
    walk_length) until `budget_s` seconds of step-loop time are spent.  Adjacency rows are built
    lazily and outside the timed region (one-off in the reference; the GPU `value` excludes its
    build too)."""
    import multiprocessing as mp
    global _CSR
    from oracle import ref_driver
    w = WORKLOADS[name]
    src, dst = graph if graph is not None else make_graph_host(name)
    n = int(max(src.max(), dst.max())) + 1
    _CSR = csr_host(src, dst, n)
    rw, root = ref_driver.load_reference()
    procs = procs or os.cpu_count() or 1
    rng = np.random.default_rng(0)
    starts = rng.permutation(np.flatnonzero(np.diff(_CSR[0]) > 0))
    chunks = [c.tolist() for c in np.array_split(starts, procs) if len(c)]
    nw = min(int(w["num_walks"]), 10)
    args = [(c, nw, w["walk_length"], w["p"], w["q"], budget_s) for c in chunks]
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        res = pool.map(_ref_walk_chunk, args)
    steps = int(sum(r[0] for r in res))
    rate = float(sum(r[0] / max(r[1], 1e-9) for r in res))       # processes run concurrently on their own cores
    kind = "reference" if rw is not None else "port"
    what = (f"unmodified node2vec.randomwalk functions from {os.path.relpath(root, ROOT) if root.startswith(ROOT) else root} "
            "chained as fugue.py:130-153 does, pandas merges for the two joins per step"
            if rw is not None else "oracle/ref_walk.py port of the reference's per-row algorithm (reference not importable)")
    graph_note = "" if name != "rmat26" else " on an R-MAT scale-20 graph of the same generator parameters"
    return {"value": rate, "unit": "walk-steps/s", "cores": len(chunks), "kind": kind,
            "sample": f"{int(sum(r[2] for r in res))} start vertices x {nw} walks x {w['walk_length']} steps = {steps} "
                      f"walker-steps{graph_note}, {max(r[1] for r in res):.1f}s of step-loop time per process on "
                      f"{len(chunks)} processes ({what}; adjacency rows prebuilt outside the timed region)"}


def cpu_baseline_walk_c(name, graph=None, budget_s=5.0):
    """Same algorithm, C port (oracle/csrc/n2v_oracle.c: per walker per step it re-derives the biased
    weights and rebuilds the alias table, like the reference) on all host cores -- context for how
    much of the Python baseline is interpreter overhead."""
    from oracle import clib
    w = WORKLOADS[name]
    src, dst = graph if graph is not None else make_graph_host(name)
    n = int(max(src.max(), dst.max())) + 1
    row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, None, n)
    rng = np.random.default_rng(0)
    starts = rng.permutation(np.flatnonzero(np.diff(row_ptr) > 0)).astype(np.int32)
    threads = os.cpu_count() or 1
    steps, dt, n_start = 0, 0.0, 64 * threads
    lo = 0
    while dt < budget_s and lo < len(starts):
        t0 = time.perf_counter()
        walks, alive = clib.reference_walk(row_ptr, col, ws, starts[lo:lo + n_start], 1, w["walk_length"], w["p"],
                                           w["q"], "naive", None, threads=threads)
        dt += time.perf_counter() - t0
        steps += int(alive.sum()) * w["walk_length"]
        lo += n_start
        n_start *= 2
    return {"value": steps / dt, "unit": "walk-steps/s", "cores": threads, "kind": "port",
            "sample": f"{steps} steps in {dt:.1f}s, {threads} threads (C port of the reference's per-row algorithm)"}


def cpu_baseline_sgns(walks, n_rows, dim, budget_s=10.0, threads=None):
    """gensim-3.8 SGNS restatement (oracle/csrc/sgns_ref.c) with lock-free host threads on a
    bounded sample of the same walk matrix: as many leading walks as fit the time budget."""
    from oracle import clib
    threads = threads or os.cpu_count() or 1
    walks = np.ascontiguousarray(walks[: min(len(walks), 1 << 20)])
    counts = np.bincount(walks.reshape(-1), minlength=n_rows)
    syn0, syn1 = clib.sgns_init(n_rows, dim, 1)
    n = min(len(walks), 2000 * threads)
    pairs, dt = 0, 0.0
    lo = seen = 0
    while dt < budget_s:                      # wraps around a small sample (further epochs over the same walks)
        t0 = time.perf_counter()
        pairs += clib.sgns_train(walks[lo:lo + n], counts, syn0, syn1, epochs=1, seed=1, batch_words=10000,
                                 threads=threads, **SGNS_HP)
        dt += time.perf_counter() - t0
        seen += len(walks[lo:lo + n])
        lo = lo + n if lo + n < len(walks) else 0
    return {"value": pairs / dt, "unit": "pairs/s", "cores": threads, "kind": "port",
            "sample": f"{seen} walks x {walks.shape[1]} tokens (a sample of {len(walks)} walks, re-read when exhausted) = "
                      f"{pairs} pairs in {dt:.1f}s, {threads} lock-free threads (oracle gensim-3.8 restatement; gensim "
                      f"itself is not installable)"}


def workload_config(name, w, n_gpus):
    """The `config` dict shared by both arms (same keys, so the driver can compare them)."""
    cfg = {"workload": name, **{k: w[k] for k in ("graph", "p", "q", "num_walks", "walk_length", "dim")},
           "sgns": dict(SGNS_HP), "n_gpus": n_gpus,
           "l2": "GPU arm: L2 flushed (256 MiB write) between timed iterations and the inputs exceed L2 "
                 "(except the L2-resident secondary block); CPU reference arm: host memory"}
    return cfg


def run_reference(args):
    """--impl reference: the reference's CPU path for the same metric/config (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = args.workload
    w = WORKLOADS[name]
    graph = make_graph_host(name)
    vals, base = [], None
    for _ in range(max(1, min(args.steps, 2))):
        base = cpu_baseline_walk(name, graph, budget_s=12.0)
        vals.append(base["value"])
    v = float(np.mean(vals))
    # SGNS leg: walks for the sample come from the oracle's C port of the reference walk
    from oracle import clib
    src, dst = graph
    n = int(max(src.max(), dst.max())) + 1
    row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, None, n)
    # a few thousand walks are enough to time SGNS (the C port of the reference walk costs O(degree) per step)
    starts = np.random.default_rng(1).permutation(np.flatnonzero(np.diff(row_ptr) > 0).astype(np.int32))
    starts = np.sort(starts[: 4096 if name in ("rmat20", "rmat26") else 50000])
    walks_cpu, alive = clib.reference_walk(row_ptr, col, ws, starts, 2, w["walk_length"], w["p"], w["q"], "naive",
                                           None, threads=os.cpu_count() or 1)
    sgns_base = cpu_baseline_sgns(walks_cpu[alive], n, w["dim"])
    steps_per_pass = int((np.diff(row_ptr) > 0).sum()) * w["num_walks"] * w["walk_length"]
    line = {
        "impl": "reference", "metric": "walk_steps_per_s", "value": v, "unit": "walk-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * steps_per_pass / v, "higher_is_better": True,
        "scaling": "weak" if args.gpus == 1 else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(name, w, args.gpus),
        "note": "each step is a bounded sample of the workload (cpu_baseline.sample); ms_per_step extrapolates the "
                "sampled rate to one full pass",
        "cpu_baseline": {**base, "value": v},
        "e2e": {"value": v, "unit": "walk-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sgns": {"metric": "sgns_pairs_per_s", "value": sgns_base["value"], "unit": "pairs/s",
                 "cpu_baseline": sgns_base,
                 "e2e": {"value": sgns_base["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0,
                         "d2h_bytes_per_step": 0}},
    }
    emit(line)


# ======================================================================================
# small frame stand-ins for the e2e passes
# ======================================================================================
class _IdFrame:
    """walk_seed stand-in: a frame with an `id` column."""

    def __init__(self, ids):
        import pandas as pd
        self._df = pd.DataFrame({"id": ids})
        self.schema = ["id"]

    def as_pandas(self):
        return self._df


class _HostWalks:
    """A [src, walk] frame stand-in whose walk matrix is a pinned host tensor."""

    def __init__(self, t):
        self._t = t

    def __getitem__(self, key):
        assert key == "walk"
        return _Col(self._t)


class _Col:
    def __init__(self, t):
        self._t = t

    def tolist(self):
        return self._t.numpy()


_REAL_STDOUT = None


def _quiet_stdout():
    """Route fd 1 to stderr while we work (NCCL / libraries may print banners on stdout); the
    one JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_AFFINITY0 = None


def pin_to_gpu_numa(index):
    """Run this process on the cores NVML reports as local to GPU `index` (its NUMA node) before any
    pinned host buffer is allocated: host<->device copies then stay on the GPU's own root complex
    (round-1 review: eight ranks sharing one node's PCIe path bent the end-to-end curve).  Best effort:
    silently does nothing where NVML or sched_setaffinity is unavailable or the set would be empty."""
    global _AFFINITY0
    try:
        import pynvml
        if _AFFINITY0 is None:
            _AFFINITY0 = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        try:                   # NVML numbers physical devices; CUDA_VISIBLE_DEVICES may have remapped `index`
            import torch
            pr = torch.cuda.get_device_properties(int(index))
            h = pynvml.nvmlDeviceGetHandleByPciBusId(b"%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
        except Exception:  # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(int(index))
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1} & set(_AFFINITY0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:  # noqa: BLE001
        pass
    return None


def unpin():
    """Give the CPU legs (one process per core) the original affinity back."""
    if _AFFINITY0 is not None:
        try:
            os.sched_setaffinity(0, _AFFINITY0)
        except OSError:
            pass


# ======================================================================================
# single GPU: replicated graph
# ======================================================================================
def timed_walk(torch, g, start, nw, w, out, flush, steps, warmup):
    """`warmup` untimed + `steps` timed full passes; returns per-pass milliseconds (CUDA events on the
    launching stream, L2 flushed between iterations outside the events)."""
    def one_pass():
        g.walk(start, nw, w["walk_length"], w["p"], w["q"], seed=SEED, out=out)
    for _ in range(warmup):
        flush.fill_(1)
        one_pass()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        one_pass()
        b.record()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev]


def timed_sgns_epochs(torch, m, walks, flush, steps, warmup):
    for _ in range(warmup):
        flush.fill_(1)
        m.train(walks, epochs=1)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    pairs = 0
    for a, b in ev:
        flush.fill_(1)
        a.record()
        m.train(walks, epochs=1)       # one n2v_sgns_train launch
        b.record()
        pairs += m.train_stats["pairs"]
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ev], pairs


def sgns_block(torch, walks, host_walks, w, name, flush, steps, warmup, e2e_passes, table_note):
    """SGNS half on one GPU: `steps` epochs over the walk matrix in HBM; e2e = Node2VecGensim(host
    walks).fit() + host vectors."""
    from node2vec_b200.embedding import Node2VecGensim
    from node2vec_b200.sgns import Word2Vec
    dim, K = w["dim"], SGNS_HP["negative"]
    m = Word2Vec(size=dim, sg=1, iter=steps + warmup, seed=1, batch_words=10000, **SGNS_HP)
    m.build_vocab(walks)
    ms, pairs = timed_sgns_epochs(torch, m, walks, flush, steps, warmup)
    n_rows = int(m.syn0.shape[0])
    del m
    torch.cuda.empty_cache()

    def e2e_pass():
        n2v = Node2VecGensim(_HostWalks(host_walks), {"sg": 1, "iter": 1, "size": dim, **SGNS_HP}, random_seed=1)
        model = n2v.fit()
        vec = model.wv.vectors          # host numpy (D2H inside)
        return model.train_stats["pairs"], vec.nbytes
    e2e_pass()
    e2e_times, p_e2e, d2h = [], 0, 0
    for _ in range(e2e_passes):
        t0 = time.perf_counter()
        p_e2e, d2h = e2e_pass()
        e2e_times.append(time.perf_counter() - t0)
    log("sgns e2e pass times (ms):", [round(t * 1e3, 1) for t in e2e_times])
    peak, peak_src = measured_peaks()
    b_pair = sgns_bytes_per_pair(dim, K)
    kernel_ms = float(np.mean(ms))
    per_launch = pairs / steps
    achieved = per_launch * b_pair / (kernel_ms * 1e-3) / 1e9
    return {
        "metric": "sgns_pairs_per_s", "value": pairs / (sum(ms) * 1e-3), "unit": "pairs/s",
        "ms_per_step": float(sum(ms)) / steps, "steps": steps, "dtype": "f32", "gpu_launches": steps,
        "config": {"dim": dim, **SGNS_HP, "walks_per_gpu": int(walks.shape[0]), "tokens_per_walk": int(walks.shape[1]),
                   "table_rows": n_rows, "updates": "red.global.add.v4.f32", "sync": "none",
                   "l2": "flushed between timed iterations; " + table_note},
        "e2e": {"value": p_e2e / float(np.median(e2e_times)), "unit": "pairs/s",
                "h2d_bytes_per_step": int(host_walks.numel() * 4), "d2h_bytes_per_step": int(d2h),
                "what": "Node2VecGensim(host walks).fit(): H2D + vocab_count + sgns_prepare + init + 1 epoch + D2H vectors; "
                        "median of %d passes" % e2e_passes},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(name, "sgns_kernel"), "kernel": "sgns_kernel",
                     "traffic_frac_of_peak": (ncu_traffic(name, "sgns_kernel") / (kernel_ms * 1e-3) / 1e9 / peak)
                     if ncu_traffic(name, "sgns_kernel") else None,
                     "bytes_per_pair": b_pair, "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_launch": per_launch * b_pair, "pairs_per_launch": per_launch,
                     "peak_source": peak_src, "note": table_note},
    }


def walk_block(torch, name, src, dst, w, dev, flush, steps, warmup, e2e_passes, level, note, clocks=None):
    """Walk half on one GPU for workload `name` whose arcs (src, dst) are on the device."""
    from node2vec_b200 import fugue
    from node2vec_b200.graph import DeviceGraph
    g = DeviceGraph.from_arcs(src, dst, None, n_vertices=w["n"])
    start = g.start_vertices()
    nw, L = w["num_walks"], w["walk_length"]
    W = int(start.numel()) * nw
    steps_per_pass = W * L
    pitch = (L + 1 + 7) // 8 * 8
    out = torch.empty((W, pitch), dtype=torch.int32, device=dev)
    _, _, stats = g.walk(start, nw, L, w["p"], w["q"], seed=SEED, collect_stats=True)
    if clocks is not None:
        clocks.start()
    ms = timed_walk(torch, g, start, nw, w, out, flush, steps, warmup)
    if clocks is not None:
        clocks.stop()
    kernel_ms = float(np.mean(ms))
    deg = g.degrees()
    graph_info = {"vertices": int(g.n_vertices), "arcs": int(g.n_arcs), "graph_bytes": int(g.nbytes()),
                  "max_degree": int(deg.max()), "flags": int(g.flags), "walkers": W}
    # ---- end to end through the public API, host buffers in, host walk matrix out
    src_pin, dst_pin = src.cpu().pin_memory(), dst.cpu().pin_memory()
    host_out = torch.empty((W, L + 1), dtype=torch.int32, pin_memory=True)
    params = {"num_walks": nw, "walk_length": L, "return_param": w["p"], "inout_param": w["q"]}
    graph_bytes = g.nbytes()
    del g
    torch.cuda.empty_cache()

    def e2e_pass():
        res = fugue.random_walk(None, (src_pin, dst_pin), dict(params), None, random_seed=SEED, out=host_out,
                                n_vertices=w["n"])
        torch.cuda.synchronize()
        assert res.walks.shape == (W, L + 1) and res.walks[0, 0] >= 0
    e2e_pass()
    e2e_times = []
    for _ in range(e2e_passes):
        t0 = time.perf_counter()
        e2e_pass()
        e2e_times.append(time.perf_counter() - t0)
    log(f"[{name}] walk e2e pass times (ms):", [round(t * 1e3, 2) for t in e2e_times])
    block = {
        "metric": "walk_steps_per_s", "value": steps_per_pass * steps / (sum(ms) * 1e-3), "unit": "walk-steps/s",
        "ms_per_step": float(sum(ms)) / steps, "steps": steps, "dtype": "u32", "gpu_launches": steps,
        "graph": graph_info,
        "e2e": {"value": steps_per_pass / float(np.median(e2e_times)), "unit": "walk-steps/s",
                "h2d_bytes_per_step": int(src_pin.numel() * 4 + dst_pin.numel() * 4),
                "d2h_bytes_per_step": int(host_out.numel() * 4),
                "pass_ms": [round(1e3 * t, 2) for t in e2e_times],
                "what": "fugue.random_walk(host arcs) = H2D + csr/hash/alias build + walk, D2H of the walk matrix "
                        "pipelined under the walk; median of %d passes (mean %.2f ms, best %.2f ms; the copies share "
                        "the host's PCIe with other tenants of the box)" % (e2e_passes, 1e3 * float(np.mean(e2e_times)),
                                                                             1e3 * float(np.min(e2e_times)))},
        "roofline": walk_roofline(stats, steps_per_pass, kernel_ms, level, name, note),
        "walk_stats": stats,
    }
    return block, out[:, : L + 1], host_out, graph_bytes


def pandas_e2e(torch, src, dst, w):
    """One pass with the reference's own frame types on BOTH sides: a pandas arc frame in, the
    [src, walk] pandas frame (Python lists) out -- what a caller of the reference receives."""
    import pandas as pd
    from node2vec_b200 import fugue
    df = pd.DataFrame({"src": src.cpu().numpy(), "dst": dst.cpu().numpy()})
    params = {"num_walks": w["num_walks"], "walk_length": w["walk_length"], "return_param": w["p"],
              "inout_param": w["q"]}
    fugue.random_walk(None, df, dict(params), None, random_seed=SEED).as_pandas()
    t0 = time.perf_counter()
    res = fugue.random_walk(None, df, dict(params), None, random_seed=SEED).as_pandas()
    dt = time.perf_counter() - t0
    steps = len(res) * w["walk_length"]
    return {"value": steps / dt, "unit": "walk-steps/s", "seconds": dt, "rows": len(res),
            "what": "fugue.random_walk(pandas frame).as_pandas(): pandas in, [src, walk] frame of Python lists out"}


def bench_single(args):
    import torch
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    numa_cpus = pin_to_gpu_numa(0)
    name = args.workload
    w = WORKLOADS[name]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    clocks = ClockSampler(0)
    prep = {}
    if name == "rmat20":
        src, dst = config3_arcs_device(w, dev, prep)
        level, note = "hbm", ("graph %.2f GB, tables 2 x %.2f GB: DRAM-resident; the walk is a stream of random 32-byte "
                              "sector gathers, so the HBM fraction by algorithmic bytes is bounded by the random-sector "
                              "ceiling (roofline.gather), not by the copy peak")
    else:
        from node2vec_b200 import synth  # noqa: F401
        s, d = make_graph_host(name)
        src, dst = torch.as_tensor(s, device=dev), torch.as_tensor(d, device=dev)
        level, note = "l2", ("graph %.3f GB, tables 2 x %.3f GB: L2-resident, so the 'HBM' fraction is an "
                             "effective-bandwidth figure (ncu dram bytes = the walk matrix only)")
    e2e_passes = 5
    walk, walks_dev, host_out, graph_bytes = walk_block(torch, name, src, dst, w, dev, flush, args.steps, args.warmup,
                                                        e2e_passes, level, "", clocks)
    table_gb = w["n"] * w["dim"] * 4 / 1e9
    walk["roofline"]["note"] = note % (graph_bytes / 1e9, table_gb)
    sgns = None
    if not args.no_sgns:
        clocks.start()
        sgns = sgns_block(torch, walks_dev, host_out, w, name, flush, args.steps, args.warmup,
                          2 if name == "rmat20" else 3,
                          "tables 2 x %.2f GB (%s)" % (table_gb, "beyond L2" if table_gb > 0.1 else "L2-resident"))
        clocks.stop()
    line = {
        "metric": "walk_steps_per_s", "value": walk["value"], "unit": "walk-steps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": walk["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(name, w, 1),
        "detail": {**walk["graph"], "sharding": "one GPU, replicated CSR",
                   "note": "the N = 1 line is configs[2] (the largest configuration quoted on one B200); --gpus N > 1 runs "
                           "configs[4] (RMAT-26, vertex-partitioned, strong scaling over N = 2/4/8): compare N >= 2 lines "
                           "with each other, not with this one",
                   "host_cores_used_for_e2e": len(numa_cpus) if numa_cpus else None,
                   **({"prep": prep} if prep else {})},
        "gpu_launches": walk["gpu_launches"] + (sgns["gpu_launches"] if sgns else 0),
        "e2e": walk["e2e"], "roofline": walk["roofline"], "clocks": clocks.summary(), "walk_stats": walk["walk_stats"],
    }
    if sgns:
        sgns.pop("gpu_launches")
        line["sgns"] = sgns
    host_graph = (src.cpu().numpy(), dst.cpu().numpy())
    host_sample = host_out.numpy()
    del walks_dev
    torch.cuda.empty_cache()
    # ---- secondary block: configs[1] (L2-resident) on the same GPU, fewer steps
    if name == "rmat20" and not args.no_secondary:
        w2 = WORKLOADS["blogcatalog_like"]
        s, d = make_graph_host("blogcatalog_like")
        s2, d2 = torch.as_tensor(s, device=dev), torch.as_tensor(d, device=dev)
        k2 = max(3, min(args.steps, 10))
        b2, walks2, host2, gb2 = walk_block(torch, "blogcatalog_like", s2, d2, w2, dev, flush, k2, 3, 5, "l2", "")
        b2["roofline"]["note"] = ("graph %.3f GB: L2-resident, effective-bandwidth figure" % (gb2 / 1e9))
        b2["config"] = workload_config("blogcatalog_like", w2, 1)
        b2["e2e_pandas"] = pandas_e2e(torch, s2, d2, w2)
        if not args.no_sgns:
            sg2 = sgns_block(torch, walks2, host2, w2, "blogcatalog_like", flush, max(3, min(args.steps, 5)), 3, 3,
                             "tables 2 x 0.005 GB (L2-resident): effective-bandwidth figure, can exceed the HBM peak")
            sg2.pop("gpu_launches")
            b2["sgns"] = sg2
        b2.pop("gpu_launches")
        line["secondary"] = b2
        del walks2, host2
    unpin()
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_walk(name, host_graph)
        line["cpu_baseline_c"] = cpu_baseline_walk_c(name, host_graph)
        if sgns:
            line["sgns"]["cpu_baseline"] = cpu_baseline_sgns(host_sample, w["n"], w["dim"])
    emit(line)


# ======================================================================================
# N > 1: vertex-partitioned graph + data-parallel SGNS (BASELINE configs[4]), strong scaling
# ======================================================================================
def multi_gpu_parity(torch, dist, dev, rank, world):
    """Sharding must not change results (rows are independent, fugue.py:146-150): the partitioned
    graph's walks equal the replicated graph's, bit for bit, on every rank."""
    from node2vec_b200 import synth
    from node2vec_b200.graph import DeviceGraph, PartitionedGraph
    scale = 16
    V = 1 << scale
    src, dst = synth.rmat_device(scale, 8, seed=3, device=dev)         # same seed => same graph on every rank
    full = DeviceGraph.from_arcs(src, dst, None, n_vertices=V)
    S = (V + world - 1) // world
    mine = (src >= rank * S) & (src < (rank + 1) * S)
    part = PartitionedGraph.from_local_arcs(src[mine], dst[mine], None, V, group=dist.group.WORLD, assume_symmetric=True)
    start = part.start_vertices()
    a, alive_a, _ = part.walk(start, 4, 30, 0.25, 4.0, seed=11)
    b, alive_b, _ = full.walk(start, 4, 30, 0.25, 4.0, seed=11)
    ok = torch.tensor([1 if (torch.equal(a, b) and torch.equal(alive_a, alive_b) and part.flags == full.flags) else 0],
                      device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.barrier()
    part.close()
    del full, part, a, b
    torch.cuda.empty_cache()
    return bool(ok.item())


def tables_identical(torch, dist, m, dev, world):
    """After an averaging step every rank must hold bit-identical tables."""
    def checksum(t):                       # int64 sum of the fp32 bit patterns, in row blocks (no table-sized temporary)
        bits, tot = t.view(torch.int32), torch.zeros((), dtype=torch.int64, device=dev)
        for lo in range(0, bits.shape[0], 1 << 20):
            tot += bits[lo:lo + (1 << 20)].sum(dtype=torch.int64)
        return tot
    sig = torch.stack([checksum(m.syn0), checksum(m.syn1neg), m.syn0.view(torch.int32)[:: 4097].sum(dtype=torch.int64)])
    allsig = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(allsig, sig)
    return all(bool(torch.equal(allsig[0], s)) for s in allsig)


def bench_partitioned(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pin_to_gpu_numa(local_rank)
    dist.init_process_group("nccl", device_id=dev)
    from node2vec_b200 import fugue, synth
    from node2vec_b200.embedding import Node2VecGensim
    from node2vec_b200.graph import PartitionedGraph
    from node2vec_b200.sgns import Word2Vec
    name = args.workload
    w = dict(WORKLOADS[name])
    if args.rmat_scale:
        w["scale"] = args.rmat_scale
        w["graph"] = f"rmat(scale {args.rmat_scale}, edge factor {w['edge_factor']}), vertex-partitioned"
    V = 1 << w["scale"]
    S = (V + world - 1) // world
    nw, L, dim, K = w["num_walks"], w["walk_length"], w["dim"], SGNS_HP["negative"]
    t_all = time.perf_counter()

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()

    def reduce_max(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    parity_walk = multi_gpu_parity(torch, dist, dev, rank, world)
    log(f"[rank {rank}] partitioned == replicated walks: {parity_walk}")

    # ---- input synthesis on the devices (not timed)
    t0 = time.perf_counter()
    src, dst = synth.rmat_partition_device(w["scale"], w["edge_factor"], rank, world, dev)
    sync_all()
    t_gen = time.perf_counter() - t0
    n_local_arcs = int(src.numel())
    src_pin, dst_pin = src.cpu().pin_memory(), dst.cpu().pin_memory()
    del src, dst
    torch.cuda.empty_cache()
    log(f"[rank {rank}] generated {n_local_arcs} local arcs in {t_gen:.1f}s")

    # ---- e2e walk: host arcs -> public API -> host walk matrix (graph build + walk + D2H inside)
    params = {"num_walks": nw, "walk_length": L, "return_param": w["p"], "inout_param": w["q"]}
    W_cap = S * nw
    host_out = torch.empty((W_cap, L + 1), dtype=torch.int32, pin_memory=True)
    e2e_times, W = [], 0
    for it in range(1 + args.e2e_passes):
        sync_all()
        t0 = time.perf_counter()
        res = fugue.random_walk(None, (src_pin, dst_pin), dict(params), None, random_seed=SEED, out=host_out,
                                process_group=dist.group.WORLD, n_vertices=V, assume_symmetric=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        W = int(res.walks_device.shape[0])
        del res
        torch.cuda.empty_cache()
        if it > 0:
            e2e_times.append(reduce_max(dt))
    log(f"[rank {rank}] walk e2e pass times (s): {[round(t, 2) for t in e2e_times]}")

    # ---- the resident graph
    sync_all()
    t0 = time.perf_counter()
    g = PartitionedGraph.from_local_arcs(src_pin.to(dev), dst_pin.to(dev), None, V, group=dist.group.WORLD,
                                         assume_symmetric=True, keep_weight=False)
    torch.cuda.empty_cache()
    sync_all()
    t_build = time.perf_counter() - t0
    start = g.start_vertices()
    W = int(start.numel()) * nw
    steps_per_pass = W * L
    pitch = (L + 1 + 7) // 8 * 8
    out = torch.empty((W, pitch), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    n_stat = min(int(start.numel()), 200000)
    wk, _, stats = g.walk(start[:n_stat], nw, L, w["p"], w["q"], seed=SEED, collect_stats=True)
    remote = float((torch.div(wk[:, 1:], S, rounding_mode="floor") != rank).float().mean().item())
    del wk
    clocks = ClockSampler(local_rank)
    sync_all()
    if rank == 0:
        clocks.start()
    ms = timed_walk(torch, g, start, nw, w, out, flush, args.steps, args.warmup)
    dist.barrier()
    if rank == 0:
        clocks.stop()
    total_ms = reduce_max(sum(ms))
    job_steps = reduce_sum(steps_per_pass)
    value = job_steps * args.steps / (total_ms * 1e-3)
    kernel_ms = float(np.mean(ms))
    graph_bytes = int(g.nbytes())
    log(f"[rank {rank}] walk {kernel_ms:.1f} ms/pass, {steps_per_pass / kernel_ms / 1e6:.2f} G steps/s on this rank, "
        f"{remote:.0%} remote hops")

    # ---- informational: the same graph REPLICATED from the parts (PartitionedGraph.localize: peers' arc records
    # and hash sets copied over NVLink, 40 B/arc) -- what the product does when the graph fits beside the walk
    # matrix; configs[4] names the partitioned form, so `value` above stays the peer-read walk
    localized = None
    if not args.no_localized:
        sync_all()
        t0 = time.perf_counter()
        fits = g.localize()
        sync_all()
        t_copy = time.perf_counter() - t0
        if fits:
            k_loc = max(3, min(args.steps, 5))
            check_rows = min(W, 200000)
            ref_rows = out[:check_rows].clone()
            ms_l = timed_walk(torch, g, start, nw, w, out, flush, k_loc, 3)
            same = torch.tensor([1 if torch.equal(out[:check_rows], ref_rows) else 0], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            del ref_rows
            tot_l = reduce_max(sum(ms_l))
            localized = {"value": job_steps * k_loc / (tot_l * 1e-3), "unit": "walk-steps/s", "ms_per_step": tot_l / k_loc,
                         "steps": k_loc, "copy_s": t_copy, "copied_bytes_per_gpu": int(sum(
                             (a.numel() + h.numel()) * 4 for a, h in g._local_parts.values())),
                         "same_walks_as_partitioned": bool(same.item()),
                         "what": "PartitionedGraph.localize(): every rank copies its peers' arc records + hash sets over "
                                 "NVLink and walks on local HBM only (replicated CSR assembled from the partitioned build)"}
            log(f"[rank {rank}] localized walk {float(np.mean(ms_l)):.1f} ms/pass (copy {t_copy:.2f}s)")
        else:
            localized = {"unavailable": "the peers' parts do not fit beside the walk matrix on every rank"}

    # the graph is not needed any more: unmap the peers' parts and free this rank's (collective) before the
    # 2 x 34 GB tables are allocated
    n_arcs_total = int(g.n_arcs)
    sync_all()
    g.close()
    del g
    torch.cuda.empty_cache()

    # ---- SGNS: data-parallel over the rank's own walks, replicated tables averaged by NCCL
    sgns = None
    parity_tables = None
    walks = out[:, : L + 1]
    if not args.no_sgns:
        slices, sync_steps = args.sgns_slices, args.sync_steps
        total_steps = args.steps + args.warmup
        epochs = (total_steps + slices - 1) // slices
        m = Word2Vec(size=dim, sg=1, iter=epochs, seed=1, batch_words=10000, process_group=dist.group.WORLD, **SGNS_HP)
        t0 = time.perf_counter()
        m.build_vocab(walks)
        torch.cuda.empty_cache()
        sync_all()
        t_vocab = time.perf_counter() - t0
        bounds = [W * i // slices for i in range(slices + 1)]
        ev, sync_ev, pairs, n_sync = [], [], 0, 0
        if rank == 0:
            clocks.start()
        for step in range(total_steps):
            timed = step >= args.warmup
            if step == args.warmup:
                sync_all()
            sl = step % slices
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            m.train(walks, epochs=epochs, epoch_range=(step // slices, step // slices + 1),
                    walk_range=(bounds[sl], bounds[sl + 1]), sync=False)
            due = (step + 1) % sync_steps == 0
            if due:
                c, d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c.record()
                m.average_tables()                  # NCCL sum-allreduce of both tables + n2v_scale(1/G)
                d.record()
            b.record()
            if timed:
                ev.append((a, b))
                pairs += m.train_stats["pairs"]
                if due:
                    sync_ev.append((c, d))
                    n_sync += 1
        sync_all()
        if rank == 0:
            clocks.stop()
        sg_ms = [a.elapsed_time(b) for a, b in ev]
        sync_ms = [c.elapsed_time(d) for c, d in sync_ev]
        sg_total = reduce_max(sum(sg_ms))
        job_pairs = reduce_sum(pairs)
        sync_total = reduce_max(sum(sync_ms)) if sync_ms else 0.0
        m.average_tables()
        parity_tables = tables_identical(torch, dist, m, dev, world)
        table_bytes = int(2 * m.syn0.numel() * 4)
        n_rows = int(m.syn0.shape[0])
        del m
        torch.cuda.empty_cache()
        # e2e: host walks in, fit() on every rank (H2D + vocab + init + ONE full epoch + the averaging), vectors out on rank 0
        del out, walks
        torch.cuda.empty_cache()
        sync_all()
        t0 = time.perf_counter()
        n2v = Node2VecGensim(_HostWalks(host_out[:W]), {"sg": 1, "iter": 1, "size": dim, "process_group": dist.group.WORLD,
                                                       **SGNS_HP}, random_seed=1)
        model = n2v.fit()
        p_e2e = model.train_stats["pairs"]
        d2h = 0
        if rank == 0:
            vec = model.syn0.cpu()                  # the averaged table is identical on every rank: one copy out
            d2h = int(vec.numel() * 4)
            del vec
        torch.cuda.synchronize()
        e2e_sg_t = reduce_max(time.perf_counter() - t0)
        e2e_sg_pairs = reduce_sum(p_e2e)
        del model, n2v
        torch.cuda.empty_cache()
        peak, peak_src = measured_peaks()
        b_pair = sgns_bytes_per_pair(dim, K)
        compute_ms = (sum(sg_ms) - sum(sync_ms)) / len(sg_ms)            # this rank's kernel time per step
        per_launch = pairs / len(sg_ms)
        achieved = per_launch * b_pair / (compute_ms * 1e-3) / 1e9
        sgns = {
            "metric": "sgns_pairs_per_s", "value": job_pairs / (sg_total * 1e-3), "unit": "pairs/s",
            "ms_per_step": sg_total / args.steps, "steps": args.steps, "dtype": "f32",
            "gpu_launches": args.steps + 2 * n_sync,
            "config": {"dim": dim, **SGNS_HP, "walks_per_gpu": W, "tokens_per_walk": L + 1, "table_rows": n_rows,
                       "step": f"1/{slices} of the epoch on every rank", "sync": f"NCCL allreduce(sum) of both tables + x1/G "
                       f"every {sync_steps} steps, inside the timed region", "updates": "red.global.add.v4.f32",
                       "l2": "flushed between timed iterations; tables 2 x %.1f GB beyond L2" % (table_bytes / 2e9)},
            "allreduce": {"syncs_in_timed_region": n_sync, "bytes_per_sync": table_bytes,
                          "ms_per_sync": (sync_total / n_sync) if n_sync else None,
                          "share_of_timed_region": sync_total / sg_total if sg_total else None,
                          "algbw_GBps": (table_bytes / (sync_total / n_sync * 1e-3) / 1e9) if n_sync else None},
            "e2e": {"value": e2e_sg_pairs / e2e_sg_t, "unit": "pairs/s", "h2d_bytes_per_step": int(W * (L + 1) * 4),
                    "d2h_bytes_per_step": d2h, "seconds": e2e_sg_t,
                    "what": "Node2VecGensim(host walks, process_group).fit() on every rank: H2D of the rank's walks + "
                            "vocab (count allreduce) + init + ONE full epoch + table averaging, then rank 0 copies the "
                            "table to the host; one pass"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(name, "sgns_kernel"), "kernel": "sgns_kernel", "bytes_per_pair": b_pair,
                         "kernel_ms": compute_ms, "algorithmic_bytes_per_launch": per_launch * b_pair,
                         "pairs_per_launch": per_launch, "peak_source": peak_src,
                         "note": "per-GPU figure of rank 0 (kernel time without the allreduce)"},
            "vocab_s": t_vocab,
        }
    sync_all()
    if rank == 0:
        line = {
            "metric": "walk_steps_per_s", "value": value, "unit": "walk-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(name, w, world),
            "detail": {"vertices": V, "arcs": n_arcs_total,
                       "arcs_per_gpu": n_local_arcs, "graph_bytes_per_gpu": graph_bytes, "walkers_per_gpu": W,
                       "sharding": "vertex-partitioned CSR (rank r owns vertices [r*S, (r+1)*S) and their arc records / "
                                   "hash sets), peers read over NVLink with plain LDG.256; walkers stay on the rank of "
                                   "their start vertex; no collective in the walk",
                       "note": "strong scaling of ONE fixed graph over N = 2/4/8; the N = 1 line is a different workload "
                               "(configs[2])"},
            "multi_gpu_parity": bool(parity_walk and (parity_tables is not False)),
            "multi_gpu_parity_detail": {"partitioned_walks_equal_replicated": parity_walk,
                                        "tables_identical_across_ranks_after_averaging": parity_tables},
            "gpu_launches": args.steps + (sgns["gpu_launches"] if sgns else 0),
            "remote_hop_fraction": remote,
            "walk_localized": localized,
            "e2e": {"value": job_steps / float(np.median(e2e_times)), "unit": "walk-steps/s",
                    "h2d_bytes_per_step": int(n_local_arcs * 8), "d2h_bytes_per_step": int(W * (L + 1) * 4),
                    "seconds": float(np.median(e2e_times)),
                    "what": "fugue.random_walk(host arcs of this rank, process_group) on every rank = H2D + partitioned "
                            "build (csr/hash/alias, header all-gather, VMM handle exchange) + walk with D2H of the rank's "
                            "walk matrix pipelined under it; max over ranks, median of %d passes; bytes are per rank"
                            % args.e2e_passes},
            "roofline": walk_roofline(stats, steps_per_pass, kernel_ms, "nvlink_peer", name,
                                      "per-GPU figure of rank 0; %.0f%% of the gathers cross NVLink, so the bound is the "
                                      "peer random-sector ceiling (roofline.gather), not local HBM" % (100 * remote), remote),
            "clocks": clocks.summary(), "walk_stats": stats,
            "one_off_s": {"generate": t_gen, "partitioned_build": t_build},
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "total_s": time.perf_counter() - t_all,
        }
        if sgns:
            sgns.pop("gpu_launches")
            line["sgns"] = sgns
        emit(line)
    dist.barrier()
    dist.destroy_process_group()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--rmat-scale", type=int, default=0, help="shrink configs[4] (tests only; the default is scale 26)")
    ap.add_argument("--sgns-slices", type=int, default=8, help="N > 1: SGNS steps per epoch")
    ap.add_argument("--sync-steps", type=int, default=4, help="N > 1: average the tables every this many steps")
    ap.add_argument("--e2e-passes", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sgns", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-localized", action="store_true", help="N > 1: skip the replicated-from-parts walk block")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        args.workload = "rmat20" if max(args.gpus, world) == 1 else "rmat26"
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "rmat26" or world > 1:
        if args.workload != "rmat26":
            raise SystemExit("multi-GPU runs use the vertex-partitioned workload rmat26")
        return bench_partitioned(args)
    return bench_single(args)


if __name__ == "__main__":
    main()
