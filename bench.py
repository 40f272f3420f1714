#!/usr/bin/env python
"""bench.py -- the node2vec hot path on B200: walk-steps/s (+ SGNS pairs/s).

Contract (see the task prompt): `python bench.py --gpus N --steps K --warmup W` prints ONE
JSON line on rank 0.  A "step" is one pass of the hot path over one batch of synthetic
input: every start vertex of the workload graph x num_walks walkers x walk_length steps
(then, when the SGNS half is built, one SGNS epoch over those walks).

  value     whole-job walk-steps/s with the graph already packed in HBM (kernel only)
  e2e       the same metric through the public API (node2vec_b200.fugue.random_walk) with
            HOST buffers: H2D of the arc list, CSR + alias build, walk, D2H of the walk matrix
  roofline  HBM model of the walk kernel: achieved = steps/s * B_step(T, l) with the trial
            and probe counts the kernel itself reports
  cpu_baseline  the reference's algorithm (oracle port, test infrastructure) on the host cores

`--impl reference` times the reference's own CPU path (oracle port) instead.
Multi-GPU: walkers shard by start vertex with a replicated graph, no data-path collective
(`"scaling": "weak"`: every rank walks the full per-GPU workload from its own start-vertex
shard of an N-times larger walker set).
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
    # BASELINE.json configs[1] -- the largest configuration quoted on ONE B200
    "blogcatalog_like": dict(graph="blogcatalog_like(10k vertices, 334k edges)", n=10000, m=334000,
                             p=0.25, q=4.0, num_walks=80, walk_length=40, dim=128),
    # BASELINE.json configs[2] -- power-law RMAT scale 20 (1M vertices, 16M edges): far larger than L2,
    # the DRAM-bound regime of the same kernels
    "rmat20": dict(graph="rmat(scale 20, edge factor 16)", n=1 << 20, m=16 << 20, p=0.25, q=4.0, num_walks=10,
                   walk_length=80, dim=128),
    # BASELINE.json configs[0] -- the reference's own CPU-runnable case
    "er_10k": dict(graph="erdos_renyi(10k vertices, 100k edges)", n=10000, m=100000,
                   p=1.0, q=0.5, num_walks=10, walk_length=20, dim=128),
}


def make_graph(name):
    from node2vec_b200 import synth
    w = WORKLOADS[name]
    if name == "blogcatalog_like":
        return synth.blogcatalog_like(w["n"], w["m"], seed=42)
    if name == "rmat20":
        import torch
        if torch.cuda.is_available():
            src, dst = synth.rmat_device(20, 16, seed=42)
            return src.cpu().numpy(), dst.cpu().numpy()
        return synth.rmat_host(20, 16, seed=42)
    return synth.erdos_renyi(w["n"], w["m"], seed=42)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, kernel):
    """DRAM bytes per launch measured by ncu for this workload/kernel (profiles/ncu_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload, {}).get(kernel)
    except OSError:
        return None


# measured on B200 with scripts/gather_peak.cu (profiles/r01_gather_peak.txt): random 32-byte
# LDG.256 gathers, G sectors/s
GATHER_CEILING = {"l2": 235.0, "hbm": 40.0}


def gather_view(stats, steps, kernel_ms, graph_bytes):
    """The walk as what it is -- a stream of random 32-byte sector gathers (one arc record per
    trial, one hash bucket per membership test): achieved G sectors/s against the measured
    random-gather ceiling of the level the graph lives in (L2 if it fits, else HBM)."""
    sectors = (stats["trials"] + stats["probes"]) / max(stats["steps"], 1) * steps
    achieved = sectors / (kernel_ms * 1e-3) / 1e9
    level = "l2" if graph_bytes < 100e6 else "hbm"
    return {"achieved_gsectors_per_s": achieved, "ceiling_gsectors_per_s": GATHER_CEILING[level], "level": level,
            "frac": achieved / GATHER_CEILING[level],
            "bytes_per_step_layout": 32.0 * (stats["trials"] + stats["probes"]) / max(stats["steps"], 1) + 4.0,
            "note": "ceiling = scripts/gather_peak.cu on this GPU type; hub sectors that hit L2 let a DRAM-sized graph exceed the HBM figure slightly"}


def walk_bytes_per_step(stats):
    """16 B vertex record + T * (16 B arc record + 4 B * probes) + 4 B store.
    (The arc record is 16 B here, not the 12 B of the survey's sketch.)"""
    steps = max(stats["steps"], 1)
    T = stats["trials"] / steps
    probes_per_trial = stats["probes"] / max(stats["trials"], 1)
    return 16.0 + T * (16.0 + 4.0 * probes_per_trial) + 4.0, T, probes_per_trial


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

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def resume(self):
        """Sample a second timed region into the same record."""
        self._stop = threading.Event()
        self.__enter__()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------
def cpu_baseline_walk(name, budget_s=12.0, procs=None):
    """The reference's algorithm on the host cores (oracle Python port: per walker per step
    it re-derives the biased weights and rebuilds the alias table, exactly like
    next_step_random_walk).  A bounded sample of the same workload: every process walks
    seeded start vertices of the same graph (2 walkers each, full walk_length) until the
    time budget is spent.  Adjacency construction is outside the timed region."""
    import multiprocessing as mp
    global _ADJ
    from oracle import ref_walk
    w = WORKLOADS[name]
    src, dst = make_graph(name)
    procs = procs or os.cpu_count() or 1
    _ADJ = ref_walk.build_adjacency(src.tolist(), dst.tolist(), [1.0] * len(src))
    rng = np.random.default_rng(0)
    starts = rng.permutation(np.unique(src))
    chunks = [c.tolist() for c in np.array_split(starts, procs) if len(c)]
    args = [(c, 2, w["walk_length"], w["p"], w["q"], 1000 + i, budget_s) for i, c in enumerate(chunks)]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(chunks)) as pool:
        res = pool.map(_cpu_walk_chunk, args)
    steps = int(sum(r[0] for r in res))
    dt = max(r[1] for r in res)
    n_starts = int(sum(r[2] for r in res))
    return {"value": steps / dt, "unit": "walk-steps/s", "cores": len(chunks), "kind": "port",
            "sample": f"{n_starts} start vertices x 2 walks x {w['walk_length']} steps = {steps} steps in {dt:.1f}s on "
                      f"{len(chunks)} processes (oracle/ref_walk.py: the reference's per-row algorithm, "
                      f"adjacency prebuilt, Fugue joins and pickle/base64 decoding not included)"}


def cpu_baseline_walk_c(name, budget_s=6.0):
    """Same algorithm, C port (oracle/csrc/n2v_oracle.c: per walker per step it re-derives the biased
    weights and rebuilds the alias table, like the reference) on all host cores -- context for how
    much of the Python baseline is interpreter overhead."""
    from oracle import clib
    w = WORKLOADS[name]
    src, dst = make_graph(name)
    row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, None, w["n"])
    starts = np.flatnonzero(np.diff(row_ptr) > 0).astype(np.int32)
    threads = os.cpu_count() or 1
    steps, dt, nw = 0, 0.0, 1
    while dt < budget_s and nw <= w["num_walks"]:
        t0 = time.perf_counter()
        walks, alive = clib.reference_walk(row_ptr, col, ws, starts, nw, w["walk_length"], w["p"], w["q"], "naive",
                                           None, threads=threads)
        dt += time.perf_counter() - t0
        steps += int(alive.sum()) * w["walk_length"]
        nw *= 2
    return {"value": steps / dt, "unit": "walk-steps/s", "cores": threads, "kind": "port",
            "sample": f"{steps} steps in {dt:.1f}s, {threads} threads (C port of the reference's per-row algorithm)"}


_ADJ = None


def _cpu_walk_chunk(a):
    import random
    from oracle import ref_walk
    starts, num_walks, L, p, q, seed, budget = a
    rng = random.Random(seed)
    steps = done = 0
    t0 = time.perf_counter()
    for v in starts:
        rows = ref_walk.start_rows([v], num_walks)
        for _ in range(L):
            rows = [ref_walk.step_row(r, _ADJ, p, q, rng.random(), rng.random()) for r in rows if r["dst"] in _ADJ]
        steps += len(rows) * L
        done += 1
        if time.perf_counter() - t0 > budget:
            break
    return steps, time.perf_counter() - t0, done


SGNS_HP = dict(window=5, negative=5, alpha=0.025, min_alpha=1e-4, min_count=1, sample=1e-3)


def sgns_bytes_per_pair(dim, negative):
    """(K+2) rows read + (K+2) rows written, D fp32 each, no reuse assumed."""
    return 8.0 * dim * (negative + 2)


def cpu_baseline_sgns(walks, n_rows, dim, budget_s=12.0, threads=None):
    """gensim-3.8 SGNS restatement (oracle/csrc/sgns_ref.c) with lock-free host threads on a
    bounded sample of the same walk matrix: as many leading walks as fit the time budget."""
    from oracle import clib
    threads = threads or os.cpu_count() or 1
    counts = np.bincount(walks.reshape(-1), minlength=n_rows)
    syn0, syn1 = clib.sgns_init(n_rows, dim, 1)
    n = min(len(walks), 2000 * threads)
    pairs, dt = 0, 0.0
    lo = 0
    while dt < budget_s and lo < len(walks):
        t0 = time.perf_counter()
        pairs += clib.sgns_train(walks[lo:lo + n], counts, syn0, syn1, epochs=1, seed=1, batch_words=10000,
                                 threads=threads, **SGNS_HP)
        dt += time.perf_counter() - t0
        lo += n
    return {"value": pairs / dt, "unit": "pairs/s", "cores": threads, "kind": "port",
            "sample": f"first {min(lo, len(walks))} walks x {walks.shape[1]} tokens, 1 epoch = {pairs} pairs in {dt:.1f}s, "
                      f"{threads} lock-free threads (oracle gensim-3.8 restatement; gensim itself is not installable)"}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) for the same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    w = WORKLOADS[name]
    vals = []
    base = None
    for _ in range(max(1, min(args.steps, 3))):
        base = cpu_baseline_walk(name)
        vals.append(base["value"])
    v = float(np.mean(vals))
    steps_per_pass = len(np.unique(make_graph(name)[0])) * w["num_walks"] * w["walk_length"]
    # SGNS leg: walks for the sample come from the oracle's C port of the reference walk
    from oracle import clib
    src, dst = make_graph(name)
    row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, None, w["n"])
    starts = np.flatnonzero(np.diff(row_ptr) > 0).astype(np.int32)
    walks_cpu, alive = clib.reference_walk(row_ptr, col, ws, starts, 4, w["walk_length"], w["p"], w["q"], "naive",
                                           None, threads=os.cpu_count() or 1)
    sgns_base = cpu_baseline_sgns(walks_cpu[alive], w["n"], w["dim"])
    line = {
        "impl": "reference", "metric": "walk_steps_per_s", "value": v, "unit": "walk-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * steps_per_pass / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, **{k: w[k] for k in ("graph", "p", "q", "num_walks", "walk_length")},
                   "note": "ms_per_step extrapolates the sampled rate to one full pass"},
        "cpu_baseline": {**base, "value": v},
        "e2e": {"value": v, "unit": "walk-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sgns": {"metric": "sgns_pairs_per_s", "value": sgns_base["value"], "unit": "pairs/s",
                 "cpu_baseline": sgns_base,
                 "e2e": {"value": sgns_base["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0,
                         "d2h_bytes_per_step": 0}},
    }
    emit(line)


# --------------------------------------------------------------------------------------
def bench_sgns(args, torch, dist, dev, world, rank, walks, w, flush, host_walks, clocks=None):
    """Timed region: `steps` SGNS epochs (one kernel launch each; + the NCCL model-averaging
    allreduce when world > 1) over the walk matrix already in HBM.  e2e: host walk matrix in,
    Node2VecGensim.fit() (H2D + vocab + tables + init + 1 epoch), host embeddings out."""
    from node2vec_b200.embedding import Node2VecGensim
    from node2vec_b200.sgns import Word2Vec
    dim, K = w["dim"], SGNS_HP["negative"]
    group = dist.group.WORLD if world > 1 else None
    m = Word2Vec(size=dim, sg=1, iter=args.steps + args.warmup, seed=1, batch_words=10000, process_group=group,
                 **SGNS_HP)
    m.build_vocab(walks)
    for _ in range(args.warmup):
        flush.fill_(1)
        m.train(walks, epochs=1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pairs = 0
    if clocks is not None:
        clocks.resume()
    for a, b in ev:
        flush.fill_(1)
        a.record()
        m.train(walks, epochs=1)       # one n2v_sgns_train launch (+ allreduce/scale when world > 1)
        b.record()
        pairs += m.train_stats["pairs"]
    torch.cuda.synchronize()
    if clocks is not None:
        clocks.__exit__()
    if world > 1:
        dist.barrier()
    ms = [a.elapsed_time(b) for a, b in ev]
    tot = torch.tensor([float(sum(ms)), float(pairs)], device=dev, dtype=torch.float64)
    if world > 1:
        t_max = tot[:1].clone()
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        p_sum = tot[1:].clone()
        dist.all_reduce(p_sum, op=dist.ReduceOp.SUM)
        total_ms, total_pairs = float(t_max.item()), float(p_sum.item())
    else:
        total_ms, total_pairs = float(tot[0].item()), float(tot[1].item())
    value = total_pairs / (total_ms * 1e-3)

    # end to end through the reference-shaped API, host buffers both ways (rank-local)
    def e2e_pass():
        n2v = Node2VecGensim(_HostWalks(host_walks), {"sg": 1, "iter": 1, "size": dim, **SGNS_HP}, random_seed=1)
        model = n2v.fit()
        vec = model.wv.vectors          # host numpy (D2H inside fit)
        return model.train_stats["pairs"], vec.nbytes
    e2e_pass()
    # median of 3 passes: a single pass occasionally stalls on the host for 100+ ms (allocator growth, GC)
    e2e_times, p_e2e = [], 0
    for _ in range(3):
        t0 = time.perf_counter()
        p_e2e, d2h = e2e_pass()
        e2e_times.append(time.perf_counter() - t0)
    print("sgns e2e pass times (ms):", [round(t * 1e3, 1) for t in e2e_times], file=sys.stderr)
    e2e_t = torch.tensor([float(np.median(e2e_times))], device=dev, dtype=torch.float64)
    e2e_p = torch.tensor([float(p_e2e)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_p, op=dist.ReduceOp.SUM)
    peak, peak_src = measured_peaks()
    b_pair = sgns_bytes_per_pair(dim, K)
    kernel_ms = float(np.mean(ms))
    achieved = (pairs / args.steps) * b_pair / (kernel_ms * 1e-3) / 1e9
    shared = None
    if world == 1 and dim <= 128 and K == 5:
        # informational: the opt-in window-shared-negatives kernel (csrc/sgns_shared.cu) on the same matrix.
        # A different sampling scheme (K negatives drawn once per centre), so it is NOT `value`.
        try:
            ms_ = Word2Vec(size=dim, sg=1, iter=4, seed=1, batch_words=10000, share_negatives=True, **SGNS_HP)
            ms_.build_vocab(walks)
            ms_.train(walks, epochs=1)
            torch.cuda.synchronize()
            t_sh, p_sh = [], 0
            for _ in range(3):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ms_.train(walks, epochs=1)
                b.record()
                torch.cuda.synchronize()
                t_sh.append(a.elapsed_time(b))
                p_sh = ms_.train_stats["pairs"]
            shared = {"value": p_sh / (float(np.mean(t_sh)) * 1e-3), "unit": "pairs/s", "ms_per_step": float(np.mean(t_sh)),
                      "steps": 3, "note": "opt-in Word2Vec(share_negatives=True): negatives drawn once per centre, "
                                          "target rows register-resident across the window; AUC within +-0.01 of the "
                                          "per-pair kernel (tests/test_gpu_sgns_shared.py); not the parity path"}
            del ms_
        except Exception as exc:      # never let the experiment disturb the bench line
            shared = {"unavailable": repr(exc)[:200]}
    return {
        "shared_negatives": shared,
        "metric": "sgns_pairs_per_s", "value": value, "unit": "pairs/s", "ms_per_step": total_ms / args.steps,
        "dtype": "f32", "gpu_launches": args.steps,
        "config": {"dim": dim, **SGNS_HP, "walks_per_gpu": int(walks.shape[0]), "tokens_per_walk": int(walks.shape[1]),
                   "updates": "red.global.add.v4.f32", "sync": "allreduce(avg) of both tables every epoch" if world > 1 else "none",
                   "l2": "flushed between timed iterations; tables (2 x %.1f MB) are L2-resident" % (w["n"] * dim * 4 / 1e6)},
        "e2e": {"value": float(e2e_p.item()) / float(e2e_t.item()), "unit": "pairs/s",
                "h2d_bytes_per_step": int(host_walks.numel() * 4), "d2h_bytes_per_step": int(d2h),
                "what": "Node2VecGensim(host walks).fit(): H2D + vocab_count + sgns_prepare + init + 1 epoch + D2H vectors; "
                        "median of 3 passes"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(args.workload, "sgns_kernel"), "kernel": "sgns_kernel",
                     "bytes_per_pair": b_pair, "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_launch": (pairs / args.steps) * b_pair,
                     "pairs_per_launch": pairs / args.steps, "peak_source": peak_src},
    }


class _IdFrame:
    """walk_seed stand-in: a frame with an `id` column (this rank's start vertices)."""

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


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="blogcatalog_like", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sgns", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from node2vec_b200 import fugue
    from node2vec_b200.graph import DeviceGraph
    name = args.workload
    w = WORKLOADS[name]
    src, dst = make_graph(name)
    g = DeviceGraph.from_arcs(src, dst, None, n_vertices=w["n"])
    from node2vec_b200 import dist as n2v_dist
    # weak scaling, sharded by start vertex: the N-GPU job walks num_walks * N walkers from every
    # start vertex; rank r owns the r-th contiguous shard of the start-vertex list, so every GPU
    # keeps ~the single-GPU walker count.  Philox is keyed by the global walk id: no collective.
    seed = 42
    nw = w["num_walks"] * world
    start = n2v_dist.shard_start_vertices(g.start_vertices(), rank, world)
    W = int(start.numel()) * nw
    steps_per_pass = W * w["walk_length"]
    pitch = (w["walk_length"] + 1 + 7) // 8 * 8
    out = torch.empty((W, pitch), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def one_pass():
        g.walk(start, nw, w["walk_length"], w["p"], w["q"], seed=seed, out=out)

    _, _, stats = g.walk(start, nw, w["walk_length"], w["p"], w["q"], seed=seed, collect_stats=True)
    for _ in range(args.warmup):
        flush.fill_(1)
        one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clocks = ClockSampler(local_rank)
    clocks.__enter__()                   # sampled through BOTH timed regions (walk, then SGNS)
    for a, b in ev:
        flush.fill_(1)                   # L2 flush between timed iterations (outside the events)
        a.record()
        one_pass()
        b.record()
    torch.cuda.synchronize()
    clocks.__exit__()                    # nvidia-smi polling perturbs host-driven work: not during the e2e passes
    if world > 1:
        dist.barrier()
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([float(sum(ms))], device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    tot_steps = torch.tensor([float(steps_per_pass)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot_steps, op=dist.ReduceOp.SUM)
    job_steps = float(tot_steps.item())
    value = job_steps / (ms_per_step * 1e-3)

    # ---- end to end through the public API, host buffers in, host walk matrix out
    src_pin = torch.as_tensor(src).pin_memory()
    dst_pin = torch.as_tensor(dst).pin_memory()
    host_out = torch.empty((W, w["walk_length"] + 1), dtype=torch.int32).pin_memory()
    params = {"num_walks": nw, "walk_length": w["walk_length"], "return_param": w["p"],
              "inout_param": w["q"]}
    seeds_df = _IdFrame(start.cpu().numpy()) if world > 1 else None

    def e2e_pass():
        # out=: rows land in the pinned host matrix, D2H of chunk k overlapped with the kernel of chunk k+1
        res = fugue.random_walk(None, (src_pin, dst_pin), dict(params), seeds_df, random_seed=seed, out=host_out)
        torch.cuda.synchronize()
        assert res.walks.shape == (W, w["walk_length"] + 1) and res.walks[0, 0] >= 0
        return res

    e2e_pass()
    if world > 1:
        dist.barrier()
    n_e2e = max(3, min(args.steps, 7))
    e2e_times = []
    for _ in range(n_e2e):
        t0 = time.perf_counter()
        e2e_pass()
        e2e_times.append(time.perf_counter() - t0)
    print("e2e pass times (ms):", [round(t * 1e3, 2) for t in e2e_times], file=sys.stderr)
    # median over passes: single passes occasionally stall on the host (allocator growth, GC) for tens of ms
    e2e_s = torch.tensor([float(np.median(e2e_times))], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = job_steps / float(e2e_s.item())

    # ---- SGNS half: one epoch of skip-gram negative sampling over this rank's walk matrix
    sgns = None
    if not args.no_sgns:
        sgns = bench_sgns(args, torch, dist, dev, world, rank, out[:, : w["walk_length"] + 1], w, flush, host_out,
                          clocks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    b_step, T, lp = walk_bytes_per_step(stats)
    kernel_ms = float(np.mean(ms))
    achieved = steps_per_pass * b_step / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": "walk_steps_per_s", "value": value, "unit": "walk-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": name, **{k: w[k] for k in ("graph", "p", "q", "num_walks", "walk_length")},
                   "walkers_per_gpu": W, "arcs": int(g.n_arcs), "l2": "flushed between timed iterations (256 MiB write)",
                   "sharding": "replicated CSR, walkers sharded by (start vertex, walk number); no collective"},
        "gpu_launches": args.steps + (sgns["gpu_launches"] if sgns else 0),
        "e2e": {"value": e2e_value, "unit": "walk-steps/s",
                "h2d_bytes_per_step": int(src_pin.numel() * 4 + dst_pin.numel() * 4),
                "d2h_bytes_per_step": int(host_out.numel() * 4),
                "what": "fugue.random_walk(host arcs) = H2D + csr/hash/alias build + walk, then D2H of the walk matrix; "
                        "median of %d passes (mean %.2f ms, median %.2f ms)" % (n_e2e, 1e3 * float(np.mean(e2e_times)),
                                                                              1e3 * float(np.median(e2e_times)))},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(name, "walk_kernel"), "kernel": "walk_kernel", "bytes_per_step": b_step,
                     "algorithmic_bytes_per_launch": steps_per_pass * b_step,
                     "sectors_per_step": (stats["trials"] + stats["probes"]) / max(stats["steps"], 1),
                     "gather": gather_view(stats, steps_per_pass, kernel_ms, g.nbytes()),
                     "trials_per_step": T, "probes_per_trial": lp, "kernel_ms": kernel_ms, "peak_source": peak_src,
                     "note": "graph (10 MB) is L2-resident: effective-bandwidth figure, see profiles/"},
        "clocks": clocks.summary(),
        "walk_stats": stats,
    }
    if sgns:
        sgns.pop("gpu_launches")
        line["sgns"] = sgns
    if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only (contract)
        line["cpu_baseline"] = cpu_baseline_walk(name)
        line["cpu_baseline_c"] = cpu_baseline_walk_c(name)
        if sgns:
            sample = host_out.numpy()
            line["sgns"]["cpu_baseline"] = cpu_baseline_sgns(sample, w["n"], w["dim"])
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
