"""ctypes front-end of the C oracle (oracle/csrc/n2v_oracle.c).  TEST INFRASTRUCTURE.

``load()`` builds ``oracle/_build/libn2v_oracle.so`` with the Makefile when it is
missing (gcc only) and returns thin numpy wrappers.  Threaded variants fan ranges out
over Python threads (ctypes drops the GIL) because this image has no OpenMP runtime.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libn2v_oracle.so")

SUM_MODE = {"naive": 0, "neumaier": 1}
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, "csrc", f) for f in ("n2v_oracle.c", "sgns_ref.c")]
    stale = (not os.path.exists(SO)) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", HERE, "-B"], stdout=subprocess.DEVNULL)
    return SO


class WalkConsts(C.Structure):
    _fields_ = [("t_ret", C.c_uint64), ("t_nbr", C.c_uint64), ("t_far", C.c_uint64),
                ("fold_gain", C.c_float), ("fold_mode", C.c_int32), ("max_trials", C.c_int32),
                ("mix_qm1", C.c_float)]


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.orc_alias_tables_csr.restype = C.c_int64
        _lib.orc_reference_walk.restype = C.c_int64
    return _lib


def csr_from_arcs(src, dst, wt, n_vertices=None):
    """(src, dst)-sorted CSR with stable ties -- numpy twin of build_adjacency."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    wt = np.ones(len(src), dtype=np.float64) if wt is None else np.asarray(wt, dtype=np.float64)
    if n_vertices is None:
        n_vertices = int(max(src.max(initial=-1), dst.max(initial=-1)) + 1)
    order = np.lexsort((dst, src))  # stable
    deg = np.bincount(src, minlength=n_vertices)
    row_ptr = np.zeros(n_vertices + 1, dtype=np.int64)
    np.cumsum(deg, out=row_ptr[1:])
    return row_ptr, dst[order].astype(np.int32), wt[order].copy(), order


def alias_tables(weights, mode="naive"):
    lib = load()
    w = np.ascontiguousarray(weights, dtype=np.float64)
    alias = np.zeros(len(w), dtype=np.int32)
    probs = np.zeros(len(w), dtype=np.float64)
    rc = lib.orc_alias_tables(_ptr(w, C.c_double), C.c_int64(len(w)), SUM_MODE[mode],
                              _ptr(alias, C.c_int32), _ptr(probs, C.c_double))
    if rc != 0:
        raise ZeroDivisionError("float division by zero")
    return alias, probs


def alias_tables_csr(row_ptr, weights, mode="naive", threads=1):
    lib = load()
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    alias = np.zeros(len(w), dtype=np.int32)
    probs = np.zeros(len(w), dtype=np.float64)
    nv = len(row_ptr) - 1

    def run(rng):
        return lib.orc_alias_tables_csr(_ptr(row_ptr, C.c_int64), _ptr(w, C.c_double), C.c_int64(rng[0]),
                                        C.c_int64(rng[1]), SUM_MODE[mode], _ptr(alias, C.c_int32),
                                        _ptr(probs, C.c_double))
    bounds = [(nv * i // threads, nv * (i + 1) // threads) for i in range(threads)]
    if threads == 1:
        bad = run(bounds[0])
    else:
        with ThreadPoolExecutor(threads) as ex:
            bad = sum(ex.map(run, bounds))
    return alias, probs, int(bad)


def edge_alias_tables(row_ptr, col, w, prev, cur, p, q, mode="naive"):
    lib = load()
    n = int(row_ptr[cur + 1] - row_ptr[cur])
    alias = np.zeros(n, dtype=np.int32)
    probs = np.zeros(n, dtype=np.float64)
    rc = lib.orc_edge_alias_tables(_ptr(row_ptr, C.c_int64), _ptr(col, C.c_int32), _ptr(w, C.c_double),
                                   C.c_int32(prev), C.c_int32(cur), C.c_double(p), C.c_double(q),
                                   SUM_MODE[mode], _ptr(alias, C.c_int32), _ptr(probs, C.c_double))
    if rc == -2:
        raise ValueError(f"Zero return ({p}) or inout ({q}) parameter!")
    if rc != 0:
        raise ZeroDivisionError("float division by zero")
    return alias, probs


def mt_random(seed, n):
    lib = load()
    out = np.zeros(n, dtype=np.float64)
    lib.orc_mt_random(C.c_uint64(seed), C.c_int64(n), _ptr(out, C.c_double))
    return out


def reference_walk(row_ptr, col, w, start, num_walks, walk_length, p, q, mode="naive",
                   random_seed=None, threads=1):
    """The reference's walk (C port).  threads == 1 and a seed reproduce the Python oracle /
    the reference exactly; threads > 1 is for timing (one MT stream per range)."""
    lib = load()
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    start = np.ascontiguousarray(start, dtype=np.int32)
    W = len(start) * num_walks
    walks = np.empty((W, walk_length + 1), dtype=np.int32)
    alive = np.empty(W, dtype=np.uint8)
    max_deg = int(np.diff(row_ptr).max(initial=0))
    has_seed = random_seed is not None
    base_seed = int(random_seed) if has_seed else int.from_bytes(os.urandom(4), "little")

    def run(i):
        lo, hi = W * i // threads, W * (i + 1) // threads
        return lib.orc_reference_walk(
            _ptr(row_ptr, C.c_int64), _ptr(col, C.c_int32), _ptr(w, C.c_double), C.c_int64(max_deg),
            _ptr(start, C.c_int32), C.c_int64(len(start)), C.c_int32(num_walks), C.c_int32(walk_length),
            C.c_double(p), C.c_double(q), SUM_MODE[mode], C.c_int(1 if has_seed else 0),
            C.c_uint64(base_seed + 100 * i), C.c_int64(lo), C.c_int64(hi),
            _ptr(walks, C.c_int32), _ptr(alive, C.c_uint8))
    if threads == 1:
        run(0)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(run, range(threads)))
    return walks, alive.astype(bool)


def philox(key, ctr):
    lib = load()
    k = (C.c_uint32 * 2)(*key)
    c = (C.c_uint32 * 4)(*ctr)
    o = (C.c_uint32 * 4)()
    lib.orc_philox4x32_10(k, c, o)
    return list(o)


def walk_consts(p, q, graph_flags, has_ratio=False):
    lib = load()
    c = WalkConsts()
    if lib.orc_walk_consts_for(C.c_double(p), C.c_double(q), C.c_uint32(graph_flags), C.c_int(1 if has_ratio else 0),
                               C.byref(c)) != 0:
        raise ValueError(f"Zero return ({p}) or inout ({q}) parameter!")
    return c


def return_ratios(row_ptr, col, weight):
    """{fwd, rev} per arc as include/n2v_b200.h defines them for the general fold (small graphs:
    plain Python loops, sequential fp64 sums, fp32 division)."""
    n = len(row_ptr) - 1
    wsum = np.zeros(n, dtype=np.float32)
    tot = {}
    for v in range(n):
        s = 0.0
        for e in range(row_ptr[v], row_ptr[v + 1]):
            s = s + float(weight[e])
            key = (v, int(col[e]))
            tot[key] = tot.get(key, 0.0) + float(weight[e])
        wsum[v] = np.float32(s)
    out = np.zeros((len(col), 2), dtype=np.float32)
    for v in range(n):
        for e in range(row_ptr[v], row_ptr[v + 1]):
            x = int(col[e])
            out[e, 0] = np.float32(tot[(v, x)]) / wsum[v]
            back = tot.get((x, v), 0.0)
            out[e, 1] = np.float32(back) / wsum[x] if back > 0.0 else np.float32(0.0)
    return out


def replay_walk(base, deg, arc_thr, arc_dst, arc_alias_dst, col, weight, graph_flags, p, q, start,
                num_walks, walk_length, seed, pitch=None, threads=1, alias_idx=None, ratio=None):
    """Host replay of the device sampler (bit-exact twin of n2v_walk)."""
    lib = load()
    base = np.ascontiguousarray(base, dtype=np.uint64)
    deg = np.ascontiguousarray(deg, dtype=np.uint32)
    arc_thr = np.ascontiguousarray(arc_thr, dtype=np.uint32)
    arc_dst = np.ascontiguousarray(arc_dst, dtype=np.int32)
    arc_alias_dst = np.ascontiguousarray(arc_alias_dst, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    weight = np.ascontiguousarray(weight, dtype=np.float64)
    start = np.ascontiguousarray(start, dtype=np.int32)
    if pitch is None:
        pitch = (walk_length + 1 + 7) // 8 * 8
    W = len(start) * num_walks
    walks = np.empty((W, pitch), dtype=np.int32)
    alive = np.empty(W, dtype=np.uint8)
    consts = walk_consts(p, q, graph_flags, ratio is not None)
    stats = np.zeros((threads, 8), dtype=np.uint64)
    alias_idx = np.zeros(len(col), dtype=np.int32) if alias_idx is None else np.ascontiguousarray(alias_idx, dtype=np.int32)
    ratio_arr = None if ratio is None else np.ascontiguousarray(ratio, dtype=np.float32)
    ratio_ptr = C.c_void_p(0) if ratio_arr is None else ratio_arr.ctypes.data_as(C.c_void_p)

    def run(i):
        lo, hi = W * i // threads, W * (i + 1) // threads
        st = stats[i]
        lib.orc_replay_walk(
            _ptr(base, C.c_uint64), _ptr(deg, C.c_uint32), _ptr(arc_thr, C.c_uint32), _ptr(arc_dst, C.c_int32),
            _ptr(arc_alias_dst, C.c_int32), _ptr(alias_idx, C.c_int32), ratio_ptr, _ptr(col, C.c_int32),
            _ptr(weight, C.c_double), C.byref(consts),
            C.c_double(p), C.c_double(q), _ptr(start, C.c_int32), C.c_int64(len(start)), C.c_int32(num_walks),
            C.c_int32(walk_length), C.c_uint64(seed), C.c_int64(lo), C.c_int64(hi), _ptr(walks, C.c_int32),
            C.c_int64(pitch), _ptr(alive, C.c_uint8), _ptr(st, C.c_uint64))
    if threads == 1:
        run(0)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(run, range(threads)))
    names = ["steps", "trials", "probes", "searches", "fold_hits", "fallbacks", "dead", "reserved"]
    return walks, alive.astype(bool), dict(zip(names, stats.sum(axis=0).tolist()))


# ---------------------------------------------------------------------------------------
# SGNS (gensim 3.8 restatement, oracle/csrc/sgns_ref.c) -- PARITY UNPINNED, see the C header
# ---------------------------------------------------------------------------------------
def sgns_exp_table():
    lib = load()
    out = np.zeros(1000, dtype=np.float32)
    lib.orc_sgns_exp_table(_ptr(out, C.c_float))
    return out


def sgns_vocab(counts, min_count=5, sample=1e-3, ns_exponent=0.75):
    """prepare_vocab + make_cum_table.  Returns (keep_int[V] uint64, order[n] ids by
    descending count, cum_table[n] uint32)."""
    lib = load()
    lib.orc_sgns_vocab.restype = C.c_int64
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    V = len(counts)
    keep = np.zeros(V, dtype=np.uint64)
    order = np.zeros(V, dtype=np.int32)
    cum = np.zeros(V, dtype=np.uint32)
    n = lib.orc_sgns_vocab(_ptr(counts, C.c_int64), C.c_int64(V), C.c_int64(min_count), C.c_double(sample),
                           C.c_double(ns_exponent), _ptr(keep, C.c_uint64), _ptr(order, C.c_int32),
                           _ptr(cum, C.c_uint32))
    return keep, order[:n].copy(), cum[:n].copy()


def sgns_init(n_rows, dim, seed):
    """gensim reset_weights law: (U[0,1) - 0.5) / dim per component, zeros for syn1neg."""
    rng = np.random.RandomState(seed & 0xFFFFFFFF)
    syn0 = ((rng.rand(n_rows, dim) - 0.5) / dim).astype(np.float32)
    return syn0, np.zeros((n_rows, dim), dtype=np.float32)


def sgns_train(walks, counts, syn0, syn1neg, window=5, negative=5, alpha=0.025, min_alpha=1e-4, epochs=5,
               min_count=5, sample=1e-3, ns_exponent=0.75, batch_words=10000, seed=1, threads=1):
    """gensim-3.8 skip-gram negative sampling over a rectangular walk matrix (in-place on
    syn0 / syn1neg).  threads > 1 = lock-free workers on the same tables, like gensim."""
    lib = load()
    lib.orc_sgns_train.restype = C.c_int64
    walks = np.ascontiguousarray(walks, dtype=np.int32)
    W, length = walks.shape
    keep, order, cum = sgns_vocab(counts, min_count, sample, ns_exponent)
    D = syn0.shape[1]
    pairs = 0
    for ep in range(epochs):
        def run(i):
            lo, hi = W * i // threads, W * (i + 1) // threads
            return lib.orc_sgns_train(
                _ptr(walks, C.c_int32), C.c_int64(W), C.c_int64(length), C.c_int64(length), _ptr(keep, C.c_uint64),
                _ptr(order, C.c_int32), _ptr(cum, C.c_uint32), C.c_int64(len(order)), _ptr(syn0, C.c_float),
                _ptr(syn1neg, C.c_float), C.c_int64(D), C.c_int(window), C.c_int(negative), C.c_double(alpha),
                C.c_double(min_alpha), C.c_int(ep), C.c_int(epochs), C.c_int64(batch_words), C.c_uint64(seed),
                C.c_int64(lo), C.c_int64(hi))
        if threads == 1:
            pairs += run(0)
        else:
            with ThreadPoolExecutor(threads) as ex:
                pairs += sum(ex.map(run, range(threads)))
    return int(pairs)


def sgns_apply_trace(trace, alphas, n_pairs, K, syn0, syn1neg):
    lib = load()
    trace = np.ascontiguousarray(trace, dtype=np.int32)
    alphas = np.ascontiguousarray(alphas, dtype=np.float32)
    lib.orc_sgns_apply_trace(_ptr(trace, C.c_int32), C.c_int64(n_pairs), C.c_int(K), _ptr(alphas, C.c_float),
                             _ptr(syn0, C.c_float), _ptr(syn1neg, C.c_float), C.c_int64(syn0.shape[1]))
