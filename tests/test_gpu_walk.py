"""GPU parity tests for the walk half (K0 csr_build, K1 alias_build, a4/a5/a6, K2 walk).

Everything goes through the C ABI (ctypes -> libn2v_b200.so).  Integer/index results and
the fp64 alias tables are compared BIT-EXACTLY with the oracle and with the golden
fixtures produced by the unmodified reference; walk distributions are compared with the
reference's transition law by chi-square at significance 1e-4 per cell group (the stated
tolerance), and every GPU walk must equal its seeded host replay bit for bit.
"""
import random

import numpy as np
import pytest

from oracle import clib, ref_walk
from tests.helpers import chi_square_ok, graph_flags, load_golden, pack_arcs, unhex

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def n2v():
    import torch
    assert torch.cuda.is_available()
    from node2vec_b200 import _lib, fugue, graph, randomwalk
    _lib.load()

    class NS:
        pass
    ns = NS()
    ns.torch, ns.lib, ns.graph, ns.rw, ns.fugue = torch, _lib, graph, randomwalk, fugue
    return ns


def _random_arcs(rng, n, m, weighted=True, sym=False, multi=False):
    src = rng.integers(0, n, m)
    dst = rng.integers(0, n, m)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    if not multi:
        key = np.unique(src.astype(np.int64) << 32 | dst)
        src, dst = (key >> 32).astype(np.int64), (key & 0xFFFFFFFF).astype(np.int64)
        perm = rng.permutation(len(src))
        src, dst = src[perm], dst[perm]
    w = rng.uniform(0.1, 2.0, len(src)) if weighted else np.ones(len(src))
    if sym:
        a, b = np.minimum(src, dst), np.maximum(src, dst)
        key, idx = np.unique(a.astype(np.int64) << 32 | b, return_index=True)
        a, b, w = a[idx], b[idx], w[idx]
        src, dst, w = np.concatenate([a, b]), np.concatenate([b, a]), np.concatenate([w, w])
    return src, dst, w


# ---------------------------------------------------------------------------------- K0
@pytest.mark.parametrize("weighted,sym,multi", [(True, False, True), (False, True, False), (True, True, False),
                                                (False, False, False)])
def test_csr_build_matches_numpy(n2v, weighted, sym, multi):
    rng = np.random.default_rng(5)
    src, dst, w = _random_arcs(rng, 500, 6000, weighted, sym, multi)
    g = n2v.graph.DeviceGraph.from_arcs(src, dst, w if weighted else None, n_vertices=520, keep_perm=True)
    row_ptr, col, ws, order = clib.csr_from_arcs(src, dst, w, 520)
    h = g.to_host()
    deg = np.diff(row_ptr)
    assert h["deg"].tolist() == deg.tolist()
    assert h["base"][deg > 0].tolist() == row_ptr[:-1][deg > 0].tolist()
    assert h["col"].tolist() == col.tolist()
    assert h["weight"].view(np.uint64).tolist() == ws.view(np.uint64).tolist()
    assert g.perm.cpu().numpy().tolist() == order.tolist()          # stable ties
    assert g.flags == graph_flags(row_ptr, col, ws)
    assert g.start_vertices().cpu().numpy().tolist() == np.flatnonzero(deg > 0).tolist()


def test_csr_build_rejects_bad_ids(n2v):
    with pytest.raises(ValueError):
        n2v.graph.DeviceGraph.from_arcs([0, 1, 7], [1, 0, 2], None, n_vertices=5)


def test_start_ids_outside_the_graph_are_dropped(n2v):
    g = n2v.graph.DeviceGraph.from_arcs([0, 1, 2], [1, 2, 0], None, n_vertices=3)
    walks, alive, _ = g.walk(np.array([0, 7, -3, 2], dtype=np.int32), 2, 4, seed=1)
    assert alive.cpu().tolist() == [True, True, False, False, False, False, True, True]
    w = walks.cpu().numpy()
    assert (w[2:6, 1:] == -1).all() and w[0].tolist() == [0, 1, 2, 0, 1]


def test_empty_graph(n2v):
    g = n2v.graph.DeviceGraph.from_arcs(np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32), None, n_vertices=4)
    assert g.n_arcs == 0 and g.start_vertices().numel() == 0
    walks, alive, _ = g.walk(g.start_vertices(), 3, 4)
    assert walks.shape == (0, 5)


def test_hash_sets_answer_membership_exactly(n2v):
    """K0b: emulate the bucket lookup of include/n2v_b200.h on the host for members and
    non-members of every vertex (multi-arcs and a 30k hub included)."""
    rng = np.random.default_rng(9)
    src, dst, w = _random_arcs(rng, 3000, 50000, False, False, True)
    src = np.concatenate([src, np.full(30000, 11)]); dst = np.concatenate([dst, rng.integers(0, 3000, 30000)])
    g = n2v.graph.DeviceGraph.from_arcs(src, dst, None, n_vertices=3000)
    h = g.to_host()
    table = h["hash"]

    def lookup(t, x):
        deg = int(h["deg"][t])
        nb = (deg + 3) >> 2
        b = ((np.uint32(x) * np.uint32(0x9E3779B1)).astype(np.uint64) * nb) >> 32 if False else \
            (((int(x) * 0x9E3779B1) & 0xFFFFFFFF) * nb) >> 32
        n = 0
        while True:
            bucket = table[int(h["hbase"][t]) + b]
            n += 1
            if (bucket == x).any():
                return True, n
            if bucket[7] == -1:
                return False, n
            b = 0 if b + 1 == nb else b + 1
    nbrs = {}
    for a, b in zip(src.tolist(), dst.tolist()):
        nbrs.setdefault(a, set()).add(b)
    probes = 0
    count = 0
    for t in list(nbrs)[:400] + [11]:
        for x in list(nbrs[t])[:50]:
            ok, n = lookup(t, x)
            assert ok
            probes += n; count += 1
        for x in rng.integers(0, 3000, 50).tolist():
            ok, n = lookup(t, x)
            assert ok == (x in nbrs[t])
            probes += n; count += 1
    assert probes / count < 1.3
    # the hub (one CTA per vertex above 2048 arcs) and a few warp-built vertices: the table holds exactly the
    # distinct neighbours, each once, used slots contiguous from slot 0, and every member is found
    for t in [11] + list(nbrs)[:20]:
        nb = (int(h["deg"][t]) + 3) >> 2
        rows = table[int(h["hbase"][t]): int(h["hbase"][t]) + nb]
        used = rows != -1
        assert (used[:, :-1] | ~used[:, 1:]).all()
        assert sorted(rows[used].tolist()) == sorted(nbrs[t])
        assert all(lookup(t, x)[0] for x in nbrs[t])
    assert np.array_equal(h["hbase"], ((h["base"].astype(np.int64) >> 2) + np.arange(len(h["deg"]))).astype(np.uint32))


# ---------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("mode", ["naive", "neumaier"])
def test_alias_build_bit_exact_vs_oracle(n2v, mode):
    rng = np.random.default_rng(6)
    src, dst, w = _random_arcs(rng, 2000, 60000, True, False, True)
    w[rng.integers(0, len(w), 500)] = 1.0                    # exact ties / probs == 1.0 cases
    # hubs on every build path: 255/256 (thread vs CTA boundary), 5000, 12288/12289 (CTA vs global), 30000
    for v, deg in ((3, 255), (4, 256), (8, 5000), (9, 12288), (10, 12289), (7, 30000)):
        src = np.concatenate([src, np.full(deg, v)]); dst = np.concatenate([dst, rng.integers(0, 2000, deg)])
        w = np.concatenate([w, rng.pareto(1.5, deg) + 0.01])
    g = n2v.graph.DeviceGraph.from_arcs(src, dst, w, n_vertices=2000, sum_mode=mode, keep_tables=True)
    row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, w, 2000)
    alias, probs, bad = clib.alias_tables_csr(row_ptr, ws, mode, threads=4)
    assert bad == 0
    h = g.to_host()
    assert np.array_equal(h["alias"], alias)
    assert np.array_equal(h["probs"].view(np.uint64), probs.view(np.uint64))
    thr, adst, aalias = pack_arcs(row_ptr, col, alias, probs)
    assert np.array_equal(h["thr"], thr) and np.array_equal(h["dst"], adst)
    assert np.array_equal(h["alias_dst"], aalias) and np.array_equal(h["alias_idx"], alias)
    assert np.array_equal(h["dst_base"], h["base"][adst].astype(np.uint32)) and np.array_equal(h["dst_deg"], h["deg"][adst])
    assert np.array_equal(h["adst_base"], h["base"][aalias].astype(np.uint32))
    assert np.array_equal(h["adst_deg"], h["deg"][aalias])
    wsum = np.array([np.float32(ref_walk.float_sum(ws[a:b].tolist())) for a, b in zip(row_ptr[:-1], row_ptr[1:])],
                    dtype=np.float32)
    assert np.array_equal(h["wsum"][np.diff(row_ptr) > 0], wsum[np.diff(row_ptr) > 0])


@pytest.mark.parametrize("label", ["native", "naive"])
def test_alias_build_golden(n2v, label):
    """Every golden weight vector of the reference becomes one vertex of a graph."""
    fx = load_golden("alias_tables.json")
    mode = "naive" if label == "naive" else fx["native_sum_mode"]
    src, dst, w = [], [], []
    for i, case in enumerate(fx["cases"]):
        ws = unhex(case["weights"])
        src += [i] * len(ws); dst += list(range(len(ws))); w += ws
    nv = max(max(dst) + 1, len(fx["cases"]))
    g = n2v.graph.DeviceGraph.from_arcs(src, dst, w, n_vertices=nv, sum_mode=mode, keep_tables=True)
    h = g.to_host()
    for i, case in enumerate(fx["cases"]):
        b, n = int(h["base"][i]), int(h["deg"][i])
        assert h["alias"][b:b + n].tolist() == case[label]["alias"]
        assert [x.hex() for x in h["probs"][b:b + n].tolist()] == case[label]["probs"]


def test_zero_weight_vertex_raises(n2v):
    with pytest.raises(ValueError):
        n2v.graph.DeviceGraph.from_arcs([0, 0, 1], [1, 2, 0], [0.0, 0.0, 1.0], n_vertices=3)
    with pytest.raises(ZeroDivisionError):
        n2v.rw.generate_alias_tables([0.0, 0.0])
    with pytest.raises(ZeroDivisionError):
        n2v.rw.generate_alias_tables([])


# ---------------------------------------------------------------- a3/a4/a5/a6 via the facade
def test_reference_kats_through_facade(n2v):
    """The reference's own tests/test_randomwalk.py vectors, against node2vec_b200.randomwalk."""
    rw = n2v.rw
    for weights, alias, probs in [([0.5, 0.8, 1.0], [2, 0, 1], [0.6521739, 1.0, 0.9565217]),
                                  ([0.5, 0.2], [0, 0], [1.0, 0.5714285714285715]),
                                  ([0.2], [0], [1.0]), ([1.0], [0], [1.0])]:
        a, p = rw.generate_alias_tables(weights)
        assert a == alias
        np.testing.assert_almost_equal(p, probs, decimal=7)
    for prev, shared, nbrs, p_, q_, alias, probs in [
            (0, {2}, ([0, 2], [0.5, 0.2]), 1.0, 1.0, [0, 0], [1.0, 0.5714285714285715]),
            (1, set(), ([1], [0.2]), 0.8, 1.5, [0], [1.0]),
            (3, set(), ([1, 3], [0.5, 1.0]), 2.0, 4.0, [1, 0], [0.4, 1.0])]:
        a, p = rw.generate_edge_alias_tables(prev, shared, nbrs, p_, q_)
        assert a == alias
        np.testing.assert_almost_equal(p, probs, decimal=7)
        pytest.raises(ValueError, rw.generate_edge_alias_tables, prev, shared, nbrs, 0)
        pytest.raises(ValueError, rw.generate_edge_alias_tables, prev, shared, nbrs, 1.0, 0)
        pytest.raises(ValueError, rw.generate_edge_alias_tables, prev, shared, (nbrs[0], nbrs[1][:-1]))
    # AliasProb / RandomPath at seed 20 (tests/test_randomwalk.py:53-128)
    nbs = rw.Neighbors(([11, 22], [1.0, 0.5]))
    jq = rw.AliasProb(([1, 0], [0.6666666666666666, 1.0]))
    random.seed(20)
    r1 = random.random()
    assert nbs.dst_id[jq.sampling_from_alias_wiki(r1)] == 22
    assert nbs.dst_id[jq.sampling_from_alias(r1, random.random())] == 22
    for path, ids, alias, probs, result in [([-1, 0], [1, 3], [1, 0], [0.6666666666666666, 1.0], [0, 3]),
                                            ([2, 1], [0, 2], [0, 0], [1.0, 0.5714285714285715], [2, 1, 0]),
                                            ([0, 3], [0], [0], [1.0], [0, 3, 0])]:
        random.seed(20)
        assert rw.RandomPath(path).append(ids, rw.AliasProb((alias, probs)), random.random()).path == result


def test_edge_alias_tables_golden_bit_exact(n2v):
    fx = load_golden("edge_alias_tables.json")
    cs = fx["cases"]
    for label in ("native", "naive"):
        mode = "naive" if label == "naive" else fx["native_sum_mode"]
        for c in cs:  # p, q differ per case: one launch each
            a, p = n2v.rw.edge_alias_tables_batch([c["prev"]], [set(c["prev_out"])], [c["ids"]],
                                                  [unhex(c["weights"])], c["p"], c["q"], mode)[0]
            assert a == c[label]["alias"]
            assert [x.hex() for x in p] == c[label]["probs"]


def test_samplers_golden(n2v):
    fx = load_golden("samplers.json")
    d = fx["draws"]
    two = n2v.rw.alias_draw([c["alias"] for c in d], [unhex(c["probs"]) for c in d],
                            [float.fromhex(c["r1"]) for c in d], [float.fromhex(c["r2"]) for c in d])
    one = n2v.rw.alias_draw([c["alias"] for c in d], [unhex(c["probs"]) for c in d],
                            [float.fromhex(c["r1"]) for c in d], None)
    assert two.tolist() == [c["two"] for c in d]
    assert one.tolist() == [c["one"] for c in d]
    for c in fx["appends"]:
        r2 = None if c["r2"] is None else float.fromhex(c["r2"])
        got = n2v.rw.RandomPath(c["path"]).append(c["ids"], n2v.rw.AliasProb((c["alias"], unhex(c["probs"]))),
                                                  float.fromhex(c["r1"]), r2).path
        assert got == c["out"]


def test_next_step_random_walk_reference_rows(n2v):
    """tests/test_randomwalk.py:268-306 against the facade (seeded MT, device tables + draws)."""
    rw = n2v.rw
    dst_neighbors = [rw.Neighbors(([0, 2, 4], [0.5, 0.9, 1.0])).serialize(),
                     rw.Neighbors(([0, 3], [1.2, 0.9])).serialize(),
                     rw.Neighbors(([0, 3], [1.2, 0.9])).serialize()]
    src_neighbors = [rw.Neighbors(([2], [1.0])).serialize(), None, None]
    src, dst, path = [0, 0, -1], [1, 2, 2], [[3, 0, 1], [2, 0, 2], [-1, 2]]
    rows = [{"src": src[i], "dst": dst[i], "path": path[i], "dst_neighbors": dst_neighbors[i],
             "src_neighbors": src_neighbors[i]} for i in range(3)]
    for i, seed, want in [(0, 1000, (1, 4, [3, 0, 1, 4])), (1, 10, (2, 3, [2, 0, 2, 3])), (2, 20, (2, 3, [2, 3]))]:
        ans = list(rw.next_step_random_walk([rows[i]], 1.0, 1.0, seed))[0]
        assert (ans["src"], ans["dst"], ans["path"]) == want
    fx = load_golden("walks.json")["single_rows"]
    assert [(r["src"], r["dst"], r["path"]) for r in fx] == [(1, 4, [3, 0, 1, 4]), (2, 3, [2, 0, 2, 3]), (2, 3, [2, 3])]


def test_whole_reference_walks_through_row_facade(n2v):
    """Chain the facade's transformer-level functions exactly as fugue.random_walk does and
    reproduce the reference's seeded walks (golden, naive sum mode)."""
    rw = n2v.rw
    fx = load_golden("walks.json")
    for rec in fx["walks"][:4]:
        import pandas as pd
        df = pd.DataFrame({"src": rec["src"], "dst": rec["dst"], "weight": unhex(rec["weight"])})
        adj = {}
        for s, part in df.groupby("src", sort=True):
            part = part.sort_values("dst", kind="stable").reset_index(drop=True)
            row = next(iter(rw.get_vertex_neighbors(part)))
            adj[int(row["id"])] = row["neighbors"]
        starts = sorted(adj)
        if rec["walk_seed"] is not None:
            starts = [v for v in starts if v in set(rec["walk_seed"])]
        P = {"num_walks": 10, "walk_length": 20, "return_param": 1.0, "inout_param": 1.0, **rec["params"]}
        rows = [dict(r) for r in rw.initiate_random_walk([{"id": v} for v in starts], P["num_walks"])]
        for _ in range(P["walk_length"]):
            joined = [{"src": r["src"], "path": r["path"], "src_neighbors": adj.get(r["src"]),
                       "dst_neighbors": adj[r["dst"]]} for r in rows if r["dst"] in adj]
            rows = list(rw.next_step_random_walk(joined, P["return_param"], P["inout_param"], rec["random_seed"]))
        assert [r["walk"] for r in rw.to_path(rows)] == rec["naive"], rec["graph"]


# ---------------------------------------------------------------------------------- K2
def _replay(g, p, q, start, num_walks, L, seed, threads=4):
    h = g.to_host()
    ratio = None
    if g.ratio is not None:      # general fold in use: ratios recomputed independently on the host
        row_ptr = np.concatenate([[0], np.cumsum(h["deg"].astype(np.int64))])
        ratio = clib.return_ratios(row_ptr, h["col"], h["weight"])
        assert np.array_equal(ratio.view(np.uint32), h["ratio"].view(np.uint32))
    return clib.replay_walk(h["base"], h["deg"], h["thr"], h["dst"], h["alias_dst"], h["col"], h["weight"],
                            g.flags, p, q, start, num_walks, L, seed, threads=threads, alias_idx=h["alias_idx"],
                            ratio=ratio)


WALK_CASES = [
    # n, m, weighted, sym, multi, p, q, num_walks, L
    (300, 3000, True, False, True, 1.0, 1.0, 3, 9),
    (300, 3000, False, True, False, 1.0, 0.5, 4, 20),
    (300, 3000, False, True, False, 0.25, 4.0, 4, 40),      # fold path
    (300, 3000, True, True, False, 0.25, 4.0, 2, 17),       # weighted symmetric: general fold (per-arc ratios)
    (300, 3000, True, False, True, 0.2, 1.0, 3, 25),         # weighted directed multi-arc: general fold, sinks
    (300, 2000, True, False, False, 4.0, 0.25, 2, 8),       # directed with sinks
    (50, 400, False, True, False, 100.0, 1000.0, 6, 5),      # fallback scans
    (2000, 40000, False, True, False, 0.5, 2.0, 2, 80),
    (300, 3000, False, True, False, 0.25, 1.0, 4, 30),      # unit symmetric, q = 1: rho = 1/deg fold (mode 1)
    (300, 3000, False, True, False, 0.2, 0.5, 4, 30),       # unit symmetric, p < q < 1: mode 1 with lookups
    (300, 3000, False, True, False, 4.0, 2.0, 4, 30),       # mixture sampler, p > q > 1: x == prev thinned in the bulk
    (400, 900, False, True, False, 1.0, 8.0, 6, 30),        # mixture sampler on a sparse graph (deg ~ 4): common-neighbour proposals mostly fail
]


@pytest.mark.parametrize("case", WALK_CASES)
def test_walk_equals_host_replay(n2v, case):
    n, m, weighted, sym, multi, p, q, nw, L = case
    rng = np.random.default_rng(hash(case) % 2 ** 31)
    src, dst, w = _random_arcs(rng, n, m, weighted, sym, multi)
    g = n2v.graph.DeviceGraph.from_arcs(src, dst, w if weighted else None, n_vertices=n + 3)
    need_ratio = (g.flags & 7) != 7 and 1.0 / p > max(1.0, 1.0 / q)
    consts = n2v.graph.walk_consts(p, q, g.flags, need_ratio)
    oc = clib.walk_consts(p, q, g.flags, need_ratio)
    assert (consts.t_ret, consts.t_nbr, consts.t_far, consts.fold_mode, consts.max_trials) == \
        (oc.t_ret, oc.t_nbr, oc.t_far, oc.fold_mode, oc.max_trials)
    assert np.float32(consts.fold_gain) == np.float32(oc.fold_gain)
    start = g.start_vertices()
    walks, alive, stats = g.walk(start, nw, L, p, q, seed=4242, collect_stats=True)
    rw, ra, rs = _replay(g, p, q, start.cpu().numpy(), nw, L, 4242)
    full = walks._base if walks._base is not None else walks
    assert np.array_equal(alive.cpu().numpy(), ra)
    assert np.array_equal(full.cpu().numpy(), rw)                       # whole pitch-padded matrix, -1 padding too
    for k in ("steps", "trials", "searches", "fold_hits", "fallbacks", "dead"):
        assert stats[k] == rs[k], k
    assert stats["searches"] <= stats["probes"] <= 1.25 * stats["searches"] + 8     # ~1 bucket per lookup
    # stats-free launch (the timed variant) gives the same walks
    walks2, alive2, _ = g.walk(start, nw, L, p, q, seed=4242, collect_stats=False)
    assert n2v.torch.equal(walks, walks2) and n2v.torch.equal(alive, alive2)
    assert np.float32(consts.mix_qm1) == np.float32(oc.mix_qm1)
    if not weighted and sym and not multi:
        assert consts.fold_mode == (3 if q > 1.0 else 1 if p < min(1.0, q) else 0)
    if case[5] in (0.25, 0.2) and not weighted:
        assert stats["fold_hits"] > 0
    if need_ratio:
        assert consts.fold_mode == 2 and g.ratio is not None and stats["fold_hits"] > 0
    if p == 100.0:
        assert stats["fallbacks"] > 0
    if not sym:
        assert stats["dead"] == int((~ra).sum())


def _pair_counts(walks, pos):
    out = {}
    for row in walks:
        key = (int(row[pos - 1]), int(row[pos]))
        d = out.setdefault(key, {})
        d[int(row[pos + 1])] = d.get(int(row[pos + 1]), 0) + 1
    return out


@pytest.mark.parametrize("p,q,weighted,sym", [(1.0, 1.0, True, False), (1.0, 0.5, False, True),
                                              (0.25, 4.0, False, True), (4.0, 0.25, True, True),
                                              (0.25, 4.0, True, False)])
def test_walk_frequencies_match_reference_law(n2v, p, q, weighted, sym):
    """chi-square (alpha = 1e-4) of device transition frequencies against the reference's
    law, i.e. the normalised biased weights fed to generate_alias_tables."""
    rng = np.random.default_rng(21)
    src, dst, w = _random_arcs(rng, 14, 80, weighted, sym, False)
    g = n2v.graph.DeviceGraph.from_arcs(src, dst, w if weighted else None, n_vertices=14)
    adj = ref_walk.build_adjacency(src.tolist(), dst.tolist(), w.tolist())
    start = g.start_vertices()
    walks, alive, _ = g.walk(start, 20000, 3, p, q, seed=77)
    walks = walks[alive].cpu().numpy()
    for v in start.cpu().numpy()[:5]:
        law = ref_walk.transition_law(adj, None, int(v), p, q)
        rows = walks[walks[:, 0] == v]
        ids = sorted(law)
        ok, pval = chi_square_ok([(rows[:, 1] == x).sum() for x in ids], [law[x] for x in ids])
        assert ok, (v, pval)
    for pos in (1, 2):
        table = _pair_counts(walks, pos)
        for (t, v) in sorted(table, key=lambda k: -sum(table[k].values()))[:15]:
            law = ref_walk.transition_law(adj, t, v, p, q)
            ids = sorted(law)
            assert set(table[(t, v)]) <= set(ids)
            ok, pval = chi_square_ok([table[(t, v)].get(x, 0) for x in ids], [law[x] for x in ids])
            assert ok, (t, v, pval)


def test_walk_frequencies_match_reference_sampler_empirically(n2v):
    """Same (t, v) pair, same graph: device frequencies vs frequencies of the REFERENCE's
    sampler (oracle C port of next_step_random_walk, MT stream) -- two-sample chi-square."""
    from scipy import stats as sst
    rng = np.random.default_rng(8)
    src, dst, w = _random_arcs(rng, 10, 50, True, False, False)
    row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, w, 10)
    g = n2v.graph.DeviceGraph.from_arcs(src, dst, w, n_vertices=10)
    start = g.start_vertices().cpu().numpy()
    gw, ga, _ = g.walk(start, 30000, 2, 0.5, 2.0, seed=5)
    gw = gw[ga].cpu().numpy()
    rw, ra = clib.reference_walk(row_ptr, col, ws, start, 30000, 2, 0.5, 2.0, "naive", None)
    rw = rw[ra]
    tg, tr = _pair_counts(gw, 1), _pair_counts(rw, 1)
    for key in sorted(tg, key=lambda k: -sum(tg[k].values()))[:10]:
        ids = sorted(set(tg[key]) | set(tr.get(key, {})))
        a = np.array([tg[key].get(x, 0) for x in ids]); b = np.array([tr[key].get(x, 0) for x in ids])
        keep = (a + b) >= 10
        if keep.sum() < 2:
            continue
        _, pval, _, _ = sst.chi2_contingency(np.vstack([a[keep], b[keep]]))
        assert pval > 1e-4, (key, pval)


# ------------------------------------------------------------------ fugue-level facade
def test_random_walk_facade_reference_graph(n2v):
    """tests/test_fugue.py:58-76 of the reference, plus what it leaves unchecked."""
    import pandas as pd
    graph = [[0, 2, 0.41], [0, 4, 0.85], [3, 4, 0.36], [2, 0, 0.68], [4, 0, 0.1], [4, 3, 0.37]]
    df = pd.DataFrame(graph, columns=["src", "dst", "weight"]).astype({"src": int, "dst": int})
    params = {"num_walks": 2, "walk_length": 3, "return_param": 0.5}
    res = n2v.fugue.random_walk(None, df, params, random_seed=3)
    assert res is not None and params["inout_param"] == 1.0            # defaults merged in place
    out = res.as_pandas()
    assert list(out.columns) == ["src", "walk"] and len(out) == 4 * 2
    arcs = {(a, b) for a, b, _ in graph}
    for s, walk in zip(out["src"], out["walk"]):
        assert len(walk) == 4 and walk[0] == s
        assert all((a, b) in arcs for a, b in zip(walk[:-1], walk[1:]))
    seeds = pd.DataFrame({"id": [0, 4, 99]})
    res = n2v.fugue.random_walk(None, n2v.fugue.Frame(df), params, seeds, random_seed=3)
    assert sorted(set(res.as_pandas()["src"])) == [0, 4] and res.count() == 4
    # duplicated seed ids multiply walkers, rows in the inner join's order (tests/test_fugue.py:73-75;
    # tests/golden/fugue_verbatim.json records the unmodified reference doing so)
    dup = pd.DataFrame({"id": [0, 0, 3, 2, 4, 4, 4]})
    r1 = n2v.fugue.random_walk(None, df, dict(params), dup, random_seed=3).walks
    assert r1[:, 0].tolist() == [0] * 4 + [2] * 2 + [3] * 2 + [4] * 6
    r2 = n2v.fugue.random_walk(None, df, dict(params), dup, random_seed=3).walks
    assert np.array_equal(r1, r2)
    one = n2v.fugue.random_walk(None, df, dict(params), pd.DataFrame({"id": [0, 2, 3, 4]}), random_seed=3).walks
    assert np.array_equal(r1[[0, 1, 4, 5, 6, 7, 8, 9]], one)           # layer 0 is the plain run
    big = n2v.fugue.random_walk(None, df, {"num_walks": 200, "walk_length": 6}, pd.DataFrame({"id": [4, 4]}),
                                random_seed=5).walks
    assert big.shape[0] == 400 and not np.array_equal(big[:200], big[200:])   # copies are independent walkers
    with pytest.raises(ValueError):
        n2v.fugue.random_walk(None, df, params, df)                     # walk_seed without "id"
    with pytest.raises(ValueError):
        n2v.fugue.random_walk(None, df, {"return_param": 0})
    # determinism and seed sensitivity
    a = n2v.fugue.random_walk(None, df, dict(params), random_seed=9).walks
    b = n2v.fugue.random_walk(None, df, dict(params), random_seed=9).walks
    c = n2v.fugue.random_walk(None, df, {"num_walks": 50, "walk_length": 8}, random_seed=10).walks
    d = n2v.fugue.random_walk(None, df, {"num_walks": 50, "walk_length": 8}, random_seed=11).walks
    assert np.array_equal(a, b) and not np.array_equal(c, d)


def test_random_walk_pipelined_host_delivery(n2v):
    """out= (pinned host matrix): chunked walk with overlapped D2H gives the rows of one launch,
    also when walkers die at sinks, and WalkFrame.walks is a view of the caller's buffer."""
    import pandas as pd
    torch = n2v.torch
    rng = np.random.default_rng(12)
    for n_src, n_dst in ((3000, 3000), (1500, 3000)):                    # second graph: ids >= 1500 are sinks
        src, dst = rng.integers(0, n_src, 20000), rng.integers(0, n_dst, 20000)
        df = pd.DataFrame({"src": src, "dst": dst, "weight": rng.uniform(0.1, 2.0, 20000)})
        prm = {"num_walks": 7, "walk_length": 11, "return_param": 0.5, "inout_param": 2.0}
        want = n2v.fugue.random_walk(None, df, dict(prm), random_seed=4).walks
        n_start = len(set(src.tolist()))
        host = torch.empty((n_start * 7, 12), dtype=torch.int32).pin_memory()
        host.fill_(-7)
        g = n2v.graph.DeviceGraph.from_arcs(src, dst, df["weight"].to_numpy())
        from node2vec_b200.graph import walk_to_host
        wd, alive = walk_to_host(g, g.start_vertices(), 7, 11, 0.5, 2.0, 4, host, chunk_walkers=4096)   # many chunks
        assert np.array_equal(host.numpy()[alive.cpu().numpy()], want) and np.array_equal(wd[alive].cpu().numpy(), want)
        host.fill_(-7)
        res = n2v.fugue.random_walk(None, df, dict(prm), random_seed=4, out=host)
        assert np.array_equal(res.walks, want) and np.array_equal(res.walks_device.cpu().numpy(), want)
        assert res.walks.ctypes.data == host.numpy().ctypes.data        # no extra host copy
        assert len(want) < n_start * 7                                  # some walkers died at sinks
        seeds = pd.DataFrame({"id": np.arange(0, n_src, 2)})
        a = n2v.fugue.random_walk(None, df, dict(prm), seeds, random_seed=4, out=host).walks.copy()
        b = n2v.fugue.random_walk(None, df, dict(prm), seeds, random_seed=4).walks
        assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        n2v.fugue.random_walk(None, df, dict(prm), random_seed=4, out=torch.empty((n_start * 7, 12), dtype=torch.int32))
    with pytest.raises(ValueError):
        n2v.fugue.random_walk(None, df, dict(prm), random_seed=4, out=host[:10])


def test_plain_c_consumer():
    """examples/c_consumer.c: CSR + hash + alias build and a walk driven from plain C through the C
    ABI only (no Python, no torch in the process); it validates every hop on the host itself."""
    import subprocess
    from node2vec_b200 import build as nb
    exe = nb.build_c_consumer()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    assert out.stdout.startswith("c_consumer OK abi=7") and "dropped_at_sink=" in out.stdout
    fields = dict(kv.split("=") for kv in out.stdout.split()[2:])
    assert int(fields["walkers"]) == 999 * 4 and int(fields["alive"]) + int(fields["dropped_at_sink"]) == 999 * 4
    assert int(fields["dropped_at_sink"]) > 0 and int(fields["alive"]) > 0


def test_random_walk_drops_walkers_at_sinks(n2v):
    import pandas as pd
    fx = [r for r in load_golden("walks.json")["walks"] if r["graph"] == "sink_multi"][0]
    df = pd.DataFrame({"src": fx["src"], "dst": fx["dst"], "weight": unhex(fx["weight"])})
    res = n2v.fugue.random_walk(None, df, {"num_walks": 200, "walk_length": 5, "return_param": 2.0,
                                           "inout_param": 0.5}, random_seed=1)
    w = res.walks
    assert 0 < len(w) < 4 * 200
    has_out = set(fx["src"])
    assert all(all(int(v) in has_out for v in row[:-1]) for row in w)   # only the last vertex may be a sink
    assert (w >= 0).all()


def test_trim_index_facade(n2v):
    import pandas as pd
    graph = [[0, 2, 0.41], [0, 4, 0.85], [3, 4, 0.36], [2, 0, 0.68], [4, 0, 0.1], [4, 3, 0.37]]
    df = pd.DataFrame(graph, columns=["src", "dst", "weight"]).astype({"src": int, "dst": int})
    r, name_id = n2v.fugue.trim_index(None, df, indexed=True)
    assert len(r.as_pandas()) == 6 and name_id is None
    r, name_id = n2v.fugue.trim_index(None, df, indexed=True, max_out_deg=1)
    assert len(r.as_pandas()) == 4 and name_id is None
    d1 = pd.DataFrame({"src": ["a1", "a1", "a1", "a2", "b2"], "dst": ["a2", "b1", "b2", "b1", "a2"]})
    r, name_id = n2v.fugue.trim_index(None, d1, indexed=False)
    assert len(r.as_pandas()) == 5 and len(name_id.as_pandas()) == 4
    d2 = pd.DataFrame({"dst": ["a2", "b1", "b2", "a1"], "weight": [0.8, 1.1, 1.0, 0.3]})
    with pytest.raises(ValueError):
        n2v.fugue.trim_index(None, d2, False)
    fx = load_golden("trim.json")
    big = pd.DataFrame({"src": fx["src"], "dst": fx["dst"], "weight": unhex(fx["weight"])})
    for c in fx["cases"]:
        if c["random_seed"] is None and c["max_out_degree"] > 0:
            continue
        r, _ = n2v.fugue.trim_index(None, big, indexed=True, max_out_deg=c["max_out_degree"],
                                    random_seed=c["random_seed"])
        out = r.as_pandas()
        assert out["src"].tolist() == c["src"] and out["dst"].tolist() == c["dst"]
        assert [float(x).hex() for x in out["weight"]] == c["weight"]


# ---------------------------------------------- BASELINE-size properties (configs[1] shape)
def test_full_size_properties_config2(n2v):
    """10k vertices / ~334k edges, p=0.25 q=4, 80 walks x 40: every hop is an arc, rows start
    where they should, nothing dies on a symmetric graph, results are seed-deterministic."""
    torch = n2v.torch
    from node2vec_b200 import synth
    src, dst = synth.blogcatalog_like(seed=42)
    g = n2v.graph.DeviceGraph.from_arcs(src, dst, None, n_vertices=10000)
    need = 1 | 2 | 4
    assert g.flags & need == need
    start = g.start_vertices()
    walks, alive, stats = g.walk(start, 80, 40, 0.25, 4.0, seed=42, collect_stats=True)
    assert bool(alive.all()) and walks.shape == (start.numel() * 80, 41)
    assert torch.equal(walks[:, 0], start.repeat_interleave(80))
    keys = (torch.as_tensor(src, device=walks.device).long() << 32) | torch.as_tensor(dst, device=walks.device).long()
    keys = torch.sort(keys).values
    hop = (walks[:, :-1].long() << 32) | walks[:, 1:].long()
    pos = torch.searchsorted(keys, hop.reshape(-1)).clamp(max=keys.numel() - 1)
    assert bool((keys[pos] == hop.reshape(-1)).all())
    assert stats["steps"] == walks.shape[0] * 40 and stats["fallbacks"] == 0
    w2, _, _ = g.walk(start, 80, 40, 0.25, 4.0, seed=42)
    assert torch.equal(walks, w2)
    sub, _, _ = g.walk(start[100:140], 80, 40, 0.25, 4.0, seed=42)      # sharding-invariant rows
    assert torch.equal(sub, walks[100 * 80:140 * 80])
    rw, ra, rs = _replay(g, 0.25, 4.0, start[:50].cpu().numpy(), 80, 40, 42)
    assert np.array_equal(rw[:, :41], walks[:50 * 80].cpu().numpy())


def test_device_trim_and_symmetrise(n2v):
    """preprocess.py on tensors: the reference's trimming law and undirected expansion."""
    torch = n2v.torch
    from node2vec_b200 import preprocess
    rng = np.random.default_rng(12)
    src = np.concatenate([rng.integers(0, 50, 400), np.full(5000, 7), np.full(300, 9)])
    dst = rng.integers(0, 6000, len(src))
    w = rng.uniform(0.1, 2.0, len(src))
    ts, td, tw = (torch.as_tensor(x, device="cuda") for x in (src, dst, w))
    s2, d2, w2 = preprocess.trim_hotspots_device(ts, td, tw, max_out_deg=100, seed=3)
    deg = np.bincount(s2.cpu().numpy(), minlength=50)
    want = np.minimum(np.bincount(src, minlength=50), 100)
    assert deg.tolist() == want.tolist()
    kept = set(zip(s2.cpu().tolist(), d2.cpu().tolist(), w2.cpu().tolist()))
    assert kept <= set(zip(src.tolist(), dst.tolist(), w.tolist()))           # a sub-multiset of the input
    a = preprocess.trim_hotspots_device(ts, td, tw, 100, seed=3)[1]
    b = preprocess.trim_hotspots_device(ts, td, tw, 100, seed=4)[1]
    assert torch.equal(a, d2) and not torch.equal(a, b)                        # seeded
    # uniformity of the sample: each of vertex 7's 5000 arcs survives with probability 100/5000
    pos7 = np.flatnonzero(src == 7)
    hits = np.zeros(len(pos7))
    for seed in range(60):
        keep_d = preprocess.trim_hotspots_device(ts[pos7], td[pos7], tw[pos7], 100, seed=seed)[2].cpu().numpy()
        hits += np.isin(w[pos7], keep_d)
    assert abs(hits.mean() - 60 * 100 / len(pos7)) < 1e-9 and hits.max() <= 12 and (hits > 0).mean() > 0.6
    # pass-through and the reference's "<= 0 means 100000"
    assert preprocess.trim_hotspots_device(ts, td, tw, 0)[0].numel() == len(src)
    # symmetrise: same triples as the pandas indexer's undirected expansion
    import pandas as pd
    from node2vec_b200.indexer import index_graph_pandas
    g = pd.DataFrame({"src": [0, 1, 2, 2, 3], "dst": [1, 0, 3, 3, 2], "weight": [1.0, 1.0, 0.5, 0.5, 0.7]})
    e, _ = index_graph_pandas(g.copy(), False)
    s3, d3, w3 = preprocess.symmetrise_device(*(torch.as_tensor(g[c].to_numpy(), device="cuda") for c in ("src", "dst", "weight")))
    # ids here are already 0..3 and first-occurrence order maps them to themselves except none: compare as sets
    name = {v: k for k, v in enumerate([0, 1, 2, 3])}
    assert set(zip(s3.cpu().tolist(), d3.cpu().tolist(), w3.cpu().tolist())) == \
        set(zip(g["src"].tolist() + g["dst"].tolist(), g["dst"].tolist() + g["src"].tolist(), g["weight"].tolist() * 2))
    assert len(s3) == len(e)
    out, none = n2v.fugue.trim_index(None, (ts, td, tw), indexed=True, directed=True, max_out_deg=100, random_seed=3)
    assert none is None and torch.equal(out[1], d2)


def test_device_trim_bit_exact_with_reference_sampler(n2v):
    """K5 n2v_trim_sample: the kept positions are numpy's RandomState(seed).permutation(deg)[:cap]
    (what pandas' DataFrame.sample evaluates to), and the tensor path of trim_index returns the same
    rows in the same order as the pandas path, which the reference goldens pin."""
    import ctypes as C
    import pandas as pd
    torch = n2v.torch
    lib = n2v.lib.load()
    for seed, cap in ((0, 1), (5, 7), (2 ** 32 - 1, 100)):
        degs = [cap + 1, 2 * cap + 3, 1000, 1000, 4097, 70001]
        deg = torch.as_tensor(degs, dtype=torch.int64, device="cuda")
        off = torch.cumsum(deg, 0) - deg
        scratch = torch.empty(int(deg.sum()), dtype=torch.int32, device="cuda")
        picked = torch.full((len(degs), cap), -1, dtype=torch.int32, device="cuda")
        n2v.lib.check(lib.n2v_trim_sample(n2v.lib.ptr(deg), n2v.lib.ptr(off), len(degs), cap, C.c_uint32(seed),
                                          n2v.lib.ptr(scratch), n2v.lib.ptr(picked), n2v.lib.current_stream_ptr()))
        got = picked.cpu().numpy()
        for h, d in enumerate(degs):
            assert got[h].tolist() == np.random.RandomState(seed).permutation(d)[:cap].tolist(), (seed, cap, d)
    # the WHOLE permutation (cap >= deg), from two elements to a BASELINE-configs[2] hotspot (2^18 arcs): every
    # batch shape of the warp-parallel shuffle -- MT refill boundaries, clashing partners, the last few indices
    degs = [2, 3, 5, 33, 624, 625, 1249, 20011, (1 << 18) + 1]
    cap = max(degs)
    deg = torch.as_tensor(degs, dtype=torch.int64, device="cuda")
    off = torch.cumsum(deg, 0) - deg
    for seed in (0, 1, 123456789):
        scratch = torch.empty(int(deg.sum()), dtype=torch.int32, device="cuda")
        picked = torch.full((len(degs), cap), -1, dtype=torch.int32, device="cuda")
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        n2v.lib.check(lib.n2v_trim_sample(n2v.lib.ptr(deg), n2v.lib.ptr(off), len(degs), cap, C.c_uint32(seed),
                                          n2v.lib.ptr(scratch), n2v.lib.ptr(picked), n2v.lib.current_stream_ptr()))
        t1.record()
        got = picked.cpu().numpy()
        print(f"n2v_trim_sample, largest vertex 2^18 + 1 arcs: {t0.elapsed_time(t1):.2f} ms")
        for h, d in enumerate(degs):
            assert np.array_equal(got[h, :d], np.random.RandomState(seed).permutation(d)), (seed, d)
            assert (got[h, d:] == -1).all()
    # whole trim_index, tensor path vs pandas path (multi-arcs, weights, several hot vertices, untouched ones)
    rng = np.random.default_rng(8)
    src = np.concatenate([rng.integers(0, 40, 600), np.full(3000, 11), np.full(900, 3), np.full(901, 25)])
    perm = rng.permutation(len(src))
    src = src[perm]
    dst = rng.integers(0, 5000, len(src))
    w = rng.uniform(0.1, 2.0, len(src))
    df = pd.DataFrame({"src": src, "dst": dst, "weight": w})
    for seed, cap in ((3, 50), (77, 800)):
        want = n2v.fugue.trim_index(None, df, indexed=True, max_out_deg=cap, random_seed=seed)[0].as_pandas()
        out, none = n2v.fugue.trim_index(None, tuple(torch.as_tensor(x, device="cuda") for x in (src, dst, w)),
                                         indexed=True, max_out_deg=cap, random_seed=seed)
        assert none is None
        assert out[0].cpu().tolist() == want["src"].tolist() and out[1].cpu().tolist() == want["dst"].tolist()
        assert out[2].cpu().numpy().tobytes() == want["weight"].to_numpy().tobytes()
    with pytest.raises(ValueError):
        n2v.fugue.trim_index(None, (torch.as_tensor(src, device="cuda"), torch.as_tensor(dst, device="cuda")),
                             indexed=True, max_out_deg=50, random_seed=2 ** 32)


def test_full_size_properties_config3(n2v):
    """BASELINE configs[2] shape: RMAT scale 20 with 16 injected hotspots of degree 2^18, trimmed to
    max_out_deg = 10000 on the device, 10 walks x 80.  The trimmed graph is directed (hotspots keep
    10k out-arcs but all their in-arcs), so this exercises the general fold at scale."""
    torch = n2v.torch
    from node2vec_b200 import synth
    src, dst = synth.rmat_device(20, 16, seed=42, hotspots=16, hotspot_degree=1 << 18)
    (src, dst), _ = n2v.fugue.trim_index(None, (src, dst), indexed=True, directed=True, max_out_deg=10000,
                                         random_seed=1)
    g = n2v.graph.DeviceGraph.from_arcs(src, dst, None, n_vertices=1 << 20)
    deg = g.degrees()
    assert int(deg.max()) == 10000 and int((deg == 10000).sum()) >= 16
    assert g.flags & 1 and g.flags & 4 and not (g.flags & 2)            # unit weights, simple, NOT symmetric
    start = g.start_vertices()
    walks, alive, stats = g.walk(start, 10, 80, 0.25, 4.0, seed=7, collect_stats=True)
    assert g.ratio is not None and stats["fold_hits"] > 0                # general fold engaged
    assert stats["steps"] == int(alive.sum()) * 80 + 0 * stats["dead"] or stats["dead"] > 0
    keys = torch.sort((src.long() << 32) | dst.long()).values
    sample = walks[alive][:: 64]
    hop = ((sample[:, :-1].long() << 32) | sample[:, 1:].long()).reshape(-1)
    pos = torch.searchsorted(keys, hop).clamp(max=keys.numel() - 1)
    assert bool((keys[pos] == hop).all())                                # every sampled hop is an arc
    assert torch.equal(walks[:, 0], start.repeat_interleave(10))
    w2, a2, _ = g.walk(start[:1000], 10, 80, 0.25, 4.0, seed=7)
    assert torch.equal(w2, walks[:10000]) and torch.equal(a2, alive[:10000])
    rw, ra, rs = _replay(g, 0.25, 4.0, start[:40].cpu().numpy(), 10, 80, 7)
    assert np.array_equal(rw[:, :81], walks[:400].cpu().numpy())


def test_full_size_properties_config4(n2v):
    """BASELINE configs[3] shape: 2.4 M vertices / ~62 M edges (124 M arcs, a 6.5 GB packed graph),
    walk length 80; 2 walks per vertex here (the bench shape is 10) to keep the test short."""
    torch = n2v.torch
    from node2vec_b200 import synth
    src, dst = synth.products_like_device(seed=42)
    V = 2449029
    assert 120_000_000 < src.numel() < 124_000_000
    g = n2v.graph.DeviceGraph.from_arcs(src, dst, None, n_vertices=V)
    assert g.flags & 7 == 7 and g.nbytes() > 6e9
    start = g.start_vertices()
    walks, alive, stats = g.walk(start, 2, 80, 0.5, 2.0, seed=3, collect_stats=True)
    assert bool(alive.all()) and stats["steps"] == walks.shape[0] * 80 and stats["fallbacks"] == 0
    keys = torch.sort((src.long() << 32) | dst.long()).values
    del src, dst
    sample = walks[:: 97]
    hop = ((sample[:, :-1].long() << 32) | sample[:, 1:].long()).reshape(-1)
    pos = torch.searchsorted(keys, hop).clamp(max=keys.numel() - 1)
    assert bool((keys[pos] == hop).all())
    del keys, hop, pos
    w2, _, _ = g.walk(start[50000:50500], 2, 80, 0.5, 2.0, seed=3)
    assert torch.equal(w2, walks[100000:101000])
    rw, ra, rs = _replay(g, 0.5, 2.0, start[:100].cpu().numpy(), 2, 80, 3)
    assert np.array_equal(rw[:, :81], walks[:200].cpu().numpy())
    # SGNS at the config's width (D = 256) on a slice of these walks: trains and stays finite
    from node2vec_b200.sgns import Word2Vec
    m = Word2Vec(size=256, sg=1, negative=5, min_count=1, iter=1, seed=2)
    m.build_vocab(walks[:200000])
    m.train(walks[:200000])
    assert m.train_stats["pairs"] > 200000 * 81 * 4 and bool(torch.isfinite(m.syn0).all())
