"""Pin the C oracle (oracle/csrc/n2v_oracle.c) to the reference through the golden
fixtures, and validate the DESIGN of the device sampler on the CPU: its host replay must
draw from the reference's transition law (chi-square against oracle.ref_walk.transition_law).
"""
import numpy as np
import pytest

from oracle import clib, ref_walk
from tests.helpers import chi_square_ok, graph_flags, load_golden, pack_arcs, unhex


def _mode_of(fx, label):
    return "naive" if label == "naive" else fx["native_sum_mode"]


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert clib.philox((0, 0), (0, 0, 0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert clib.philox((0xFFFFFFFF,) * 2, (0xFFFFFFFF,) * 4) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert clib.philox((0xA4093822, 0x299F31D0), (0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344)) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_mersenne_twister_matches_cpython():
    import random
    for seed in (0, 1, 20, 1000, 2 ** 40 + 5):
        random.seed(seed)
        want = [random.random() for _ in range(700)]
        assert clib.mt_random(seed, 700).tolist() == want


@pytest.mark.parametrize("label", ["native", "naive"])
def test_alias_tables_bit_exact(label):
    fx = load_golden("alias_tables.json")
    mode = _mode_of(fx, label)
    for case in fx["cases"]:
        alias, probs = clib.alias_tables(unhex(case["weights"]), mode)
        assert alias.tolist() == case[label]["alias"]
        assert [p.hex() for p in probs.tolist()] == case[label]["probs"]
    for case in fx["errors"]:
        with pytest.raises(ZeroDivisionError):
            clib.alias_tables(unhex(case["weights"]), mode)


@pytest.mark.parametrize("label", ["native", "naive"])
def test_edge_alias_tables_bit_exact(label):
    fx = load_golden("edge_alias_tables.json")
    mode = _mode_of(fx, label)
    for c in fx["cases"]:
        # a 2-vertex-plus CSR: vertex `cur` = 61 owns the arcs, vertex prev owns prev_out
        prev, cur = c["prev"], 61
        src = [cur] * len(c["ids"]) + [prev] * len(c["prev_out"])
        dst = c["ids"] + c["prev_out"]
        w = unhex(c["weights"]) + [1.0] * len(c["prev_out"])
        row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, w, 62)
        alias, probs = clib.edge_alias_tables(row_ptr, col, ws, prev, cur, c["p"], c["q"], mode)
        assert alias.tolist() == c[label]["alias"]
        assert [x.hex() for x in probs.tolist()] == c[label]["probs"]
    with pytest.raises(ValueError):
        clib.edge_alias_tables(row_ptr, col, ws, prev, cur, 0.0, 1.0)


@pytest.mark.parametrize("label", ["native", "naive"])
def test_whole_walks_match_reference(label):
    fx = load_golden("walks.json")
    mode = _mode_of(fx, label)
    for rec in fx["walks"]:
        row_ptr, col, w, _ = clib.csr_from_arcs(rec["src"], rec["dst"], unhex(rec["weight"]))
        starts = np.flatnonzero(np.diff(row_ptr) > 0)
        if rec["walk_seed"] is not None:
            starts = np.array([v for v in starts if v in set(rec["walk_seed"])])
        P = {"num_walks": 10, "walk_length": 20, "return_param": 1.0, "inout_param": 1.0, **rec["params"]}
        walks, alive = clib.reference_walk(row_ptr, col, w, starts, P["num_walks"], P["walk_length"],
                                           P["return_param"], P["inout_param"], mode, rec["random_seed"])
        assert walks[alive].tolist() == rec[label], rec["graph"]


def test_c_port_equals_python_port_on_random_graph():
    rng = np.random.default_rng(3)
    n, m = 300, 3000
    src = rng.integers(0, n, m)
    dst = rng.integers(0, n, m)
    w = rng.uniform(0.1, 2.0, m)
    row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, w, n)
    starts = np.flatnonzero(np.diff(row_ptr) > 0)[:40]
    got, alive = clib.reference_walk(row_ptr, col, ws, starts, 2, 6, 0.5, 2.0, "naive", 77)
    want = ref_walk.random_walk(src.tolist(), dst.tolist(), w.tolist(),
                                {"num_walks": 2, "walk_length": 6, "return_param": 0.5, "inout_param": 2.0},
                                starts.tolist(), 77)
    assert got[alive].tolist() == want


# ---- the device sampler's design, checked on the CPU --------------------------------------
def _replay_setup(src, dst, w, n):
    row_ptr, col, ws, _ = clib.csr_from_arcs(src, dst, w, n)
    alias, probs, bad = clib.alias_tables_csr(row_ptr, ws)
    assert bad == 0
    thr, adst, aalias = pack_arcs(row_ptr, col, alias, probs)
    flags = graph_flags(row_ptr, col, ws)
    return row_ptr, col, ws, thr, adst, aalias, flags


def _second_order_counts(walks, alive, pos):
    """(prev, cur) -> {next: count} at one step index of every surviving walk."""
    out = {}
    for row in walks[alive]:
        key = (int(row[pos - 1]), int(row[pos]))
        out.setdefault(key, {}).setdefault(int(row[pos + 1]), 0)
        out[key][int(row[pos + 1])] += 1
    return out


@pytest.mark.parametrize("p,q,weighted,sym,use_ratio", [
    (1.0, 1.0, True, False, False), (1.0, 0.5, False, True, False), (0.25, 4.0, False, True, False),
    (4.0, 0.25, True, True, False), (0.25, 4.0, True, False, False), (0.5, 2.0, False, False, False),
    (0.25, 4.0, True, False, True), (0.2, 1.0, True, True, True), (0.5, 2.0, False, False, True),
    (4.0, 2.0, False, True, False), (1.0, 3.0, False, True, False), (0.1, 1.5, False, True, False),
])
def test_replay_draws_from_reference_law(p, q, weighted, sym, use_ratio):
    rng = np.random.default_rng(11)
    n = 12
    pairs = {(int(a), int(b)) for a, b in rng.integers(0, n, (60, 2)) if a != b}
    if sym:
        pairs |= {(b, a) for a, b in pairs}
    pairs = sorted(pairs)
    wmap = {}
    for a, b in pairs:
        key = (min(a, b), max(a, b)) if sym else (a, b)
        wmap.setdefault(key, float(rng.uniform(0.2, 2.0)) if weighted else 1.0)
    src = [a for a, _ in pairs]
    dst = [b for _, b in pairs]
    w = [wmap[(min(a, b), max(a, b)) if sym else (a, b)] for a, b in pairs]
    row_ptr, col, ws, thr, adst, aalias, flags = _replay_setup(src, dst, w, n)
    consts = clib.walk_consts(p, q, flags)
    plain = not weighted and sym                   # unit-weight symmetric simple graph
    assert consts.fold_mode == (3 if (plain and q > 1.0) else 1 if (plain and p < min(1.0, q)) else 0)
    starts = np.flatnonzero(np.diff(row_ptr) > 0).astype(np.int32)
    ratio = alias_idx = None
    if use_ratio:
        alias_idx, _, _ = clib.alias_tables_csr(row_ptr, ws)
        ratio = clib.return_ratios(row_ptr, col, ws)
        c2 = clib.walk_consts(p, q, flags, True)
        assert c2.fold_mode == (2 if (p < min(1.0, q) and consts.fold_mode == 0) else consts.fold_mode)
    walks, alive, stats = clib.replay_walk(row_ptr[:-1], np.diff(row_ptr), thr, adst, aalias, col, ws, flags,
                                           p, q, starts, 6000, 3, seed=1234, alias_idx=alias_idx, ratio=ratio)
    adj = ref_walk.build_adjacency(src, dst, w)
    # first step: unbiased law
    for v in starts[:4]:
        law = ref_walk.transition_law(adj, None, int(v), p, q)
        rows = walks[alive & (walks[:, 0] == v)]
        ids = sorted(law)
        counts = [(rows[:, 1] == x).sum() for x in ids]
        ok, pval = chi_square_ok(counts, [law[x] for x in ids])
        assert ok, (v, pval)
    # second and third steps: biased law, the most visited (prev, cur) pairs
    for pos in (1, 2):
        table = _second_order_counts(walks, alive, pos)
        top = sorted(table, key=lambda k: -sum(table[k].values()))[:12]
        for (t, v) in top:
            law = ref_walk.transition_law(adj, t, v, p, q)
            ids = sorted(law)
            assert set(table[(t, v)]) <= set(ids)
            counts = [table[(t, v)].get(x, 0) for x in ids]
            ok, pval = chi_square_ok(counts, [law[x] for x in ids])
            assert ok, (t, v, pval, stats)
    assert stats["steps"] == alive.sum() * 3 + 0 * stats["dead"] or stats["dead"] > 0


def test_replay_fallback_and_sinks():
    # hub with extreme q: almost every proposal is rejected -> the exact scan must kick in
    n = 40
    src, dst = [], []
    for i in range(1, n):
        src += [0, i]
        dst += [i, 0]
    src += [1, 2]; dst += [2, 1]
    src += [5]; dst += [n]          # n is a sink
    row_ptr, col, ws, thr, adst, aalias, flags = _replay_setup(src, dst, [1.0] * len(src), n + 1)
    starts = np.array([3, 5], dtype=np.int32)
    walks, alive, stats = clib.replay_walk(row_ptr[:-1], np.diff(row_ptr), thr, adst, aalias, col, ws, flags,
                                           1e6, 1e6, starts, 4000, 2, seed=9)
    assert stats["fallbacks"] > 0 and stats["dead"] > 0
    assert (~alive).sum() == stats["dead"]
    assert (walks[~alive][:, 2:] == -1).all()
    adj = ref_walk.build_adjacency(src, dst, [1.0] * len(src))
    rows = walks[alive & (walks[:, 0] == 3) & (walks[:, 1] == 0)]
    assert stats["fallbacks"] >= len(rows) * 0.9
    law = ref_walk.transition_law(adj, 3, 0, 1e6, 1e6)
    ids = sorted(law)
    ok, pval = chi_square_ok([(rows[:, 2] == x).sum() for x in ids], [law[x] for x in ids])
    assert ok, pval
