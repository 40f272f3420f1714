"""Pin the CPU oracle (oracle/ref_walk.py, oracle/ref_indexer.py) to the reference.

Fixtures under tests/golden/ were produced by running the unmodified reference
(tests/golden/make_golden.py).  Everything here is bit-exact: floats are compared
through their hex strings.  Also restates the known-answer vectors of the
reference's own tests/test_randomwalk.py.
"""
import os
import random

import numpy as np
import pytest

from oracle import ref_indexer, ref_walk
from tests.helpers import load_golden, unhex

MODES = {"naive": "naive"}


def _mode_of(fixture, label):
    return "naive" if label == "naive" else fixture["native_sum_mode"]


@pytest.mark.parametrize("label", ["native", "naive"])
def test_alias_tables_bit_exact(label):
    fx = load_golden("alias_tables.json")
    mode = _mode_of(fx, label)
    for case in fx["cases"]:
        alias, probs = ref_walk.alias_tables(unhex(case["weights"]), mode)
        assert alias == case[label]["alias"]
        assert [p.hex() for p in probs] == case[label]["probs"]


def test_alias_tables_sum_modes_really_differ():
    fx = load_golden("alias_tables.json")
    assert fx["native_sum_mode"] == "neumaier"
    assert any(c["native"]["probs"] != c["naive"]["probs"] for c in fx["cases"])


def test_alias_tables_errors():
    fx = load_golden("alias_tables.json")
    for case in fx["errors"]:
        assert case["raises"] == "ZeroDivisionError"
        with pytest.raises(ZeroDivisionError):
            ref_walk.alias_tables(unhex(case["weights"]))


@pytest.mark.parametrize("label", ["native", "naive"])
def test_edge_alias_tables_bit_exact(label):
    fx = load_golden("edge_alias_tables.json")
    mode = _mode_of(fx, label)
    for c in fx["cases"]:
        alias, probs = ref_walk.edge_alias_tables(
            c["prev"], set(c["prev_out"]), (c["ids"], unhex(c["weights"])), c["p"], c["q"], mode
        )
        assert alias == c[label]["alias"]
        assert [p.hex() for p in probs] == c[label]["probs"]


def test_samplers_and_append():
    fx = load_golden("samplers.json")
    for c in fx["draws"]:
        probs = unhex(c["probs"])
        r1, r2 = float.fromhex(c["r1"]), float.fromhex(c["r2"])
        assert ref_walk.draw_two_uniform(c["alias"], probs, r1, r2) == c["two"]
        assert ref_walk.draw_one_uniform(c["alias"], probs, r1) == c["one"]
    for c in fx["appends"]:
        r2 = None if c["r2"] is None else float.fromhex(c["r2"])
        out = ref_walk.extend_path(c["path"], c["ids"], c["alias"], unhex(c["probs"]),
                                   float.fromhex(c["r1"]), r2)
        assert out == c["out"]
    for seed, (a, b) in fx["mt_seeds"].items():
        random.seed(int(seed))
        assert random.random().hex() == a and random.random().hex() == b


@pytest.mark.parametrize("label", ["native", "naive"])
def test_whole_walks_match_reference(label):
    fx = load_golden("walks.json")
    mode = _mode_of(fx, label)
    for rec in fx["walks"]:
        walks = ref_walk.random_walk(
            rec["src"], rec["dst"], unhex(rec["weight"]),
            {**{"num_walks": 10, "walk_length": 20, "return_param": 1.0, "inout_param": 1.0}, **rec["params"]},
            rec["walk_seed"], rec["random_seed"], mode,
        )
        assert walks == rec[label], rec["graph"]
        L = rec["params"].get("walk_length", 20)
        assert all(len(w) == L + 1 for w in walks)


def test_sink_drops_walkers():
    fx = load_golden("walks.json")
    rec = [r for r in fx["walks"] if r["graph"] == "sink_multi"][0]
    n_start = len(set(rec["src"])) * rec["params"]["num_walks"]
    assert 0 < len(rec["naive"]) < n_start          # some walkers died at the sink
    adj_srcs = set(rec["src"])
    for w in rec["naive"]:
        assert all(v in adj_srcs for v in w[:-1])    # only the last vertex may be a sink


# ---- the reference's own known-answer tests (tests/test_randomwalk.py) ----------------
@pytest.mark.parametrize("weights,alias,probs", [
    ([0.5, 0.8, 1.0], [2, 0, 1], [0.6521739, 1.0, 0.9565217]),      # :135
    ([0.5, 0.2], [0, 0], [1.0, 0.5714285714285715]),                # :136
    ([0.2], [0], [1.0]),                                            # :137
    ([1.0], [0], [1.0]),                                            # :138
])
def test_reference_kat_alias(weights, alias, probs):
    for mode in ref_walk.SUM_MODES:
        a, p = ref_walk.alias_tables(weights, mode)
        assert a == alias
        np.testing.assert_almost_equal(p, probs, decimal=7)


@pytest.mark.parametrize("prev,shared,nbrs,p,q,alias,probs", [
    (0, {2}, ([0, 2], [0.5, 0.2]), 1.0, 1.0, [0, 0], [1.0, 0.5714285714285715]),   # :158
    (1, set(), ([1], [0.2]), 0.8, 1.5, [0], [1.0]),                               # :159
    (3, set(), ([1, 3], [0.5, 1.0]), 2.0, 4.0, [1, 0], [0.4, 1.0]),               # :160
])
def test_reference_kat_edge_alias(prev, shared, nbrs, p, q, alias, probs):
    a, pr = ref_walk.edge_alias_tables(prev, shared, nbrs, p, q)
    assert a == alias
    np.testing.assert_almost_equal(pr, probs, decimal=7)
    with pytest.raises(ValueError):                                                # :184-189
        ref_walk.edge_alias_tables(prev, shared, nbrs, 0)
    with pytest.raises(ValueError):
        ref_walk.edge_alias_tables(prev, shared, nbrs, 1.0, 0)
    with pytest.raises(ValueError):
        ref_walk.edge_alias_tables(prev, shared, (nbrs[0], nbrs[1][:-1]))


def test_reference_kat_single_rows():
    """tests/test_randomwalk.py:268-306 -- three seeded rows at p=q=1."""
    fx = load_golden("walks.json")["single_rows"]
    adj = {1: ([0, 2, 4], [0.5, 0.9, 1.0]), 2: ([0, 3], [1.2, 0.9]), 0: ([2], [1.0])}
    rows = [
        ({"src": 0, "dst": 1, "path": [3, 0, 1]}, adj, 1000),
        ({"src": 0, "dst": 2, "path": [2, 0, 2]}, {1: adj[1], 2: adj[2]}, 10),
        ({"src": -1, "dst": 2, "path": [-1, 2]}, {2: adj[2]}, 20),
    ]
    expect = [(1, 4, [3, 0, 1, 4]), (2, 3, [2, 0, 2, 3]), (2, 3, [2, 3])]
    for (row, a, seed), want, ref in zip(rows, expect, fx):
        random.seed(seed)
        got = ref_walk.step_row(row, a, 1.0, 1.0, random.random(), random.random())
        assert (got["src"], got["dst"], got["path"]) == want
        assert (ref["src"], ref["dst"], ref["path"]) == want


def test_start_rows():
    """tests/test_randomwalk.py:245-264."""
    rows = ref_walk.start_rows([3, 2], 3)
    assert [(r["src"], r["dst"], r["path"]) for r in rows] == [
        (-1, 3, [-1, 3]), (-2, 3, [-2, 3]), (-3, 3, [-3, 3]),
        (-1, 2, [-1, 2]), (-2, 2, [-2, 2]), (-3, 2, [-3, 2])]


# ---- indexer and trimming ----------------------------------------------------------
def test_indexer_bit_exact():
    fx = load_golden("indexer.json")
    for c in fx["cases"]:
        s, d, w, names, ids = ref_indexer.index_graph(c["src"], c["dst"], c["weight"], c["directed"])
        assert s.tolist() == c["edge_src"]
        assert d.tolist() == c["edge_dst"]
        assert [x.hex() for x in w.tolist()] == c["edge_weight"]
        assert ids.tolist() == c["vertex_id"]
        assert names == c["vertex_name"]
        assert c["edge_columns"] == ["src", "dst", "weight"]
        assert c["name_id_columns"] == ["vertex_id", "vertex_name"]
        assert c["edge_weight_dtype"] == "float64"


def test_indexer_sparse_ids_example():
    """src=[a1..a4], dst=[a2,b1,b2,a1] -> a1:0 a2:1 a3:2 a4:3 b1:5 b2:6."""
    s, d, w, names, ids = ref_indexer.index_graph(
        ["a1", "a2", "a3", "a4"], ["a2", "b1", "b2", "a1"], None, False)
    assert dict(zip(names, ids.tolist())) == {"a1": 0, "a2": 1, "a3": 2, "a4": 3, "b1": 5, "b2": 6}
    assert len(s) == 8 and w.dtype == np.float64


def test_trim_hotspot_matches_reference():
    fx = load_golden("trim.json")
    for c in fx["cases"]:
        if c["random_seed"] is None and c["max_out_degree"] > 0:
            continue
        df = ref_indexer.trim_hotspot(fx["src"], fx["dst"], unhex(fx["weight"]),
                                      c["max_out_degree"], c["random_seed"])
        assert df["src"].tolist() == c["src"]
        assert df["dst"].tolist() == c["dst"]
        assert [x.hex() for x in df["weight"].tolist()] == c["weight"]


# ---- the reference's own entry points, run verbatim on the Fugue stand-in ------------
@pytest.mark.parametrize("label", ["native", "naive"])
def test_verbatim_fugue_random_walk(label):
    """fugue_verbatim.json holds node2vec.fugue.random_walk outputs (the real DAG code: joins,
    per-step re-seeding, sink drops, duplicated walk_seed ids); the oracle must reproduce them."""
    fx = load_golden("fugue_verbatim.json")
    mode = _mode_of(fx, label)
    assert len(fx["walks"]) >= 5
    for rec in fx["walks"]:
        params = dict(rec["params"])
        for k, v in {"num_walks": 10, "walk_length": 20, "return_param": 1.0, "inout_param": 1.0}.items():
            params.setdefault(k, v)                              # constants.py:15-20
        assert params == rec["params_after"]
        got = ref_walk.random_walk(rec["src"], rec["dst"], unhex(rec["weight"]), params, rec["walk_seed"],
                                   rec["random_seed"], sum_mode=mode, rng=random.Random())
        assert got == rec[label]["walk"], rec["graph"]
        assert [w[0] for w in got] == rec[label]["src"]


def test_verbatim_fugue_duplicate_seed_multiplicity():
    fx = load_golden("fugue_verbatim.json")
    rec = next(r for r in fx["walks"] if r["walk_seed"] and len(set(r["walk_seed"])) < len(r["walk_seed"]))
    heads = [w[0] for w in rec["naive"]["walk"]]
    nw = rec["params"]["num_walks"]
    assert heads == [v for v in sorted(set(rec["walk_seed"])) for _ in range(rec["walk_seed"].count(v) * nw)]


def test_verbatim_fugue_trim_index():
    """trim_index = trim per src partition FIRST (on the raw names, fugue.py:57-67), THEN index
    (fugue.py:69-77)."""
    fx = load_golden("fugue_verbatim.json")
    assert len(fx["trim_index"]) >= 5
    for c in fx["trim_index"]:
        kw, inp = c["kwargs"], c["input"]
        has_w = "weight" in inp
        wt = inp["weight"] if has_w else [1.0] * len(inp["src"])
        df = ref_indexer.trim_hotspot(inp["src"], inp["dst"], wt, kw.get("max_out_deg", 0), kw.get("random_seed"))
        s, d, w = df["src"].tolist(), df["dst"].tolist(), df["weight"].to_numpy(dtype=np.float64)
        if kw["indexed"]:
            assert c["name_id"] is None
        else:
            s, d, w, names, ids = ref_indexer.index_graph(s, d, w if has_w else None, kw.get("directed", True))
            assert c["name_id"] == {"vertex_id": ids.tolist(), "vertex_name": names}
            s, d = s.tolist(), d.tolist()
        assert s == c["src"] and d == c["dst"], kw
        assert [float(x).hex() for x in np.asarray(w).tolist()] == c["weight"]


def test_numpy_legacy_permutation_restatement():
    """oracle/ref_indexer.py::numpy_legacy_permutation (the algorithm K5 runs on the device) against
    numpy itself and against pandas.DataFrame.sample, the call the reference makes."""
    import pandas as pd
    for seed in (0, 5, 2 ** 32 - 1):
        for n in (1, 2, 9, 624, 625, 3000):
            assert ref_indexer.numpy_legacy_permutation(n, seed) == np.random.RandomState(seed).permutation(n).tolist()
    df = pd.DataFrame({"x": range(700)})
    assert df.sample(n=30, random_state=4)["x"].tolist() == ref_indexer.numpy_legacy_permutation(700, 4)[:30]
    with pytest.raises(ValueError):
        ref_indexer.numpy_legacy_permutation(5, 2 ** 32)


@pytest.mark.skipif(not os.path.isdir(os.environ.get("N2V_REFERENCE", "/root/reference")),
                    reason="the reference checkout only exists in the build container")
def test_fixtures_regenerate_from_the_unmodified_reference(tmp_path):
    """Re-run tests/golden/make_golden.py against the live reference checkout (a fresh interpreter:
    the script stubs pyspark / shims DataFrame.append) and require byte-identical fixtures: the
    committed goldens are what the unmodified reference computes, today."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, "tests", "golden", "make_golden.py"), str(tmp_path)],
                   check=True, capture_output=True, timeout=600, env={**os.environ, "PYTHONDONTWRITEBYTECODE": "1"})
    names = sorted(f for f in os.listdir(tmp_path) if f.endswith(".json"))
    assert names == sorted(f for f in os.listdir(os.path.join(root, "tests", "golden")) if f.endswith(".json"))
    for name in names:
        with open(os.path.join(tmp_path, name), "rb") as a, open(os.path.join(root, "tests", "golden", name), "rb") as b:
            assert a.read() == b.read(), name
