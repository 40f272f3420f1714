#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py            # rewrites tests/golden/*.json

Nothing from the reference is copied: its functions are imported from where they lie
and fed seeded inputs; inputs and outputs are written as JSON (floats as C99 hex
strings so they round-trip bit-for-bit).

Interpreter note: CPython >= 3.12 evaluates ``sum(list_of_floats)`` with Neumaier
compensation, older interpreters (the reference's supported 3.6/3.7) add left to
right.  Each alias fixture is therefore produced twice: ``native`` (reference as it
runs under this interpreter) and ``naive`` (module global ``sum`` shadowed by a
left-to-right loop, i.e. the reference as it runs under its own supported
interpreters).  The reference source is not touched in either case.
"""
import json
import os
import random
import sys
import types

import numpy as np
import pandas as pd

REF = os.environ.get("N2V_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

import node2vec.randomwalk as ref_rw  # noqa: E402
from node2vec.constants import NODE2VEC_PARAMS  # noqa: E402


def hx(xs):
    return [float(x).hex() for x in xs]


def naive_sum(xs):
    total = 0
    for x in xs:
        total = total + x
    return total


class sum_mode:
    """Shadow ``sum`` inside the reference module's namespace (not its source)."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        if self.mode == "naive":
            ref_rw.sum = naive_sum

    def __exit__(self, *a):
        if hasattr(ref_rw, "sum"):
            del ref_rw.sum


def native_mode_name():
    return "neumaier" if sys.version_info >= (3, 12) else "naive"


# ----------------------------------------------------------------------------------
def weight_vectors(rng):
    vecs = [
        [0.5, 0.8, 1.0], [0.5, 0.2], [0.2], [1.0], [1.0, 2.0, 3.0, 4.0],
        [1.0] * 7, [3.0] * 64, [0.1] * 10, [0.1, 0.2, 0.3, 0.4, 0.5, 0.6],
        [1e-300, 1.0], [1e300, 1.0, 1.0], [0.0, 1.0], [0.0, 0.0, 2.5],
        [1.0, 1.0, 1.0, 1e-9], [2.0, 2.0, 1.0, 1.0, 1.0, 1.0],
    ]
    for _ in range(60):
        n = rng.randint(1, 48)
        vecs.append([rng.uniform(0.1, 2.0) for _ in range(n)])
    for _ in range(10):  # ties and exact-1.0 probs
        n = rng.randint(2, 30)
        vecs.append([float(rng.choice([1, 1, 2, 3])) for _ in range(n)])
    for n in (257, 1000, 2053):  # hubs
        vecs.append([rng.paretovariate(1.5) for _ in range(n)])
    return vecs


def gen_alias(rng):
    cases = []
    for w in weight_vectors(rng):
        case = {"weights": hx(w)}
        for label in ("native", "naive"):
            with sum_mode(label):
                alias, probs = ref_rw.generate_alias_tables(list(w))
            case[label] = {"alias": alias, "probs": hx(probs)}
        cases.append(case)
    errors = []
    for w in ([], [0.0, 0.0]):
        try:
            ref_rw.generate_alias_tables(list(w))
            errors.append({"weights": hx(w), "raises": None})
        except Exception as e:  # noqa: BLE001
            errors.append({"weights": hx(w), "raises": type(e).__name__})
    return {"native_sum_mode": native_mode_name(), "cases": cases, "errors": errors}


def gen_edge_alias(rng):
    cases = []
    for _ in range(80):
        n = rng.randint(1, 24)
        ids = sorted(rng.sample(range(60), n))
        wts = [rng.uniform(0.1, 2.0) for _ in range(n)] if rng.random() < 0.7 else [1.0] * n
        prev = rng.choice(ids) if rng.random() < 0.7 else rng.randrange(60)
        prev_out = set(rng.sample(range(60), rng.randint(0, 20)))
        p = rng.choice([0.25, 0.5, 1.0, 2.0, 4.0])
        q = rng.choice([0.25, 0.5, 1.0, 2.0, 4.0])
        case = {"prev": prev, "prev_out": sorted(prev_out), "ids": ids, "weights": hx(wts),
                "p": p, "q": q}
        for label in ("native", "naive"):
            with sum_mode(label):
                alias, probs = ref_rw.generate_edge_alias_tables(prev, prev_out, (ids, wts), p, q)
            case[label] = {"alias": alias, "probs": hx(probs)}
        cases.append(case)
    return {"native_sum_mode": native_mode_name(), "cases": cases}


def gen_samplers(rng):
    cases = []
    for _ in range(200):
        n = rng.randint(1, 12)
        alias, probs = ref_rw.generate_alias_tables([rng.uniform(0.1, 2.0) for _ in range(n)])
        ap = ref_rw.AliasProb((alias, probs))
        r1, r2 = rng.random(), rng.random()
        cases.append({"alias": alias, "probs": hx(probs), "r1": r1.hex(), "r2": r2.hex(),
                      "two": ap.sampling_from_alias(r1, r2),
                      "one": ap.sampling_from_alias_wiki(r1)})
    paths = []
    for _ in range(60):
        n = rng.randint(1, 8)
        ids = sorted(rng.sample(range(40), n))
        alias, probs = ref_rw.generate_alias_tables([rng.uniform(0.1, 2.0) for _ in range(n)])
        if rng.random() < 0.4:
            path = [-rng.randint(1, 5), rng.randrange(40)]
        else:
            path = [rng.randrange(40) for _ in range(rng.randint(2, 6))]
        r1 = rng.random()
        r2 = rng.random() if rng.random() < 0.7 else None
        out = ref_rw.RandomPath(list(path)).append(ids, ref_rw.AliasProb((alias, probs)), r1, r2).path
        paths.append({"path": path, "ids": ids, "alias": alias, "probs": hx(probs),
                      "r1": r1.hex(), "r2": None if r2 is None else r2.hex(), "out": out})
    seeds = {}
    for s in (20, 10, 1000):
        random.seed(s)
        seeds[str(s)] = [random.random().hex(), random.random().hex()]
    return {"draws": cases, "appends": paths, "mt_seeds": seeds}


# ----------------------------------------------------------------------------------
def reference_random_walk(src, dst, wt, params, walk_seed=None, random_seed=None):
    """fugue.random_walk (fugue.py:119-155) re-driven without Fugue: the reference's own
    transformer functions, chained with dict joins in place of the two DataFrame joins.
    Row order = start vertex ascending x walk number, preserved through the joins (as
    pandas left/inner merges do on one partition)."""
    for k, v in NODE2VEC_PARAMS.items():
        params.setdefault(k, v)
    df = pd.DataFrame({"src": src, "dst": dst, "weight": wt})
    adj = {}
    for s, part in df.groupby("src", sort=True):
        part = part.sort_values("dst", kind="stable").reset_index(drop=True)
        row = next(iter(ref_rw.get_vertex_neighbors(part)))
        adj[int(row["id"])] = row["neighbors"]
    starts = sorted(adj)
    if walk_seed is not None:
        starts = [v for v in starts if v in set(walk_seed)]
    rows = [dict(r) for r in ref_rw.initiate_random_walk([{"id": v} for v in starts], params["num_walks"])]
    for _ in range(params["walk_length"]):
        joined = []
        for r in rows:
            if r["dst"] not in adj:      # inner join with df_dst drops walkers at sinks
                continue
            joined.append({"src": r["src"], "path": r["path"],
                           "src_neighbors": adj.get(r["src"]),   # left join: None if absent
                           "dst_neighbors": adj[r["dst"]]})
        rows = [dict(r) for r in ref_rw.next_step_random_walk(
            joined, params["return_param"], params["inout_param"], random_seed)]
    return [r["walk"] for r in ref_rw.to_path(rows)]


def small_graphs(rng):
    g = {}
    # the reference's own test graph (tests/test_fugue.py:65-66)
    g["ref_test_graph"] = ([0, 0, 3, 2, 4, 4], [2, 4, 4, 0, 0, 3], [0.41, 0.85, 0.36, 0.68, 0.1, 0.37])
    # with a sink (vertex 9 has no out-arcs) and a multi-arc
    g["sink_multi"] = ([0, 0, 1, 1, 2, 2, 2, 3, 3, 3], [1, 9, 0, 2, 0, 3, 3, 2, 9, 1],
                       [1.0, 0.5, 1.0, 2.0, 0.3, 1.0, 0.25, 1.5, 0.1, 0.7])
    # weighted undirected ER
    n, m = 40, 120
    pairs = set()
    while len(pairs) < m:
        a, b = rng.randrange(n), rng.randrange(n)
        if a != b:
            pairs.add((min(a, b), max(a, b)))
    s, d, w = [], [], []
    for a, b in sorted(pairs):
        ww = rng.uniform(0.1, 2.0)
        s += [a, b]; d += [b, a]; w += [ww, ww]
    g["er40_weighted"] = (s, d, w)
    # unweighted directed with sparse ids
    s = [rng.randrange(0, 300, 7) for _ in range(150)]
    d = [rng.randrange(0, 300, 7) for _ in range(150)]
    g["sparse_directed"] = (s, d, [1.0] * 150)
    return g


def gen_walks(rng):
    out = []
    graphs = small_graphs(rng)
    settings = [
        ("ref_test_graph", {"num_walks": 2, "walk_length": 3, "return_param": 0.5}, None, 7),
        ("ref_test_graph", {"num_walks": 3, "walk_length": 6, "return_param": 0.25, "inout_param": 4.0}, [0, 4], 11),
        ("sink_multi", {"num_walks": 3, "walk_length": 5, "return_param": 2.0, "inout_param": 0.5}, None, 3),
        ("er40_weighted", {"num_walks": 2, "walk_length": 8, "return_param": 1.0, "inout_param": 0.5}, None, 42),
        ("er40_weighted", {"num_walks": 1, "walk_length": 5, "return_param": 4.0, "inout_param": 0.25}, None, 5),
        ("sparse_directed", {"num_walks": 2, "walk_length": 6}, None, 99),
    ]
    for name, params, walk_seed, seed in settings:
        s, d, w = graphs[name]
        for label in ("native", "naive"):
            with sum_mode(label):
                walks = reference_random_walk(list(s), list(d), list(w), dict(params), walk_seed, seed)
            if label == "native":
                rec = {"graph": name, "src": s, "dst": d, "weight": hx(w), "params": params,
                       "walk_seed": walk_seed, "random_seed": seed}
            rec[label] = walks
        out.append(rec)
    # the three seeded single rows of tests/test_randomwalk.py:268-306, re-run here
    rows = [
        {"src": 0, "path": [3, 0, 1], "dst_neighbors": ref_rw.Neighbors(([0, 2, 4], [0.5, 0.9, 1.0])).serialize(),
         "src_neighbors": ref_rw.Neighbors(([2], [1.0])).serialize()},
        {"src": 0, "path": [2, 0, 2], "dst_neighbors": ref_rw.Neighbors(([0, 3], [1.2, 0.9])).serialize(),
         "src_neighbors": None},
        {"src": -1, "path": [-1, 2], "dst_neighbors": ref_rw.Neighbors(([0, 3], [1.2, 0.9])).serialize(),
         "src_neighbors": None},
    ]
    it = iter(ref_rw.next_step_random_walk(rows, 1.0, 1.0, 1000))
    single = [next(it)]
    random.seed(10)
    single.append(next(it))
    random.seed(20)
    single.append(next(it))
    return {"native_sum_mode": native_mode_name(), "walks": out, "single_rows": single}


# ----------------------------------------------------------------------------------
def import_reference_indexer():
    """indexer.py imports pyspark at module top (:4-6) and uses DataFrame.append (:28,48),
    removed in pandas 2.  Stub the former, shim the latter; the function body runs as is."""
    for name in ("pyspark", "pyspark.sql", "pyspark.sql.functions"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pyspark.sql"].DataFrame = object
    sys.modules["pyspark.sql"].Row = object
    sys.modules["pyspark.sql"].functions = sys.modules["pyspark.sql.functions"]
    if not hasattr(pd.DataFrame, "append"):
        def _append(self, other, ignore_index=False):
            return pd.concat([self, other], ignore_index=ignore_index)
        pd.DataFrame.append = _append
    import node2vec.indexer as ref_ix
    return ref_ix


def gen_indexer(rng):
    ref_ix = import_reference_indexer()
    inputs = [
        {"src": ["a1", "a2", "a3", "a4"], "dst": ["a2", "b1", "b2", "a1"], "weight": None},
        {"src": ["a1", "a1", "a1", "a2", "b2"], "dst": ["a2", "b1", "b2", "b1", "a2"], "weight": None},
        {"src": ["x", "y", "x", "z", "y"], "dst": ["y", "x", "y", "x", "z"], "weight": [1.0, 1.0, 2.0, 0.5, 0.5]},
        {"src": [10, 20, 30, 10, 50], "dst": [20, 10, 10, 30, 50], "weight": [1, 1, 2, 2, 3]},
    ]
    names = [f"v{i}" for i in range(30)]
    s = [rng.choice(names) for _ in range(80)]
    d = [rng.choice(names) for _ in range(80)]
    inputs.append({"src": s, "dst": d, "weight": [rng.choice([0.5, 1.0, 1.5]) for _ in range(80)]})
    inputs.append({"src": s, "dst": d, "weight": None})
    cases = []
    for inp in inputs:
        for directed in (True, False):
            cols = {"src": list(inp["src"]), "dst": list(inp["dst"])}
            if inp["weight"] is not None:
                cols["weight"] = list(inp["weight"])
            df_edge, name_id = ref_ix.index_graph_pandas(pd.DataFrame(cols), directed)
            cases.append({
                "src": inp["src"], "dst": inp["dst"], "weight": inp["weight"], "directed": directed,
                "edge_columns": list(df_edge.columns), "name_id_columns": list(name_id.columns),
                "edge_src": [int(x) for x in df_edge["src"]], "edge_dst": [int(x) for x in df_edge["dst"]],
                "edge_weight": hx(df_edge["weight"]), "edge_weight_dtype": str(df_edge["weight"].dtype),
                "vertex_id": [int(x) for x in name_id["vertex_id"]],
                "vertex_name": [x if isinstance(x, str) else int(x) for x in name_id["vertex_name"]],
            })
    return {"cases": cases}


def gen_trim(rng):
    cases = []
    src = [rng.randrange(6) for _ in range(60)]
    dst = [rng.randrange(50) for _ in range(60)]
    wt = [rng.uniform(0.1, 2.0) for _ in range(60)]
    df = pd.DataFrame({"src": src, "dst": dst, "weight": wt})
    for max_deg, seed in ((0, None), (3, 20), (8, 5), (100, 1)):
        rows = []
        for _, part in df.groupby("src", sort=True):
            rows += list(ref_rw.trim_hotspot_vertices(part.reset_index(drop=True), max_deg, seed))
        cases.append({"max_out_degree": max_deg, "random_seed": seed,
                      "src": [int(r["src"]) for r in rows], "dst": [int(r["dst"]) for r in rows],
                      "weight": hx([r["weight"] for r in rows])})
    return {"src": src, "dst": dst, "weight": hx(wt), "cases": cases}


def gen_fugue_verbatim(rng):
    """The reference's OWN entry points, node2vec.fugue.trim_index / random_walk, run verbatim on
    the pandas stand-in for Fugue under tests/golden/fugue_shim (Fugue is not installable here)."""
    sys.path.insert(0, os.path.join(HERE, "fugue_shim"))
    import_reference_indexer()                      # pyspark stub + DataFrame.append shim
    from fugue import ArrayDataFrame, NativeExecutionEngine, PandasDataFrame
    import node2vec.fugue as ref_fugue
    graphs = small_graphs(random.Random(7))
    walks = []
    settings = [
        ("ref_test_graph", {"num_walks": 2, "walk_length": 3, "return_param": 0.5}, None, 7),
        ("ref_test_graph", {"num_walks": 2, "walk_length": 4, "return_param": 0.5}, [0, 0, 3, 2, 4, 4], 7),  # tests/test_fugue.py:73-75: duplicated seeds
        ("sink_multi", {"num_walks": 4, "walk_length": 6, "return_param": 0.5, "inout_param": 2.0}, None, 13),
        ("er40_weighted", {"num_walks": 2, "walk_length": 10, "return_param": 0.25, "inout_param": 4.0}, list(range(0, 40, 3)), 21),
        ("sparse_directed", {"num_walks": 3, "walk_length": 5, "return_param": 2.0, "inout_param": 0.5}, None, 1),
    ]
    for name, params, seeds, seed in settings:
        s, d, w = graphs[name]
        rec = {"graph": name, "src": s, "dst": d, "weight": hx(w), "params": dict(params), "walk_seed": seeds,
               "random_seed": seed}
        for label in ("native", "naive"):
            with sum_mode(label):
                df = PandasDataFrame(pd.DataFrame({"src": s, "dst": d, "weight": w}))
                sd = None if seeds is None else PandasDataFrame(pd.DataFrame({"id": seeds}))
                p = dict(params)
                res = ref_fugue.random_walk(NativeExecutionEngine(), df, p, sd, random_seed=seed).as_pandas()
            rec[label] = {"src": [int(x) for x in res["src"]], "walk": [list(map(int, x)) for x in res["walk"]]}
            rec["params_after"] = p                 # defaults merged into the caller's dict in place
        walks.append(rec)
    # trim_index: the reference's own test inputs (tests/test_fugue.py:13-56) plus a seeded trim
    trims = []
    graph = [[0, 2, 0.41], [0, 4, 0.85], [3, 4, 0.36], [2, 0, 0.68], [4, 0, 0.1], [4, 3, 0.37]]
    for kwargs in ({"indexed": True}, {"indexed": True, "max_out_deg": 1, "random_seed": 5}):
        df = ArrayDataFrame(graph, schema="src:int,dst:int,weight:double")
        r, nid = ref_fugue.trim_index(NativeExecutionEngine(), df, **kwargs)
        rp = r.as_pandas()
        trims.append({"input": {"src": [g[0] for g in graph], "dst": [g[1] for g in graph], "weight": [g[2] for g in graph]},
                      "kwargs": kwargs, "src": [int(x) for x in rp["src"]], "dst": [int(x) for x in rp["dst"]],
                      "weight": hx(rp["weight"]), "name_id": None})
    dat1 = {"src": ["a1", "a1", "a1", "a2", "b2"], "dst": ["a2", "b1", "b2", "b1", "a2"]}
    for kwargs in ({"indexed": False}, {"indexed": False, "directed": False}, {"indexed": False, "max_out_deg": 2, "random_seed": 3}):
        r, nid = ref_fugue.trim_index(NativeExecutionEngine(), PandasDataFrame(pd.DataFrame(dat1)), **kwargs)
        rp, npd = r.as_pandas(), nid.as_pandas()
        trims.append({"input": dat1, "kwargs": kwargs, "src": [int(x) for x in rp["src"]], "dst": [int(x) for x in rp["dst"]],
                      "weight": hx(rp["weight"]),
                      "name_id": {"vertex_id": [int(x) for x in npd["vertex_id"]], "vertex_name": list(npd["vertex_name"])}})
    return {"native_sum_mode": native_mode_name(), "walks": walks, "trim_index": trims}


def main(out_dir=None):
    out_dir = out_dir or HERE
    rng = random.Random(20261017)
    files = {
        "alias_tables.json": gen_alias(rng),
        "edge_alias_tables.json": gen_edge_alias(rng),
        "samplers.json": gen_samplers(rng),
        "walks.json": gen_walks(rng),
        "indexer.json": gen_indexer(rng),
        "trim.json": gen_trim(rng),
        "fugue_verbatim.json": gen_fugue_verbatim(rng),
    }
    meta = {"python": sys.version.split()[0], "pandas": pd.__version__, "numpy": np.__version__,
            "reference": "graph-embedding/node2vec 0.3.5 (node2vec-fugue), imported from " + REF}
    for name, payload in files.items():
        payload["_meta"] = meta
        with open(os.path.join(out_dir, name), "w") as f:
            json.dump(payload, f, separators=(",", ":"))
        print(name, os.path.getsize(os.path.join(out_dir, name)), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
