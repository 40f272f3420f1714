"""CPU-only checks of the boundary: the shared library loads, exports every symbol that
include/n2v_b200.h declares, host-only entry points behave, and the host facade validates
arguments like the reference does.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "n2v_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = "\n".join(l for l in text.splitlines() if "static inline" not in l)   # header-only helpers
    return sorted(set(re.findall(r"\b(n2v_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from node2vec_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    names = _declared_functions()
    assert "n2v_walk" in names and "n2v_alias_build" in names and len(names) >= 9
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/n2v_b200.h but not exported"


def test_abi_version_and_struct_sizes(lib):
    from node2vec_b200 import _lib
    assert lib.n2v_abi_version() == 7
    assert C.sizeof(_lib.GraphPart) == 48
    assert C.sizeof(_lib.Graph) == 8 + 8 + 4 + 4 + 8 + 48 * 16
    assert C.sizeof(_lib.WalkConsts) == 40


def test_header_is_plain_c_and_c_consumer_links():
    """The boundary is a C ABI: the header must compile as C11 with all warnings as errors, and a
    C program using every walk-path entry point must link against the library (it RUNS in the gpu
    suite, tests/test_gpu_walk.py::test_plain_c_consumer)."""
    import subprocess
    from node2vec_b200 import build as nb
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c",
                           os.path.join(root, "include", "n2v_b200.h")])
    exe = nb.build_c_consumer(force=True)
    assert os.access(exe, os.X_OK)
    needed = subprocess.check_output(["readelf", "-d", exe], text=True)
    assert "libn2v_b200.so" in needed and "python" not in needed.lower() and "torch" not in needed.lower()


def test_walk_consts_host_only(lib):
    from node2vec_b200 import graph
    from oracle import clib
    for p, q in [(1.0, 1.0), (1.0, 0.5), (0.25, 4.0), (4.0, 0.25), (0.5, 0.5), (1e-3, 1e3), (3.0, 7.0)]:
        for flags, ratio in ((0, False), (7, False), (3, False), (5, True), (0, True)):
            a, b = graph.walk_consts(p, q, flags, ratio), clib.walk_consts(p, q, flags, ratio)
            assert (a.t_ret, a.t_nbr, a.t_far, a.fold_mode, a.max_trials) == \
                   (b.t_ret, b.t_nbr, b.t_far, b.fold_mode, b.max_trials)
            assert a.fold_gain == b.fold_gain and a.mix_qm1 == b.mix_qm1
            assert 1 <= a.t_ret <= 2 ** 32 and 1 <= a.t_nbr <= 2 ** 32 and 1 <= a.t_far <= 2 ** 32
    c = graph.walk_consts(1.0, 1.0, 0)
    assert c.t_ret == c.t_nbr == c.t_far == 2 ** 32 and c.fold_mode == 0
    c = graph.walk_consts(0.25, 1.0, 7)                     # fold: envelope stays at max(1, 1/q) = 1
    assert c.fold_mode == 1 and c.t_nbr == 2 ** 32 and c.t_far == 2 ** 32 and abs(c.fold_gain - 3.0) < 1e-6
    c = graph.walk_consts(0.25, 4.0, 7)                     # unit symmetric simple graph, q > 1: mixture sampler
    assert c.fold_mode == 3 and c.t_ret == 2 ** 32 and c.fold_gain == 15.0 and c.mix_qm1 == 3.0
    c = graph.walk_consts(8.0, 2.0, 7)                      # p > q > 1: no return excess, x == prev thinned in the bulk
    assert c.fold_mode == 3 and c.t_ret == 2 ** 30 and c.fold_gain == 0.0 and c.mix_qm1 == 1.0
    c = graph.walk_consts(0.25, 4.0, 5)                     # not symmetric: no mixture
    assert c.fold_mode == 0
    c = graph.walk_consts(0.25, 4.0, 0)                     # no fold possible: envelope 1/p = 4
    assert c.fold_mode == 0 and c.t_ret == 2 ** 32 and c.t_nbr == 2 ** 30 and c.t_far == 2 ** 28
    c = graph.walk_consts(0.25, 4.0, 0, True)               # general fold through per-arc ratios
    assert c.fold_mode == 2 and c.t_nbr == 2 ** 32 and c.t_far == 2 ** 30 and abs(c.fold_gain - 3.0) < 1e-6
    with pytest.raises(ValueError):
        graph.walk_consts(0.0, 1.0, 0)
    with pytest.raises(ValueError):
        graph.walk_consts(1.0, 0.0, 0)
    assert b"Zero return" in lib.n2v_last_error()


def test_product_has_no_cpu_fallback():
    """Without a GPU every compute entry of the host facade must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from node2vec_b200 import _lib, fugue, randomwalk
    with pytest.raises(_lib.N2VError):
        randomwalk.generate_alias_tables([0.5, 0.8, 1.0])
    df = pd.DataFrame({"src": [0, 1], "dst": [1, 0], "weight": [1.0, 1.0]})
    with pytest.raises(_lib.N2VError):
        fugue.random_walk(None, df, {})


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "node2vec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libn2v_oracle" not in text and "oracle._build" not in text, f


def test_facade_argument_validation_matches_reference():
    from node2vec_b200 import fugue, randomwalk
    from node2vec_b200.constants import GENSIM_PARAMS, NODE2VEC_PARAMS, WORD2VEC_PARAMS
    df = pd.DataFrame({"dst": ["a2", "b1"], "weight": [0.8, 1.1]})
    with pytest.raises(ValueError):
        fugue.trim_index(None, df, False)
    good = pd.DataFrame({"src": [0, 1], "dst": [1, 0], "weight": [1.0, 1.0]})
    params = {"num_walks": 2}
    with pytest.raises(ValueError):
        fugue.random_walk(None, good, params, good)                 # walk_seed lacks "id"
    assert params["walk_length"] == 20 and params["return_param"] == 1.0   # defaults merged in place first
    with pytest.raises(ValueError):
        randomwalk.generate_edge_alias_tables(0, set(), ([1, 2], [1.0]))
    with pytest.raises(ValueError):
        randomwalk.generate_edge_alias_tables(0, set(), ([1], [1.0]), 0)
    # tests/test_constants.py of the reference
    assert isinstance(NODE2VEC_PARAMS["num_walks"], int) and isinstance(NODE2VEC_PARAMS["walk_length"], int)
    assert isinstance(NODE2VEC_PARAMS["return_param"], float) and isinstance(NODE2VEC_PARAMS["inout_param"], float)
    assert isinstance(WORD2VEC_PARAMS["stepSize"], float)
    for k in ["minCount", "numPartitions", "maxIter", "maxSentenceLength", "windowSize", "vectorSize"]:
        assert isinstance(WORD2VEC_PARAMS[k], int)
    assert isinstance(GENSIM_PARAMS["alpha"], float)
    for k in ["min_count", "iter", "batch_words", "window", "size", "negative", "workers"]:
        assert isinstance(GENSIM_PARAMS[k], int)


def test_value_types_reference_goldens():
    """tests/test_randomwalk.py:16-49,53-90,94-128 serialisation goldens (pickle protocol 3)."""
    from node2vec_b200.randomwalk import AliasProb, Neighbors, RandomPath
    idx, weight = [0, 1, 2], [1.0, 0.2, 1.4]
    code64 = "gANdcQAoSwBLAUsCZV1xAShHP/AAAAAAAABHP8mZmZmZmZpHP/ZmZmZmZmZlhnECLg=="
    for nbs in (Neighbors((idx, weight)), Neighbors(code64)):
        assert nbs.dst_id == idx and nbs.dst_wt == weight
        assert list(nbs.items()) == [(0, 1.0), (1, 0.2), (2, 1.4)]
        assert nbs.serialize() == code64
        assert nbs.as_pandas().equals(pd.DataFrame({"dst": idx, "weight": weight}))
    nbs = Neighbors(pd.DataFrame({"dst": [1, 2, 3], "weight": [0.1, 1.2, 0.8]}))
    assert nbs.serialize() == "gANdcQAoSwFLAksDZV1xAShHP7mZmZmZmZpHP/MzMzMzMzNHP+mZmZmZmZplhnECLg=="
    code = "gANdcQAoSwFLAGVdcQEoRz/lVVVVVVVVRz/wAAAAAAAAZYZxAi4="
    for jq in (AliasProb(([1, 0], [0.6666666666666666, 1.0])), AliasProb(code),
               AliasProb(pd.DataFrame({"alias": [1, 0], "probs": [0.6666666666666666, 1.0]}))):
        assert jq.alias == [1, 0] and jq.probs == [0.6666666666666666, 1.0] and jq.serialize() == code
    for path, code in [([-1, 0], "gANdcQAoSv////9LAGUu"), ([2, 1], "gANdcQAoSwJLAWUu"), ([0, 3], "gANdcQAoSwBLA2Uu")]:
        for rp in (RandomPath(path), RandomPath(code)):
            assert rp.path == path and rp.last_edge == (path[-2], path[-1])
            assert rp.serialize() == code and str(rp) == str(path)


def test_host_transformers():
    from node2vec_b200.randomwalk import get_vertex_neighbors, initiate_random_walk, to_path, trim_hotspot_vertices
    df = pd.DataFrame({"src": [3, 3, 3], "dst": [0, 1, 2], "weight": [1.0, 0.2, 1.4]})
    res = next(iter(get_vertex_neighbors(df)))
    assert res["id"] == 3
    assert res["neighbors"] == "gANdcQAoSwBLAUsCZV1xAShHP/AAAAAAAABHP8mZmZmZmZpHP/ZmZmZmZmZlhnECLg=="
    rows = list(trim_hotspot_vertices(df))
    assert [r["dst"] for r in rows] == [0, 1, 2]
    assert len(list(trim_hotspot_vertices(df, max_out_degree=2, random_seed=20))) == 2
    it = iter(initiate_random_walk([{"id": 3}, {"id": 2}], 3))
    for s in (3, 2):
        for i in range(3):
            a = next(it)
            assert (a["src"], a["dst"], a["path"]) == (-1 - i, s, [-1 - i, s])
    out = list(to_path([{"src": 0, "dst": 2, "path": [1, 0, 2]}]))
    assert out == [{"src": 1, "walk": [1, 0, 2]}]


def test_indexer_facade_bit_exact():
    from node2vec_b200.indexer import index_graph_dense, index_graph_pandas
    from tests.helpers import load_golden
    for c in load_golden("indexer.json")["cases"]:
        cols = {"src": c["src"], "dst": c["dst"]}
        if c["weight"] is not None:
            cols["weight"] = c["weight"]
        e, n = index_graph_pandas(pd.DataFrame(cols), c["directed"])
        assert e["src"].tolist() == c["edge_src"] and e["dst"].tolist() == c["edge_dst"]
        assert [float(x).hex() for x in e["weight"]] == c["edge_weight"]
        assert n["vertex_id"].tolist() == c["vertex_id"] and n["vertex_name"].tolist() == c["vertex_name"]
    # reference tests/test_indexer.py:8-25
    g = pd.DataFrame({"src": ["a1", "a2", "a3", "a4"], "dst": ["a2", "b1", "b2", "a1"]})
    e, vid = index_graph_pandas(g, False)
    assert len(vid) == 6 and len(e) == 8
    with pytest.raises(ValueError):
        index_graph_pandas(pd.DataFrame({"dst": ["a"], "weight": [1.0]}), True)
    from node2vec_b200.indexer import index_graph_spark
    with pytest.raises(NotImplementedError):
        index_graph_spark(g, False)
    e, vid = index_graph_dense(pd.DataFrame({"src": ["b", "a"], "dst": ["c", "b"]}), True)
    assert vid["id"].tolist() == [0, 1, 2] and vid["name"].tolist() == ["a", "b", "c"]
    assert e["src"].tolist() == [1, 0] and e["dst"].tolist() == [2, 1]


def test_trim_index_facade_matches_verbatim_reference():
    """node2vec_b200.fugue.trim_index (host pandas code) against node2vec.fugue.trim_index outputs
    recorded by tests/golden/make_golden.py::gen_fugue_verbatim."""
    from node2vec_b200.fugue import trim_index
    from tests.helpers import load_golden
    for c in load_golden("fugue_verbatim.json")["trim_index"]:
        inp = dict(c["input"])
        res, name_id = trim_index(None, pd.DataFrame(inp), **c["kwargs"])
        out = res.as_pandas()
        assert out["src"].tolist() == c["src"] and out["dst"].tolist() == c["dst"], c["kwargs"]
        assert [float(x).hex() for x in out["weight"]] == c["weight"]
        if c["name_id"] is None:
            assert name_id is None
        else:
            nid = name_id.as_pandas()
            assert nid["vertex_id"].tolist() == c["name_id"]["vertex_id"]
            assert nid["vertex_name"].tolist() == c["name_id"]["vertex_name"]


def test_example_pipeline_index_stage(tmp_path):
    """examples/pipeline.py, stage 1 (host only): parquet in, indexed graph + name map out."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("n2v_example_pipeline", os.path.join(root, "examples", "pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    pd.DataFrame({"src": ["a", "a", "b", "c", "a"], "dst": ["b", "c", "c", "a", "b"],
                  "weight": [1.0, 2.0, 0.5, 1.5, 1.0]}).to_parquet(tmp_path / "input_graph.parquet")
    mod.stage_index(str(tmp_path))
    g = pd.read_parquet(tmp_path / "graph_indexed.parquet")
    nid = pd.read_parquet(tmp_path / "graph_name2id.parquet")
    assert list(g.columns) == ["src", "dst", "weight"] and len(g) == 4                    # the duplicate row is gone
    # the reference's ids: position of a name's first occurrence in [src..., dst...] (indexer.py:26-35)
    assert dict(zip(nid["vertex_name"], nid["vertex_id"])) == {"a": 0, "b": 2, "c": 3}
    assert sorted(zip(g["src"], g["dst"])) == [(0, 2), (0, 3), (2, 3), (3, 0)]


def test_reference_test_trim_index_against_facade_with_fugue_like_frames():
    """The engine-independent half of the reference's tests/test_fugue.py::test_trim_index (:13-56), run
    against THIS package with Fugue-shaped inputs (the pandas stand-in for Fugue's ArrayDataFrame /
    PandasDataFrame / NativeExecutionEngine under tests/golden/fugue_shim): same calls, same assertions."""
    import sys
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fugue_shim")
    sys.path.insert(0, shim)
    try:
        from fugue import ArrayDataFrame, NativeExecutionEngine, PandasDataFrame
    finally:
        sys.path.remove(shim)
    from node2vec_b200.fugue import trim_index

    graph = [[0, 2, 0.41], [0, 4, 0.85], [3, 4, 0.36], [2, 0, 0.68], [4, 0, 0.1], [4, 3, 0.37]]
    df = ArrayDataFrame(graph, schema="src:int,dst:int,weight:double")
    df_res, name_id = trim_index(NativeExecutionEngine(), df, indexed=True)
    assert len(df_res.as_pandas()) == 6 and name_id is None
    df_res, name_id = trim_index(NativeExecutionEngine(), df, indexed=True, max_out_deg=1)
    assert len(df_res.as_pandas()) == 4 and name_id is None
    dat1 = {"src": ["a1", "a1", "a1", "a2", "b2"], "dst": ["a2", "b1", "b2", "b1", "a2"]}
    dat2 = {"dst": ["a2", "b1", "b2", "a1"], "weight": [0.8, 1.1, 1.0, 0.3]}
    df_res, name_id = trim_index(NativeExecutionEngine(), PandasDataFrame(pd.DataFrame.from_dict(dat1)), indexed=False)
    assert len(df_res.as_pandas()) == 5 and len(name_id.as_pandas()) == 4
    df_res, name_id = trim_index(NativeExecutionEngine(), PandasDataFrame(pd.DataFrame.from_dict(dat1)),
                                 indexed=False, max_out_deg=2)
    assert df_res.count() == 4 and name_id.count() == 4                 # the Spark half's numbers (:35-38)
    pytest.raises(ValueError, trim_index, NativeExecutionEngine(), PandasDataFrame(pd.DataFrame.from_dict(dat2)), False)

    class SparkExecutionEngine:                                          # Spark engines are refused, loudly
        pass
    pytest.raises(NotImplementedError, trim_index, SparkExecutionEngine(), df, True)


def test_keyed_vectors_word2vec_text_format(tmp_path):
    """The embedding wire format (embedding.py:166-178 -> wv.save_word2vec_format / load_word2vec_format):
    header "<n> <dim>", one "<token> <floats>" line per vertex, most frequent first; fp32 survives the text
    round trip bit for bit (shortest-repr floats); loaded counts follow gensim's n - i convention."""
    import numpy as np
    from node2vec_b200.sgns import KeyedVectors, Vocab
    rng = np.random.default_rng(0)
    kv = KeyedVectors(6)
    kv.vectors = rng.standard_normal((4, 6)).astype(np.float32) * np.float32(1e-3)
    kv.index2word = ["17", "3", "250", "8"]
    kv.vocab = {w: Vocab(count=100 - 10 * i, index=i) for i, w in enumerate(kv.index2word)}
    path = str(tmp_path / "vectors.txt")
    kv.save_word2vec_format(path)
    lines = open(path).read().splitlines()
    assert lines[0] == "4 6" and [ln.split(" ")[0] for ln in lines[1:]] == kv.index2word
    assert all(len(ln.split(" ")) == 7 for ln in lines[1:])
    back = KeyedVectors.load_word2vec_format(path)
    assert back.index2word == kv.index2word and back.vector_size == 6 and len(back) == 4
    assert back.vectors.dtype == np.float32 and np.array_equal(back.vectors, kv.vectors)
    assert [back.vocab[w].count for w in back.index2word] == [4, 3, 2, 1]
    assert np.array_equal(back["250"], kv.vectors[2]) and np.array_equal(back[250], kv.vectors[2])
    assert np.array_equal(back[["3", "8"]], kv.vectors[[1, 3]])
    assert "17" in back and 17 in back and "18" not in back
    with pytest.raises(KeyError):
        back["18"]


def test_walk_frame_parquet_round_trip(tmp_path):
    """The [src, walk] wire format (host logic only: a WalkFrame over a CPU tensor)."""
    import numpy as np
    import torch
    from node2vec_b200.fugue import WalkFrame, read_walks_parquet
    walks = np.array([[3, 1, 4, 1], [5, 9, 2, 6], [5, 3, 5, 8]], dtype=np.int32)
    wf = WalkFrame(torch.as_tensor(walks))
    assert wf.count() == 3 and wf.schema == ["src", "walk"]
    df = wf.as_pandas()
    assert df["src"].tolist() == [3, 5, 5] and df["walk"].tolist() == walks.tolist()
    path = str(tmp_path / "walks.parquet")
    wf.to_parquet(path)
    back = read_walks_parquet(path)
    assert back["src"].tolist() == [3, 5, 5] and back["walk"].tolist() == walks.tolist()


def test_walk_matrix_inputs_host_logic():
    """What Node2VecGensim hands the SGNS engine: decimal-string tokens (embedding.py:125), ints, tensors;
    ragged walks are rejected (the reference's np.array(...) would not build a matrix either)."""
    import numpy as np
    import torch
    from node2vec_b200.sgns import _walk_matrix, neg_top_entries, NEG_CHUNK
    cpu = torch.device("cpu")
    a = _walk_matrix(np.array([["3", "1", "4"], ["1", "5", "9"]]), cpu)
    assert a.dtype == torch.int32 and a.tolist() == [[3, 1, 4], [1, 5, 9]]
    assert _walk_matrix([[1, 2], [3, 4]], cpu).tolist() == [[1, 2], [3, 4]]
    t = torch.tensor([[7, 8, 9]], dtype=torch.int64)
    assert _walk_matrix(t, cpu).dtype == torch.int32
    with pytest.raises(ValueError):
        _walk_matrix([[1, 2, 3], [4, 5]], cpu)
    assert neg_top_entries(65536) == 0 and neg_top_entries(65537) == 9 and neg_top_entries(1 << 26) == (1 << 26) // NEG_CHUNK


def test_node2vec_gensim_constructor_contract_needs_no_gpu():
    """embedding.py:75-118 of the reference: defaults merged into the CALLER's dict, seed = random_seed or
    minutes since the epoch, window in [5, 30] and size in [32, 1024] else ValueError -- raised after the
    earlier keys were already written, as the reference does."""
    import pandas as pd
    from node2vec_b200.constants import GENSIM_PARAMS
    from node2vec_b200.embedding import Node2VecGensim
    df = pd.DataFrame({"src": [0], "walk": [[0, 1]]})
    p = {}
    m = Node2VecGensim(df, p, window_size=5, vector_size=32, random_seed=7)
    assert m.w2v_params is p and p["seed"] == 7 and p["window"] == 5 and p["size"] == 32
    assert all(k in p for k in GENSIM_PARAMS)
    p2 = {"iter": 3}
    Node2VecGensim(df, p2)
    assert p2["iter"] == 3 and p2["seed"] > 0 and p2["window"] == GENSIM_PARAMS["window"]
    for kw in ({"window_size": 4}, {"window_size": 31}, {"vector_size": 31}, {"vector_size": 1025}):
        with pytest.raises(ValueError):
            Node2VecGensim(df, {}, **kw)
    q = {}
    with pytest.raises(ValueError):
        Node2VecGensim(df, q, window_size=6, vector_size=8)
    assert q["window"] == 6 and "seed" in q and "size" in q
    with pytest.raises(ValueError):
        m.embedding()                                   # before fit()
