#!/usr/bin/env python
"""The reference's three-stage example (examples/fugue_spark.py: index -> walk -> embed, parquet in,
parquet out) on node2vec_b200.  Same calls, same parameter dictionaries; the Spark session and the
Fugue engine are gone (the compute engine argument is accepted and ignored).

    python examples/pipeline.py index DIR      # DIR/input_graph.parquet [src, dst(, weight)] -> graph_indexed / graph_name2id
    python examples/pipeline.py walk  DIR      # graph_indexed.parquet -> graph_walks.parquet [src, walk]
    python examples/pipeline.py embed DIR      # graph_walks (+ graph_name2id) -> graph_embedding.parquet [name|id, vector]
    python examples/pipeline.py all   DIR
"""
import logging
import os
import sys

import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from node2vec_b200.embedding import Node2VecGensim                  # noqa: E402
from node2vec_b200.fugue import random_walk, read_walks_parquet, trim_index   # noqa: E402

logging.basicConfig(format="%(asctime)s %(levelname)s:: %(message)s", level=logging.INFO)

# refer to constants.NODE2VEC_PARAMS / WORD2VEC_PARAMS for the defaults (same as the reference's)
N2V_PARAMS = {"num_walks": 30, "walk_length": 10, "return_param": 1.0, "inout_param": 1.0}
W2V_PARAMS = {"sg": 1, "negative": 5, "min_count": 1, "iter": 5}


def stage_index(d: str, max_out_deg: int = 10000) -> None:
    df = pd.read_parquet(f"{d}/input_graph.parquet").drop_duplicates()
    # assume the input graph is not indexed, and is directed (as the reference's example does)
    df_index, name_id = trim_index(None, df, indexed=False, directed=True, max_out_deg=max_out_deg)
    name_id.as_pandas().to_parquet(f"{d}/graph_name2id.parquet")
    df_index.as_pandas().to_parquet(f"{d}/graph_indexed.parquet")


def stage_walk(d: str, random_seed=None) -> None:
    df = pd.read_parquet(f"{d}/graph_indexed.parquet").drop_duplicates()
    walks = random_walk(None, df, n2v_params=dict(N2V_PARAMS), random_seed=random_seed)
    walks.to_parquet(f"{d}/graph_walks.parquet")          # [src, walk], straight from the walk matrix


def stage_embed(d: str, random_seed=None) -> None:
    df_walk = read_walks_parquet(f"{d}/graph_walks.parquet")
    name_id = None
    if os.path.exists(f"{d}/graph_name2id.parquet"):
        # the indexer names its columns (vertex_id, vertex_name); Node2Vec* expects (id, name)
        name_id = pd.read_parquet(f"{d}/graph_name2id.parquet").rename(columns={"vertex_id": "id", "vertex_name": "name"})
    g2v = Node2VecGensim(df_walk, dict(W2V_PARAMS), name_id, window_size=5, vector_size=128, random_seed=random_seed)
    g2v.fit()
    logging.info("model fitting done!")
    g2v.embedding().to_parquet(f"{d}/graph_embedding.parquet")
    g2v.save_vectors(d, "graph_vectors.txt")


if __name__ == "__main__":
    stage = sys.argv[1] if len(sys.argv) > 1 else "all"
    where = sys.argv[2] if len(sys.argv) > 2 else "."
    if stage in ("index", "all"):
        stage_index(where)
    if stage in ("walk", "all"):
        stage_walk(where)
    if stage in ("embed", "all"):
        stage_embed(where)
