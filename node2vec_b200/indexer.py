"""Vertex-name -> integer id remap.  Mirrors ``node2vec/indexer.py`` of the reference.

``index_graph_pandas`` reproduces the reference's pandas indexer bit for bit
(indexer.py:9-49): ids are the position of a name's first occurrence in the
concatenation ``[all src..., all dst...]`` -- sparse, up to 2E-1 -- arcs keep input
order, weights become float64, and the undirected option appends the reversed arcs and
drops exact duplicate triples.  ``index_graph_dense`` offers the Spark indexer's id rule
(dense rank of the sorted names, indexer.py:66-71) for callers that want compact ids.
This is one-off host-side string work, not part of the GPU hot path.
"""
import logging
from typing import Tuple

import numpy as np
import pandas as pd


def _check(df_graph: pd.DataFrame) -> None:
    if "src" not in df_graph.columns or "dst" not in df_graph.columns:
        raise ValueError(f"Input graph NOT in the right format: {df_graph.columns}")


def index_graph_pandas(df_graph: pd.DataFrame, directed: bool) -> Tuple[pd.DataFrame, pd.DataFrame]:
    """Returns (df_edge[src, dst, weight], name_id[vertex_id, vertex_name])."""
    _check(df_graph)
    if "weight" not in df_graph.columns:
        df_graph["weight"] = 1.0          # the reference adds the column to the caller's frame too
    src = df_graph["src"].to_numpy()
    dst = df_graph["dst"].to_numpy()
    weight = df_graph["weight"].to_numpy().astype(np.float64)

    names = pd.Series(np.concatenate([src.astype(object), dst.astype(object)]))
    if src.dtype == dst.dtype and src.dtype != object:
        names = names.astype(src.dtype)
    is_first = ~names.duplicated().to_numpy()
    vertex_id = np.flatnonzero(is_first).astype(np.int64)
    vertex_name = names[is_first]
    name_id = pd.DataFrame({"vertex_id": vertex_id, "vertex_name": vertex_name.to_numpy()})
    logging.info(f"Num of indexed vertices: {len(name_id)}")

    lookup = pd.Series(vertex_id, index=pd.Index(vertex_name.to_numpy()))
    df_edge = pd.DataFrame({
        "src": lookup.reindex(src).to_numpy(),
        "dst": lookup.reindex(dst).to_numpy(),
        "weight": weight,
    })
    logging.info(f"Num of indexed edges: {len(df_edge)}")
    if directed is not True:
        mirror = pd.DataFrame({"src": df_edge["dst"], "dst": df_edge["src"], "weight": df_edge["weight"]})
        df_edge = pd.concat([df_edge, mirror]).drop_duplicates()
    return df_edge, name_id


def index_graph_dense(df_graph: pd.DataFrame, directed: bool) -> Tuple[pd.DataFrame, pd.DataFrame]:
    """The Spark indexer's id rule (indexer.py:66-71): id = rank of the name among the
    sorted distinct names; name table has columns [name, id]."""
    _check(df_graph)
    weight = df_graph["weight"].to_numpy().astype(np.float64) if "weight" in df_graph.columns \
        else np.ones(len(df_graph))
    names = np.unique(np.concatenate([df_graph["src"].to_numpy(), df_graph["dst"].to_numpy()]))
    name_id = pd.DataFrame({"name": names, "id": np.arange(len(names), dtype=np.int64)})
    df_edge = pd.DataFrame({
        "src": np.searchsorted(names, df_graph["src"].to_numpy()),
        "dst": np.searchsorted(names, df_graph["dst"].to_numpy()),
        "weight": weight,
    })
    if directed is not True:
        mirror = pd.DataFrame({"src": df_edge["dst"], "dst": df_edge["src"], "weight": df_edge["weight"]})
        df_edge = pd.concat([df_edge, mirror]).drop_duplicates()
    return df_edge, name_id


def index_graph_spark(df_graph, directed: bool):
    """The reference's Spark indexer (indexer.py:52-85) runs on a SparkSession, which is outside
    the hot path; ``index_graph_dense`` applies its id rule (dense rank of the sorted names) to a
    pandas frame."""
    raise NotImplementedError("index_graph_spark needs Spark (out of scope): use index_graph_dense for the same ids")
