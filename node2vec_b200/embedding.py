"""Drop-in for the reference's ``node2vec/embedding.py``: same classes, constructor
arguments, parameter-dict semantics and errors -- the Word2Vec underneath is the B200 SGNS
engine (``node2vec_b200.sgns.Word2Vec``) instead of gensim.

``Node2VecGensim`` keeps the reference's class name so that
``from node2vec_b200.embedding import Node2VecGensim`` is a one-line switch;
``Node2VecB200`` is the same class.  ``Node2VecSpark`` (Spark ML Word2Vec: JVM,
hierarchical softmax) is out of scope and raises NotImplementedError.
"""
import logging
import time
from typing import Any, Dict, List, Optional, Union

import numpy as np
import pandas as pd

from .constants import GENSIM_PARAMS
from .sgns import KeyedVectors, Word2Vec


class Node2VecBase(object):
    """Abstract interface (reference embedding.py:22-66)."""

    def __init__(self):
        pass

    def fit(self):
        raise NotImplementedError()

    def embedding(self):
        raise NotImplementedError()

    def get_vector(self, vertex_id: Union[str, int]):
        raise NotImplementedError()

    def save_model(self, file_path: str, file_name: str):
        raise NotImplementedError()

    def load_model(self, file_path: str, file_name: str):
        raise NotImplementedError()


class Node2VecGensim(Node2VecBase):
    """Vertex embedding from random walks (reference embedding.py:70-178).

    :param df_walks: two-column frame [src, walk] (pandas, or the WalkFrame returned by
        ``node2vec_b200.fugue.random_walk`` -- then the walk matrix never leaves HBM)
    :param w2v_params: gensim-3.8 Word2Vec keyword dict; GENSIM_PARAMS defaults are merged
        into it IN PLACE (embedding.py:105-107).  The SGNS hot path needs ``{"sg": 1,
        "negative": K}``; with the reference's defaults (negative=0) nothing is trained.
    :param name_id: optional frame [name, id] to map ids back to names in ``embedding()``
    :param window_size: in [5, 30] else ValueError (embedding.py:109-112)
    :param vector_size: in [32, 1024] else ValueError (embedding.py:113-116)
    :param random_seed: ``seed``; None/0 -> minutes since the epoch (embedding.py:108)
    """

    def __init__(
        self,
        df_walks: Any,
        w2v_params: Dict[str, Any],
        name_id: Optional[pd.DataFrame] = None,
        window_size: Optional[int] = None,
        vector_size: Optional[int] = None,
        random_seed: Optional[int] = None,
    ) -> None:
        super().__init__()
        self.walks, self.name_id = df_walks, name_id
        self.model: Optional[Word2Vec] = None
        # the caller's dict IS the parameter set (the reference merges into it in place, embedding.py:105-107)
        for key, default in GENSIM_PARAMS.items():
            w2v_params.setdefault(key, default)
        w2v_params["seed"] = random_seed or int(time.time()) // 60        # None / 0: minutes since the epoch (:108)
        for value, lo, hi, key, what in ((window_size, 5, 30, "window", "context window size"),
                                         (vector_size, 32, 1024, "size", "vector dimension")):
            if value is None:
                continue
            if not lo <= value <= hi:
                raise ValueError(f"Inappropriate {what} {value}!")
            w2v_params[key] = value
        self.w2v_params = w2v_params
        logging.info("Node2VecGensim: word2vec parameters %s", w2v_params)

    def _sentences(self):
        dev = getattr(self.walks, "walks_device", None)
        if dev is not None:
            return dev
        walks = self.walks["walk"] if not hasattr(self.walks, "as_pandas") else self.walks.as_pandas()["walk"]
        return np.asarray(walks.tolist())

    def fit(self) -> Word2Vec:
        self.model = Word2Vec(sentences=self._sentences(), **self.w2v_params)
        return self.model

    def embedding(self) -> pd.DataFrame:
        if self.model is None:
            raise ValueError("Model is not available. Please run fit()")
        ids = [int(t) for t in self.model.wv.vocab]
        vectors = [list(self.model.wv[t]) for t in self.model.wv.vocab]
        if self.name_id is not None:
            dic = self.name_id.set_index("id").to_dict()["name"]
            names = [dic[i] for i in ids]
            return pd.DataFrame.from_dict({"name": names, "vector": vectors})
        return pd.DataFrame.from_dict({"id": ids, "vector": vectors})

    def get_vector(self, vertex_id: Union[str, int]) -> List[float]:
        if isinstance(vertex_id, int):
            vertex_id = str(vertex_id)
        return list(self.model.wv[vertex_id])  # type: ignore

    def save_model(self, file_path: str, file_name: str) -> None:
        self.model.save(file_path + "/" + file_name + ".model")  # type: ignore

    def load_model(self, file_path: str, file_name: str) -> Word2Vec:
        self.model = Word2Vec.load(file_path + "/" + file_name + ".model")
        return self.model

    def save_vectors(self, file_path: str, file_name: str) -> None:
        self.model.wv.save_word2vec_format(file_path + "/" + file_name)  # type: ignore

    @staticmethod
    def load_vectors(file_path: str, file_name: str) -> KeyedVectors:
        return KeyedVectors.load_word2vec_format(file_path + "/" + file_name)


Node2VecB200 = Node2VecGensim


class Node2VecSpark(Node2VecBase):
    """Out of scope: Spark ML Word2Vec runs in a JVM and optimises hierarchical softmax, a
    different objective from the SGNS hot path."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("Node2VecSpark (Spark ML Word2Vec) is out of scope; use Node2VecGensim")
