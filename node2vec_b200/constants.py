"""Default parameter dicts -- same keys, values and Python types as the reference's
``node2vec/constants.py:6-68`` (they are part of the boundary contract and are merged
into the caller's dict in place, fugue.py:120-122, embedding.py:105-107)."""
from typing import Any, Dict

# default cap on a vertex's out-degree (constants.py:6)
MAX_OUT_DEGREES: int = 100000

# kept for signature compatibility; partitioning is a Spark notion (constants.py:10)
NUM_PARTITIONS: int = 3000

# constants.py:14-27
NODE2VEC_PARAMS: Dict[str, Any] = {
    "num_walks": 10,        # walks started from every vertex
    "walk_length": 20,      # steps per walk (a walk has walk_length + 1 vertices)
    "return_param": 1.0,    # p
    "inout_param": 1.0,     # q
}

# constants.py:31-46 (Spark ML names; accepted by Node2VecSpark-style callers)
WORD2VEC_PARAMS: Dict[str, Any] = {
    "minCount": 10,
    "numPartitions": 100,
    "stepSize": 0.025,
    "maxIter": 10,
    "seed": None,
    "maxSentenceLength": 10000,
    "windowSize": 5,
    "vectorSize": 128,
}

# constants.py:50-68 (gensim 3.8 names)
GENSIM_PARAMS: Dict[str, Any] = {
    "min_count": 10,
    "alpha": 0.025,
    "iter": 10,
    "seed": None,
    "batch_words": 1000,
    "window": 5,
    "size": 128,
    "negative": 0,
    "workers": 16,
}
