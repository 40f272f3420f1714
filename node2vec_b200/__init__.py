"""node2vec_b200 -- B200-native drop-in for the hot path of node2vec-fugue 0.3.5.

Mirrors the reference's module layout for the path it replaces:
``node2vec_b200.fugue`` (trim_index, random_walk), ``node2vec_b200.randomwalk``,
``node2vec_b200.indexer``, ``node2vec_b200.embedding``, ``node2vec_b200.constants``.
All arithmetic on the path runs in hand-written sm_100a CUDA kernels behind the C ABI
declared in ``include/n2v_b200.h``; there is no CPU fallback.
"""
__version__ = "0.1.0"
