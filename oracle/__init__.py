"""CPU oracle for the node2vec hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Everything under ``oracle/`` is a CPU restatement of what the reference
(graph-embedding/node2vec 0.3.5, ``/root/reference``) computes on the hot path,
written so the CUDA kernels in ``node2vec_b200/csrc`` can be checked against it.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import or execute anything in this package, and there
only as the checker / the timed CPU baseline.  ``node2vec_b200`` never imports it.

Parity pinning (see DESIGN.md "Oracle"):
  * walk half  -- PINNED: ``tests/golden/*.json`` were produced by importing the
    unmodified reference functions (``tests/golden/make_golden.py``) and the
    restatement reproduces every one of them bit-for-bit, plus the known-answer
    tests of ``tests/test_randomwalk.py`` in the reference.
  * indexer    -- PINNED the same way (reference ``index_graph_pandas`` run under a
    pyspark stub).
  * SGNS half  -- PARITY UNPINNED: the arithmetic lives in gensim ~=3.8.2
    (``requirements.txt:27``), which is neither vendored nor installable here and
    whose numerics the reference's tests never assert (``tests/test_embedding.py:50-62``).
    ``oracle/csrc/sgns_ref.c`` restates the published word2vec/gensim SGNS algorithm.
"""
