#!/usr/bin/env bash
# round-2 GPU session 21 (1 GPU): the embedding facade after the constructor rewrite
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_sgns.py -m gpu -q --timeout 30 -k "embedding or reference_test or save or load or vectors or gensim_facade" > gpurun_out/r2u_embedding_tests.log 2>&1; tail -2 gpurun_out/r2u_embedding_tests.log
echo done
