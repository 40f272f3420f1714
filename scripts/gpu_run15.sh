#!/usr/bin/env bash
# round-2 GPU session 15 (1 GPU): hub path of the hash build (one CTA per hub) + per-call build times on configs[2]
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 300 --durations 8 -k "hash or csr or full_size or replay or giant" > gpurun_out/r2o_hash_tests.log 2>&1; tail -14 gpurun_out/r2o_hash_tests.log
timeout 200 python scripts/build_stages.py rmat20 > gpurun_out/r2o_build_stages.txt 2>&1; cat gpurun_out/r2o_build_stages.txt
echo done
