#!/usr/bin/env bash
# round-2 GPU session 16 (1 GPU): arc-parallel alias build on all-unit graphs + per-call build times on configs[2]
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 300 --durations 5 -k "alias or hash or replay or golden or zero_weight or c_consumer or facade" > gpurun_out/r2p_alias_tests.log 2>&1; tail -12 gpurun_out/r2p_alias_tests.log
timeout 200 python scripts/build_stages.py rmat20 > gpurun_out/r2p_build_stages.txt 2>&1; cat gpurun_out/r2p_build_stages.txt
echo done
