#!/usr/bin/env bash
# round-2 GPU session 17 (1 GPU): warp-parallel trim_sample (bit-exact vs numpy, timing) + build times with the retained scratch pool
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -s --timeout 150 -k "trim" > gpurun_out/r2q_trim_tests.log 2>&1; grep -E "n2v_trim_sample|passed|failed|Error|error" gpurun_out/r2q_trim_tests.log | tail -12
timeout 200 python scripts/build_stages.py rmat20 > gpurun_out/r2q_build_stages.txt 2>&1; cat gpurun_out/r2q_build_stages.txt
echo done
