#!/usr/bin/env bash
# round-2 GPU session 6 (1 GPU): SGNS latency-hiding modes A/B, then the whole GPU test suite
mkdir -p gpurun_out
timeout 900 python scripts/sgns_modes.py > gpurun_out/r2f_sgns_modes.txt 2>&1; tail -15 gpurun_out/r2f_sgns_modes.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2f_gpu_tests.log 2>&1; tail -5 gpurun_out/r2f_gpu_tests.log
echo done
