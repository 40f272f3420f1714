#!/usr/bin/env bash
# round-2 GPU session 12 (1 GPU): SGNS mode 5 (one-pair lookahead): parity test + timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 600 -k "latency_hiding" > gpurun_out/r2l_mode_tests.log 2>&1; tail -3 gpurun_out/r2l_mode_tests.log
N2V_MODES=4,5,2 timeout 900 python scripts/sgns_modes.py > gpurun_out/r2l_sgns_modes.txt 2>&1; tail -16 gpurun_out/r2l_sgns_modes.txt
echo done
