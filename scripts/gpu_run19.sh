#!/usr/bin/env bash
# round-2 GPU session 19 (1 GPU): last sanity of the committed tree -- the experimental shared-negatives kernel's
# gates (arithmetic now in the default suite), the facade / string-name paths, smoke
mkdir -p gpurun_out
N2V_EXPERIMENTAL=1 timeout 150 python -m pytest tests/test_gpu_sgns_shared.py -m gpu -q -s --timeout 120 > gpurun_out/r2s_sgns_shared.log 2>&1; grep -E "AUC|epoch ms|passed|failed" gpurun_out/r2s_sgns_shared.log | tail -5
timeout 100 python -m pytest tests -m gpu -q --timeout 90 -k "facade or trim_index_frame or first_occurrence_on_string or embedding" > gpurun_out/r2s_facade_tests.log 2>&1; tail -2 gpurun_out/r2s_facade_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; tail -1 gpurun_out/r2s_smoke.log
echo done
