#!/usr/bin/env bash
# round-2 GPU session 13 (1 GPU): final validation of the committed tree: smoke + the whole GPU suite
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; tail -2 gpurun_out/r2m_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2m_gpu_tests.log 2>&1; tail -3 gpurun_out/r2m_gpu_tests.log
echo done
