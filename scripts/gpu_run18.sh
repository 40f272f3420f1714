#!/usr/bin/env bash
# round-2 GPU session 18 (1 GPU): validation of the committed tree -- smoke, the whole GPU suite (incl. the
# two-rank one-GPU partitioned check), per-call build times and a short N = 1 bench
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; tail -2 gpurun_out/r2r_smoke.log
timeout 600 python -m pytest tests -m gpu -q --timeout 400 --durations 12 > gpurun_out/r2r_gpu_tests.log 2>&1; tail -22 gpurun_out/r2r_gpu_tests.log
timeout 120 python scripts/build_stages.py rmat20 > gpurun_out/r2r_build_stages.txt 2>&1; cat gpurun_out/r2r_build_stages.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench_1gpu_short.json 2> gpurun_out/r2r_bench_1gpu_short.err; tail -5 gpurun_out/r2r_bench_1gpu_short.err; head -c 600 gpurun_out/r2r_bench_1gpu_short.json
echo done
