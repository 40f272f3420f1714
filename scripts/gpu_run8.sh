#!/usr/bin/env bash
# round-2 GPU session 8 (1 GPU): the driver's N = 1 command lines (both arms), then the whole GPU suite
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2h_bench_1gpu_reference.json 2> gpurun_out/r2h_bench_1gpu_reference.err
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2h_bench_1gpu.json 2> gpurun_out/r2h_bench_1gpu.err; tail -4 gpurun_out/r2h_bench_1gpu.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -2 gpurun_out/r2h_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2h_gpu_tests.log 2>&1; tail -3 gpurun_out/r2h_gpu_tests.log
echo done
