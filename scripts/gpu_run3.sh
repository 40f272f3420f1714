#!/usr/bin/env bash
# round-2 GPU session 3 (1 GPU): L2 fetch granularity experiments, row-gather micro-benchmark, re-run of fixed tests
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_granule scripts/gather_granule.cu
for g in "" 32 64 128; do timeout 200 /tmp/gather_granule $g; done > gpurun_out/r2c_gather_granule_l2fetch.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/row_gather scripts/row_gather.cu && timeout 300 /tmp/row_gather > gpurun_out/r2c_row_gather.txt 2>&1
for g in 0 32 64; do
  N2V_L2_FETCH=$g timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_l2fetch$g.json 2> gpurun_out/r2c_bench_l2fetch$g.err
done
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_walk.py -m gpu -q --timeout 900 > gpurun_out/r2c_gpu_tests.log 2>&1
tail -5 gpurun_out/r2c_gpu_tests.log
echo done
