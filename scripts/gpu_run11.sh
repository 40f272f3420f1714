#!/usr/bin/env bash
# round-2 GPU session 11 (1 GPU): whole GPU suite, ncu sections of the SGNS kernel (default mode) on rmat20, final N = 1 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2k_gpu_tests.log 2>&1; tail -3 gpurun_out/r2k_gpu_tests.log
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:sgns_kernel -s 1 -c 1 -f -o gpurun_out/r2k_sgns_rmat20 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2k_ncu_sgns.log 2>&1
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2k_bench_1gpu.json 2> gpurun_out/r2k_bench_1gpu.err; tail -4 gpurun_out/r2k_bench_1gpu.err
echo done
