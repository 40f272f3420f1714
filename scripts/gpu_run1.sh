#!/usr/bin/env bash
# round-2 GPU session 1 (1 GPU): micro-benchmark, GPU tests, bench N=1, DP-AUC study, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt 2>&1
free -g >> gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_granule scripts/gather_granule.cu && timeout 300 /tmp/gather_granule > gpurun_out/r2_gather_granule.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_gpu_tests.log 2>&1
tail -5 gpurun_out/r2_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -3 gpurun_out/r2_bench1.err
timeout 300 python scripts/dp_averaging_auc.py --graph blog --epochs 3 > gpurun_out/r2_dp_blog.txt 2>&1
timeout 300 python scripts/dp_averaging_auc.py --graph er10k --epochs 3 > gpurun_out/r2_dp_er.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_under_ncu.log 2>&1
echo done
