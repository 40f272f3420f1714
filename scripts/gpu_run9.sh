#!/usr/bin/env bash
# round-2 GPU session 9 (2 GPUs): multi-GPU check (incl. localize, random_walk(process_group), DP AUC), bench N = 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r2i_multi_gpu_check.log 2>&1; tail -8 gpurun_out/r2i_multi_gpu_check.log
timeout 1200 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err; tail -5 gpurun_out/r2i_bench_2gpu.err
echo done
