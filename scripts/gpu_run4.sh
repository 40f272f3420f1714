#!/usr/bin/env bash
# round-2 GPU session 4 (2 GPUs): multi-GPU check, partitioned bench at scales 22 / 24 / 26, SGNS prefetch A/B on one GPU
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r2d_multi_gpu_check.log 2>&1; tail -4 gpurun_out/r2d_multi_gpu_check.log
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 --rmat-scale 22 > gpurun_out/r2d_bench_2gpu_s22.json 2> gpurun_out/r2d_bench_2gpu_s22.err; tail -3 gpurun_out/r2d_bench_2gpu_s22.err
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --rmat-scale 24 > gpurun_out/r2d_bench_2gpu_s24.json 2> gpurun_out/r2d_bench_2gpu_s24.err; tail -3 gpurun_out/r2d_bench_2gpu_s24.err
for f in 0 1; do
  N2V_SGNS_PREFETCH=$f CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2d_bench_prefetch$f.json 2> gpurun_out/r2d_bench_prefetch$f.err
done
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_sgns.py -m gpu -q --timeout 600 -k "prefetch or sgns or indexer or trim_index or first_occurrence" > gpurun_out/r2d_gpu_tests.log 2>&1; tail -3 gpurun_out/r2d_gpu_tests.log
timeout 1500 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2d_bench_2gpu_s26.json 2> gpurun_out/r2d_bench_2gpu_s26.err; tail -5 gpurun_out/r2d_bench_2gpu_s26.err
nvidia-smi --query-gpu=memory.used --format=csv >> gpurun_out/r2d_bench_2gpu_s26.err
echo done
