#!/usr/bin/env bash
# round-2 GPU session 10 (all visible GPUs): the partitioned bench, 10 timed steps
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
timeout 1200 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2j_bench_${N}gpu.json 2> gpurun_out/r2j_bench_${N}gpu.err; tail -4 gpurun_out/r2j_bench_${N}gpu.err
echo done
