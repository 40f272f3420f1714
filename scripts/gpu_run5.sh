#!/usr/bin/env bash
# round-2 GPU session 5 (N GPUs = all visible): the driver's own command line for the partitioned workload
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 1500 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2e_bench_${N}gpu.json 2> gpurun_out/r2e_bench_${N}gpu.err; tail -6 gpurun_out/r2e_bench_${N}gpu.err
timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2e_bench_${N}gpu_reference.json 2> gpurun_out/r2e_bench_${N}gpu_reference.err
echo done
