#!/usr/bin/env bash
# round-2 GPU session 2 (1 GPU): tests, bench N=1 with the mixture sampler, launch list, ncu captures on rmat20
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2b_gpu_tests.log 2>&1
tail -5 gpurun_out/r2b_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -f -o gpurun_out/r2b_walk_rmat20 python bench.py --steps 1 --warmup 3 --no-sgns --no-cpu-baseline --no-secondary > gpurun_out/r2b_ncu_walk.log 2>&1
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --clock-control none -k regex:sgns_kernel -s 1 -c 1 -f -o gpurun_out/r2b_sgns_rmat20 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2b_ncu_sgns.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
