#!/usr/bin/env bash
# round-2 GPU session 20 (1 GPU): ncu launch list (durations only) of the build kernels on configs[2], final tree
mkdir -p gpurun_out
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_build_launches.csv \
  -k regex:'fill_|alias_|unit_scan|check_symmetric|trim_sample|pack_keys|scatter_sorted|close_runs|RadixSort|insert_rows|lookup_rows' \
  python scripts/build_once.py > gpurun_out/r2t_build_once.log 2>&1
tail -2 gpurun_out/r2t_build_once.log; wc -l gpurun_out/r2t_build_launches.csv
echo done
