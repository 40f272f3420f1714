#!/bin/bash
# Kernel-tuning helper (run under gpurun): builds libn2v_b200 variants with different
# walk-kernel occupancy targets and times the walk on the bench workload.
set -e
cd "$(dirname "$0")/.."
CS=node2vec_b200/csrc
for bps in ${BPS_LIST:-8 6 5}; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -shared \
       -cudart static -DN2V_WALK_BLOCKS_PER_SM=$bps ${EXTRA_NVCC} -o /tmp/libn2v_bps$bps.so \
       $CS/abi.cu $CS/peer_mem.cu $CS/csr_build.cu $CS/hash_build.cu $CS/alias_build.cu $CS/walk.cu $(ls $CS/vocab.cu $CS/sgns.cu 2>/dev/null)
  echo "== blocks/SM $bps"
  N2V_B200_LIB=/tmp/libn2v_bps$bps.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('value %.3e steps/s  kernel_ms %.3f  T %.2f probes/trial %.2f  e2e %.3e' % (d['value'], r['kernel_ms'], r['trials_per_step'], r['probes_per_trial'], d['e2e']['value']))"
done
