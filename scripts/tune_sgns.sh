#!/bin/bash
# Kernel-tuning helper (run under gpurun): SGNS kernel occupancy variants.
set -e
cd "$(dirname "$0")/.."
CS=node2vec_b200/csrc
for mb in ${MB_LIST:-1 3 4 5 6}; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -shared \
       -cudart static -DN2V_SGNS_MIN_BLOCKS=$mb ${EXTRA_NVCC} -o /tmp/libn2v_mb$mb.so \
       $CS/abi.cu $CS/peer_mem.cu $CS/csr_build.cu $CS/hash_build.cu $CS/alias_build.cu $CS/walk.cu $CS/vocab.cu $CS/sgns.cu
  echo "== min blocks/SM $mb"
  N2V_B200_LIB=/tmp/libn2v_mb$mb.so python scripts/sgns_experiments.py | grep "atomic=True"
done
