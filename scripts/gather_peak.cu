// Micro-benchmark: random 32-byte gather throughput of HBM3e on B200 (the denominator that
// bounds a walk kernel: every trial is one random sector).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_peak scripts/gather_peak.cu && /tmp/gather_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Int8 { int a[8]; };
__device__ __forceinline__ Int8 ld256(const void* p) {
  Int8 r;
  asm volatile("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.a[0]), "=r"(r.a[1]), "=r"(r.a[2]), "=r"(r.a[3]), "=r"(r.a[4]), "=r"(r.a[5]), "=r"(r.a[6]), "=r"(r.a[7]) : "l"(p));
  return r;
}

// DEP = 1: each load's address depends on the previous load (pointer chase, one in flight per thread)
// DEP = 0: independent addresses from a hash (UNROLL in flight per thread)
template <int DEP, int UNROLL>
__global__ void gather(const int* __restrict__ buf, uint64_t n_sectors, int iters, unsigned long long* out) {
  uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
  unsigned long long acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (DEP) {
      x = x * 6364136223846793005ull + 1442695040888963407ull;
      const uint64_t s = ((x >> 20) + (acc & 1)) % n_sectors;
      Int8 v = ld256(buf + s * 8);
      acc += (unsigned)v.a[0] + (unsigned)v.a[7];
    } else {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        const uint64_t s = (x >> 20) % n_sectors;
        Int8 v = ld256(buf + s * 8);
        acc += (unsigned)v.a[0] + (unsigned)v.a[7];
      }
    }
  }
  if (acc == 0x1234567) *out = acc;
}

template <int DEP, int UNROLL>
void run(const char* name, const int* buf, uint64_t n_sectors, int blocks_per_sm, unsigned long long* out) {
  const int iters = DEP ? 256 : 256 / UNROLL;
  const int grid = 148 * blocks_per_sm, block = 256;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  gather<DEP, UNROLL><<<grid, block>>>(buf, n_sectors, iters, out);
  cudaEventRecord(a);
  gather<DEP, UNROLL><<<grid, block>>>(buf, n_sectors, iters, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double loads = (double)grid * block * (DEP ? iters : iters * UNROLL);
  printf("%-34s footprint %6.2f GB  blocks/SM %d  %8.2f G sectors/s  %8.1f GB/s\n", name, n_sectors * 32 / 1e9,
         blocks_per_sm, loads / ms / 1e6, loads * 32 / ms / 1e6);
}

int main() {
  unsigned long long* out; cudaMalloc(&out, 8);
  for (double gb : {0.03, 1.0, 4.0, 16.0}) {
    const uint64_t n_sectors = (uint64_t)(gb * 1e9 / 32);
    int* buf; cudaMalloc(&buf, n_sectors * 32); cudaMemset(buf, 1, n_sectors * 32);
    for (int bps : {6, 8}) {
      run<1, 1>("dependent chain (1 in flight)", buf, n_sectors, bps, out);
      run<0, 4>("independent x4", buf, n_sectors, bps, out);
      run<0, 8>("independent x8", buf, n_sectors, bps, out);
    }
    cudaFree(buf);
  }
  return 0;
}
