// Micro-benchmark for the walk kernel's data layout (round 2): random gathers of GRANULE bytes
// (32 / 64 / 128, granule-aligned, by ONE thread with LDG.256s) over footprints from L2-sized to
// DRAM-sized.  Answers two layout questions: (1) does a 64-byte record cost more DRAM time than a
// 32-byte one (HBM3e access granularity)?  (2) how fast does the random-gather rate fall once the
// footprint exceeds L2 (how much of the graph must be compact to stay L2-resident)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_granule scripts/gather_granule.cu && /tmp/gather_granule
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

struct Int8 { int a[8]; };
__device__ __forceinline__ Int8 ld256(const void* p) {
  Int8 r;
  asm volatile("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.a[0]), "=r"(r.a[1]), "=r"(r.a[2]), "=r"(r.a[3]), "=r"(r.a[4]), "=r"(r.a[5]), "=r"(r.a[6]), "=r"(r.a[7]) : "l"(p));
  return r;
}

template <int SECTORS>
__global__ void gather(const int* __restrict__ buf, uint64_t n_granules, int iters, unsigned long long* out) {
  uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
  unsigned long long acc = 0;
  for (int i = 0; i < iters; ++i) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    const uint64_t g = ((x >> 20) + (acc & 1)) % n_granules;     // dependent chain, like a walker
    const int* p = buf + g * (8 * SECTORS);
#pragma unroll
    for (int s = 0; s < SECTORS; ++s) {
      Int8 v = ld256(p + 8 * s);
      acc += (unsigned)v.a[0] + (unsigned)v.a[7];
    }
  }
  if (acc == 0x1234567) *out = acc;
}

template <int SECTORS>
void run(const int* buf, double gb, unsigned long long* out) {
  const uint64_t n_granules = (uint64_t)(gb * 1e9 / (32 * SECTORS));
  const int iters = 256, grid = 148 * 8, block = 256;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  gather<SECTORS><<<grid, block>>>(buf, n_granules, iters, out);
  cudaEventRecord(a);
  gather<SECTORS><<<grid, block>>>(buf, n_granules, iters, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double loads = (double)grid * block * iters;
  printf("granule %4d B  footprint %7.3f GB  %8.2f G granules/s  %8.1f GB/s\n", 32 * SECTORS, gb, loads / ms / 1e6,
         loads * 32 * SECTORS / ms / 1e6);
}

int main(int argc, char** argv) {
  // optional argument: cudaLimitMaxL2FetchGranularity in bytes (32 / 64 / 128); the default leaves the driver's value
  if (argc > 1) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
    printf("cudaDeviceSetLimit(MaxL2FetchGranularity, %s) -> %s\n", argv[1], cudaGetErrorString(e));
  }
  size_t gran = 0;
  cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
  printf("L2 fetch granularity limit: %zu bytes\n", gran);
  unsigned long long* out; cudaMalloc(&out, 8);
  const double max_gb = 4.0;
  int* buf; cudaMalloc(&buf, (size_t)(max_gb * 1e9) + 4096); cudaMemset(buf, 1, (size_t)(max_gb * 1e9));
  for (double gb : {0.03, 0.12, 0.3, 1.0, 4.0}) {
    run<1>(buf, gb, out);
    run<2>(buf, gb, out);
    run<4>(buf, gb, out);
  }
  return 0;
}
