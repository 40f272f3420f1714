// Micro-benchmark: random 32-byte gathers from a PEER GPU's memory (single process, cudaMalloc +
// cudaDeviceEnablePeerAccess) as a function of the remote footprint.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/peer_gather scripts/peer_gather.cu && /tmp/peer_gather
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
__global__ void gather(const int* __restrict__ buf, uint64_t n_sectors, int iters, unsigned long long* out) {
  uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
  unsigned long long acc = 0;
  for (int i = 0; i < iters; ++i) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    const uint64_t s = ((x >> 20) + (acc & 1)) % n_sectors;
    Int8 v = ld256(buf + s * 8);
    acc += (unsigned)v.a[0] + (unsigned)v.a[7];
  }
  if (acc == 0x1234567) *out = acc;
}
int main() {
  int n = 0; cudaGetDeviceCount(&n);
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  cudaSetDevice(0);
  cudaDeviceEnablePeerAccess(1, 0);
  unsigned long long* out; cudaMalloc(&out, 8);
  for (double gb : {0.25, 1.0, 2.0, 4.0, 8.0, 16.0}) {
    const uint64_t n_sectors = (uint64_t)(gb * 1e9 / 32);
    int* remote; cudaSetDevice(1); cudaMalloc(&remote, n_sectors * 32); cudaMemset(remote, 1, n_sectors * 32); cudaDeviceSynchronize();
    cudaSetDevice(0);
    const int iters = 64, grid = 148 * 6, block = 256;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather<<<grid, block>>>(remote, n_sectors, iters, out);
    cudaEventRecord(a);
    gather<<<grid, block>>>(remote, n_sectors, iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double loads = (double)grid * block * iters;
    printf("peer gather  footprint %6.2f GB  %8.2f G sectors/s  %8.1f GB/s  (%s)\n", gb, loads / ms / 1e6, loads * 32 / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
    cudaSetDevice(1); cudaFree(remote); cudaSetDevice(0);
  }
  return 0;
}
