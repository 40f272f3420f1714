// Micro-benchmark for the SGNS kernel's table traffic (round 2): one warp gathers a random 512-byte
// row (32 lanes x float4) and -- optionally -- reduces 512 bytes back into another random row with
// red.global.add.v4.f32, over table footprints from 1 GB to most of the HBM.  Question: do random
// row gathers slow down once the table exceeds the TLB reach (config 5 trains 2 x 34 GB tables)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/row_gather scripts/row_gather.cu && /tmp/row_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <bool RED>
__global__ void rows(float* __restrict__ table, uint64_t n_rows, int iters, float* out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  uint64_t x = warp * 0x9E3779B97F4A7C15ull + 12345;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < iters; ++i) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    const uint64_t r = (x >> 20) % n_rows;
    const float4 v = *reinterpret_cast<const float4*>(table + r * 128 + lane * 4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    if (RED) {
      x = x * 6364136223846793005ull + 1442695040888963407ull;
      const uint64_t w = (x >> 20) % n_rows;
      atomicAdd(reinterpret_cast<float4*>(table + w * 128 + lane * 4), make_float4(1e-9f * v.x, 0.f, 0.f, 0.f));
    }
  }
  if (acc.x == 1.2345f) *out = acc.y + acc.z + acc.w;
}

template <bool RED>
void run(float* table, double gb, float* out) {
  const uint64_t n_rows = (uint64_t)(gb * 1e9 / 512);
  const int iters = 128, grid = 148 * 8, block = 256;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  rows<RED><<<grid, block>>>(table, n_rows, iters, out);
  cudaEventRecord(a);
  rows<RED><<<grid, block>>>(table, n_rows, iters, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double n = (double)grid * block / 32 * iters;
  printf("%-18s table %7.1f GB  %7.2f G rows/s  %8.1f GB/s %s\n", RED ? "gather + red.add" : "gather only", gb, n / ms / 1e6,
         n * 512 * (RED ? 2 : 1) / ms / 1e6, RED ? "(read + reduced bytes)" : "");
}

int main() {
  size_t free_b, total_b;
  cudaMemGetInfo(&free_b, &total_b);
  const double max_gb = (double)free_b / 1e9 - 6.0;
  float* table; float* out;
  cudaMalloc(&out, 4);
  if (cudaMalloc(&table, (size_t)(max_gb * 1e9)) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(table, 0, (size_t)(max_gb * 1e9));
  for (double gb : {0.5, 1.0, 4.0, 16.0, 32.0, 64.0, 128.0, 160.0}) {
    if (gb > max_gb) break;
    run<false>(table, gb, out);
    run<true>(table, gb, out);
  }
  return 0;
}
