// Internal helpers shared by the .cu files of libn2v_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/n2v_b200.h"

namespace n2v {

void set_error(const char* fmt, ...);

#define N2V_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      n2v::set_error(__VA_ARGS__);          \
      return N2V_ERR_INVALID;               \
    }                                       \
  } while (0)

#define N2V_CUDA(call)                                                                    \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      n2v::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return N2V_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

// launch-error check that does not synchronise
#define N2V_LAUNCH_OK() N2V_CUDA(cudaGetLastError())

// SMs of the current device (B200: 148 = 2 dies x 74); grids are sized in multiples of this.
// Queried once per host thread and device; 148 if the query fails.
inline int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    cached = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    cached_dev = dev;
  }
  return cached;
}

// Small stream-ordered scratch (counters, hub lists).  The device's default memory pool hands every
// free block back to the OS at the next synchronisation unless its release threshold is raised, and each
// entry point that reads a counter back synchronises -- so every build paid a cuMemCreate / cuMemMap
// round trip (milliseconds) for a few hundred KB.  Keep up to 64 MB cached per device instead (raised
// once per host thread and device, never lowered).
inline cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t stream) {
  static thread_local int tuned_dev = -1;
  int dev = 0;
  const cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev != tuned_dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t cur = 0, want = 64ull << 20;
      if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < want)
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &want);
    }
    tuned_dev = dev;
  }
  return cudaMallocAsync(p, bytes, stream);
}

// ---- 16-byte record loads (one LDG.128 each, read-only path) ------------------------
// {base, deg, hbase, wsum-bits} of a vertex in one load
__device__ __forceinline__ uint4 load_vtx(const n2v_vertex_t* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// 32 bytes in ONE request (LDG.E.256, new on sm_100): an arc record or a hash bucket is
// exactly one L2/DRAM sector.  Read-only (.nc) path.
struct Int8 {
  int32_t a[8];
};
__device__ __forceinline__ Int8 load_sector(const void* p) {
  Int8 r;
  asm volatile("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.a[0]), "=r"(r.a[1]), "=r"(r.a[2]), "=r"(r.a[3]), "=r"(r.a[4]), "=r"(r.a[5]), "=r"(r.a[6]),
                 "=r"(r.a[7])
               : "l"(p));
  return r;
}

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based ------------------------------
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0,
                                                        uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

}  // namespace n2v
