// K5: hotspot trimming, bit-exact with the reference's sampler.
//
// trim_hotspot_vertices (reference randomwalk.py:238-262) keeps, for a vertex with more than
// max_out_degree out-arcs, `DataFrame.sample(n=max_out_degree, random_state=seed)` of them, which
// pandas evaluates as numpy's legacy  RandomState(seed).permutation(deg)[:n]  -- a Fisher-Yates
// shuffle of arange(deg) from the top index down, each swap partner drawn by masked rejection from
// 32-bit MT19937 outputs (numpy random_interval).  Every partition re-creates RandomState(seed), so
// the kept POSITIONS depend on (seed, deg) only.  The construction is inherently sequential per
// vertex; hot vertices are few, so each gets one warp: all lanes fill arange and copy the result,
// lane 0 runs the generator (state in shared memory) and the swaps.
#include "n2v_internal.cuh"

namespace {

constexpr int kMtN = 624, kMtM = 397;
constexpr int kTrimWarps = 4;

__device__ void mt_seed(uint32_t* key, uint32_t seed) {   // numpy mt19937_seed == init_genrand
  for (int pos = 0; pos < kMtN; ++pos) {
    key[pos] = seed;
    seed = 1812433253u * (seed ^ (seed >> 30)) + static_cast<uint32_t>(pos) + 1u;
  }
}

__device__ void mt_refill(uint32_t* key) {
  int i = 0;
  for (; i < kMtN - kMtM; ++i) {
    const uint32_t y = (key[i] & 0x80000000u) | (key[i + 1] & 0x7fffffffu);
    key[i] = key[i + kMtM] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  for (; i < kMtN - 1; ++i) {
    const uint32_t y = (key[i] & 0x80000000u) | (key[i + 1] & 0x7fffffffu);
    key[i] = key[i + (kMtM - kMtN)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  const uint32_t y = (key[kMtN - 1] & 0x80000000u) | (key[0] & 0x7fffffffu);
  key[kMtN - 1] = key[kMtM - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_next(uint32_t* key, int& pos) {
  if (pos == kMtN) {
    mt_refill(key);
    pos = 0;
  }
  uint32_t y = key[pos++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

__global__ void __launch_bounds__(kTrimWarps * 32)
trim_sample_kernel(const int64_t* __restrict__ deg, const int64_t* __restrict__ offset, int64_t n_hot, int32_t cap,
                   uint32_t seed, int32_t* __restrict__ scratch, int32_t* __restrict__ picked) {
  __shared__ uint32_t mt[kTrimWarps][kMtN];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t h = static_cast<int64_t>(blockIdx.x) * kTrimWarps + warp;
  if (h >= n_hot) return;
  const int64_t n = deg[h];
  int32_t* arr = scratch + offset[h];
  for (int64_t k = lane; k < n; k += 32) arr[k] = static_cast<int32_t>(k);
  __syncwarp();
  if (lane == 0) {
    uint32_t* key = mt[warp];
    mt_seed(key, seed);
    int pos = kMtN;
    for (int64_t i = n - 1; i >= 1; --i) {
      uint32_t mask = static_cast<uint32_t>(i);     // smallest 2^k - 1 >= i
      mask |= mask >> 1;
      mask |= mask >> 2;
      mask |= mask >> 4;
      mask |= mask >> 8;
      mask |= mask >> 16;
      uint32_t j;
      do {
        j = mt_next(key, pos) & mask;
      } while (j > static_cast<uint32_t>(i));
      const int32_t t = arr[j];
      arr[j] = arr[i];
      arr[i] = t;
    }
  }
  __syncwarp();
  const int64_t take = n < cap ? n : cap;
  for (int64_t k = lane; k < take; k += 32) picked[h * cap + k] = arr[k];
}

}  // namespace

extern "C" int n2v_trim_sample(const int64_t* deg, const int64_t* scratch_offset, int64_t n_hot, int32_t cap,
                               uint32_t seed, int32_t* scratch, int32_t* picked, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(n_hot >= 0 && cap >= 1, "n2v_trim_sample: bad sizes (n_hot %lld, cap %d)", static_cast<long long>(n_hot), cap);
  if (n_hot == 0) return N2V_OK;
  N2V_CHECK_ARG(deg && scratch_offset && scratch && picked, "n2v_trim_sample: NULL buffer");
  const int64_t grid = (n_hot + kTrimWarps - 1) / kTrimWarps;
  N2V_CHECK_ARG(grid <= 0x7fffffff, "n2v_trim_sample: too many hot vertices");
  trim_sample_kernel<<<static_cast<unsigned int>(grid), kTrimWarps * 32, 0, stream>>>(deg, scratch_offset, n_hot, cap, seed,
                                                                                      scratch, picked);
  N2V_LAUNCH_OK();
  return N2V_OK;
}
