// K5: hotspot trimming, bit-exact with the reference's sampler.
//
// trim_hotspot_vertices (reference randomwalk.py:238-262) keeps, for a vertex with more than
// max_out_degree out-arcs, `DataFrame.sample(n=max_out_degree, random_state=seed)` of them, which
// pandas evaluates as numpy's legacy  RandomState(seed).permutation(deg)[:n]  -- a Fisher-Yates
// shuffle of arange(deg) from the top index down, each swap partner drawn by masked rejection from
// 32-bit MT19937 outputs (numpy random_interval).  Every partition re-creates RandomState(seed), so
// the kept POSITIONS depend on (seed, deg) only.
//
// One warp per hot vertex.  The shuffle is a sequential algorithm, but almost none of its steps
// depend on each other, and the warp exploits that 32 steps at a time:
//   * MT19937: the 624-word refill runs 32 words per step (a word depends on its right neighbour's OLD
//     value and on a word 227 places back / 397 ahead, so ascending 32-word chunks with a
//     read-then-write barrier are exact); lane l tempers draw l of the batch;
//   * rejection: draw l is accepted iff (draw & mask(i)) <= i where i = top index - (accepted draws
//     before l).  Start from "all accepted" and iterate ballot -> prefix count -> decision: after t
//     rounds the first t lanes are final, a fixed point is the sequential answer; two rounds almost always;
//   * swaps: the batch's (i, j) pairs touch distinct positions unless two partners coincide or a
//     partner lands in the batch's own index range [i_low, i_top] -- then lane 0 replays the batch
//     in order (a few per cent of the batches); otherwise all lanes load, then store, in parallel.
// One L2 round trip per ~24 swaps instead of one per swap: 89 ms -> single-digit ms for the 2^18-arc
// hotspots of BASELINE configs[2].
#include "n2v_internal.cuh"

namespace {

constexpr int kMtN = 624, kMtM = 397;
constexpr int kTrimWarps = 4;
constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ void mt_seed(uint32_t* key, uint32_t seed) {   // numpy mt19937_seed == init_genrand
  for (int pos = 0; pos < kMtN; ++pos) {
    key[pos] = seed;
    seed = 1812433253u * (seed ^ (seed >> 30)) + static_cast<uint32_t>(pos) + 1u;
  }
}

// genrand's state refill by the whole warp; bit-identical to the sequential loop (see above)
__device__ void mt_refill_warp(uint32_t* key, int lane) {
  for (int c = 0; c < kMtN; c += 32) {
    const int i = c + lane;
    uint32_t v = 0;
    if (i < kMtN) {
      const uint32_t y = (key[i] & 0x80000000u) | (key[i + 1 == kMtN ? 0 : i + 1] & 0x7fffffffu);
      v = key[i < kMtN - kMtM ? i + kMtM : i + kMtM - kMtN] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    __syncwarp();
    if (i < kMtN) key[i] = v;
    __syncwarp();
  }
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
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
  if (h >= n_hot) return;                            // warp-uniform
  const int64_t n = deg[h];
  int32_t* arr = scratch + offset[h];
  for (int64_t k = lane; k < n; k += 32) arr[k] = static_cast<int32_t>(k);
  uint32_t* key = mt[warp];
  if (lane == 0) mt_seed(key, seed);
  __syncwarp();
  int pos = kMtN;
  int64_t i_top = n - 1;                             // the shuffle's next index; warp-uniform
  while (i_top >= 1) {
    if (pos == kMtN) {
      mt_refill_warp(key, lane);
      pos = 0;
    }
    const int m = min(32, kMtN - pos);               // draws in this batch
    const bool has = lane < m;
    const uint32_t d = has ? mt_temper(key[pos + lane]) : 0u;
    pos += m;
    // which draws does the sequential rejection loop accept, and for which index?
    unsigned acc = m == 32 ? kFull : ((1u << m) - 1u);
    bool a;
    uint32_t j = 0, i_k = 0;
    for (;;) {
      const int64_t ii = i_top - __popc(acc & ((1u << lane) - 1u));
      a = false;
      if (has && ii >= 1) {
        uint32_t mask = static_cast<uint32_t>(ii);   // smallest 2^k - 1 >= ii
        mask |= mask >> 1;
        mask |= mask >> 2;
        mask |= mask >> 4;
        mask |= mask >> 8;
        mask |= mask >> 16;
        j = d & mask;
        i_k = static_cast<uint32_t>(ii);
        a = j <= i_k;
      }
      const unsigned now = __ballot_sync(kFull, a);
      if (now == acc) break;
      acc = now;
    }
    const int cnt = __popc(acc);
    const uint32_t i_low = static_cast<uint32_t>(i_top - cnt + 1);
    // the batch's swaps
    int32_t x = 0, y = 0;
    if (a) {
      x = __ldcg(arr + j);
      y = __ldcg(arr + i_k);
    }
    const unsigned same_j = __match_any_sync(kFull, a ? j : (0x80000000u | static_cast<uint32_t>(lane)));
    const bool clash = a && ((same_j & ~(1u << lane)) != 0u || (j >= i_low && j != i_k));
    if (__any_sync(kFull, clash)) {                  // rare: replay the batch in order on lane 0
      for (int q = 0; q < 32; ++q) {
        const uint32_t jq = __shfl_sync(kFull, j, q), iq = __shfl_sync(kFull, i_k, q);
        if (lane == 0 && ((acc >> q) & 1u)) {
          const int32_t t = __ldcg(arr + jq);
          arr[jq] = __ldcg(arr + iq);
          arr[iq] = t;
        }
      }
    } else if (a) {
      arr[j] = y;
      arr[i_k] = x;
    }
    __syncwarp();                                    // this batch's stores before the next batch's loads
    i_top -= cnt;
  }
  const int64_t take = n < cap ? n : cap;
  for (int64_t k = lane; k < take; k += 32) picked[h * cap + k] = __ldcg(arr + k);
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
