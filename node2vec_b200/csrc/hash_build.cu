// K0b: per-vertex neighbour hash sets (8-slot, 32-byte buckets, load <= 0.5).
// Replaces `set(Neighbors(src_nbs).dst_id)` (reference randomwalk.py:318), which the
// reference rebuilds for every walker at every step, by a one-off build whose lookup is a
// single 32-byte gather in the walk kernel.
#include "n2v_internal.cuh"

namespace {

constexpr int kBlock = 128;

inline int grid_for(int64_t n_warps_needed) {
  const int64_t need = (n_warps_needed * 32 + kBlock - 1) / kBlock;
  const int64_t cap = int64_t(n2v::sm_count()) * 16;
  return static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
}

// A vertex of kHubDeg arcs or more gets a whole 1024-thread CTA (fill_hub_buckets): one warp per
// vertex left the 232k-arc hotspots of BASELINE configs[2] as a 20 ms tail of a sub-millisecond kernel.
constexpr uint32_t kHubDeg = 2048;
constexpr int kHubBlock = 1024;

// Insert x into the table of a vertex with nb buckets.  Slots of a bucket fill in order (a thread
// moves to slot k+1 only after seeing slot k taken) and buckets only ever fill up, so the final
// table satisfies the lookup invariant "x lives in the first bucket of its probe sequence that had
// room" whatever the interleaving of the inserting threads.
__device__ __forceinline__ void insert_neighbour(int32_t* __restrict__ table, uint32_t nb, int32_t x) {
  uint32_t b = __umulhi(static_cast<uint32_t>(x) * N2V_HASH_MULT, nb);
  for (;;) {
    int32_t* slot = table + b * N2V_HASH_SLOTS;
    for (int k = 0; k < N2V_HASH_SLOTS; ++k) {
      const int32_t old = atomicCAS(slot + k, N2V_HASH_EMPTY, x);
      if (old == N2V_HASH_EMPTY || old == x) return;
    }
    b = (b + 1 == nb) ? 0 : b + 1;
  }
}

// One WARP per vertex below kHubDeg: lanes clear the vertex's buckets, then insert its arcs in
// parallel with atomicCAS.  Every vertex gets its hbase here; hubs are only LISTED (their ids cluster --
// R-MAT's heavy vertices are the ids with few set bits -- so the list, not the id range, is what
// fill_hub_buckets strides over).
__global__ void fill_buckets(n2v_vertex_t* __restrict__ vtx, const int32_t* __restrict__ col,
                             int64_t n_vertices, int32_t* __restrict__ hash, int32_t* __restrict__ hubs,
                             unsigned int* __restrict__ n_hubs) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * int64_t(kBlock) + threadIdx.x) >> 5;
  const int64_t n_warps = (int64_t(gridDim.x) * kBlock) >> 5;
  for (int64_t v = warp0; v < n_vertices; v += n_warps) {
    const uint32_t deg = vtx[v].deg, base = vtx[v].base;
    const uint32_t hb = n2v_hash_base(base, static_cast<uint32_t>(v));
    if (lane == 0) vtx[v].hbase = hb;
    if (deg == 0) continue;
    if (deg >= kHubDeg) {
      if (lane == 0) hubs[atomicAdd(n_hubs, 1u)] = static_cast<int32_t>(v);
      continue;
    }
    const uint32_t nb = n2v_hash_nbuckets(deg);
    int32_t* table = hash + static_cast<size_t>(hb) * N2V_HASH_SLOTS;
    for (uint32_t i = lane; i < nb * N2V_HASH_SLOTS; i += 32) table[i] = N2V_HASH_EMPTY;
    __syncwarp();
    const int32_t* c = col + base;
    for (uint32_t i = lane; i < deg; i += 32) {
      const int32_t x = c[i];
      if (i > 0 && c[i - 1] == x) continue;  // multi-arc: one entry per distinct neighbour (col is sorted)
      insert_neighbour(table, nb, x);
    }
    __syncwarp();
  }
}

// One CTA per listed hub (the count stays on the device: the CTAs stride over the list and leave at
// once when it is empty).
__global__ void __launch_bounds__(kHubBlock) fill_hub_buckets(const n2v_vertex_t* __restrict__ vtx,
                                                              const int32_t* __restrict__ col,
                                                              int32_t* __restrict__ hash,
                                                              const int32_t* __restrict__ hubs,
                                                              const unsigned int* __restrict__ n_hubs_ptr) {
  const unsigned int n_hubs = *n_hubs_ptr;
  for (unsigned int h = blockIdx.x; h < n_hubs; h += gridDim.x) {
    const int64_t v = hubs[h];
    const uint32_t deg = vtx[v].deg, base = vtx[v].base;
    const uint32_t nb = n2v_hash_nbuckets(deg);
    int32_t* table = hash + static_cast<size_t>(n2v_hash_base(base, static_cast<uint32_t>(v))) * N2V_HASH_SLOTS;
    for (uint32_t i = threadIdx.x; i < nb * N2V_HASH_SLOTS; i += kHubBlock) table[i] = N2V_HASH_EMPTY;
    __syncthreads();
    const int32_t* c = col + base;
    for (uint32_t i = threadIdx.x; i < deg; i += kHubBlock) {
      const int32_t x = c[i];
      if (i > 0 && c[i - 1] == x) continue;
      insert_neighbour(table, nb, x);
    }
  }
}

}  // namespace

extern "C" int64_t n2v_hash_buckets_bound(int64_t n_arcs, int64_t n_vertices) {
  // vertex v starts at floor(base/4) + v and needs ceil(deg/4) buckets
  return n_arcs / 4 + n_vertices + 1;
}

extern "C" int n2v_hash_build(n2v_vertex_t* vtx, const int32_t* col, int64_t n_vertices, int64_t n_arcs,
                              int32_t* hash, int64_t n_buckets_cap, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_vertices <= 0) return N2V_OK;
  N2V_CHECK_ARG(vtx && (n_arcs == 0 || (col && hash)), "n2v_hash_build: NULL buffer");
  N2V_CHECK_ARG(n_buckets_cap >= n2v_hash_buckets_bound(n_arcs, n_vertices) && n_buckets_cap < (int64_t(1) << 32),
                "n2v_hash_build: bucket capacity %lld out of range", static_cast<long long>(n_buckets_cap));
  N2V_CHECK_ARG((reinterpret_cast<uintptr_t>(hash) & 31) == 0, "n2v_hash_build: hash must be 32-byte aligned");
  // device scratch: the hub count, then the hub list (at most n_arcs / kHubDeg entries)
  const int64_t max_hubs = n_arcs / kHubDeg + 1;
  unsigned int* d_nhubs = nullptr;
  N2V_CUDA(n2v::scratch_alloc(reinterpret_cast<void**>(&d_nhubs), 8 + sizeof(int32_t) * static_cast<size_t>(max_hubs), stream));
  N2V_CUDA(cudaMemsetAsync(d_nhubs, 0, 8, stream));
  int32_t* d_hubs = reinterpret_cast<int32_t*>(d_nhubs + 2);
  fill_buckets<<<grid_for(n_vertices), kBlock, 0, stream>>>(vtx, col, n_vertices, hash, d_hubs, d_nhubs);
  N2V_LAUNCH_OK();
  if (n_arcs >= kHubDeg) {
    const int64_t hub_cap = int64_t(n2v::sm_count()) * 2;
    fill_hub_buckets<<<static_cast<int>(max_hubs < hub_cap ? max_hubs : hub_cap), kHubBlock, 0, stream>>>(
        vtx, col, hash, d_hubs, d_nhubs);
    N2V_LAUNCH_OK();
  }
  N2V_CUDA(cudaFreeAsync(d_nhubs, stream));
  return N2V_OK;
}
