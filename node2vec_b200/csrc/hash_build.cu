// K0b: per-vertex neighbour hash sets (8-slot, 32-byte buckets, load <= 0.5).
// Replaces `set(Neighbors(src_nbs).dst_id)` (reference randomwalk.py:318), which the
// reference rebuilds for every walker at every step, by a one-off build whose lookup is a
// single 32-byte gather in the walk kernel.
#include <cub/device/device_scan.cuh>

#include "n2v_internal.cuh"

namespace {

constexpr int kBlock = 128;

inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

inline int grid_for(int64_t n) {
  const int64_t need = (n + kBlock - 1) / kBlock;
  const int64_t cap = int64_t(n2v::kSmCount) * 16;
  return static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
}

__global__ void count_buckets(const n2v_vertex_t* __restrict__ vtx, int64_t n_vertices,
                              uint32_t* __restrict__ nb) {
  for (int64_t v = blockIdx.x * int64_t(kBlock) + threadIdx.x; v < n_vertices; v += int64_t(gridDim.x) * kBlock)
    nb[v] = n2v_hash_nbuckets(vtx[v].deg);
}

// one thread per vertex, sequential insertion in col[] order => deterministic layout
__global__ void fill_buckets(n2v_vertex_t* __restrict__ vtx, const int32_t* __restrict__ col,
                             const uint32_t* __restrict__ hbase, int64_t n_vertices,
                             int32_t* __restrict__ hash) {
  for (int64_t v = blockIdx.x * int64_t(kBlock) + threadIdx.x; v < n_vertices; v += int64_t(gridDim.x) * kBlock) {
    const uint32_t deg = vtx[v].deg, hb = hbase[v];
    vtx[v].hbase = hb;
    if (deg == 0) continue;
    const uint32_t nb = n2v_hash_nbuckets(deg);
    int32_t* table = hash + static_cast<size_t>(hb) * N2V_HASH_SLOTS;
    for (uint32_t i = 0; i < nb * N2V_HASH_SLOTS; ++i) table[i] = N2V_HASH_EMPTY;
    const int32_t* c = col + vtx[v].base;
    for (uint32_t i = 0; i < deg; ++i) {
      const int32_t x = c[i];
      if (i > 0 && c[i - 1] == x) continue;  // multi-arc: already present (col is sorted)
      uint32_t b = __umulhi(static_cast<uint32_t>(x) * N2V_HASH_MULT, nb);
      for (;;) {
        int32_t* slot = table + b * N2V_HASH_SLOTS;
        int k = 0;
        while (k < N2V_HASH_SLOTS && slot[k] != N2V_HASH_EMPTY) ++k;
        if (k < N2V_HASH_SLOTS) { slot[k] = x; break; }
        b = (b + 1 == nb) ? 0 : b + 1;
      }
    }
  }
}

struct Layout {
  int64_t nb, hbase, cub, total;
  size_t cub_bytes;
};

Layout layout_for(int64_t n_vertices) {
  Layout L{};
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, static_cast<uint32_t*>(nullptr),
                                static_cast<uint32_t*>(nullptr), n_vertices + 1);  // +1: total at the end
  L.cub_bytes = cub_bytes;
  int64_t off = 0;
  L.nb = off; off += align256((n_vertices + 1) * 4);
  L.hbase = off; off += align256((n_vertices + 1) * 4);
  L.cub = off; off += align256(static_cast<int64_t>(cub_bytes));
  L.total = off;
  return L;
}

}  // namespace

extern "C" int64_t n2v_hash_buckets_bound(int64_t n_arcs, int64_t n_vertices) {
  // sum over non-empty vertices of ceil(deg / 4) <= n_arcs / 4 + min(n_vertices, n_arcs)
  return n_arcs / 4 + (n_vertices < n_arcs ? n_vertices : n_arcs) + 1;
}

extern "C" size_t n2v_hash_scratch_bytes(int64_t n_vertices) {
  return static_cast<size_t>(layout_for(n_vertices).total);
}

extern "C" int n2v_hash_build(n2v_vertex_t* vtx, const int32_t* col, int64_t n_vertices, int64_t n_arcs,
                              int32_t* hash, int64_t n_buckets_cap, void* scratch, size_t scratch_bytes,
                              int64_t* n_buckets_host, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_buckets_host) *n_buckets_host = 0;
  if (n_vertices <= 0) return N2V_OK;
  N2V_CHECK_ARG(vtx && scratch && (n_arcs == 0 || (col && hash)), "n2v_hash_build: NULL buffer");
  N2V_CHECK_ARG(n_buckets_cap >= n2v_hash_buckets_bound(n_arcs, n_vertices) && n_buckets_cap < (int64_t(1) << 32),
                "n2v_hash_build: bucket capacity %lld out of range", static_cast<long long>(n_buckets_cap));
  N2V_CHECK_ARG((reinterpret_cast<uintptr_t>(hash) & 31) == 0, "n2v_hash_build: hash must be 32-byte aligned");
  const Layout L = layout_for(n_vertices);
  if (scratch_bytes < static_cast<size_t>(L.total)) {
    n2v::set_error("n2v_hash_build: scratch %zu < required %lld", scratch_bytes, static_cast<long long>(L.total));
    return N2V_ERR_SCRATCH;
  }
  char* base = static_cast<char*>(scratch);
  uint32_t* nb = reinterpret_cast<uint32_t*>(base + L.nb);
  uint32_t* hb = reinterpret_cast<uint32_t*>(base + L.hbase);
  count_buckets<<<grid_for(n_vertices), kBlock, 0, stream>>>(vtx, n_vertices, nb);
  N2V_LAUNCH_OK();
  N2V_CUDA(cudaMemsetAsync(nb + n_vertices, 0, 4, stream));
  size_t cub_bytes = L.cub_bytes;
  N2V_CUDA(cub::DeviceScan::ExclusiveSum(base + L.cub, cub_bytes, nb, hb, n_vertices + 1, stream));
  fill_buckets<<<grid_for(n_vertices), kBlock, 0, stream>>>(vtx, col, hb, n_vertices, hash);
  N2V_LAUNCH_OK();
  uint32_t total = 0;
  N2V_CUDA(cudaMemcpyAsync(&total, hb + n_vertices, 4, cudaMemcpyDeviceToHost, stream));
  N2V_CUDA(cudaStreamSynchronize(stream));
  if (n_buckets_host) *n_buckets_host = total;
  return N2V_OK;
}
