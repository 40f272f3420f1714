// K0: arc list -> CSR sorted by (src, dst), stable.
// Replaces `partition(by=["src"], presort="dst")` + get_vertex_neighbors
// (reference fugue.py:130, randomwalk.py:266-275).
#include <cub/device/device_radix_sort.cuh>

#include "n2v_internal.cuh"

namespace {

constexpr int kBlock = 256;

inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

inline int grid_for(int64_t n, int per_sm = 8) {
  const int64_t need = (n + kBlock - 1) / kBlock;
  const int64_t cap = int64_t(n2v::sm_count()) * per_sm;
  return static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
}

__global__ void pack_keys(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                          int64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx,
                          int64_t n_vertices, int64_t n_dst_vertices, unsigned int* __restrict__ bad) {
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock) {
    const int32_t s = src[i], d = dst[i];
    if (s < 0 || d < 0 || s >= n_vertices || d >= n_dst_vertices) atomicOr(bad, 1u);
    keys[i] = (static_cast<uint64_t>(static_cast<uint32_t>(s)) << 32) | static_cast<uint32_t>(d);
    idx[i] = static_cast<uint32_t>(i);
  }
}

// run starts -> vtx.base ; also col / weight / perm gather and the SIMPLE / UNIT flags
__global__ void scatter_sorted(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx,
                               const double* __restrict__ weight, int64_t n,
                               n2v_vertex_t* __restrict__ vtx, int32_t* __restrict__ col,
                               double* __restrict__ w_sorted, int64_t* __restrict__ perm,
                               unsigned int* __restrict__ not_flags, const unsigned int* __restrict__ bad) {
  if (*bad) return;   // pack_keys saw an id outside the graph: nothing below may index vtx[] with it
  unsigned int local = 0;
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock) {
    const uint64_t k = keys[i];
    const uint32_t s = static_cast<uint32_t>(k >> 32);
    col[i] = static_cast<int32_t>(static_cast<uint32_t>(k));
    const uint32_t j = idx[i];
    const double w = weight ? weight[j] : 1.0;
    w_sorted[i] = w;
    if (perm) perm[i] = j;
    if (w != 1.0) local |= N2V_GRAPH_UNIT_WEIGHT;
    if (i > 0) {
      const uint64_t kp = keys[i - 1];
      if (kp == k) local |= N2V_GRAPH_SIMPLE;
      if (static_cast<uint32_t>(kp >> 32) != s) vtx[s].base = static_cast<uint32_t>(i);
    } else {
      vtx[s].base = 0;
    }
  }
  if (local) atomicOr(not_flags, local);
}

// run ends -> vtx.deg (needs every base written: separate launch)
__global__ void close_runs(const uint64_t* __restrict__ keys, int64_t n, n2v_vertex_t* __restrict__ vtx,
                           const unsigned int* __restrict__ bad) {
  if (*bad) return;
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock) {
    const uint32_t s = static_cast<uint32_t>(keys[i] >> 32);
    if (i + 1 == n || static_cast<uint32_t>(keys[i + 1] >> 32) != s)
      vtx[s].deg = static_cast<uint32_t>(i + 1) - vtx[s].base;
  }
}

// SYMMETRIC: every arc (a,b,w) has a mirror (b,a,w).  Only meaningful on SIMPLE graphs (distinct arcs),
// which is what makes a ONE-SIDED search exact: order the two ends of an arc by (degree, id); the arc
// whose head is the smaller end is the pair's "searcher" and looks its mirror up in the head's row --
// the SHORTER of the two rows, so an arc between a leaf and a 232k-arc hub costs a probe or two instead
// of 18 -- and every other arc (self-loops aside) must be the mirror some searcher found.  Mirrors of
// distinct searchers are distinct arcs, so "every search succeeded and #searchers == #others" means the
// mirror map is a bijection.  balance accumulates #searchers - #others.
__global__ void check_symmetric(const uint64_t* __restrict__ keys, const double* __restrict__ w_sorted,
                                int64_t n, const n2v_vertex_t* __restrict__ vtx,
                                unsigned int* __restrict__ not_flags, const unsigned int* __restrict__ bad_ids,
                                unsigned long long* __restrict__ balance) {
  if (*bad_ids) return;
  bool bad = false;
  long long local = 0;
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock) {
    const uint64_t k = keys[i];
    const uint32_t a = static_cast<uint32_t>(k >> 32), b = static_cast<uint32_t>(k);
    if (a == b) continue;                              // a self-loop is its own mirror
    const uint32_t deg_a = vtx[a].deg, deg_b = vtx[b].deg;
    if (!(deg_b < deg_a || (deg_b == deg_a && b < a))) {
      --local;                                         // to be found by its mirror
      continue;
    }
    ++local;
    const uint64_t want = (static_cast<uint64_t>(b) << 32) | a;
    const uint64_t lo0 = vtx[b].base;
    uint64_t lo = lo0, hi = lo0 + deg_b;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      if (keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    if (lo >= lo0 + deg_b || keys[lo] != want || w_sorted[lo] != w_sorted[i]) bad = true;
  }
  if (bad) atomicOr(not_flags, N2V_GRAPH_SYMMETRIC);
  if (local) atomicAdd(balance, static_cast<unsigned long long>(local));
}

struct Layout {
  int64_t keys_a, keys_b, idx_a, idx_b, flags, cub, total;
  size_t cub_bytes;
};

Layout layout_for(int64_t n_arcs) {
  Layout L{};
  size_t cub_bytes = 0;
  cub::DoubleBuffer<uint64_t> dk(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, dk, dv, n_arcs > 0 ? n_arcs : 1, 0, 64);
  L.cub_bytes = cub_bytes;
  int64_t off = 0;
  L.keys_a = off; off += align256(n_arcs * 8);
  L.keys_b = off; off += align256(n_arcs * 8);
  L.idx_a = off; off += align256(n_arcs * 4);
  L.idx_b = off; off += align256(n_arcs * 4);
  L.flags = off; off += 256;
  L.cub = off; off += align256(static_cast<int64_t>(cub_bytes));
  L.total = off;
  return L;
}

}  // namespace

extern "C" size_t n2v_csr_scratch_bytes(int64_t n_arcs, int64_t /*n_vertices*/) {
  return static_cast<size_t>(layout_for(n_arcs).total);
}

extern "C" int n2v_csr_build(const int32_t* src, const int32_t* dst, const double* weight,
                             int64_t n_arcs, int64_t n_vertices, int64_t n_dst_vertices, n2v_vertex_t* vtx,
                             int32_t* col, double* weight_sorted, int64_t* perm, void* scratch,
                             size_t scratch_bytes, uint32_t* flags_host, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(n_arcs >= 0 && n_vertices >= 0, "n2v_csr_build: negative size");
  N2V_CHECK_ARG(n_arcs < (int64_t(1) << 32), "n2v_csr_build: %lld arcs exceed the 2^32 per-part limit",
                static_cast<long long>(n_arcs));
  N2V_CHECK_ARG(n_vertices <= (int64_t(1) << 31) && n_dst_vertices <= (int64_t(1) << 31) && n_dst_vertices >= 0,
                "n2v_csr_build: vertex ids must fit int32");
  const bool replicated = n_dst_vertices == n_vertices;
  N2V_CHECK_ARG(vtx && (n_arcs == 0 || (src && dst && col && weight_sorted && scratch)),
                "n2v_csr_build: NULL buffer");
  const Layout L = layout_for(n_arcs);
  if (scratch_bytes < static_cast<size_t>(L.total)) {
    n2v::set_error("n2v_csr_build: scratch %zu < required %lld", scratch_bytes, static_cast<long long>(L.total));
    return N2V_ERR_SCRATCH;
  }
  N2V_CUDA(cudaMemsetAsync(vtx, 0, sizeof(n2v_vertex_t) * static_cast<size_t>(n_vertices), stream));
  uint32_t flags = N2V_GRAPH_UNIT_WEIGHT | N2V_GRAPH_SYMMETRIC | N2V_GRAPH_SIMPLE;
  if (n_arcs == 0) {
    if (flags_host) *flags_host = flags;
    return N2V_OK;
  }
  char* base = static_cast<char*>(scratch);
  uint64_t* keys_a = reinterpret_cast<uint64_t*>(base + L.keys_a);
  uint64_t* keys_b = reinterpret_cast<uint64_t*>(base + L.keys_b);
  uint32_t* idx_a = reinterpret_cast<uint32_t*>(base + L.idx_a);
  uint32_t* idx_b = reinterpret_cast<uint32_t*>(base + L.idx_b);
  unsigned int* dflags = reinterpret_cast<unsigned int*>(base + L.flags);  // [0]=bad ids, [1]=NOT-flags
  N2V_CUDA(cudaMemsetAsync(dflags, 0, 256, stream));

  const int grid = grid_for(n_arcs);
  pack_keys<<<grid, kBlock, 0, stream>>>(src, dst, n_arcs, keys_a, idx_a, n_vertices, n_dst_vertices, dflags);
  N2V_LAUNCH_OK();

  int end_bit = 32;
  while (end_bit < 64 && (int64_t(1) << (end_bit - 32)) < n_vertices) ++end_bit;  // src bits; dst uses all low 32
  cub::DoubleBuffer<uint64_t> dk(keys_a, keys_b);
  cub::DoubleBuffer<uint32_t> dv(idx_a, idx_b);
  size_t cub_bytes = L.cub_bytes;
  N2V_CUDA(cub::DeviceRadixSort::SortPairs(base + L.cub, cub_bytes, dk, dv, n_arcs, 0, end_bit, stream));

  scatter_sorted<<<grid, kBlock, 0, stream>>>(dk.Current(), dv.Current(), weight, n_arcs, vtx, col,
                                              weight_sorted, perm, dflags + 1, dflags);
  N2V_LAUNCH_OK();
  close_runs<<<grid, kBlock, 0, stream>>>(dk.Current(), n_arcs, vtx, dflags);
  N2V_LAUNCH_OK();
  if (replicated) {
    check_symmetric<<<grid, kBlock, 0, stream>>>(dk.Current(), weight_sorted, n_arcs, vtx, dflags + 1, dflags,
                                                 reinterpret_cast<unsigned long long*>(dflags + 2));
    N2V_LAUNCH_OK();
  }

  unsigned int h[4] = {0, 0, 0, 0};                    // [0] bad ids, [1] NOT-flags, [2..3] searcher balance (u64)
  N2V_CUDA(cudaMemcpyAsync(h, dflags, sizeof(h), cudaMemcpyDeviceToHost, stream));
  N2V_CUDA(cudaStreamSynchronize(stream));
  N2V_CHECK_ARG(h[0] == 0, "n2v_csr_build: vertex id outside [0, %lld) / [0, %lld)",
                static_cast<long long>(n_vertices), static_cast<long long>(n_dst_vertices));
  flags &= ~h[1];
  if (h[2] != 0 || h[3] != 0) flags &= ~uint32_t(N2V_GRAPH_SYMMETRIC);   // an arc nobody's mirror search found
  if (!replicated) flags &= ~uint32_t(N2V_GRAPH_SYMMETRIC);
  if (!(flags & N2V_GRAPH_SIMPLE)) flags &= ~uint32_t(N2V_GRAPH_SYMMETRIC);
  if (flags_host) *flags_host = flags;
  return N2V_OK;
}
