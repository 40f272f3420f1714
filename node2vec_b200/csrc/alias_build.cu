// K1: per-vertex first-order alias tables, bit-exact fp64 against the reference's
// generate_alias_tables (randomwalk.py:157-190); a4: generate_edge_alias_tables
// (randomwalk.py:193-232) for explicit (prev, cur) pairs; a5/a6: the two samplers
// (randomwalk.py:70-99) on explicit fp64 uniforms.
//
// The construction is inherently sequential per vertex (LIFO work-lists, donor pushed
// back, left-to-right fp64 sum), so one thread owns one vertex; vertices are independent.
// fp64 is kept un-contracted (explicit __d*_rn intrinsics) so every intermediate rounds
// exactly as CPython's float arithmetic does.
#include "alias_core.cuh"
#include "n2v_internal.cuh"

namespace {

constexpr int kBlock = 128;

__device__ __forceinline__ uint32_t prob_to_thr(double pr) {
  // u32 < thr  <=>  u32 / 2^32 < pr   (exact for pr < 1); pr >= 1 saturates (alias_dst == dst there)
  if (pr >= 1.0) return 0xFFFFFFFFu;
  const double s = ceil(pr * 4294967296.0);
  return s >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(s);
}

__global__ void alias_build_kernel(n2v_vertex_t* __restrict__ vtx, const n2v_vertex_t* lookup,
                                   const int32_t* __restrict__ col,
                                   const double* __restrict__ weight, int64_t n_vertices, int sum_mode,
                                   int32_t* __restrict__ alias, double* __restrict__ probs,
                                   n2v_arc_t* __restrict__ arcs, int32_t* __restrict__ scratch,
                                   unsigned long long* __restrict__ n_zero) {
  for (int64_t v = blockIdx.x * int64_t(kBlock) + threadIdx.x; v < n_vertices;
       v += int64_t(gridDim.x) * kBlock) {
    const uint64_t base = vtx[v].base;
    const uint32_t n = vtx[v].deg;
    if (n == 0) continue;
    double* pr = probs + base;
    n2v_arc_t* out = arcs + base;
    // left-to-right fp64 sum of the raw weights, kept as float for the weighted return-edge fold
    double wsum = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
      const double w = weight[base + i];
      pr[i] = w;
      wsum = __dadd_rn(wsum, w);
    }
    vtx[v].wsum = static_cast<float>(wsum);
    const bool ok = n2v::build_alias_one(pr, n, sum_mode, scratch + base,
                                    [&](uint32_t i, int32_t a) { out[i].alias_idx = a; });
    if (!ok) {
      atomicAdd(n_zero, 1ull);
      for (uint32_t i = 0; i < n; ++i) {
        pr[i] = 0.0;
        const int32_t x = col[base + i];
        out[i] = n2v_arc_t{0xFFFFFFFFu, x, x, 0, lookup[x].base, lookup[x].deg, lookup[x].base, lookup[x].deg};
        if (alias) alias[base + i] = 0;
      }
      continue;
    }
    for (uint32_t i = 0; i < n; ++i) {
      const int32_t a = out[i].alias_idx;
      const double p = pr[i];
      const int32_t self = col[base + i];
      n2v_arc_t rec;
      rec.thr = prob_to_thr(p);
      rec.dst = self;
      rec.alias_dst = (p >= 1.0) ? self : col[base + a];
      rec.alias_idx = a;
      // adjacency headers of both possible landing vertices (deg is final; base too: K0 ran before)
      rec.dst_base = lookup[rec.dst].base;
      rec.dst_deg = lookup[rec.dst].deg;
      rec.adst_base = lookup[rec.alias_dst].base;
      rec.adst_deg = lookup[rec.alias_dst].deg;
      out[i] = rec;
      if (alias) alias[base + i] = a;
    }
  }
}

// membership of x in the ascending id slice col[0..n)
__device__ __forceinline__ bool contains_sorted(const int32_t* __restrict__ col, uint32_t n, int32_t x) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (col[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo < n && col[lo] == x;
}

__global__ void edge_alias_kernel(const n2v_vertex_t* __restrict__ vtx, const int32_t* __restrict__ col,
                                  const double* __restrict__ weight, const int32_t* __restrict__ prev,
                                  const int32_t* __restrict__ cur, int64_t n_pairs, double p, double q,
                                  int sum_mode, const int64_t* __restrict__ out_offset,
                                  int32_t* __restrict__ alias_out, double* __restrict__ probs_out,
                                  int32_t* __restrict__ scratch) {
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n_pairs;
       i += int64_t(gridDim.x) * kBlock) {
    const int32_t t = prev[i], v = cur[i];
    const uint64_t base = vtx[v].base;
    const uint32_t n = vtx[v].deg;
    if (n == 0) continue;
    const int64_t off = out_offset[i];
    double* pr = probs_out + off;
    int32_t* al = alias_out + off;
    const int32_t* tcol = t >= 0 ? col + vtx[t].base : nullptr;
    const uint32_t tdeg = t >= 0 ? vtx[t].deg : 0;
    for (uint32_t k = 0; k < n; ++k) {
      const int32_t x = col[base + k];
      const double w = weight[base + k];
      double bw;
      if (t < 0) bw = w;                                       // first step: unbiased (:320-321)
      else if (x == t) bw = __ddiv_rn(w, p);                   // back to prev        (:223-224)
      else if (contains_sorted(tcol, tdeg, x)) bw = w;         // into N_out(prev)    (:226-227)
      else bw = __ddiv_rn(w, q);                               // anywhere else       (:229-230)
      pr[k] = bw;
    }
    n2v::build_alias_one(pr, n, sum_mode, scratch + off, [&](uint32_t k, int32_t a) { al[k] = a; });
  }
}

__global__ void alias_draw_kernel(const int32_t* __restrict__ alias, const double* __restrict__ probs,
                                  const int64_t* __restrict__ offset, const double* __restrict__ first,
                                  const double* __restrict__ second, int64_t n_draws,
                                  int32_t* __restrict__ picked) {
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n_draws;
       i += int64_t(gridDim.x) * kBlock) {
    const int64_t off = offset[i];
    const int64_t n = offset[i + 1] - off;
    const double r1 = first[i];
    const double scaled = __dmul_rn(r1, static_cast<double>(n));  // int(r1 * n) == int(n * r1)
    const int64_t k = static_cast<int64_t>(scaled);               // truncation toward zero
    double gate;
    if (second) gate = second[i];                                 // two-uniform sampler (:95-99)
    else gate = __dsub_rn(scaled, static_cast<double>(k));        // one-uniform "wiki" sampler (:79-84)
    picked[i] = gate < probs[off + k] ? static_cast<int32_t>(k) : alias[off + k];
  }
}

inline int grid_for(int64_t n) {
  const int64_t need = (n + kBlock - 1) / kBlock;
  const int64_t cap = int64_t(n2v::kSmCount) * 16;
  return static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace

extern "C" int n2v_alias_build(n2v_vertex_t* vtx, const n2v_vertex_t* vtx_lookup, const int32_t* col,
                               const double* weight_sorted, int64_t n_vertices, int64_t n_arcs, int sum_mode,
                               int32_t* alias, double* probs, n2v_arc_t* arcs, int32_t* scratch,
                               int64_t* n_zero_host, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(sum_mode == N2V_SUM_NAIVE || sum_mode == N2V_SUM_NEUMAIER, "n2v_alias_build: bad sum_mode %d", sum_mode);
  N2V_CHECK_ARG(n_vertices >= 0 && n_arcs >= 0, "n2v_alias_build: negative size");
  if (n_zero_host) *n_zero_host = 0;
  if (n_arcs == 0 || n_vertices == 0) return N2V_OK;
  N2V_CHECK_ARG(vtx && col && weight_sorted && probs && arcs && scratch, "n2v_alias_build: NULL buffer");
  unsigned long long* d_zero = nullptr;
  N2V_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_zero), sizeof(unsigned long long), stream));
  N2V_CUDA(cudaMemsetAsync(d_zero, 0, sizeof(unsigned long long), stream));
  alias_build_kernel<<<grid_for(n_vertices), kBlock, 0, stream>>>(vtx, vtx_lookup ? vtx_lookup : vtx, col,
                                                                   weight_sorted, n_vertices, sum_mode,
                                                                   alias, probs, arcs, scratch, d_zero);
  N2V_LAUNCH_OK();
  unsigned long long h = 0;
  N2V_CUDA(cudaMemcpyAsync(&h, d_zero, sizeof(h), cudaMemcpyDeviceToHost, stream));
  N2V_CUDA(cudaFreeAsync(d_zero, stream));
  N2V_CUDA(cudaStreamSynchronize(stream));
  if (n_zero_host) *n_zero_host = static_cast<int64_t>(h);
  if (h != 0) {
    n2v::set_error("n2v_alias_build: %llu vertices have out-weights summing to zero", h);
    return N2V_ERR_ZERO_WEIGHT;
  }
  return N2V_OK;
}

extern "C" int n2v_edge_alias_build(const n2v_graph_t* graph, const int32_t* prev, const int32_t* cur,
                                    int64_t n_pairs, double return_param, double inout_param, int sum_mode,
                                    const int64_t* out_offset, int32_t* alias_out, double* probs_out,
                                    int32_t* scratch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(graph != nullptr && graph->n_parts == 1, "n2v_edge_alias_build: needs a single-part graph");
  N2V_CHECK_ARG(return_param != 0.0 && inout_param != 0.0, "Zero return (%g) or inout (%g) parameter!",
                return_param, inout_param);
  N2V_CHECK_ARG(sum_mode == N2V_SUM_NAIVE || sum_mode == N2V_SUM_NEUMAIER, "n2v_edge_alias_build: bad sum_mode %d", sum_mode);
  if (n_pairs <= 0) return N2V_OK;
  N2V_CHECK_ARG(prev && cur && out_offset && alias_out && probs_out && scratch, "n2v_edge_alias_build: NULL buffer");
  const n2v_graph_part_t& P = graph->parts[0];
  edge_alias_kernel<<<grid_for(n_pairs), kBlock, 0, stream>>>(P.vtx, P.col, P.weight, prev, cur, n_pairs,
                                                              return_param, inout_param, sum_mode, out_offset,
                                                              alias_out, probs_out, scratch);
  N2V_LAUNCH_OK();
  return N2V_OK;
}

extern "C" int n2v_alias_draw(const int32_t* alias, const double* probs, const int64_t* offset,
                              const double* first, const double* second, int64_t n_draws, int32_t* picked,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_draws <= 0) return N2V_OK;
  N2V_CHECK_ARG(alias && probs && offset && first && picked, "n2v_alias_draw: NULL buffer");
  alias_draw_kernel<<<grid_for(n_draws), kBlock, 0, stream>>>(alias, probs, offset, first, second, n_draws, picked);
  N2V_LAUNCH_OK();
  return N2V_OK;
}
