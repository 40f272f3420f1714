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
// vertices with kHubMin <= deg <= kHubMax are built by one CTA each with probs / alias / work
// lists in shared memory (16 B per arc, <= 192 KB): the sequential part then runs at
// shared-memory latency instead of L2 latency
constexpr uint32_t kHubMin = 256;
constexpr uint32_t kHubMax = 12288;
constexpr int kHubBlock = 256;

__device__ __forceinline__ uint32_t prob_to_thr(double pr) {
  // u32 < thr  <=>  u32 / 2^32 < pr   (exact for pr < 1); pr >= 1 saturates (alias_dst == dst there)
  if (pr >= 1.0) return 0xFFFFFFFFu;
  const double s = ceil(pr * 4294967296.0);
  return s >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(s);
}

// the sequential per-vertex construction on ONE thread (global-memory work lists)
__device__ void build_vertex_sequential(n2v_vertex_t* __restrict__ vtx, const n2v_vertex_t* lookup,
                                        const int32_t* __restrict__ col, const double* __restrict__ weight,
                                        int64_t v, int sum_mode, int32_t* __restrict__ alias,
                                        double* __restrict__ probs, n2v_arc_t* __restrict__ arcs,
                                        int32_t* __restrict__ scratch, unsigned long long* __restrict__ n_zero) {
  const uint64_t base = vtx[v].base;
  const uint32_t n = vtx[v].deg;
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
    return;
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

// Does any arc carry a weight other than exactly 1.0?  One streaming pass; the answer stays on the device.
__global__ void unit_scan_kernel(const double* __restrict__ weight, int64_t n_arcs, unsigned int* __restrict__ not_unit) {
  int other = 0;
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n_arcs; i += int64_t(gridDim.x) * kBlock)
    other |= (weight[i] != 1.0) ? 1 : 0;
  if (__syncthreads_or(other) && threadIdx.x == 0) atomicOr(not_unit, 1u);
}

// Unit-weight graphs (every BASELINE config): generate_alias_tables on n ones is sum = n exactly,
// mean = 1.0, probs = 1.0 / 1.0 = 1.0, nothing on the underfull list, every alias 0 -- in both sum modes.
// The records then depend on the arc alone: one thread per arc, coalesced, no per-vertex sequential pass.
__global__ void alias_unit_kernel(n2v_vertex_t* __restrict__ vtx, const n2v_vertex_t* lookup,
                                  const int32_t* __restrict__ col, int64_t n_vertices, int64_t n_arcs,
                                  int32_t* __restrict__ alias, double* __restrict__ probs,
                                  n2v_arc_t* __restrict__ arcs, const unsigned int* __restrict__ not_unit) {
  if (*not_unit != 0) return;
  const int64_t t0 = blockIdx.x * int64_t(kBlock) + threadIdx.x, stride = int64_t(gridDim.x) * kBlock;
  for (int64_t i = t0; i < n_arcs; i += stride) {
    const int32_t self = col[i];
    const uint32_t b = lookup[self].base, d = lookup[self].deg;
    arcs[i] = n2v_arc_t{0xFFFFFFFFu, self, self, 0, b, d, b, d};
    probs[i] = 1.0;
    if (alias) alias[i] = 0;
  }
  for (int64_t v = t0; v < n_vertices; v += stride) {
    const uint32_t n = vtx[v].deg;
    if (n) vtx[v].wsum = static_cast<float>(static_cast<double>(n));   // the left-to-right fp64 sum of n ones
  }
}

__global__ void alias_build_kernel(n2v_vertex_t* __restrict__ vtx, const n2v_vertex_t* lookup,
                                   const int32_t* __restrict__ col,
                                   const double* __restrict__ weight, int64_t n_vertices, int sum_mode,
                                   int32_t* __restrict__ alias, double* __restrict__ probs,
                                   n2v_arc_t* __restrict__ arcs, int32_t* __restrict__ scratch,
                                   unsigned long long* __restrict__ n_zero, int32_t* __restrict__ hubs,
                                   unsigned int* __restrict__ n_hubs, int32_t* __restrict__ giants,
                                   unsigned int* __restrict__ n_giants, const unsigned int* __restrict__ not_unit) {
  if (*not_unit == 0) return;            // every weight is 1.0: alias_unit_kernel wrote the records
  for (int64_t v = blockIdx.x * int64_t(kBlock) + threadIdx.x; v < n_vertices;
       v += int64_t(gridDim.x) * kBlock) {
    const uint32_t n = vtx[v].deg;
    if (n == 0) continue;
    if (n >= kHubMin && n <= kHubMax) {  // staged in shared memory by alias_hub_kernel
      const unsigned int slot = atomicAdd(n_hubs, 1u);
      hubs[slot] = static_cast<int32_t>(v);
      continue;
    }
    if (n > kHubMax) {                   // one CTA each in alias_giant_kernel
      const unsigned int slot = atomicAdd(n_giants, 1u);
      giants[slot] = static_cast<int32_t>(v);
      continue;
    }
    build_vertex_sequential(vtx, lookup, col, weight, v, sum_mode, alias, probs, arcs, scratch, n_zero);
  }
}

// One CTA per vertex above the shared-memory limit (deg > kHubMax; the reference's default trim
// cap is 100000, constants.py:6).  If every weight is exactly 1.0 the reference's construction is
// trivial -- sum = n exactly, probs = 1.0, nothing on the underfull list, every alias stays 0 --
// and the records are written in parallel.  Otherwise thread 0 runs the sequential construction.
__global__ void __launch_bounds__(kHubBlock)
alias_giant_kernel(n2v_vertex_t* __restrict__ vtx, const n2v_vertex_t* lookup, const int32_t* __restrict__ col,
                   const double* __restrict__ weight, int sum_mode, int32_t* __restrict__ alias,
                   double* __restrict__ probs, n2v_arc_t* __restrict__ arcs, int32_t* __restrict__ scratch,
                   const int32_t* __restrict__ giants, const unsigned int* __restrict__ n_giants_ptr,
                   unsigned long long* __restrict__ n_zero) {
  const unsigned int n_giants = *n_giants_ptr;       // 0 on all-unit graphs (alias_build_kernel listed nothing)
  for (unsigned int h = blockIdx.x; h < n_giants; h += gridDim.x) {
    const int64_t v = giants[h];
    const uint32_t base = vtx[v].base, n = vtx[v].deg;
    int not_unit = 0;
    for (uint32_t i = threadIdx.x; i < n; i += kHubBlock) not_unit |= (weight[base + i] != 1.0) ? 1 : 0;
    const bool unit = __syncthreads_or(not_unit) == 0;
    if (unit) {
      if (threadIdx.x == 0) vtx[v].wsum = static_cast<float>(static_cast<double>(n));
      for (uint32_t i = threadIdx.x; i < n; i += kHubBlock) {
        const int32_t self = col[base + i];
        const uint32_t b = lookup[self].base, d = lookup[self].deg;
        arcs[base + i] = n2v_arc_t{0xFFFFFFFFu, self, self, 0, b, d, b, d};
        probs[base + i] = 1.0;
        if (alias) alias[base + i] = 0;
      }
    } else if (threadIdx.x == 0) {
      build_vertex_sequential(vtx, lookup, col, weight, v, sum_mode, alias, probs, arcs, scratch, n_zero);
    }
    __syncthreads();
  }
}

// One CTA per hub vertex.  Bit-exact with the sequential reference: the fp64 sum and the LIFO
// pairing run on thread 0 in index order; only order-free work (load, divide, pack) and the
// order-PRESERVING list construction (chunked prefix sum) are parallel.
__global__ void __launch_bounds__(kHubBlock)
alias_hub_kernel(n2v_vertex_t* __restrict__ vtx, const n2v_vertex_t* lookup, const int32_t* __restrict__ col,
                 const double* __restrict__ weight, int sum_mode, int32_t* __restrict__ alias,
                 double* __restrict__ probs, n2v_arc_t* __restrict__ arcs, const int32_t* __restrict__ hubs,
                 const unsigned int* __restrict__ n_hubs_ptr, unsigned long long* __restrict__ n_zero) {
  const unsigned int n_hubs = *n_hubs_ptr;   // written by alias_build_kernel earlier on the same stream
  extern __shared__ __align__(16) unsigned char hub_smem[];
  __shared__ double s_mean;
  __shared__ int s_counts[kHubBlock + 1];
  __shared__ int s_ok;
  for (unsigned int h = blockIdx.x; h < n_hubs; h += gridDim.x) {
    const int64_t v = hubs[h];
    const uint32_t base = vtx[v].base, n = vtx[v].deg;
    double* pr = reinterpret_cast<double*>(hub_smem);
    int32_t* al = reinterpret_cast<int32_t*>(pr + n);
    int32_t* stack = al + n;
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < n; i += kHubBlock) {
      pr[i] = weight[base + i];
      al[i] = 0;
    }
    __syncthreads();
    if (tid == 0) {
      double wsum = 0.0;
      for (uint32_t i = 0; i < n; ++i) wsum = __dadd_rn(wsum, pr[i]);
      vtx[v].wsum = static_cast<float>(wsum);
      const double total = sum_mode == N2V_SUM_NAIVE ? wsum
                                                     : n2v::python_sum([&](uint32_t i) { return pr[i]; }, n, sum_mode);
      s_mean = __ddiv_rn(total, static_cast<double>(n));
      s_ok = (s_mean != 0.0) ? 1 : 0;
    }
    __syncthreads();
    const bool ok = s_ok != 0;
    if (ok) {
      const double mean = s_mean;
      // probs = w / mean, then index-ordered small / large lists via a chunked prefix sum
      const uint32_t chunk = (n + kHubBlock - 1) / kHubBlock;
      const uint32_t lo = min(n, tid * chunk), hi = min(n, lo + chunk);
      int n_small = 0;
      for (uint32_t i = lo; i < hi; ++i) {
        const double q = __ddiv_rn(pr[i], mean);
        pr[i] = q;
        n_small += (q < 1.0) ? 1 : 0;
      }
      s_counts[tid] = n_small;
      __syncthreads();
      if (tid == 0) {
        int run = 0;
        for (int t = 0; t < kHubBlock; ++t) { const int c = s_counts[t]; s_counts[t] = run; run += c; }
        s_counts[kHubBlock] = run;
      }
      __syncthreads();
      {
        int ps = s_counts[tid];                          // smalls before this chunk
        int pl = static_cast<int>(lo) - ps;              // larges before this chunk
        for (uint32_t i = lo; i < hi; ++i) {
          if (pr[i] < 1.0) stack[ps++] = static_cast<int32_t>(i);
          else stack[n - 1 - (pl++)] = static_cast<int32_t>(i);
        }
      }
      __syncthreads();
      if (tid == 0) {
        int64_t ns = s_counts[kHubBlock], nl = static_cast<int64_t>(n) - ns;
        while (ns > 0 && nl > 0) {
          const int32_t a = stack[--ns];
          const int32_t b = stack[n - 1 - (--nl)];
          al[a] = b;
          const double ph = __dsub_rn(__dadd_rn(pr[b], pr[a]), 1.0);
          pr[b] = ph;
          if (ph < 1.0) stack[ns++] = b;
          else stack[n - 1 - (nl++)] = b;
        }
      }
      __syncthreads();
    } else if (tid == 0) {
      atomicAdd(n_zero, 1ull);
    }
    for (uint32_t i = tid; i < n; i += kHubBlock) {
      const int32_t self = col[base + i];
      const double p = ok ? pr[i] : 0.0;
      const int32_t a = ok ? al[i] : 0;
      n2v_arc_t rec;
      rec.thr = ok ? prob_to_thr(p) : 0xFFFFFFFFu;
      rec.dst = self;
      rec.alias_dst = (!ok || p >= 1.0) ? self : col[base + a];
      rec.alias_idx = a;
      rec.dst_base = lookup[rec.dst].base;
      rec.dst_deg = lookup[rec.dst].deg;
      rec.adst_base = lookup[rec.alias_dst].base;
      rec.adst_deg = lookup[rec.alias_dst].deg;
      arcs[base + i] = rec;
      probs[base + i] = p;
      if (alias) alias[base + i] = a;
    }
    __syncthreads();
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

// {fwd, rev} return-mass ratios per arc (general fold); one thread per arc
__global__ void ratio_kernel(const n2v_vertex_t* __restrict__ vtx, const int32_t* __restrict__ col,
                             const double* __restrict__ weight, int64_t n_vertices, float2* __restrict__ out) {
  for (int64_t v = blockIdx.x * int64_t(kBlock) + threadIdx.x; v < n_vertices; v += int64_t(gridDim.x) * kBlock) {
    const uint32_t base = vtx[v].base, n = vtx[v].deg;
    const float wv = vtx[v].wsum;
    uint32_t i = 0;
    while (i < n) {
      const int32_t x = col[base + i];
      uint32_t j = i;
      double tot = 0.0;
      while (j < n && col[base + j] == x) tot = __dadd_rn(tot, weight[base + j++]);   // parallel arcs v -> x
      const float fwd = __fdiv_rn(static_cast<float>(tot), wv);
      // reverse mass: arcs x -> v (lower bound of v in x's ascending slice, then the run)
      const uint32_t xb = vtx[x].base, xn = vtx[x].deg;
      uint32_t lo = 0, hi = xn;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (col[xb + mid] < static_cast<int32_t>(v)) lo = mid + 1; else hi = mid;
      }
      double back = 0.0;
      while (lo < xn && col[xb + lo] == static_cast<int32_t>(v)) back = __dadd_rn(back, weight[xb + lo++]);
      const float rev = back > 0.0 ? __fdiv_rn(static_cast<float>(back), vtx[x].wsum) : 0.0f;
      for (; i < j; ++i) out[base + i] = make_float2(fwd, rev);
    }
  }
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
  const int64_t cap = int64_t(n2v::sm_count()) * 16;
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
  // small device scratch: [0] zero-weight counter, [1] hub / giant counters (2 x u32), [2] "some weight is
  // not 1.0" flag, then the hub and giant lists
  const int64_t max_hubs = n_arcs / kHubMin + 1;
  const int64_t max_giants = n_arcs / (kHubMax + 1) + 1;
  unsigned long long* d_zero = nullptr;
  N2V_CUDA(n2v::scratch_alloc(reinterpret_cast<void**>(&d_zero), 24 + sizeof(int32_t) * (max_hubs + max_giants), stream));
  N2V_CUDA(cudaMemsetAsync(d_zero, 0, 24, stream));
  unsigned int* d_nhubs = reinterpret_cast<unsigned int*>(d_zero + 1);
  unsigned int* d_ngiants = d_nhubs + 1;
  unsigned int* d_not_unit = reinterpret_cast<unsigned int*>(d_zero + 2);
  int32_t* d_hubs = reinterpret_cast<int32_t*>(d_zero + 3);
  int32_t* d_giants = d_hubs + max_hubs;
  const n2v_vertex_t* lookup = vtx_lookup ? vtx_lookup : vtx;
  // all-unit graphs take the arc-parallel path; the per-vertex kernels below then return at once (the
  // decision stays on the device: no host round trip)
  unit_scan_kernel<<<grid_for(n_arcs), kBlock, 0, stream>>>(weight_sorted, n_arcs, d_not_unit);
  alias_unit_kernel<<<grid_for(n_arcs > n_vertices ? n_arcs : n_vertices), kBlock, 0, stream>>>(
      vtx, lookup, col, n_vertices, n_arcs, alias, probs, arcs, d_not_unit);
  N2V_LAUNCH_OK();
  alias_build_kernel<<<grid_for(n_vertices), kBlock, 0, stream>>>(vtx, lookup, col, weight_sorted, n_vertices,
                                                                   sum_mode, alias, probs, arcs, scratch, d_zero,
                                                                   d_hubs, d_nhubs, d_giants, d_ngiants, d_not_unit);
  N2V_LAUNCH_OK();
  if (n_arcs > kHubMax) {
    const int64_t grid = max_giants < 4 * n2v::sm_count() ? max_giants : 4 * n2v::sm_count();
    alias_giant_kernel<<<static_cast<unsigned int>(grid), kHubBlock, 0, stream>>>(
        vtx, lookup, col, weight_sorted, sum_mode, alias, probs, arcs, scratch, d_giants, d_ngiants, d_zero);
    N2V_LAUNCH_OK();
  }
  if (n_arcs >= kHubMin) {
    // the hub count stays on the device (no host round trip): one CTA per SM strides over the list and
    // exits at once when it is empty
    const size_t smem = static_cast<size_t>(kHubMax) * 16;
    N2V_CUDA(cudaFuncSetAttribute(alias_hub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int64_t grid = max_hubs < n2v::sm_count() ? max_hubs : n2v::sm_count();
    alias_hub_kernel<<<static_cast<unsigned int>(grid), kHubBlock, smem, stream>>>(
        vtx, lookup, col, weight_sorted, sum_mode, alias, probs, arcs, d_hubs, d_nhubs, d_zero);
    N2V_LAUNCH_OK();
  }
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

extern "C" int n2v_ratio_build(const n2v_graph_t* graph, float* ratio_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(graph != nullptr && graph->n_parts == 1, "n2v_ratio_build: needs a single-part graph");
  if (graph->n_arcs == 0) return N2V_OK;
  N2V_CHECK_ARG(ratio_out != nullptr, "n2v_ratio_build: NULL buffer");
  const n2v_graph_part_t& P = graph->parts[0];
  ratio_kernel<<<grid_for(graph->n_vertices), kBlock, 0, stream>>>(P.vtx, P.col, P.weight, graph->n_vertices,
                                                                  reinterpret_cast<float2*>(ratio_out));
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
