// K4: vocabulary statistics, sub-sampling thresholds and the negative-sampling table.
// Replaces gensim 3.8 Word2Vec.build_vocab (scan_vocab / prepare_vocab / make_cum_table) as
// reached from Node2VecGensim.fit (reference embedding.py:126).  Tokens are vertex ids.
#include <math.h>

#include "alias_core.cuh"
#include "n2v_internal.cuh"

namespace {

constexpr int kBlock = 256;

inline int grid_for(int64_t n) {
  const int64_t need = (n + kBlock - 1) / kBlock;
  const int64_t cap = int64_t(n2v::sm_count()) * 8;
  return static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
}

__global__ void count_tokens(const int32_t* __restrict__ walks, int64_t n_walks, int32_t len, int64_t pitch,
                             int64_t n_vertices, int64_t pos_offset, unsigned long long* __restrict__ counts,
                             long long* __restrict__ first_pos, unsigned int* __restrict__ bad) {
  const int64_t total = n_walks * len;
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < total; i += int64_t(gridDim.x) * kBlock) {
    const int64_t row = i / len;
    const int32_t c = static_cast<int32_t>(i - row * len);
    const int32_t tok = walks[row * pitch + c];
    if (tok < 0) continue;
    if (tok >= n_vertices) { atomicOr(bad, 1u); continue; }
    atomicAdd(counts + tok, 1ull);
    atomicMin(first_pos + tok, static_cast<long long>(pos_offset + i));
  }
}

__global__ void retain_totals(const int64_t* __restrict__ counts, int64_t n, int64_t min_count,
                              unsigned long long* __restrict__ out) {
  unsigned long long tot = 0, ids = 0;
  for (int64_t v = blockIdx.x * int64_t(kBlock) + threadIdx.x; v < n; v += int64_t(gridDim.x) * kBlock) {
    const int64_t c = counts[v];
    if (c > 0 && c >= min_count) { tot += static_cast<unsigned long long>(c); ++ids; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
    ids += __shfl_xor_sync(0xffffffffu, ids, o);
  }
  if ((threadIdx.x & 31) == 0 && ids) { atomicAdd(out, tot); atomicAdd(out + 1, ids); }
}

// gensim prepare_vocab: keep probability (sqrt(c/t) + 1) * (t/c), t = sample * retained total
__global__ void thresholds(const int64_t* __restrict__ counts, int64_t n, int64_t min_count, double sample,
                           double ns_exponent, const unsigned long long* __restrict__ totals,
                           uint32_t* __restrict__ keep_thr, double* __restrict__ weight) {
  const double retain_total = static_cast<double>(totals[0]);
  double threshold;
  if (sample == 0.0) threshold = retain_total;
  else if (sample < 1.0) threshold = sample * retain_total;
  else threshold = floor(sample * (3.0 + sqrt(5.0)) / 2.0);
  for (int64_t v = blockIdx.x * int64_t(kBlock) + threadIdx.x; v < n; v += int64_t(gridDim.x) * kBlock) {
    const int64_t c = counts[v];
    if (c <= 0 || c < min_count) { keep_thr[v] = 0u; weight[v] = 0.0; continue; }
    const double cd = static_cast<double>(c);
    const double p = (sqrt(cd / threshold) + 1.0) * (threshold / cd);
    uint32_t thr = 0xFFFFFFFFu;
    if (p < 1.0) {
      const double s = rint(p * 4294967296.0);   // gensim: int(round(p * 2**32))
      thr = s >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(s);
    }
    keep_thr[v] = thr;
    weight[v] = pow(cd, ns_exponent);
  }
}

// ---- negative-sampling alias table over all ids ----------------------------------------------
// Not a reference-parity object (gensim bisects a cumulative table; any exact alias table of
// count^0.75 gives the same law), so it is built for speed: ids are split into "small"
// (scaled prob < 1) and "large" lists, their probabilities are packed next to them,
// and ONE thread runs Vose's two-queue merge over the two packed streams.  The merge is
// sequential but its loads are contiguous and independent of the arithmetic (a few ns per id,
// ~1 s for 67 M ids) instead of a dependent global-memory round trip per id.
// Deterministic: ONE CTA walks the ids in order and compacts them with ballot + block scan, so both
// lists come out in ascending id order and the table (hence every negative drawn from a given
// random stream) is reproducible run to run.  The level it runs over has at most 262,144 entries
// (ids beyond 65,536 go through the two-level table), i.e. <= 256 iterations.
constexpr int kSplitBlock = 1024;
__global__ void __launch_bounds__(kSplitBlock)
scale_and_split(const double* __restrict__ weight, int64_t n, const double* __restrict__ total,
                int32_t* __restrict__ ids, double* __restrict__ packed,
                unsigned long long* __restrict__ counters /* [0] smalls, [1] larges */) {
  __shared__ unsigned int w_small[32], w_large[32];
  __shared__ unsigned int tot_small, tot_large;
  if (blockIdx.x != 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double scale = static_cast<double>(n) / *total;
  unsigned long long base_small = 0, base_large = 0;   // same value in every thread
  for (int64_t v0 = 0; v0 < n; v0 += kSplitBlock) {
    const int64_t v = v0 + tid;
    const bool valid = v < n;
    const double p = valid ? weight[v] * scale : 0.0;
    const bool small = valid && p < 1.0;
    const bool large = valid && !small;
    const unsigned int m_small = __ballot_sync(0xffffffffu, small);
    const unsigned int m_large = __ballot_sync(0xffffffffu, large);
    if (lane == 0) { w_small[warp] = __popc(m_small); w_large[warp] = __popc(m_large); }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the 32 per-warp counts
      const unsigned int cs = w_small[lane], cl = w_large[lane];
      unsigned int is = cs, il = cl;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int ts = __shfl_up_sync(0xffffffffu, is, o), tl = __shfl_up_sync(0xffffffffu, il, o);
        if (lane >= o) { is += ts; il += tl; }
      }
      w_small[lane] = is - cs;
      w_large[lane] = il - cl;
      if (lane == 31) { tot_small = is; tot_large = il; }
    }
    __syncthreads();
    if (valid) {
      const unsigned int below = (1u << lane) - 1u;
      // smalls grow from the front, larges from the back (vose_merge reads them as ids[n-1-j])
      const int64_t slot = small ? static_cast<int64_t>(base_small + w_small[warp] + __popc(m_small & below))
                                 : n - 1 - static_cast<int64_t>(base_large + w_large[warp] + __popc(m_large & below));
      ids[slot] = static_cast<int32_t>(v);
      packed[slot] = p;
    }
    base_small += tot_small;
    base_large += tot_large;
    __syncthreads();   // the scan arrays are rewritten by the next iteration
  }
  if (tid == 0) { counters[0] = base_small; counters[1] = base_large; }
}

__device__ __forceinline__ uint32_t neg_thr(double p) {
  if (p >= 1.0) return 0xFFFFFFFFu;
  const double s = ceil(p * 4294967296.0);
  return s >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(s);
}

// One warp: lanes stream the two packed lists through shared-memory windows (coalesced 1 KiB
// refills), lane 0 runs the inherently sequential two-queue merge out of shared memory.
constexpr int kWin = 1024;   // entries per window
__global__ void __launch_bounds__(32) vose_merge(const int32_t* __restrict__ ids, const double* __restrict__ packed,
                                                 int64_t n, const unsigned long long* __restrict__ counters,
                                                 int32_t* __restrict__ table) {
  __shared__ int32_t s_id[kWin], l_id[kWin];
  __shared__ double s_p[kWin], l_p[kWin];
  __shared__ int64_t sh_i, sh_j;
  __shared__ int sh_carry, sh_carry_id, sh_done;
  __shared__ double sh_carry_p;
  if (blockIdx.x != 0) return;
  const int lane = threadIdx.x;
  int2* out = reinterpret_cast<int2*>(table);
  const int64_t ns = static_cast<int64_t>(counters[0]), nl = static_cast<int64_t>(counters[1]);
  if (lane == 0) { sh_i = 0; sh_j = 0; sh_carry = 0; sh_carry_id = 0; sh_carry_p = 0.0; sh_done = 0; }
  __syncwarp();
  int64_t s_base = 0, l_base = 0;       // list positions of the windows' first entries
  int64_t s_have = 0, l_have = 0;       // entries valid in each window
  for (;;) {
    // refill whichever window lane 0 exhausted (both on the first pass)
    const int64_t i = sh_i, j = sh_j;
    if (i >= s_base + s_have && i < ns) {
      s_base = i;
      s_have = (ns - i) < kWin ? (ns - i) : kWin;
      for (int k = lane; k < s_have; k += 32) { s_id[k] = ids[i + k]; s_p[k] = packed[i + k]; }
    }
    if (j >= l_base + l_have && j < nl) {
      l_base = j;
      l_have = (nl - j) < kWin ? (nl - j) : kWin;
      for (int k = lane; k < l_have; k += 32) { l_id[k] = ids[n - 1 - (j + k)]; l_p[k] = packed[n - 1 - (j + k)]; }
    }
    __syncwarp();
    if (lane == 0) {
      int64_t ii = sh_i, jj = sh_j;
      bool carry = sh_carry != 0;
      int32_t carry_id = sh_carry_id;
      double carry_p = sh_carry_p;
      // big_p lives in the window (updated in place) so a refill never loses it
      while (jj < nl && jj < l_base + l_have && (carry || (ii < ns && ii < s_base + s_have))) {
        int32_t sid;
        double sp;
        if (carry) { sid = carry_id; sp = carry_p; carry = false; }
        else { sid = s_id[ii - s_base]; sp = s_p[ii - s_base]; ++ii; }
        const int lk = static_cast<int>(jj - l_base);
        const int32_t big_id = l_id[lk];
        out[sid] = make_int2(static_cast<int32_t>(neg_thr(sp)), big_id);
        const double bp = (l_p[lk] + sp) - 1.0;
        l_p[lk] = bp;
        if (bp < 1.0) { carry = true; carry_id = big_id; carry_p = bp; ++jj; }
      }
      sh_i = ii; sh_j = jj; sh_carry = carry ? 1 : 0; sh_carry_id = carry_id; sh_carry_p = carry_p;
      // finished when the larges are exhausted, or no small is left to place
      sh_done = (jj >= nl || (!carry && ii >= ns)) ? 1 : 0;
    }
    __syncwarp();
    if (sh_done) break;
  }
  // leftovers (fp residue): they keep themselves with probability 1
  const int64_t ii = sh_i, jj = sh_j;
  if (lane == 0 && sh_carry) out[sh_carry_id] = make_int2(static_cast<int32_t>(0xFFFFFFFFu), sh_carry_id);
  for (int64_t k = ii + lane; k < ns; k += 32) out[ids[k]] = make_int2(static_cast<int32_t>(0xFFFFFFFFu), ids[k]);
  for (int64_t k = jj + lane; k < nl; k += 32) {
    const int32_t id = ids[n - 1 - k];
    out[id] = make_int2(static_cast<int32_t>(0xFFFFFFFFu), id);
  }
}

// ---- two-level table for large id spaces: one thread per N2V_NEG_CHUNK-id chunk --------------------------
// chunk alias table (global ids as aliases) + the chunk's mass; same Vose core as the per-vertex
// graph tables (alias_core.cuh), probs / work list in the caller's scratch
__global__ void chunk_tables(double* __restrict__ probs, int64_t n, int32_t* __restrict__ work,
                             int32_t* __restrict__ table, double* __restrict__ chunk_mass) {
  const int64_t n_chunks = (n + N2V_NEG_CHUNK - 1) / N2V_NEG_CHUNK;
  int2* out = reinterpret_cast<int2*>(table);
  for (int64_t c = blockIdx.x * int64_t(kBlock) + threadIdx.x; c < n_chunks; c += int64_t(gridDim.x) * kBlock) {
    const int64_t lo = c * N2V_NEG_CHUNK;
    const uint32_t m = static_cast<uint32_t>((n - lo) < N2V_NEG_CHUNK ? (n - lo) : N2V_NEG_CHUNK);
    double* pr = probs + lo;
    double mass = 0.0;
    for (uint32_t i = 0; i < m; ++i) mass += pr[i];
    chunk_mass[c] = mass;
    if (!(mass > 0.0)) {   // a chunk of dropped ids: never selected by the top level
      for (uint32_t i = 0; i < m; ++i) out[lo + i] = make_int2(static_cast<int32_t>(0xFFFFFFFFu), static_cast<int32_t>(lo + i));
      continue;
    }
    n2v::build_alias_one(pr, m, N2V_SUM_NAIVE, work + lo,
                         [&](uint32_t i, int32_t a) { out[lo + i].y = static_cast<int32_t>(lo) + a; });
    for (uint32_t i = 0; i < m; ++i) {
      const double p = pr[i];
      if (p >= 1.0) out[lo + i].y = static_cast<int32_t>(lo + i);
      out[lo + i].x = static_cast<int32_t>(neg_thr(p));
    }
  }
}

__global__ void sum_weights(const double* __restrict__ w, int64_t n, double* __restrict__ partial) {
  // deterministic two-stage sum: one partial per block (fixed grid), added in block order by the caller kernel
  __shared__ double sh[kBlock];
  double acc = 0.0;
  for (int64_t v = blockIdx.x * int64_t(kBlock) + threadIdx.x; v < n; v += int64_t(gridDim.x) * kBlock) acc += w[v];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kBlock / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void finish_sum(double* __restrict__ partial, int n_partial) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < n_partial; ++k) t += partial[k];
    partial[0] = t;
  }
}

__global__ void init_rows(float* __restrict__ syn0, int64_t n_vertices, int32_t dim, uint32_t k0, uint32_t k1) {
  const int64_t quads = (static_cast<int64_t>(dim) + 3) / 4;
  const int64_t total = n_vertices * quads;
  const float inv = 1.0f / static_cast<float>(dim);
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < total; i += int64_t(gridDim.x) * kBlock) {
    const int64_t v = i / quads;
    const int32_t qd = static_cast<int32_t>(i - v * quads);
    const uint4 r = n2v::philox4x32_10(k0, k1, static_cast<uint32_t>(v), static_cast<uint32_t>(v >> 32),
                                       static_cast<uint32_t>(qd), 0x494E4954u /* "INIT" */);
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
    for (int k = 0; k < 4; ++k) {
      const int32_t d = qd * 4 + k;
      if (d < dim) syn0[v * dim + d] = (static_cast<float>(u[k] >> 8) * (1.0f / 16777216.0f) - 0.5f) * inv;
    }
  }
}

__global__ void scale_kernel(float* __restrict__ x, int64_t n, float f) {
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock) x[i] *= f;
}

}  // namespace

extern "C" int n2v_vocab_count(const int32_t* walks, int64_t n_walks, int32_t len, int64_t pitch,
                               int64_t n_vertices, int64_t pos_offset, int64_t* counts, int64_t* first_pos,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(n_walks >= 0 && len >= 1 && pitch >= len && n_vertices >= 0, "n2v_vocab_count: bad shape");
  if (n_walks == 0) return N2V_OK;
  N2V_CHECK_ARG(walks && counts && first_pos, "n2v_vocab_count: NULL buffer");
  unsigned int* bad = nullptr;
  N2V_CUDA(n2v::scratch_alloc(reinterpret_cast<void**>(&bad), sizeof(unsigned int), stream));
  N2V_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned int), stream));
  count_tokens<<<grid_for(n_walks * len), kBlock, 0, stream>>>(walks, n_walks, len, pitch, n_vertices, pos_offset,
                                                              reinterpret_cast<unsigned long long*>(counts),
                                                              reinterpret_cast<long long*>(first_pos), bad);
  N2V_LAUNCH_OK();
  unsigned int h = 0;
  N2V_CUDA(cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, stream));
  N2V_CUDA(cudaFreeAsync(bad, stream));
  N2V_CUDA(cudaStreamSynchronize(stream));
  N2V_CHECK_ARG(h == 0, "n2v_vocab_count: token id outside [0, %lld)", static_cast<long long>(n_vertices));
  return N2V_OK;
}

extern "C" int n2v_sgns_prepare(const int64_t* counts, int64_t n_vertices, int64_t min_count, double sample,
                                double ns_exponent, uint32_t* keep_thr, int32_t* neg_table, void* scratch,
                                int64_t* totals_host, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(n_vertices > 0 && n_vertices < (int64_t(1) << 31), "n2v_sgns_prepare: n_vertices out of range");
  N2V_CHECK_ARG(counts && keep_thr && neg_table && scratch, "n2v_sgns_prepare: NULL buffer");
  N2V_CHECK_ARG(sample >= 0.0, "n2v_sgns_prepare: negative sample");
  double* probs = static_cast<double*>(scratch);   // count^ns_exponent per id
  unsigned long long* d_tot = nullptr;
  N2V_CUDA(n2v::scratch_alloc(reinterpret_cast<void**>(&d_tot), 4 * sizeof(unsigned long long), stream));
  N2V_CUDA(cudaMemsetAsync(d_tot, 0, 4 * sizeof(unsigned long long), stream));
  retain_totals<<<grid_for(n_vertices), kBlock, 0, stream>>>(counts, n_vertices, min_count, d_tot);
  N2V_LAUNCH_OK();
  thresholds<<<grid_for(n_vertices), kBlock, 0, stream>>>(counts, n_vertices, min_count, sample, ns_exponent, d_tot,
                                                         keep_thr, probs);
  N2V_LAUNCH_OK();
  unsigned long long h[4] = {0, 0, 0, 0};
  N2V_CUDA(cudaMemcpyAsync(h, d_tot, sizeof(h), cudaMemcpyDeviceToHost, stream));
  N2V_CUDA(cudaStreamSynchronize(stream));
  if (totals_host) { totals_host[0] = static_cast<int64_t>(h[0]); totals_host[1] = static_cast<int64_t>(h[1]); }
  if (h[1] == 0) {
    N2V_CUDA(cudaFreeAsync(d_tot, stream));
    n2v::set_error("n2v_sgns_prepare: no token reaches min_count=%lld (empty vocabulary)", static_cast<long long>(min_count));
    return N2V_ERR_INVALID;
  }
  // scratch layout (T = n + top entries): [w f64 T | packed f64 T | ids i32 T]
  const int64_t n_top = n2v_neg_top_entries(n_vertices);
  const int64_t T = n_vertices + n_top;
  double* packed = probs + T;
  int32_t* ids = reinterpret_cast<int32_t*>(probs + 2 * T);
  const double* level_w = probs;          // weights of the level the one-thread merge runs over
  int64_t level_n = n_vertices;
  int32_t* level_table = neg_table;
  if (n_top > 0) {
    // per-chunk tables in parallel; their masses become the top level's weights
    double* mass = probs + n_vertices;
    chunk_tables<<<grid_for(n_top), kBlock, 0, stream>>>(probs, n_vertices, ids, neg_table, mass);
    N2V_LAUNCH_OK();
    level_w = mass;
    level_n = n_top;
    level_table = neg_table + 2 * n_vertices;
    packed = packed + n_vertices;          // disjoint from the region chunk_tables used
    ids = ids + n_vertices;
  }
  // one table over `level_n` entries: scaled probs -> small / large streams -> one-warp Vose merge
  double* partial = packed;                                                 // reused before scale_and_split fills it
  const int sum_grid = grid_for(level_n);
  sum_weights<<<sum_grid, kBlock, 0, stream>>>(level_w, level_n, partial);
  N2V_LAUNCH_OK();
  finish_sum<<<1, 32, 0, stream>>>(partial, sum_grid);
  N2V_LAUNCH_OK();
  double* d_total = reinterpret_cast<double*>(d_tot + 2);
  N2V_CUDA(cudaMemcpyAsync(d_total, partial, sizeof(double), cudaMemcpyDeviceToDevice, stream));
  N2V_CUDA(cudaMemsetAsync(d_tot, 0, 2 * sizeof(unsigned long long), stream));   // reuse as the two list counters
  scale_and_split<<<1, kSplitBlock, 0, stream>>>(level_w, level_n, d_total, ids, packed, d_tot);
  N2V_LAUNCH_OK();
  vose_merge<<<1, 32, 0, stream>>>(ids, packed, level_n, d_tot, level_table);
  N2V_LAUNCH_OK();
  N2V_CUDA(cudaFreeAsync(d_tot, stream));
  return N2V_OK;
}

extern "C" int n2v_sgns_init(float* syn0, int64_t n_vertices, int32_t dim, uint64_t seed, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(dim >= 1 && n_vertices >= 0, "n2v_sgns_init: bad shape");
  if (n_vertices == 0) return N2V_OK;
  N2V_CHECK_ARG(syn0 != nullptr, "n2v_sgns_init: NULL buffer");
  init_rows<<<grid_for(n_vertices * ((dim + 3) / 4)), kBlock, 0, stream>>>(
      syn0, n_vertices, dim, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  N2V_LAUNCH_OK();
  return N2V_OK;
}

extern "C" int n2v_sgns_exp_table(float* out) {
  N2V_CHECK_ARG(out != nullptr, "n2v_sgns_exp_table: NULL buffer");
  for (int i = 0; i < 1000; ++i) {
    const float e = static_cast<float>(exp((i / 1000.0 * 2 - 1) * 6.0));
    out[i] = e / (e + 1.0f);
  }
  return N2V_OK;
}

extern "C" int n2v_scale(float* x, int64_t n, float factor, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n <= 0) return N2V_OK;
  N2V_CHECK_ARG(x != nullptr, "n2v_scale: NULL buffer");
  scale_kernel<<<grid_for(n), kBlock, 0, stream>>>(x, n, factor);
  N2V_LAUNCH_OK();
  return N2V_OK;
}
