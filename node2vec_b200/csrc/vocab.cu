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
  const int64_t cap = int64_t(n2v::kSmCount) * 8;
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

// one alias table over all ids (sequential Vose, same core as the per-vertex tables)
__global__ void neg_table_kernel(double* __restrict__ probs, int64_t n, int32_t* __restrict__ work,
                                 int32_t* __restrict__ table, int* __restrict__ ok) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int2* out = reinterpret_cast<int2*>(table);
  const bool good = n2v::build_alias_one(probs, static_cast<uint32_t>(n), N2V_SUM_NAIVE, work,
                                         [&](uint32_t i, int32_t a) { out[i].y = a; });
  *ok = good ? 1 : 0;
  if (!good) return;
  for (int64_t i = 0; i < n; ++i) {
    const double p = probs[i];
    uint32_t thr = 0xFFFFFFFFu;
    if (p < 1.0) {
      const double s = ceil(p * 4294967296.0);
      thr = s >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(s);
    } else {
      out[i].y = static_cast<int32_t>(i);
    }
    out[i].x = static_cast<int32_t>(thr);
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
  N2V_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&bad), sizeof(unsigned int), stream));
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
  double* probs = static_cast<double*>(scratch);
  int32_t* work = reinterpret_cast<int32_t*>(probs + n_vertices);
  unsigned long long* d_tot = nullptr;
  N2V_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_tot), 4 * sizeof(unsigned long long), stream));
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
  int* ok = reinterpret_cast<int*>(d_tot + 2);
  neg_table_kernel<<<1, 32, 0, stream>>>(probs, n_vertices, work, neg_table, ok);
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
