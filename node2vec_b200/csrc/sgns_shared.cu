// K3s (EXPERIMENTAL, opt-in): skip-gram negative sampling with WINDOW-SHARED negatives.
//
// Same model, same update rule and same per-target arithmetic as sgns.cu (gensim 3.8
// w2v_fast_sentence_sg_neg as reached from the reference's embedding.py:120-127), but the K
// negatives are drawn ONCE PER CENTRE and shared by the ~2b pairs of that centre's window instead
// of being re-drawn per pair (gensim).  Marginally every pair still sees K negatives ~ count^0.75;
// inside one window they are the same K words (the negative-sharing idea of pWord2Vec, Ji et al.
// 2016).  A duplicate draw inside one K-set is dropped (its second register copy would diverge).
// This is a different sampling scheme, so it is NOT the parity path: it has its own AUC gate and
// is only reachable through Word2Vec(share_negatives=True) / n2v_sgns_train_shared.
//
// Why: in sgns.cu a pair costs 6.2 row reads + 6.2 row reductions, every negative read sits on the
// dependent chain alias entry -> row -> dot, and the kernel runs at the L2 sector rate.  Here the
// K + 1 target rows of a centre live in REGISTERS for the whole window (read once, updated pair by
// pair in registers -- the warp sees its own updates in gensim's sequential order -- and reduced
// to memory once per centre): a pair is 1 row read + 1 row reduction + 6 register dots, a centre is
// 6 row reads + 6 row reductions.  Row operations per pair: 12.4 -> ~4.2; no sampling and no
// dependent gather inside the pair loop.
//
// Kept in its own translation unit so that the production kernel's code is untouched by it.
// Supports dim <= 128 (one float4 per lane), negative == N2V_SHARED_K, vector-reduction updates.
#include "n2v_internal.cuh"

namespace {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
constexpr int KS = 5;          // negatives per centre held in registers
constexpr float kMaxExp = 6.0f;
constexpr int kExpTable = 1000;

struct SharedArgs {
  const int32_t* walks;
  const uint32_t* keep_thr;
  const int2* neg_table;
  const float* exp_table;
  float* syn0;
  float* syn1neg;
  unsigned long long* stats;
  int32_t* trace;
  float* trace_alpha;
  int64_t trace_cap;
  int64_t n_walks, pitch, walk_offset, total_walks;
  uint32_t n_vertices;
  uint32_t n_top;
  int32_t len, len_cap, dim, window, epochs, epoch, batch_words;
  float alpha, min_alpha;
  uint32_t key0, key1;
};

__device__ __forceinline__ uint32_t pcg_next(uint32_t& state) {   // PCG-RXS-M-XS-32, as sgns.cu
  const uint32_t s = state;
  state = s * 747796405u + 2891336453u;
  const uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
  return (w >> 22u) ^ w;
}

template <bool FULL>
__device__ __forceinline__ float4 load_row(const float* base, int32_t dim, int lane) {
  const int d = lane * 4;
  return (FULL || d < dim) ? *reinterpret_cast<const float4*>(base + d) : make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ float dot_row(const float4& a, const float4& b) {
  float p = a.x * b.x;
  p = fmaf(a.y, b.y, p);
  p = fmaf(a.z, b.z, p);
  p = fmaf(a.w, b.w, p);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  return p;
}

__device__ __forceinline__ void axpy(float4& acc, float g, const float4& x) {
  acc.x = fmaf(g, x.x, acc.x);
  acc.y = fmaf(g, x.y, acc.y);
  acc.z = fmaf(g, x.z, acc.z);
  acc.w = fmaf(g, x.w, acc.w);
}

template <bool FULL>
__device__ __forceinline__ void red_row(float* base, int32_t dim, int lane, const float4& delta) {
  const int d = lane * 4;
  if (FULL || d < dim) atomicAdd(reinterpret_cast<float4*>(base + d), delta);   // red.global.add.v4.f32
}

__device__ __forceinline__ float gradient(const float* table, float f, float label, float alpha, bool ok) {
  const float fc = fminf(fmaxf(f, -kMaxExp), kMaxExp - 1e-3f);
  const float s = table[static_cast<int>((fc + kMaxExp) * (kExpTable / kMaxExp / 2.0f))];
  return ok ? (label - s) * alpha : 0.0f;
}

template <bool TRACE, bool FULL>
__global__ void __launch_bounds__(kBlock, 2) sgns_shared_kernel(const __grid_constant__ SharedArgs A) {
  extern __shared__ int32_t smem[];
  __shared__ float exp_table[kExpTable];
  for (int i = threadIdx.x; i < kExpTable; i += kBlock) exp_table[i] = __ldg(A.exp_table + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  int32_t* sent = smem + wib * A.len_cap;
  if (TRACE && (blockIdx.x != 0 || wib != 0)) return;   // trace mode: ONE warp, walks in order
  const int64_t n_warps = TRACE ? 1 : static_cast<int64_t>(gridDim.x) * kWarps;
  unsigned long long c_pairs = 0, c_kept = 0;
  uint32_t c_negskip = 0, c_clip = 0;
  int64_t trace_pos = 0;
  const double total_words = static_cast<double>(A.total_walks) * static_cast<double>(A.len);

  for (int64_t s = TRACE ? 0 : static_cast<int64_t>(blockIdx.x) * kWarps + wib; s < A.n_walks; s += n_warps) {
    const int64_t gs = A.walk_offset + s;
    float alpha;   // learning rate of this walk's job: same schedule as sgns.cu
    {
      const int64_t job_start_words = (gs * A.len / A.batch_words) * static_cast<int64_t>(A.batch_words);
      const double progress = (static_cast<double>(A.epoch) + static_cast<double>(job_start_words) / total_words) /
                              static_cast<double>(A.epochs);
      const double a = static_cast<double>(A.alpha) - (static_cast<double>(A.alpha) - static_cast<double>(A.min_alpha)) * progress;
      alpha = static_cast<float>(a < static_cast<double>(A.min_alpha) ? static_cast<double>(A.min_alpha) : a);
    }
    // sub-sampling, identical to sgns.cu (same Philox counters => same kept tokens)
    int n = 0;
    for (int k0 = 0; k0 < A.len; k0 += 32) {
      const int k = k0 + lane;
      int32_t tok = -1;
      bool keep = false;
      if (k < A.len) {
        tok = __ldg(A.walks + s * A.pitch + k);
        if (tok >= 0 && static_cast<uint32_t>(tok) < A.n_vertices) {   // ids beyond the table are out-of-vocabulary (gensim skips them)
          const uint32_t thr = __ldg(A.keep_thr + tok);
          if (thr == 0xFFFFFFFFu) keep = true;
          else if (thr != 0u) {
            const uint4 r = n2v::philox4x32_10(A.key0, A.key1, static_cast<uint32_t>(gs), static_cast<uint32_t>(gs >> 32),
                                               static_cast<uint32_t>(A.epoch), 0x53554200u + static_cast<uint32_t>(k));
            keep = r.x < thr;
          }
        }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      if (keep) sent[n + __popc(m & ((1u << lane) - 1u))] = tok;
      n += __popc(m);
    }
    __syncwarp();
    c_kept += static_cast<unsigned long long>(n);
    uint32_t rnd;
    {
      const uint4 r = n2v::philox4x32_10(A.key0, A.key1, static_cast<uint32_t>(gs), static_cast<uint32_t>(gs >> 32),
                                         static_cast<uint32_t>(A.epoch), 0x50434700u);
      rnd = r.x;
    }
    for (int i = 0; i < n; ++i) {
      const uint32_t b = __umulhi(pcg_next(rnd), static_cast<uint32_t>(A.window));
      int j0 = i - A.window + static_cast<int>(b), j1 = i + A.window + 1 - static_cast<int>(b);
      if (j0 < 0) j0 = 0;
      if (j1 > n) j1 = n;
      if (j1 - j0 <= 1) continue;
      const int32_t wi = sent[i];
      float* pos_ptr = A.syn1neg + static_cast<int64_t>(wi) * A.dim;
      float4 pos = load_row<FULL>(pos_ptr, A.dim, lane);
      float4 pos_d = make_float4(0.f, 0.f, 0.f, 0.f);
      // the centre's K negatives: drawn once, rows + their accumulated updates in registers
      int32_t tgt[KS];
      bool live[KS];
      float4 neg[KS], neg_d[KS];
#pragma unroll
      for (int d = 0; d < KS; ++d) {
        uint32_t lo = 0, span = A.n_vertices;
        if (A.n_top) {
          const uint32_t c0 = __umulhi(pcg_next(rnd), A.n_top);
          const int2 te = __ldg(A.neg_table + A.n_vertices + c0);
          const uint32_t c = (pcg_next(rnd) < static_cast<uint32_t>(te.x)) ? c0 : static_cast<uint32_t>(te.y);
          lo = c * N2V_NEG_CHUNK;
          span = min(static_cast<uint32_t>(N2V_NEG_CHUNK), A.n_vertices - lo);
        }
        const uint32_t u1 = pcg_next(rnd);
        const uint32_t u2 = pcg_next(rnd);
        const uint32_t slot = lo + __umulhi(u1, span);
        const int2 e = __ldg(A.neg_table + slot);
        tgt[d] = (u2 < static_cast<uint32_t>(e.x)) ? static_cast<int32_t>(slot) : e.y;
        bool ok = tgt[d] != wi;                 // gensim: a negative equal to the centre is skipped
#pragma unroll
        for (int e2 = 0; e2 < KS; ++e2)
          if (e2 < d && live[e2] && tgt[e2] == tgt[d]) ok = false;   // duplicate inside the K-set: dropped
        live[d] = ok;
        if (TRACE) c_negskip += ok ? 0u : 1u;
        neg[d] = ok ? load_row<FULL>(A.syn1neg + static_cast<int64_t>(tgt[d]) * A.dim, A.dim, lane)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
        neg_d[d] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int j = j0; j < j1; ++j) {
        if (j == i) continue;
        const int32_t wj = sent[j];
        float* in_ptr = A.syn0 + static_cast<int64_t>(wj) * A.dim;
        const float4 in = load_row<FULL>(in_ptr, A.dim, lane);
        float4 work = make_float4(0.f, 0.f, 0.f, 0.f);
        if (TRACE && trace_pos < A.trace_cap && lane == 0) {
          int32_t* trow = A.trace + trace_pos * (2 + KS);
          trow[0] = wi;
          trow[1] = wj;
#pragma unroll
          for (int d = 0; d < KS; ++d) trow[2 + d] = live[d] ? tgt[d] : -1;
          A.trace_alpha[trace_pos] = alpha;
        }
        {   // positive target: the centre word
          const float f = dot_row(in, pos);
          const bool ok = f > -kMaxExp && f < kMaxExp;
          const float g = gradient(exp_table, f, 1.0f, alpha, ok);
          if (TRACE) c_clip += ok ? 0u : 1u;
          axpy(work, g, pos);
          axpy(pos, g, in);
          axpy(pos_d, g, in);
        }
#pragma unroll
        for (int d = 0; d < KS; ++d) {
          const float f = dot_row(in, neg[d]);
          const bool in_range = f > -kMaxExp && f < kMaxExp;
          const float g = gradient(exp_table, f, 0.0f, alpha, in_range && live[d]);
          if (TRACE) c_clip += (live[d] && !in_range) ? 1u : 0u;
          axpy(work, g, neg[d]);
          axpy(neg[d], g, in);
          axpy(neg_d[d], g, in);
        }
        red_row<FULL>(in_ptr, A.dim, lane, work);
        ++c_pairs;
        if (TRACE) ++trace_pos;
      }
      // one reduction per target row per centre
      red_row<FULL>(pos_ptr, A.dim, lane, pos_d);
#pragma unroll
      for (int d = 0; d < KS; ++d)
        if (live[d]) red_row<FULL>(A.syn1neg + static_cast<int64_t>(tgt[d]) * A.dim, A.dim, lane, neg_d[d]);
    }
    __syncwarp();
  }
  if (A.stats && lane == 0) {
    if (c_pairs) atomicAdd(A.stats + 0, c_pairs);
    if (c_kept) atomicAdd(A.stats + 1, c_kept);
    if (c_negskip) atomicAdd(A.stats + 2, static_cast<unsigned long long>(c_negskip));
    if (c_clip) atomicAdd(A.stats + 3, static_cast<unsigned long long>(c_clip));
  }
}

}  // namespace

extern "C" int n2v_sgns_train_shared(const int32_t* walks, int64_t n_walks, int32_t len, int64_t pitch,
                                     const uint32_t* keep_thr, const int32_t* neg_table, int64_t n_vertices,
                                     float* syn0, float* syn1neg, const float* exp_table,
                                     const n2v_sgns_params_t* P, uint64_t* stats, int32_t* trace, float* trace_alpha,
                                     int64_t trace_cap, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(P != nullptr, "n2v_sgns_train_shared: params is NULL");
  N2V_CHECK_ARG(P->dim >= 4 && P->dim <= 128 && P->dim % 4 == 0,
                "n2v_sgns_train_shared: dim %d must be a multiple of 4 in [4, 128]", P->dim);
  N2V_CHECK_ARG(P->negative == KS, "n2v_sgns_train_shared: negative must be %d (got %d)", KS, P->negative);
  N2V_CHECK_ARG(P->atomic_updates != 0, "n2v_sgns_train_shared: vector-reduction updates only");
  N2V_CHECK_ARG(P->window >= 1, "n2v_sgns_train_shared: window (%d) must be >= 1", P->window);
  N2V_CHECK_ARG(P->epochs >= 1 && P->epoch >= 0 && P->epoch < P->epochs && P->batch_words >= 1,
                "n2v_sgns_train_shared: bad epoch schedule");
  N2V_CHECK_ARG(n_walks >= 0 && len >= 1 && pitch >= len && len <= 10000,
                "n2v_sgns_train_shared: bad walk matrix shape (sentences are capped at 10000 tokens)");
  N2V_CHECK_ARG(n_vertices > 0 && n_vertices < (int64_t(1) << 31), "n2v_sgns_train_shared: n_vertices out of range");
  if (n_walks == 0) return N2V_OK;
  N2V_CHECK_ARG(walks && keep_thr && neg_table && syn0 && syn1neg && exp_table, "n2v_sgns_train_shared: NULL buffer");
  N2V_CHECK_ARG(trace == nullptr || trace_alpha != nullptr, "n2v_sgns_train_shared: trace needs trace_alpha");
  N2V_CHECK_ARG((reinterpret_cast<uintptr_t>(syn0) & 15) == 0 && (reinterpret_cast<uintptr_t>(syn1neg) & 15) == 0,
                "n2v_sgns_train_shared: tables must be 16-byte aligned");
  SharedArgs A{};
  A.walks = walks;
  A.keep_thr = keep_thr;
  A.neg_table = reinterpret_cast<const int2*>(neg_table);
  A.exp_table = exp_table;
  A.syn0 = syn0;
  A.syn1neg = syn1neg;
  A.stats = reinterpret_cast<unsigned long long*>(stats);
  A.trace = trace;
  A.trace_alpha = trace_alpha;
  A.trace_cap = trace_cap;
  A.n_walks = n_walks;
  A.pitch = pitch;
  A.walk_offset = P->walk_offset;
  A.total_walks = P->total_walks > 0 ? P->total_walks : n_walks;
  A.n_vertices = static_cast<uint32_t>(n_vertices);
  A.n_top = static_cast<uint32_t>(n2v_neg_top_entries(n_vertices));
  A.len = len;
  A.len_cap = (len + 31) & ~31;
  A.dim = P->dim;
  A.window = P->window;
  A.epochs = P->epochs;
  A.epoch = P->epoch;
  A.batch_words = P->batch_words;
  A.alpha = P->alpha;
  A.min_alpha = P->min_alpha;
  A.key0 = static_cast<uint32_t>(P->seed);
  A.key1 = static_cast<uint32_t>(P->seed >> 32);
  const size_t smem = static_cast<size_t>(A.len_cap) * kWarps * sizeof(int32_t);
  N2V_CHECK_ARG(smem <= 200 * 1024, "n2v_sgns_train_shared: walk too long for shared-memory staging");
  int grid;
  if (trace) grid = 1;
  else {
    const int64_t need = (n_walks + kWarps - 1) / kWarps;
    const int64_t cap = int64_t(n2v::sm_count()) * 8;
    grid = static_cast<int>(need < cap ? need : cap);
  }
  const bool full = P->dim == 128;
#define N2V_SHARED_GO(TR, FU)                                                                                      \
  do {                                                                                                             \
    if (smem > 48 * 1024)                                                                                          \
      cudaFuncSetAttribute(sgns_shared_kernel<TR, FU>, cudaFuncAttributeMaxDynamicSharedMemorySize,               \
                           static_cast<int>(smem));                                                                \
    sgns_shared_kernel<TR, FU><<<grid, kBlock, smem, stream>>>(A);                                                \
  } while (0)
  if (trace) {
    if (full) N2V_SHARED_GO(true, true); else N2V_SHARED_GO(true, false);
  } else {
    if (full) N2V_SHARED_GO(false, true); else N2V_SHARED_GO(false, false);
  }
#undef N2V_SHARED_GO
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    n2v::set_error("n2v_sgns_train_shared: launch failed: %s", cudaGetErrorString(err));
    return N2V_ERR_CUDA;
  }
  return N2V_OK;
}
