// K3: skip-gram negative-sampling SGD over the walk matrix.
//
// Replaces gensim 3.8's train_batch_sg / w2v_fast_sentence_sg_neg as reached from
// Node2VecGensim.fit (reference embedding.py:120-127) with sg=1, negative=K:
//   per sentence : drop out-of-vocabulary and sub-sampled tokens; per centre i a reduced
//                  window b in [0, window); contexts j in [i-window+b, i+window-b], j != i
//   per pair     : input row syn0[walk[j]]; targets walk[i] (label 1) then K negatives drawn
//                  ~ count^0.75 (label 0, skipped when equal to walk[i]); f = <in, target>;
//                  skipped when |f| >= 6; s = EXP_TABLE[f]; g = (label - s) * alpha;
//                  work += g * target; target += g * in; finally in += work.
// All arithmetic fp32.
//
// Execution model.  One WARP per walk (sentence); the sentence lives in shared memory after
// sub-sampling; rows are D fp32 held NV float4-per-lane in registers, so one row is one
// coalesced 512-byte gather per 128 dims, the dot is NV fused multiply-adds per lane plus a
// five-step butterfly, and the centre's positive row stays in registers across its whole
// context window (read once, its updates summed and reduced once per centre: within the warp
// this is exactly gensim's sequential order; other warps see the centre row's change a window
// later, which is ordinary Hogwild staleness).  Row updates are Hogwild: `red.global.add.v4.f32` (no lost updates under
// tens of thousands of concurrent warps) or plain stores (gensim's own lock-free behaviour).
// This is a gather/scatter path -- (K+2) rows in, (K+2) rows out per pair -- so no tensor
// cores; the walk buffer is read once per epoch.
#include <stdlib.h>

#include "n2v_internal.cuh"

namespace {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
#ifndef N2V_SGNS_MIN_BLOCKS
#define N2V_SGNS_MIN_BLOCKS 4  // measured: 4 blocks (64 regs) beats 1/3 (fewer warps) and 5/6 (spills)
#endif
#ifndef N2V_SGNS_MIN_BLOCKS_NV2
#define N2V_SGNS_MIN_BLOCKS_NV2 3
#endif
constexpr float kMaxExp = 6.0f;
constexpr int kDefaultMode = 0;   // see sgns_kernel MODE; chosen by measurement (profiles/README.md)

struct SgnsArgs {
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
  uint32_t n_top;   // entries of the top-level negative table (0 = single level)
  int32_t len, len_cap, dim, window, negative, epochs, epoch, batch_words;
  float alpha, min_alpha;
  uint32_t key0, key1;
};

// Per-walk generator for windows and negatives: PCG-RXS-M-XS-32 (O'Neill 2014) -- one 32-bit
// multiply-add per draw; the stream is seeded per (walk, epoch) from Philox.  (gensim uses a
// 48-bit LCG here; any uniform source gives the same law.)
__device__ __forceinline__ uint32_t pcg_next(uint32_t& state) {
  const uint32_t s = state;
  state = s * 747796405u + 2891336453u;
  const uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
  return (w >> 22u) ^ w;
}

template <int NV>
struct Row {
  float4 v[NV];
};

// FULL: dim == NV * 128, every lane owns NV whole float4 (no tail predicate)
template <int NV, bool FULL>
__device__ __forceinline__ Row<NV> load_row(const float* __restrict__ base, int32_t dim, int lane) {
  Row<NV> r;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int d = (i * 32 + lane) * 4;
    r.v[i] = (FULL || d < dim) ? *reinterpret_cast<const float4*>(base + d) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return r;
}

template <int NV>
__device__ __forceinline__ float dot_rows(const Row<NV>& a, const Row<NV>& b) {
  float p = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    p = fmaf(a.v[i].x, b.v[i].x, p);
    p = fmaf(a.v[i].y, b.v[i].y, p);
    p = fmaf(a.v[i].z, b.v[i].z, p);
    p = fmaf(a.v[i].w, b.v[i].w, p);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  return p;
}

// per-lane part of dot_rows (same FMA order), without the butterfly
template <int NV>
__device__ __forceinline__ float dot_partial(const Row<NV>& a, const Row<NV>& b) {
  float p = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    p = fmaf(a.v[i].x, b.v[i].x, p);
    p = fmaf(a.v[i].y, b.v[i].y, p);
    p = fmaf(a.v[i].z, b.v[i].z, p);
    p = fmaf(a.v[i].w, b.v[i].w, p);
  }
  return p;
}

// acc += g * x
template <int NV>
__device__ __forceinline__ void axpy(Row<NV>& acc, float g, const Row<NV>& x) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc.v[i].x = fmaf(g, x.v[i].x, acc.v[i].x);
    acc.v[i].y = fmaf(g, x.v[i].y, acc.v[i].y);
    acc.v[i].z = fmaf(g, x.v[i].z, acc.v[i].z);
    acc.v[i].w = fmaf(g, x.v[i].w, acc.v[i].w);
  }
}

// global row += g * x  (ATOMIC: vector reduction; else read-modify-write of the value we loaded)
template <int NV, bool ATOMIC, bool FULL>
__device__ __forceinline__ void update_row(float* __restrict__ base, int32_t dim, int lane, float g,
                                           const Row<NV>& x, const Row<NV>& loaded) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int d = (i * 32 + lane) * 4;
    if (FULL || d < dim) {
      float4* p = reinterpret_cast<float4*>(base + d);
      if (ATOMIC) {
        atomicAdd(p, make_float4(g * x.v[i].x, g * x.v[i].y, g * x.v[i].z, g * x.v[i].w));
      } else {
        *p = make_float4(fmaf(g, x.v[i].x, loaded.v[i].x), fmaf(g, x.v[i].y, loaded.v[i].y),
                         fmaf(g, x.v[i].z, loaded.v[i].z), fmaf(g, x.v[i].w, loaded.v[i].w));
      }
    }
  }
}

// global row += delta
template <int NV, bool ATOMIC, bool FULL>
__device__ __forceinline__ void add_row(float* __restrict__ base, int32_t dim, int lane, const Row<NV>& delta,
                                        const Row<NV>& loaded) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int d = (i * 32 + lane) * 4;
    if (FULL || d < dim) {
      float4* p = reinterpret_cast<float4*>(base + d);
      if (ATOMIC) {
        atomicAdd(p, delta.v[i]);
      } else {
        *p = make_float4(loaded.v[i].x + delta.v[i].x, loaded.v[i].y + delta.v[i].y,
                         loaded.v[i].z + delta.v[i].z, loaded.v[i].w + delta.v[i].w);
      }
    }
  }
}

constexpr int kExpTable = 1000;

// EXP_TABLE lookup (shared-memory copy of the host-built table; uniform index => broadcast)
__device__ __forceinline__ float sigmoid_table(const float* table, float f) {
  return table[static_cast<int>((f + kMaxExp) * (kExpTable / kMaxExp / 2.0f))];
}

// g = (label - sigmoid(f)) * alpha, or 0 when the target is skipped / clipped (|f| >= 6).
// Branch-free: the index is clamped into the table, the result is zeroed by `ok`.
__device__ __forceinline__ float gradient(const float* table, float f, float label, float alpha, bool ok) {
  const float fc = fminf(fmaxf(f, -kMaxExp), kMaxExp - 1e-3f);
  const float g = (label - sigmoid_table(table, fc)) * alpha;
  return ok ? g : 0.0f;
}

// one negative ~ count^0.75 from the (two-level) alias table; consumes 2 (4) draws of the walk's stream
__device__ __forceinline__ int32_t draw_negative(const SgnsArgs& A, uint32_t& rnd) {
  uint32_t lo = 0, span = A.n_vertices;
  if (A.n_top) {   // two-level table: chunk ~ mass first (top level is L2-resident)
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
  return (u2 < static_cast<uint32_t>(e.x)) ? static_cast<int32_t>(slot) : e.y;
}

// draw_negative split in two so that the alias-entry gather can stay in flight: `issue` consumes the draws
// and starts the gather, `resolve` (first use of the loaded entry) picks slot or alias
struct PendingNegative {
  int2 entry;
  uint32_t slot, u2;
};
__device__ __forceinline__ PendingNegative issue_negative(const SgnsArgs& A, uint32_t& rnd) {
  uint32_t lo = 0, span = A.n_vertices;
  if (A.n_top) {
    const uint32_t c0 = __umulhi(pcg_next(rnd), A.n_top);
    const int2 te = __ldg(A.neg_table + A.n_vertices + c0);
    const uint32_t c = (pcg_next(rnd) < static_cast<uint32_t>(te.x)) ? c0 : static_cast<uint32_t>(te.y);
    lo = c * N2V_NEG_CHUNK;
    span = min(static_cast<uint32_t>(N2V_NEG_CHUNK), A.n_vertices - lo);
  }
  PendingNegative p;
  const uint32_t u1 = pcg_next(rnd);
  p.u2 = pcg_next(rnd);
  p.slot = lo + __umulhi(u1, span);
  p.entry = __ldg(A.neg_table + p.slot);
  return p;
}
__device__ __forceinline__ int32_t resolve_negative(const PendingNegative& p) {
  return (p.u2 < static_cast<uint32_t>(p.entry.x)) ? static_cast<int32_t>(p.slot) : p.entry.y;
}

// MODE (latency hiding for tables beyond L2; every mode makes the same draws and the same arithmetic):
//   0  one negative at a time: draw -> alias entry -> row -> update (2 dependent round trips per negative)
//   1  the pair's K negatives drawn first (sequentially), their rows prefetched into L2 by one warp-wide
//      prefetch, then processed one by one
//   2  as 1, but lane d draws negative d from the walk's PCG stream jumped ahead to draw d's position (the
//      LCG state after n steps is s * a^n + c * (a^n - 1) / (a - 1)): the K alias-entry gathers become ONE
//      warp-wide gather instead of K dependent ones
//   3  as 2, and the next target row is loaded while the current one is processed (double buffering)
//   4  as 2 without the prefetch, and -- for K == 5 distinct targets -- all K rows are loaded at once into
//      registers (K row gathers in flight instead of a chain of K), their K dot products reduced by K
//      interleaved butterflies, then applied in the sequential order (positive, then negatives 0..K-1):
//      identical floating-point results; a pair with a repeated target takes the one-by-one loop
//   5  as 4, and the NEXT pair of the same centre is drawn one pair ahead (its position in the walk's stream
//      is known: K * draws_per_negative further on): its alias gathers fly during this pair's arithmetic and
//      its K rows are prefetched into L2 at the end of this pair
constexpr int kBatchK = 5;
template <int NV, bool ATOMIC, bool TRACE, bool FULL, int MODE>
__global__ void __launch_bounds__(kBlock, MODE >= 4 ? (NV == 1 ? 3 : 2)
                                                    : NV == 1 ? N2V_SGNS_MIN_BLOCKS : (NV == 2 ? N2V_SGNS_MIN_BLOCKS_NV2 : 1))
sgns_kernel(const __grid_constant__ SgnsArgs A) {
  extern __shared__ int32_t smem[];
  __shared__ float exp_table[kExpTable];
  for (int i = threadIdx.x; i < kExpTable; i += kBlock) exp_table[i] = __ldg(A.exp_table + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  int32_t* sent = smem + wib * A.len_cap;
  constexpr bool tracing = TRACE;
  if (tracing && (blockIdx.x != 0 || wib != 0)) return;   // trace mode: ONE warp, walks in order
  const int64_t n_warps = tracing ? 1 : static_cast<int64_t>(gridDim.x) * kWarps;
  unsigned long long c_pairs = 0, c_kept = 0;
  uint32_t c_negskip = 0, c_clip = 0;   // per-warp; far below 2^32 per launch share
  int64_t trace_pos = 0;
  const int K = A.negative;
  // PCG jump-ahead constants: (jA, jC) takes the stream from a pair's first negative draw to lane d's
  // negative (d * draws_per_negative steps), (kA, kC) past all K negatives
  uint32_t jA = 1u, jC = 0u, kA = 1u, kC = 0u;
  if (MODE >= 2) {
    const int per = A.n_top ? 4 : 2;
    for (int n = 0; n < K * per; ++n) {
      if (n == lane * per) { jA = kA; jC = kC; }
      kA *= 747796405u;
      kC = kC * 747796405u + 2891336453u;
    }
  }
  const double total_words = static_cast<double>(A.total_walks) * static_cast<double>(A.len);

  for (int64_t s = tracing ? 0 : static_cast<int64_t>(blockIdx.x) * kWarps + wib; s < A.n_walks; s += n_warps) {
    const int64_t gs = A.walk_offset + s;
    // learning rate of this walk's job (gensim: linear decay, refreshed per batch_words words)
    float alpha;
    {
      const int64_t job_start_words = (gs * A.len / A.batch_words) * static_cast<int64_t>(A.batch_words);
      const double progress = (static_cast<double>(A.epoch) + static_cast<double>(job_start_words) / total_words) /
                              static_cast<double>(A.epochs);
      const double a = static_cast<double>(A.alpha) - (static_cast<double>(A.alpha) - static_cast<double>(A.min_alpha)) * progress;
      alpha = static_cast<float>(a < static_cast<double>(A.min_alpha) ? static_cast<double>(A.min_alpha) : a);
    }
    // sub-sampling: position k is kept iff its token is in the vocabulary and u32 < keep_thr
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
    // per-walk PCG stream seeded from Philox; identical in every lane
    uint32_t rnd;
    {
      const uint4 r = n2v::philox4x32_10(A.key0, A.key1, static_cast<uint32_t>(gs), static_cast<uint32_t>(gs >> 32),
                                         static_cast<uint32_t>(A.epoch), 0x50434700u);
      rnd = r.x;
    }
    for (int i = 0; i < n; ++i) {
      const uint32_t b = __umulhi(pcg_next(rnd), static_cast<uint32_t>(A.window));   // uniform on [0, window)
      int j0 = i - A.window + static_cast<int>(b), j1 = i + A.window + 1 - static_cast<int>(b);
      if (j0 < 0) j0 = 0;
      if (j1 > n) j1 = n;
      if (j1 - j0 <= 1) continue;
      const int32_t wi = sent[i];
      float* pos_ptr = A.syn1neg + static_cast<int64_t>(wi) * A.dim;
      // the centre's row stays in registers across the window: every pair of this centre sees the
      // updates of the previous ones (gensim's sequential semantics inside the sentence); the sum of
      // its updates goes out as ONE reduction after the window instead of one per pair
      Row<NV> pos = load_row<NV, FULL>(pos_ptr, A.dim, lane);
      Row<NV> pos_delta;
#pragma unroll
      for (int q = 0; q < NV; ++q) pos_delta.v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      int32_t tgt_lookahead = 0;          // MODE 5: lane d's target of the NEXT pair, drawn one pair ahead
      bool have_lookahead = false;
      for (int j = j0; j < j1; ++j) {
        if (j == i) continue;
        const int32_t wj = sent[j];
        float* in_ptr = A.syn0 + static_cast<int64_t>(wj) * A.dim;
        // K negatives ~ count^0.75, one 8-byte alias gather each.  MODE >= 1: the K targets of the pair are
        // drawn up front (the same draws in the same order as drawing them one by one: nothing else consumes
        // the walk's stream in between), lane d keeps target d, and their rows are prefetched into L2
        // with ONE warp-wide prefetch (a lane per 128-byte line) while the input row is loaded and the
        // positive target is processed.
        int32_t my_tgt = 0;
        if (MODE >= 1) {
          if (MODE == 1) {
            for (int d = 0; d < K; ++d) {
              const int32_t tgt = draw_negative(A, rnd);
              if (lane == d) my_tgt = tgt;
            }
          } else if (MODE == 5 && have_lookahead) {
            my_tgt = tgt_lookahead;                           // drawn during the previous pair
            rnd = rnd * kA + kC;
          } else {
            uint32_t mine = rnd * jA + jC;                    // the stream at lane d's negative
            if (lane < K) my_tgt = draw_negative(A, mine);    // K alias gathers in one warp-wide load
            rnd = rnd * kA + kC;                              // past all K negatives
          }
          const int lines = (A.dim * 4 + 127) >> 7;          // 128-byte lines per row
          for (int first = 0; MODE < 4 && first < K * lines; first += 32) {
            const int idx = first + lane;
            const int d = min(idx / lines, K - 1);
            const int32_t t = __shfl_sync(0xffffffffu, my_tgt, d);
            if (idx < K * lines) {
              const float* line = A.syn1neg + static_cast<int64_t>(t) * A.dim + (idx - d * lines) * 32;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
            }
          }
          if (j + 1 < j1 && lane < lines) {                  // next context row of this centre (skips i itself)
            const int jn = (j + 1 == i) ? j + 2 : j + 1;
            if (jn < j1) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.syn0 + static_cast<int64_t>(sent[jn]) * A.dim + lane * 32));
          }
        }
        // MODE 5: start the next pair's draws now (rnd already stands at the next pair's first draw)
        PendingNegative pending;
        pending.slot = 0; pending.u2 = 0; pending.entry = make_int2(0, 0);
        if (MODE == 5) {
          const int jn = (j + 1 == i) ? j + 2 : j + 1;
          have_lookahead = jn < j1;
          if (have_lookahead && lane < K) {
            uint32_t mine = rnd * jA + jC;
            pending = issue_negative(A, mine);
          }
        }
        // MODE >= 4: the K target rows, all in flight at once (only when the K targets are distinct)
        Row<NV> rows[MODE >= 4 ? kBatchK : 1];
        bool batched = false;
        if (MODE >= 4) {
          const unsigned same = __match_any_sync(0xffffffffu, lane < K ? my_tgt : -1 - lane);
          batched = K == kBatchK && !__any_sync(0xffffffffu, lane < K && __popc(same) > 1);
          if (batched) {
#pragma unroll
            for (int d = 0; d < kBatchK; ++d)
              rows[d] = load_row<NV, FULL>(A.syn1neg + static_cast<int64_t>(__shfl_sync(0xffffffffu, my_tgt, d)) * A.dim,
                                           A.dim, lane);
          }
        }
        const Row<NV> in = load_row<NV, FULL>(in_ptr, A.dim, lane);
        Row<NV> work;
#pragma unroll
        for (int q = 0; q < NV; ++q) work.v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        int32_t* trow = nullptr;
        if (TRACE && trace_pos < A.trace_cap) {
          trow = A.trace + trace_pos * (2 + K);
          if (lane == 0) { trow[0] = wi; trow[1] = wj; A.trace_alpha[trace_pos] = alpha; }
        }
        // positive target: the centre word (row in registers).  Branch-free: a clipped target
        // (|f| >= 6, gensim skips it) gets g = 0 and its reduction adds zeros.
        {
          const float f = dot_rows<NV>(in, pos);
          const bool ok = f > -kMaxExp && f < kMaxExp;
          const float g = gradient(exp_table, f, 1.0f, alpha, ok);
          if (TRACE) c_clip += ok ? 0u : 1u;
          axpy<NV>(work, g, pos);
          axpy<NV>(pos, g, in);
          if (ATOMIC) axpy<NV>(pos_delta, g, in);
        }
        if (MODE >= 4 && batched) {
          float f[kBatchK];
#pragma unroll
          for (int d = 0; d < kBatchK; ++d) f[d] = dot_partial<NV>(in, rows[d]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int d = 0; d < kBatchK; ++d) f[d] += __shfl_xor_sync(0xffffffffu, f[d], o);
          }
#pragma unroll
          for (int d = 0; d < kBatchK; ++d) {
            const int32_t tgt = __shfl_sync(0xffffffffu, my_tgt, d);
            const bool skip = tgt == wi;
            if (TRACE && trow && lane == 0) trow[2 + d] = skip ? -1 : tgt;
            const bool in_range = f[d] > -kMaxExp && f[d] < kMaxExp;
            const bool ok = in_range && !skip;
            const float g = gradient(exp_table, f[d], 0.0f, alpha, ok);
            if (TRACE) {
              c_negskip += skip ? 1u : 0u;
              c_clip += (!skip && !in_range) ? 1u : 0u;
            }
            axpy<NV>(work, g, rows[d]);
            float* t_ptr = A.syn1neg + static_cast<int64_t>(tgt) * A.dim;
            if (ATOMIC || ok) update_row<NV, ATOMIC, FULL>(t_ptr, A.dim, lane, g, in, rows[d]);
          }
        }
        // the K negative targets, in order (drawn at the top of the pair when MODE >= 1, else one by one here)
        Row<NV> ahead;                                        // MODE 3: row of the NEXT target, loaded early
        int32_t tgt_ahead = 0;
        if (MODE == 3) {
          tgt_ahead = __shfl_sync(0xffffffffu, my_tgt, 0);
          ahead = load_row<NV, FULL>(A.syn1neg + static_cast<int64_t>(tgt_ahead) * A.dim, A.dim, lane);
        }
        for (int d = 0; d < (MODE >= 4 && batched ? 0 : K); ++d) {
          const int32_t tgt = MODE == 3 ? tgt_ahead : MODE >= 1 ? __shfl_sync(0xffffffffu, my_tgt, d) : draw_negative(A, rnd);
          const bool skip = tgt == wi;                      // gensim: a negative equal to the centre is skipped
          if (TRACE && trow && lane == 0) trow[2 + d] = skip ? -1 : tgt;
          float* t_ptr = A.syn1neg + static_cast<int64_t>(tgt) * A.dim;
          Row<NV> tr;
          if (MODE == 3) {
            tr = ahead;
            if (d + 1 < K) {
              tgt_ahead = __shfl_sync(0xffffffffu, my_tgt, d + 1);
              // a repeated target must see this target's update: it is (re)loaded after the update below
              if (tgt_ahead != tgt)
                ahead = load_row<NV, FULL>(A.syn1neg + static_cast<int64_t>(tgt_ahead) * A.dim, A.dim, lane);
            }
          } else {
            tr = load_row<NV, FULL>(t_ptr, A.dim, lane);
          }
          const float f = dot_rows<NV>(in, tr);
          const bool in_range = f > -kMaxExp && f < kMaxExp;
          const bool ok = in_range && !skip;
          const float g = gradient(exp_table, f, 0.0f, alpha, ok);
          if (TRACE) {   // diagnostics only in the single-warp trace build
            c_negskip += skip ? 1u : 0u;
            c_clip += (!skip && !in_range) ? 1u : 0u;
          }
          axpy<NV>(work, g, tr);
          if (ATOMIC || ok) update_row<NV, ATOMIC, FULL>(t_ptr, A.dim, lane, g, in, tr);
          if (MODE == 3 && d + 1 < K && tgt_ahead == tgt)
            ahead = load_row<NV, FULL>(t_ptr, A.dim, lane);
        }
        add_row<NV, ATOMIC, FULL>(in_ptr, A.dim, lane, work, in);
        if (MODE == 5 && have_lookahead) {                   // the next pair's targets: resolve, prefetch their rows
          tgt_lookahead = lane < K ? resolve_negative(pending) : 0;
          const int lines = (A.dim * 4 + 127) >> 7;
          for (int first = 0; first < K * lines; first += 32) {
            const int idx = first + lane;
            const int d = min(idx / lines, K - 1);
            const int32_t t = __shfl_sync(0xffffffffu, tgt_lookahead, d);
            if (idx < K * lines) {
              const float* line = A.syn1neg + static_cast<int64_t>(t) * A.dim + (idx - d * lines) * 32;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
            }
          }
        }
        ++c_pairs;
        if (TRACE) ++trace_pos;
      }
      // centre row: += sum of its updates (ATOMIC), or the register copy stored back (plain Hogwild)
      if (ATOMIC) {
        add_row<NV, true, FULL>(pos_ptr, A.dim, lane, pos_delta, pos);
      } else {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          const int d = (q * 32 + lane) * 4;
          if (FULL || d < A.dim) *reinterpret_cast<float4*>(pos_ptr + d) = pos.v[q];
        }
      }
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

template <int NV, bool FULL, int MODE>
cudaError_t launch_full(const SgnsArgs& A, bool atomic, int grid, size_t smem, cudaStream_t stream) {
#define N2V_SGNS_GO(AT, TR)                                                                                        \
  do {                                                                                                             \
    if (smem > 48 * 1024)                                                                                          \
      cudaFuncSetAttribute(sgns_kernel<NV, AT, TR, FULL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                           static_cast<int>(smem));                                                                \
    sgns_kernel<NV, AT, TR, FULL, MODE><<<grid, kBlock, smem, stream>>>(A);                                   \
  } while (0)
  if (A.trace) {
    if (atomic) N2V_SGNS_GO(true, true); else N2V_SGNS_GO(false, true);
  } else {
    if (atomic) N2V_SGNS_GO(true, false); else N2V_SGNS_GO(false, false);
  }
#undef N2V_SGNS_GO
  return cudaGetLastError();
}

template <int NV, bool FULL>
cudaError_t launch_mode(const SgnsArgs& A, bool atomic, int mode, int grid, size_t smem, cudaStream_t stream) {
  switch (mode) {
    case 1: return launch_full<NV, FULL, 1>(A, atomic, grid, smem, stream);
    case 2: return launch_full<NV, FULL, 2>(A, atomic, grid, smem, stream);
    case 3: return launch_full<NV, FULL, 3>(A, atomic, grid, smem, stream);
    case 4: return NV <= 2 ? launch_full<NV, FULL, (NV <= 2 ? 4 : 2)>(A, atomic, grid, smem, stream)
                           : launch_full<NV, FULL, 2>(A, atomic, grid, smem, stream);
    case 5: return NV <= 2 ? launch_full<NV, FULL, (NV <= 2 ? 5 : 2)>(A, atomic, grid, smem, stream)
                           : launch_full<NV, FULL, 2>(A, atomic, grid, smem, stream);
    default: return launch_full<NV, FULL, 0>(A, atomic, grid, smem, stream);
  }
}

template <int NV>
cudaError_t launch(const SgnsArgs& A, bool atomic, int mode, int grid, size_t smem, cudaStream_t stream) {
  return A.dim == NV * 128 ? launch_mode<NV, true>(A, atomic, mode, grid, smem, stream)
                           : launch_mode<NV, false>(A, atomic, mode, grid, smem, stream);
}

}  // namespace

extern "C" int n2v_sgns_train(const int32_t* walks, int64_t n_walks, int32_t len, int64_t pitch,
                              const uint32_t* keep_thr, const int32_t* neg_table, int64_t n_vertices,
                              float* syn0, float* syn1neg, const float* exp_table,
                              const n2v_sgns_params_t* P, uint64_t* stats, int32_t* trace, float* trace_alpha,
                              int64_t trace_cap, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(P != nullptr, "n2v_sgns_train: params is NULL");
  N2V_CHECK_ARG(P->dim >= 4 && P->dim <= 1024 && P->dim % 4 == 0,
                "n2v_sgns_train: dim %d must be a multiple of 4 in [4, 1024]", P->dim);
  N2V_CHECK_ARG(P->window >= 1 && P->negative >= 1 && P->negative <= 64,
                "n2v_sgns_train: window (%d) must be >= 1 and negative (%d) in [1, 64]", P->window, P->negative);
  N2V_CHECK_ARG(P->epochs >= 1 && P->epoch >= 0 && P->epoch < P->epochs && P->batch_words >= 1,
                "n2v_sgns_train: bad epoch schedule");
  N2V_CHECK_ARG(n_walks >= 0 && len >= 1 && pitch >= len && len <= 10000,
                "n2v_sgns_train: bad walk matrix shape (sentences are capped at 10000 tokens)");
  N2V_CHECK_ARG(n_vertices > 0 && n_vertices < (int64_t(1) << 31), "n2v_sgns_train: n_vertices out of range");
  if (n_walks == 0) return N2V_OK;
  N2V_CHECK_ARG(walks && keep_thr && neg_table && syn0 && syn1neg && exp_table, "n2v_sgns_train: NULL buffer");
  N2V_CHECK_ARG(trace == nullptr || trace_alpha != nullptr, "n2v_sgns_train: trace needs trace_alpha");
  N2V_CHECK_ARG((reinterpret_cast<uintptr_t>(syn0) & 15) == 0 && (reinterpret_cast<uintptr_t>(syn1neg) & 15) == 0,
                "n2v_sgns_train: tables must be 16-byte aligned");

  SgnsArgs A{};
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
  A.negative = P->negative;
  A.epochs = P->epochs;
  A.epoch = P->epoch;
  A.batch_words = P->batch_words;
  A.alpha = P->alpha;
  A.min_alpha = P->min_alpha;
  A.key0 = static_cast<uint32_t>(P->seed);
  A.key1 = static_cast<uint32_t>(P->seed >> 32);

  const size_t smem = static_cast<size_t>(A.len_cap) * kWarps * sizeof(int32_t);
  N2V_CHECK_ARG(smem <= 200 * 1024, "n2v_sgns_train: walk too long for shared-memory staging");
  int grid;
  if (trace) grid = 1;  // the kernel lets only warp 0 work in trace mode
  else {
    const int64_t need = (n_walks + kWarps - 1) / kWarps;
    const int64_t cap = int64_t(n2v::sm_count()) * 8;
    grid = static_cast<int>(need < cap ? need : cap);
  }
  const int nv = (P->dim + 127) / 128;
  cudaError_t err;
  const bool atomic = P->atomic_updates != 0;
  // latency-hiding mode (see sgns_kernel), chosen by measurement (profiles/r02_sgns_modes.txt, ms per epoch
  // for modes 0 / 2 / 4): configs[1] tables 10 MB 144.9 / 152.9 / 140.2; configs[2] 1.07 GB 5109 / 4250 /
  // 4117 (D = 256: 9142 / 7925 / 7283); 16 M-row tables 17 GB 6552 / 4611 / 4761.  So: K = 5 and D <= 256
  // and tables <= 4 GB -> mode 4 (all K rows in flight); larger tables -> mode 2 (lane-parallel draws +
  // L2 prefetch); otherwise L2-resident tables keep mode 0.  N2V_SGNS_MODE=0..4 overrides (tests, tuning).
  const double table_bytes = 2.0 * static_cast<double>(n_vertices) * P->dim * 4.0;
  int mode = (P->negative == kBatchK && P->dim <= 256 && table_bytes <= 4.0e9) ? 4 : table_bytes > 96.0e6 ? 2 : kDefaultMode;
  if (const char* e = getenv("N2V_SGNS_MODE")) mode = atoi(e);
  if (mode < 0 || mode > 5 || P->negative > 32) mode = 0;
#define N2V_SGNS_LAUNCH(NVV) err = launch<NVV>(A, atomic, mode, grid, smem, stream)
  if (nv <= 1) N2V_SGNS_LAUNCH(1);
  else if (nv <= 2) N2V_SGNS_LAUNCH(2);
  else if (nv <= 4) N2V_SGNS_LAUNCH(4);
  else N2V_SGNS_LAUNCH(8);
#undef N2V_SGNS_LAUNCH
  if (err != cudaSuccess) {
    n2v::set_error("n2v_sgns_train: launch failed: %s", cudaGetErrorString(err));
    return N2V_ERR_CUDA;
  }
  return N2V_OK;
}
