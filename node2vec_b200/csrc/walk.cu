// K2: second-order (p, q) biased random walks over the packed CSR.
//
// Replaces, for every walker at once, the reference's
//   initiate_random_walk                      (randomwalk.py:279-296)
//   [left join on src, inner join on dst]     (fugue.py:147)
//   next_step_random_walk  x walk_length      (randomwalk.py:300-339), which per row
//       unpickles two adjacency strings, builds a Python set of N_out(prev), rebuilds an
//       O(deg) alias table (generate_edge_alias_tables :193-232) and draws once (:86-99)
//   to_path                                   (randomwalk.py:343-349)
//
// Sampling.  The reference's law for the next vertex x, given the previous vertex t and
// the current vertex v, is  P(x) ~ w(v,x) * alpha(t,x)  with alpha = 1/p (x == t),
// 1 (x in N_out(t)), 1/q (otherwise) (randomwalk.py:223-230).  We draw from exactly that
// law without building the per-(t,v) table: propose x from v's FIRST-ORDER alias table
// (one 32-byte gather that also carries x's adjacency header, so an accepted step needs no
// further lookup), accept with alpha/cap; membership in N_out(t) is one 32-byte gather
// from t's bucketed hash set and is skipped whenever the accept draw already decides
// (u below both thresholds / above both).  Both gathers are single LDG.256 requests.  The return arc's excess mass (1/p above
// cap) is a separate mixture component (the "fold"), so small p does not inflate the
// envelope.
//
// Execution model.  One thread per walker, lanes fully asynchronous: the loop body is ONE
// proposal; a lane whose proposal is accepted advances its own step counter, a lane that
// finishes its walk picks up the next walker (grid-stride).  Nobody waits for a
// neighbour's rejections.  Walk rows are staged 8 ids at a time per lane and written as
// whole 32-byte sectors.
//
// Determinism.  Philox4x32-10, key = seed, counter = (walk_id lo, walk_id hi, step,
// trial); all sampling decisions are integer compares, the fold threshold is three IEEE
// fp32 ops.  oracle/csrc/n2v_oracle.c (orc_replay_walk) restates this on the host bit-for-bit.
#include "n2v_internal.cuh"

namespace {

#ifndef N2V_WALK_BLOCK
#define N2V_WALK_BLOCK 256
#endif
constexpr int kBlock = N2V_WALK_BLOCK;
#ifndef N2V_WALK_BLOCKS_PER_SM
#define N2V_WALK_BLOCKS_PER_SM 6  // measured on B200 (v3 kernel): 6 blocks = 40 regs beats 8 (spills) and 5
#endif
constexpr int kBlocksPerSm = N2V_WALK_BLOCKS_PER_SM;
constexpr int kStage = 8;  // ids per lane per flush = one 32 B sector

struct WalkArgs {
  const int32_t* start;
  int32_t* walks;
  uint8_t* alive;
  unsigned long long* stats;
  int64_t pitch;
  uint32_t n_walkers;  // per launch (< 2^31; the host chunks larger jobs)
  uint32_t walker0;    // global index of this launch's first walker
  int32_t num_walks;
  int32_t walk_length;
  uint32_t key0, key1;
  // accept iff u32 <= *_m1
  uint32_t ret_m1, nbr_m1, far_m1, lo_m1, hi_m1;
  float fold_gain;
  float mix_qm1;
  int32_t max_trials;
  double inv_p, inv_q;
};

// x in N_out(t)?  t's hash set: nb buckets of 8 slots starting at bucket hbase
__device__ __forceinline__ bool member(const int32_t* __restrict__ hash, uint32_t hbase, uint32_t deg,
                                       int32_t x, uint32_t& probes) {
  const uint32_t nb = n2v_hash_nbuckets(deg);
  uint32_t b = __umulhi(static_cast<uint32_t>(x) * N2V_HASH_MULT, nb);
  for (;;) {
    const n2v::Int8 s = n2v::load_sector(hash + (static_cast<size_t>(hbase) + b) * N2V_HASH_SLOTS);
    ++probes;
    if (s.a[0] == x || s.a[1] == x || s.a[2] == x || s.a[3] == x || s.a[4] == x || s.a[5] == x || s.a[6] == x ||
        s.a[7] == x)
      return true;
    if (s.a[7] == N2V_HASH_EMPTY) return false;  // bucket not full: x would have been here
    b = (b + 1 == nb) ? 0u : b + 1;
  }
}

// lower_bound membership in an ascending slice (cold path: exact fallback only)
__device__ bool member_sorted(const int32_t* __restrict__ col, uint32_t n, int32_t x) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (col[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo < n && col[lo] == x;
}

// Exact O(deg log deg) draw from the biased law; used only after max_trials rejections so
// that extreme (p, q) on a hub cannot stall a lane.  Sequential fp64, no contraction.
__device__ __noinline__ uint32_t exact_draw(const int32_t* __restrict__ vcol, const double* __restrict__ vw,
                                           uint32_t deg, int32_t t, const int32_t* __restrict__ tcol,
                                           uint32_t tdeg, double inv_p, double inv_q, uint32_t r0,
                                           uint32_t r1) {
  double total = 0.0;
  for (uint32_t i = 0; i < deg; ++i) {
    const int32_t x = vcol[i];
    const double a = (x == t) ? inv_p : (member_sorted(tcol, tdeg, x) ? 1.0 : inv_q);
    total = __dadd_rn(total, __dmul_rn(vw ? vw[i] : 1.0, a));   // vw == NULL: unit-weight graph stored without weights
  }
  // 53-bit uniform in [0,1) from two 32-bit lanes
  const double u = __dmul_rn(__dadd_rn(__dmul_rn(static_cast<double>(r0 >> 5), 67108864.0),
                                       static_cast<double>(r1 >> 6)),
                             1.0 / 9007199254740992.0);
  const double target = __dmul_rn(u, total);
  double run = 0.0;
  uint32_t last = deg - 1;
  for (uint32_t i = 0; i < deg; ++i) {
    const int32_t x = vcol[i];
    const double a = (x == t) ? inv_p : (member_sorted(tcol, tdeg, x) ? 1.0 : inv_q);
    const double m = __dmul_rn(vw ? vw[i] : 1.0, a);
    run = __dadd_rn(run, m);
    if (m > 0.0) last = i;
    if (target < run) return i;
  }
  return last;   // index into v's arc slice
}

__device__ __forceinline__ uint32_t fold_threshold(float gain, uint32_t deg) {
  // P(return via fold) = gain / (deg + gain) on a unit-weight symmetric simple graph
  const float pr = __fdiv_rn(gain, __fadd_rn(static_cast<float>(deg), gain));
  return pr >= 1.0f ? 0xFFFFFFFFu : __float2uint_rz(__fmul_rn(pr, 4294967296.0f));
}

// general fold: P(return via fold) = e*rev / (1 + e*rev), rev = return mass / out-weight of v
__device__ __forceinline__ uint32_t fold_threshold_ratio(float gain, float rev) {
  const float num = __fmul_rn(gain, rev);
  const float pr = __fdiv_rn(num, __fadd_rn(1.0f, num));
  return pr >= 1.0f ? 0xFFFFFFFFu : __float2uint_rz(__fmul_rn(pr, 4294967296.0f));
}

// mixture sampler (mode 3, n2v_b200.h): component thresholds at (prev t, cur v) from the two degrees
__device__ __forceinline__ void mix_thresholds(float gain_r, float qm1, uint32_t deg_v, uint32_t deg_t,
                                               uint32_t& thr_ret, uint32_t& thr_out) {
  const float fo = __fmul_rn(static_cast<float>(min(deg_v, deg_t)), qm1);
  const float tot = __fadd_rn(__fadd_rn(static_cast<float>(deg_v), fo), gain_r);
  const float pr = __fdiv_rn(gain_r, tot);
  const float po = __fdiv_rn(fo, tot);
  thr_ret = __float2uint_rz(__fmul_rn(pr, 4294967296.0f));
  const float s = __fmul_rn(__fadd_rn(pr, po), 4294967296.0f);
  thr_out = s >= 4294967296.0f ? 0xFFFFFFFFu : __float2uint_rz(s);
}

// FOLD: 0 = off, 1 = unit-weight symmetric simple graph, 2 = general (per-arc {fwd, rev} ratios),
//       3 = mixture sampler (unit-weight symmetric simple graph, q > 1)
template <int FOLD, bool MULTI, bool STATS>
__global__ void __launch_bounds__(kBlock, kBlocksPerSm)
walk_kernel(const __grid_constant__ n2v_graph_t g, const __grid_constant__ WalkArgs A) {
  __shared__ int32_t stage[kStage * kBlock];  // [slot][thread]: conflict-free
  const uint32_t tid = threadIdx.x;
  const uint32_t stride = gridDim.x * kBlock;
  uint32_t w = blockIdx.x * kBlock + tid;

  uint32_t c_steps = 0, c_trials = 0, c_probes = 0, c_search = 0, c_fold = 0, c_fb = 0, c_dead = 0;

  // walker state (offsets are 32-bit: a part holds < 2^32 arcs / buckets)
  int32_t t = -1, v = 0, pos = 0;
  uint32_t deg_t = 0, deg_v = 0, base_v = 0, base_t = 0;
  uint32_t trial = 0, thr_out = 0, thr_ret = 0, wid_lo = 0, wid_hi = 0;
  uint32_t part_v = 0, part_t = 0;
  float r_fwd = 0.f, r_rev = 0.f;   // FOLD == 2: ratios of the arc that brought the walker to v
  bool active = false;
  const int32_t L = A.walk_length;

  auto row_ptr = [&](int chunk) {
    return reinterpret_cast<int4*>(A.walks + static_cast<int64_t>(w) * A.pitch + chunk * kStage);
  };
  auto flush_chunk = [&](int chunk) {
    int4 a, b;
    a.x = stage[0 * kBlock + tid]; a.y = stage[1 * kBlock + tid];
    a.z = stage[2 * kBlock + tid]; a.w = stage[3 * kBlock + tid];
    b.x = stage[4 * kBlock + tid]; b.y = stage[5 * kBlock + tid];
    b.z = stage[6 * kBlock + tid]; b.w = stage[7 * kBlock + tid];
    int4* dst = row_ptr(chunk);
    __stcs(dst, a);       // streaming: the walk matrix is write-once
    __stcs(dst + 1, b);
  };
  // end of a walk (complete or dropped): pad the row with -1 and release the lane
  auto finish = [&](bool is_alive) {
    for (int s = (pos & (kStage - 1)) + 1; s < kStage; ++s) stage[s * kBlock + tid] = -1;
    int chunk = pos >> 3;
    flush_chunk(chunk);
    const int4 neg = make_int4(-1, -1, -1, -1);
    for (++chunk; chunk * kStage < A.pitch; ++chunk) {
      int4* dst = row_ptr(chunk);
      __stcs(dst, neg);
      __stcs(dst + 1, neg);
    }
    A.alive[w] = is_alive ? 1 : 0;
    active = false;
    w += stride;
  };
  auto part_of = [&](int32_t x) -> uint32_t { return MULTI ? static_cast<uint32_t>(x / g.part_size) : 0u; };
  auto local_of = [&](int32_t x, uint32_t part) -> uint32_t {
    return MULTI ? static_cast<uint32_t>(x - part * g.part_size) : static_cast<uint32_t>(x);
  };
  // adjacency header by vertex id: only at walk start and after the (rare) exact fallback
  auto enter = [&](int32_t x, uint32_t& part, uint32_t& base, uint32_t& deg) {
    part = part_of(x);
    const uint4 rec = n2v::load_vtx(g.parts[part].vtx + local_of(x, part));
    base = rec.x;
    deg = rec.y;
  };

  for (;;) {
    if (!active) {
      if (w >= A.n_walkers) break;
      const uint32_t gw = A.walker0 + w;
      const uint32_t s = gw / static_cast<uint32_t>(A.num_walks);
      const uint32_t r = gw - s * static_cast<uint32_t>(A.num_walks);
      v = __ldg(A.start + s);
      const uint64_t walk_id = static_cast<uint64_t>(static_cast<uint32_t>(v)) * static_cast<uint32_t>(A.num_walks) + r;
      wid_lo = static_cast<uint32_t>(walk_id);
      wid_hi = static_cast<uint32_t>(walk_id >> 32);
      if (v >= 0 && v < g.n_vertices) {
        enter(v, part_v, base_v, deg_v);
      } else {   // a start id outside the graph has no adjacency row: dropped like a sink
        part_v = 0;
        base_v = 0;
        deg_v = 0;
      }
      t = -1;
      pos = 0;
      trial = 0;
      stage[tid] = v;
      active = true;
    }
    if (deg_v == 0) {  // inner join with df_dst drops the row (fugue.py:147)
      if (STATS) ++c_dead;
      finish(false);
      continue;
    }
    const uint4 rnd = n2v::philox4x32_10(A.key0, A.key1, wid_lo, wid_hi, static_cast<uint32_t>(pos), trial);
    const bool first = (pos == 0);
    bool accept;
    int32_t x;
    uint32_t base_x, deg_x;   // adjacency header of x, delivered with the proposal
    uint32_t arc_index = 0xFFFFFFFFu;   // part-local index of the arc taken (FOLD == 2)
    if (FOLD == 3 && !first) {
      if (rnd.x < thr_ret) {               // return component
        x = t;
        base_x = base_t;
        deg_x = deg_t;
        accept = true;
        if (STATS) ++c_fold;
      } else {
        const bool common = rnd.x < thr_out;          // common-neighbour proposal, else bulk
        const bool from_t = common && deg_t < deg_v;  // propose from the smaller adjacency
        const uint32_t k = __umulhi(rnd.y, from_t ? deg_t : deg_v);
        const n2v::Int8 arc = n2v::load_sector(g.parts[MULTI ? (from_t ? part_t : part_v) : 0].arcs +
                                               (from_t ? base_t : base_v) + k);
        const bool self = rnd.z < static_cast<uint32_t>(arc.a[0]);
        x = self ? arc.a[1] : arc.a[2];
        base_x = static_cast<uint32_t>(self ? arc.a[4] : arc.a[6]);
        deg_x = static_cast<uint32_t>(self ? arc.a[5] : arc.a[7]);
        if (STATS) ++c_trials;
        if (!common) {
          accept = (x != t) || (rnd.w <= A.ret_m1);
        } else {
          if (STATS) ++c_search;
          uint32_t probes = 0;
          const bool in = from_t
              ? member(g.parts[MULTI ? part_v : 0].hash, n2v_hash_base(base_v, local_of(v, part_v)), deg_v, x, probes)
              : member(g.parts[MULTI ? part_t : 0].hash, n2v_hash_base(base_t, local_of(t, part_t)), deg_t, x, probes);
          if (STATS) c_probes += probes;
          accept = in && x != t;
        }
      }
    } else if (FOLD != 0 && FOLD != 3 && !first && rnd.x < thr_out) {
      x = t;
      base_x = base_t;
      deg_x = deg_t;
      accept = true;
      if (STATS) ++c_fold;
    } else {
      const uint32_t k = __umulhi(rnd.y, deg_v);
      const n2v::Int8 arc = n2v::load_sector(g.parts[MULTI ? part_v : 0].arcs + base_v + k);
      const bool self = rnd.z < static_cast<uint32_t>(arc.a[0]);
      x = self ? arc.a[1] : arc.a[2];
      base_x = static_cast<uint32_t>(self ? arc.a[4] : arc.a[6]);
      deg_x = static_cast<uint32_t>(self ? arc.a[5] : arc.a[7]);
      if (FOLD == 2) arc_index = base_v + (self ? k : static_cast<uint32_t>(arc.a[3]));
      if (STATS) ++c_trials;
      if (first) accept = true;                       // unbiased first step (randomwalk.py:320-321)
      else if (x == t) accept = rnd.w <= A.ret_m1;
      else if (rnd.w <= A.lo_m1) accept = true;       // below both thresholds: no lookup needed
      else if (rnd.w > A.hi_m1) accept = false;       // above both
      else {
        if (STATS) ++c_search;
        uint32_t probes = 0;
        const bool in = member(g.parts[MULTI ? part_t : 0].hash, n2v_hash_base(base_t, local_of(t, part_t)), deg_t,
                               x, probes);
        if (STATS) c_probes += probes;
        accept = rnd.w <= (in ? A.nbr_m1 : A.far_m1);
      }
    }
    if (!accept) {
      if (++trial < static_cast<uint32_t>(A.max_trials)) continue;
      const uint4 r2 = n2v::philox4x32_10(A.key0, A.key1, wid_lo, wid_hi, static_cast<uint32_t>(pos), 0xFFFFFFFFu);
      const n2v_graph_part_t& PV = g.parts[MULTI ? part_v : 0];
      const n2v_graph_part_t& PT = g.parts[MULTI ? part_t : 0];
      const uint32_t pick = exact_draw(PV.col + base_v, PV.weight ? PV.weight + base_v : nullptr, deg_v, t, PT.col + base_t, deg_t,
                                       A.inv_p, A.inv_q, r2.x, r2.y);
      x = PV.col[base_v + pick];
      arc_index = base_v + pick;
      uint32_t px;
      enter(x, px, base_x, deg_x);
      if (STATS) ++c_fb;
    }
    // advance: v becomes the previous vertex
    t = v;
    base_t = base_v;
    deg_t = deg_v;
    part_t = part_v;
    v = x;
    base_v = base_x;
    deg_v = deg_x;
    if (MULTI) part_v = part_of(x);
    if (FOLD == 2) {
      if (arc_index != 0xFFFFFFFFu) {   // arrived through an arc: its ratios ride along
        const float2 r = __ldg(reinterpret_cast<const float2*>(g.parts[0].ratio) + arc_index);
        r_fwd = r.x;
        r_rev = r.y;
      } else {                           // arrived through the fold (back along the same arc): swap
        const float tmp = r_fwd;
        r_fwd = r_rev;
        r_rev = tmp;
      }
    }
    ++pos;
    if (STATS) ++c_steps;
    trial = 0;
    stage[(pos & (kStage - 1)) * kBlock + tid] = v;
    if (pos == L) {
      finish(true);
      continue;
    }
    if ((pos & (kStage - 1)) == kStage - 1) flush_chunk(pos >> 3);
    if (FOLD == 1) thr_out = fold_threshold(A.fold_gain, deg_v);
    if (FOLD == 2) thr_out = fold_threshold_ratio(A.fold_gain, r_rev);
    if (FOLD == 3) mix_thresholds(A.fold_gain, A.mix_qm1, deg_v, deg_t, thr_ret, thr_out);
  }

  if (STATS) {  // one atomic per counter per warp
    unsigned long long vals[7] = {c_steps, c_trials, c_probes, c_search, c_fold, c_fb, c_dead};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      unsigned long long x = vals[i];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if ((tid & 31) == 0 && x) atomicAdd(A.stats + i, x);
    }
  }
}

}  // namespace

extern "C" int n2v_walk(const n2v_graph_t* graph, const int32_t* start, int64_t n_start,
                        int32_t num_walks, int32_t walk_length, double return_param,
                        double inout_param, uint64_t seed, int32_t* walks, int64_t pitch,
                        uint8_t* alive, n2v_walk_stats_t* stats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(graph != nullptr, "n2v_walk: graph is NULL");
  N2V_CHECK_ARG(graph->n_parts >= 1 && graph->n_parts <= N2V_MAX_PARTS, "n2v_walk: n_parts %d out of range", graph->n_parts);
  N2V_CHECK_ARG(num_walks >= 1 && walk_length >= 1, "n2v_walk: num_walks (%d) and walk_length (%d) must be >= 1",
                num_walks, walk_length);
  N2V_CHECK_ARG(pitch >= walk_length + 1 && pitch % kStage == 0,
                "n2v_walk: pitch %lld must be a multiple of 8 and >= walk_length+1", static_cast<long long>(pitch));
  N2V_CHECK_ARG(n_start >= 0, "n2v_walk: negative n_start");
  n2v_walk_consts_t C;
  const int has_ratio = (graph->n_parts == 1 && graph->parts[0].ratio != nullptr) ? 1 : 0;
  const int rc = n2v_walk_consts(return_param, inout_param, graph->flags, has_ratio, &C);
  if (rc != N2V_OK) return rc;
  if (n_start == 0) return N2V_OK;
  N2V_CHECK_ARG(start && walks && alive, "n2v_walk: NULL buffer");
  N2V_CHECK_ARG((reinterpret_cast<uintptr_t>(walks) & 31) == 0, "n2v_walk: walks must be 32-byte aligned");
  const int64_t total = n_start * num_walks;
  N2V_CHECK_ARG(total < (int64_t(1) << 32), "n2v_walk: %lld walkers per call exceed 2^32; shard the start list",
                static_cast<long long>(total));

  WalkArgs A{};
  A.start = start;
  A.num_walks = num_walks;
  A.walk_length = walk_length;
  A.key0 = static_cast<uint32_t>(seed);
  A.key1 = static_cast<uint32_t>(seed >> 32);
  A.pitch = pitch;
  A.stats = reinterpret_cast<unsigned long long*>(stats);
  A.ret_m1 = static_cast<uint32_t>(C.t_ret - 1);
  A.nbr_m1 = static_cast<uint32_t>(C.t_nbr - 1);
  A.far_m1 = static_cast<uint32_t>(C.t_far - 1);
  A.lo_m1 = A.nbr_m1 < A.far_m1 ? A.nbr_m1 : A.far_m1;
  A.hi_m1 = A.nbr_m1 < A.far_m1 ? A.far_m1 : A.nbr_m1;
  A.fold_gain = C.fold_gain;
  A.mix_qm1 = C.mix_qm1;
  A.max_trials = C.max_trials;
  A.inv_p = 1.0 / return_param;
  A.inv_q = 1.0 / inout_param;

  const int fold = C.fold_mode;
  const bool multi = graph->n_parts > 1;
  const bool st = stats != nullptr;
  const int64_t chunk_max = int64_t(1) << 30;  // walkers per launch
  for (int64_t w0 = 0; w0 < total; w0 += chunk_max) {
    const int64_t n = total - w0 < chunk_max ? total - w0 : chunk_max;
    A.walker0 = static_cast<uint32_t>(w0);
    A.n_walkers = static_cast<uint32_t>(n);
    A.walks = walks + w0 * pitch;
    A.alive = alive + w0;
    const int64_t need = (n + kBlock - 1) / kBlock;
    const int64_t cap = int64_t(n2v::sm_count()) * kBlocksPerSm;
    const int grid = static_cast<int>(need < cap ? need : cap);
#define N2V_LAUNCH_WALK(F, M, S) walk_kernel<F, M, S><<<grid, kBlock, 0, stream>>>(*graph, A)
    switch (fold * 4 + (multi ? 2 : 0) + (st ? 1 : 0)) {
      case 0: N2V_LAUNCH_WALK(0, false, false); break;
      case 1: N2V_LAUNCH_WALK(0, false, true); break;
      case 2: N2V_LAUNCH_WALK(0, true, false); break;
      case 3: N2V_LAUNCH_WALK(0, true, true); break;
      case 4: N2V_LAUNCH_WALK(1, false, false); break;
      case 5: N2V_LAUNCH_WALK(1, false, true); break;
      case 6: N2V_LAUNCH_WALK(1, true, false); break;
      case 7: N2V_LAUNCH_WALK(1, true, true); break;
      case 8: N2V_LAUNCH_WALK(2, false, false); break;
      case 9: N2V_LAUNCH_WALK(2, false, true); break;    // general fold is single-part only
      case 12: N2V_LAUNCH_WALK(3, false, false); break;
      case 13: N2V_LAUNCH_WALK(3, false, true); break;
      case 14: N2V_LAUNCH_WALK(3, true, false); break;
      case 15: N2V_LAUNCH_WALK(3, true, true); break;
      default:
        n2v::set_error("n2v_walk: fold mode %d is not available on a %d-part graph", fold, graph->n_parts);
        return N2V_ERR_INVALID;
    }
#undef N2V_LAUNCH_WALK
    N2V_LAUNCH_OK();
  }
  return N2V_OK;
}
