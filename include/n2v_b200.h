/*
 * n2v_b200.h -- C ABI of libn2v_b200.so: the B200 (sm_100a) replacement for the hot
 * path of graph-embedding/node2vec 0.3.5 (node2vec-fugue).
 *
 * The reference is pure Python and has no FFI; the "interface each entry point
 * replaces" is therefore the reference FUNCTION whose work it takes over, cited as
 * file:line into the reference tree.  The Python host (node2vec_b200/) binds these
 * with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, PODs.  No torch / C++ types cross the boundary.
 *   - every pointer is a DEVICE pointer unless the name ends in _host.
 *   - the library owns no persistent memory: callers (torch tensors) own every buffer,
 *     scratch included; sizes come from the *_scratch_bytes() queries.
 *   - `stream` is a cudaStream_t passed as void* (torch's current stream); calls are
 *     asynchronous on it and never synchronise unless documented.
 *   - return value: 0 = OK, non-zero = error; n2v_last_error() gives the message for
 *     the calling thread.  No exception crosses the ABI.
 *   - re-entrant, no global state; one host thread per GPU/process is the intended use.
 */
#ifndef N2V_B200_H_
#define N2V_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define N2V_ABI_VERSION 7
#define N2V_MAX_PARTS 16

/* error codes */
enum {
  N2V_OK = 0,
  N2V_ERR_INVALID = 1,   /* bad argument (the Python host raises ValueError) */
  N2V_ERR_CUDA = 2,      /* CUDA runtime error; message has the cudaError string */
  N2V_ERR_SCRATCH = 3,   /* scratch buffer too small */
  N2V_ERR_ZERO_WEIGHT = 4 /* a vertex's out-weights sum to 0 (reference: ZeroDivisionError, randomwalk.py:172-173) */
};

/* how `sum(node_weights)` (randomwalk.py:172) is evaluated -- see DESIGN.md "sum modes" */
enum {
  N2V_SUM_NAIVE = 0,     /* left-to-right fp64: CPython <= 3.11, the reference's supported 3.6/3.7 */
  N2V_SUM_NEUMAIER = 1   /* CPython >= 3.12 builtin sum() float path */
};

/* graph property flags (set by n2v_csr_build, consumed by n2v_walk) */
enum {
  N2V_GRAPH_UNIT_WEIGHT = 1u,  /* every arc weight == 1.0 */
  N2V_GRAPH_SYMMETRIC = 2u,    /* arc (a,b,w) present  <=>  arc (b,a,w) present */
  N2V_GRAPH_SIMPLE = 4u        /* no repeated (src,dst) pair */
};

/* One vertex of the CSR: replaces one row of df_adj, i.e. one base64(pickle) adjacency
 * string (randomwalk.py:266-275, Neighbors :17-41).  16 B, loaded as one LDG.128.
 * Offsets are local to the vertex's part; a part holds fewer than 2^32 arcs. */
typedef struct n2v_vertex {
  uint32_t base;  /* index of the first out-arc in arcs[] / col[] / weight[] */
  uint32_t deg;   /* number of out-arcs */
  uint32_t hbase; /* first 32-byte bucket of this vertex's neighbour hash set: (base >> 2) + local vertex index */
  float wsum;     /* (float) of the fp64 left-to-right sum of out-weights (weighted return-edge fold) */
} n2v_vertex_t;

/* Neighbour hash set of a vertex: the membership test "x in N_out(t)" of
 * generate_edge_alias_tables (randomwalk.py:226, a Python set) as ONE 32-byte gather.
 * Vertex t owns n2v_hash_nbuckets(deg) consecutive buckets of 8 int32 slots (load <= 0.5)
 * starting at bucket n2v_hash_base(base, local index) -- derivable from what a walker
 * already holds, so no extra gather -- filled by linear probing over buckets in col[]
 * order; empty slots hold N2V_HASH_EMPTY and a bucket's used slots are contiguous from
 * slot 0.   bucket(x) = mulhi32(x * 0x9E3779B1, nbuckets) */
#define N2V_HASH_SLOTS 8
#define N2V_HASH_EMPTY (-1)
#define N2V_HASH_MULT 0x9E3779B1u
#ifdef __CUDACC__
#define N2V_HD __host__ __device__
#else
#define N2V_HD
#endif
static inline N2V_HD uint32_t n2v_hash_nbuckets(uint32_t deg) { return (deg + 3u) >> 2; }
/* floor(base/4) + v leaves every vertex at least ceil(deg/4) buckets: total <= n_arcs/4 + n_vertices */
static inline N2V_HD uint32_t n2v_hash_base(uint32_t base, uint32_t local_vertex) { return (base >> 2) + local_vertex; }

/* One out-arc with its FIRST-ORDER alias-table entry folded in: replaces the
 * (alias[k], probs[k]) pair of generate_alias_tables (randomwalk.py:157-190), the
 * neighbour id AND the adjacency header of whichever vertex the draw lands on (the
 * reference's next join, fugue.py:147), so that one walk trial is ONE 32-byte gather
 * (one DRAM/L2 sector, one LDG.256 on sm_100a) with no dependent vertex lookup.
 *   thr       = min(ceil(probs[k] * 2^32), 2^32 - 1): `u32 < thr`  <=>  `r2 < probs[k]`
 *   dst       = neighbour id at index k          (taken when u32 <  thr)
 *   alias_dst = neighbour id at index alias[k]   (taken when u32 >= thr); == dst when probs[k] >= 1
 *   alias_idx = alias[k] exactly as the reference computes it
 *   dst_base / dst_deg, adst_base / adst_deg = vtx[dst] / vtx[alias_dst] (part-local base) */
typedef struct n2v_arc {
  uint32_t thr;
  int32_t dst;
  int32_t alias_dst;
  int32_t alias_idx;
  uint32_t dst_base;
  uint32_t dst_deg;
  uint32_t adst_base;
  uint32_t adst_deg;
} n2v_arc_t;

/* One vertex-range shard of the CSR.  A replicated graph has exactly one part that
 * covers [0, n_vertices).  With a vertex-partitioned CSR the pointers of remote parts
 * are CUDA-IPC-mapped peer addresses (NVLink loads). */
typedef struct n2v_graph_part {
  const n2v_vertex_t* vtx;  /* [v_hi - v_lo], indexed by (v - v_lo); base is local to this part */
  const n2v_arc_t* arcs;    /* [part arcs] */
  const int32_t* col;       /* [part arcs] neighbour ids, ascending within a vertex */
  const double* weight;     /* [part arcs] fp64 weights in col order */
  const int32_t* hash;      /* [part buckets][8] neighbour hash sets */
  const float* ratio;       /* [part arcs][2] {fwd, rev} return-mass ratios (n2v_ratio_build) or NULL */
} n2v_graph_part_t;

typedef struct n2v_graph {
  int64_t n_vertices;   /* ids are 0 .. n_vertices-1; ids absent from the arc list have deg 0 */
  int64_t n_arcs;       /* total over all parts */
  uint32_t flags;       /* N2V_GRAPH_* */
  int32_t n_parts;      /* 1 = replicated */
  int64_t part_size;    /* vertices per part (last may be short); part(v) = v / part_size */
  n2v_graph_part_t parts[N2V_MAX_PARTS];
} n2v_graph_t;

/* Counters returned by n2v_walk (device memory, 8 x uint64, accumulated with atomics;
 * the caller zeroes them).  They feed the roofline's algorithmic-bytes formula. */
typedef struct n2v_walk_stats {
  uint64_t steps;        /* walker-steps taken */
  uint64_t trials;       /* alias proposals drawn (>= steps; == steps when p == q == 1) */
  uint64_t probes;       /* 32-byte hash buckets read for membership tests (>= searches) */
  uint64_t searches;     /* membership tests "x in N_out(prev)" the accept draw could not skip */
  uint64_t fold_hits;    /* steps resolved by the return-edge fold without a proposal */
  uint64_t fallbacks;    /* steps resolved by the exact O(deg) scan after N2V_MAX_TRIALS rejections */
  uint64_t dead;         /* walkers dropped at a vertex with no out-arcs (fugue.py:147 inner join) */
  uint64_t reserved;
} n2v_walk_stats_t;

/* ---- housekeeping ----------------------------------------------------------------- */
int n2v_abi_version(void);
const char* n2v_last_error(void);

/* ---- K0: arcs -> sorted CSR -------------------------------------------------------
 * Replaces `partition(by=["src"], presort="dst")` + get_vertex_neighbors
 * (fugue.py:130, randomwalk.py:266-275).
 * src: [n_arcs] int32 ids in [0, n_vertices) (part-LOCAL ids when this is one part of a
 * vertex-partitioned CSR); dst: ids in [0, n_dst_vertices) (always global ids;
 * n_dst_vertices == n_vertices for a replicated graph); weight: [n_arcs] fp64 or NULL (=1.0).
 * The SYMMETRIC flag is only computed for a replicated graph (mirrors live in other parts).
 * Out: vtx[n_vertices] (base, deg; wsum filled by n2v_alias_build), col[n_arcs],
 * weight_sorted[n_arcs], perm[n_arcs] (input position of each sorted arc; may be NULL),
 * *flags_host (N2V_GRAPH_*; this call synchronises the stream to return it).
 * Arcs are ordered by (src, dst); equal pairs keep input order (stable). */
size_t n2v_csr_scratch_bytes(int64_t n_arcs, int64_t n_vertices);
int n2v_csr_build(const int32_t* src, const int32_t* dst, const double* weight, int64_t n_arcs,
                  int64_t n_vertices, int64_t n_dst_vertices, n2v_vertex_t* vtx, int32_t* col,
                  double* weight_sorted, int64_t* perm, void* scratch, size_t scratch_bytes,
                  uint32_t* flags_host, void* stream);

/* ---- K0b: neighbour hash sets -----------------------------------------------------
 * Replaces `set(Neighbors(src_nbs).dst_id)` (randomwalk.py:318), built once per vertex
 * instead of once per walker per step.  hash: [n_buckets_cap][8] int32 with
 * n_buckets_cap >= n2v_hash_buckets_bound() = n_arcs/4 + n_vertices + 1; fills vtx[].hbase.
 * Asynchronous. */
int64_t n2v_hash_buckets_bound(int64_t n_arcs, int64_t n_vertices);
int n2v_hash_build(n2v_vertex_t* vtx, const int32_t* col, int64_t n_vertices, int64_t n_arcs,
                   int32_t* hash, int64_t n_buckets_cap, void* stream);

/* ---- K1: per-vertex first-order alias tables, bit-exact fp64 ------------------------
 * Replaces generate_alias_tables (randomwalk.py:157-190) for every vertex at once (the
 * reference re-runs it per walker per step, :320-321).
 * Out: alias[n_arcs] int32 and probs[n_arcs] fp64 exactly as the reference returns them
 * per vertex (alias may be NULL), arcs[n_arcs] packed records, vtx[].wsum.
 * vtx_lookup: headers indexed by GLOBAL vertex id, read for the landing-vertex fields of the
 * arc records; NULL = vtx itself (replicated graph).  In a vertex-partitioned CSR it is the
 * all-gathered header array.
 * scratch: n_arcs int32 (the two LIFO work-lists share each vertex's slice).
 * *n_zero_host: number of vertices with deg > 0 whose weights sum to 0 (their tables are
 * left zeroed; the host raises).  Synchronises the stream to return it. */
int n2v_alias_build(n2v_vertex_t* vtx, const n2v_vertex_t* vtx_lookup, const int32_t* col,
                    const double* weight_sorted, int64_t n_vertices, int64_t n_arcs, int sum_mode,
                    int32_t* alias, double* probs, n2v_arc_t* arcs, int32_t* scratch,
                    int64_t* n_zero_host, void* stream);

/* ---- a4: second-order (p,q) alias tables for explicit (prev, cur) pairs -------------
 * Replaces generate_edge_alias_tables (randomwalk.py:193-232).  The walk kernel never
 * materialises these (that is the point); this entry point exists so the bit-exact
 * parity of the biasing + table construction can be shown on the device.
 * prev/cur: [n_pairs]; prev < 0 means "first step" (unbiased, randomwalk.py:320-321).
 * Table i is written at out_offset[i] .. out_offset[i] + deg(cur[i]) in alias_out /
 * probs_out; scratch has the same layout (int32).  Single-part graphs only. */
int n2v_edge_alias_build(const n2v_graph_t* graph, const int32_t* prev, const int32_t* cur,
                         int64_t n_pairs, double return_param, double inout_param, int sum_mode,
                         const int64_t* out_offset, int32_t* alias_out, double* probs_out,
                         int32_t* scratch, void* stream);

/* ---- a5/a6: the two alias samplers on explicit fp64 uniforms ------------------------
 * Replaces AliasProb.sampling_from_alias (randomwalk.py:86-99; second != NULL) and
 * sampling_from_alias_wiki (:70-84; second == NULL).  Table i is
 * alias[offset[i] .. offset[i+1]).  Out: picked index per draw. */
int n2v_alias_draw(const int32_t* alias, const double* probs, const int64_t* offset,
                   const double* first, const double* second, int64_t n_draws, int32_t* picked,
                   void* stream);

/* ---- K2: second-order biased random walks ------------------------------------------
 * Replaces initiate_random_walk + [two joins + next_step_random_walk] x walk_length +
 * to_path (randomwalk.py:279-349, fugue.py:137-153).
 * Walker w = s * num_walks + r  (s indexes start[], r = 0..num_walks-1) writes row w of
 * walks[n_start*num_walks][pitch] (int32; pitch >= walk_length+1, pitch % 8 == 0):
 * walk_length+1 vertex ids, entries past the end of a dropped walk are -1.
 * alive[w] = 1 if the walker took all walk_length steps, 0 if it was dropped at a vertex
 * without out-arcs (the reference's inner join drops the whole row, fugue.py:147).
 * Randomness: Philox4x32-10, key = seed, counter = (walk_id lo, walk_id hi, step, trial)
 * with walk_id = start[s] * num_walks + r, so results do not depend on sharding.
 * stats: device n2v_walk_stats_t, accumulated (caller zeroes); may be NULL. */
int n2v_walk(const n2v_graph_t* graph, const int32_t* start, int64_t n_start, int32_t num_walks,
             int32_t walk_length, double return_param, double inout_param, uint64_t seed,
             int32_t* walks, int64_t pitch, uint8_t* alive, n2v_walk_stats_t* stats,
             void* stream);

/* The sampling constants n2v_walk derives from (p, q, graph flags); exposed so the host
 * replay in oracle/ and the roofline calculator use exactly the same numbers.
 *
 * Mode 3 (mixture sampler).  On a unit-weight symmetric simple graph with q > 1 the reference's
 * law at (prev t, cur v) (randomwalk.py:223-230), divided by 1/q, is
 *     alpha(x) * q = 1  +  (q - 1) * [x in N(t) and N(v)]  +  (q/p - 1) * [x == t]        (x in N(v))
 * i.e. a mixture of  (a) "bulk": x uniform on N(v), mass deg(v), always accepted (x == t only
 * with probability min(1, q/p));  (b) "common neighbours": mass |N(t) & N(v)| * (q - 1), drawn by
 * proposing x uniformly from the SMALLER of N(t), N(v) under the envelope mass
 * min(deg t, deg v) * (q - 1) and accepting iff x is in the other set (one hash-bucket gather)
 * and x != t;  (c) "return": mass max(0, q/p - 1), x = t.  One Philox draw picks the component
 * with the fp32 thresholds below; a failed (b) proposal rejects the whole trial.  Proposals are
 * the same 32-byte arc-record gathers as in the other modes, so an accepted x arrives with its
 * adjacency header.  Compared with proposing from N(v) only this needs fewer trials whenever
 * deg(t) < deg(v) and no membership test at all for bulk trials.
 *     fo = fmul(float(min(dv, dt)), mix_qm1); tot = fadd(fadd(float(dv), fo), fold_gain);
 *     thr_ret = rz(fmul(fdiv(fold_gain, tot), 2^32));
 *     thr_out = sat(rz(fmul(fadd(fdiv(fold_gain, tot), fdiv(fo, tot)), 2^32)))      (all fp32, rn)
 *     u32 < thr_ret: return;  u32 < thr_out: common-neighbour proposal;  else bulk. */
typedef struct n2v_walk_consts {
  uint64_t t_ret;   /* accept x == prev      iff u32 < t_ret   (values in [0, 2^32]) */
  uint64_t t_nbr;   /* accept x in N_out(prev) iff u32 < t_nbr */
  uint64_t t_far;   /* accept otherwise      iff u32 < t_far */
  float fold_gain;  /* modes 1/2: e' = max(0, 1/p - cap) / cap; mode 3: max(0, q/p - 1); 0 = return fold off */
  int32_t fold_mode;/* 0 off, 1 unit-weight symmetric simple graph (rho = 1/deg), 2 general (per-arc ratios),
                     * 3 "mixture" (unit-weight symmetric simple graph AND q > 1), see below */
  int32_t max_trials;
  float mix_qm1;    /* mode 3: q - 1 */
} n2v_walk_consts_t;
int n2v_walk_consts(double return_param, double inout_param, uint32_t graph_flags, int has_ratio,
                    n2v_walk_consts_t* out_host);

/* ---- return-mass ratios for the general (weighted / directed / multi-arc) fold ----------
 * For arc e = (v -> x):  fwd = tot(v->x) / wsum(v),  rev = tot(x->v) / wsum(x)  (0 when x has no
 * arc back to v), tot = left-to-right fp64 sum over the parallel arcs, cast to fp32 and divided
 * by the fp32 vtx[].wsum.  The walk carries {fwd, rev} of the arc it took, so the probability of
 * returning through the fold, e*rev / (1 + e*rev), needs one 8-byte gather per step and no
 * search.  Single-part graphs; ratio_out: [n_arcs][2] fp32.  Run after n2v_alias_build. */
int n2v_ratio_build(const n2v_graph_t* graph, float* ratio_out, void* stream);

/* ---- device tuning: L2 fetch granularity ------------------------------------------------
 * The walk is a stream of random 32-byte gathers (one arc record / one hash bucket each); ncu on
 * the RMAT-20 walk shows 3.4 DRAM sectors read per requested sector (profiles/r02_walk_ncu.txt).
 * cudaLimitMaxL2FetchGranularity (a context-wide performance hint) is exposed for tuning runs;
 * measured on B200 it changes nothing (42.6 G random sectors/s at 32, 64 = default and 128,
 * profiles/r02_gather_granule_l2fetch.txt), so nothing in the package sets it by default.
 * bytes: 32, 64 or 128.  get returns the current limit (or -1 on error). */
int n2v_set_l2_fetch_granularity(int bytes);
int n2v_get_l2_fetch_granularity(void);

/* ---- K6: first-occurrence positions (graph indexer: name -> id remap, undirected dedupe) ----
 * Replaces the two order-sensitive pandas steps of index_graph_pandas (indexer.py:9-49):
 *   vertex id of a name = POSITION of its first occurrence in the concatenation
 *   [all src..., all dst...] (:26-35, append + drop_duplicates + reset_index: ids are sparse,
 *   up to 2E - 1), and the undirected expansion keeps the first occurrence of every
 *   (src, dst, weight) triple in frame order (:45-48).
 * first_out[i] = min { j : key[j] == key[i] } for rows keyed by 1-3 int64 columns (key1 / key2
 * may be NULL; a weight column is passed as its IEEE bit pattern).  No sort: an open-addressing
 * table of row positions, `n_slots` = n2v_first_occurrence_slots(n_rows) u32 entries of caller
 * scratch.  n_rows < 2^32 - 1.  Row i is a first occurrence iff first_out[i] == i. */
int64_t n2v_first_occurrence_slots(int64_t n_rows);
int n2v_first_occurrence(const int64_t* key0, const int64_t* key1, const int64_t* key2, int64_t n_rows,
                         uint32_t* table, int64_t n_slots, int64_t* first_out, void* stream);
/* The same for STRING vertex names in Arrow layout (row i = data[offsets[i] .. offsets[i+1]), UTF-8 bytes,
 * int64 offsets as in pyarrow's large_string): keys are compared byte for byte, so the name column of a
 * frame goes to the device as it is -- no host-side factorisation of the 2E names. */
int n2v_first_occurrence_bytes(const int64_t* offsets, const uint8_t* data, int64_t n_rows, uint32_t* table,
                               int64_t n_slots, int64_t* first_out, void* stream);

/* ---- K5: hotspot trimming, bit-exact with the reference's sampler ------------------------
 * Replaces trim_hotspot_vertices (randomwalk.py:238-262): a vertex with more than `cap` out-arcs
 * keeps `DataFrame.sample(n=cap, random_state=seed)` of them, i.e. numpy's legacy
 * RandomState(seed).permutation(deg)[:cap] (Fisher-Yates from the top index down, swap partners by
 * masked rejection from 32-bit MT19937 outputs; every partition re-seeds, so the kept positions
 * depend on (seed, deg) only).
 * deg[n_hot]: out-degrees of the hot vertices (each > cap is the intended use; any deg >= 1 works);
 * scratch_offset[n_hot]: start of vertex h's slice in scratch (exclusive prefix sum of deg);
 * scratch: int32[sum(deg)];  picked[n_hot][cap]: picked[h][k] = position (0-based, in the vertex's
 * input order) of the k-th kept arc, in the order pandas returns them.  Asynchronous. */
int n2v_trim_sample(const int64_t* deg, const int64_t* scratch_offset, int64_t n_hot, int32_t cap,
                    uint32_t seed, int32_t* scratch, int32_t* picked, void* stream);


/* ---- peer-shareable device buffers (vertex-partitioned CSR over NVLink) ----------------
 * The one place the library allocates.  A rank allocates its part's arcs / hash / col / weight
 * with n2v_ipc_alloc (CUDA virtual-memory-management allocation, 2 MiB pages, mapped read/write
 * on the current device), exports a 64-byte handle per buffer
 *     { int32 fd, int32 owner device, uint64 mapped size, zeros }
 * whose POSIX file descriptor the HOST transfers to the peer processes (SCM_RIGHTS); a peer
 * rewrites the fd field with its own copy and calls n2v_ipc_open, which maps the allocation
 * with its native page size on the caller's current device (NVLink peer loads), and puts the
 * address into n2v_graph_t.parts[owner].  (cudaIpc* mappings were measured ~50x slower for
 * random gathers beyond ~1 GB of remote footprint.) */
#define N2V_IPC_HANDLE_BYTES 64
int n2v_ipc_alloc(size_t bytes, void** ptr_host);
int n2v_ipc_free(void* ptr);
int n2v_ipc_export(const void* ptr, unsigned char* handle_host /* [64] */);
int n2v_ipc_open(const unsigned char* handle_host /* [64] */, void** ptr_host);
int n2v_ipc_close(void* ptr);

/* ================================ SGNS half ========================================
 * Replaces gensim.models.Word2Vec(sentences=all_walks, sg=1, negative=K, ...) as called by
 * Node2VecGensim.fit (embedding.py:120-127): vocabulary statistics, sub-sampling
 * thresholds, negative-sampling table, weight init and the skip-gram negative-sampling
 * SGD itself.  Tokens are vertex ids and index the embedding tables directly (row v of
 * syn0 / syn1neg belongs to vertex v); ids whose count is below min_count never occur in a
 * training sentence, exactly like gensim's out-of-vocabulary words. */

/* K4a: per-id token count and first flat position (row * len + column) over a walk matrix
 * walks[n_walks][pitch] (first `len` columns valid, negative entries ignored).
 * counts / first_pos: [n_vertices] int64, ACCUMULATED / MIN-ed into (caller initialises:
 * counts = 0, first_pos = INT64_MAX); pos_offset is added to positions (multi-shard). */
int n2v_vocab_count(const int32_t* walks, int64_t n_walks, int32_t len, int64_t pitch,
                    int64_t n_vertices, int64_t pos_offset, int64_t* counts, int64_t* first_pos,
                    void* stream);

/* K4b: gensim prepare_vocab / make_cum_table from counts.
 *   keep_thr[v]  : token v survives sub-sampling iff u32 < keep_thr[v] or keep_thr[v] == 2^32-1;
 *                  0 for ids with count < min_count (dropped from every sentence)
 *   neg_table    : alias table(s) over ids, P(v) ~ count^ns_exponent (0 for dropped ids), entries
 *                  {thr u32, alias i32}, n_vertices + n2v_neg_top_entries(n_vertices) of them.
 *                  n_vertices <= 65536: ONE table, one 8-byte gather per negative.  Larger: ids
 *                  are cut into chunks of 8192, each with its own alias table (built in parallel,
 *                  one thread per chunk), plus a top-level alias table over the chunk masses
 *                  stored after them: chunk ~ mass, then id within the chunk -- the same law
 *                  exactly, two gathers (the top level stays L2-resident).
 * scratch: (n_vertices + n2v_neg_top_entries(n_vertices)) * 20 bytes.
 * totals_host[0] = retained token total, [1] = retained ids.  Synchronises the stream. */
int n2v_sgns_prepare(const int64_t* counts, int64_t n_vertices, int64_t min_count, double sample,
                     double ns_exponent, uint32_t* keep_thr, int32_t* neg_table, void* scratch,
                     int64_t* totals_host, void* stream);

#define N2V_NEG_CHUNK 8192
static inline N2V_HD int64_t n2v_neg_top_entries(int64_t n_vertices) {
  return n_vertices > 65536 ? (n_vertices + N2V_NEG_CHUNK - 1) / N2V_NEG_CHUNK : 0;
}

/* weight init: syn0[v][d] = (U[0,1) - 0.5) / dim from Philox(seed; v, d/4) -- gensim's
 * reset_weights law; syn1neg is the caller's zero-filled buffer. */
int n2v_sgns_init(float* syn0, int64_t n_vertices, int32_t dim, uint64_t seed, void* stream);

/* the 1000-entry sigmoid table of word2vec (EXP_TABLE), host pointer out[1000] */
int n2v_sgns_exp_table(float* out_host);

typedef struct n2v_sgns_params {
  int32_t dim;          /* embedding width; multiple of 4, <= 1024 */
  int32_t window;       /* gensim `window` */
  int32_t negative;     /* gensim `negative` (K >= 1) */
  int32_t epochs;       /* gensim `iter`: total epochs of the schedule */
  int32_t epoch;        /* epoch this call runs (0-based) */
  int32_t batch_words;  /* gensim `batch_words`: alpha is refreshed per job of this many words */
  int32_t atomic_updates; /* 1 = red.global.add row updates, 0 = plain Hogwild stores like gensim */
  int32_t reserved;
  float alpha;          /* gensim `alpha` */
  float min_alpha;      /* gensim `min_alpha` */
  uint64_t seed;
  int64_t walk_offset;  /* global index of this shard's first walk (schedule + RNG streams) */
  int64_t total_walks;  /* walks per epoch over all shards */
} n2v_sgns_params_t;

/* K3: one epoch of skip-gram negative sampling over walks[n_walks][pitch] (`len` tokens each).
 * One warp per walk: sub-sample, then for every centre i draw the reduced window and for every
 * context j update (syn0[walk[j]], syn1neg[walk[i]], syn1neg[K negatives]).
 * stats (device, 4 x uint64, accumulated): [0] pairs trained, [1] tokens kept; and, counted only
 * in trace mode (diagnostics), [2] negatives skipped (== centre), [3] targets clipped (|f| >= 6).
 * trace (device int32[trace_cap][2 + negative], may be NULL): when given, ONE warp runs the
 * walks in order and records every pair {centre, context, negatives (-1 = skipped)} and its
 * alpha in trace_alpha; pairs beyond trace_cap are trained but not recorded. */
int n2v_sgns_train(const int32_t* walks, int64_t n_walks, int32_t len, int64_t pitch,
                   const uint32_t* keep_thr, const int32_t* neg_table, int64_t n_vertices,
                   float* syn0, float* syn1neg, const float* exp_table, const n2v_sgns_params_t* params,
                   uint64_t* stats, int32_t* trace, float* trace_alpha, int64_t trace_cap,
                   void* stream);

/* K3s (EXPERIMENTAL, opt-in; NOT the parity path): the same epoch with WINDOW-SHARED negatives --
 * the K negatives are drawn once per centre and shared by the pairs of its window instead of
 * re-drawn per pair as gensim does (a duplicate draw inside one K-set is dropped).  The centre row
 * and the K negative rows stay in registers across the window and are reduced to memory once per
 * centre: ~4.2 row operations per pair instead of ~12.4.  Same arguments as n2v_sgns_train;
 * requires params->dim <= 128, params->negative == 5, params->atomic_updates != 0.  The trace lists
 * the centre's K-set for every pair (-1 = skipped or duplicate). */
int n2v_sgns_train_shared(const int32_t* walks, int64_t n_walks, int32_t len, int64_t pitch,
                          const uint32_t* keep_thr, const int32_t* neg_table, int64_t n_vertices,
                          float* syn0, float* syn1neg, const float* exp_table, const n2v_sgns_params_t* params,
                          uint64_t* stats, int32_t* trace, float* trace_alpha, int64_t trace_cap,
                          void* stream);

/* x[i] *= factor (the 1/G of model averaging after an NCCL sum-allreduce) */
int n2v_scale(float* x, int64_t n, float factor, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* N2V_B200_H_ */
