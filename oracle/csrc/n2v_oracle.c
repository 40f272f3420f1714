/*
 * n2v_oracle.c -- CPU oracle for the node2vec walk path.  TEST INFRASTRUCTURE, NOT PRODUCT:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  node2vec_b200 never does.
 *
 * Part A restates the REFERENCE's algorithm (graph-embedding/node2vec 0.3.5) in C so
 * that large cases finish in seconds: alias tables (randomwalk.py:157-190), (p,q) edge
 * biasing (:193-232), the two-uniform sampler (:86-99), the per-row step (:300-339) and
 * the step loop with its drop-at-sink join (fugue.py:146-150), driven by CPython's own
 * Mersenne Twister so seeded runs reproduce the reference's walks exactly.  It is pinned
 * to the reference through the fixtures under tests/golden (tests/test_oracle_c.py).
 *
 * Part B is the host replay of the DEVICE sampler (node2vec_b200/csrc/walk.cu): same
 * Philox4x32-10 stream, same integer decisions, written independently of the CUDA
 * source.  A GPU walk must equal its replay bit for bit.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, so fp64
 * and fp32 round exactly as CPython / the device intrinsics do).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_SUM_NAIVE 0
#define ORC_SUM_NEUMAIER 1

/* ================================ Part A: the reference ============================ */

/* builtins.sum over floats starting from int 0 (randomwalk.py:172) */
static double py_float_sum(const double* v, int64_t n, int mode) {
  if (n == 0) return 0.0;
  double total = v[0];
  if (mode == ORC_SUM_NAIVE) {
    for (int64_t i = 1; i < n; ++i) total = total + v[i];
    return total;
  }
  double comp = 0.0; /* CPython >= 3.12: Neumaier */
  for (int64_t i = 1; i < n; ++i) {
    const double x = v[i];
    const double t = total + x;
    if (fabs(total) >= fabs(x)) comp += (total - t) + x;
    else comp += (x - t) + total;
    total = t;
  }
  if (comp != 0.0 && isfinite(comp)) total += comp;
  return total;
}

/* generate_alias_tables (randomwalk.py:157-190).  probs[] = weights on entry.
 * work[] : n ints.  Returns 0, or -1 when the mean is zero (ZeroDivisionError). */
static int alias_tables_inplace(double* probs, int32_t* alias, int64_t n, int mode, int32_t* work) {
  if (n == 0) return -1;
  const double mean = py_float_sum(probs, n, mode) / (double)n;
  if (mean == 0.0) return -1;
  int64_t ns = 0, nl = 0; /* small list grows up from work[0], large list down from work[n-1] */
  for (int64_t i = 0; i < n; ++i) {
    probs[i] = probs[i] / mean;
    alias[i] = 0;
    if (probs[i] < 1.0) work[ns++] = (int32_t)i;
    else work[n - 1 - nl++] = (int32_t)i;
  }
  while (ns > 0 && nl > 0) {
    const int32_t lo = work[--ns];
    const int32_t hi = work[n - 1 - (--nl)];
    alias[lo] = hi;
    probs[hi] = probs[hi] + probs[lo] - 1.0;
    if (probs[hi] < 1.0) work[ns++] = hi;
    else work[n - 1 - nl++] = hi;
  }
  return 0;
}

int orc_alias_tables(const double* weights, int64_t n, int mode, int32_t* alias, double* probs) {
  int32_t* work = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  memcpy(probs, weights, sizeof(double) * (size_t)n);
  const int rc = alias_tables_inplace(probs, alias, n, mode, work);
  free(work);
  return rc;
}

/* alias tables for vertices [v_lo, v_hi) of a CSR (row_ptr[V+1]); returns #vertices with
 * zero mean.  Threading: callers run disjoint ranges from several host threads (ctypes
 * releases the GIL); there is no OpenMP runtime in this image. */
int64_t orc_alias_tables_csr(const int64_t* row_ptr, const double* weights, int64_t v_lo, int64_t v_hi, int mode,
                             int32_t* alias, double* probs) {
  int64_t bad = 0;
  {
    int32_t* work = NULL;
    int64_t cap = 0;
    for (int64_t v = v_lo; v < v_hi; ++v) {
      const int64_t b = row_ptr[v], n = row_ptr[v + 1] - b;
      if (n == 0) continue;
      if (n > cap) {
        free(work);
        cap = n * 2;
        work = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
      }
      memcpy(probs + b, weights + b, sizeof(double) * (size_t)n);
      if (alias_tables_inplace(probs + b, alias + b, n, mode, work) != 0) {
        ++bad;
        for (int64_t i = 0; i < n; ++i) { probs[b + i] = 0.0; alias[b + i] = 0; }
      }
    }
    free(work);
  }
  return bad;
}

static int sorted_contains(const int32_t* a, int64_t n, int32_t x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo < n && a[lo] == x;
}

/* the biased weights of generate_edge_alias_tables (randomwalk.py:219-231); prev < 0 = first step */
static void biased_weights(const int64_t* row_ptr, const int32_t* col, const double* w, int32_t prev,
                           int32_t cur, double p, double q, double* out) {
  const int64_t b = row_ptr[cur], n = row_ptr[cur + 1] - b;
  if (prev < 0) {
    memcpy(out, w + b, sizeof(double) * (size_t)n);
    return;
  }
  const int32_t* pc = col + row_ptr[prev];
  const int64_t pn = row_ptr[prev + 1] - row_ptr[prev];
  for (int64_t i = 0; i < n; ++i) {
    const int32_t x = col[b + i];
    if (x == prev) out[i] = w[b + i] / p;
    else if (sorted_contains(pc, pn, x)) out[i] = w[b + i];
    else out[i] = w[b + i] / q;
  }
}

int orc_edge_alias_tables(const int64_t* row_ptr, const int32_t* col, const double* w, int32_t prev,
                          int32_t cur, double p, double q, int mode, int32_t* alias, double* probs) {
  if (p == 0.0 || q == 0.0) return -2;
  const int64_t n = row_ptr[cur + 1] - row_ptr[cur];
  biased_weights(row_ptr, col, w, prev, cur, p, q, probs);
  int32_t* work = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  const int rc = alias_tables_inplace(probs, alias, n, mode, work);
  free(work);
  return rc;
}

/* ---- CPython's random module: MT19937, seed(int) = init_by_array, random() = 53 bits ---- */
typedef struct { uint32_t mt[624]; int idx; } orc_mt;

static void mt_init_genrand(orc_mt* s, uint32_t seed) {
  s->mt[0] = seed;
  for (int i = 1; i < 624; ++i) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->idx = 624;
}

static void mt_init_by_array(orc_mt* s, const uint32_t* key, int len) {
  mt_init_genrand(s, 19650218u);
  int i = 1, j = 0;
  for (int k = (624 > len ? 624 : len); k; --k) {
    s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    ++i; ++j;
    if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
    if (j >= len) j = 0;
  }
  for (int k = 623; k; --k) {
    s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    ++i;
    if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
  }
  s->mt[0] = 0x80000000u;
  s->idx = 624;
}

static void mt_seed_u64(orc_mt* s, uint64_t seed) { /* random.seed(non-negative int) */
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  mt_init_by_array(s, key, key[1] ? 2 : 1);
}

static uint32_t mt_next(orc_mt* s) {
  if (s->idx >= 624) {
    uint32_t* mt = s->mt;
    for (int k = 0; k < 624; ++k) {
      const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
      mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s->idx = 0;
  }
  uint32_t y = s->mt[s->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

static double mt_random(orc_mt* s) {
  const uint32_t a = mt_next(s) >> 5, b = mt_next(s) >> 6;
  return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

void orc_mt_random(uint64_t seed, int64_t n, double* out) {
  orc_mt s;
  mt_seed_u64(&s, seed);
  for (int64_t i = 0; i < n; ++i) out[i] = mt_random(&s);
}

/* one row of next_step_random_walk (randomwalk.py:316-339): build the table, draw with
 * the two-uniform sampler.  buf_p / buf_a / buf_w: deg(cur)-sized scratch. */
static int32_t reference_step(const int64_t* row_ptr, const int32_t* col, const double* w, int32_t prev,
                              int32_t cur, double p, double q, int mode, double r1, double r2,
                              double* buf_p, int32_t* buf_a, int32_t* buf_w) {
  const int64_t b = row_ptr[cur], n = row_ptr[cur + 1] - b;
  biased_weights(row_ptr, col, w, prev, cur, p, q, buf_p);
  alias_tables_inplace(buf_p, buf_a, n, mode, buf_w);
  const int64_t k = (int64_t)(r1 * (double)n);
  const int64_t pick = (r2 < buf_p[k]) ? k : buf_a[k];
  return col[b + pick];
}

/*
 * fugue.random_walk (fugue.py:119-155) over a CSR (row_ptr, col sorted, w):
 *   walkers = start[] x (1..num_walks) in that order; each of walk_length steps first
 *   drops walkers standing on a vertex without out-arcs, then steps the rest in row
 *   order drawing (r1, r2) per row.  has_seed: re-seed the Mersenne Twister at every step
 *   (randomwalk.py:314-315); otherwise one stream seeded from `seed`.
 * walks: [n_start*num_walks][walk_length+1]; alive[w] says whether row w survived.
 * Only walkers [w_lo, w_hi) are processed (one "partition"): for timing, several host
 * threads each run their own range with their own stream, which is the reference's
 * behaviour on more than one partition.  Parity runs use the full range.
 * Returns the number of surviving walkers in the range.
 */
int64_t orc_reference_walk(const int64_t* row_ptr, const int32_t* col, const double* w, int64_t max_deg,
                           const int32_t* start, int64_t n_start, int32_t num_walks, int32_t walk_length,
                           double p, double q, int mode, int has_seed, uint64_t seed, int64_t w_lo, int64_t w_hi,
                           int32_t* walks, uint8_t* alive) {
  const int64_t W = n_start * num_walks;
  const int64_t pitch = walk_length + 1;
  if (w_lo < 0) w_lo = 0;
  if (w_hi > W) w_hi = W;
  for (int64_t i = w_lo; i < w_hi; ++i) {
    walks[i * pitch] = start[i / num_walks];
    for (int64_t j = 1; j < pitch; ++j) walks[i * pitch + j] = -1;
    alive[i] = 1;
  }
  {
    const int64_t lo = w_lo, hi = w_hi;
    double* buf_p = (double*)malloc(sizeof(double) * (size_t)(max_deg + 1));
    int32_t* buf_a = (int32_t*)malloc(sizeof(int32_t) * (size_t)(max_deg + 1));
    int32_t* buf_w = (int32_t*)malloc(sizeof(int32_t) * (size_t)(max_deg + 1));
    orc_mt rng;
    mt_seed_u64(&rng, seed);
    for (int32_t s = 0; s < walk_length; ++s) {
      if (has_seed) mt_seed_u64(&rng, seed);
      for (int64_t i = lo; i < hi; ++i) {
        if (!alive[i]) continue;
        const int32_t cur = walks[i * pitch + s];
        if (row_ptr[cur + 1] == row_ptr[cur]) { alive[i] = 0; continue; } /* inner join drops it */
        const int32_t prev = s == 0 ? -1 : walks[i * pitch + s - 1];
        const double r1 = mt_random(&rng), r2 = mt_random(&rng);
        walks[i * pitch + s + 1] = reference_step(row_ptr, col, w, prev, cur, p, q, mode, r1, r2, buf_p, buf_a, buf_w);
      }
    }
    free(buf_p); free(buf_a); free(buf_w);
  }
  int64_t n_alive = 0;
  for (int64_t i = w_lo; i < w_hi; ++i) n_alive += alive[i];
  return n_alive;
}

/* ============================ Part B: replay of the device sampler ================== */

static void philox4x32_10(uint32_t k0, uint32_t k1, const uint32_t ctr[4], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t key[2], const uint32_t ctr[4], uint32_t out[4]) {
  philox4x32_10(key[0], key[1], ctr, out);
}

typedef struct orc_walk_consts {
  uint64_t t_ret, t_nbr, t_far; /* accept iff u32 < t, t in [1, 2^32] */
  float fold_gain;
  int32_t fold_mode;
  int32_t max_trials;
  float mix_qm1; /* mode 3: q - 1 */
} orc_walk_consts;

static uint64_t accept_thr(double a) {
  if (a >= 1.0) return 4294967296ull;
  double s = floor(a * 4294967296.0 + 0.5);
  if (s < 1.0) s = 1.0;
  if (s > 4294967296.0) s = 4294967296.0;
  return (uint64_t)s;
}

/* envelope and fold constants: rejection sampling of w(v,x)*alpha(t,x) from the
 * first-order table; graph_flags bit0 unit weights, bit1 symmetric, bit2 simple */
int orc_walk_consts_for(double p, double q, uint32_t graph_flags, int has_ratio, orc_walk_consts* c) {
  if (!(p > 0.0) || !(q > 0.0) || !isfinite(p) || !isfinite(q)) return -1;
  const double ip = 1.0 / p, iq = 1.0 / q;
  double cap = iq > 1.0 ? iq : 1.0;
  memset(c, 0, sizeof(*c));
  if ((graph_flags & 7u) == 7u && q > 1.0) { /* mixture sampler (include/n2v_b200.h, mode 3) */
    const double g = q / p - 1.0;
    const double a_ret = q / p;
    c->fold_mode = 3;
    c->fold_gain = (float)(g > 0.0 ? g : 0.0);
    c->mix_qm1 = (float)(q - 1.0);
    c->t_ret = accept_thr(a_ret < 1.0 ? a_ret : 1.0);
    c->t_nbr = c->t_far = 4294967296ull;
    c->max_trials = 256;
    return 0;
  }
  if (ip > cap) {
    if ((graph_flags & 7u) == 7u) { c->fold_mode = 1; c->fold_gain = (float)((ip - cap) / cap); }
    else if (has_ratio) { c->fold_mode = 2; c->fold_gain = (float)((ip - cap) / cap); }
    else cap = ip;
  }
  c->t_ret = accept_thr((ip < cap ? ip : cap) / cap);
  c->t_nbr = accept_thr(1.0 / cap);
  c->t_far = accept_thr(iq / cap);
  c->max_trials = 256;
  return 0;
}

static int member_probe(const int32_t* col, uint32_t n, int32_t x, uint64_t* probes) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    ++*probes;
    if (col[mid] < x) lo = mid + 1; else hi = mid;
  }
  if (lo >= n) return 0;
  ++*probes;
  return col[lo] == x;
}

static uint32_t exact_draw(const int32_t* vcol, const double* vw, uint32_t deg, int32_t t, const int32_t* tcol,
                           uint32_t tdeg, double inv_p, double inv_q, uint32_t r0, uint32_t r1, uint64_t* probes) {
  double total = 0.0;
  for (uint32_t i = 0; i < deg; ++i) {
    const int32_t x = vcol[i];
    const double a = (x == t) ? inv_p : (member_probe(tcol, tdeg, x, probes) ? 1.0 : inv_q);
    total = total + vw[i] * a;
  }
  const double u = ((double)(r0 >> 5) * 67108864.0 + (double)(r1 >> 6)) * (1.0 / 9007199254740992.0);
  const double target = u * total;
  double run = 0.0;
  uint32_t last = deg - 1;
  for (uint32_t i = 0; i < deg; ++i) {
    const int32_t x = vcol[i];
    const double a = (x == t) ? inv_p : (member_probe(tcol, tdeg, x, probes) ? 1.0 : inv_q);
    const double m = vw[i] * a;
    run = run + m;
    if (m > 0.0) last = i;
    if (target < run) return i;
  }
  return last; /* index into v's arc slice */
}

/*
 * Host replay of n2v_walk on a single-part graph given as plain arrays:
 *   base[V] u64, deg[V] u32, arc_thr / arc_dst / arc_alias_dst [A], col[A], weight[A].
 * Same output layout as the device: walks[W][pitch] (-1 padded), alive[W], stats[8]
 * (added into; a caller running ranges on several threads passes one stats[] per thread).
 */
int orc_replay_walk(const uint64_t* base, const uint32_t* deg, const uint32_t* arc_thr, const int32_t* arc_dst,
                    const int32_t* arc_alias_dst, const int32_t* arc_alias_idx, const float* ratio /* [A][2] or NULL */,
                    const int32_t* col, const double* weight,
                    const orc_walk_consts* c, double p, double q, const int32_t* start, int64_t n_start,
                    int32_t num_walks, int32_t walk_length, uint64_t seed, int64_t w_lo, int64_t w_hi,
                    int32_t* walks, int64_t pitch, uint8_t* alive, uint64_t* stats) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t ret_m1 = (uint32_t)(c->t_ret - 1), nbr_m1 = (uint32_t)(c->t_nbr - 1), far_m1 = (uint32_t)(c->t_far - 1);
  const uint32_t lo_m1 = nbr_m1 < far_m1 ? nbr_m1 : far_m1, hi_m1 = nbr_m1 < far_m1 ? far_m1 : nbr_m1;
  const double inv_p = 1.0 / p, inv_q = 1.0 / q;
  const int64_t W = n_start * num_walks;
  uint64_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (w_lo < 0) w_lo = 0;
  if (w_hi > W) w_hi = W;
  for (int64_t w = w_lo; w < w_hi; ++w) {
    int32_t* row = walks + w * pitch;
    for (int64_t j = 0; j < pitch; ++j) row[j] = -1;
    int32_t v = start[w / num_walks], t = -1;
    const uint64_t walk_id = (uint64_t)(uint32_t)v * (uint32_t)num_walks + (uint64_t)(w % num_walks);
    row[0] = v;
    uint8_t ok = 1;
    float r_fwd = 0.0f, r_rev = 0.0f; /* general fold: ratios of the arc that brought us to v */
    for (int32_t pos = 0; pos < walk_length; ++pos) {
      const uint32_t dv = deg[v];
      if (dv == 0) { ok = 0; ++st[6]; break; }
      uint32_t thr_out = 0, thr_ret = 0;
      if (c->fold_mode == 3 && pos > 0) {
        const uint32_t dt = deg[t];
        const float fo = (float)(dv < dt ? dv : dt) * c->mix_qm1;
        const float tot = ((float)dv + fo) + c->fold_gain;
        const float pr = c->fold_gain / tot, po = fo / tot;
        const float s = (pr + po) * 4294967296.0f;
        thr_ret = (uint32_t)(pr * 4294967296.0f);
        thr_out = s >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)s;
      } else if (c->fold_mode == 1 && pos > 0) {
        const float pr = c->fold_gain / ((float)dv + c->fold_gain);
        thr_out = pr >= 1.0f ? 0xFFFFFFFFu : (uint32_t)(pr * 4294967296.0f);
      } else if (c->fold_mode == 2 && pos > 0) {
        const float num = c->fold_gain * r_rev;
        const float pr = num / (1.0f + num);
        thr_out = pr >= 1.0f ? 0xFFFFFFFFu : (uint32_t)(pr * 4294967296.0f);
      }
      int32_t x = -1;
      int accepted = 0;
      int64_t arc_index = -1;
      for (uint32_t trial = 0; trial < (uint32_t)c->max_trials; ++trial) {
        const uint32_t ctr[4] = {(uint32_t)walk_id, (uint32_t)(walk_id >> 32), (uint32_t)pos, trial};
        uint32_t r[4];
        philox4x32_10(k0, k1, ctr, r);
        if (c->fold_mode == 3 && pos > 0) {
          if (r[0] < thr_ret) { x = t; accepted = 1; arc_index = -1; ++st[4]; break; }
          const int common = r[0] < thr_out;
          const int from_t = common && deg[t] < dv;
          const int32_t side = from_t ? t : v;
          const uint64_t e = base[side] + (uint64_t)(((uint64_t)r[1] * deg[side]) >> 32);
          const int self = r[2] < arc_thr[e];
          x = self ? arc_dst[e] : arc_alias_dst[e];
          ++st[1];
          if (!common) accepted = (x != t) || (r[3] <= ret_m1);
          else {
            const int32_t other = from_t ? v : t;
            ++st[3];
            accepted = member_probe(col + base[other], deg[other], x, &st[2]) && x != t;
          }
          if (accepted) break;
          continue;
        }
        if (c->fold_mode != 0 && pos > 0 && r[0] < thr_out) { x = t; accepted = 1; arc_index = -1; ++st[4]; break; }
        const uint64_t e = base[v] + (uint64_t)(((uint64_t)r[1] * dv) >> 32);
        const int self = r[2] < arc_thr[e];
        x = self ? arc_dst[e] : arc_alias_dst[e];
        arc_index = self ? (int64_t)e : (int64_t)(base[v] + (uint64_t)arc_alias_idx[e]);
        ++st[1];
        if (pos == 0) accepted = 1;
        else if (x == t) accepted = r[3] <= ret_m1;
        else if (r[3] <= lo_m1) accepted = 1;
        else if (r[3] > hi_m1) accepted = 0;
        else {
          ++st[3];
          accepted = r[3] <= (member_probe(col + base[t], deg[t], x, &st[2]) ? nbr_m1 : far_m1);
        }
        if (accepted) break;
      }
      if (!accepted) {
        const uint32_t ctr[4] = {(uint32_t)walk_id, (uint32_t)(walk_id >> 32), (uint32_t)pos, 0xFFFFFFFFu};
        uint32_t r[4];
        philox4x32_10(k0, k1, ctr, r);
        const uint32_t pick = exact_draw(col + base[v], weight + base[v], dv, t, col + base[t], deg[t], inv_p, inv_q,
                                         r[0], r[1], &st[2]);
        x = col[base[v] + pick];
        arc_index = (int64_t)(base[v] + pick);
        ++st[5];
      }
      if (c->fold_mode == 2) {
        if (arc_index >= 0) { r_fwd = ratio[2 * arc_index]; r_rev = ratio[2 * arc_index + 1]; }
        else { const float tmp = r_fwd; r_fwd = r_rev; r_rev = tmp; }
      }
      t = v;
      v = x;
      row[pos + 1] = v;
      ++st[0];
    }
    alive[w] = ok;
  }
  if (stats) for (int i = 0; i < 8; ++i) stats[i] += st[i];
  return 0;
}

