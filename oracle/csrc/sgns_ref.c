/*
 * sgns_ref.c -- CPU restatement of gensim 3.8.x skip-gram negative sampling.
 * TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/__init__.py).
 *
 * PARITY UNPINNED: the reference reaches this arithmetic only through
 * `gensim.models.Word2Vec(sentences=all_walks, **w2v_params)` (node2vec/embedding.py:126);
 * gensim ~=3.8.2 (requirements.txt:27) is neither vendored nor installable here and the
 * reference's tests never assert embedding numerics (tests/test_embedding.py:50-62).  What
 * follows restates the published algorithm of gensim 3.8 `word2vec.py` / `word2vec_inner.pyx`
 * (itself the word2vec.c algorithm of Mikolov et al.):
 *   vocabulary   : count tokens, drop count < min_count, index by descending count
 *                  (stable), keep-probability (sqrt(c/t)+1)*(t/c) with t = sample*total,
 *                  stored as round(p * 2^32)                       [prepare_vocab]
 *   negatives    : cumulative count^0.75 scaled to 2^31-1, drawn by bisect_left of
 *                  (next_random >> 16) % cum_table[-1]              [make_cum_table]
 *   random       : 48-bit LCG  r = r * 25214903917 + 11 (mod 2^48)
 *   per sentence : drop out-of-vocabulary and sub-sampled tokens; per position a reduced
 *                  window b in [0, window); contexts j in [i-window+b, i+window-b], j != i
 *   per pair     : input row syn0[word_j]; targets word_i (label 1) then `negative` draws
 *                  (label 0, skipped when equal to word_i); f = dot; skip when |f| >= 6;
 *                  sigmoid from a 1000-entry table over [-6, 6); g = (label - s) * alpha;
 *                  work += g * syn1neg[target]; syn1neg[target] += g * syn0[word_j];
 *                  finally syn0[word_j] += work                       [w2v_fast_sentence_sg_neg]
 *   schedule     : alpha decays linearly alpha -> min_alpha over `iter` epochs, updated per
 *                  job of batch_words words
 * All arithmetic fp32.  Threads: callers run disjoint sentence ranges from several host
 * threads on the SAME tables (lock-free, like gensim's `workers`).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXP_TABLE_SIZE 1000
#define MAX_EXP 6.0f

static float g_exp_table[EXP_TABLE_SIZE];
static int g_exp_ready = 0;

void orc_sgns_exp_table(float* out) {
  for (int i = 0; i < EXP_TABLE_SIZE; ++i) {
    float e = (float)exp((i / (double)EXP_TABLE_SIZE * 2 - 1) * MAX_EXP);
    out[i] = e / (e + 1.0f);
  }
}

static void ensure_exp(void) {
  if (!g_exp_ready) { orc_sgns_exp_table(g_exp_table); g_exp_ready = 1; }
}

/* prepare_vocab + make_cum_table over per-id counts (ids are the tokens).
 * Out: keep_int[V] = round(p*2^32) as uint64 (0 for dropped words), order[] = ids by
 * descending count (stable), cum_table[n_vocab] in that order.  Returns n_vocab. */
int64_t orc_sgns_vocab(const int64_t* counts, int64_t V, int64_t min_count, double sample, double ns_exponent,
                       uint64_t* keep_int, int32_t* order, uint32_t* cum_table) {
  int64_t n = 0, retain_total = 0;
  for (int64_t v = 0; v < V; ++v) {
    keep_int[v] = 0;
    if (counts[v] >= min_count && counts[v] > 0) { order[n++] = (int32_t)v; retain_total += counts[v]; }
  }
  /* stable sort by descending count: bottom-up merge sort (ties keep id order), O(n log n) */
  if (n > 1) {
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *a = order, *b = tmp;
    for (int64_t width = 1; width < n; width *= 2) {
      for (int64_t lo = 0; lo < n; lo += 2 * width) {
        const int64_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
        int64_t i = lo, j = mid, k = lo;
        while (i < mid && j < hi) b[k++] = (counts[a[j]] > counts[a[i]]) ? a[j++] : a[i++];
        while (i < mid) b[k++] = a[i++];
        while (j < hi) b[k++] = a[j++];
      }
      int32_t* t = a; a = b; b = t;
    }
    if (a != order) memcpy(order, a, sizeof(int32_t) * (size_t)n);
    free(tmp);
  }
  double threshold;
  if (sample == 0.0) threshold = (double)retain_total;
  else if (sample < 1.0) threshold = sample * (double)retain_total;
  else threshold = (double)(int64_t)(sample * (3.0 + sqrt(5.0)) / 2.0);
  double pow_total = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double c = (double)counts[order[i]];
    double p = (sqrt(c / threshold) + 1.0) * (threshold / c);
    if (p > 1.0) p = 1.0;
    keep_int[order[i]] = (uint64_t)llround(p * 4294967296.0);
    pow_total += pow(c, ns_exponent);
  }
  double cum = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    cum += pow((double)counts[order[i]], ns_exponent);
    cum_table[i] = (uint32_t)llround(cum / pow_total * 2147483647.0);
  }
  return n;
}

static inline uint64_t lcg_next(uint64_t r) { return (r * 25214903917ULL + 11ULL) & 281474976710655ULL; }

static int64_t bisect_left_u32(const uint32_t* a, uint64_t x, int64_t lo, int64_t hi) {
  while (hi > lo) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] >= x) hi = mid; else lo = mid + 1;
  }
  return lo;
}

/* one (centre, context) pair; targets given explicitly (targets[0] = centre word, label 1) */
static void pair_update(float* syn0, float* syn1neg, int64_t D, int32_t word_i, int32_t word_j,
                        const int32_t* negs, int n_negs, float alpha, float* work) {
  float* in = syn0 + (int64_t)word_j * D;
  memset(work, 0, sizeof(float) * (size_t)D);
  for (int d = 0; d <= n_negs; ++d) {
    int32_t target;
    float label;
    if (d == 0) { target = word_i; label = 1.0f; }
    else {
      target = negs[d - 1];
      if (target == word_i || target < 0) continue;
      label = 0.0f;
    }
    float* out = syn1neg + (int64_t)target * D;
    float f = 0.0f;
    for (int64_t k = 0; k < D; ++k) f += in[k] * out[k];
    if (f <= -MAX_EXP || f >= MAX_EXP) continue;
    const float s = g_exp_table[(int)((f + MAX_EXP) * (EXP_TABLE_SIZE / MAX_EXP / 2))];
    const float g = (label - s) * alpha;
    for (int64_t k = 0; k < D; ++k) work[k] += g * out[k];
    for (int64_t k = 0; k < D; ++k) out[k] += g * in[k];
  }
  for (int64_t k = 0; k < D; ++k) in[k] += work[k];
}

/*
 * gensim-3.8 train_batch_sg over sentences [s_lo, s_hi) of a rectangular corpus
 * walks[W][pitch] (len tokens per sentence, ids index the tables directly; vocab_index[id]
 * = rank in the cum_table order or -1 when dropped).  One epoch of the range.  alpha for a
 * sentence follows the per-job linear decay: progress = (epoch + done_words/total_words)/epochs
 * evaluated at job boundaries of batch_words words.  Returns pairs trained.
 */
int64_t orc_sgns_train(const int32_t* walks, int64_t W, int64_t len, int64_t pitch, const uint64_t* keep_int,
                       const int32_t* order, const uint32_t* cum_table, int64_t n_vocab, float* syn0, float* syn1neg,
                       int64_t D, int window, int negative, double alpha0, double min_alpha, int epoch, int epochs,
                       int64_t batch_words, uint64_t seed, int64_t s_lo, int64_t s_hi) {
  ensure_exp();
  float* work = (float*)malloc(sizeof(float) * (size_t)D);
  int32_t* sent = (int32_t*)malloc(sizeof(int32_t) * (size_t)len);
  int32_t* negs = (int32_t*)malloc(sizeof(int32_t) * (size_t)(negative > 0 ? negative : 1));
  uint64_t next_random = ((seed + 1) * 2654435761ULL + (uint64_t)s_lo * 97ULL + (uint64_t)epoch * 7919ULL) & 281474976710655ULL;
  int64_t pairs = 0;
  const double total_words = (double)W * (double)len;
  float alpha = (float)alpha0;
  int64_t job_words = batch_words; /* forces an alpha refresh at the first sentence */
  for (int64_t s = s_lo; s < s_hi; ++s) {
    if (job_words >= batch_words) {
      const double progress = ((double)epoch + ((double)s * (double)len) / total_words) / (double)epochs;
      double a = alpha0 - (alpha0 - min_alpha) * progress;
      if (a < min_alpha) a = min_alpha;
      alpha = (float)a;
      job_words = 0;
    }
    job_words += len;
    int64_t n = 0;
    for (int64_t k = 0; k < len; ++k) {
      const int32_t tok = walks[s * pitch + k];
      if (tok < 0 || keep_int[tok] == 0) continue;
      const uint64_t r32 = next_random >> 16;
      next_random = lcg_next(next_random);
      if (keep_int[tok] < r32) continue; /* sub-sampled away */
      sent[n++] = tok;
    }
    for (int64_t i = 0; i < n; ++i) {
      const uint64_t rb = next_random >> 16;
      next_random = lcg_next(next_random);
      const int64_t b = (int64_t)(rb % (uint64_t)window);
      int64_t j0 = i - window + b, j1 = i + window + 1 - b;
      if (j0 < 0) j0 = 0;
      if (j1 > n) j1 = n;
      for (int64_t j = j0; j < j1; ++j) {
        if (j == i) continue;
        for (int d = 0; d < negative; ++d) {
          const int64_t idx = bisect_left_u32(cum_table, (next_random >> 16) % cum_table[n_vocab - 1], 0, n_vocab);
          next_random = lcg_next(next_random);
          negs[d] = order[idx < n_vocab ? idx : n_vocab - 1];
        }
        pair_update(syn0, syn1neg, D, sent[i], sent[j], negs, negative, alpha, work);
        ++pairs;
      }
    }
  }
  free(work); free(sent); free(negs);
  return pairs;
}

/*
 * Apply an explicit pair trace with gensim's per-pair arithmetic, sequentially:
 * trace[n_pairs][2 + K] = {centre word, context word, K negatives (-1 = none)}.
 * Used to check the CUDA kernel's arithmetic on the very pairs it sampled.
 */
void orc_sgns_apply_trace(const int32_t* trace, int64_t n_pairs, int K, const float* alphas, float* syn0,
                          float* syn1neg, int64_t D) {
  ensure_exp();
  float* work = (float*)malloc(sizeof(float) * (size_t)D);
  for (int64_t p = 0; p < n_pairs; ++p) {
    const int32_t* t = trace + p * (2 + K);
    pair_update(syn0, syn1neg, D, t[0], t[1], t + 2, K, alphas[p], work);
  }
  free(work);
}

int orc_sgns_abi(void) { return 1; }
