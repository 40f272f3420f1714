// Device core of the alias-table construction, shared by alias_build.cu (per-vertex tables)
// and vocab.cu (the negative-sampling table).  Bit-exact restatement of the reference's
// generate_alias_tables (randomwalk.py:157-190); see alias_build.cu for the commentary.
#pragma once
#include "n2v_internal.cuh"

namespace n2v {

// sum(list) starting from int 0, as the interpreter evaluates it (see n2v_b200.h sum modes)
template <typename F>
__device__ double python_sum(F value, uint32_t n, int sum_mode) {
  double total = value(0);  // 0 + v0 is exact
  if (sum_mode == N2V_SUM_NAIVE) {
    for (uint32_t i = 1; i < n; ++i) total = __dadd_rn(total, value(i));
    return total;
  }
  double comp = 0.0;
  for (uint32_t i = 1; i < n; ++i) {
    const double v = value(i);
    const double t = __dadd_rn(total, v);
    if (fabs(total) >= fabs(v))
      comp = __dadd_rn(comp, __dadd_rn(__dsub_rn(total, t), v));
    else
      comp = __dadd_rn(comp, __dadd_rn(__dsub_rn(v, t), total));
    total = t;
  }
  if (comp != 0.0 && isfinite(comp)) total = __dadd_rn(total, comp);
  return total;
}

// probs[] holds the raw weights on entry and the alias "probs" on exit.
// stack[] is an n-slot int32 slice: the underfull list grows up from slot 0, the overfull
// list grows down from slot n-1 (every index is on exactly one list, so they never meet).
// Returns false when the weights sum to zero (reference: ZeroDivisionError).
template <typename AliasStore>
__device__ bool build_alias_one(double* __restrict__ probs, uint32_t n, int sum_mode,
                                int32_t* __restrict__ stack, AliasStore store_alias) {
  const double total = python_sum([&](uint32_t i) { return probs[i]; }, n, sum_mode);
  const double mean = __ddiv_rn(total, static_cast<double>(n));
  if (!(mean != 0.0)) return false;
  int64_t n_small = 0, n_large = 0;  // list sizes
  for (uint32_t i = 0; i < n; ++i) {
    const double pr = __ddiv_rn(probs[i], mean);
    probs[i] = pr;
    store_alias(i, 0);
    if (pr < 1.0) stack[n_small++] = static_cast<int32_t>(i);
    else stack[n - 1 - (n_large++)] = static_cast<int32_t>(i);
  }
  while (n_small > 0 && n_large > 0) {
    const int32_t lo = stack[--n_small];
    const int32_t hi = stack[n - 1 - (--n_large)];
    store_alias(static_cast<uint32_t>(lo), hi);
    const double ph = __dsub_rn(__dadd_rn(probs[hi], probs[lo]), 1.0);
    probs[hi] = ph;
    if (ph < 1.0) stack[n_small++] = hi;
    else stack[n - 1 - (n_large++)] = hi;
  }
  return true;
}

}  // namespace n2v
