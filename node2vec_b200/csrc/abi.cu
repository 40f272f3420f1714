// ABI housekeeping: version, thread-local error string, walk sampling constants.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "n2v_internal.cuh"

namespace n2v {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace n2v

extern "C" int n2v_abi_version(void) { return N2V_ABI_VERSION; }

extern "C" const char* n2v_last_error(void) { return n2v::g_err; }

// accept iff u32 < T, T in [1, 2^32]
static uint64_t accept_threshold(double a) {
  if (a >= 1.0) return 4294967296ull;
  const double s = floor(a * 4294967296.0 + 0.5);
  if (s < 1.0) return 1ull;
  if (s >= 4294967296.0) return 4294967296ull;
  return static_cast<uint64_t>(s);
}

// Rejection sampling of  P(x) ~ w(v,x) * alpha(t,x)  by proposing x ~ w(v,.) from the
// first-order alias table and accepting with alpha/cap (randomwalk.py:223-230 defines
// alpha: 1/p back to t, 1 into N_out(t), 1/q elsewhere).  When 1/p is the only thing
// above cap' = max(1, 1/q), its excess mass on the single return arc is drawn as a
// separate mixture component ("fold") so the envelope stays at cap'.
extern "C" int n2v_walk_consts(double return_param, double inout_param, uint32_t graph_flags, int has_ratio,
                               n2v_walk_consts_t* out) {
  N2V_CHECK_ARG(out != nullptr, "n2v_walk_consts: out is NULL");
  N2V_CHECK_ARG(return_param > 0.0 && inout_param > 0.0 && isfinite(return_param) &&
                    isfinite(inout_param),
                "Zero return (%g) or inout (%g) parameter!", return_param, inout_param);
  const double ip = 1.0 / return_param, iq = 1.0 / inout_param;
  const uint32_t need = N2V_GRAPH_UNIT_WEIGHT | N2V_GRAPH_SYMMETRIC | N2V_GRAPH_SIMPLE;
  double cap = iq > 1.0 ? iq : 1.0;
  memset(out, 0, sizeof(*out));
  if ((graph_flags & need) == need && inout_param > 1.0) {
    // mixture sampler (mode 3, see n2v_b200.h): bulk + common-neighbour + return components
    const double g = inout_param / return_param - 1.0;
    out->fold_mode = 3;
    out->fold_gain = static_cast<float>(g > 0.0 ? g : 0.0);
    out->mix_qm1 = static_cast<float>(inout_param - 1.0);
    const double a_ret = inout_param / return_param;      // bulk acceptance of x == prev: min(1, (1/p) / (1/q))
    out->t_ret = accept_threshold(a_ret < 1.0 ? a_ret : 1.0);
    out->t_nbr = out->t_far = 4294967296ull;
    out->max_trials = 256;
    return N2V_OK;
  }
  if (ip > cap) {
    if ((graph_flags & need) == need) {
      out->fold_mode = 1;
      out->fold_gain = static_cast<float>((ip - cap) / cap);
    } else if (has_ratio) {
      out->fold_mode = 2;
      out->fold_gain = static_cast<float>((ip - cap) / cap);
    } else {
      cap = ip;  // no fold available: widen the envelope instead
    }
  }
  const double a_ret = (ip < cap ? ip : cap) / cap;
  out->t_ret = accept_threshold(a_ret);
  out->t_nbr = accept_threshold(1.0 / cap);
  out->t_far = accept_threshold(iq / cap);
  out->max_trials = 256;
  return N2V_OK;
}

extern "C" int n2v_set_l2_fetch_granularity(int bytes) {
  N2V_CHECK_ARG(bytes == 32 || bytes == 64 || bytes == 128, "n2v_set_l2_fetch_granularity: %d is not 32, 64 or 128", bytes);
  N2V_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, static_cast<size_t>(bytes)));
  return N2V_OK;
}

extern "C" int n2v_get_l2_fetch_granularity(void) {
  size_t v = 0;
  if (cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity) != cudaSuccess) return -1;
  return static_cast<int>(v);
}
