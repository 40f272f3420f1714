/* A plain-C11 consumer of libn2v_b200.so: no Python, no torch, only the CUDA runtime for device
 * memory.  It is what a cgo / JNI / N-API binding of the reference's hot path would do:
 *
 *   arcs -> n2v_csr_build -> n2v_hash_build -> n2v_alias_build -> n2v_walk -> (walks, alive)
 *
 * on a small weighted directed graph (a ring with chords plus one sink), then checks on the host
 * that every hop of every surviving walk is an arc of the graph, that walks start where asked, that
 * walkers dropped at the sink are flagged, and that a second run with the same seed is identical.
 * Build + run: tests/test_gpu_walk.py::test_plain_c_consumer (or see INTEGRATION.md section 4).
 * Exit code 0 and a line "c_consumer OK ..." on success.
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "n2v_b200.h"

#define CU(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s:%d CUDA: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));     \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)
#define N2V(x)                                                                             \
  do {                                                                                     \
    int rc_ = (x);                                                                         \
    if (rc_ != 0) {                                                                        \
      fprintf(stderr, "%s:%d n2v rc=%d: %s\n", __FILE__, __LINE__, rc_, n2v_last_error()); \
      return 3;                                                                            \
    }                                                                                      \
  } while (0)

enum { V = 1000, CHORDS = 3, NUM_WALKS = 4, WALK_LEN = 12 };

static int has_arc(const int32_t* src, const int32_t* dst, int64_t n, int32_t a, int32_t b) {
  for (int64_t i = 0; i < n; ++i)
    if (src[i] == a && dst[i] == b) return 1;
  return 0;
}

static int run_walk(const n2v_graph_t* g, const int32_t* d_start, int64_t n_start, uint64_t seed, int32_t* h_walks,
                    uint8_t* h_alive, int64_t pitch, cudaStream_t stream) {
  const int64_t W = n_start * NUM_WALKS;
  int32_t* d_walks;
  uint8_t* d_alive;
  CU(cudaMalloc((void**)&d_walks, (size_t)W * pitch * sizeof(int32_t)));
  CU(cudaMalloc((void**)&d_alive, (size_t)W));
  N2V(n2v_walk(g, d_start, n_start, NUM_WALKS, WALK_LEN, 0.5, 2.0, seed, d_walks, pitch, d_alive, NULL, stream));
  CU(cudaMemcpyAsync(h_walks, d_walks, (size_t)W * pitch * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  CU(cudaMemcpyAsync(h_alive, d_alive, (size_t)W, cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));
  CU(cudaFree(d_walks));
  CU(cudaFree(d_alive));
  return 0;
}

int main(void) {
  if (n2v_abi_version() != N2V_ABI_VERSION) {
    fprintf(stderr, "ABI mismatch: header %d, library %d\n", N2V_ABI_VERSION, n2v_abi_version());
    return 1;
  }
  /* vertex V-1 is a sink; everybody else has a ring arc and CHORDS chords (one may hit the sink) */
  const int64_t A = (int64_t)(V - 1) * (1 + CHORDS);
  int32_t* src = malloc(A * sizeof(int32_t));
  int32_t* dst = malloc(A * sizeof(int32_t));
  double* wgt = malloc(A * sizeof(double));
  uint32_t lcg = 12345u;
  int64_t k = 0;
  for (int32_t v = 0; v < V - 1; ++v) {
    src[k] = v, dst[k] = (v + 1) % (V - 1), wgt[k] = 1.0, ++k;
    for (int c = 0; c < CHORDS; ++c) {
      lcg = lcg * 1664525u + 1013904223u;
      const int32_t to = (c == 0 && v % 50 == 0) ? V - 1 : (int32_t)((lcg >> 8) % V); /* some arcs into the sink */
      src[k] = v, dst[k] = to, wgt[k] = 0.25 + (double)((lcg >> 4) & 15) / 8.0, ++k;
    }
  }

  cudaStream_t stream;
  CU(cudaSetDevice(0));
  CU(cudaStreamCreate(&stream));
  int32_t *d_src, *d_dst, *d_col, *d_hash, *d_work, *d_start;
  double *d_w, *d_ws, *d_probs;
  n2v_vertex_t* d_vtx;
  n2v_arc_t* d_arcs;
  void* d_scratch;
  const size_t scratch_bytes = n2v_csr_scratch_bytes(A, V);
  const int64_t n_buckets = n2v_hash_buckets_bound(A, V);
  CU(cudaMalloc((void**)&d_src, A * sizeof(int32_t)));
  CU(cudaMalloc((void**)&d_dst, A * sizeof(int32_t)));
  CU(cudaMalloc((void**)&d_w, A * sizeof(double)));
  CU(cudaMalloc((void**)&d_vtx, V * sizeof(n2v_vertex_t)));
  CU(cudaMalloc((void**)&d_col, A * sizeof(int32_t)));
  CU(cudaMalloc((void**)&d_ws, A * sizeof(double)));
  CU(cudaMalloc(&d_scratch, scratch_bytes));
  CU(cudaMalloc((void**)&d_hash, (size_t)n_buckets * 8 * sizeof(int32_t)));
  CU(cudaMalloc((void**)&d_arcs, A * sizeof(n2v_arc_t)));
  CU(cudaMalloc((void**)&d_probs, A * sizeof(double)));
  CU(cudaMalloc((void**)&d_work, A * sizeof(int32_t)));
  CU(cudaMemcpyAsync(d_src, src, A * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync(d_dst, dst, A * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync(d_w, wgt, A * sizeof(double), cudaMemcpyHostToDevice, stream));

  uint32_t flags = 0;
  int64_t n_zero = 0;
  N2V(n2v_csr_build(d_src, d_dst, d_w, A, V, V, d_vtx, d_col, d_ws, NULL, d_scratch, scratch_bytes, &flags, stream));
  N2V(n2v_hash_build(d_vtx, d_col, V, A, d_hash, n_buckets, stream));
  N2V(n2v_alias_build(d_vtx, NULL, d_col, d_ws, V, A, N2V_SUM_NAIVE, NULL, d_probs, d_arcs, d_work, &n_zero, stream));
  if (n_zero != 0) {
    fprintf(stderr, "unexpected zero-weight vertices: %lld\n", (long long)n_zero);
    return 4;
  }

  n2v_graph_t g;
  memset(&g, 0, sizeof(g));
  g.n_vertices = V, g.n_arcs = A, g.flags = flags, g.n_parts = 1, g.part_size = V;
  g.parts[0].vtx = d_vtx, g.parts[0].arcs = d_arcs, g.parts[0].col = d_col, g.parts[0].weight = d_ws;
  g.parts[0].hash = d_hash, g.parts[0].ratio = NULL;

  /* walk_start = every vertex with an out-arc (fugue.py:132) */
  const int64_t n_start = V - 1, W = n_start * NUM_WALKS, pitch = (WALK_LEN + 8) / 8 * 8;
  int32_t* start = malloc(n_start * sizeof(int32_t));
  for (int32_t v = 0; v < n_start; ++v) start[v] = v;
  CU(cudaMalloc((void**)&d_start, n_start * sizeof(int32_t)));
  CU(cudaMemcpyAsync(d_start, start, n_start * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  int32_t* walks = malloc(W * pitch * sizeof(int32_t));
  int32_t* again = malloc(W * pitch * sizeof(int32_t));
  uint8_t* alive = malloc(W);
  uint8_t* alive2 = malloc(W);
  int rc = run_walk(&g, d_start, n_start, 7, walks, alive, pitch, stream);
  if (rc) return rc;
  rc = run_walk(&g, d_start, n_start, 7, again, alive2, pitch, stream);
  if (rc) return rc;

  int64_t n_alive = 0, n_dead = 0, hops = 0;
  for (int64_t w = 0; w < W; ++w) {
    const int32_t* row = walks + w * pitch;
    if (memcmp(row, again + w * pitch, (WALK_LEN + 1) * sizeof(int32_t)) != 0 || alive[w] != alive2[w]) {
      fprintf(stderr, "walk %lld not reproducible under the same seed\n", (long long)w);
      return 5;
    }
    if (row[0] != start[w / NUM_WALKS]) {
      fprintf(stderr, "walk %lld starts at %d, expected %d\n", (long long)w, row[0], start[w / NUM_WALKS]);
      return 6;
    }
    int len = 1;
    while (len <= WALK_LEN && row[len] >= 0) ++len;
    for (int i = 0; i + 1 < len; ++i, ++hops)
      if (!has_arc(src, dst, A, row[i], row[i + 1])) {
        fprintf(stderr, "walk %lld hop %d: (%d -> %d) is not an arc\n", (long long)w, i, row[i], row[i + 1]);
        return 7;
      }
    if (alive[w]) {
      if (len != WALK_LEN + 1) return 8;
      ++n_alive;
    } else {
      if (row[len - 1] != V - 1) {    /* the only way to die is to stand on the sink */
        fprintf(stderr, "walk %lld dropped at %d, which is not the sink\n", (long long)w, row[len - 1]);
        return 9;
      }
      ++n_dead;
    }
  }
  /* a bad argument must come back as a status code + message, not a crash */
  if (n2v_walk(&g, d_start, n_start, NUM_WALKS, WALK_LEN, 0.0, 1.0, 7, NULL, pitch, NULL, NULL, stream) == 0 ||
      strlen(n2v_last_error()) == 0) {
    fprintf(stderr, "p = 0 was accepted\n");
    return 10;
  }
  printf("c_consumer OK abi=%d flags=0x%x walkers=%lld alive=%lld dropped_at_sink=%lld hops_checked=%lld\n",
         n2v_abi_version(), flags, (long long)W, (long long)n_alive, (long long)n_dead, (long long)hops);
  return 0;
}
