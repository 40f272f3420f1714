// K6: "position of the first occurrence" of every row's key, on the device, without a sort.
//
// Replaces the two order-sensitive pandas steps of the reference's graph indexer
// (node2vec/indexer.py): the vertex-name -> id remap, where a name's id is the POSITION of its
// first occurrence in the concatenation [all src..., all dst...] (:26-35,
// append(ignore_index=True).drop_duplicates().reset_index()), and the undirected expansion,
// which appends the reversed arcs and keeps the FIRST occurrence of every (src, dst, weight)
// triple in frame order (:45-48, drop_duplicates()).  Both are "for row i, the smallest j with
// key[j] == key[i]".
//
// Method.  An open-addressing table whose slots hold ROW POSITIONS (u32), not keys: a slot's key
// is read through the position it stores, so keys of any width (1-3 int64 columns here) need only
// 4 bytes per slot and one 32-bit CAS.  insert(i): probe from hash(key_i); an EMPTY slot is
// claimed with atomicCAS; a slot whose stored row has an equal key takes atomicMin(i); a slot with
// a different key is skipped.  Slots never empty and a slot's key never changes, so every key
// owns exactly one slot and ends up with its smallest row position.  lookup(i) repeats the probe
// and reads that minimum.  HBM-bound random 4-byte atomics + key gathers: ~3 sectors per row.
#include "n2v_internal.cuh"

namespace {

constexpr int kBlock = 256;
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

struct Keys {
  const int64_t* c0;
  const int64_t* c1;   // may be NULL
  const int64_t* c2;   // may be NULL
};

__device__ __forceinline__ uint64_t mix64(uint64_t x) {   // splitmix64 finaliser
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

template <int NCOL>
__device__ __forceinline__ uint64_t hash_row(const Keys& K, int64_t i) {
  uint64_t h = mix64(static_cast<uint64_t>(K.c0[i]));
  if (NCOL > 1) h = mix64(h ^ static_cast<uint64_t>(K.c1[i]));
  if (NCOL > 2) h = mix64(h ^ static_cast<uint64_t>(K.c2[i]));
  return h;
}

template <int NCOL>
__device__ __forceinline__ bool same_key(const Keys& K, int64_t a, int64_t b) {
  if (K.c0[a] != K.c0[b]) return false;
  if (NCOL > 1 && K.c1[a] != K.c1[b]) return false;
  if (NCOL > 2 && K.c2[a] != K.c2[b]) return false;
  return true;
}

inline int grid_for(int64_t n) {
  const int64_t need = (n + kBlock - 1) / kBlock;
  const int64_t cap = int64_t(n2v::sm_count()) * 16;
  return static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
}

// keys made of 1-3 int64 columns
template <int NCOL>
struct ColumnKeys {
  Keys K;
  __device__ __forceinline__ uint64_t hash(int64_t i) const { return hash_row<NCOL>(K, i); }
  __device__ __forceinline__ bool same(int64_t a, int64_t b) const { return same_key<NCOL>(K, a, b); }
};

// variable-length byte-string keys in Arrow layout: row i is data[offsets[i] .. offsets[i+1])
struct ByteKeys {
  const int64_t* offsets;
  const uint8_t* data;
  __device__ __forceinline__ uint64_t hash(int64_t i) const {
    uint64_t h = 0xCBF29CE484222325ull;                 // FNV-1a over the bytes, then the splitmix64 finaliser
    for (int64_t k = offsets[i]; k < offsets[i + 1]; ++k) h = (h ^ data[k]) * 0x100000001B3ull;
    return mix64(h ^ static_cast<uint64_t>(offsets[i + 1] - offsets[i]));
  }
  __device__ __forceinline__ bool same(int64_t a, int64_t b) const {
    const int64_t la = offsets[a + 1] - offsets[a];
    if (la != offsets[b + 1] - offsets[b]) return false;
    const uint8_t* pa = data + offsets[a];
    const uint8_t* pb = data + offsets[b];
    for (int64_t k = 0; k < la; ++k)
      if (pa[k] != pb[k]) return false;
    return true;
  }
};

template <typename KeyOps>
__global__ void insert_rows(KeyOps ops, int64_t n, uint32_t* __restrict__ table, uint64_t mask) {
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock) {
    uint64_t h = ops.hash(i) & mask;
    for (;;) {
      uint32_t cur = *reinterpret_cast<volatile uint32_t*>(table + h);
      if (cur == kEmpty) {
        cur = atomicCAS(table + h, kEmpty, static_cast<uint32_t>(i));
        if (cur == kEmpty) break;                      // claimed: we are this key's first row so far
      }
      if (ops.same(cur, i)) {
        if (static_cast<uint32_t>(i) < cur) atomicMin(table + h, static_cast<uint32_t>(i));
        break;
      }
      h = (h + 1) & mask;
    }
  }
}

template <typename KeyOps>
__global__ void lookup_rows(KeyOps ops, int64_t n, const uint32_t* __restrict__ table, uint64_t mask,
                            int64_t* __restrict__ first) {
  for (int64_t i = blockIdx.x * int64_t(kBlock) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock) {
    uint64_t h = ops.hash(i) & mask;
    for (;;) {
      const uint32_t cur = table[h];
      if (cur == kEmpty) {                             // cannot happen after insert_rows; never read out of bounds
        first[i] = -1;
        break;
      }
      if (ops.same(cur, i)) {                          // every key was inserted: the probe always ends here
        first[i] = static_cast<int64_t>(cur);
        break;
      }
      h = (h + 1) & mask;
    }
  }
}

template <typename KeyOps>
int run_first_occurrence(const KeyOps& ops, int64_t n_rows, uint32_t* table, int64_t n_slots, int64_t* first_out,
                         cudaStream_t stream) {
  N2V_CUDA(cudaMemsetAsync(table, 0xFF, sizeof(uint32_t) * static_cast<size_t>(n_slots), stream));
  const uint64_t mask = static_cast<uint64_t>(n_slots - 1);
  const int grid = grid_for(n_rows);
  insert_rows<KeyOps><<<grid, kBlock, 0, stream>>>(ops, n_rows, table, mask);
  lookup_rows<KeyOps><<<grid, kBlock, 0, stream>>>(ops, n_rows, table, mask, first_out);
  N2V_LAUNCH_OK();
  return N2V_OK;
}

}  // namespace

extern "C" int64_t n2v_first_occurrence_slots(int64_t n_rows) {
  int64_t cap = 1024;
  while (cap < 2 * n_rows) cap <<= 1;                   // load factor <= 0.5
  return cap;
}

extern "C" int n2v_first_occurrence(const int64_t* key0, const int64_t* key1, const int64_t* key2, int64_t n_rows,
                                    uint32_t* table, int64_t n_slots, int64_t* first_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(n_rows >= 0 && n_rows < int64_t(0xFFFFFFFF), "n2v_first_occurrence: %lld rows exceed the 2^32 - 1 limit",
                static_cast<long long>(n_rows));
  if (n_rows == 0) return N2V_OK;
  N2V_CHECK_ARG(key0 && table && first_out, "n2v_first_occurrence: NULL buffer");
  N2V_CHECK_ARG(key1 != nullptr || key2 == nullptr, "n2v_first_occurrence: key2 without key1");
  N2V_CHECK_ARG(n_slots >= 2 * n_rows && (n_slots & (n_slots - 1)) == 0,
                "n2v_first_occurrence: n_slots %lld must be a power of two >= 2 * n_rows", static_cast<long long>(n_slots));
  const Keys K{key0, key1, key2};
  if (key2) return run_first_occurrence(ColumnKeys<3>{K}, n_rows, table, n_slots, first_out, stream);
  if (key1) return run_first_occurrence(ColumnKeys<2>{K}, n_rows, table, n_slots, first_out, stream);
  return run_first_occurrence(ColumnKeys<1>{K}, n_rows, table, n_slots, first_out, stream);
}

extern "C" int n2v_first_occurrence_bytes(const int64_t* offsets, const uint8_t* data, int64_t n_rows,
                                          uint32_t* table, int64_t n_slots, int64_t* first_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  N2V_CHECK_ARG(n_rows >= 0 && n_rows < int64_t(0xFFFFFFFF), "n2v_first_occurrence_bytes: %lld rows exceed the 2^32 - 1 limit",
                static_cast<long long>(n_rows));
  if (n_rows == 0) return N2V_OK;
  N2V_CHECK_ARG(offsets && table && first_out, "n2v_first_occurrence_bytes: NULL buffer");   // data may be NULL: all rows empty
  N2V_CHECK_ARG(n_slots >= 2 * n_rows && (n_slots & (n_slots - 1)) == 0,
                "n2v_first_occurrence_bytes: n_slots %lld must be a power of two >= 2 * n_rows", static_cast<long long>(n_slots));
  return run_first_occurrence(ByteKeys{offsets, data}, n_rows, table, n_slots, first_out, stream);
}
