// Peer-shareable device buffers for the vertex-partitioned CSR.
//
// Built on the CUDA virtual-memory-management API (cuMemCreate / cuMemMap with 2 MiB
// granularity, POSIX-fd shareable handles) rather than cudaIpc*: measured on B200, random
// 32-byte gathers from cudaIpcOpenMemHandle-mapped peer memory collapse by ~50x once the remote
// footprint exceeds ~1-2 GB (translation thrash), while peer memory mapped with the
// allocation's own 2 MiB pages sustains ~11 G sectors/s at any footprint
// (profiles/r01_peer_gather.txt).  Driver entry points are fetched through
// cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda (it must load,
// and export its symbols, on a box without a driver).
//
// Handle layout (64 bytes, host): int32 fd | int32 owner device | uint64 mapped size | zeros.
// The fd inside a handle passed to n2v_ipc_open must be valid in the CALLING process: the host
// side transfers it (SCM_RIGHTS over a Unix socket, node2vec_b200/graph.py).
#include <cuda.h>
#include <string.h>
#include <unistd.h>

#include <map>
#include <mutex>

#include "n2v_internal.cuh"

namespace {

struct Rec {
  CUmemGenericAllocationHandle handle;
  size_t size;
};
std::mutex g_mu;
std::map<void*, Rec> g_recs;

struct Driver {
  CUresult (*memGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*memUnmap)(CUdeviceptr, size_t);
  CUresult (*memAddressFree)(CUdeviceptr, size_t);
  CUresult (*memRelease)(CUmemGenericAllocationHandle);
  CUresult (*memExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*memImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  bool ok;
};

bool fetch(const char* name, void** fn) {
  cudaDriverEntryPointQueryResult st;
  return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess &&
         *fn != nullptr;
}

const Driver& driver() {
  static Driver d = [] {
    Driver x{};
    x.ok = fetch("cuMemGetAllocationGranularity", reinterpret_cast<void**>(&x.memGetAllocationGranularity)) &&
           fetch("cuMemCreate", reinterpret_cast<void**>(&x.memCreate)) &&
           fetch("cuMemAddressReserve", reinterpret_cast<void**>(&x.memAddressReserve)) &&
           fetch("cuMemMap", reinterpret_cast<void**>(&x.memMap)) &&
           fetch("cuMemSetAccess", reinterpret_cast<void**>(&x.memSetAccess)) &&
           fetch("cuMemUnmap", reinterpret_cast<void**>(&x.memUnmap)) &&
           fetch("cuMemAddressFree", reinterpret_cast<void**>(&x.memAddressFree)) &&
           fetch("cuMemRelease", reinterpret_cast<void**>(&x.memRelease)) &&
           fetch("cuMemExportToShareableHandle", reinterpret_cast<void**>(&x.memExportToShareableHandle)) &&
           fetch("cuMemImportFromShareableHandle", reinterpret_cast<void**>(&x.memImportFromShareableHandle));
    return x;
  }();
  return d;
}

#define N2V_CU(call)                                                              \
  do {                                                                            \
    CUresult r__ = (call);                                                        \
    if (r__ != CUDA_SUCCESS) {                                                    \
      n2v::set_error("%s:%d: %s -> CUresult %d", __FILE__, __LINE__, #call, (int)r__); \
      return N2V_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

CUmemAllocationProp prop_for(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

int map_handle(const Driver& D, CUmemGenericAllocationHandle h, size_t size, int access_device, void** out) {
  CUdeviceptr va = 0;
  N2V_CU(D.memAddressReserve(&va, size, 0, 0, 0));
  N2V_CU(D.memMap(va, size, 0, h, 0));
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = access_device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  N2V_CU(D.memSetAccess(va, size, &acc, 1));
  *out = reinterpret_cast<void*>(va);
  std::lock_guard<std::mutex> lk(g_mu);
  g_recs[*out] = Rec{h, size};
  return N2V_OK;
}

int unmap(void* ptr) {
  const Driver& D = driver();
  Rec rec;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_recs.find(ptr);
    if (it == g_recs.end()) {
      n2v::set_error("n2v_ipc: %p is not a buffer of this library", ptr);
      return N2V_ERR_INVALID;
    }
    rec = it->second;
    g_recs.erase(it);
  }
  N2V_CUDA(cudaDeviceSynchronize());
  N2V_CU(D.memUnmap(reinterpret_cast<CUdeviceptr>(ptr), rec.size));
  N2V_CU(D.memAddressFree(reinterpret_cast<CUdeviceptr>(ptr), rec.size));
  N2V_CU(D.memRelease(rec.handle));
  return N2V_OK;
}

}  // namespace

extern "C" int n2v_ipc_alloc(size_t bytes, void** ptr) {
  N2V_CHECK_ARG(ptr != nullptr, "n2v_ipc_alloc: NULL out pointer");
  *ptr = nullptr;
  N2V_CUDA(cudaFree(0));  // make sure the primary context exists
  const Driver& D = driver();
  N2V_CHECK_ARG(D.ok, "n2v_ipc_alloc: CUDA virtual-memory-management entry points unavailable");
  int device = 0;
  N2V_CUDA(cudaGetDevice(&device));
  const CUmemAllocationProp prop = prop_for(device);
  size_t gran = 0;
  N2V_CU(D.memGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (gran == 0) gran = size_t(2) << 20;
  const size_t size = ((bytes ? bytes : 1) + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h;
  N2V_CU(D.memCreate(&h, size, &prop, 0));
  return map_handle(D, h, size, device, ptr);
}

extern "C" int n2v_ipc_free(void* ptr) { return ptr ? unmap(ptr) : N2V_OK; }

extern "C" int n2v_ipc_export(const void* ptr, unsigned char* handle) {
  N2V_CHECK_ARG(ptr && handle, "n2v_ipc_export: NULL argument");
  const Driver& D = driver();
  N2V_CHECK_ARG(D.ok, "n2v_ipc_export: CUDA virtual-memory-management entry points unavailable");
  Rec rec;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_recs.find(const_cast<void*>(ptr));
    N2V_CHECK_ARG(it != g_recs.end(), "n2v_ipc_export: %p was not allocated by n2v_ipc_alloc", ptr);
    rec = it->second;
  }
  int fd = -1;
  N2V_CU(D.memExportToShareableHandle(&fd, rec.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  int device = 0;
  N2V_CUDA(cudaGetDevice(&device));
  memset(handle, 0, N2V_IPC_HANDLE_BYTES);
  const int32_t fd32 = fd, dev32 = device;
  const uint64_t size64 = rec.size;
  memcpy(handle, &fd32, 4);
  memcpy(handle + 4, &dev32, 4);
  memcpy(handle + 8, &size64, 8);
  return N2V_OK;
}

extern "C" int n2v_ipc_open(const unsigned char* handle, void** ptr) {
  N2V_CHECK_ARG(ptr && handle, "n2v_ipc_open: NULL argument");
  *ptr = nullptr;
  N2V_CUDA(cudaFree(0));
  const Driver& D = driver();
  N2V_CHECK_ARG(D.ok, "n2v_ipc_open: CUDA virtual-memory-management entry points unavailable");
  int32_t fd32;
  uint64_t size64;
  memcpy(&fd32, handle, 4);
  memcpy(&size64, handle + 8, 8);
  CUmemGenericAllocationHandle h;
  N2V_CU(D.memImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<intptr_t>(fd32)),
                                        CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  int device = 0;
  N2V_CUDA(cudaGetDevice(&device));
  return map_handle(D, h, static_cast<size_t>(size64), device, ptr);
}

extern "C" int n2v_ipc_close(void* ptr) { return ptr ? unmap(ptr) : N2V_OK; }
