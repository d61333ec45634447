// Multi-GPU exchange of the scoring path: ONE NCCL all-gather of compact per-pair scores, enqueued
// on the compute stream.  Replaces the reference's dist.barrier() + 2-6 dense N x N all_reduce(SUM) of -100-filled
// matrices (retrieval_utils.py:252-262).
//
// NCCL is resolved at run time (dlopen) instead of at link time: a process that already carries an NCCL (PyTorch's
// bundled libnccl.so.2) must not get a second copy, and the single-GPU product must load on boxes without NCCL.
// Order: the copy already loaded into the process (RTLD_NOLOAD), $BLIM_NCCL_LIB, then the dynamic linker's search path.
// Only the few entry points used here are declared; their signatures are NCCL's public C API (nccl.h, stable since 2.x).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>

namespace blim {

struct NcclUniqueId { char internal[128]; };   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;                         // ncclComm_t (opaque pointer)
enum { kNcclSuccess = 0, kNcclFloat32 = 7 };    // ncclSuccess, ncclFloat32

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string error;

  bool load() {
    if (AllGather) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
      if (!handle) handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);     // the NCCL this process already uses, if any
    if (!handle)
      if (const char* env = getenv("BLIM_NCCL_LIB")) handle = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    for (const char* n : names)
      if (!handle) handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (!handle) {
      error = std::string("NCCL not found (dlopen libnccl.so.2; set BLIM_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
      return false;
    }
    auto sym = [&](const char* s) { return dlsym(handle, s); };
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    GetVersion = reinterpret_cast<decltype(GetVersion)>(sym("ncclGetVersion"));
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather || !GetErrorString) {
      error = "the loaded NCCL lacks ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy / ncclAllGather";
      AllGather = nullptr;
      return false;
    }
    return true;
  }
  std::string describe(int rc) const { return GetErrorString ? std::string(GetErrorString(rc)) : std::string("nccl error ") + std::to_string(rc); }
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  return api;
}

}  // namespace blim
