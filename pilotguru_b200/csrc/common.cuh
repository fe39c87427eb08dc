// Shared host-side plumbing for libpgb200.so: error reporting, launch accounting, small RAII helpers.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: no-ops unless a profiler is attached

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/pgb200.h"

namespace pgb {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define PGB_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return pgb::fail(PGB_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define PGB_LAUNCHED() (pgb::g_launches.fetch_add(1, std::memory_order_relaxed))

#define PGB_CHECK_LAUNCH()                                                                             \
  do {                                                                                                 \
    PGB_LAUNCHED();                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                               \
    if (_e != cudaSuccess)                                                                             \
      return pgb::fail(PGB_ERR_CUDA, "%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

// Picks the device and verifies it can run sm_100a code.  No fallback of any kind.
int use_device(int device);

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    PGB_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
    n = count;
    return PGB_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// Per-call scratch for the entry points that have no handle to keep buffers in: memory from the device's stream-ordered
// pool (cudaMallocAsync / cudaFreeAsync on the call's stream).  The pool keeps what it is given back (release threshold
// raised once per device), so steady-state calls do not reach the driver's allocator -- ten cudaMalloc/cudaFree pairs
// per call cost more than the kernels they serve (pgb_pose_optimization: 9.4 -> 5.0 ms per 256-frame batch).
// Usage: `TempScope scope(device, stream);` first, then TempBuf<T> objects with DevBuf's interface.
int keep_pool_memory(int device);
struct TempScope {
  static cudaStream_t& current() { static thread_local cudaStream_t s = nullptr; return s; }
  cudaStream_t prev;
  int rc;
  TempScope(int device, cudaStream_t s) : prev(current()), rc(keep_pool_memory(device)) { current() = s; }
  ~TempScope() { current() = prev; }
};
template <typename T>
struct TempBuf {
  T* p = nullptr;
  cudaStream_t s = nullptr;
  int alloc(size_t count) {
    release();
    s = TempScope::current();
    PGB_CUDA(cudaMallocAsync((void**)&p, (count ? count : 1) * sizeof(T), s));
    return PGB_OK;
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
  }
  ~TempBuf() { release(); }
  TempBuf() {}
  TempBuf(const TempBuf&) = delete;
  TempBuf& operator=(const TempBuf&) = delete;
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// NVTX range around a host-side phase (kernel launches of a stage, a collective, a solver call): shows up as a named span in
// Nsight Systems / ncu --nvtx captures (tracing hooks).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (function, device) pair, not to a handle: one limit object
// per kernel keeps the process-wide maximum ever requested on each device and only ever raises the attribute, so a
// handle with a smaller capacity can never lower it under another handle (or thread) that relies on the larger value.
struct DynSmemLimit {
  std::atomic<size_t> perDevice[64];
  DynSmemLimit() { for (auto& v : perDevice) v.store(0); }
  template <typename F>
  int ensure(F fn, size_t bytes) {
    if (bytes <= 48 * 1024) return PGB_OK;
    int dev = 0;
    PGB_CUDA(cudaGetDevice(&dev));
    std::atomic<size_t>& cur = perDevice[dev & 63];
    size_t seen = cur.load();
    while (bytes > seen) {
      PGB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      if (cur.compare_exchange_weak(seen, bytes)) break;
    }
    return PGB_OK;
  }
};

}  // namespace pgb
