// Shared host-side plumbing for libpgb200.so: error reporting, launch accounting, small RAII helpers.
#pragma once
#include <cuda_runtime.h>

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

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace pgb
