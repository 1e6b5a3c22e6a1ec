// Shared host/device helpers for the msg_b200 C-ABI library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "msg_b200.h"

namespace msg {

// ---- error reporting ----------------------------------------------------------------------------
extern thread_local char g_last_error[512];
extern std::atomic<uint64_t> g_launch_count;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

inline void count_launch(uint64_t n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

// Checks the launch (not the execution): the library never synchronises.
#define MSG_CHECK_LAUNCH(what)                                                            \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess)                                                               \
      return ::msg::fail(MSG_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e__)); \
    ::msg::count_launch();                                                                \
  } while (0)

#define MSG_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return ::msg::fail(MSG_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));  \
  } while (0)

inline int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
  }
  return cached;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// per-device "already configured" flags (function attributes belong to the device's context)
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  return dev;
}

// ---- device helpers -----------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  `scratch` holds >= 32 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? scratch[threadIdx.x] : T(0);
  if (wid == 0) v = warp_sum(v);
  return v;
}

}  // namespace msg
