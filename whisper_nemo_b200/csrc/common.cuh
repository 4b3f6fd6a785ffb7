// Shared helpers for libb200d.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/b200d.h"

namespace b200d {

extern thread_local char g_last_error[512];

inline int set_error(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_last_error, sizeof(g_last_error), fmt, a, b);
  return code;
}

#define B200D_CHECK_ARG(cond)                                                        \
  do {                                                                               \
    if (!(cond)) return b200d::set_error(B200D_EINVAL, "%s: invalid argument: %s", __func__, #cond); \
  } while (0)

#define B200D_CHECK_LAUNCH()                                                          \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess)                                                           \
      return b200d::set_error(B200D_ELAUNCH, "%s: %s", __func__, cudaGetErrorString(e__)); \
  } while (0)

#define B200D_CHECK_CUDA(expr)                                                        \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess)                                                           \
      return b200d::set_error(B200D_ELAUNCH, "%s: %s", __func__, cudaGetErrorString(e__)); \
  } while (0)

constexpr int kNumSMs = 148;
constexpr int kMaxDevices = 16;  // per-device one-time setup flags (function attributes, occupancy queries)
inline int current_device() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace b200d
