// Shared helpers for the l2i_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/l2i_b200.h"

namespace l2i {

// ---- error reporting -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return L2I_ERR_CUDA;
  }
  return L2I_OK;
}

#define L2I_CUDA_TRY(expr)                                                        \
  do {                                                                            \
    cudaError_t e__ = (expr);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      ::l2i::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));          \
      return L2I_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

#define L2I_TRY(...)               \
  do {                             \
    int rc__ = (__VA_ARGS__);      \
    if (rc__ != L2I_OK) return rc__; \
  } while (0)

#define L2I_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::l2i::set_error(__VA_ARGS__);    \
      return L2I_ERR_INVALID_ARG;       \
    }                                   \
  } while (0)

// ---- scalar type traits ----------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<double>(double v) { return (float)v; }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

}  // namespace l2i
