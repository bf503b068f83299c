// common.cuh -- shared helpers for libb200woq (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/b200woq.h"

namespace b200woq {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);  // kernels launched through this library since load (b200woq_launch_count)

#define WOQ_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      b200woq::set_error(__VA_ARGS__);    \
      return B200WOQ_EINVAL;              \
    }                                     \
  } while (0)

#define WOQ_CUDA(call)                                                                           \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      b200woq::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200WOQ_ECUDA;                                                                      \
    }                                                                                            \
  } while (0)

#define WOQ_LAUNCH_CHECK()                                                                        \
  do {                                                                                            \
    b200woq::count_launch(1);                                                                     \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess) {                                                                      \
      b200woq::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200WOQ_ECUDA;                                                                       \
    }                                                                                             \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int num_sms();  // cached SM count of the current device (148 on B200)

// ---- element-type helpers: load as float, round a float through the storage type ----
template <typename T> struct ElemTraits;
template <> struct ElemTraits<float> {
  static __device__ __forceinline__ float load(const float* p) { return *p; }
  static __device__ __forceinline__ float round(float v) { return v; }
  static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
};
template <> struct ElemTraits<__half> {
  static __device__ __forceinline__ float load(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ float round(float v) { return __half2float(__float2half_rn(v)); }
  static __device__ __forceinline__ void store(__half* p, float v) { *p = __float2half_rn(v); }
};
template <> struct ElemTraits<__nv_bfloat16> {
  static __device__ __forceinline__ float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ float round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// effective group size and group count as quant_tensor computes them (utility.py:303-305)
static inline int eff_group(int64_t K, int group_size) {
  return (group_size <= 0 || (int64_t)group_size > K) ? (int)K : group_size;
}

}  // namespace b200woq
