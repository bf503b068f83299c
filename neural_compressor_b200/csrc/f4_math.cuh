// f4_math.cuh -- per-element math of the 4-bit TABLE data types (nf4, fp4 = fp4_e2m1_bnb, fp4_e2m1).
//
// Reference: neural_compressor/torch/algorithms/weight_only/utility.py
//     FLOAT_MAPPING / INT_MAPPING :52-103     quantize_4bit :121-160
// The functions are __host__ __device__ and free of CUDA types so that tests/test_f4_math_cpu.py can compile this very
// header with g++ (B200WOQ_F4_HOST) and check it element by element against tensors written by the live reference --
// the kernels in float4.cu only add the grouping, the group reduction and the loads / stores around it.
//
// Rounding model.  torch evaluates every step of quantize_4bit in the weight's storage type T with float opmath, i.e.
// each op's float result is rounded through T.  `R::round` is that rounding (identity for fp32).  The scalar operands
// (quantile, max level, the mid points, the level values) are Python doubles that torch converts to float first.
#pragma once
#include <stdint.h>

#ifdef B200WOQ_F4_HOST
#define F4_HD inline
#define F4_MUL(a, b) ((a) * (b))
#define F4_DIV(a, b) ((a) / (b))
#else
#define F4_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define F4_MUL(a, b) __fmul_rn((a), (b))
#define F4_DIV(a, b) __fdiv_rn((a), (b))
#else
#define F4_MUL(a, b) ((a) * (b))
#define F4_DIV(a, b) ((a) / (b))
#endif
#endif

#include "../../include/b200woq.h"

namespace b200woq {

// scale = absmax * quantile / max(levels)   (utility.py:138), each op rounded through T
template <typename R>
F4_HD float f4_group_scale(float absmax, float quantile, float max_level) {
  return R::round(F4_DIV(R::round(F4_MUL(absmax, quantile)), max_level));
}

// Index of the level the reference selects for w / scale, or -1 when no interval matches (w / scale is NaN for an
// all-zero group: every `torch.where` condition is false and the element keeps code 0 / value 0).  utility.py:141-150:
//   i == 0      : t <= mid[0]
//   0 < i < n-1 : mid[i-1] < t <= mid[i]
//   i == n-1    : t > mid[n-2]
template <typename R>
F4_HD int f4_select(float w, float scale, const b200woq_f4_table& t) {
  const float v = R::round(F4_DIV(w, scale));
  if (!(v == v)) return -1;
  int i = 0;
  while (i < t.n - 1 && !(v <= R::round(t.mid[i]))) ++i;
  return i;
}

// integer code the reference leaves in the tensor with return_int=True (INT_MAPPING)
F4_HD int f4_code(int idx, const b200woq_f4_table& t) { return idx < 0 ? 0 : t.code[idx]; }

// fake-quantised value: q_tensor (T) accumulates the fp32 level, then tensor.mul_(scale)   (utility.py:146-155)
template <typename R>
F4_HD float f4_fake(int idx, float scale, const b200woq_f4_table& t) {
  const float q = idx < 0 ? 0.f : R::round(t.level[idx]);
  return R::round(F4_MUL(q, scale));
}

// recover(): the nibble's sign-extended integer -> level (modules.py:392-396), times the fp32 scale (:437-440)
F4_HD float f4_recover(uint32_t nibble, float scale, const float* nibble_level) {
  return F4_MUL(nibble_level[nibble & 0xFu], scale);
}

}  // namespace b200woq
