// rtn_math.cuh -- per-group RTN parameter / code math shared by rtn_pack.cu and the GPTQ double-quant kernel.
#pragma once
#include "common.cuh"

namespace b200woq {

// ------------------------------------------------------------------------------------------------
// per-group parameter math.  R() rounds a float through the weight's storage type so that fp16/bf16
// weights see the same per-op rounding as torch CPU ops on Half/BFloat16 tensors (opmath = float).
// ------------------------------------------------------------------------------------------------
struct QRange {
  float minq, maxq;
};

__host__ __device__ __forceinline__ QRange qrange(int bits, bool sym) {
  QRange r;
  if (sym) {
    r.maxq = (float)((1 << (bits - 1)) - 1);
    r.minq = -(float)(1 << (bits - 1));
    if (bits == 1) {  // utility.py:223-225
      r.maxq = 1.f;
      r.minq = 0.f;
    }
  } else {
    r.maxq = (float)((1 << bits) - 1);
    r.minq = 0.f;
  }
  return r;
}

template <typename T>
__device__ __forceinline__ void rtn_group_params(float mn, float mx, int bits, bool sym, bool full_range,
                                                 float quantile, float& scale, float& zp) {
  using E = ElemTraits<T>;
  if (sym) {  // utility.py:226-240
    const QRange r = qrange(bits, true);
    const bool flip = fabsf(mx) > fabsf(mn);
    float amax = E::round(__fmul_rn(fmaxf(fabsf(mx), fabsf(mn)), quantile));
    if (amax == 0.f) amax = 1.f;
    if (full_range) {
      scale = E::round(__fdiv_rn(amax, -r.minq));
      if (flip) scale = -scale;
    } else {
      scale = E::round(__fdiv_rn(amax, r.maxq));
    }
    zp = 0.f;
  } else {  // utility.py:176-187 ; min/max against float32 zeros promote to fp32
    const QRange r = qrange(bits, false);
    float lo = __fmul_rn(fminf(mn, 0.f), quantile);
    float hi = __fmul_rn(fmaxf(mx, 0.f), quantile);
    if (lo == 0.f && hi == 0.f) {
      lo = -1.f;
      hi = 1.f;
    }
    scale = __fdiv_rn(__fsub_rn(hi, lo), r.maxq);
    zp = rintf(__fdiv_rn(-lo, scale));
  }
}

// integer-valued float code exactly as quant_tensor(return_int=True) leaves it in the tensor
template <typename T>
__device__ __forceinline__ float rtn_code(float w, float scale, float zp, bool sym, QRange r) {
  using E = ElemTraits<T>;
  float v = rintf(E::round(__fdiv_rn(w, scale)));
  if (!sym) v = E::round(__fadd_rn(v, zp));
  return fminf(fmaxf(v, r.minq), r.maxq);
}


}  // namespace b200woq
