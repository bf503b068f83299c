// Host build of the product's f4_math.cuh (test infrastructure for tests/test_f4_math_cpu.py): the very same
// per-element functions the CUDA kernels in csrc/float4.cu call, wrapped in the kernels' grouping loop, so the math can
// be checked against tensors written by the live reference without a GPU.  Not part of libb200woq.so.
#define B200WOQ_F4_HOST
#include "../../neural_compressor_b200/csrc/f4_math.cuh"

#include <cmath>
#include <cstring>

namespace {
struct RoundF32 {
  static float round(float v) { return v; }
};
struct RoundF16 {
  static float round(float v) { return (float)(_Float16)v; }
};
struct RoundBF16 {
  static float round(float v) {
    uint32_t u;
    std::memcpy(&u, &v, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return v;  // NaN
    u += 0x7fffu + ((u >> 16) & 1u);                // round to nearest even
    u &= 0xffff0000u;
    std::memcpy(&v, &u, 4);
    return v;
  }
};

template <typename R>
void quantize(const float* W, int64_t N, int64_t K, int g, const b200woq_f4_table& t, float quantile, int8_t* codes,
              float* scale, float* fake) {
  const int64_t G = (K + g - 1) / g;
  for (int64_t n = 0; n < N; ++n)
    for (int64_t gi = 0; gi < G; ++gi) {
      const int64_t k0 = gi * g, k1 = (k0 + g < K) ? k0 + g : K;
      float amax = 0.f;
      for (int64_t k = k0; k < k1; ++k) amax = std::fmax(amax, std::fabs(W[n * K + k]));
      const float s = b200woq::f4_group_scale<R>(amax, quantile, t.max_level);
      scale[n * G + gi] = s;
      for (int64_t k = k0; k < k1; ++k) {
        const int idx = b200woq::f4_select<R>(W[n * K + k], s, t);
        codes[n * K + k] = (int8_t)b200woq::f4_code(idx, t);
        fake[n * K + k] = b200woq::f4_fake<R>(idx, s, t);
      }
    }
}
}  // namespace

extern "C" void f4_quantize_host(const float* W, int64_t N, int64_t K, int g, const b200woq_f4_table* t, float quantile,
                                 int rounding, int8_t* codes, float* scale, float* fake) {
  if (rounding == 1)
    quantize<RoundF16>(W, N, K, g, *t, quantile, codes, scale, fake);
  else if (rounding == 2)
    quantize<RoundBF16>(W, N, K, g, *t, quantile, codes, scale, fake);
  else
    quantize<RoundF32>(W, N, K, g, *t, quantile, codes, scale, fake);
}

extern "C" void f4_recover_host(const int32_t* qweight, const float* scales, const float* nibble_levels, int64_t N,
                                int64_t K, int g, float* out) {
  const int64_t Kw = (K + 7) / 8, G = (K + g - 1) / g;
  for (int64_t n = 0; n < N; ++n)
    for (int64_t k = 0; k < K; ++k) {
      const uint32_t word = (uint32_t)qweight[n * Kw + k / 8];
      out[n * K + k] = b200woq::f4_recover(word >> (4 * (k % 8)), scales[n * G + k / g], nibble_levels);
    }
}
