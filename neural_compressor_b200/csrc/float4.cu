// float4.cu -- RTN for the 4-bit table data types (nf4 / fp4 / fp4_e2m1): group quantisation, row-major packing and
// de-quantisation.
//
// Reference (neural_compressor/torch/algorithms/weight_only/):
//   utility.py:121-160   quantize_4bit            (per-group absmax scale, nearest level by mid points)
//   utility.py:272-376   quant_tensor             (grouping incl. the ragged tail group)
//   modules.py:214-222   INCWeightOnlyLinear      (table dtypes force use_optimum_format=False: qweight int32 [N, ceil(K/8)]
//                                                   packed along K, scales fp32 [N, G], no qzeros)
//   modules.py:445-466   pack_tensor_with_torch   (field e of a word = (v & 0xF) << 4e)
//   modules.py:377-443   unpack / recover         (sign-extended nibble -> level via INT/FLOAT_MAPPING, * scale, fp32)
// All three kernels are HBM-bound streaming passes: W is read twice by the same warp (the second read hits L1/L2), codes
// and words are written once, coalesced along K.
#include "common.cuh"
#include "f4_math.cuh"

namespace b200woq {

template <typename T>
struct StorageRound {
  static __host__ __device__ __forceinline__ float round(float v) {
#ifdef __CUDA_ARCH__
    return ElemTraits<T>::round(v);
#else
    return v;
#endif
  }
};

// one warp per (row, group): absmax -> scale -> level index per element
template <typename T>
__global__ void __launch_bounds__(256)
    f4_quantize_kernel(const T* __restrict__ W, int64_t N, int64_t K, int g, int64_t G, b200woq_f4_table table,
                       float quantile, int8_t* __restrict__ codes, float* __restrict__ scale_out, T* __restrict__ fake) {
  using R = StorageRound<T>;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t task = warp; task < N * G; task += nwarps) {
    const int64_t n = task / G, gi = task % G;
    const int64_t k0 = gi * g;
    const int64_t k1 = (k0 + g < K) ? k0 + g : K;
    const T* row = W + n * K;
    float amax = 0.f;
    for (int64_t k = k0 + lane; k < k1; k += 32) amax = fmaxf(amax, fabsf(ElemTraits<T>::load(row + k)));
    amax = warp_max(amax);
    const float s = f4_group_scale<R>(amax, quantile, table.max_level);
    if (lane == 0 && scale_out) scale_out[task] = s;
    for (int64_t k = k0 + lane; k < k1; k += 32) {
      const int idx = f4_select<R>(ElemTraits<T>::load(row + k), s, table);
      if (codes) codes[n * K + k] = (int8_t)f4_code(idx, table);
      if (fake) ElemTraits<T>::store(fake + n * K + k, f4_fake<R>(idx, s, table));
    }
  }
}

// codes int8 [N,K] (two's complement fields) -> words int32 [N, ceil(K/n_pack)], one thread per word
__global__ void __launch_bounds__(256)
    pack_rows_kernel(const int8_t* __restrict__ codes, int64_t N, int64_t K, int bits, int32_t* __restrict__ qweight) {
  const int n_pack = 32 / bits;
  const int64_t Kw = (K + n_pack - 1) / n_pack;
  const uint32_t mask = (1u << bits) - 1u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N * Kw; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / Kw, j = i % Kw;
    uint32_t word = 0;
    for (int e = 0; e < n_pack; ++e) {
      const int64_t k = j * n_pack + e;
      if (k < K) word |= ((uint32_t)(int32_t)codes[n * K + k] & mask) << (bits * e);
    }
    qweight[i] = (int32_t)word;
  }
}

struct NibbleLevels {
  float v[16];
};

// words [N, ceil(K/8)] + scales fp32 [N, G] -> fp32 [N, K]; one thread per word, 8 consecutive outputs
__global__ void __launch_bounds__(256)
    f4_dequantize_kernel(const int32_t* __restrict__ qweight, const float* __restrict__ scales, NibbleLevels lv,
                         int64_t N, int64_t K, int g, int64_t G, float* __restrict__ out) {
  const int64_t Kw = (K + 7) / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N * Kw; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / Kw, j = i % Kw;
    const uint32_t word = (uint32_t)qweight[i];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int64_t k = j * 8 + e;
      if (k < K) out[n * K + k] = f4_recover(word >> (4 * e), scales[n * G + k / g], lv.v);
    }
  }
}

static int f4_grid(int64_t work_items, int threads) {
  int64_t b = ceil_div(work_items, threads);
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int check_table(const b200woq_f4_table* t) {
  if (!t || t->n < 2 || t->n > 16) {
    set_error("f4 table: n must be in [2,16]");
    return B200WOQ_EINVAL;
  }
  for (int i = 0; i + 1 < t->n; ++i)
    if (!(t->level[i] < t->level[i + 1]) || !(t->level[i] <= t->mid[i] && t->mid[i] <= t->level[i + 1])) {
      set_error("f4 table: levels must ascend with mid points in between (entry %d)", i);
      return B200WOQ_EINVAL;
    }
  return 0;
}

}  // namespace b200woq

using namespace b200woq;

extern "C" int b200woq_f4_quantize(const void* W, int w_dtype, int64_t N, int64_t K, int group_size,
                                   const b200woq_f4_table* host_table, float quantile, int8_t* codes_out,
                                   float* scale_out, void* fake_out, void* stream) {
  WOQ_CHECK_ARG(W && N > 0 && K > 0, "f4_quantize: null pointer or empty shape");
  WOQ_CHECK_ARG(codes_out || scale_out || fake_out, "f4_quantize: no output requested");
  if (check_table(host_table)) return B200WOQ_EINVAL;
  const int g = eff_group(K, group_size);
  const int64_t G = ceil_div(K, g);
  const int blocks = f4_grid(N * G * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  const b200woq_f4_table table = *host_table;
  switch (w_dtype) {
    case B200WOQ_F32:
      f4_quantize_kernel<float><<<blocks, 256, 0, st>>>((const float*)W, N, K, g, G, table, quantile, codes_out,
                                                        scale_out, (float*)fake_out);
      break;
    case B200WOQ_F16:
      f4_quantize_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)W, N, K, g, G, table, quantile, codes_out,
                                                         scale_out, (__half*)fake_out);
      break;
    case B200WOQ_BF16:
      f4_quantize_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)W, N, K, g, G, table, quantile,
                                                                codes_out, scale_out, (__nv_bfloat16*)fake_out);
      break;
    default:
      set_error("unsupported dtype %d", w_dtype);
      return B200WOQ_EINVAL;
  }
  WOQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200woq_pack_rows(const int8_t* codes, int64_t N, int64_t K, int bits, int32_t* qweight_out,
                                 void* stream) {
  WOQ_CHECK_ARG(codes && qweight_out && N > 0 && K > 0, "pack_rows: null pointer or empty shape");
  WOQ_CHECK_ARG(bits >= 1 && bits <= 8, "pack_rows: bits must be in [1,8], got %d", bits);
  const int n_pack = 32 / bits;
  pack_rows_kernel<<<f4_grid(N * ceil_div(K, n_pack), 256), 256, 0, (cudaStream_t)stream>>>(codes, N, K, bits,
                                                                                           qweight_out);
  WOQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200woq_f4_dequantize(const int32_t* qweight, const float* scales, const float* host_nibble_levels,
                                     int64_t N, int64_t K, int group_size, float* w_out, void* stream) {
  WOQ_CHECK_ARG(qweight && scales && host_nibble_levels && w_out && N > 0 && K > 0,
                "f4_dequantize: null pointer or empty shape");
  const int g = eff_group(K, group_size);
  const int64_t G = ceil_div(K, g);
  NibbleLevels lv;
  for (int i = 0; i < 16; ++i) lv.v[i] = host_nibble_levels[i];
  f4_dequantize_kernel<<<f4_grid(N * ceil_div(K, 8), 256), 256, 0, (cudaStream_t)stream>>>(qweight, scales, lv, N, K,
                                                                                          g, G, w_out);
  WOQ_LAUNCH_CHECK();
  return 0;
}
