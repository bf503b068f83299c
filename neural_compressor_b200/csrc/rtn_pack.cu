// rtn_pack.cu -- K4: RTN group quantisation, bit packing, unpacking, de-quantisation.
//
// Reference semantics (neural_compressor/torch/algorithms/weight_only/):
//   utility.py:162-244  qdq_weight_asym / qdq_weight_sym   (per-group scale / zero-point / rounding)
//   utility.py:272-376  quant_tensor                        (grouping incl. ragged tail)
//   modules.py:321-375  INCWeightOnlyLinear.pack            (optimum layout, +2^(b-1), zp-1)
//   modules.py:377-443  unpack / recover
// All kernels are HBM-bound byte movers: W is read coalesced along K, the packed words are written
// coalesced along N, the [n][k] -> [k/n_pack][n] transpose goes through shared memory.
#include "common.cuh"
#include "rtn_math.cuh"

namespace b200woq {

// ------------------------------------------------------------------------------------------------
// B1: one warp per (row, group): min/max -> scale, zp
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) rtn_params_kernel(const T* __restrict__ W, int64_t N, int64_t K, int g,
                                                        int64_t G, int bits, int sym, int full_range,
                                                        float quantile, const float* __restrict__ col_scale,
                                                        float* __restrict__ scale, float* __restrict__ zp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t task = warp; task < N * G; task += nwarps) {
    const int64_t n = task / G, gi = task % G;
    const int64_t k0 = gi * g;
    const int64_t k1 = (k0 + g < K) ? k0 + g : K;
    const T* row = W + n * K;
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t k = k0 + lane; k < k1; k += 32) {
      float v = ElemTraits<T>::load(row + k);
      if (col_scale) v = __fmul_rn(v, col_scale[k]);
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) {
      float s, z;
      if (col_scale)
        rtn_group_params<float>(mn, mx, bits, sym, full_range, quantile, s, z);
      else
        rtn_group_params<T>(mn, mx, bits, sym, full_range, quantile, s, z);
      scale[task] = s;
      if (zp) zp[task] = z;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// B2: quantise (or read codes) + pack.  Tile = 32 rows x (32 words * n_pack) columns.
//   SRC = 0: codes come from W + scale/zp arrays ; SRC = 1: codes come from a uint8 [N,K] array
// ------------------------------------------------------------------------------------------------
constexpr int kTileN = 32;
constexpr int kTileWords = 32;

template <typename T, int SRC>
__global__ void __launch_bounds__(256)
    quant_pack_kernel(const T* __restrict__ W, const uint8_t* __restrict__ codes_in, int64_t N, int64_t K, int g,
                      int64_t G, int bits, int sym, const float* __restrict__ scale, const float* __restrict__ zp,
                      int32_t* __restrict__ qweight, uint8_t* __restrict__ codes_out) {
  extern __shared__ uint8_t smem[];
  const int n_pack = 32 / bits;
  const int tile_k = kTileWords * n_pack;
  const int ld = tile_k + 4;  // bytes per smem row; (ld/4) odd for the column reads below
  const int64_t n0 = (int64_t)blockIdx.y * kTileN;
  const int64_t kw0 = (int64_t)blockIdx.x * kTileWords;
  const int64_t k0 = kw0 * n_pack;
  const int64_t Kw = (K + n_pack - 1) / n_pack;
  const QRange r = qrange(bits, sym != 0);
  const int bias = sym ? (1 << (bits - 1)) : 0;
  const uint32_t mask = (1u << bits) - 1u;

  // phase 1: coalesced read along K, one row at a time per 256-thread sweep
  for (int rn = 0; rn < kTileN; ++rn) {
    const int64_t n = n0 + rn;
    for (int kk = threadIdx.x; kk < tile_k; kk += blockDim.x) {
      const int64_t k = k0 + kk;
      uint8_t c = 0;
      if (n < N && k < K) {
        if (SRC == 0) {
          const int64_t gi = k / g;
          const float s = scale[n * G + gi];
          const float z = sym ? 0.f : zp[n * G + gi];
          const float q = rtn_code<T>(ElemTraits<T>::load(W + n * K + k), s, z, sym != 0, r);
          c = (uint8_t)(((int)q + bias) & mask);
          if (codes_out) codes_out[n * K + k] = c;
        } else {
          c = codes_in[n * K + k] & mask;
        }
      }
      smem[rn * ld + kk] = c;
    }
  }
  __syncthreads();
  // phase 2: word (kw, n) = OR_e code[n][kw*n_pack+e] << bits*e ; lanes sweep n -> coalesced stores
  const int tn = threadIdx.x & 31;
  for (int w = threadIdx.x >> 5; w < kTileWords; w += (blockDim.x >> 5)) {
    const int64_t kw = kw0 + w, n = n0 + tn;
    if (kw < Kw && n < N) {
      uint32_t word = 0;
      const uint8_t* p = smem + tn * ld + w * n_pack;
      for (int e = 0; e < n_pack; ++e) word |= (uint32_t)p[e] << (bits * e);
      qweight[kw * N + n] = (int32_t)word;
    }
  }
}

// scales fp32 [N,G] -> fp16 [G,N] ; zp -> qzeros [G, ceil(N/n_pack)] with (zp-1)&mask  (modules.py:345-371)
__global__ void pack_params_kernel(const float* __restrict__ scale, const float* __restrict__ zp, int64_t N,
                                   int64_t G, int bits, __half* __restrict__ scales16,
                                   int32_t* __restrict__ qzeros) {
  const int n_pack = 32 / bits;
  const int64_t Nw = (N + n_pack - 1) / n_pack;
  const uint32_t mask = (1u << bits) - 1u;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (scales16)
    for (int64_t i = tid; i < N * G; i += stride) {
      const int64_t gi = i / N, n = i % N;
      scales16[i] = __float2half_rn(scale[n * G + gi]);
    }
  if (qzeros)
    for (int64_t i = tid; i < G * Nw; i += stride) {
      const int64_t gi = i / Nw, j = i % Nw;
      uint32_t word = 0;
      for (int e = 0; e < n_pack; ++e) {
        const int64_t n = j * n_pack + e;
        if (n < N) {
          const int z = zp ? (int)zp[n * G + gi] : (1 << (bits - 1));
          word |= ((uint32_t)(z - 1) & mask) << (bits * e);
        }
      }
      qzeros[i] = (int32_t)word;
    }
}

// fake quantisation (quant_tensor(return_int=False)); optional AWQ column scale
template <typename T>
__global__ void __launch_bounds__(256)
    fake_quant_kernel(const T* __restrict__ W, int64_t N, int64_t K, int g, int64_t G, int bits, int sym,
                      const float* __restrict__ scale, const float* __restrict__ zp,
                      const float* __restrict__ col_scale, T* __restrict__ out) {
  const QRange r = qrange(bits, sym != 0);
  const int64_t total = N * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / K, k = i % K;
    const int64_t gi = k / g;
    const float s = scale[n * G + gi];
    const float z = sym ? 0.f : zp[n * G + gi];
    float w = ElemTraits<T>::load(W + i);
    float res;
    if (col_scale) {  // awq.py:326-335 runs in fp32 (W.mul(fp32 scales) promotes)
      const float cs = col_scale[k];
      float q = rtn_code<float>(__fmul_rn(w, cs), s, z, sym != 0, r);
      if (!sym) q = __fsub_rn(q, z);
      res = __fdiv_rn(__fmul_rn(q, s), cs);
    } else {
      float q = rtn_code<T>(w, s, z, sym != 0, r);
      if (!sym) q = ElemTraits<T>::round(__fsub_rn(q, z));
      res = __fmul_rn(q, s);
    }
    ElemTraits<T>::store(out + i, res);
  }
}

// unpack stored codes / zero-points  (modules.py:377-411)
__global__ void unpack_kernel(const int32_t* __restrict__ qweight, const int32_t* __restrict__ qzeros, int64_t N,
                              int64_t K, int64_t G, int bits, uint8_t* __restrict__ codes,
                              uint8_t* __restrict__ zps) {
  const int n_pack = 32 / bits;
  const uint32_t mask = (1u << bits) - 1u;
  const int64_t Nw = (N + n_pack - 1) / n_pack;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (codes)
    for (int64_t i = tid; i < N * K; i += stride) {
      const int64_t n = i / K, k = i % K;
      const uint32_t w = (uint32_t)qweight[(k / n_pack) * N + n];
      codes[i] = (uint8_t)((w >> (bits * (k % n_pack))) & mask);
    }
  if (zps)
    for (int64_t i = tid; i < N * G; i += stride) {
      const int64_t n = i / G, gi = i % G;
      const uint32_t w = (uint32_t)qzeros[gi * Nw + n / n_pack];
      uint32_t z = ((w >> (bits * (n % n_pack))) & mask) + 1u;
      if (z > mask) z = 0;  // modules.py:409-410
      zps[i] = (uint8_t)z;
    }
}

// recover(): W_fp16[n,k] = fp16(int8(q - zp) * scale_fp16). Tile transpose through smem:
// read words coalesced along N, write halves coalesced along K.
__global__ void __launch_bounds__(256)
    dequantize_kernel(const int32_t* __restrict__ qweight, const int32_t* __restrict__ qzeros,
                      const __half* __restrict__ scales, const int32_t* __restrict__ g_idx, int64_t N, int64_t K,
                      int bits, int g, __half* __restrict__ out) {
  extern __shared__ uint8_t smem[];
  const int n_pack = 32 / bits;
  const int tile_k = kTileWords * n_pack;
  const int ld = tile_k + 4;
  const uint32_t mask = (1u << bits) - 1u;
  const int64_t n0 = (int64_t)blockIdx.y * kTileN, kw0 = (int64_t)blockIdx.x * kTileWords;
  const int64_t Kw = (K + n_pack - 1) / n_pack, Nw = (N + n_pack - 1) / n_pack;
  const int tn = threadIdx.x & 31;
  for (int w = threadIdx.x >> 5; w < kTileWords; w += (blockDim.x >> 5)) {
    const int64_t kw = kw0 + w, n = n0 + tn;
    uint32_t word = (kw < Kw && n < N) ? (uint32_t)qweight[kw * N + n] : 0u;
    uint8_t* p = smem + tn * ld + w * n_pack;
    for (int e = 0; e < n_pack; ++e) p[e] = (uint8_t)((word >> (bits * e)) & mask);
  }
  __syncthreads();
  for (int rn = 0; rn < kTileN; ++rn) {
    const int64_t n = n0 + rn;
    if (n >= N) break;
    for (int kk = threadIdx.x; kk < tile_k; kk += blockDim.x) {
      const int64_t k = kw0 * n_pack + kk;
      if (k >= K) continue;
      const int64_t gi = g_idx ? g_idx[k] : k / g;
      const uint32_t zw = (uint32_t)qzeros[gi * Nw + n / n_pack];
      uint32_t z = ((zw >> (bits * (n % n_pack))) & mask) + 1u;
      if (z > mask) z = 0;
      const int8_t d = (int8_t)((int)smem[rn * ld + kk] - (int)z);  // .to(torch.int8) wraps (modules.py:435)
      out[n * K + k] = __hmul(__int2half_rn((int)d), scales[gi * N + n]);
    }
  }
}

static int grid_for(int64_t work_items, int threads) {
  int64_t b = ceil_div(work_items, threads);
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace b200woq

using namespace b200woq;

#define DISPATCH_DTYPE(dt, ...)                         \
  switch (dt) {                                         \
    case B200WOQ_F32: {                                 \
      using T = float;                                  \
      __VA_ARGS__;                                      \
    } break;                                            \
    case B200WOQ_F16: {                                 \
      using T = __half;                                 \
      __VA_ARGS__;                                      \
    } break;                                            \
    case B200WOQ_BF16: {                                \
      using T = __nv_bfloat16;                          \
      __VA_ARGS__;                                      \
    } break;                                            \
    default:                                            \
      set_error("unsupported dtype %d", dt);            \
      return B200WOQ_EINVAL;                            \
  }

static int check_bits(int bits) {
  if (bits < 1 || bits > 8) {
    set_error("bits must be in [1,8], got %d", bits);
    return B200WOQ_EINVAL;
  }
  return 0;
}

static int rtn_params_impl(const void* W, int w_dtype, int64_t N, int64_t K, int bits, int group_size, int sym,
                           int full_range, float quantile, const float* col_scale, float* scale, float* zp,
                           cudaStream_t st) {
  const int g = eff_group(K, group_size);
  const int64_t G = ceil_div(K, g);
  const int blocks = grid_for(N * G * 32, 256);
  DISPATCH_DTYPE(w_dtype, rtn_params_kernel<T><<<blocks, 256, 0, st>>>((const T*)W, N, K, g, G, bits, sym, full_range,
                                                                       quantile, col_scale, scale, sym ? nullptr : zp));
  WOQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200woq_rtn_params(const void* W, int w_dtype, int64_t N, int64_t K, int bits, int group_size, int sym,
                                  int full_range, float quantile, float* scale, float* zp, void* stream) {
  if (check_bits(bits)) return B200WOQ_EINVAL;
  WOQ_CHECK_ARG(W && scale && N > 0 && K > 0, "rtn_params: null pointer or empty shape");
  WOQ_CHECK_ARG(sym || zp, "rtn_params: zp output required for asym");
  return rtn_params_impl(W, w_dtype, N, K, bits, group_size, sym, full_range, quantile, nullptr, scale, zp,
                         (cudaStream_t)stream);
}

extern "C" int b200woq_rtn_quant_pack(const void* W, int w_dtype, int64_t N, int64_t K, int bits, int group_size,
                                      int sym, const float* scale, const float* zp, int32_t* qweight,
                                      uint8_t* codes_out, void* stream) {
  if (check_bits(bits)) return B200WOQ_EINVAL;
  WOQ_CHECK_ARG(W && scale && qweight && N > 0 && K > 0, "rtn_quant_pack: null pointer or empty shape");
  WOQ_CHECK_ARG(sym || zp, "rtn_quant_pack: zp required for asym");
  const int g = eff_group(K, group_size);
  const int64_t G = ceil_div(K, g);
  const int n_pack = 32 / bits;
  const int64_t Kw = ceil_div(K, n_pack);
  dim3 grid((unsigned)ceil_div(Kw, kTileWords), (unsigned)ceil_div(N, kTileN));
  const size_t smem = (size_t)kTileN * (kTileWords * n_pack + 4);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DTYPE(w_dtype, (quant_pack_kernel<T, 0><<<grid, 256, smem, st>>>((const T*)W, nullptr, N, K, g, G, bits, sym,
                                                                            scale, zp, qweight, codes_out)));
  WOQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200woq_pack_codes(const uint8_t* codes, int64_t N, int64_t K, int bits, int32_t* qweight,
                                  void* stream) {
  if (check_bits(bits)) return B200WOQ_EINVAL;
  WOQ_CHECK_ARG(codes && qweight && N > 0 && K > 0, "pack_codes: null pointer or empty shape");
  const int n_pack = 32 / bits;
  const int64_t Kw = ceil_div(K, n_pack);
  dim3 grid((unsigned)ceil_div(Kw, kTileWords), (unsigned)ceil_div(N, kTileN));
  const size_t smem = (size_t)kTileN * (kTileWords * n_pack + 4);
  quant_pack_kernel<float, 1><<<grid, 256, smem, (cudaStream_t)stream>>>(nullptr, codes, N, K, (int)K, 1, bits, 0,
                                                                         nullptr, nullptr, qweight, nullptr);
  WOQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200woq_pack_params(const float* scale, const float* zp, int64_t N, int64_t G, int bits,
                                   void* scales16_out, int32_t* qzeros_out, void* stream) {
  if (check_bits(bits)) return B200WOQ_EINVAL;
  WOQ_CHECK_ARG(N > 0 && G > 0 && (scales16_out == nullptr || scale != nullptr), "pack_params: bad arguments");
  pack_params_kernel<<<grid_for(N * G, 256), 256, 0, (cudaStream_t)stream>>>(scale, zp, N, G, bits,
                                                                             (__half*)scales16_out, qzeros_out);
  WOQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200woq_rtn_fake_quant(const void* W, int w_dtype, int64_t N, int64_t K, int bits, int group_size,
                                      int sym, int full_range, float quantile, const float* col_scale, void* out,
                                      void* stream) {
  if (check_bits(bits)) return B200WOQ_EINVAL;
  WOQ_CHECK_ARG(W && out && N > 0 && K > 0, "rtn_fake_quant: null pointer or empty shape");
  const int g = eff_group(K, group_size);
  const int64_t G = ceil_div(K, g);
  cudaStream_t st = (cudaStream_t)stream;
  // group parameters live in a stream-ordered temporary (cudaMallocAsync pool, no sync)
  float* params = nullptr;
  WOQ_CUDA(cudaMallocAsync((void**)&params, sizeof(float) * 2 * N * G, st));
  int rc = rtn_params_impl(W, w_dtype, N, K, bits, group_size, sym, full_range, quantile, col_scale, params,
                           params + N * G, st);
  if (rc == 0) {
    DISPATCH_DTYPE(w_dtype, fake_quant_kernel<T><<<grid_for(N * K, 256), 256, 0, st>>>(
                                (const T*)W, N, K, g, G, bits, sym, params, params + N * G, col_scale, (T*)out));
    count_launch(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("fake_quant launch failed: %s", cudaGetErrorString(e));
      rc = B200WOQ_ECUDA;
    }
  }
  cudaFreeAsync(params, st);
  return rc;
}

extern "C" int b200woq_unpack(const int32_t* qweight, const int32_t* qzeros, int64_t N, int64_t K, int64_t G,
                              int bits, uint8_t* codes_out, uint8_t* zp_out, void* stream) {
  if (check_bits(bits)) return B200WOQ_EINVAL;
  WOQ_CHECK_ARG((codes_out == nullptr || qweight) && (zp_out == nullptr || qzeros), "unpack: missing input");
  unpack_kernel<<<grid_for(N * K, 256), 256, 0, (cudaStream_t)stream>>>(qweight, qzeros, N, K, G, bits, codes_out,
                                                                        zp_out);
  WOQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200woq_dequantize(const int32_t* qweight, const int32_t* qzeros, const void* scales16,
                                  const int32_t* g_idx, int64_t N, int64_t K, int bits, int group_size,
                                  void* w_fp16_out, void* stream) {
  if (check_bits(bits)) return B200WOQ_EINVAL;
  WOQ_CHECK_ARG(qweight && qzeros && scales16 && w_fp16_out, "dequantize: null pointer");
  const int g = eff_group(K, group_size);
  const int n_pack = 32 / bits;
  const int64_t Kw = ceil_div(K, n_pack);
  dim3 grid((unsigned)ceil_div(Kw, kTileWords), (unsigned)ceil_div(N, kTileN));
  const size_t smem = (size_t)kTileN * (kTileWords * n_pack + 4);
  dequantize_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(qweight, qzeros, (const __half*)scales16, g_idx, N, K,
                                                               bits, g, (__half*)w_fp16_out);
  WOQ_LAUNCH_CHECK();
  return 0;
}
