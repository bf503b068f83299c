// w8a8.cu -- K7: SmoothQuant W8A8 static INT8 linear (BASELINE configs[3]) on the 5th-generation tensor cores.
//
// Reference semantics (neural_compressor/torch/algorithms/smooth_quant/utility.py):
//   SQLinearWrapper (:2559-2662)      x' = x * input_scale (= x / s, the smoothing scale folded as a mul), static per-tensor
//                                     asym uint8 activation qparams from the calibrated range (:2607-2631)
//   quant_dequant_x_v1 (:726-755)     q_x = clamp(round(x'/scale_x + zp_x), 0, 255)
//   quant_dequant_w_v1 (:652-690)     per-out-channel sym: scale_w[n] = max|W'_n| / 127.5, q_w = clamp(round(W'/scale_w), -128, 127)
//   cal_scale / _scale_layer_weight   W' = W * s (per input channel)
// The INT8 GEMM itself lives in IPEX/oneDNN (not in the reference tree, SURVEY §8c: parity UNPINNED); it is
//       y[m,n] = (sum_k q_x[m,k] q_w[n,k]  -  zp_x * sum_k q_w[n,k]) * scale_x * scale_w[n] + bias[n]
// which is what this file computes, with the integer part exact.
//
//   sq_smooth_quant_weight_kernel   one CTA per out-channel row: W*s -> absmax -> q_w (int8), scale_w, row sums
//   sq_quantize_act_kernel          x -> u8 codes, 16 elements per thread (HBM bound; IEEE division like torch)
//   w8a8_gemm_kernel                tcgen05.mma.kind::i8: A = q_w tile (128 out-channels x 128 k, K-major SWIZZLE_128B, TMA),
//                                   B = q_x tile (NT tokens x 128 k), D = s32 [128 x NT] in TMEM; warp 0 = TMA producer,
//                                   warp 1 = MMA issuer, warps 2-5 = epilogue (tcgen05.ld -> dequant -> coalesced stores).
//                                   Out-channels are MMA-M so decode batches (M = 1..64 tokens) waste no tensor rows; K is
//                                   split over gridDim.z when the tile grid is smaller than the chip: partial s32 sums are
//                                   added with integer atomics (exact, order-free => deterministic) into a zeroed workspace
//                                   and the last CTA of a tile runs the epilogue and re-zeroes it.
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include <algorithm>

namespace b200woq {
namespace w8a8 {

constexpr int TM = 128;          // out-channels per CTA (MMA M)
constexpr int BKB = 128;         // bytes (= int8 elements) of K per pipeline stage: one 128-byte swizzle row
constexpr int STAGES = 4;
constexpr int NUM_EPI_THREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor; canonical layout
// ((8,n),2):((8,SBO),1) in 16-byte units): rows are 128 B apart inside an 8-row atom, atoms SBO = 1024 B apart.
__device__ __forceinline__ uint64_t make_desc_k(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) /*LBO (unused for swizzled K-major)*/ |
         ((uint64_t)(1024 >> 4) << 32) /*SBO*/ | (1ull << 46) /*version*/ | (2ull << 61) /*SWIZZLE_128B*/;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct GemmParams {
  int M, N, K;            // tokens, out-channels, in-channels
  int NT;                 // token tile (MMA N): multiple of 16, <= 256
  int kb_per_split;       // k-blocks (of 128) per z-slice
  const int32_t* wsum;    // [N] sum_k q_w[n,k]
  const float* w_scale;   // [N]
  const float* x_scale;   // [1]
  const float* x_zp;      // [1] (integer-valued float, like the reference keeps it)
  const void* bias;       // [N] or null
  int bias_dtype;
  void* y;                // [M, N]
  int y_dtype;
  int32_t* acc_ws;        // [tiles][splits][NT][128] s32 partial tiles (split-K only; plain scratch)
  int* counters;          // [tiles], zero on entry and on exit
};

__device__ __forceinline__ float load_as_float(const void* p, int dtype, int64_t i) {
  if (dtype == B200WOQ_F32) return ((const float*)p)[i];
  if (dtype == B200WOQ_F16) return __half2float(((const __half*)p)[i]);
  return __bfloat162float(((const __nv_bfloat16*)p)[i]);
}
__device__ __forceinline__ void store_from_float(void* p, int dtype, int64_t i, float v) {
  if (dtype == B200WOQ_F32)
    ((float*)p)[i] = v;
  else if (dtype == B200WOQ_F16)
    ((__half*)p)[i] = __float2half_rn(v);
  else
    ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
}

// grid (N/128 out-channel tiles, ceil(M/NT) token tiles, splits); block 192
template <int NT_MAX>
__global__ void __launch_bounds__(192, 1)
    w8a8_gemm_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const GemmParams p,
                     uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int NT = p.NT;
  const uint32_t a_bytes = TM * BKB, b_bytes = (uint32_t)NT * BKB, stage_bytes = a_bytes + b_bytes;
  const uint32_t bars = base + STAGES * (a_bytes + (uint32_t)NT_MAX * BKB);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t accum_full = bars + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  int* flag_ptr = reinterpret_cast<int*>(smem_raw + (tmem_slot + 8 - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TM, m0 = blockIdx.y * NT;
  const int nkb_total = p.K / BKB;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int nkb = min(p.kb_per_split, nkb_total - kb0);
  const uint32_t tmem_cols = NT_MAX <= 32 ? 32 : NT_MAX <= 64 ? 64 : NT_MAX <= 128 ? 128 : 256;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), stage_bytes);
        const uint32_t sa = base + s * (a_bytes + (uint32_t)NT_MAX * BKB);
        const int k0 = (kb0 + i) * BKB;
        tma_load_2d(sa, &map_w, full_bar(s), k0, n0);              // 128 out-channel rows x 128 B
        tma_load_2d(sa + a_bytes, &map_x, full_bar(s), k0, m0);    // NT token rows x 128 B (rows past M are zero-filled)
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = base + s * (a_bytes + (uint32_t)NT_MAX * BKB);
#pragma unroll
        for (int k4 = 0; k4 < BKB / 32; ++k4) {   // K = 32 int8 per instruction: advance 32 B inside the swizzle row
          umma_i8(tmem_base, make_desc_k(sa + k4 * 32), make_desc_k(sa + a_bytes + k4 * 32), idesc, (i | k4) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_full);
    }
  } else {
    // epilogue warps: warp w owns TMEM lanes [32*(w%4), +32) = out-channels n0 + 32*(w%4) + lane
    const int q = warp & 3;
    const int n = n0 + q * 32 + lane;
    mbar_wait(accum_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int splits = gridDim.z;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int nl = q * 32 + lane;
    const float sx = *p.x_scale, zp = *p.x_zp;
    const float sw = (n < p.N) ? p.w_scale[n] * sx : 0.f;
    const int32_t zsum = (n < p.N) ? (int32_t)zp * p.wsum[n] : 0;
    const float b = (p.bias && n < p.N) ? load_as_float(p.bias, p.bias_dtype, n) : 0.f;
    if (splits == 1) {
      for (int c = 0; c < NT; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
#pragma unroll
        for (int v = 0; v < 16; ++v) {
          const int m = m0 + c + v;
          if (m < p.M && n < p.N) store_from_float(p.y, p.y_dtype, (int64_t)m * p.N + n, fmaf((float)((int32_t)r[v] - zsum), sw, b));
        }
      }
    } else {
      // split-K: s32 partial tiles [tile][split][NT][128] (coalesced over out-channels); the LAST CTA of the tile
      // (atomic counter, self-resetting) sums them -- integer adds, exact in any order
      int32_t* mine = p.acc_ws + ((size_t)(tile * splits + blockIdx.z) * NT) * TM + nl;
      for (int c = 0; c < NT; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
#pragma unroll
        for (int v = 0; v < 16; ++v)
          if (m0 + c + v < p.M) mine[(size_t)(c + v) * TM] = (int32_t)r[v];
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) *flag_ptr = atomicAdd(p.counters + tile, 1);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*flag_ptr == splits - 1) {
        __threadfence();
        const int32_t* t0 = p.acc_ws + ((size_t)(tile * splits) * NT) * TM + nl;
        const int mlim = min(NT, p.M - m0);
        for (int m = 0; m < mlim; ++m) {
          int32_t acc = 0;
          for (int s2 = 0; s2 < splits; ++s2) acc += __ldcg(t0 + ((size_t)s2 * NT + m) * TM);
          if (n < p.N) store_from_float(p.y, p.y_dtype, (int64_t)(m0 + m) * p.N + n, fmaf((float)(acc - zsum), sw, b));
        }
        if (threadIdx.x == 64) p.counters[tile] = 0;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- activation quantisation: q = clamp(rint(x * in_scale / sx + zp), 0, 255), written K-padded to a multiple of 128
__global__ void sq_quantize_act_kernel(const void* __restrict__ x, int x_dtype, int64_t M, int64_t K, int64_t Kp,
                                       const float* __restrict__ in_scale, const float* __restrict__ x_scale,
                                       const float* __restrict__ x_zp, uint8_t* __restrict__ out) {
  const float sx = *x_scale, zp = *x_zp;
  const int64_t chunks_per_row = Kp / 16;
  const int64_t total = M * chunks_per_row;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = idx / chunks_per_row, k0 = (idx - m * chunks_per_row) * 16;
    uint32_t w[4] = {0, 0, 0, 0};
    float v16[16];
    const bool vec = (x_dtype != B200WOQ_F32) && ((K & 7) == 0) && (k0 + 16 <= K);
    if (vec) {  // two 16-byte loads of 8 halves each
      const uint4 a = *reinterpret_cast<const uint4*>((const uint16_t*)x + m * K + k0);
      const uint4 b = *reinterpret_cast<const uint4*>((const uint16_t*)x + m * K + k0 + 8);
      const uint32_t h[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (x_dtype == B200WOQ_F16) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
          v16[2 * i] = f.x;
          v16[2 * i + 1] = f.y;
        } else {
          v16[2 * i] = __uint_as_float(h[i] << 16);
          v16[2 * i + 1] = __uint_as_float(h[i] & 0xffff0000u);
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) v16[e] = (k0 + e < K) ? load_as_float(x, x_dtype, m * K + k0 + e) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int64_t k = k0 + e;
      uint32_t q = 0;
      if (k < K) {
        float v = v16[e];
        if (in_scale) v = __fmul_rn(v, in_scale[k]);
        const float t = rintf(__fadd_rn(__fdiv_rn(v, sx), zp));
        q = (uint32_t)fminf(fmaxf(t, 0.f), 255.f);
      }
      w[e >> 2] |= q << (8 * (e & 3));
    }
    *reinterpret_cast<uint4*>(out + m * Kp + k0) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---- weight side, once per layer: W' = W * s (per input channel), per-row absmax -> scale, int8 codes, row sums
__global__ void __launch_bounds__(256) sq_smooth_quant_weight_kernel(const void* __restrict__ W, int w_dtype, int64_t N, int64_t K,
                                                                    int64_t Kp, const float* __restrict__ smooth,
                                                                    int8_t* __restrict__ qw, float* __restrict__ w_scale,
                                                                    int32_t* __restrict__ wsum) {
  __shared__ float red[8];
  __shared__ int redi[8];
  const int64_t n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float amax = 0.f;
  for (int64_t k = threadIdx.x; k < K; k += 256) {
    float v = load_as_float(W, w_dtype, n * K + k);
    if (smooth) v = __fmul_rn(v, smooth[k]);
    amax = fmaxf(amax, fabsf(v));
  }
  amax = warp_max(amax);
  if (lane == 0) red[warp] = amax;
  __syncthreads();
  if (warp == 0) {
    float v = lane < 8 ? red[lane] : 0.f;
    v = warp_max(v);
    if (lane == 0) red[0] = v;
  }
  __syncthreads();
  // quant_dequant_w_v1: scale = max|w| / ((q_max - q_min) / 2) = max|w| / 127.5, clipped to fp32 eps
  const float scale = fmaxf(__fdiv_rn(red[0], 127.5f), 1.1920928955078125e-07f);
  int sum = 0;
  for (int64_t k = threadIdx.x; k < Kp; k += 256) {
    int q = 0;
    if (k < K) {
      float v = load_as_float(W, w_dtype, n * K + k);
      if (smooth) v = __fmul_rn(v, smooth[k]);
      q = (int)fminf(fmaxf(rintf(__fdiv_rn(v, scale)), -128.f), 127.f);
    }
    qw[n * Kp + k] = (int8_t)q;
    sum += q;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) redi[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w2 = 0; w2 < 8; ++w2) t += redi[w2];
    wsum[n] = t;
    w_scale[n] = scale;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}
static bool encode_u8_2d(CUtensorMap* m, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_outer) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld};
  const cuuint32_t box[2] = {(cuuint32_t)BKB, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace w8a8
}  // namespace b200woq

using namespace b200woq;

static inline int64_t pad_k(int64_t K) { return ceil_div(K, 128) * 128; }

extern "C" int64_t b200woq_w8a8_padded_k(int64_t K) { return pad_k(K); }

extern "C" int b200woq_sq_smooth_quant_weight(const void* W, int w_dtype, int64_t N, int64_t K, const float* smooth,
                                              int8_t* qweight, float* w_scale, int32_t* wsum, void* stream) {
  WOQ_CHECK_ARG(W && qweight && w_scale && wsum && N > 0 && K > 0, "sq_smooth_quant_weight: bad arguments");
  WOQ_CHECK_ARG(w_dtype >= 0 && w_dtype <= 2, "sq_smooth_quant_weight: bad dtype");
  w8a8::sq_smooth_quant_weight_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(W, w_dtype, N, K, pad_k(K), smooth, qweight,
                                                                                    w_scale, wsum);
  WOQ_LAUNCH_CHECK();
  return 0;
}

// workspace: [q_x: M x Kp u8][pad to 256][counters: tiles ints -- zero on entry, left zero][pad][s32 partial tiles]
static void w8a8_plan(int64_t M, int64_t N, int64_t K, int* NT, int* splits, int* kb_per_split, int64_t* tiles) {
  const int64_t Kp = pad_k(K);
  int nt = (int)std::min<int64_t>(256, ceil_div(M, 16) * 16);
  *NT = nt;
  const int64_t t = ceil_div(N, 128) * ceil_div(M, nt);
  *tiles = t;
  const int nkb = (int)(Kp / 128);
  int s = 1;
  const int sms = num_sms();
  if (t < sms) s = (int)std::min<int64_t>(std::min<int64_t>(nkb / 4 > 0 ? nkb / 4 : 1, 16), std::max<int64_t>(1, sms / t));
  int per = (int)ceil_div(nkb, s);
  s = (int)ceil_div(nkb, per);
  *splits = s;
  *kb_per_split = per;
}

extern "C" int64_t b200woq_w8a8_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  int NT, splits, per;
  int64_t tiles;
  w8a8_plan(M, N, K, &NT, &splits, &per, &tiles);
  const int64_t q = ((M * pad_k(K) + 255) / 256) * 256;
  return q + tiles * 4 + 256 + tiles * splits * 128 * NT * 4 + 256;
}

extern "C" int64_t b200woq_w8a8_workspace_zeroed_offset(int64_t M, int64_t K) { return ((M * pad_k(K) + 255) / 256) * 256; }

extern "C" int b200woq_w8a8_linear_forward(const void* x, int x_dtype, int64_t M, int64_t K, int64_t N, const int8_t* qweight,
                                           const float* w_scale, const int32_t* wsum, const float* input_scale,
                                           const float* x_scale, const float* x_zp, const void* bias, int bias_dtype, void* y,
                                           int y_dtype, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace w8a8;
  WOQ_CHECK_ARG(x && qweight && w_scale && wsum && x_scale && x_zp && y && workspace, "w8a8_linear_forward: null pointer");
  WOQ_CHECK_ARG(M > 0 && K > 0 && N > 0, "w8a8_linear_forward: empty shape");
  WOQ_CHECK_ARG(x_dtype >= 0 && x_dtype <= 2 && y_dtype >= 0 && y_dtype <= 2, "w8a8_linear_forward: bad dtype");
  WOQ_CHECK_ARG(workspace_bytes >= b200woq_w8a8_workspace_bytes(M, N, K), "w8a8_linear_forward: workspace too small");
  WOQ_CHECK_ARG((((uintptr_t)workspace) & 255) == 0 && (((uintptr_t)qweight) & 15) == 0, "w8a8_linear_forward: alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t Kp = pad_k(K);
  uint8_t* qx = (uint8_t*)workspace;
  int NT, splits, per;
  int64_t tiles;
  w8a8_plan(M, N, K, &NT, &splits, &per, &tiles);
  int* counters = (int*)(qx + b200woq_w8a8_workspace_zeroed_offset(M, K));   // zero on entry, left zero
  int32_t* acc_ws = (int32_t*)((uint8_t*)counters + ((tiles * 4 + 255) / 256) * 256);
  {
    const int64_t total = M * (Kp / 16);
    int64_t b = ceil_div(total, 256);
    const int64_t cap = (int64_t)num_sms() * 16;
    sq_quantize_act_kernel<<<(unsigned)(b > cap ? cap : b), 256, 0, st>>>(x, x_dtype, M, K, Kp, input_scale, x_scale, x_zp, qx);
    WOQ_LAUNCH_CHECK();
  }
  // tensor maps are pure functions of (pointer, shape, box): a small thread-local cache avoids two driver encodes per call
  struct MapKey { const void* p; int64_t a, b; int box; };
  struct MapSlot { MapKey k; CUtensorMap m; bool used; };
  static thread_local MapSlot cache[16] = {};
  static thread_local int next_slot = 0;
  auto get_map = [&](const void* base_, int64_t rows, int box, CUtensorMap* out) -> bool {
    for (int i = 0; i < 16; ++i)
      if (cache[i].used && cache[i].k.p == base_ && cache[i].k.a == Kp && cache[i].k.b == rows && cache[i].k.box == box) {
        *out = cache[i].m;
        return true;
      }
    if (!encode_u8_2d(out, base_, Kp, rows, Kp, box)) return false;
    MapSlot& sl = cache[next_slot];
    next_slot = (next_slot + 1) & 15;
    sl.k = MapKey{base_, Kp, rows, box};
    sl.m = *out;
    sl.used = true;
    return true;
  };
  CUtensorMap map_w, map_x;
  if (!get_map(qweight, N, 128, &map_w) || !get_map(qx, M, NT, &map_x)) {
    set_error("w8a8_linear_forward: cuTensorMapEncodeTiled failed");
    return B200WOQ_ECUDA;
  }
  GemmParams p = {};
  p.M = (int)M; p.N = (int)N; p.K = (int)Kp; p.NT = NT; p.kb_per_split = per;
  p.wsum = wsum; p.w_scale = w_scale; p.x_scale = x_scale; p.x_zp = x_zp; p.bias = bias; p.bias_dtype = bias_dtype;
  p.y = y; p.y_dtype = y_dtype; p.acc_ws = acc_ws; p.counters = counters;
  // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D = S32 (2), A = INT8 (1) weights, B = UINT8 (0)
  // activations, both K-major, N = NT, M = 128
  const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div(M, NT), (unsigned)splits);
#define W8A8_LAUNCH(NTM)                                                                                              \
  do {                                                                                                                \
    const size_t smem = (size_t)STAGES * (TM * BKB + (NTM) * BKB) + 1024 + 256;                                       \
    WOQ_CUDA(cudaFuncSetAttribute(w8a8_gemm_kernel<NTM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    w8a8_gemm_kernel<NTM><<<grid, 192, smem, st>>>(map_w, map_x, p, idesc);                                           \
  } while (0)
  if (NT <= 32) W8A8_LAUNCH(32);
  else if (NT <= 64) W8A8_LAUNCH(64);
  else if (NT <= 128) W8A8_LAUNCH(128);
  else W8A8_LAUNCH(256);
#undef W8A8_LAUNCH
  WOQ_LAUNCH_CHECK();
  return 0;
}
