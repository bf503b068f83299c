// woq_gemm.cu -- K6: INCWeightOnlyLinear.forward as ONE fused kernel (unpack + dequant + matmul).
//
// Reference: modules.py:594-610 forward, :413-443 recover, :377-411 unpack.  The reference de-quantises
// the whole weight once (fp16( int8(q - zp) * scale_fp16 )), caches it as fp32 and calls F.linear; here
// the packed words are streamed from HBM exactly once per call and never materialised as fp weights.
//
// Layout (optimum format): qweight int32 [K/n_pack, N] -- field e of word (kw, n) is the code of
// W[n, kw*n_pack + e]; consecutive n are contiguous, so a warp reading 128 B-512 B segments of one kw
// row is fully coalesced.  scales fp16 [G, N], qzeros int32 [G, N/n_pack] (stored zp-1).
//
// Fast path (bits 4/8, no g_idx, N % 32 == 0, group % (4*k_per_word) == 0): batch M <= 64 is HBM-bound
// (ridge at M ~ 70), so the design goal is bytes in flight, not flops:
//   * every lane keeps D = 8 independent 16-byte ld.global.nc (L1::no_allocate) loads in flight,
//     16 warps/SM -> 64 KB/SM outstanding;
//   * the 4-bit -> fp16 conversion uses the 0x6400 magic-number trick (2 LOP3 + 1 SHF per 4 codes, exact
//     subtraction of 1024+zp in half2), and the multiply-accumulate runs on the tensor cores through
//     mma.sync.m16n8k16 with the out-channel dimension as the MMA "M" (16 rows = 4 lanes-groups x 4 n) and
//     the batch as the MMA "N" (8 columns), so CUDA-core issue slots are spent only on the dequant;
//   * k is consumed in a permuted order (the dot product is permutation invariant): lane t of a quad
//     owns qweight row kw0+t, so one 16-byte load feeds 4 MMAs without any shuffles; x is staged once per
//     CTA into shared memory in the matching permuted order;
//   * M <= 8: group scale applied after the per-group accumulation in fp32 (POST mode, 9 ALU ops/word);
//     M  > 8: scale applied in half2 before the MMA, reproducing the reference's fp16 weight exactly;
//   * split-K across CTAs with a deterministic "last CTA reduces" epilogue (fixed summation order).
// The general path (g_idx / ragged shapes / other bit widths) is a plain CUDA-core kernel.
#include <cooperative_groups.h>

#include "common.cuh"
#include <cstdlib>

namespace b200woq {

struct GemmParams {
  const void* x;
  int x_dtype;
  int64_t M, K, N;  // M = rows of this chunk (<= 64 on the fast path)
  const int32_t* qweight;
  const int32_t* qzeros;
  const __half* scales;
  const int32_t* g_idx;
  const void* bias;
  int bias_dtype;
  const float* input_scale;
  void* y;
  int y_dtype;
  int bits, g;
  int64_t G;
  int S;      // split-K factor == cluster size along x (1..8)
  int slice;  // columns of the 128-wide tile reduced by each cluster rank (multiple of 4)
  int gmax;   // max groups per CTA
  int xs_ld;  // halves per smem x row
  int pdl;
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

__device__ __forceinline__ int4 ldg_nc_v4(const int32_t* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

constexpr int kPrefetch = 8;  // 16-byte loads in flight per lane
constexpr int kRedPerM = 320; // floats of split-K exchange buffer per batch row: 2*S*slice <= 288 for S in 1..8

// BITS = 4: word -> P0=(c0,c4) P1=(c1,c5) P2=(c2,c6) P3=(c3,c7) as half2 of (code - zp)
// BITS = 8: word -> P0=(c0,c2) P1=(c1,c3)
template <int BITS>
struct Dequant;

template <>
struct Dequant<4> {
  static constexpr int kPairs = 4;
  // zc = half2(1024 + zp), zn = half2(-(64 + zp))
  static __device__ __forceinline__ void run(uint32_t w, __half2 zc, __half2 zn, __half2 (&P)[4]) {
    const __half2 sixteenth = __float2half2_rn(0.0625f);
    const uint32_t w8 = w >> 8;
    P[0] = __hsub2(u32_as_h2((w & 0x000f000fu) | 0x64006400u), zc);
    P[1] = __hfma2(u32_as_h2((w & 0x00f000f0u) | 0x64006400u), sixteenth, zn);
    P[2] = __hsub2(u32_as_h2((w8 & 0x000f000fu) | 0x64006400u), zc);
    P[3] = __hfma2(u32_as_h2((w8 & 0x00f000f0u) | 0x64006400u), sixteenth, zn);
  }
};

template <>
struct Dequant<8> {
  static constexpr int kPairs = 2;
  static __device__ __forceinline__ void run(uint32_t w, __half2 zc, __half2 /*zn*/, __half2 (&P)[4]) {
    P[0] = __hsub2(u32_as_h2((w & 0x00ff00ffu) | 0x64006400u), zc);
    P[1] = __hsub2(u32_as_h2(((w >> 8) & 0x00ff00ffu) | 0x64006400u), zc);
    // .to(torch.int8) in recover() wraps q - zp into [-128, 127] (modules.py:435)
    const __half2 hi = __float2half2_rn(127.f), lo = __float2half2_rn(-128.f), wrap = __float2half2_rn(256.f);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      P[i] = __hsub2(P[i], __hmul2(__hgt2(P[i], hi), wrap));
      P[i] = __hadd2(P[i], __hmul2(__hlt2(P[i], lo), wrap));
    }
  }
};

// 8 consecutive activations -> fp16, permuted to the MMA k order of one (BITS=4) or two (BITS=8) words
template <int BITS>
__device__ __forceinline__ uint4 permute_pack8(const float (&v)[8]) {
  __half2 h[4];
  if (BITS == 4) {  // [x0,x4,x1,x5,x2,x6,x3,x7]
    h[0] = __floats2half2_rn(v[0], v[4]);
    h[1] = __floats2half2_rn(v[1], v[5]);
    h[2] = __floats2half2_rn(v[2], v[6]);
    h[3] = __floats2half2_rn(v[3], v[7]);
  } else {  // two 4-code words: [x0,x2,x1,x3 | x4,x6,x5,x7]
    h[0] = __floats2half2_rn(v[0], v[2]);
    h[1] = __floats2half2_rn(v[1], v[3]);
    h[2] = __floats2half2_rn(v[4], v[6]);
    h[3] = __floats2half2_rn(v[5], v[7]);
  }
  return make_uint4(h2_as_u32(h[0]), h2_as_u32(h[1]), h2_as_u32(h[2]), h2_as_u32(h[3]));
}

__device__ __forceinline__ void load8_as_float(const void* base, int dtype, int64_t idx, float (&v)[8]) {
  if (dtype == B200WOQ_F32) {
    const float4 a = *reinterpret_cast<const float4*>((const float*)base + idx);
    const float4 b = *reinterpret_cast<const float4*>((const float*)base + idx + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 r = *reinterpret_cast<const uint4*>((const uint16_t*)base + idx);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (dtype == B200WOQ_F16) {
        const float2 f = __half22float2(u32_as_h2(w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
      } else {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
    }
  }
}

// grid = (S, n_tiles) with thread-block clusters of S CTAs along x: cluster rank r owns the k-groups
// [r*G/S, (r+1)*G/S) of the n tile.  Split-K partial sums never touch global memory: every warp writes its partial
// straight into the distributed shared memory of the rank that owns its columns (each rank reduces 128/S columns of
// the tile over the 2*S sources in a fixed order -> deterministic, no atomics, no global scratch).
//
// SUB (4-bit, M <= 16): the codes enter the tensor core as fp16 SUBNORMALS.  `w & 0x000f000f` *is* the half2
// (q_lo * 2^-24, q_hi * 2^-24) -- no magic-number add, no half2 arithmetic at all: 7 integer ops per 8 codes.
// Products q*2^-24*x are exact in the fp32 accumulator; per group the zero-point is applied algebraically,
//     sum_k (q_k - z) x_k = 2^24 * acc_g - z * X_g ,   X_g = sum_{k in group} x_k  (fp32, once per CTA),
// and the group scale multiplies the result in fp32.  This is the exact (q-z)*scale product, i.e. slightly MORE
// accurate than the reference's fp16-rounded weight fp16((q-z)*scale); the difference is <= 2^-11 relative per weight.
// !SUB: codes -> (q - z) via the 0x6400 magic number, times the fp16 scale in half2 = the reference's fp16 weight.
template <int BITS, int MT, bool SUB, bool NI4>
__global__ void __launch_bounds__(256, (MT <= 2) ? 2 : 1) woq_gemm_mma_kernel(const GemmParams p) {
  static_assert(!SUB || BITS == 4, "subnormal-code path is 4-bit only");
  constexpr int KPW = 32 / BITS;   // codes per word
  constexpr int KSTEP = 4 * KPW;   // k covered by the 4 quad-lanes' words
  constexpr int NSTEP = (BITS == 4) ? 2 : 1;  // MMA k16 steps per 16-byte load
  constexpr int ZW = 128 / KPW;    // qzeros words per group per 128-column tile
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // smem: [ red: M*kRedPerM floats ][ xs: M*xs_ld halves ][ xsum: M*gmax floats ][ sc_s: gmax*128 halves ][ zr_s ]
  float* red = reinterpret_cast<float*>(smem_raw);
  __half* xs = reinterpret_cast<__half*>(smem_raw + (size_t)p.M * kRedPerM * sizeof(float));
  float* xsum = reinterpret_cast<float*>(xs + (size_t)p.M * p.xs_ld);
  __half* sc_s = reinterpret_cast<__half*>(xsum + (((size_t)p.M * p.gmax + 3) & ~(size_t)3));  // keep 16-byte alignment
  uint32_t* zr_s = reinterpret_cast<uint32_t*>(sc_s + (size_t)p.gmax * 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;
  const int strip = warp & 3, khalf = warp >> 2;
  const int N = (int)p.N;
  const int n_tile0 = blockIdx.y * 128;
  const int n_in_tile = strip * 32 + 4 * gq;
  const int n0 = n_tile0 + n_in_tile;
  const bool strip_valid = n_tile0 + strip * 32 < N;
  const int g = p.g;
  const int NI = NI4 ? 4 : g / KSTEP;  // 16-byte loads per group per lane (NI4: group_size == 4*KSTEP)
  const int S = p.S;
  const int rank = blockIdx.x;  // == cluster rank (cluster dims = (S,1,1), gridDim.x == S)
  // all CTAs of the cluster must be resident before anyone writes into a peer's shared memory: arrive now,
  // wait just before the exchange (the barrier completes in the background while we stream weights)
  if (S > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");

  // CTA group range and this warp's half of it
  const int Gt = (int)p.G;
  const int gb = rank * Gt / S, ge = (rank + 1) * Gt / S;
  const int gmid = gb + (ge - gb + 1) / 2;
  const int g0 = khalf == 0 ? gb : gmid, g1 = khalf == 0 ? gmid : ge;
  const int total = strip_valid ? (g1 - g0) * NI : 0;

  // (1) weights: word row of iteration `it` is g0*g/KPW + 4*it + t (groups are contiguous in k)
  const int32_t* wptr = p.qweight + ((int64_t)g0 * (g / KPW) + t) * N + n0;
  const int64_t wstep = 4 * (int64_t)N;
  int4 buf[kPrefetch];
#pragma unroll
  for (int j = 0; j < kPrefetch; ++j)
    if (j < total) buf[j] = ldg_nc_v4(wptr + (int64_t)j * wstep);
  wptr += (int64_t)kPrefetch * wstep;

  // (2) scales / zero-points of the CTA's groups -> smem (cp.async, 16-byte chunks)
  {
    const int ng = ge - gb;
    const int sc_chunks = ng * 16, z_chunks = ng * (ZW / 4);
    const int Nw = N / KPW;
    for (int c = threadIdx.x; c < sc_chunks + z_chunks; c += 256) {
      if (c < sc_chunks) {
        const int gl = c >> 4, ch = c & 15;
        if (n_tile0 + ch * 8 < N)
          cp_async16(sc_s + gl * 128 + ch * 8, p.scales + (int64_t)(gb + gl) * N + n_tile0 + ch * 8);
      } else {
        const int cz = c - sc_chunks;
        const int gl = cz / (ZW / 4), ch = cz % (ZW / 4);
        if (n_tile0 + ch * 4 * KPW < N)
          cp_async16(zr_s + gl * ZW + ch * 4, p.qzeros + (int64_t)(gb + gl) * Nw + n_tile0 / KPW + ch * 4);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // everything above is constant data; x may be produced by the previous kernel
  if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");

  // (3) x: 8-element chunks (m outer, chunk inner: no divisions), vector loads, convert / permute / store
  {
    const int64_t kbase = (int64_t)gb * g;
    const int ksz8 = (ge - gb) * g / 8;
    for (int m = 0; m < (int)p.M; ++m) {
      const int64_t row = (int64_t)m * p.K + kbase;
      for (int kc = threadIdx.x; kc < ksz8; kc += 256) {
        float v[8];
        load8_as_float(p.x, p.x_dtype, row + kc * 8, v);
        if (p.input_scale) {
          float sc8[8];
          load8_as_float(p.input_scale, B200WOQ_F32, kbase + kc * 8, sc8);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] *= sc8[i];
        }
        *reinterpret_cast<uint4*>(xs + m * p.xs_ld + kc * 8) = permute_pack8<BITS>(v);
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  // let the next kernel in the stream start its own (constant) weight prefetch while we compute
  if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (SUB) {  // X_g[m] = sum of the (fp16-rounded) activations of each group, fp32
    const int ng = ge - gb;
    for (int task = warp; task < (int)p.M * ng; task += 8) {
      const int m = task / ng, gl = task - m * ng;
      float sum = 0.f;
      for (int e = lane * 4; e < g; e += 128) {
        const uint2 v = *reinterpret_cast<const uint2*>(xs + m * p.xs_ld + gl * g + e);
        const float2 a = __half22float2(u32_as_h2(v.x)), b = __half22float2(u32_as_h2(v.y));
        sum += (a.x + a.y) + (b.x + b.y);
      }
      sum = warp_sum(sum);
      if (lane == 0) xsum[m * p.gmax + gl] = sum;
    }
    __syncthreads();
  }

  float acc[2][MT][4];
  float accg[2][MT][4];  // SUB only: per-group partial sums (dead code otherwise)
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < MT; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[a][b][c] = 0.f;
        accg[a][b][c] = 0.f;
      }

  __half2 zc[4], zn[4], sh[4];
  float sf[4], zf[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    zc[r] = zn[r] = sh[r] = __float2half2_rn(0.f);
    sf[r] = zf[r] = 0.f;
  }
  const int zshift = (BITS == 4) ? ((n0 & 7) * 4) : 0;
  int gl = g0 - gb;
  int xoff = (g0 - gb) * g + t * KPW;  // halves; advances by 4*KPW per iteration

  auto group_start = [&]() {
    const uint2 sc = *reinterpret_cast<const uint2*>(sc_s + gl * 128 + n_in_tile);
    const uint32_t zw = zr_s[gl * ZW + n_in_tile / KPW] >> zshift;
    const __half2 s01 = u32_as_h2(sc.x), s23 = u32_as_h2(sc.y);
    const __half sr[4] = {__low2half(s01), __high2half(s01), __low2half(s23), __high2half(s23)};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      // stored zp-1, +1 wraps to 0 past the maximum (modules.py:363, 409-410)
      const uint32_t z = (((zw >> (BITS * r)) & ((1u << BITS) - 1u)) + 1u) & ((1u << BITS) - 1u);
      sf[r] = __half2float(sr[r]);
      if (SUB) {
        zf[r] = (float)z;
      } else {
        zc[r] = __float2half2_rn(1024.f + (float)z);
        zn[r] = __float2half2_rn(-(64.f + (float)z));
        sh[r] = __half2half2(sr[r]);
      }
    }
  };
  auto do_iter = [&](const int4 wv) {
    uint32_t P[4][4];
    const uint32_t wr[4] = {(uint32_t)wv.x, (uint32_t)wv.y, (uint32_t)wv.z, (uint32_t)wv.w};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (SUB) {
        P[r][0] = wr[r] & 0x000f000fu;
        P[r][1] = (wr[r] >> 4) & 0x000f000fu;
        P[r][2] = (wr[r] >> 8) & 0x000f000fu;
        P[r][3] = (wr[r] >> 12) & 0x000f000fu;
      } else {
        __half2 h[4];
        Dequant<BITS>::run(wr[r], zc[r], zn[r], h);
#pragma unroll
        for (int q = 0; q < Dequant<BITS>::kPairs; ++q) P[r][q] = h2_as_u32(__hmul2(h[q], sh[r]));
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int m = gq + 8 * mt;
      uint32_t xb[4] = {0u, 0u, 0u, 0u};
      if (m < p.M) {
        if (BITS == 4) {
          const uint4 v = *reinterpret_cast<const uint4*>(xs + m * p.xs_ld + xoff);
          xb[0] = v.x; xb[1] = v.y; xb[2] = v.z; xb[3] = v.w;
        } else {
          const uint2 v = *reinterpret_cast<const uint2*>(xs + m * p.xs_ld + xoff);
          xb[0] = v.x; xb[1] = v.y;
        }
      }
#pragma unroll
      for (int st = 0; st < NSTEP; ++st) {
        if (SUB) {
          mma_16816(accg[0][mt], P[0][2 * st], P[1][2 * st], P[0][2 * st + 1], P[1][2 * st + 1], xb[2 * st], xb[2 * st + 1]);
          mma_16816(accg[1][mt], P[2][2 * st], P[3][2 * st], P[2][2 * st + 1], P[3][2 * st + 1], xb[2 * st], xb[2 * st + 1]);
        } else {
          mma_16816(acc[0][mt], P[0][2 * st], P[1][2 * st], P[0][2 * st + 1], P[1][2 * st + 1], xb[2 * st], xb[2 * st + 1]);
          mma_16816(acc[1][mt], P[2][2 * st], P[3][2 * st], P[2][2 * st + 1], P[3][2 * st + 1], xb[2 * st], xb[2 * st + 1]);
        }
      }
    }
    xoff += 4 * KPW;
  };
  auto group_end = [&]() {
    if (SUB) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int m0 = 8 * mt + 2 * t;
        const float x0 = (m0 < p.M) ? xsum[m0 * p.gmax + gl] : 0.f;
        const float x1 = (m0 + 1 < p.M) ? xsum[(m0 + 1) * p.gmax + gl] : 0.f;
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const float sa = sf[2 * tile], sb = sf[2 * tile + 1], za = zf[2 * tile], zb = zf[2 * tile + 1];
          acc[tile][mt][0] = fmaf(fmaf(accg[tile][mt][0], 16777216.f, -za * x0), sa, acc[tile][mt][0]);
          acc[tile][mt][1] = fmaf(fmaf(accg[tile][mt][1], 16777216.f, -za * x1), sa, acc[tile][mt][1]);
          acc[tile][mt][2] = fmaf(fmaf(accg[tile][mt][2], 16777216.f, -zb * x0), sb, acc[tile][mt][2]);
          acc[tile][mt][3] = fmaf(fmaf(accg[tile][mt][3], 16777216.f, -zb * x1), sb, acc[tile][mt][3]);
          accg[tile][mt][0] = accg[tile][mt][1] = accg[tile][mt][2] = accg[tile][mt][3] = 0.f;
        }
      }
    }
    ++gl;
  };

  int it_load = kPrefetch;  // next iteration whose words get loaded into the ring
  if (NI4) {                // 4 loads per group: two groups per pass over the 8-deep ring, no per-iteration tests
    const int ngw = total / 4;
#pragma unroll 1
    for (int gi = 0; gi < ngw; gi += 2) {
      group_start();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int4 wv = buf[j];
        if (it_load < total) buf[j] = ldg_nc_v4(wptr);
        wptr += wstep;
        ++it_load;
        do_iter(wv);
      }
      group_end();
      if (gi + 1 < ngw) {
        group_start();
#pragma unroll
        for (int j = 4; j < 8; ++j) {
          const int4 wv = buf[j];
          if (it_load < total) buf[j] = ldg_nc_v4(wptr);
          wptr += wstep;
          ++it_load;
          do_iter(wv);
        }
        group_end();
      }
    }
  } else {
    int i_in_g = 0;
    for (int base = 0; base < total; base += kPrefetch) {
#pragma unroll
      for (int j = 0; j < kPrefetch; ++j) {
        const int it = base + j;
        if (it < total) {
          const int4 wv = buf[j];
          if (it_load < total) buf[j] = ldg_nc_v4(wptr);
          wptr += wstep;
          ++it_load;
          if (i_in_g == 0) group_start();
          do_iter(wv);
          if (++i_in_g == NI) {
            i_in_g = 0;
            group_end();
          }
        }
      }
    }
  }

  // ---- epilogue ----
  // lane holds y[m = 8mt + 2t + {0,1}][n0 + {0,1,2,3}]:
  // acc[tile][mt][c]: c0,c1 -> row gq (n0 + 2*tile), cols 2t,2t+1 ; c2,c3 -> row gq+8 (n0 + 2*tile + 1)
  // columns of the tile are owned by cluster ranks in slices of 128/S; every warp writes its partial of a slice
  // into the owner's `red` ([source = 2*rank + khalf][m][128/S]) through distributed shared memory
  // slices are multiples of 4 columns (a lane's float4) and cover the 128-column tile: slice_w = ceil(128/S) -> x4
  const int slice = p.slice;
  if (S > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (strip_valid) {
    const int owner = n_in_tile / slice;
    float* owner_red = (S == 1) ? red : cluster.map_shared_rank(red, owner);
    const int src = 2 * rank + khalf;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int m = 8 * mt + 2 * t + half;
        if (m < p.M) {
          // n0+0: tile0 c(half) ; n0+1: tile0 c(2+half) ; n0+2: tile1 c(half) ; n0+3: tile1 c(2+half)
          const float4 v = make_float4(acc[0][mt][half], acc[0][mt][2 + half], acc[1][mt][half], acc[1][mt][2 + half]);
          *reinterpret_cast<float4*>(owner_red + (src * (int)p.M + m) * slice + (n_in_tile - owner * slice)) = v;
        }
      }
  }
  if (S > 1) cluster.sync(); else __syncthreads();
  {
    // red: [2S sources][M][slice].  One half-warp per output element: lane l < 2S loads source l, xor-shuffle tree
    // (fixed order -> deterministic), lane 0 / 16 stores.
    const int nbase = n_tile0 + rank * slice;
    const int width = min(slice, 128 - rank * slice);   // the last rank's slice may be shorter
    const int nsrc = 2 * S;
    const int hw = lane >> 4, hl = lane & 15;
    for (int e = warp * 2 + hw; e < (int)p.M * width; e += 16) {
      const int m = e / width, nl = e - m * width;
      float v = (hl < nsrc) ? red[(hl * (int)p.M + m) * slice + nl] : 0.f;
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      const int n = nbase + nl;
      if (hl == 0 && n < N) {
        if (p.bias) v += load_as_float(p.bias, p.bias_dtype, n);
        store_from_float(p.y, p.y_dtype, (int64_t)m * N + n, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// general path: any bits in [1,8], g_idx, ragged K / group.  32 lanes <-> 32 out-channels, 8 warps split K,
// 8 batch rows per CTA.y.  Weight value = fp16(int8(q - zp) * scale) exactly as recover() computes it.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) woq_gemm_general_kernel(const GemmParams p) {
  __shared__ float red[8][8][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * 32 + lane;
  const int64_t m0 = (int64_t)blockIdx.y * 8;
  const int n_pack = 32 / p.bits;
  const uint32_t mask = (1u << p.bits) - 1u;
  const int64_t Kw = (p.K + n_pack - 1) / n_pack, Nw = (p.N + n_pack - 1) / n_pack;
  const int64_t per = (Kw + 7) / 8;
  const int64_t kw_begin = warp * per, kw_end = (kw_begin + per < Kw) ? kw_begin + per : Kw;
  float acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.f;
  if (n < p.N) {
    for (int64_t kw = kw_begin; kw < kw_end; ++kw) {
      const uint32_t word = (uint32_t)p.qweight[kw * p.N + n];
      for (int e = 0; e < n_pack; ++e) {
        const int64_t k = kw * n_pack + e;
        if (k >= p.K) break;
        const int64_t gi = p.g_idx ? p.g_idx[k] : k / p.g;
        const uint32_t zw = (uint32_t)p.qzeros[gi * Nw + n / n_pack];
        uint32_t z = ((zw >> (p.bits * (n % n_pack))) & mask) + 1u;
        if (z > mask) z = 0;
        const int8_t d = (int8_t)((int)((word >> (p.bits * e)) & mask) - (int)z);
        const float wv = __half2float(__hmul(__int2half_rn((int)d), p.scales[gi * p.N + n]));
        const float is = p.input_scale ? p.input_scale[k] : 1.f;
#pragma unroll
        for (int m = 0; m < 8; ++m)
          if (m0 + m < p.M) acc[m] = fmaf(load_as_float(p.x, p.x_dtype, (m0 + m) * p.K + k) * is, wv, acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) red[warp][m][lane] = acc[m];
  __syncthreads();
  if (warp == 0 && n < p.N) {
    for (int m = 0; m < 8 && m0 + m < p.M; ++m) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += red[w][m][lane];
      if (p.bias) s += load_as_float(p.bias, p.bias_dtype, n);
      store_from_float(p.y, p.y_dtype, (m0 + m) * p.N + n, s);
    }
  }
}

static bool fast_path_ok(int64_t N, int64_t K, int bits, int g, const int32_t* g_idx) {
  if (g_idx) return false;
  if (bits != 4 && bits != 8) return false;
  const int kstep = 4 * (32 / bits);
  return (N % 32 == 0) && (g % kstep == 0) && (K % g == 0);
}

// split-K factor == cluster size in 1..8: as many CTAs as fit in ONE wave of 2 CTAs/SM (a second, nearly empty
// wave would double the kernel time), at least 2 groups per CTA so both k-halves of a CTA have work
static int choose_split(int64_t M, int64_t N, int64_t K, int g) {
  const int64_t n_tiles = ceil_div(N, 128);
  const int64_t G = K / g;
  const int64_t slots = 2 * (int64_t)num_sms();
  int best = 1;
  for (int S = 1; S <= 8; ++S) {
    if (G / S < 2 && S > 1) break;
    if (n_tiles * S <= slots) best = S;
  }
  (void)M;
  return best;
}

static size_t fast_smem_bytes(int64_t Mc, int mt, int64_t gmax, int g, int bits) {
  const int kpw = 32 / bits;
  const size_t red = (size_t)Mc * kRedPerM * sizeof(float);
  const size_t xs = (size_t)Mc * (gmax * g + 32) * sizeof(__half);
  const size_t xsum = (((size_t)Mc * gmax + 3) & ~(size_t)3) * sizeof(float);
  const size_t sc = (size_t)gmax * 128 * sizeof(__half), zr = (size_t)gmax * (128 / kpw) * sizeof(uint32_t);
  (void)mt;
  return red + xs + xsum + sc + zr + 64;
}

}  // namespace b200woq

using namespace b200woq;

namespace b200woq {
// woq_tc.cu: tcgen05 dequant-GEMM for batches of 9..128 rows
bool woq_tc_shape_ok(int64_t M, int64_t N, int64_t K, int bits, int g, const int32_t* g_idx);
int64_t woq_tc_workspace_bytes(int64_t M, int64_t N, int64_t K);
int woq_tc_forward(const void* x, int x_dtype, int64_t M, int64_t K, int64_t N, const int32_t* qweight, const int32_t* qzeros,
                   const __half* scales, const void* bias, int bias_dtype, const float* input_scale, void* y, int y_dtype,
                   int g, void* workspace, int64_t workspace_bytes, int pdl, cudaStream_t st);
}  // namespace b200woq

static int tc_min_rows() {
  static const int v = getenv("B200WOQ_TC_MIN_ROWS") ? atoi(getenv("B200WOQ_TC_MIN_ROWS")) : 5;
  return v;
}

// The cluster kernel keeps its split-K partials in distributed shared memory; the tensor-core kernel (9..128 rows) needs
// tile counters (zero on first use, left zero) + fp32 partial tiles in global memory.
extern "C" int64_t b200woq_linear_workspace_bytes(int64_t M, int64_t N, int64_t K, int bits, int group_size) {
  const int g = eff_group(K, group_size);
  if (M >= tc_min_rows() && woq_tc_shape_ok(M, N, K, bits, g, nullptr)) return woq_tc_workspace_bytes(M, N, K);
  return 256;
}

template <int BITS>
static int launch_fast(GemmParams& p, int64_t Mc, cudaStream_t st) {
  const int mt = Mc <= 8 ? 1 : Mc <= 16 ? 2 : Mc <= 32 ? 4 : 8;
  // shrink the split until the tile fits in shared memory
  while (true) {
    p.gmax = (int)ceil_div(p.G, p.S);
    if (fast_smem_bytes(Mc, mt, p.gmax, p.g, BITS) <= 200 * 1024 || p.S >= 8 || p.G / (p.S + 1) < 2) break;
    p.S += 1;
  }
  p.slice = (int)((ceil_div(128, p.S) + 3) & ~3);
  p.xs_ld = p.gmax * p.g + 32;
  // the k-half exchange buffer aliases xs: make sure the row count covers it
  size_t smem = fast_smem_bytes(Mc, mt, p.gmax, p.g, BITS);
  if (smem > 220 * 1024) {
    set_error("linear_forward: K slice does not fit in shared memory (M=%lld K=%lld)", (long long)Mc, (long long)p.K);
    return B200WOQ_EUNSUPPORTED;
  }
  dim3 grid((unsigned)p.S, (unsigned)ceil_div(p.N, 128));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = (unsigned)p.S;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (p.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const bool ni4 = (p.g == 16 * (32 / BITS));  // 4 sixteen-byte loads per group per lane (g=128 @4-bit, 64 @8-bit)
#define WOQ_LAUNCH(MT_, POST_)                                                                              \
  do {                                                                                                      \
    auto kern = ni4 ? woq_gemm_mma_kernel<BITS, MT_, POST_, true> : woq_gemm_mma_kernel<BITS, MT_, POST_, false>; \
    if (smem > 48 * 1024) WOQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    WOQ_CUDA(cudaLaunchKernelEx(&cfg, kern, p));                                                            \
    count_launch(1);                                                                                        \
  } while (0)
  constexpr bool kSub = (BITS == 4);  // subnormal-code path for the small-batch 4-bit kernels
  switch (mt) {
    case 1: WOQ_LAUNCH(1, kSub); break;
    case 2: WOQ_LAUNCH(2, kSub); break;
    case 4: WOQ_LAUNCH(4, false); break;
    default: WOQ_LAUNCH(8, false); break;
  }
#undef WOQ_LAUNCH
  return 0;
}

extern "C" int b200woq_linear_forward(const void* x, int x_dtype, int64_t M, int64_t K, int64_t N,
                                      const int32_t* qweight, const int32_t* qzeros, const void* scales16,
                                      const int32_t* g_idx, const void* bias, int bias_dtype,
                                      const float* input_scale, void* y, int y_dtype, int bits, int group_size,
                                      void* workspace, int64_t workspace_bytes, int flags, void* stream) {
  WOQ_CHECK_ARG(x && qweight && qzeros && scales16 && y, "linear_forward: null pointer");
  WOQ_CHECK_ARG(M > 0 && K > 0 && N > 0, "linear_forward: empty shape");
  WOQ_CHECK_ARG(bits >= 1 && bits <= 8, "linear_forward: bits must be in [1,8]");
  WOQ_CHECK_ARG(x_dtype >= 0 && x_dtype <= 2 && y_dtype >= 0 && y_dtype <= 2, "linear_forward: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = eff_group(K, group_size);
  GemmParams p = {};
  p.x_dtype = x_dtype;
  p.K = K;
  p.N = N;
  p.qweight = qweight;
  p.qzeros = qzeros;
  p.scales = (const __half*)scales16;
  p.g_idx = g_idx;
  p.bias = bias;
  p.bias_dtype = bias_dtype;
  p.input_scale = input_scale;
  p.y_dtype = y_dtype;
  p.bits = bits;
  p.g = g;
  p.G = ceil_div(K, g);
  p.pdl = (flags & 2) ? 1 : 0;
  const size_t xes = x_dtype == B200WOQ_F32 ? 4 : 2, yes = y_dtype == B200WOQ_F32 ? 4 : 2;
  if (!(flags & 1) && x_dtype == B200WOQ_F16 && !input_scale && (((uintptr_t)x) & 15) == 0 && M >= tc_min_rows() &&
      woq_tc_shape_ok(M, N, K, bits, g, g_idx) && workspace && workspace_bytes >= woq_tc_workspace_bytes(M, N, K))
    return woq_tc_forward(x, x_dtype, M, K, N, qweight, qzeros, (const __half*)scales16, bias, bias_dtype, input_scale, y,
                          y_dtype, g, workspace, workspace_bytes, p.pdl, st);
  const bool fast = !(flags & 1) && fast_path_ok(N, K, bits, g, g_idx);
  if (!fast) {
    p.x = x;
    p.y = y;
    p.M = M;
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, 8));
    woq_gemm_general_kernel<<<grid, 256, 0, st>>>(p);
    WOQ_LAUNCH_CHECK();
    return 0;
  }
  // largest batch chunk whose x slice fits in shared memory at the maximum split
  int64_t cap = 64;
  while (cap > 8) {
    const int mtc = cap <= 8 ? 1 : cap <= 16 ? 2 : cap <= 32 ? 4 : 8;
    if (fast_smem_bytes(cap, mtc, ceil_div(p.G, 8), g, bits) <= 200 * 1024) break;
    cap /= 2;
  }
  for (int64_t m0 = 0; m0 < M; m0 += cap) {
    const int64_t Mc = (M - m0) < cap ? (M - m0) : cap;
    p.x = (const char*)x + (size_t)m0 * K * xes;
    p.y = (char*)y + (size_t)m0 * N * yes;
    p.M = Mc;
    p.S = choose_split(Mc, N, K, g);
    int rc = (bits == 4) ? launch_fast<4>(p, Mc, st) : launch_fast<8>(p, Mc, st);
    if (rc) return rc;
  }
  return 0;
}
