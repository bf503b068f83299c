// woq_stream.cu -- K6 at small batch (M <= 16, 4-bit): the packed weights are streamed HBM -> shared memory by the
// TMA engine (cp.async.bulk + mbarrier ring) from a B200-native "stream layout" of the optimum-format tensors.
//
// Why a derived layout: in the reference's optimum format consecutive k-rows of qweight are N*4 bytes apart, so a
// CTA that owns 128 columns touches one 512-byte piece of every row (DRAM-row unfriendly, and every 16-byte piece
// needs its own load instruction).  The stream layout stores, for every (32-column strip, quantisation group), ONE
// contiguous record
//     [ NI iterations x 32 lanes x int4 packed words | 32 fp16 scales | 32 u8 zero-points ]      (2144 B at g=128)
// with the words already in the order the MMA lanes consume them, strips outermost so that the groups a warp owns are
// one contiguous span.  Every WARP runs its own ring of 8 records (lane 0 issues cp.async.bulk, an mbarrier per stage
// signals arrival): bytes in flight are independent of registers, warps never wait for each other, weights are
// constants so under programmatic dependent launch the rings fill while the previous layer is still finishing, and
// the lanes read their int4 with conflict-free LDS.128.  No clusters, no producer warp, one __syncthreads per CTA.  The derived copy is built once per module (like the reference caches its
// de-quantised fp32 weight at first forward, modules.py:603-604); the checkpoint tensors stay in optimum format.
//
// Arithmetic = the SUB path of woq_gemm.cu: codes enter mma.sync.m16n8k16 as fp16 subnormals; here the nibbles at
// bit positions 4-7 / 12-15 are used in place (value q*2^-20) and the matching activations are pre-scaled by 2^-4
// when x is staged, which removes the shifts: 5 integer ops per 8 codes.
#include <cooperative_groups.h>

#include <type_traits>

#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace b200woq {
namespace stream {

constexpr int kMaxStages = 8;  // records in flight per warp (runtime nst <= this)

struct Params {
  const void* x;
  int x_dtype;
  int M, K, N;
  const uint8_t* recs;  // [n_strips][G] records
  int rec_bytes;        // NI*512 + 64 + 32
  int NI;               // 16-byte loads per lane per group = g / 32
  int g, G;
  const void* bias;
  int bias_dtype;
  const float* input_scale;
  void* y;
  int y_dtype;
  int wpc;      // warps per CTA (each owns a contiguous range of groups)
  int gw_max;   // max groups per warp
  int xs_ld;    // halves per staged x row (per warp)
  int nst;      // ring depth per warp
  int warp_stride;  // bytes of shared memory per warp
  int pdl;
};

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
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
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
__device__ __forceinline__ void load8(const void* base, int dtype, int64_t idx, float (&v)[8]) {
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
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
      } else {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
    }
  }
}

__device__ __forceinline__ void mma_16816_zero(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}

// grid = N/32 CTAs (one 32-column strip each); block = wpc warps; warp w streams the groups
// [w*gw_max, min((w+1)*gw_max, G)) of the strip through ITS OWN ring of nst records (lane 0 issues the bulk copies, a
// per-stage mbarrier signals arrival).  Warps never synchronise with each other until the final in-CTA reduction (one
// __syncthreads).  The kernel is issue-bound, not bandwidth-bound, so everything per-warp and per-group is kept to a
// minimum: no integer division, the first MMA of a group takes C = 0 instead of zeroing accumulators, batch rows that
// do not exist read a 320-byte zero pad with a zero stride.
// MODE: 0 -> M == 1, 1 -> M <= 8.   NI_T: 16-byte loads per lane per group (g/32) when known, 0 = runtime.
template <int MODE, int NI_T>
__global__ void __launch_bounds__(1024) woq_gemm_stream_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;
  const int strip = blockIdx.x;
  const int NI = NI_T ? NI_T : p.NI;
  const int g = NI * 32;
  const int g0 = min(warp * p.gw_max, p.G);
  const int ngw = min(p.gw_max, p.G - g0);
  // per-warp smem: [ ring: nst*rec ][ bars ][ xs: M rows x xs_ld f16 ][ xsum: M*gw_max f32 ]; per CTA: red, zero pad
  const int nst = p.nst;
  const int rec_bytes = NI * 512 + 96;
  uint8_t* ring = smem_raw + (size_t)warp * p.warp_stride;
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t bars = ring_u32 + nst * rec_bytes;
  __half* xs = reinterpret_cast<__half*>(ring + nst * rec_bytes + ((nst * 8 + 15) & ~15));
  float* xsum = reinterpret_cast<float*>(xs + p.M * p.xs_ld);
  float* red = reinterpret_cast<float*>(smem_raw + (size_t)p.wpc * p.warp_stride);  // [wpc][M][32]
  // g+32 zero halves read by the MMA lanes whose batch row does not exist.  Shared by the CTA: every warp stores the
  // same zeros (benign) and only needs its own stores to be visible, so no CTA-wide barrier is required.
  __half* zpad = reinterpret_cast<__half*>(red + (size_t)p.wpc * p.M * 32);

  const uint8_t* src = p.recs + ((size_t)strip * p.G + g0) * rec_bytes;
  if (lane == 0) {
    for (int s = 0; s < nst; ++s) mbar_init(bars + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // weights are constants: fill the ring right away (under PDL this overlaps the previous kernel's tail)
    const int pre = min(nst, ngw);
    for (int i = 0; i < pre; ++i) {
      mbar_expect_tx(bars + 8u * i, (uint32_t)rec_bytes);
      bulk_load(ring_u32 + i * rec_bytes, src + (size_t)i * rec_bytes, (uint32_t)rec_bytes, bars + 8u * i);
    }
  }
  if (p.pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  // stage this warp's slice of x: fp16, order [x0,x4,x1/16,x5/16,x2,x6,x3/16,x7/16] per 8 (the /16 lets the nibbles at
  // bits 4-7 / 12-15 be used in place as q*2^-20), plus X_g = sum of each group's (fp16-rounded) activations in fp32
  {
    const int64_t kbase = (int64_t)g0 * g;
    const int ksz8 = ngw * (g >> 3);
    const int cpg = g >> 3;  // 8-element chunks per group (4, 8, 16, 32, ...)
    const __half2 sixteenth = __float2half2_rn(0.0625f);
    for (int m = 0; m < p.M; ++m) {
      const int64_t row = (int64_t)m * p.K + kbase;
      float gsum = 0.f;
      int gcur = 0;
      for (int kb = 0; kb < ksz8; kb += 32) {
        const int kc = kb + lane;
        float csum = 0.f;
        if (kc < ksz8) {
          float v[8];
          load8(p.x, p.x_dtype, row + kc * 8, v);
          if (p.input_scale) {
            float sc8[8];
            load8(p.input_scale, B200WOQ_F32, kbase + kc * 8, sc8);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= sc8[i];
          }
          __half2 h0 = __floats2half2_rn(v[0], v[4]), h1 = __floats2half2_rn(v[1], v[5]);
          __half2 h2 = __floats2half2_rn(v[2], v[6]), h3 = __floats2half2_rn(v[3], v[7]);
          const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
          csum = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
          h1 = __hmul2(h1, sixteenth);
          h3 = __hmul2(h3, sixteenth);
          uint4 o;
          o.x = *reinterpret_cast<uint32_t*>(&h0);
          o.y = *reinterpret_cast<uint32_t*>(&h1);
          o.z = *reinterpret_cast<uint32_t*>(&h2);
          o.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(xs + m * p.xs_ld + kc * 8) = o;
        }
        if (NI_T == 4) {  // 16 chunks per group: two groups per 32-lane pass
          csum += __shfl_xor_sync(0xffffffffu, csum, 8);
          csum += __shfl_xor_sync(0xffffffffu, csum, 4);
          csum += __shfl_xor_sync(0xffffffffu, csum, 2);
          csum += __shfl_xor_sync(0xffffffffu, csum, 1);
          if ((lane & 15) == 0 && kc < ksz8) xsum[m * p.gw_max + (kc >> 4)] = csum;
        } else if (cpg >= 32) {
          gsum += warp_sum(csum);
          if (((kb + 32) % cpg) == 0) {
            if (lane == 0) xsum[m * p.gw_max + gcur] = gsum;
            gsum = 0.f;
            ++gcur;
          }
        } else {
          for (int o = cpg >> 1; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
          if ((lane % cpg) == 0 && kc < ksz8) xsum[m * p.gw_max + kc / cpg] = csum;
        }
      }
    }
    for (int e = lane * 8; e < g + 32; e += 256) *reinterpret_cast<uint4*>(zpad + e) = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();

  float acc[2][4], accg[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = accg[a][c] = 0.f;
  // B fragment rows are batch rows: lanes whose row does not exist read the zero pad and never advance
  const bool live = gq < p.M;
  const __half* xptr = (live ? xs + gq * p.xs_ld : zpad) + t * 8;
  const int xstep = live ? g : 0;
  const int m0 = 2 * t;
  const float* xsum0 = xsum + min(m0, p.M - 1) * p.gw_max;
  const float* xsum1 = xsum + min(m0 + 1, p.M - 1) * p.gw_max;
  const float keep0 = (m0 < p.M) ? 1.f : 0.f, keep1 = (MODE != 0 && m0 + 1 < p.M) ? 1.f : 0.f;

  int s = 0;
  uint32_t phase = 0;
  for (int i = 0; i < ngw; ++i) {
    mbar_wait(bars + 8u * s, phase);
    const uint8_t* rec = ring + s * rec_bytes;
    const uint2 sc = *reinterpret_cast<const uint2*>(rec + NI * 512 + gq * 8);
    const uint32_t zq = *reinterpret_cast<const uint32_t*>(rec + NI * 512 + 64 + gq * 4);
    const float x0 = xsum0[i] * keep0;
    const float x1 = (MODE != 0) ? xsum1[i] * keep1 : 0.f;
    const uint4* wsrc = reinterpret_cast<const uint4*>(rec) + lane;
    auto step = [&](int it, auto first) {
      const uint4 wv = wsrc[it * 32];
      const uint4 v = *reinterpret_cast<const uint4*>(xptr + it * 32);
      const uint32_t wr[4] = {wv.x, wv.y, wv.z, wv.w};
      uint32_t P[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t w8 = wr[r] >> 8;
        P[r][0] = wr[r] & 0x000f000fu;  // (c0, c4) * 2^-24
        P[r][1] = wr[r] & 0x00f000f0u;  // (c1, c5) * 2^-20
        P[r][2] = w8 & 0x000f000fu;     // (c2, c6) * 2^-24
        P[r][3] = w8 & 0x00f000f0u;     // (c3, c7) * 2^-20
      }
      if (decltype(first)::value) {
        mma_16816_zero(accg[0], P[0][0], P[1][0], P[0][1], P[1][1], v.x, v.y);
        mma_16816_zero(accg[1], P[2][0], P[3][0], P[2][1], P[3][1], v.x, v.y);
      } else {
        mma_16816(accg[0], P[0][0], P[1][0], P[0][1], P[1][1], v.x, v.y);
        mma_16816(accg[1], P[2][0], P[3][0], P[2][1], P[3][1], v.x, v.y);
      }
      mma_16816(accg[0], P[0][2], P[1][2], P[0][3], P[1][3], v.z, v.w);
      mma_16816(accg[1], P[2][2], P[3][2], P[2][3], P[3][3], v.z, v.w);
    };
    step(0, std::true_type{});
    if (NI_T) {
#pragma unroll
      for (int it = 1; it < NI_T; ++it) step(it, std::false_type{});
    } else {
      for (int it = 1; it < NI; ++it) step(it, std::false_type{});
    }
    xptr += xstep;
    // the stage is consumed (its words are in registers): refill it with record i + nst
    if (i + nst < ngw) {
      __syncwarp();
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bars + 8u * s, (uint32_t)rec_bytes);
        bulk_load(ring_u32 + s * rec_bytes, src + (size_t)(i + nst) * rec_bytes, (uint32_t)rec_bytes, bars + 8u * s);
      }
    }
    if (++s == nst) {
      s = 0;
      phase ^= 1u;
    }
    // group epilogue: acc += scale * (2^24 * acc_g - zp * X_g)
    const __half2 s01 = *reinterpret_cast<const __half2*>(&sc.x), s23 = *reinterpret_cast<const __half2*>(&sc.y);
    const float sf[4] = {__low2float(s01), __high2float(s01), __low2float(s23), __high2float(s23)};
    const float zf[4] = {(float)(zq & 0xffu), (float)((zq >> 8) & 0xffu), (float)((zq >> 16) & 0xffu), (float)(zq >> 24)};
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      const float sa = sf[2 * tile], sb = sf[2 * tile + 1], za = zf[2 * tile], zb = zf[2 * tile + 1];
      acc[tile][0] = fmaf(fmaf(accg[tile][0], 16777216.f, -za * x0), sa, acc[tile][0]);
      acc[tile][2] = fmaf(fmaf(accg[tile][2], 16777216.f, -zb * x0), sb, acc[tile][2]);
      if (MODE != 0) {  // batch rows 2t+1 exist only when M > 1
        acc[tile][1] = fmaf(fmaf(accg[tile][1], 16777216.f, -za * x1), sa, acc[tile][1]);
        acc[tile][3] = fmaf(fmaf(accg[tile][3], 16777216.f, -zb * x1), sb, acc[tile][3]);
      }
    }
  }

  // in-CTA reduction over the wpc warps (fixed order -> deterministic)
#pragma unroll
  for (int half = 0; half < (MODE == 0 ? 1 : 2); ++half) {
    const int m = 2 * t + half;
    if (m < p.M)
      *reinterpret_cast<float4*>(red + ((size_t)warp * p.M + m) * 32 + 4 * gq) =
          make_float4(acc[0][half], acc[0][2 + half], acc[1][half], acc[1][2 + half]);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < p.M * 32; e += blockDim.x) {
    const int m = e >> 5, nl = e & 31;
    const int n = strip * 32 + nl;
    float v = 0.f;
    for (int w = 0; w < p.wpc; ++w) v += red[((size_t)w * p.M + m) * 32 + nl];
    if (p.bias) v += load_as_float(p.bias, p.bias_dtype, n);
    store_from_float(p.y, p.y_dtype, (int64_t)m * p.N + n, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant (B200WOQ_STREAM_IMPL=1; measured SLOWER than the strip-per-CTA kernel above on the Llama-2-7B
// layers -- 5.5 vs 3.6 us at 4096x4096, equal at 22016x4096 -- because a decode layer is dominated by fixed per-launch
// latency, not by balance; kept for the record and for M <= 8): one CTA per SM, each owning a contiguous range of WHOLE strips (no cross-CTA
// reduction, deterministic); the CTA's records (strip-major = one contiguous byte range) are cut into W equal
// contiguous pieces, one per warp, regardless of strip boundaries, so every warp streams the same number of bytes
// whatever G is (the strip-per-CTA kernel leaves warps idle when G % warps != 0 and whole SMs idle when
// strips % 148 != 0).  A warp that crosses a strip boundary parks its partial sums in a slot; one __syncthreads at
// the end, then a fixed-order reduction.  The footprint is kept under HALF an SM (<= 512 threads, <= ~110 KB of shared
// memory) so that under programmatic dependent launch the NEXT layer's CTAs are co-resident and fill their rings from
// HBM while this layer is still computing: the ~2 us dependent-launch latency and the ring-fill latency are hidden
// behind the previous layer instead of serialising with it.
struct PParams {
  const void* x;
  int x_dtype;
  int M, K, N;
  const uint8_t* recs;
  int NI, g, G;
  const void* bias;
  int bias_dtype;
  const float* input_scale;
  void* y;
  int y_dtype;
  int W;          // warps per CTA
  int nst;        // ring depth per warp
  int n_strips;
  int xs_ld;      // halves per staged x row (K + 32: 64-byte skew between batch rows)
  int pdl;
};
constexpr int kSlots = 4;  // strip segments a warp's record range may touch

template <int MODE, int NI_T>
__global__ void __launch_bounds__(1024) woq_gemm_persist_kernel(const PParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;
  const int NI = NI_T ? NI_T : p.NI;
  const int g = NI * 32;
  const int rec_bytes = NI * 512 + 96;
  const int W = p.W, nst = p.nst, G = p.G;
  // strips of this CTA, records of this warp (CTA-linear, strip-major)
  const int s0 = (int)(((int64_t)p.n_strips * blockIdx.x) / gridDim.x);
  const int s1 = (int)(((int64_t)p.n_strips * (blockIdx.x + 1)) / gridDim.x);
  const int R = (s1 - s0) * G;
  const int wa = (int)(((int64_t)R * warp) / W), wb = (int)(((int64_t)R * (warp + 1)) / W);
  const int nrec = wb - wa;
  // shared memory: [W rings: nst*rec_bytes][W*nst barriers][xs: M x xs_ld f16][xsum: M x G f32][slots][seg_first][zpad]
  uint8_t* ring = smem_raw + (size_t)warp * nst * rec_bytes;
  const uint32_t ring_u32 = smem_u32(ring);
  uint8_t* after_rings = smem_raw + (size_t)W * nst * rec_bytes;
  const uint32_t bars = smem_u32(after_rings) + (uint32_t)warp * nst * 8u;
  __half* xs = reinterpret_cast<__half*>(after_rings + (((size_t)W * nst * 8 + 127) & ~(size_t)127));
  float* xsum = reinterpret_cast<float*>(xs + (size_t)p.M * p.xs_ld);
  float* slots = xsum + (((size_t)p.M * G + 3) & ~(size_t)3);                  // [W][kSlots][M][32]
  int* seg_first = reinterpret_cast<int*>(slots + (size_t)W * kSlots * p.M * 32);  // [W] first local strip, [W] count
  __half* zpad = reinterpret_cast<__half*>(seg_first + ((2 * W + 3) & ~3));  // keep the 16-byte alignment

  const uint8_t* src = p.recs + ((size_t)s0 * G + wa) * rec_bytes;
  if (lane == 0) {
    for (int s = 0; s < nst; ++s) mbar_init(bars + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int pre = min(nst, nrec);
    for (int i = 0; i < pre; ++i) {  // weights are constants: stream them before the dependency is resolved
      mbar_expect_tx(bars + 8u * i, (uint32_t)rec_bytes);
      bulk_load(ring_u32 + i * rec_bytes, src + (size_t)i * rec_bytes, (uint32_t)rec_bytes, bars + 8u * i);
    }
  }
  if (p.pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  // ---- stage x once per CTA (all warps, 8-element chunks round-robin): fp16 in the permuted / pre-scaled order of
  // the strip kernel, plus X_g = per-group sums of the fp16-rounded activations
  {
    const int cpg = g >> 3;                 // chunks per group
    const int nchunk = p.K >> 3;
    const __half2 sixteenth = __float2half2_rn(0.0625f);
    for (int m = 0; m < p.M; ++m) {
      const int64_t row = (int64_t)m * p.K;
      for (int kb = warp * 32; kb < nchunk; kb += W * 32) {
        const int kc = kb + lane;
        float csum = 0.f;
        if (kc < nchunk) {
          float v[8];
          load8(p.x, p.x_dtype, row + (int64_t)kc * 8, v);
          if (p.input_scale) {
            float sc8[8];
            load8(p.input_scale, B200WOQ_F32, (int64_t)kc * 8, sc8);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= sc8[i];
          }
          __half2 h0 = __floats2half2_rn(v[0], v[4]), h1 = __floats2half2_rn(v[1], v[5]);
          __half2 h2 = __floats2half2_rn(v[2], v[6]), h3 = __floats2half2_rn(v[3], v[7]);
          const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
          csum = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
          h1 = __hmul2(h1, sixteenth);
          h3 = __hmul2(h3, sixteenth);
          uint4 o;
          o.x = *reinterpret_cast<uint32_t*>(&h0);
          o.y = *reinterpret_cast<uint32_t*>(&h1);
          o.z = *reinterpret_cast<uint32_t*>(&h2);
          o.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(xs + (size_t)m * p.xs_ld + (size_t)kc * 8) = o;
        }
        // group sums: cpg chunks per group (the host only takes this kernel when cpg is a power of two <= 32)
        {
          for (int o = cpg >> 1; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
          if ((lane & (cpg - 1)) == 0 && kc < nchunk) xsum[(size_t)m * G + kc / cpg] = csum;
        }
      }
    }
    for (int e = threadIdx.x * 8; e < g + 32; e += blockDim.x * 8) *reinterpret_cast<uint4*>(zpad + e) = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();

  float acc[2][4], accg[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = accg[a][c] = 0.f;
  const bool live = gq < p.M;
  const __half* xrow = (live ? xs + (size_t)gq * p.xs_ld : zpad) + t * 8;
  const int m0 = 2 * t;
  const float* xsum0 = xsum + (size_t)min(m0, p.M - 1) * G;
  const float* xsum1 = xsum + (size_t)min(m0 + 1, p.M - 1) * G;
  const float keep0 = (m0 < p.M) ? 1.f : 0.f, keep1 = (MODE != 0 && m0 + 1 < p.M) ? 1.f : 0.f;

  int sl = wa / G;            // current local strip
  int gi = wa - sl * G;       // current group within the strip
  const int first_sl = sl;
  int nseg = 0;
  auto flush = [&]() {
    float* dst = slots + ((size_t)(warp * kSlots + nseg) * p.M) * 32;
#pragma unroll
    for (int half = 0; half < (MODE == 0 ? 1 : 2); ++half) {
      const int m = 2 * t + half;
      if (m < p.M)
        *reinterpret_cast<float4*>(dst + (size_t)m * 32 + 4 * gq) =
            make_float4(acc[0][half], acc[0][2 + half], acc[1][half], acc[1][2 + half]);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    ++nseg;
  };

  int s = 0;
  uint32_t phase = 0;
  for (int i = 0; i < nrec; ++i) {
    mbar_wait(bars + 8u * s, phase);
    const uint8_t* rec = ring + s * rec_bytes;
    const uint2 sc = *reinterpret_cast<const uint2*>(rec + NI * 512 + gq * 8);
    const uint32_t zq = *reinterpret_cast<const uint32_t*>(rec + NI * 512 + 64 + gq * 4);
    const float x0 = xsum0[gi] * keep0;
    const float x1 = (MODE != 0) ? xsum1[gi] * keep1 : 0.f;
    const uint4* wsrc = reinterpret_cast<const uint4*>(rec) + lane;
    const __half* xptr = live ? xrow + (size_t)gi * g : xrow;
    auto step = [&](int it, auto first) {
      const uint4 wv = wsrc[it * 32];
      const uint4 v = *reinterpret_cast<const uint4*>(xptr + it * 32);
      const uint32_t wr[4] = {wv.x, wv.y, wv.z, wv.w};
      uint32_t P[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t w8 = wr[r] >> 8;
        P[r][0] = wr[r] & 0x000f000fu;
        P[r][1] = wr[r] & 0x00f000f0u;
        P[r][2] = w8 & 0x000f000fu;
        P[r][3] = w8 & 0x00f000f0u;
      }
      if (decltype(first)::value) {
        mma_16816_zero(accg[0], P[0][0], P[1][0], P[0][1], P[1][1], v.x, v.y);
        mma_16816_zero(accg[1], P[2][0], P[3][0], P[2][1], P[3][1], v.x, v.y);
      } else {
        mma_16816(accg[0], P[0][0], P[1][0], P[0][1], P[1][1], v.x, v.y);
        mma_16816(accg[1], P[2][0], P[3][0], P[2][1], P[3][1], v.x, v.y);
      }
      mma_16816(accg[0], P[0][2], P[1][2], P[0][3], P[1][3], v.z, v.w);
      mma_16816(accg[1], P[2][2], P[3][2], P[2][3], P[3][3], v.z, v.w);
    };
    step(0, std::true_type{});
    if (NI_T) {
#pragma unroll
      for (int it = 1; it < NI_T; ++it) step(it, std::false_type{});
    } else {
      for (int it = 1; it < NI; ++it) step(it, std::false_type{});
    }
    if (i + nst < nrec) {  // the stage is consumed (its words are in registers): refill it
      __syncwarp();
      if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bars + 8u * s, (uint32_t)rec_bytes);
        bulk_load(ring_u32 + s * rec_bytes, src + (size_t)(i + nst) * rec_bytes, (uint32_t)rec_bytes, bars + 8u * s);
      }
    }
    if (++s == nst) {
      s = 0;
      phase ^= 1u;
    }
    const __half2 s01 = *reinterpret_cast<const __half2*>(&sc.x), s23 = *reinterpret_cast<const __half2*>(&sc.y);
    const float sf[4] = {__low2float(s01), __high2float(s01), __low2float(s23), __high2float(s23)};
    const float zf[4] = {(float)(zq & 0xffu), (float)((zq >> 8) & 0xffu), (float)((zq >> 16) & 0xffu), (float)(zq >> 24)};
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      const float sa = sf[2 * tile], sb = sf[2 * tile + 1], za = zf[2 * tile], zb = zf[2 * tile + 1];
      acc[tile][0] = fmaf(fmaf(accg[tile][0], 16777216.f, -za * x0), sa, acc[tile][0]);
      acc[tile][2] = fmaf(fmaf(accg[tile][2], 16777216.f, -zb * x0), sb, acc[tile][2]);
      if (MODE != 0) {
        acc[tile][1] = fmaf(fmaf(accg[tile][1], 16777216.f, -za * x1), sa, acc[tile][1]);
        acc[tile][3] = fmaf(fmaf(accg[tile][3], 16777216.f, -zb * x1), sb, acc[tile][3]);
      }
    }
    if (++gi == G) {  // strip finished for this warp: park the partial sums
      flush();
      gi = 0;
      ++sl;
    }
  }
  if (gi != 0 && nrec > 0) flush();
  if (lane == 0) {
    seg_first[warp] = first_sl;
    seg_first[W + warp] = nseg;
  }
  __syncthreads();
  // fixed-order reduction over the warps' segments -> deterministic
  const int nloc = s1 - s0;
  for (int e = threadIdx.x; e < nloc * p.M * 32; e += blockDim.x) {
    const int nl = e & 31, m = (e >> 5) % p.M, ls = (e >> 5) / p.M;
    float v = 0.f;
    for (int w = 0; w < W; ++w) {
      const int k = ls - seg_first[w];
      if (k >= 0 && k < seg_first[W + w]) v += slots[((size_t)(w * kSlots + k) * p.M + m) * 32 + nl];
    }
    const int n = (s0 + ls) * 32 + nl;
    if (p.bias) v += load_as_float(p.bias, p.bias_dtype, n);
    store_from_float(p.y, p.y_dtype, (int64_t)m * p.N + n, v);
  }
}

// optimum-format tensors -> strip records [N/32][G]: [NI x 32 lanes x int4 | 32 fp16 scales | 32 u8 zero-points]
__global__ void build_stream_kernel(const int32_t* __restrict__ qweight, const int32_t* __restrict__ qzeros,
                                    const __half* __restrict__ scales, int N, int K, int g, int G, int NI, int rec_bytes,
                                    uint8_t* __restrict__ out) {
  const int n_strips = N / 32;
  const int64_t pieces_per_rec = rec_bytes / 16;
  const int64_t total = (int64_t)n_strips * G * pieces_per_rec;
  const int Nw = N / 8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t rec = idx / pieces_per_rec;
    const int piece = (int)(idx - rec * pieces_per_rec);
    const int strip = (int)(rec / G), gi = (int)(rec - (int64_t)strip * G);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (piece < NI * 32) {  // packed words: [iter][lane]
      const int it = piece / 32, lane = piece & 31;
      const int gq = lane >> 2, t = lane & 3;
      const int n = strip * 32 + 4 * gq;
      const int kw = gi * (g / 8) + 4 * it + t;
      v = *reinterpret_cast<const uint4*>(qweight + (int64_t)kw * N + n);
    } else if (piece < NI * 32 + 4) {  // 32 fp16 scales
      const int n = strip * 32 + (piece - NI * 32) * 8;
      v = *reinterpret_cast<const uint4*>(scales + (int64_t)gi * N + n);
    } else {  // 32 zero-points as bytes, already +1 and wrapped (modules.py:363, 409-410)
      const int n = strip * 32 + (piece - NI * 32 - 4) * 16;
      uint32_t o[4] = {0, 0, 0, 0};
      const uint32_t w0 = (uint32_t)qzeros[(int64_t)gi * Nw + n / 8], w1 = (uint32_t)qzeros[(int64_t)gi * Nw + n / 8 + 1];
      for (int e = 0; e < 16; ++e) {
        const uint32_t nib = ((e < 8 ? w0 : w1) >> (4 * (e & 7))) & 0xfu;
        o[e >> 2] |= ((nib + 1u) & 0xfu) << (8 * (e & 3));
      }
      v = make_uint4(o[0], o[1], o[2], o[3]);
    }
    *reinterpret_cast<uint4*>(out + rec * rec_bytes + (int64_t)piece * 16) = v;
  }
}

static size_t warp_smem_bytes(int M, int rec_bytes, int gw_max, int g, int nst) {
  const size_t b = (size_t)nst * rec_bytes + ((nst * 8 + 15) & ~15) + (size_t)M * (gw_max * g + 32) * 2 +
                   (((size_t)M * gw_max + 3) & ~(size_t)3) * 4;
  return (b + 127) & ~(size_t)127;
}

}  // namespace stream
}  // namespace b200woq

using namespace b200woq;

static bool stream_shape_ok(int64_t N, int64_t K, int bits, int g) {
  return bits == 4 && g > 0 && g % 32 == 0 && g <= 1024 && K % g == 0 && N % 32 == 0 && N < (1ll << 30) && K < (1ll << 30);
}

extern "C" int64_t b200woq_stream_layout_bytes(int64_t N, int64_t K, int bits, int group_size) {
  const int g = eff_group(K, group_size);
  if (!stream_shape_ok(N, K, bits, g)) return 0;
  const int64_t rec = (int64_t)(g / 32) * 512 + 96;
  return (N / 32) * (K / g) * rec;
}

extern "C" int b200woq_build_stream_layout(const int32_t* qweight, const int32_t* qzeros, const void* scales16, int64_t N,
                                           int64_t K, int bits, int group_size, void* out, void* stream_) {
  const int g = eff_group(K, group_size);
  WOQ_CHECK_ARG(qweight && qzeros && scales16 && out, "build_stream_layout: null pointer");
  WOQ_CHECK_ARG(stream_shape_ok(N, K, bits, g), "build_stream_layout: unsupported shape (4-bit, g%%32==0, N%%32==0 only)");
  const int NI = g / 32, rec = NI * 512 + 96;
  const int64_t total = (N / 32) * (K / g) * (rec / 16);
  int64_t blocks = ceil_div(total, 256);
  if (blocks > (int64_t)num_sms() * 32) blocks = (int64_t)num_sms() * 32;
  stream::build_stream_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(
      qweight, qzeros, (const __half*)scales16, (int)N, (int)K, g, (int)(K / g), NI, rec, (uint8_t*)out);
  WOQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200woq_linear_forward_stream(const void* x, int x_dtype, int64_t M, int64_t K, int64_t N,
                                             const void* stream_layout, const void* bias, int bias_dtype,
                                             const float* input_scale, void* y, int y_dtype, int bits, int group_size,
                                             int flags, void* stream_) {
  using namespace stream;
  const int g = eff_group(K, group_size);
  WOQ_CHECK_ARG(x && stream_layout && y && M > 0, "linear_forward_stream: null pointer / empty batch");
  WOQ_CHECK_ARG(stream_shape_ok(N, K, bits, g), "linear_forward_stream: unsupported shape");
  WOQ_CHECK_ARG(M <= 8, "linear_forward_stream: M must be <= 8 (use b200woq_linear_forward)");
  Params p = {};
  p.x = x;
  p.x_dtype = x_dtype;
  p.M = (int)M;
  p.K = (int)K;
  p.N = (int)N;
  p.recs = (const uint8_t*)stream_layout;
  p.NI = g / 32;
  p.rec_bytes = p.NI * 512 + 96;
  p.g = g;
  p.G = (int)(K / g);
  p.bias = bias;
  p.bias_dtype = bias_dtype;
  p.input_scale = input_scale;
  p.y = y;
  p.y_dtype = y_dtype;
  p.pdl = (flags & 2) ? 1 : 0;
  // ---- persistent kernel (default): see woq_gemm_persist_kernel
  static const int impl = getenv("B200WOQ_STREAM_IMPL") ? atoi(getenv("B200WOQ_STREAM_IMPL")) : 0;
  if (impl == 1 || M > 4) {  // the strip kernel is tuned (and shared-memory sized) for M <= 4
    PParams q = {};
    q.x = x; q.x_dtype = x_dtype; q.M = (int)M; q.K = (int)K; q.N = (int)N;
    q.recs = (const uint8_t*)stream_layout; q.NI = g / 32; q.g = g; q.G = (int)(K / g);
    q.bias = bias; q.bias_dtype = bias_dtype; q.input_scale = input_scale; q.y = y; q.y_dtype = y_dtype;
    q.pdl = p.pdl; q.n_strips = (int)(N / 32); q.xs_ld = (int)K + 32;
    const int sms_ = num_sms();
    const int ctas = std::min(sms_, q.n_strips);
    const int max_strips = (int)ceil_div(q.n_strips, ctas);
    static const int env_w = getenv("B200WOQ_PERSIST_WARPS") ? atoi(getenv("B200WOQ_PERSIST_WARPS")) : 16;
    static const int env_n = getenv("B200WOQ_PERSIST_NST") ? atoi(getenv("B200WOQ_PERSIST_NST")) : 0;
    int Wp = std::max(1, std::min(env_w, 32));
    const int rec_b = q.NI * 512 + 96;
    // every warp needs work, and its contiguous range may touch at most kSlots strips
    while (Wp > 1 && (int64_t)max_strips * q.G < 2ll * Wp) Wp /= 2;
    auto segs = [&](int w) { return (int)(ceil_div((int64_t)max_strips * q.G, w) / q.G) + 2; };
    auto smem_for = [&](int w, int n) {
      size_t b = (size_t)w * n * rec_b;
      b += ((size_t)w * n * 8 + 127) & ~(size_t)127;
      b += (size_t)q.M * q.xs_ld * 2;
      b += (((size_t)q.M * q.G + 3) & ~(size_t)3) * 4;
      b += (size_t)w * kSlots * q.M * 32 * 4 + (size_t)((2 * w + 3) & ~3) * 4 + (size_t)(g + 32) * 2;
      return (b + 127) & ~(size_t)127;
    };
    const size_t half_sm = (size_t)(227 * 1024) / 2 - 1024, full_sm = 226 * 1024;
    const int cpg_ = g >> 3;
    const bool cpg_ok = cpg_ <= 32 && (cpg_ & (cpg_ - 1)) == 0;  // the staging pass reduces X_g with in-warp shuffles
    if (cpg_ok && segs(Wp) <= kSlots && smem_for(Wp, 1) <= full_sm) {
      int n = env_n > 0 ? env_n : 4;
      n = std::min<int>(n, kMaxStages);
      while (n > 2 && smem_for(Wp, n) > half_sm) --n;         // prefer the co-resident (half-SM) footprint
      while (n > 1 && smem_for(Wp, n) > full_sm) --n;
      q.W = Wp; q.nst = n;
      const size_t smem = smem_for(Wp, n);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)ctas);
      cfg.blockDim = dim3((unsigned)(32 * Wp));
      cfg.dynamicSmemBytes = smem;
      cfg.stream = (cudaStream_t)stream_;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = q.pdl ? 1 : 0;
#define PERSIST_LAUNCH(MODE_, NI_)                                                                                  \
  do {                                                                                                              \
    WOQ_CUDA(cudaFuncSetAttribute(woq_gemm_persist_kernel<MODE_, NI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem));                                                                      \
    WOQ_CUDA(cudaLaunchKernelEx(&cfg, woq_gemm_persist_kernel<MODE_, NI_>, q));                                     \
  } while (0)
      if (M == 1) {
        if (q.NI == 4) PERSIST_LAUNCH(0, 4); else PERSIST_LAUNCH(0, 0);
      } else {
        if (q.NI == 4) PERSIST_LAUNCH(1, 4); else PERSIST_LAUNCH(1, 0);
      }
#undef PERSIST_LAUNCH
      count_launch(1);
      return 0;
    }
  }
  // One CTA per 32-column strip; its wpc warps split K.  Every CTA of the grid must be resident at once (a second
  // wave would serialise behind the first): that caps warps per CTA at 64 / ceil(strips / SMs) and shared memory per
  // CTA at its share of the SM.  Within those caps: as many warps as keep >= 2 groups each (up to ~12 streaming warps
  // per SM), then the deepest ring that fits.
  const int n_strips = (int)(N / 32);
  const int sms = num_sms();
  static const int env_wpc = getenv("B200WOQ_STREAM_WPC") ? atoi(getenv("B200WOQ_STREAM_WPC")) : 0;
  static const int env_nst = getenv("B200WOQ_STREAM_NST") ? atoi(getenv("B200WOQ_STREAM_NST")) : 0;
  const int cta_per_sm = (int)ceil_div(n_strips, sms);
  size_t budget = (size_t)(227 * 1024) / cta_per_sm - 1024;  // 1 KB reserved per CTA by the driver
  int wpc_cap = 32;  // K = 11008: 32 warps x 2.7 groups beat 16 x 5.4 (7.4 vs 8.3 us): more chains in flight per SM
  while (wpc_cap > 1 && wpc_cap * cta_per_sm > 64) wpc_cap /= 2;
  auto total_smem = [&](int w, int nst) {
    return (size_t)w * warp_smem_bytes(p.M, p.rec_bytes, (int)ceil_div(p.G, w), g, nst) + (size_t)w * p.M * 32 * 4 +
           (size_t)(g + 32) * 2;
  };
  int wpc = std::min(4, wpc_cap);
  while (wpc < wpc_cap && (int64_t)n_strips * wpc < 12ll * sms && p.G / (wpc * 2) >= 2) wpc *= 2;
  if (env_wpc > 0) wpc = std::min(env_wpc, 32);
  while (wpc > 1 && p.G < wpc) wpc /= 2;
  int nst = 0;
  for (int pass = 0; pass < 2 && nst == 0; ++pass) {
    for (int w = wpc; w >= 1 && nst == 0; w /= 2) {
      const int need = (int)std::min<int64_t>(kMaxStages, ceil_div(p.G, w));
      int n = env_nst > 0 ? std::min(env_nst, kMaxStages) : need;
      while (n > std::min(2, need) && total_smem(w, n) > budget) --n;
      if (total_smem(w, n) <= budget) {
        wpc = w;
        nst = n;
      }
    }
    if (nst == 0) budget = 226 * 1024;  // cannot be single-wave: take what fits on an SM
  }
  if (nst == 0) {
    set_error("linear_forward_stream: K slice does not fit in shared memory");
    return B200WOQ_EUNSUPPORTED;
  }
  p.wpc = wpc;
  p.nst = nst;
  p.gw_max = (int)ceil_div(p.G, wpc);
  p.xs_ld = p.gw_max * g + 32;
  p.warp_stride = (int)warp_smem_bytes(p.M, p.rec_bytes, p.gw_max, g, nst);
  // Ask for the whole per-CTA share of the SM's shared memory: the block scheduler then places exactly cta_per_sm CTAs
  // on every SM.  With a smaller footprint, CTAs of the next (programmatically dependent) launch double up on some SMs
  // and those SMs finish late (measured: 4.5 us vs 3.6 us per 4096x4096 layer).
  const size_t smem = std::max(total_smem(wpc, nst), std::min(budget, (size_t)(226 * 1024)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n_strips);
  cfg.blockDim = dim3((unsigned)(32 * wpc));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream_;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
#define STREAM_LAUNCH(MODE_, NI_)                                                                                  \
  do {                                                                                                              \
    WOQ_CUDA(cudaFuncSetAttribute(woq_gemm_stream_kernel<MODE_, NI_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                  (int)smem));                                                                      \
    WOQ_CUDA(cudaLaunchKernelEx(&cfg, woq_gemm_stream_kernel<MODE_, NI_>, p));                                      \
  } while (0)
  if (M == 1) {
    if (p.NI == 4) STREAM_LAUNCH(0, 4); else STREAM_LAUNCH(0, 0);
  } else {
    if (p.NI == 4) STREAM_LAUNCH(1, 4); else STREAM_LAUNCH(1, 0);
  }
#undef STREAM_LAUNCH
  count_launch(1);
  return 0;
}
