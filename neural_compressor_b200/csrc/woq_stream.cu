// woq_stream.cu -- K6 at small batch (M <= 16, 4-bit): the packed weights are streamed HBM -> shared memory by the
// TMA engine (cp.async.bulk + mbarrier ring) from a B200-native "stream layout" of the optimum-format tensors.
//
// Why a derived layout: in the reference's optimum format consecutive k-rows of qweight are N*4 bytes apart, so a
// CTA that owns 128 columns touches one 512-byte piece of every row (DRAM-row unfriendly, and every 16-byte piece
// needs its own load instruction).  The stream layout stores, for every (128-column tile, quantisation group), ONE
// contiguous record
//     [ 4 strips x NI iterations x 32 lanes x int4 packed words | 128 fp16 scales | 128 u8 zero-points ]
// with the words already in the order the MMA lanes consume them.  A single elected thread keeps a ring of NST
// records in flight per CTA (NST x 8.4 KB, independent of registers), weights are constants so under programmatic
// dependent launch the ring fills while the previous layer is still finishing, and the consumer warps read their
// int4 with conflict-free LDS.128.  The derived copy is built once per module (like the reference caches its
// de-quantised fp32 weight at first forward, modules.py:603-604); the checkpoint tensors stay in optimum format.
//
// Arithmetic = the SUB path of woq_gemm.cu: codes enter mma.sync.m16n8k16 as fp16 subnormals; here the nibbles at
// bit positions 4-7 / 12-15 are used in place (value q*2^-20) and the matching activations are pre-scaled by 2^-4
// when x is staged, which removes the shifts: 5 integer ops per 8 codes.
#include <cooperative_groups.h>

#include <type_traits>

#include "common.cuh"

namespace b200woq {
namespace stream {

constexpr int kStages = 8;
constexpr int kRedPerM = 320;  // floats of split-K exchange buffer per batch row (S * slice <= 160)

struct Params {
  const void* x;
  int x_dtype;
  int M, K, N;
  const uint8_t* recs;  // [n_tiles][G] records
  int rec_bytes;        // NI*2048 + 256 + 128
  int NI;               // 16-byte loads per lane per group = g / 32
  int g, G;
  const void* bias;
  int bias_dtype;
  const float* input_scale;
  void* y;
  int y_dtype;
  int S, slice, gmax, xs_ld;
  int pdl;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

// grid = (S, n_tiles), cluster (S,1,1); block = 160: warps 0-3 consume one 32-column strip each, warp 4 produces.
template <int MT>
__global__ void __launch_bounds__(160, 3) woq_gemm_stream_kernel(const Params p) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // [ ring: kStages * rec_bytes ][ barriers 2*kStages*8 ][ red: M*kRedPerM f32 ][ xs: M*xs_ld f16 ][ xsum: M*gmax f32 ]
  uint8_t* ring = smem_raw;
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t bars = ring_u32 + kStages * p.rec_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  float* red = reinterpret_cast<float*>(ring + kStages * p.rec_bytes + 2 * kStages * 8);
  __half* xs = reinterpret_cast<__half*>(red + (size_t)p.M * kRedPerM);
  float* xsum = reinterpret_cast<float*>(xs + (size_t)p.M * p.xs_ld);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, rank = blockIdx.x;
  const int n_tile0 = blockIdx.y * 128;
  const int gb = rank * p.G / S, ge = (rank + 1) * p.G / S;
  const int ng = ge - gb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (S > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  __syncthreads();
  // the next kernel may start streaming ITS weights as soon as it finds room; it waits for our results itself
  if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 4) {
    // ---------------- producer: weights are constants, no dependency on the previous kernel ----------------
    if (lane == 0) {
      const uint8_t* src = p.recs + ((size_t)blockIdx.y * p.G + gb) * p.rec_bytes;
      for (int i = 0; i < ng; ++i) {
        const int s = i % kStages;
        mbar_wait(empty_bar(s), (((uint32_t)(i / kStages)) & 1u) ^ 1u);
        mbar_expect_tx(full_bar(s), (uint32_t)p.rec_bytes);
        bulk_load(ring_u32 + s * p.rec_bytes, src + (size_t)i * p.rec_bytes, (uint32_t)p.rec_bytes, full_bar(s));
      }
    }
  } else {
    // ---------------- consumers ----------------
    const int gq = lane >> 2, t = lane & 3;
    const int strip = warp;
    const int n_in_tile = strip * 32 + 4 * gq;
    const bool strip_valid = n_tile0 + strip * 32 < p.N;
    if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    // stage x: fp16, permuted [x0,x4,x1,x5,x2,x6,x3,x7]; the elements that meet the bit-4..7 nibbles (x1,x5,x3,x7) are
    // pre-scaled by 2^-4 so that `w & 0x00f000f0` (= q * 2^-20 as an fp16 subnormal) needs no shift
    {
      const int64_t kbase = (int64_t)gb * p.g;
      const int ksz8 = ng * p.g / 8;
      for (int m = 0; m < p.M; ++m) {
        const int64_t row = (int64_t)m * p.K + kbase;
        for (int kc = threadIdx.x; kc < ksz8; kc += 128) {
          float v[8];
          load8(p.x, p.x_dtype, row + kc * 8, v);
          if (p.input_scale) {
            float sc8[8];
            load8(p.input_scale, B200WOQ_F32, kbase + kc * 8, sc8);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= sc8[i];
          }
          // X_g uses the fp16-rounded values (what the MMA sees for the un-shifted nibbles)
          __half2 h0 = __floats2half2_rn(v[0], v[4]);
          __half2 h2 = __floats2half2_rn(v[2], v[6]);
          __half2 h1 = __floats2half2_rn(v[1], v[5]);
          __half2 h3 = __floats2half2_rn(v[3], v[7]);
          const __half2 sixteenth = __float2half2_rn(0.0625f);
          uint4 o;
          o.x = *reinterpret_cast<uint32_t*>(&h0);
          __half2 h1s = __hmul2(h1, sixteenth), h3s = __hmul2(h3, sixteenth);
          o.y = *reinterpret_cast<uint32_t*>(&h1s);
          o.z = *reinterpret_cast<uint32_t*>(&h2);
          o.w = *reinterpret_cast<uint32_t*>(&h3s);
          *reinterpret_cast<uint4*>(xs + m * p.xs_ld + kc * 8) = o;
        }
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // X_g[m] = sum_k x[m][k] over each group (undo the 2^-4 on the pre-scaled positions), fp32
    for (int task = warp; task < p.M * ng; task += 4) {
      const int m = task / ng, gl = task - m * ng;
      float sum = 0.f;
      for (int e = lane * 8; e < p.g; e += 256) {
        const uint4 v = *reinterpret_cast<const uint4*>(xs + m * p.xs_ld + gl * p.g + e);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
        const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
        sum += (a.x + a.y) + (c.x + c.y) + 16.f * ((b.x + b.y) + (d.x + d.y));
      }
      sum = warp_sum(sum);
      if (lane == 0) xsum[m * p.gmax + gl] = sum;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");

    // NSET independent per-group accumulator sets: the legacy HMMA has a long dependent-issue latency on sm_100, so
    // consecutive MMAs never target the same accumulator (step A / step B x iteration parity); summed at group end
    constexpr int NSET = (MT == 1) ? 4 : 2;
    float acc[2][MT][4], accg[NSET][2][MT][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < MT; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[a][b][c] = 0.f;
#pragma unroll
          for (int q = 0; q < NSET; ++q) accg[q][a][b][c] = 0.f;
        }

    const int NI = p.NI;
    int xoff = t * 8;
    for (int i = 0; i < ng; ++i) {
      const int s = i % kStages;
      mbar_wait(full_bar(s), ((uint32_t)(i / kStages)) & 1u);
      const uint8_t* rec = ring + s * p.rec_bytes;
      if (strip_valid) {
        const uint2 sc = *reinterpret_cast<const uint2*>(rec + NI * 2048 + n_in_tile * 2);
        const uint32_t zq = *reinterpret_cast<const uint32_t*>(rec + NI * 2048 + 256 + n_in_tile);
        const __half2 s01 = *reinterpret_cast<const __half2*>(&sc.x), s23 = *reinterpret_cast<const __half2*>(&sc.y);
        const float sf[4] = {__low2float(s01), __high2float(s01), __low2float(s23), __high2float(s23)};
        const float zf[4] = {(float)(zq & 0xffu), (float)((zq >> 8) & 0xffu), (float)((zq >> 16) & 0xffu), (float)(zq >> 24)};
        const uint4* wsrc = reinterpret_cast<const uint4*>(rec + strip * NI * 512) + lane;
        auto do_it = [&](int it, auto sa_tag) {
          constexpr int sa = decltype(sa_tag)::value, sb = sa + 1;
          const uint4 wv = wsrc[it * 32];
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
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int m = gq + 8 * mt;
            uint32_t xb[4] = {0u, 0u, 0u, 0u};
            if (m < p.M) {
              const uint4 v = *reinterpret_cast<const uint4*>(xs + m * p.xs_ld + xoff);
              xb[0] = v.x; xb[1] = v.y; xb[2] = v.z; xb[3] = v.w;
            }
            mma_16816(accg[sa][0][mt], P[0][0], P[1][0], P[0][1], P[1][1], xb[0], xb[1]);
            mma_16816(accg[sa][1][mt], P[2][0], P[3][0], P[2][1], P[3][1], xb[0], xb[1]);
            mma_16816(accg[sb][0][mt], P[0][2], P[1][2], P[0][3], P[1][3], xb[2], xb[3]);
            mma_16816(accg[sb][1][mt], P[2][2], P[3][2], P[2][3], P[3][3], xb[2], xb[3]);
          }
          xoff += 32;
        };
#pragma unroll 2
        for (int it = 0; it < NI; it += 2) {
          do_it(it, std::integral_constant<int, 0>{});
          if (it + 1 < NI) do_it(it + 1, std::integral_constant<int, (NSET == 4) ? 2 : 0>{});
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int m0 = 8 * mt + 2 * t;
          const float x0 = (m0 < p.M) ? xsum[m0 * p.gmax + i] : 0.f;
          const float x1 = (m0 + 1 < p.M) ? xsum[(m0 + 1) * p.gmax + i] : 0.f;
#pragma unroll
          for (int tile = 0; tile < 2; ++tile) {
            const float sa = sf[2 * tile], sb = sf[2 * tile + 1], za = zf[2 * tile], zb = zf[2 * tile + 1];
            float gsum[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              gsum[c] = accg[0][tile][mt][c];
#pragma unroll
              for (int q = 1; q < NSET; ++q) gsum[c] += accg[q][tile][mt][c];
#pragma unroll
              for (int q = 0; q < NSET; ++q) accg[q][tile][mt][c] = 0.f;
            }
            acc[tile][mt][0] = fmaf(fmaf(gsum[0], 16777216.f, -za * x0), sa, acc[tile][mt][0]);
            acc[tile][mt][1] = fmaf(fmaf(gsum[1], 16777216.f, -za * x1), sa, acc[tile][mt][1]);
            acc[tile][mt][2] = fmaf(fmaf(gsum[2], 16777216.f, -zb * x0), sb, acc[tile][mt][2]);
            acc[tile][mt][3] = fmaf(fmaf(gsum[3], 16777216.f, -zb * x1), sb, acc[tile][mt][3]);
          }
        }
      } else {
        xoff += 32 * NI;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(s));  // this warp is done with the stage
    }

    // partial sums -> the cluster rank that owns the columns (distributed shared memory)
    const int slice = p.slice;
    if (S > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (strip_valid) {
      const int owner = n_in_tile / slice;
      float* owner_red = (S == 1) ? red : cluster.map_shared_rank(red, owner);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int m = 8 * mt + 2 * t + half;
          if (m < p.M) {
            const float4 v = make_float4(acc[0][mt][half], acc[0][mt][2 + half], acc[1][mt][half], acc[1][mt][2 + half]);
            *reinterpret_cast<float4*>(owner_red + (rank * p.M + m) * slice + (n_in_tile - owner * slice)) = v;
          }
        }
    }
  }
  if (warp == 4 && S > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (S > 1) cluster.sync(); else __syncthreads();
  if (warp < 4) {
    const int slice = p.slice;
    const int nbase = n_tile0 + rank * slice;
    const int width = min(slice, 128 - rank * slice);
    const int hw = lane >> 3, hl = lane & 7;  // 8 lanes per output element: lane hl loads source hl (< S <= 8)
    for (int e = warp * 4 + hw; e < p.M * width; e += 16) {
      const int m = e / width, nl = e - m * width;
      float v = (hl < S) ? red[(hl * p.M + m) * slice + nl] : 0.f;
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      const int n = nbase + nl;
      if (hl == 0 && n < p.N) {
        if (p.bias) v += load_as_float(p.bias, p.bias_dtype, n);
        store_from_float(p.y, p.y_dtype, (int64_t)m * p.N + n, v);
      }
    }
  }
}

// optimum-format tensors -> stream records.  One thread per 16-byte piece.
__global__ void build_stream_kernel(const int32_t* __restrict__ qweight, const int32_t* __restrict__ qzeros,
                                    const __half* __restrict__ scales, int N, int K, int g, int G, int NI, int rec_bytes,
                                    uint8_t* __restrict__ out) {
  const int n_tiles = (N + 127) / 128;
  const int64_t pieces_per_rec = rec_bytes / 16;
  const int64_t total = (int64_t)n_tiles * G * pieces_per_rec;
  const int Nw = N / 8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t rec = idx / pieces_per_rec;
    const int piece = (int)(idx - rec * pieces_per_rec);
    const int T = (int)(rec / G), gi = (int)(rec - (int64_t)T * G);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (piece < NI * 128) {  // packed words: [strip][iter][lane]
      const int strip = piece / (NI * 32), rem = piece - strip * NI * 32;
      const int it = rem / 32, lane = rem & 31;
      const int gq = lane >> 2, t = lane & 3;
      const int n = T * 128 + strip * 32 + 4 * gq;
      const int kw = gi * (g / 8) + 4 * it + t;
      if (n < N) v = *reinterpret_cast<const uint4*>(qweight + (int64_t)kw * N + n);
    } else if (piece < NI * 128 + 16) {  // 128 fp16 scales
      const int n = T * 128 + (piece - NI * 128) * 8;
      if (n < N) v = *reinterpret_cast<const uint4*>(scales + (int64_t)gi * N + n);
    } else {  // 128 zero-points as bytes, already +1 and wrapped (modules.py:363, 409-410)
      const int n = T * 128 + (piece - NI * 128 - 16) * 16;
      uint32_t o[4] = {0, 0, 0, 0};
      if (n < N) {
        const uint32_t w0 = (uint32_t)qzeros[(int64_t)gi * Nw + n / 8];
        const uint32_t w1 = (n + 8 < N) ? (uint32_t)qzeros[(int64_t)gi * Nw + n / 8 + 1] : 0u;
        for (int e = 0; e < 16; ++e) {
          const uint32_t nib = ((e < 8 ? w0 : w1) >> (4 * (e & 7))) & 0xfu;
          o[e >> 2] |= ((nib + 1u) & 0xfu) << (8 * (e & 3));
        }
      }
      v = make_uint4(o[0], o[1], o[2], o[3]);
    }
    *reinterpret_cast<uint4*>(out + rec * rec_bytes + (int64_t)piece * 16) = v;
  }
}

static size_t smem_bytes(int M, int rec_bytes, int gmax, int g) {
  return (size_t)kStages * rec_bytes + 2 * kStages * 8 + (size_t)M * kRedPerM * 4 + (size_t)M * (gmax * g + 32) * 2 +
         (((size_t)M * gmax + 3) & ~(size_t)3) * 4 + 128;
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
  const int64_t rec = (int64_t)(g / 32) * 2048 + 256 + 128;
  return ceil_div(N, 128) * (K / g) * rec;
}

extern "C" int b200woq_build_stream_layout(const int32_t* qweight, const int32_t* qzeros, const void* scales16, int64_t N,
                                           int64_t K, int bits, int group_size, void* out, void* stream_) {
  const int g = eff_group(K, group_size);
  WOQ_CHECK_ARG(qweight && qzeros && scales16 && out, "build_stream_layout: null pointer");
  WOQ_CHECK_ARG(stream_shape_ok(N, K, bits, g), "build_stream_layout: unsupported shape (4-bit, g%%32==0, N%%32==0 only)");
  const int NI = g / 32, rec = NI * 2048 + 384;
  const int64_t total = ceil_div(N, 128) * (K / g) * (rec / 16);
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
  WOQ_CHECK_ARG(M <= 16, "linear_forward_stream: M must be <= 16 (use b200woq_linear_forward)");
  Params p = {};
  p.x = x;
  p.x_dtype = x_dtype;
  p.M = (int)M;
  p.K = (int)K;
  p.N = (int)N;
  p.recs = (const uint8_t*)stream_layout;
  p.NI = g / 32;
  p.rec_bytes = p.NI * 2048 + 384;
  p.g = g;
  p.G = (int)(K / g);
  p.bias = bias;
  p.bias_dtype = bias_dtype;
  p.input_scale = input_scale;
  p.y = y;
  p.y_dtype = y_dtype;
  p.pdl = (flags & 2) ? 1 : 0;
  // cluster size: as many CTAs as fit in one wave of 3 CTAs/SM, every CTA keeps >= 2 groups
  const int n_tiles = (int)ceil_div(N, 128);
  const int64_t slots = 3ll * num_sms();
  int S = 1;
  for (int s = 1; s <= 8; ++s) {
    if (s > 1 && p.G / s < 2) break;
    if ((int64_t)n_tiles * s <= slots) S = s;
  }
  while (S < 8 && smem_bytes(p.M, p.rec_bytes, (int)ceil_div(p.G, S), g) > 200 * 1024 && p.G / (S + 1) >= 1) ++S;
  p.S = S;
  p.slice = (int)((ceil_div(128, S) + 3) & ~3);
  p.gmax = (int)ceil_div(p.G, S);
  p.xs_ld = p.gmax * g + 32;
  const size_t smem = smem_bytes(p.M, p.rec_bytes, p.gmax, g);
  if (smem > 220 * 1024) {
    set_error("linear_forward_stream: K slice does not fit in shared memory");
    return B200WOQ_EUNSUPPORTED;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)S, (unsigned)n_tiles);
  cfg.blockDim = dim3(160);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream_;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = (unsigned)S;
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
  if (M <= 8) {
    WOQ_CUDA(cudaFuncSetAttribute(woq_gemm_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WOQ_CUDA(cudaLaunchKernelEx(&cfg, woq_gemm_stream_kernel<1>, p));
  } else {
    WOQ_CUDA(cudaFuncSetAttribute(woq_gemm_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WOQ_CUDA(cudaLaunchKernelEx(&cfg, woq_gemm_stream_kernel<2>, p));
  }
  count_launch(1);
  return 0;
}
