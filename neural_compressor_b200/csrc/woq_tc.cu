// woq_tc.cu -- K6 for batches of 9..128 rows: INT4 group-wise dequant-GEMM on the 5th-generation tensor cores.
//   reference: INCWeightOnlyLinear.forward (modules.py:594-610): recover() = fp16(int8(q - zp) * scale_fp16), F.linear.
//
// Out-channels are MMA-M (128 per CTA), batch rows are MMA-N (16..128), so small batches waste no tensor rows and the
// kernel stays HBM-bound on the packed weights.  Per CTA (448 threads, warp-specialised):
//   warp 0        TMA producer: [8 k-words x 128 out-channels] int32 boxes of the reference's OPTIMUM-format `qweight`
//                 (no derived layout: word (kw, n) already holds 8 consecutive k of out-channel n) -> 8-stage smem ring
//   warps 2-5,    two dequant groups (alternate 64-k blocks), one thread per out-channel row: 8 LDS.32 -> per word
//   warps 6-9     1 SHF + 4 LOP3 ((w & mask) | magic: fp16 1024+q / 64+q) + 4 HSUB2 (exact q - zp) + 4 HMUL2 (one rounding
//                 = the reference's fp16 weight, bit for bit) -> tcgen05.st: the dequantised A tile goes straight into
//                 TENSOR MEMORY (128 lanes x 32 columns per 64-k block), never through shared memory -- writing fp16
//                 weights (4x the packed bytes) to smem and reading them back would cap the kernel near 45 % of HBM
//                 (the LOP3 extraction yields the pairs (k0,k4),(k1,k5),(k2,k6),(k3,k7); four PRMTs per word restore the
//                 natural k order, so the activations need no permuted copy)
//   warp 0 also   TMA-loads the fp16 activation tile [batch x 64 k] (K-major SWIZZLE_128B, rows past M zero-filled) as the
//                 B operand: fully asynchronous, 4 stages in flight -- a synchronous load/convert/store stage per 64-k
//                 block costs ~0.7 us of L2 latency per block and was measured 5x slower than the rest of the pipeline
//   warps 10-13   epilogue: tcgen05.ld of D[128 x batch] fp32 -> (+bias) -> y, coalesced over out-channels
//   warp 1        MMA issuer: tcgen05.mma.kind::f16 with A from TMEM, B from smem, D fp32 in TMEM; tcgen05.commit frees
//                 the A / B stages
// Split-K over gridDim.y when the tile grid is smaller than the chip: fp32 partial tiles go to a workspace and the LAST
// CTA of a tile (atomic counter, self-resetting) sums them in fixed split order -> deterministic.
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace b200woq {
namespace woqtc {

constexpr int TMN = 128;     // out-channels per CTA (MMA M)
constexpr int KB = 64;       // k per pipeline block: 8 packed words per row, one 128-byte swizzle row of fp16 activations
constexpr int WST = 24;      // packed-weight smem stages (4 KB each): ~2.5 us of HBM latency x the per-SM share of the
                             // bandwidth = ~100 KB that must be in flight per SM (8 stages measured 3x too few)
constexpr int DG = 4;        // dequant groups of 4 warps (block i belongs to group i % DG)
constexpr int AST = 8;       // TMEM A stages (32 columns each): two per dequant group
constexpr int BST = 32;      // max activation smem stages: as many [NT x 64 k] tiles as fit in 64 KB (32 at NT = 16, 4 at NT = 128);
                             // the tiles are tiny, so only a deep ring keeps enough bytes in flight to cover the L2 latency
constexpr int W_STAGE_BYTES = 8 * TMN * 4;
constexpr int DEQ_WARPS = 4 * DG;
constexpr int EPI_WARP0 = 2 + DEQ_WARPS;   // first epilogue warp (18): 18 % 4 == 2, like warp 2, so quarters line up
constexpr int THREADS = (EPI_WARP0 + 4) * 32;

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// K-major SWIZZLE_128B smem descriptor (rows 128 B apart, 8-row atoms 1024 B apart)
__device__ __forceinline__ uint64_t make_desc_k(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
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
__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t magic) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(magic));  // (a & mask) | magic
  return d;
}
__device__ __forceinline__ uint32_t h2_sub_mul(uint32_t v, uint32_t zc, uint32_t s2) {
  __half2 h = __hsub2(*reinterpret_cast<__half2*>(&v), *reinterpret_cast<__half2*>(&zc));   // exact: small integers
  h = __hmul2(h, *reinterpret_cast<__half2*>(&s2));                                         // the one rounding of recover()
  return *reinterpret_cast<uint32_t*>(&h);
}

struct Params {
  const void* x;
  int x_dtype;
  int M, K, N;
  int NT;              // batch tile = MMA N (multiple of 16, >= M)
  int g, G;
  int kb_per_split;    // 64-k blocks per split
  const int32_t* qzeros;
  const __half* scales;
  const void* bias;
  int bias_dtype;
  const float* input_scale;
  void* y;
  int y_dtype;
  float* part_ws;      // [tiles][splits][NT][128] fp32 partial tiles (split-K only)
  int* counters;       // [tiles], zero on entry and on exit
  int pdl;
  int debug;           // B200WOQ_TC_DEBUG (measurement only): 1 = issue one MMA per block instead of four, 2 = none
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

// grid (N / 128, splits); block 448
__global__ void __launch_bounds__(THREADS, 1)
    woq_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const Params p,
                       uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int NT = p.NT;
  const uint32_t b_stage_bytes = (uint32_t)NT * 128;
  const uint32_t w_ring = base;                                   // WST x 4 KB
  const uint32_t b_ring = base + WST * W_STAGE_BYTES;             // BST x NT x 128 B (1024-aligned: NT % 16 == 0 -> multiple of 2 KB)
  const int bst = min(BST, (int)(65536u / b_stage_bytes));          // 64 KB of activation stages
  const uint32_t bars = b_ring + 4 * (uint32_t)(128 * 128);
  auto full_w = [&](int s) { return bars + 8u * s; };
  auto empty_w = [&](int s) { return bars + 8u * (WST + s); };
  auto a_full = [&](int s) { return bars + 8u * (2 * WST + s); };
  auto a_empty = [&](int s) { return bars + 8u * (2 * WST + AST + s); };
  auto b_full = [&](int s) { return bars + 8u * (2 * WST + 2 * AST + s); };
  auto b_empty = [&](int s) { return bars + 8u * (2 * WST + 2 * AST + BST + s); };
  const uint32_t d_full = bars + 8u * (2 * WST + 2 * AST + 2 * BST);
  const uint32_t tmem_slot = d_full + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  int* flag_ptr = reinterpret_cast<int*>(smem_raw + (tmem_slot + 8 - smem_u32(smem_raw)));
  // scales / zero-points of this CTA's 128 out-channels for ITS groups, staged once (they are constants): the dequant
  // warps then never wait on a global load inside the pipeline
  __half* sc_tab = reinterpret_cast<__half*>(smem_raw + (bars + 2048 - smem_u32(smem_raw)));   // [groups][128] fp16

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TMN;
  const int nkb_total = p.K / KB;
  const int kb0 = blockIdx.y * p.kb_per_split;
  const int nkb = min(p.kb_per_split, nkb_total - kb0);
  const int g_first = (kb0 * KB) / p.g, g_last = ((kb0 + nkb) * KB - 1) / p.g, n_groups = g_last - g_first + 1;
  uint8_t* z_tab = reinterpret_cast<uint8_t*>(sc_tab + (size_t)n_groups * TMN);                // [groups][128] u8 (already +1)
  // NACC independent accumulators (k-steps round-robin): successive MMAs into ONE 128 x NT tile are a dependent chain of
  // ~60 cycles each (measured), four chains hide it; the epilogue adds them up
  const uint32_t d_stride = NT <= 32 ? 32 : NT <= 64 ? 64 : 128;
  const int nacc = NT <= 64 ? 4 : 2;
  const uint32_t d_cols = d_stride * nacc;
  const uint32_t tmem_cols = 512;   // D (<= 128 columns) + 8 A stages of 32 columns

  if (threadIdx.x == 0) {
    for (int s = 0; s < WST; ++s) {
      mbar_init(full_w(s), 1);
      mbar_init(empty_w(s), 4);      // one elected arrival per dequant warp of the consuming group
    }
    for (int s = 0; s < AST; ++s) {
      mbar_init(a_full(s), 4);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < BST; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    mbar_init(d_full, NT <= 64 ? 4 : 2);   // every MMA issuer commits once at the end
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    // one vectorised round trip: per group 16 x 16-byte pieces of scales and 16 zero-point words (8 nibbles each)
    const int Nw = p.N >> 3;
    for (int idx = threadIdx.x; idx < n_groups * 16; idx += THREADS) {
      const int gl = idx >> 4, pc = idx & 15;
      const uint4 sv = *reinterpret_cast<const uint4*>(p.scales + (int64_t)(g_first + gl) * p.N + n0 + pc * 8);
      *reinterpret_cast<uint4*>(sc_tab + gl * TMN + pc * 8) = sv;
      const uint32_t zw = (uint32_t)p.qzeros[(int64_t)(g_first + gl) * Nw + (n0 >> 3) + pc];
      uint32_t lo = 0, hi = 0;   // stored minus one (modules.py:363, 409-410): +1 and wrap to 4 bits
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        lo |= ((((zw >> (4 * e)) & 0xfu) + 1u) & 0xfu) << (8 * e);
        hi |= ((((zw >> (4 * (e + 4))) & 0xfu) + 1u) & 0xfu) << (8 * e);
      }
      *reinterpret_cast<uint2*>(z_tab + gl * TMN + pc * 8) = make_uint2(lo, hi);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_a0 = tmem_base + d_cols;   // A stages behind the accumulator columns

  // One issuing thread spends ~600 cycles per 64-k block (two mbarrier waits at ~90 cycles each even when complete, four
  // tiny MMAs at ~66 cycles of issue each, two commits): measured as THE serial bottleneck of the pipeline.  Two issuers
  // take alternate blocks and own disjoint accumulators, so no ordering between them is needed.  `nacc` issuers (4 for
  // batch tiles <= 64, 2 for 128): issuer r = block index mod nacc, accumulator r.
  auto mma_role = [&](int r) {
    for (int i = r; i < nkb; i += nacc) {
      const int ta = i % AST, tb = i % bst;
      mbar_wait(a_full(ta), ((uint32_t)(i / AST)) & 1u);
      mbar_wait(b_full(tb), ((uint32_t)(i / bst)) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sb = b_ring + tb * b_stage_bytes;
#pragma unroll
      for (int k4 = 0; k4 < KB / 16; ++k4)
        umma_f16_ts(tmem_base + (uint32_t)r * d_stride, tmem_a0 + (uint32_t)(ta * 32 + k4 * 8), make_desc_k(sb + k4 * 32), idesc,
                    (i == r && k4 == 0) ? 0u : 1u);
      umma_commit(a_empty(ta));
      umma_commit(b_empty(tb));
    }
    umma_commit(d_full);
  };

  if (warp == 0) {
    // ---------------- TMA producer of the packed weights (constants: no dependency on the previous kernel)
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % WST;
        mbar_wait(empty_w(s), (((uint32_t)(i / WST)) & 1u) ^ 1u);
        mbar_expect_tx(full_w(s), W_STAGE_BYTES);
        tma_load_2d(w_ring + s * W_STAGE_BYTES, &map_w, full_w(s), n0, (kb0 + i) * 8);
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer 0 (even blocks); issuer 1 (odd blocks) is lane 0 of the second epilogue warp
    if (lane == 0) mma_role(0);
  } else if (warp < EPI_WARP0) {
    // ---------------- dequant: group dg handles blocks i = dg, dg + 2, ...; thread = out-channel row
    const int dg = (warp - 2) >> 2;            // 0 .. DG-1
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int nl = q * 32 + lane;              // row inside the tile
    const int n = n0 + nl;
    auto fetch_sz = [&](int i, uint32_t& s2, uint32_t& zlo, uint32_t& zhi) {
      const int gl = ((kb0 + i) * KB) / p.g - g_first;
      const __half s = sc_tab[gl * TMN + nl];
      const uint32_t z = z_tab[gl * TMN + nl];
      const __half2 sh = __half2half2(s);
      s2 = *reinterpret_cast<const uint32_t*>(&sh);
      const __half2 zl = __half2half2(__ushort_as_half((unsigned short)(0x6400u + z)));   // 1024 + z
      const __half2 zh = __half2half2(__ushort_as_half((unsigned short)(0x5400u + (z << 4))));   // 64 + z
      zlo = *reinterpret_cast<const uint32_t*>(&zl);
      zhi = *reinterpret_cast<const uint32_t*>(&zh);
    };
    uint32_t s2 = 0, zlo = 0, zhi = 0;
    for (int i = dg; i < nkb; i += DG) {
      fetch_sz(i, s2, zlo, zhi);
      const int s = i % WST, ta = i % AST;
      if (p.debug != 6) mbar_wait(full_w(s), ((uint32_t)(i / WST)) & 1u);
      const uint32_t* wsm = reinterpret_cast<const uint32_t*>(base_ptr + s * W_STAGE_BYTES) + nl;
      uint32_t wd[8];
      if (p.debug >= 5) {
#pragma unroll
        for (int kw = 0; kw < 8; ++kw) wd[kw] = 0;
      } else
#pragma unroll
      for (int kw = 0; kw < 8; ++kw) wd[kw] = wsm[kw * TMN];
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_w(s));  // the warp's words are in registers (128 same-address arrivals serialise)
      uint32_t a[32];
      if (p.debug >= 4) {   // measurement only: no dequant arithmetic
#pragma unroll
        for (int kw = 0; kw < 8; ++kw) a[4 * kw] = a[4 * kw + 1] = a[4 * kw + 2] = a[4 * kw + 3] = wd[kw];
      } else
#pragma unroll
      for (int kw = 0; kw < 8; ++kw) {
        const uint32_t w = wd[kw], w8 = w >> 8;
        const uint32_t p04 = h2_sub_mul(lop3_and_or(w, 0x000f000fu, 0x64006400u), zlo, s2);    // (k0, k4)
        const uint32_t p15 = h2_sub_mul(lop3_and_or(w, 0x00f000f0u, 0x54005400u), zhi, s2);    // (k1, k5)
        const uint32_t p26 = h2_sub_mul(lop3_and_or(w8, 0x000f000fu, 0x64006400u), zlo, s2);   // (k2, k6)
        const uint32_t p37 = h2_sub_mul(lop3_and_or(w8, 0x00f000f0u, 0x54005400u), zhi, s2);   // (k3, k7)
        a[4 * kw + 0] = __byte_perm(p04, p15, 0x5410);   // (k0, k1): natural k order, TMEM column = 2 consecutive k
        a[4 * kw + 1] = __byte_perm(p26, p37, 0x5410);   // (k2, k3)
        a[4 * kw + 2] = __byte_perm(p04, p15, 0x7632);   // (k4, k5)
        a[4 * kw + 3] = __byte_perm(p26, p37, 0x7632);   // (k6, k7)
      }
      mbar_wait(a_empty(ta), (((uint32_t)(i / AST)) & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (p.debug != 3 && p.debug < 5) tmem_st32(tmem_a0 + ((uint32_t)(q * 32) << 16) + (uint32_t)(ta * 32), a);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(ta));
    }
  } else {
    // ---------------- epilogue warps; lane 0 of the first one also produces the activation tiles, running as far ahead
    // as its own ring allows (independent of the weight ring: coupling the two in one thread serialised every block's MMA
    // behind an activation load that was issued one block earlier)
    const int st = threadIdx.x - EPI_WARP0 * 32;      // 0..127
    const int q = warp & 3;
    if (st == 0) {
      if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int i = 0; i < nkb; ++i) {
        const int tb = i % bst;
        mbar_wait(b_empty(tb), (((uint32_t)(i / bst)) & 1u) ^ 1u);
        mbar_expect_tx(b_full(tb), b_stage_bytes);
        tma_load_2d(b_ring + tb * b_stage_bytes, &map_x, b_full(tb), (kb0 + i) * KB, 0);
      }
    }
    if ((st & 31) == 0 && st > 0 && (st >> 5) < nacc) mma_role(st >> 5);   // MMA issuers 1..nacc-1 (lane 0 of epilogue warps 1..3)
    __syncwarp();
    // ---- epilogue: lane quarter q, thread = out-channel n0 + 32 q + lane.  The accumulator is ready only at the very
    // end: poll with a back-off instead of spinning 128 threads through the whole main loop
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(d_full), "r"(0u)
            : "memory");
        if (!done) __nanosleep(256);
      }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int nl = q * 32 + lane, n = n0 + nl;
    const int splits = gridDim.y;
    const float bias = (p.bias && n < p.N) ? load_as_float(p.bias, p.bias_dtype, n) : 0.f;
    auto load_acc = [&](int c, float (&acc)[16]) {   // sum of the independent accumulators, fixed order
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
#pragma unroll
      for (int v = 0; v < 16; ++v) acc[v] = __uint_as_float(r[v]);
      const int nacc_used = min(nacc, nkb);
      for (int a = 1; a < nacc_used; ++a) {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * d_stride + (uint32_t)c, r);
#pragma unroll
        for (int v = 0; v < 16; ++v) acc[v] += __uint_as_float(r[v]);
      }
    };
    if (splits == 1) {
      for (int c = 0; c < NT; c += 16) {
        float acc[16];
        load_acc(c, acc);
#pragma unroll
        for (int v = 0; v < 16; ++v)
          if (c + v < p.M) store_from_float(p.y, p.y_dtype, (int64_t)(c + v) * p.N + n, acc[v] + bias);
      }
    } else {
      float* mine = p.part_ws + ((size_t)(blockIdx.x * splits + blockIdx.y) * NT) * TMN + nl;
      for (int c = 0; c < NT; c += 16) {
        float acc[16];
        load_acc(c, acc);
#pragma unroll
        for (int v = 0; v < 16; ++v)
          if (c + v < p.M) mine[(size_t)(c + v) * TMN] = acc[v];
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (st == 0) *flag_ptr = atomicAdd(p.counters + blockIdx.x, 1);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*flag_ptr == splits - 1) {   // last CTA of this tile: fixed-order sum over the splits
        __threadfence();
        const float* t0 = p.part_ws + ((size_t)(blockIdx.x * splits) * NT) * TMN + nl;
        for (int mb = 0; mb < p.M; mb += 4) {       // 4 rows x all splits of loads in flight, summed in split order
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          for (int s = 0; s < splits; ++s) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (mb + u < p.M) ? __ldcg(t0 + ((size_t)s * NT + mb + u) * TMN) : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] += v[u];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (mb + u < p.M) store_from_float(p.y, p.y_dtype, (int64_t)(mb + u) * p.N + n, acc[u] + bias);
        }
        if (st == 0) p.counters[blockIdx.x] = 0;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
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

}  // namespace woqtc

// shape gate for the tensor-core path
// (activations must be fp16, 16-byte aligned and contiguous; an input_scale is applied by the caller beforehand)
bool woq_tc_shape_ok(int64_t M, int64_t N, int64_t K, int bits, int g, const int32_t* g_idx) {
  return bits == 4 && !g_idx && M >= 1 && M <= 128 && (N % 128) == 0 && (K % 64) == 0 && g > 0 && (g % 64) == 0 && (K % g) == 0;
}

static void woq_tc_plan(int64_t M, int64_t N, int64_t K, int* NT, int* splits, int* per) {
  *NT = (int)(ceil_div(M, 16) * 16);
  const int64_t tiles = N / 128;
  const int nkb = (int)(K / 64);
  const int sms = num_sms();
  int s = 1;
  if (tiles < sms) {
    s = (int)std::max<int64_t>(1, sms / tiles);
    s = std::min(s, std::max(1, nkb / 8));
    s = std::min(s, 16);
  }
  int p = (int)ceil_div(nkb, s);
  s = (int)ceil_div(nkb, p);
  *splits = s;
  *per = p;
}

int64_t woq_tc_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  int NT, s, per;
  woq_tc_plan(M, N, K, &NT, &s, &per);
  const int64_t tiles = N / 128;
  return 256 + tiles * 4 + (s > 1 ? tiles * s * NT * 128 * 4 : 0) + 256;
}

// workspace: [counters: tiles ints (zero on entry, left zero)][pad to 256][partials]
int woq_tc_forward(const void* x, int x_dtype, int64_t M, int64_t K, int64_t N, const int32_t* qweight, const int32_t* qzeros,
                   const __half* scales, const void* bias, int bias_dtype, const float* input_scale, void* y, int y_dtype,
                   int g, void* workspace, int64_t workspace_bytes, int pdl, cudaStream_t st) {
  using namespace woqtc;
  if (workspace_bytes < woq_tc_workspace_bytes(M, N, K) || !workspace) {
    set_error("woq_tc_forward: workspace too small");
    return B200WOQ_EWORKSPACE;
  }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return B200WOQ_EUNSUPPORTED;
  CUtensorMap map_w;
  const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)(K / 8)};
  const cuuint64_t strides[1] = {(cuuint64_t)N * 4};
  const cuuint32_t box[2] = {128, 8};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, const_cast<int32_t*>(qweight), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    set_error("woq_tc_forward: cuTensorMapEncodeTiled failed");
    return B200WOQ_ECUDA;
  }
  CUtensorMap map_x;
  {
    const cuuint64_t xd[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t xs[1] = {(cuuint64_t)K * 2};
    int NTb, sb, pb;
    woq_tc_plan(M, N, K, &NTb, &sb, &pb);
    const cuuint32_t xb[2] = {64, (cuuint32_t)NTb};
    if (enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x), xd, xs, xb, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_error("woq_tc_forward: cuTensorMapEncodeTiled (activations) failed");
      return B200WOQ_ECUDA;
    }
  }
  Params p = {};
  p.x = x; p.x_dtype = x_dtype; p.M = (int)M; p.K = (int)K; p.N = (int)N; p.g = g; p.G = (int)(K / g);
  p.qzeros = qzeros; p.scales = scales; p.bias = bias; p.bias_dtype = bias_dtype; p.input_scale = input_scale;
  p.y = y; p.y_dtype = y_dtype; p.pdl = pdl;
  static const int dbg = getenv("B200WOQ_TC_DEBUG") ? atoi(getenv("B200WOQ_TC_DEBUG")) : 0;
  p.debug = dbg;
  int splits;
  woq_tc_plan(M, N, K, &p.NT, &splits, &p.kb_per_split);
  p.counters = (int*)workspace;
  const int64_t tiles = N / 128;
  p.part_ws = (float*)((uint8_t*)workspace + ((tiles * 4 + 255) / 256) * 256);
  // instruction descriptor: D = F32 (1), A = B = F16 (0), both K-major, N = NT, M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const size_t groups_max = (size_t)ceil_div((int64_t)p.kb_per_split * KB, g) + 1;
  const size_t smem = (size_t)WST * W_STAGE_BYTES + (size_t)4 * 128 * 128 + 2048 + 1024 + groups_max * TMN * 3 + 256;
  WOQ_CUDA(cudaFuncSetAttribute(woq_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)tiles, (unsigned)splits);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  WOQ_CUDA(cudaLaunchKernelEx(&cfg, woq_gemm_tc_kernel, map_w, map_x, p, idesc));
  count_launch(1);
  return 0;
}

}  // namespace b200woq
