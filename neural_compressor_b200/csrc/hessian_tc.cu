// hessian_tc.cu -- K1 on the 5th-generation tensor cores: H += X^T X with tcgen05.mma, TMA and TMEM.
//
// X is [T, C] token-major (fp16 or bf16).  H[i,j] = sum_t X[t,i] X[t,j]: the contraction index is the SLOW axis of X,
// so both MMA operands are "MN-major" (for a fixed token the 64 channels of a box are contiguous).  A TMA box of
// [BK tokens x 64 channels] with the 128-byte swizzle lands in shared memory exactly as the canonical MN-major
// SWIZZLE_128B UMMA layout:  ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO))  elements,  LBO = BK*128 B (next 64 channels),
// SBO = 1024 B (next 8 tokens).  No transposes, no ldmatrix: the slabs stream HBM/L2 -> smem -> tensor core.
//
//   tile        default: a CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns a 256 x 256 tile, each CTA keeps the
//               128 x 256 fp32 accumulator of its rows in TMEM and stages its 128 A rows + half of the B columns
//               (hessian_syrk_tc2_kernel below).  B200WOQ_SYRK_PAIR=0: one CTA per 128 x 256 tile
//               (hessian_syrk_tc_kernel).  Only tiles touching the upper block triangle are computed; the rest is
//               mirrored in b200woq_hessian_finalize.
//   precision   the tensor core adds into its fp32 accumulator with truncation, so a long contraction drifts
//               (measured 1.3e-5 relative after 8192 tokens).  The token loop is therefore cut into SEGMENTS; each
//               segment accumulates in its own TMEM buffer (2 x 256 columns, double buffered) and is drained with
//               round-to-nearest fp32 adds while the next segment's MMAs run.  Pair kernel: segments of 256 tokens are
//               summed into a REGISTER-resident running tile (8 epilogue warps x 128 columns per thread), and the fp32
//               H tile in global memory is read-modify-written ONCE per launch -- 1-CTA kernel: 2048-token segments,
//               one H round trip per segment.
//   pipeline    mbarrier full/empty ring of [64 tokens x 64 channels] SWIZZLE_128B boxes: 6 stages x 32 KB per CTA in
//               pair mode (4 x 48 KB in 1-CTA mode)
//   warp roles  warp 0 = TMA producer (1 lane), warp 1 = TMEM alloc + MMA issuer (1 lane, even CTA only in pair
//               mode), warps 2-5 = epilogue
//   MMA         tcgen05.mma.kind::f16, K=16, both operands MN-major, fp32 accumulate; M=256/N=256 over the pair
//   epilogue    tcgen05.ld 32x32b.x32 -> registers -> read-modify-write of the fp32 H tile rows owned by this CTA
//
// fp16 x fp16 (11-bit mantissas) products are exact in fp32, so one pass is fp32-grade.
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace b200woq {

namespace tc {

constexpr int TM = 128, TN = 256, BK = 64, STAGES = 4;
constexpr int BOX_BYTES = BK * 128;            // [BK tokens][64 ch] 16-bit
constexpr int A_BYTES = (TM / 64) * BOX_BYTES; // 16 KB
constexpr int B_BYTES = (TN / 64) * BOX_BYTES; // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES; // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NUM_EPI_THREADS = 128;
constexpr int kSuperRowsDefault = 16;  // row tiles per rasterisation super-row
constexpr int kPairDefault = 1;        // CTA-pair (cta_group::2) kernel by default; B200WOQ_SYRK_PAIR=0 selects the 1-CTA kernel
constexpr int TMEM_COLS = 512;      // two 128x256 fp32 accumulators
constexpr int SEG_KB = 32;          // k-blocks (of BK tokens) per accumulation segment = 2048 tokens

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
// MN-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  constexpr uint64_t LBO = (uint64_t)(BOX_BYTES >> 4);  // next 64-channel atom
  constexpr uint64_t SBO = (uint64_t)(1024 >> 4);       // next 8-token group
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (LBO << 16) | (SBO << 32) | (1ull << 46) /*version*/ |
         (2ull << 61) /*SWIZZLE_128B*/;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// grid = one CTA per (128 x 256) tile of the super-row enumeration below; block = 192 threads
__global__ void __launch_bounds__(192, 1)
    hessian_syrk_tc_kernel(const __grid_constant__ CUtensorMap tmap, int64_t Ttok, int64_t C, float* __restrict__ H,
                           uint32_t idesc, int super_rows) {
  // Rasterisation: tiles are enumerated in super-rows of `super_rows` row tiles, column tile by column tile, so the
  // ~148 CTAs in flight cover a 16 x 9 patch of tiles.  They sweep the token axis roughly in step, so what a wave pulls
  // from HBM is (16*128 + 9*256) channels x T instead of (4*128 + all) channels x T with a row-major order: at
  // C = 11008 (X = 360 MB, 3x the L2) that is 2.5x less DRAM traffic.
  int id = blockIdx.x, ti0 = 0, rows, tjmin;
  const int ny = (int)((C + TM - 1) / TM), nx = (int)((C + TN - 1) / TN);
  for (;;) {
    rows = min(super_rows, ny - ti0);
    tjmin = ti0 >> 1;  // tile (ti, tj) touches the upper triangle iff ti <= 2*tj + 1
    const int cnt = rows * (nx - tjmin);
    if (id < cnt) break;
    id -= cnt;
    ti0 += super_rows;
  }
  const int tj = tjmin + id / rows, ti = ti0 + id % rows;
  if (ti > 2 * tj + 1) return;  // below the block diagonal: mirrored later
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t bars = base + STAGES * STAGE_BYTES;            // full[S], empty[S], accum, tmem slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto accum_full = [&](int b) { return bars + 8u * (2 * STAGES + b); };       // MMA -> epilogue
  auto accum_empty = [&](int b) { return bars + 8u * (2 * STAGES + 2 + b); };  // epilogue -> MMA
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)ti * TM, j0 = (int64_t)tj * TN;
  const int nk = (int)((Ttok + BK - 1) / BK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accum_full(b), 1);
      mbar_init(accum_empty(b), NUM_EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        const uint32_t sa = base + s * STAGE_BYTES;
        const int t0 = kb * BK;
#pragma unroll
        for (int b = 0; b < TM / 64; ++b) tma_load_2d(sa + b * BOX_BYTES, &tmap, full_bar(s), (int)(i0 + 64 * b), t0);
#pragma unroll
        for (int b = 0; b < TN / 64; ++b)
          tma_load_2d(sa + A_BYTES + b * BOX_BYTES, &tmap, full_bar(s), (int)(j0 + 64 * b), t0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const int nseg = (nk + SEG_KB - 1) / SEG_KB;
      for (int seg = 0; seg < nseg; ++seg) {
        const int b = seg & 1;
        // wait until the epilogue has drained this buffer (first use of each buffer passes immediately)
        mbar_wait(accum_empty(b), ((uint32_t)(seg >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * TN);
        const int kb1 = min(nk, (seg + 1) * SEG_KB);
        for (int kb = seg * SEG_KB; kb < kb1; ++kb) {
          const int s = kb % STAGES;
          const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * STAGE_BYTES;
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4) {
            const uint64_t ad = make_desc(sa + k4 * 2048);            // 16 tokens = 2 groups of 8 x 128 B rows
            const uint64_t bd = make_desc(sa + A_BYTES + k4 * 2048);
            umma_f16(tmem_d, ad, bd, idesc, (kb != seg * SEG_KB || k4 != 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));  // frees the smem stage when these MMAs retire
        }
        umma_commit(accum_full(b));   // this segment's accumulator is complete
      }
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32); H tile += segment accumulator (round-to-nearest adds)
    const int q = warp & 3;
    const int64_t row = i0 + q * 32 + lane;
    const int nseg = (nk + SEG_KB - 1) / SEG_KB;
    for (int seg = 0; seg < nseg; ++seg) {
      const int b = seg & 1;
      mbar_wait(accum_full(b), (uint32_t)(seg >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cc = 0; cc < TN / 32; ++cc) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * TN + cc * 32), r);
        const int64_t col0 = j0 + cc * 32;
        if (row < C && col0 < C) {
          float* dst = H + row * C + col0;
          if (col0 + 32 <= C && ((C & 3) == 0)) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              float4 h = *reinterpret_cast<float4*>(dst + 4 * v);
              h.x += __uint_as_float(r[4 * v + 0]);
              h.y += __uint_as_float(r[4 * v + 1]);
              h.z += __uint_as_float(r[4 * v + 2]);
              h.w += __uint_as_float(r[4 * v + 3]);
              *reinterpret_cast<float4*>(dst + 4 * v) = h;
            }
          } else {
            for (int v = 0; v < 32; ++v)
              if (col0 + v < C) dst[v] += __uint_as_float(r[v]);
          }
        }
      }
      // release the TMEM buffer to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(accum_empty(b)) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs on one TPC owns a 256 x 256 tile.  CTA r stages ITS 128
// rows of the A panel and HALF of the B panel (128 of the 256 columns); the pair's MMA (M = 256, N = 256, issued by the
// even CTA only) reads A from each CTA's own shared memory and the B halves from both, so every CTA pulls 32 KB per
// 64-token stage through L2 instead of 48 KB for the same flops -- the 1-CTA kernel is bound by exactly that traffic.
// Barriers: full[s] lives in the even CTA and collects the TMA bytes of BOTH CTAs; empty[s] / accum_full[b] are
// arrived on in both CTAs by a multicast tcgen05.commit; accum_empty[b] (even CTA) counts the epilogue threads of both.
constexpr int STAGES2 = 6;
constexpr int A2_BYTES = (128 / 64) * BOX_BYTES;   // 16 KB: this CTA's 128 rows
constexpr int B2_BYTES = (128 / 64) * BOX_BYTES;   // 16 KB: this CTA's half of the 256 columns
constexpr int STAGE2_BYTES = A2_BYTES + B2_BYTES;  // 32 KB
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address of the even CTA of a pair (cute: Sm100MmaPeerBitMask)
constexpr int kSuperRows2Default = 8;
constexpr int EPI2_WARPS = 8;                      // two warps per TMEM lane quarter, 128 columns each
constexpr int EPI2_THREADS = EPI2_WARPS * 32;
constexpr int THREADS2 = 64 + EPI2_THREADS;        // TMA warp + MMA warp + epilogue
constexpr int kSeg2KbDefault = 4;                  // k-blocks of 64 tokens per TMEM accumulation chain (256 tokens)

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_even_cta, int c_inner,
                                                 int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_even_cta), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {  // arrives on `bar` at the same offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// grid = 2 CTAs per 256 x 256 tile of the super-row enumeration; cluster (2,1,1); block = 320 threads
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS2, 1)
    hessian_syrk_tc2_kernel(const __grid_constant__ CUtensorMap tmap, int64_t Ttok, int64_t C, float* __restrict__ H,
                            uint32_t idesc, int super_rows, int SEG2_KB) {
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  int id = blockIdx.x >> 1, ti0 = 0, rows;
  const int nt = (int)((C + 255) / 256);
  for (;;) {
    rows = min(super_rows, nt - ti0);
    const int cnt = rows * (nt - ti0);  // column tiles tj >= ti0 for every row of the super-row
    if (id < cnt) break;
    id -= cnt;
    ti0 += super_rows;
  }
  const int tj = ti0 + id / rows, ti = ti0 + id % rows;
  if (ti > tj) return;  // strictly below the block diagonal (both CTAs of the pair leave together)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES2 * STAGE2_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES2 + s); };
  auto accum_full = [&](int b) { return bars + 8u * (2 * STAGES2 + b); };
  auto accum_empty = [&](int b) { return bars + 8u * (2 * STAGES2 + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES2 + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i0 = (int64_t)ti * 256 + 128 * rank;  // this CTA's accumulator rows
  const int64_t j0 = (int64_t)tj * 256;               // the pair's columns
  const int nk = (int)((Ttok + BK - 1) / BK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accum_full(b), 1);
      mbar_init(accum_empty(b), 2 * EPI2_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is signalled across the pair
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % STAGES2;
        const uint32_t ph = (uint32_t)(kb / STAGES2) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (rank == 0) mbar_expect_tx(full_bar(s), 2 * STAGE2_BYTES);  // the bytes of both CTAs land on this barrier
        const uint32_t sa = base + s * STAGE2_BYTES;
        const uint32_t fb = full_bar(s) & kPeerBitMask;
        const int t0 = kb * BK;
#pragma unroll
        for (int b = 0; b < 2; ++b) tma_load_2d_pair(sa + b * BOX_BYTES, &tmap, fb, (int)(i0 + 64 * b), t0);
#pragma unroll
        for (int b = 0; b < 2; ++b)
          tma_load_2d_pair(sa + A2_BYTES + b * BOX_BYTES, &tmap, fb, (int)(j0 + 128 * rank + 64 * b), t0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const int nseg = (nk + SEG2_KB - 1) / SEG2_KB;
      for (int seg = 0; seg < nseg; ++seg) {
        const int b = seg & 1;
        mbar_wait(accum_empty(b), ((uint32_t)(seg >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * 256);
        const int kb1 = min(nk, (seg + 1) * SEG2_KB);
        for (int kb = seg * SEG2_KB; kb < kb1; ++kb) {
          const int s = kb % STAGES2;
          const uint32_t ph = (uint32_t)(kb / STAGES2) & 1u;
          mbar_wait(full_bar(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * STAGE2_BYTES;
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4) {
            const uint64_t ad = make_desc(sa + k4 * 2048);
            const uint64_t bd = make_desc(sa + A2_BYTES + k4 * 2048);
            umma_f16_pair(tmem_d, ad, bd, idesc, (kb != seg * SEG2_KB || k4 != 0) ? 1u : 0u);
          }
          umma_commit_pair(empty_bar(s));
        }
        umma_commit_pair(accum_full(b));
      }
    }
  } else {
    // epilogue: 8 warps; warp w may touch TMEM lanes [32*(w%4), +32) and owns the column half (w-2)/4 of the tile.
    // Every 512-token segment is added (round to nearest) into a register-resident running row of 128 columns; the
    // fp32 H tile is read-modify-written once, after the last segment.
    const int q = warp & 3, chalf = (warp - 2) >> 2;
    const int64_t row = i0 + q * 32 + lane;
    const int nseg = (nk + SEG2_KB - 1) / SEG2_KB;
    float hacc[128];
#pragma unroll
    for (int v = 0; v < 128; ++v) hacc[v] = 0.f;
    for (int seg = 0; seg < nseg; ++seg) {
      const int b = seg & 1;
      mbar_wait(accum_full(b), (uint32_t)(seg >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256 + chalf * 128 + cc * 32), r);
#pragma unroll
        for (int v = 0; v < 32; ++v) hacc[cc * 32 + v] = __fadd_rn(hacc[cc * 32 + v], __uint_as_float(r[v]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(accum_empty(b) & kPeerBitMask) : "memory");
    }
    if (row < C) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int64_t col0 = j0 + chalf * 128 + cc * 32;
        if (col0 < C) {
          float* dst = H + row * C + col0;
          if (col0 + 32 <= C && ((C & 3) == 0)) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              float4 h = *reinterpret_cast<float4*>(dst + 4 * v);
              h.x += hacc[cc * 32 + 4 * v + 0];
              h.y += hacc[cc * 32 + 4 * v + 1];
              h.z += hacc[cc * 32 + 4 * v + 2];
              h.w += hacc[cc * 32 + 4 * v + 3];
              *reinterpret_cast<float4*>(dst + 4 * v) = h;
            }
          } else {
#pragma unroll
            for (int v = 0; v < 32; ++v)
              if (col0 + v < C) dst[v] += hacc[cc * 32 + v];
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // nobody leaves while the peer may still signal its barriers or read its shared memory
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
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

}  // namespace tc

// returns 0 on success, B200WOQ_EUNSUPPORTED when the shape/alignment needs the legacy kernel
int hessian_accumulate_tcgen05(const void* X, int x_dtype, int64_t T, int64_t C, int64_t ldx, float* Hsum, cudaStream_t st) {
  using namespace tc;
  if (x_dtype != B200WOQ_F16 && x_dtype != B200WOQ_BF16) return B200WOQ_EUNSUPPORTED;
  if ((C % 8) || (ldx % 8) || (((uintptr_t)X) & 15) || T >= (1ll << 31) || C >= (1ll << 31)) return B200WOQ_EUNSUPPORTED;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return B200WOQ_EUNSUPPORTED;
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)T};
  const cuuint64_t strides[1] = {(cuuint64_t)ldx * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)BK};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tmap, x_dtype == B200WOQ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                         const_cast<void*>(X), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return B200WOQ_ECUDA;
  }
  // instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): D=f32, A/B format, both MN-major, N=256, M=128
  const uint32_t fmt = (x_dtype == B200WOQ_F16) ? 0u : 1u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TN >> 3) << 17) |
                         ((uint32_t)(TM >> 4) << 24);
  // per-device attribute: set on every call (cheap), never cached per process
  WOQ_CUDA(cudaFuncSetAttribute(hessian_syrk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  static const int pair_mode = getenv("B200WOQ_SYRK_PAIR") ? atoi(getenv("B200WOQ_SYRK_PAIR")) : kPairDefault;
  if (pair_mode) {
    static const int super_rows2 =
        getenv("B200WOQ_SYRK_SUPER_ROWS") ? std::max(1, atoi(getenv("B200WOQ_SYRK_SUPER_ROWS"))) : kSuperRows2Default;
    // cta_group::2: M = 256 (128 rows per CTA), N = 256
    const uint32_t idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) |
                            ((uint32_t)(256 >> 4) << 24);
    // per-device attribute: set on every call (cheap), never cached per process
    WOQ_CUDA(cudaFuncSetAttribute(hessian_syrk_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
    const int nt = (int)ceil_div(C, 256);
    unsigned tiles = 0;
    for (int ti0 = 0; ti0 < nt; ti0 += super_rows2) tiles += (unsigned)(std::min(super_rows2, nt - ti0) * (nt - ti0));
    // shorter TMEM accumulation chains = less truncation bias in H (the tensor core adds with truncation); the epilogue
    // warps drain each chain into registers while the next one runs
    static const int seg_kb = getenv("B200WOQ_SYRK_SEG_KB") ? std::max(1, atoi(getenv("B200WOQ_SYRK_SEG_KB"))) : kSeg2KbDefault;
    hessian_syrk_tc2_kernel<<<2 * tiles, THREADS2, SMEM2_BYTES, st>>>(tmap, T, C, Hsum, idesc2, super_rows2, seg_kb);
    WOQ_LAUNCH_CHECK();
    return 0;
  }
  static const int super_rows =
      getenv("B200WOQ_SYRK_SUPER_ROWS") ? std::max(1, atoi(getenv("B200WOQ_SYRK_SUPER_ROWS"))) : kSuperRowsDefault;
  const int ny = (int)ceil_div(C, TM), nx = (int)ceil_div(C, TN);
  unsigned total = 0;
  for (int ti0 = 0; ti0 < ny; ti0 += super_rows) total += (unsigned)(std::min(super_rows, ny - ti0) * (nx - (ti0 >> 1)));
  hessian_syrk_tc_kernel<<<total, 192, SMEM_BYTES, st>>>(tmap, T, C, Hsum, idesc, super_rows);
  WOQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200woq
