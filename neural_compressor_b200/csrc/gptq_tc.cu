// gptq_tc.cu -- K3 lazy update on the tensor cores:  W[:, j0:j1] -= Err[:, 0:KK] @ Hinv[r0:r0+KK, j0:j1]   (gptq.py:1304)
//
// The update is a rank-128 GEMM per column block: 8 bytes of W traffic per 256 flops.  On CUDA cores it is issue bound
// (34 TFLOP/s); on the tensor cores it is bound by the W read-modify-write.  fp32-grade accuracy comes from the split
//     x = hi + lo,  hi = x with the low 13 mantissa bits cleared (exactly representable in TF32),  lo = x - hi  (exact)
//     A B ~= A_hi B_hi + A_lo B_hi + A_hi B_lo          (the dropped lo*lo term is < 2^-22 relative)
// with tcgen05.mma.kind::tf32: products of TF32 numbers are exact in fp32, accumulation is fp32 in TMEM.
//
// Operands: ErrT_hi/lo [blocksize, N] (written by the column-loop kernel) and Hinv_hi/lo [C, C] (split once per call).
// For a fixed k both are contiguous along the output dimension, i.e. "MN-major" like the Hessian SYRK: TMA boxes of
// [32 k x 32 floats] with the 128B/32B-atom swizzle land in shared memory as the canonical MN-major layout for 32-bit
// operands (SWIZZLE_128B_BASE32B).
//
//   grid        persistent: one CTA per SM loops over the 128 x 128 output tiles of the launch
//   pipeline    3 stages x 64 KB (E_hi, E_lo, H_hi, H_lo: 4 boxes each), mbarrier full/empty ring continuing across tiles
//   TMEM        two 128 x 128 fp32 accumulators: the epilogue subtracts tile t from W while the MMAs of tile t+1 run
//   warp roles  warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2-5 = epilogue (one W row per thread)
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace b200woq {
namespace lazytc {

constexpr int TM = 128, TN = 128, BK = 32, STAGES = 3;
constexpr int BOX_BYTES = BK * 128;          // [32 k][32 floats]
constexpr int OP_BYTES = 4 * BOX_BYTES;      // 128 rows or columns = 4 boxes = 16 KB
constexpr int STAGE_BYTES = 4 * OP_BYTES;    // E_hi, E_lo, H_hi, H_lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int NUM_EPI_THREADS = 128;
constexpr int TMEM_COLS = 256;

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
// MN-major 32-bit operands have ONE legal shared-memory layout: SWIZZLE_128B_BASE32B (layout type 1; cute:
// Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> o ((T,8,m),(4,k))): 128-byte rows, atoms of 4 k-rows, 32-byte chunks XORed
// with the row index -- what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  LBO = next 32-float atom along M/N
// (one box), SBO = next 4 k-rows (512 B); one K = 8 MMA spans two atoms.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  constexpr uint64_t LBO = (uint64_t)(BOX_BYTES >> 4);
  constexpr uint64_t SBO = (uint64_t)(512 >> 4);
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (LBO << 16) | (SBO << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
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

struct Maps {
  CUtensorMap e_hi, e_lo, h_hi, h_lo;
};

// x -> (hi, lo) with hi = x & ~0x1fff (TF32-exact), lo = x - hi
__global__ void tf32_split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    l.x = v.x - h.x;
    l.y = v.y - h.y;
    l.z = v.z - h.z;
    l.w = v.w - h.w;
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
}

// persistent: grid = min(#SMs, tiles); block = 192
__global__ void __launch_bounds__(192, 1)
    gptq_lazy_update_tc_kernel(const __grid_constant__ Maps maps, float* __restrict__ W, int64_t N, int64_t C, int e0, int r0,
                               int KK, int64_t j0, int64_t j1, int n_row_tiles, int n_col_tiles, uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto accum_full = [&](int b) { return bars + 8u * (2 * STAGES + b); };
  auto accum_empty = [&](int b) { return bars + 8u * (2 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = n_row_tiles * n_col_tiles;
  const int nks = (KK + BK - 1) / BK;  // 32-row k stages per tile

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
      int g = 0;  // running stage counter across tiles
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int rt = tile / n_col_tiles, ct = tile - rt * n_col_tiles;
        const int i0 = rt * TM;
        const int jb = (int)j0 + ct * TN;
        for (int ks = 0; ks < nks; ++ks, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (uint32_t)(g / STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), STAGE_BYTES);
          const uint32_t sa = base + s * STAGE_BYTES;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            tma_load_2d(sa + 0 * OP_BYTES + b * BOX_BYTES, &maps.e_hi, full_bar(s), i0 + 32 * b, e0 + ks * BK);
            tma_load_2d(sa + 1 * OP_BYTES + b * BOX_BYTES, &maps.e_lo, full_bar(s), i0 + 32 * b, e0 + ks * BK);
            tma_load_2d(sa + 2 * OP_BYTES + b * BOX_BYTES, &maps.h_hi, full_bar(s), jb + 32 * b, r0 + ks * BK);
            tma_load_2d(sa + 3 * OP_BYTES + b * BOX_BYTES, &maps.h_lo, full_bar(s), jb + 32 * b, r0 + ks * BK);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int g = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(accum_empty(b), ((uint32_t)(it >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * TN);
        for (int ks = 0; ks < nks; ++ks, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (uint32_t)(g / STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * STAGE_BYTES;
          const int k8n = min(BK / 8, (KK - ks * BK) / 8);  // rows beyond KK were loaded but are not multiplied
          for (int k8 = 0; k8 < k8n; ++k8) {
            const uint64_t ehi = make_desc(sa + 0 * OP_BYTES + k8 * 1024);
            const uint64_t elo = make_desc(sa + 1 * OP_BYTES + k8 * 1024);
            const uint64_t hhi = make_desc(sa + 2 * OP_BYTES + k8 * 1024);
            const uint64_t hlo = make_desc(sa + 3 * OP_BYTES + k8 * 1024);
            umma_tf32(tmem_d, ehi, hhi, idesc, (ks != 0 || k8 != 0) ? 1u : 0u);
            umma_tf32(tmem_d, elo, hhi, idesc, 1u);
            umma_tf32(tmem_d, ehi, hlo, idesc, 1u);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(accum_full(b));
      }
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int rt = tile / n_col_tiles, ct = tile - rt * n_col_tiles;
      const int64_t row = (int64_t)rt * TM + q * 32 + lane;
      const int64_t jb = j0 + (int64_t)ct * TN;
      const int b = it & 1;
      mbar_wait(accum_full(b), (uint32_t)(it >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cc = 0; cc < TN / 32; ++cc) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * TN + cc * 32), r);
        const int64_t col0 = jb + cc * 32;
        if (row < N && col0 < j1) {
          float* dst = W + row * C + col0;
          if (col0 + 32 <= j1 && ((C & 3) == 0) && ((col0 & 3) == 0)) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              float4 w = *reinterpret_cast<float4*>(dst + 4 * v);
              w.x -= __uint_as_float(r[4 * v + 0]);
              w.y -= __uint_as_float(r[4 * v + 1]);
              w.z -= __uint_as_float(r[4 * v + 2]);
              w.w -= __uint_as_float(r[4 * v + 3]);
              *reinterpret_cast<float4*>(dst + 4 * v) = w;
            }
          } else {
            for (int v = 0; v < 32; ++v)
              if (col0 + v < j1) dst[v] -= __uint_as_float(r[v]);
          }
        }
      }
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

static bool encode_f32_2d(EncodeTiledFn enc, CUtensorMap* m, const float* base, int64_t inner, int64_t outer, int64_t ld) {
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)BK};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace lazytc

// ---- host interface used by gptq.cu -------------------------------------------------------------------------------
struct LazyTcPlan {  // same layout as the declaration in gptq.cu
  CUtensorMap maps[4];  // e_hi, e_lo, h_hi, h_lo
  bool ok;
};

bool lazy_tc_shape_ok(int64_t N, int64_t C) {
  return (N % 4 == 0) && (C % 4 == 0) && N < (1ll << 31) && C < (1ll << 31) && lazytc::encode_fn() != nullptr;
}

// Hinv -> (hi, lo) once per fasterquant call
int lazy_tc_split(const float* x, float* hi, float* lo, int64_t n, cudaStream_t st) {
  const int64_t n4 = n / 4;  // callers guarantee n % 4 == 0
  int64_t blocks = std::min<int64_t>(ceil_div(n4, 256), (int64_t)num_sms() * 16);
  lazytc::tf32_split_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, hi, lo, n4);
  WOQ_LAUNCH_CHECK();
  return 0;
}

int lazy_tc_make_plan(LazyTcPlan* plan, const float* e_hi, const float* e_lo, int64_t err_rows, int64_t N, const float* h_hi,
                      const float* h_lo, int64_t C) {
  auto enc = lazytc::encode_fn();
  plan->ok = enc && lazytc::encode_f32_2d(enc, &plan->maps[0], e_hi, N, err_rows, N) &&
             lazytc::encode_f32_2d(enc, &plan->maps[1], e_lo, N, err_rows, N) &&
             lazytc::encode_f32_2d(enc, &plan->maps[2], h_hi, C, C, C) &&
             lazytc::encode_f32_2d(enc, &plan->maps[3], h_lo, C, C, C);
  return plan->ok ? 0 : B200WOQ_EUNSUPPORTED;
}

// W[:, j0:j1] -= ErrT[e0:e0+KK, :]^T @ Hinv[r0:r0+KK, j0:j1];  KK % 8 == 0
int lazy_tc_update(const LazyTcPlan* plan, float* W, int64_t N, int64_t C, int64_t e0, int64_t r0, int KK, int64_t j0,
                   int64_t j1, cudaStream_t st) {
  using namespace lazytc;
  // per-device attribute: set on every call (cheap), never cached per process
  WOQ_CUDA(cudaFuncSetAttribute(gptq_lazy_update_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const int n_row_tiles = (int)ceil_div(N, TM), n_col_tiles = (int)ceil_div(j1 - j0, TN);
  const int64_t tiles = (int64_t)n_row_tiles * n_col_tiles;
  // D = f32, A/B = TF32, both MN-major, N = 128, M = 128 (cute/arch/mma_sm100_desc.hpp: InstrDescriptor)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TN >> 3) << 17) |
                         ((uint32_t)(TM >> 4) << 24);
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, num_sms());
  Maps maps;
  maps.e_hi = plan->maps[0];
  maps.e_lo = plan->maps[1];
  maps.h_hi = plan->maps[2];
  maps.h_lo = plan->maps[3];
  gptq_lazy_update_tc_kernel<<<grid, 192, SMEM_BYTES, st>>>(maps, W, N, C, (int)e0, (int)r0, KK, j0, j1, n_row_tiles,
                                                            n_col_tiles, idesc);
  WOQ_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200woq
