// hessian.cu -- K1: GPTQ Hessian accumulation  H = (2/n) * sum_j X_j^T X_j   (gptq.py:1111-1141)
//
// The reference updates a running mean per sample (H *= n/(n+b); H += (sqrt(2/n) X)^T (sqrt(2/n) X));
// the closed form is accumulated here as raw fp32 sums of X^T X and scaled once in
// b200woq_hessian_finalize.  X is [T, C] row-major (tokens x channels), so both MMA operands are the
// SAME matrix read "transposed" (token index is the contraction dim):  H[i,j] = sum_t X[t,i] X[t,j].
//
//   * fp16 / bf16 activations (the Llama-2-7B fp16 calibration of BASELINE configs[1]): products of two
//     11-bit (8-bit) mantissas are exact in fp32, so ONE tensor-core pass with fp32 accumulation is
//     fp32-grade (SURVEY §7.2 needs fp32-accurate H to keep codes bit-stable).  128x128 output tiles,
//     only tiles on/above the diagonal (SYRK), 4-stage cp.async pipeline of [32 tokens x 128 ch] slabs,
//     ldmatrix.trans feeds mma.sync.m16n8k16 directly from the token-major slabs.
//   * fp32 activations (tiny unit-test models): exact fp32 FFMA tiles (generic kernel below).
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace b200woq {

constexpr bool kHessianDefaultTc = true;  // tcgen05 kernel is the default (parity-green on B200); mma.sync kept as fallback
constexpr int HT = 128;      // output tile edge
constexpr int HBK = 32;      // tokens per pipeline stage
constexpr int HLD = HT + 8;  // padded smem row (halves): 272 B rows -> conflict-free ldmatrix
constexpr int HSTAGES = 4;

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
template <typename T>
__device__ __forceinline__ void mma_f32_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_f32_16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_f32_16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                                             uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 16-bit inputs, tensor cores.  grid = (nt, nt), CTAs below the diagonal exit.
template <typename T>
__global__ void __launch_bounds__(256) hessian_syrk16_kernel(const T* __restrict__ X, int64_t Ttok, int64_t C,
                                                            int64_t ldx, float* __restrict__ H) {
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (ti > tj) return;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);  // [stage][2][HBK][HLD]
  const bool diag = (ti == tj);
  const int64_t i0 = (int64_t)ti * HT, j0 = (int64_t)tj * HT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 2, wn = warp & 3;  // warp tile: 64 (i) x 32 (j)
  const int nk = (int)((Ttok + HBK - 1) / HBK);

  auto tileA = [&](int s) { return sm + (size_t)(s * 2 + 0) * HBK * HLD; };
  auto tileB = [&](int s) { return diag ? tileA(s) : sm + (size_t)(s * 2 + 1) * HBK * HLD; };

  auto load_stage = [&](int s, int kb) {
    const int64_t t0 = (int64_t)kb * HBK;
    // 32 rows x 16 chunks of 16 B per tile; 256 threads -> 2 chunks per tile
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int c = threadIdx.x + rep * 256;
      const int row = c >> 4, ch = (c & 15) * 8;
      const int64_t t = t0 + row;
      {
        const int64_t col = i0 + ch;
        const bool ok = (t < Ttok) && (col < C);
        cp_async_16(tileA(s) + row * HLD + ch, ok ? (const void*)(X + t * ldx + col) : (const void*)X, ok ? 16 : 0);
      }
      if (!diag) {
        const int64_t col = j0 + ch;
        const bool ok = (t < Ttok) && (col < C);
        cp_async_16(tileB(s) + row * HLD + ch, ok ? (const void*)(X + t * ldx + col) : (const void*)X, ok ? 16 : 0);
      }
    }
  };

  float acc[4][4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

#pragma unroll
  for (int s = 0; s < HSTAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  for (int kb = 0; kb < nk; ++kb) {
    cp_async_wait<HSTAGES - 2>();
    __syncthreads();
    {
      const int nxt = kb + HSTAGES - 1;
      if (nxt < nk) load_stage(nxt % HSTAGES, nxt);
      cp_async_commit();
    }
    const T* A = tileA(kb % HSTAGES);
    const T* B = tileB(kb % HSTAGES);
    const int r = lane & 7, mat = lane >> 3;
#pragma unroll
    for (int ks = 0; ks < HBK / 16; ++ks) {
      const int k0 = ks * 16;
      uint32_t af[4][4], bf[2][4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
        ldmatrix_x4_trans(af[mi], A + (k0 + (mat >> 1) * 8 + r) * HLD + wm * 64 + mi * 16 + (mat & 1) * 8);
#pragma unroll
      for (int nj = 0; nj < 2; ++nj)
        ldmatrix_x4_trans(bf[nj], B + (k0 + (mat & 1) * 8 + r) * HLD + wn * 32 + nj * 16 + (mat >> 1) * 8);
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int n8 = 0; n8 < 4; ++n8)
          mma_f32_16816<T>(acc[mi][n8], af[mi], bf[n8 >> 1][(n8 & 1) * 2], bf[n8 >> 1][(n8 & 1) * 2 + 1]);
    }
  }
  cp_async_wait<0>();
  // epilogue: H[i,j] += acc  (tile is owned by this CTA for the whole launch)
  const int gq = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int n8 = 0; n8 < 4; ++n8)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t i = i0 + wm * 64 + mi * 16 + gq + h * 8;
        const int64_t j = j0 + wn * 32 + n8 * 8 + 2 * t4;
        if (i < C && j < C) {
          float* dst = H + i * C + j;
          if (j + 1 < C) {
            float2 v = *reinterpret_cast<float2*>(dst);
            v.x += acc[mi][n8][2 * h];
            v.y += acc[mi][n8][2 * h + 1];
            *reinterpret_cast<float2*>(dst) = v;
          } else {
            dst[0] += acc[mi][n8][2 * h];
          }
        }
      }
}

// generic exact-fp32 path (fp32 activations, or unaligned 16-bit): 64x64 tile, 4x4 per thread
template <typename T>
__global__ void __launch_bounds__(256) hessian_syrk_simt_kernel(const T* __restrict__ X, int64_t Ttok, int64_t C,
                                                               int64_t ldx, float* __restrict__ H) {
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (ti > tj) return;
  __shared__ float As[16][64 + 4], Bs[16][64 + 4];
  const int64_t i0 = (int64_t)ti * 64, j0 = (int64_t)tj * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int64_t t0 = 0; t0 < Ttok; t0 += 16) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      const int r = e >> 6, c = e & 63;
      const int64_t t = t0 + r;
      As[r][c] = (t < Ttok && i0 + c < C) ? ElemTraits<T>::load(X + t * ldx + i0 + c) : 0.f;
      Bs[r][c] = (t < Ttok && j0 + c < C) ? ElemTraits<T>::load(X + t * ldx + j0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = As[k][ty * 4 + u];
        b[u] = Bs[k][tx * 4 + u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int64_t i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i < C && j < C) H[i * C + j] += acc[u][v];
    }
}

// finalize: scale by 2/n, mirror the upper 128-tiles into the lower triangle
__global__ void hessian_scale_mirror_kernel(float* __restrict__ H, int64_t C, float factor, int tile) {
  const int64_t total = C * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / C, j = idx % C;
    if (i / tile <= j / tile) H[idx] *= factor;  // owned (computed) region
  }
}
__global__ void hessian_mirror_kernel(float* __restrict__ H, int64_t C, int tile) {
  const int64_t total = C * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / C, j = idx % C;
    if (i / tile > j / tile) H[idx] = H[j * C + i];
  }
}
// dead columns + damping (gptq.py:1189-1191, 1221-1227); single block, deterministic reduction order
__global__ void __launch_bounds__(1024) hessian_dead_damp_kernel(float* __restrict__ H, int64_t C, float percdamp,
                                                                uint8_t* __restrict__ dead, float* __restrict__ scratch) {
  __shared__ float part[1024];
  float s = 0.f;
  for (int64_t c = threadIdx.x; c < C; c += blockDim.x) {
    float d = H[c * C + c];
    const bool is_dead = (d == 0.f);
    if (is_dead) {
      d = 1.f;
      H[c * C + c] = 1.f;
    }
    if (dead) dead[c] = is_dead ? 1 : 0;
    s += d;
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  const float damp = percdamp * (part[0] / (float)C);
  if (threadIdx.x == 0 && scratch) {
    scratch[0] = damp;
    scratch[1] = part[0] / (float)C;
  }
  for (int64_t c = threadIdx.x; c < C; c += blockDim.x) H[c * C + c] += damp;
}

}  // namespace b200woq

namespace b200woq {
int hessian_accumulate_tcgen05(const void* X, int x_dtype, int64_t T, int64_t C, int64_t ldx, float* Hsum, cudaStream_t st);
}

using namespace b200woq;

// which tile edge owns the "computed" upper region for a given (dtype, alignment)
static int hessian_tile_edge(int x_dtype, int64_t C, int64_t ldx, const void* X) {
  const bool fast = (x_dtype != B200WOQ_F32) && (C % 8 == 0) && (ldx % 8 == 0) && (((uintptr_t)X & 15) == 0);
  return fast ? HT : 64;
}

extern "C" int b200woq_hessian_accumulate(const void* X, int x_dtype, int64_t T, int64_t C, int64_t ldx, float* Hsum,
                                          void* stream) {
  WOQ_CHECK_ARG(X && Hsum && T > 0 && C > 0 && ldx >= C, "hessian_accumulate: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  // NOTE: the upper region is defined on 128-tiles for BOTH kernels so finalize can mirror uniformly:
  // the SIMT kernel uses 64-tiles but (i/64 <= j/64) is a superset of the 128-tile upper region only
  // when restricted properly, so it always computes every 64-tile with (i/128 <= j/128).
  const int edge = hessian_tile_edge(x_dtype, C, ldx, X);
  // B200WOQ_HESSIAN_IMPL = "tc" (tcgen05 + TMA + TMEM, hessian_tc.cu) | "mma" (mma.sync + ldmatrix, below)
  const char* impl = getenv("B200WOQ_HESSIAN_IMPL");
  const bool want_tc = impl ? (strcmp(impl, "tc") == 0) : kHessianDefaultTc;
  if (want_tc && edge == HT) {
    const int rc = hessian_accumulate_tcgen05(X, x_dtype, T, C, ldx, Hsum, st);
    if (rc != B200WOQ_EUNSUPPORTED) return rc;
  }
  if (edge == HT) {
    const unsigned nt = (unsigned)ceil_div(C, HT);
    const size_t smem = (size_t)HSTAGES * 2 * HBK * HLD * 2;
    if (x_dtype == B200WOQ_F16) {
      WOQ_CUDA(cudaFuncSetAttribute(hessian_syrk16_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      hessian_syrk16_kernel<__half><<<dim3(nt, nt), 256, smem, st>>>((const __half*)X, T, C, ldx, Hsum);
    } else {
      WOQ_CUDA(cudaFuncSetAttribute(hessian_syrk16_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      hessian_syrk16_kernel<__nv_bfloat16><<<dim3(nt, nt), 256, smem, st>>>((const __nv_bfloat16*)X, T, C, ldx, Hsum);
    }
  } else {
    const unsigned nt = (unsigned)ceil_div(C, 64);
    dim3 grid(nt, nt);
    if (x_dtype == B200WOQ_F32)
      hessian_syrk_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)X, T, C, ldx, Hsum);
    else if (x_dtype == B200WOQ_F16)
      hessian_syrk_simt_kernel<__half><<<grid, 256, 0, st>>>((const __half*)X, T, C, ldx, Hsum);
    else if (x_dtype == B200WOQ_BF16)
      hessian_syrk_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)X, T, C, ldx, Hsum);
    else {
      set_error("hessian_accumulate: bad dtype %d", x_dtype);
      return B200WOQ_EINVAL;
    }
  }
  WOQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200woq_hessian_finalize(float* H, int64_t C, double nsamples, float percdamp, uint8_t* dead_mask,
                                        float* scratch, void* stream) {
  WOQ_CHECK_ARG(H && C > 0 && nsamples > 0, "hessian_finalize: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const float factor = (float)(2.0 / nsamples);
  int64_t b = ceil_div(C * C, 256);
  const int64_t cap = (int64_t)num_sms() * 32;
  const int blocks = (int)(b > cap ? cap : b);
  // every accumulate kernel fully computes all 64-tiles (i/64 <= j/64), a superset of that triangle at
  // 128 granularity is NOT guaranteed, so scale + mirror at the finest granularity both kernels share: 64.
  hessian_scale_mirror_kernel<<<blocks, 256, 0, st>>>(H, C, factor, 64);
  hessian_mirror_kernel<<<blocks, 256, 0, st>>>(H, C, 64);
  hessian_dead_damp_kernel<<<1, 1024, 0, st>>>(H, C, percdamp, dead_mask, scratch);
  count_launch(2);
  WOQ_LAUNCH_CHECK();
  return 0;
}
