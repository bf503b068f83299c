// cholinv.cu -- K2: upper Cholesky factor U of H^-1 (U^T U = H^-1), the last step of the fasterquant prologue
//   reference: gptq.py:1228-1231   H = cholesky(H); H = cholesky_inverse(H); H = cholesky(H, upper=True); Hinv = H
//
// U is unique, and it is the inverse of the "UL" factor of H:  H = R R^T with R upper  =>  H^-1 = R^-T R^-1, U = R^-1.
// With J the index reversal, J H J = L L^T (ordinary lower Cholesky) gives R = J L J, so
//       U = J L^-1 J
// i.e. ONE Cholesky factorisation + ONE triangular inverse (2/3 C^3 flop) instead of the reference's three LAPACK
// calls (4/3 C^3), with fewer roundings.  Everything is exact-fp32 FFMA (no tensor cores: the truncating fp32
// accumulate of tcgen05 / a 3xTF32 split is ~1e-5-grade, GPTQ codes need fp32-grade factors).
//
// Layout: the flipped matrix A = J H J lives in a workspace padded to a multiple of 128 (identity on the padding), so
// every kernel works on full 128x128 tiles with float4 loads and no bounds checks.
//   1. flip_pad              A <- J H J
//   2. for k in tiles:       potrf_inv_tile  A_kk <- inv(chol(A_kk))            (one CTA, shared memory)
//                            panel           A_ik <- A_ik * inv(L_kk)^T  (i > k)  (= the trsm, as a GEMM)
//                            trail           A_ij <- A_ij - A_ik A_jk^T  (i >= j > k)
//      the diagonal tiles end up holding inv(L_kk): nobody needs L_kk itself any more.
//   3. triangular inverse X = L^-1 by recursive doubling over aligned tile ranges [lo,mid) [mid,hi):
//         X21 = -X22 * (L21 * X11)       two batched tile GEMMs per level (phase 1 into a second buffer B, phase 2 back
//      into A), K ranges trimmed to the non-zero (lower-triangular) tiles.  log2(tiles) levels, all nodes of a level in
//      one launch -> full-chip parallelism, unlike the textbook column sweep whose critical path is tiles^2/2 tile GEMMs.
//   4. unflip                U[r,c] = X[C-1-r, C-1-c] for r <= c, 0 below the diagonal.
// A non-positive pivot sets *info = (flipped) column + 1 (the reference raises there, torch.linalg.cholesky); the host
// wrapper checks it once per block.
#include "common.cuh"

namespace b200woq {
namespace cholinv {

constexpr int TB = 128;       // tile edge
constexpr int BK = 16;        // k chunk of the FFMA tile GEMM
constexpr int LDS_A = TB + 4;  // padded smem row of the k-major operand tiles (conflict-free transposed stores)

__global__ void flip_pad_kernel(const float* __restrict__ H, int64_t C, float* __restrict__ A, int64_t Cp) {
  const int64_t total = Cp * Cp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / Cp, j = idx - i * Cp;
    float v;
    if (i < C && j < C) v = H[(C - 1 - i) * C + (C - 1 - j)];
    else v = (i == j) ? 1.f : 0.f;
    A[idx] = v;
  }
}

__global__ void unflip_kernel(const float* __restrict__ X, int64_t Cp, float* __restrict__ U, int64_t C) {
  const int64_t total = C * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / C, c = idx - r * C;
    U[idx] = (r <= c) ? X[(C - 1 - r) * Cp + (C - 1 - c)] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 128x128xK FFMA tile GEMM:  acc[8][8] (this thread's micro-tile) = sum_k Aop[row][k] * Bop[k][col]
//   A: row-major [128, K], k contiguous (lda)                     -> transposed into As[k][row]
//   B: B_NK ? row-major [128 (n), K], k contiguous : row-major [K, 128 (n)], n contiguous (ldb)
// 256 threads, thread (ty, tx) = (tid >> 4, tid & 15) owns rows {ty*4..+3, 64+ty*4..+3} x cols {tx*4..+3, 64+tx*4..+3}.
// Register double buffering of the next k chunk; K % 16 == 0.
template <bool B_NK>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                                          int K, float (&acc)[8][8], float (*As)[LDS_A], float (*Bs)[LDS_A]) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  // k-contiguous loader: row = tid >> 1, k quads {kq, kq + 8}, kq = (tid & 1) * 4
  const int lrow = tid >> 1, lkq = (tid & 1) * 4;
  // n-contiguous loader (B, !B_NK): k rows {tid >> 5, (tid >> 5) + 8}, n = (tid & 31) * 4
  const int bk = tid >> 5, bn = (tid & 31) * 4;
  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
    ra[0] = *reinterpret_cast<const float4*>(A + (int64_t)lrow * lda + k0 + lkq);
    ra[1] = *reinterpret_cast<const float4*>(A + (int64_t)lrow * lda + k0 + lkq + 8);
    if (B_NK) {
      rb[0] = *reinterpret_cast<const float4*>(B + (int64_t)lrow * ldb + k0 + lkq);
      rb[1] = *reinterpret_cast<const float4*>(B + (int64_t)lrow * ldb + k0 + lkq + 8);
    } else {
      rb[0] = *reinterpret_cast<const float4*>(B + (int64_t)(k0 + bk) * ldb + bn);
      rb[1] = *reinterpret_cast<const float4*>(B + (int64_t)(k0 + bk + 8) * ldb + bn);
    }
  };
#pragma unroll
  for (int u = 0; u < 8; ++u)
#pragma unroll
    for (int v = 0; v < 8; ++v) acc[u][v] = 0.f;
  gload(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int kk = lkq + 8 * u;
      As[kk + 0][lrow] = ra[u].x;
      As[kk + 1][lrow] = ra[u].y;
      As[kk + 2][lrow] = ra[u].z;
      As[kk + 3][lrow] = ra[u].w;
      if (B_NK) {
        Bs[kk + 0][lrow] = rb[u].x;
        Bs[kk + 1][lrow] = rb[u].y;
        Bs[kk + 2][lrow] = rb[u].z;
        Bs[kk + 3][lrow] = rb[u].w;
      } else {
        *reinterpret_cast<float4*>(&Bs[bk + 8 * u][bn]) = rb[u];
      }
    }
    __syncthreads();
    if (k0 + BK < K) gload(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
}

// epilogue: MODE 0: C = acc ; 1: C -= acc ; 2: C = -acc          (C row-major tile, ldc)
template <int MODE>
__device__ __forceinline__ void tile_store(float* __restrict__ Ct, int64_t ldc, const float (&acc)[8][8]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int r = (u < 4) ? ty * 4 + u : 64 + ty * 4 + (u - 4);
#pragma unroll
    for (int hv = 0; hv < 2; ++hv) {
      float* dst = Ct + (int64_t)r * ldc + (hv == 0 ? tx * 4 : 64 + tx * 4);
      float4 t;
      if (MODE == 1) {
        t = *reinterpret_cast<float4*>(dst);
        t.x -= acc[u][4 * hv + 0];
        t.y -= acc[u][4 * hv + 1];
        t.z -= acc[u][4 * hv + 2];
        t.w -= acc[u][4 * hv + 3];
      } else if (MODE == 2) {
        t = make_float4(-acc[u][4 * hv + 0], -acc[u][4 * hv + 1], -acc[u][4 * hv + 2], -acc[u][4 * hv + 3]);
      } else {
        t = make_float4(acc[u][4 * hv + 0], acc[u][4 * hv + 1], acc[u][4 * hv + 2], acc[u][4 * hv + 3]);
      }
      *reinterpret_cast<float4*>(dst) = t;
    }
  }
}

// A_ik <- A_ik * Dinv_k^T  (i = k+1+blockIdx.x); Dinv_k = A_kk holds inv(L_kk) (lower, zeros above).  In place: the CTA
// reads its whole tile before it writes it.
__global__ void __launch_bounds__(256, 2) chol_panel_kernel(float* __restrict__ A, int64_t Cp, int k) {
  __shared__ __align__(16) float As[BK][LDS_A];
  __shared__ __align__(16) float Bs[BK][LDS_A];
  const int64_t i = k + 1 + blockIdx.x;
  float* tile = A + i * TB * Cp + (int64_t)k * TB;
  const float* dinv = A + (int64_t)k * TB * Cp + (int64_t)k * TB;
  float acc[8][8];
  tile_gemm<true>(tile, Cp, dinv, Cp, TB, acc, As, Bs);
  tile_store<0>(tile, Cp, acc);
}

// A_ij -= A_ik A_jk^T for i >= j > k   (grid: x = j - k - 1, y = i - k - 1; CTAs above the diagonal exit)
__global__ void __launch_bounds__(256, 2) chol_trail_kernel(float* __restrict__ A, int64_t Cp, int k) {
  if (blockIdx.x > blockIdx.y) return;
  __shared__ __align__(16) float As[BK][LDS_A];
  __shared__ __align__(16) float Bs[BK][LDS_A];
  const int64_t i = k + 1 + blockIdx.y, j = k + 1 + blockIdx.x;
  float acc[8][8];
  tile_gemm<true>(A + i * TB * Cp + (int64_t)k * TB, Cp, A + j * TB * Cp + (int64_t)k * TB, Cp, TB, acc, As, Bs);
  tile_store<1>(A + i * TB * Cp + j * TB, Cp, acc);
}

// level s of the recursive triangular inverse; node z: lo = z << (s+1), mid = lo + (1 << s), hi = min(lo + (2 << s), nt)
//   phase 1: T[i,j] = sum_{m=j}^{mid-1} L[i,m] X11[m,j]     i in [mid,hi), j in [lo,mid)     -> Bbuf
//   phase 2: X[i,j] = -sum_{m=mid}^{i} X22[i,m] T[m,j]                                         -> A
template <int PHASE>
__global__ void __launch_bounds__(256, 2) trinv_level_kernel(float* __restrict__ A, float* __restrict__ Bbuf, int64_t Cp,
                                                              int nt, int s) {
  const int lo = (int)blockIdx.z << (s + 1), mid = lo + (1 << s);
  if (mid >= nt) return;
  const int hi = min(lo + (2 << s), nt);
  const int i = mid + blockIdx.y, j = lo + blockIdx.x;
  if (i >= hi) return;
  __shared__ __align__(16) float As[BK][LDS_A];
  __shared__ __align__(16) float Bs[BK][LDS_A];
  float acc[8][8];
  if (PHASE == 1) {
    tile_gemm<false>(A + (int64_t)i * TB * Cp + (int64_t)j * TB, Cp, A + (int64_t)j * TB * Cp + (int64_t)j * TB, Cp,
                     (mid - j) * TB, acc, As, Bs);
    tile_store<0>(Bbuf + (int64_t)i * TB * Cp + (int64_t)j * TB, Cp, acc);
  } else {
    tile_gemm<false>(A + (int64_t)i * TB * Cp + (int64_t)mid * TB, Cp, Bbuf + (int64_t)mid * TB * Cp + (int64_t)j * TB, Cp,
                     (i - mid + 1) * TB, acc, As, Bs);
    tile_store<2>(A + (int64_t)i * TB * Cp + (int64_t)j * TB, Cp, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Diagonal tile: A_kk <- X = inv(chol(A_kk)) in shared memory, one CTA of 256 threads (the critical path of the whole
// factorisation: nothing else runs while it does, so it is written for latency).
//   Recursive 2 x 2 splitting down to 32 x 32 base blocks:
//       [A11 .  ]      L11 = chol(A11), X11 = inv(L11)                  (recursion / base)
//       [A21 A22]      L21 = A21 X11^T ; A22 -= L21 L21^T               (shared-memory FFMA GEMMs, 4 x 4 per thread)
//                      L22 = chol(A22), X22 = inv(L22)                  (recursion / base)
//                      X21 = -X22 (L21 X11)
//   Base block (one warp, registers + shuffles, fully unrolled): lane i holds row i; right-looking Cholesky with the
//   pivot / column broadcast by __shfl, then the inverse by forward substitution with lane = column.  IEEE sqrt and
//   division.  Only X (and the L21 blocks it needs on the way) is ever materialised: L_kk itself is not needed.
constexpr int SP = TB + 1;

// C[M x N] (+)= alpha * A[M x K] * op(B); all operands in shared memory with row stride SP; M, N multiples of 4.
//   BT = false: op(B) = B[K x N];  BT = true: op(B) = B^T with B[N x K].   ACC: C += (else C =).
template <bool BT, bool ACC>
__device__ __forceinline__ void smem_gemm(const float* A, const float* B, float* Cm, int M, int N, int K, float alpha) {
  const int tiles_n = N >> 2, tiles = (M >> 2) * tiles_n;
  for (int t = threadIdx.x; t < tiles; t += 256) {
    const int i0 = (t / tiles_n) << 2, j0 = (t % tiles_n) << 2;
    float acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
    for (int k = 0; k < K; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = A[(i0 + u) * SP + k];
#pragma unroll
      for (int v = 0; v < 4; ++v) b[v] = BT ? B[(j0 + v) * SP + k] : B[k * SP + j0 + v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float* c = Cm + (i0 + u) * SP + j0 + v;
        *c = ACC ? fmaf(alpha, acc[u][v], *c) : alpha * acc[u][v];
      }
  }
}

// 32 x 32 base block at (lo, lo): reads the lower triangle of S, writes X = inv(chol(.)) (lower, zeros above) to Xo.
// One warp; fully unrolled so that the row / column live in registers.
__device__ __forceinline__ void invchol32(const float* S, float* Xo, int lo, int lane, int* info, int col_base) {
  float a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = S[(lo + lane) * SP + lo + c];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float piv = __shfl_sync(0xffffffffu, a[j], j);
    if (!(piv > 0.f)) bad = true;
    const float d = __fsqrt_rn(piv);
    const float l = (lane == j) ? d : __fdiv_rn(a[j], d);      // L[lane][j] (rows above the diagonal: unused garbage)
    a[j] = l;
#pragma unroll
    for (int c = j + 1; c < 32; ++c) {
      const float lc = __shfl_sync(0xffffffffu, l, c);         // L[c][j]
      a[c] = fmaf(-l, lc, a[c]);
    }
    if (bad && lane == 0 && j == 31) atomicCAS(info, 0, col_base + lo + 1);
  }
  // inverse, lane = column: x[r] = X[r][lane]
  float x[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    float sum = 0.f;
#pragma unroll
    for (int m = 0; m < r; ++m) sum = fmaf(__shfl_sync(0xffffffffu, a[m], r), (m >= lane) ? x[m] : 0.f, sum);
    const float lrr = __shfl_sync(0xffffffffu, a[r], r);
    x[r] = (r == lane) ? __fdiv_rn(1.f, lrr) : ((r > lane) ? __fdiv_rn(-sum, lrr) : 0.f);
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) Xo[(lo + r) * SP + lo + lane] = x[r];
}

// X = inv(chol(S[lo:lo+n, lo:lo+n])) for n = 64: two base blocks + the 32-wide off-diagonal work
__device__ __forceinline__ void invchol64(float* S, float* X, float* Tm, int lo, int* info, int col_base) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mid = lo + 32;
  if (warp == 0) invchol32(S, X, lo, lane, info, col_base);
  __syncthreads();
  smem_gemm<true, false>(S + mid * SP + lo, X + lo * SP + lo, Tm, 32, 32, 32, 1.f);       // L21 = A21 X11^T
  __syncthreads();
  smem_gemm<true, true>(Tm, Tm, S + mid * SP + mid, 32, 32, 32, -1.f);                    // A22 -= L21 L21^T
  __syncthreads();
  if (warp == 0) invchol32(S, X, mid, lane, info, col_base);
  // T2 = L21 X11 (while warp 0 factors A22: warps 1..7 could run it, kept simple: everyone after the barrier)
  __syncthreads();
  smem_gemm<false, false>(Tm, X + lo * SP + lo, Tm + 32 * SP, 32, 32, 32, 1.f);           // T2 = L21 X11
  __syncthreads();
  smem_gemm<false, false>(X + mid * SP + mid, Tm + 32 * SP, X + mid * SP + lo, 32, 32, 32, -1.f);   // X21 = -X22 T2
  __syncthreads();
}

__global__ void __launch_bounds__(256, 1) potrf_inv_tile_kernel(float* __restrict__ A, int64_t Cp, int k, int* __restrict__ info) {
  extern __shared__ float smem[];
  float* S = smem;                 // [128][129] the tile (lower triangle is maintained)
  float* X = smem + TB * SP;       // [128][129] the inverse
  float* Tm = smem + 2 * TB * SP;  // [128][129] scratch: L21 (64 rows) and L21 X11 (64 rows)
  float* tile = A + (int64_t)k * TB * Cp + (int64_t)k * TB;
  const int tid = threadIdx.x;
  for (int e = tid; e < TB * TB / 4; e += 256) {
    const int r = e >> 5, c4 = (e & 31) * 4;
    const float4 v = *reinterpret_cast<const float4*>(tile + (int64_t)r * Cp + c4);
    S[r * SP + c4 + 0] = v.x;
    S[r * SP + c4 + 1] = v.y;
    S[r * SP + c4 + 2] = v.z;
    S[r * SP + c4 + 3] = v.w;
  }
  for (int e = tid; e < TB * SP; e += 256) X[e] = 0.f;
  __syncthreads();
  const int col_base = k * TB;
  invchol64(S, X, Tm, 0, info, col_base);                                                  // X11 (64 x 64)
  smem_gemm<true, false>(S + 64 * SP, X, Tm, 64, 64, 64, 1.f);                             // L21 = A21 X11^T
  __syncthreads();
  smem_gemm<true, true>(Tm, Tm, S + 64 * SP + 64, 64, 64, 64, -1.f);                       // A22 -= L21 L21^T
  __syncthreads();
  // Tm rows 0..63 hold L21 and must survive the second half: invchol64 uses scratch rows 64..127
  invchol64(S, X, Tm + 64 * SP, 64, info, col_base);                                       // X22
  float* T2 = S;                                                                           // A11 area is dead now: reuse it
  smem_gemm<false, false>(Tm, X, T2, 64, 64, 64, 1.f);                                     // T2 = L21 X11
  __syncthreads();
  smem_gemm<false, false>(X + 64 * SP + 64, T2, X + 64 * SP, 64, 64, 64, -1.f);            // X21 = -X22 T2
  __syncthreads();
  for (int e = tid; e < TB * TB / 4; e += 256) {
    const int r = e >> 5, c4 = (e & 31) * 4;
    *reinterpret_cast<float4*>(tile + (int64_t)r * Cp + c4) =
        make_float4(X[r * SP + c4], X[r * SP + c4 + 1], X[r * SP + c4 + 2], X[r * SP + c4 + 3]);
  }
}

}  // namespace cholinv
}  // namespace b200woq

using namespace b200woq;

static inline int64_t pad128(int64_t C) { return ceil_div(C, 128) * 128; }

extern "C" int64_t b200woq_cholinv_workspace_bytes(int64_t C) {
  const int64_t Cp = pad128(C);
  return (2 * Cp * Cp) * (int64_t)sizeof(float) + 256;
}

extern "C" int b200woq_cholinv_upper(const float* H, int64_t C, float* U, void* workspace, int64_t workspace_bytes,
                                     int* info, void* stream) {
  using namespace cholinv;
  WOQ_CHECK_ARG(H && U && workspace && info && C > 0, "cholinv_upper: bad arguments");
  WOQ_CHECK_ARG(workspace_bytes >= b200woq_cholinv_workspace_bytes(C), "cholinv_upper: workspace too small");
  WOQ_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "cholinv_upper: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t Cp = pad128(C);
  const int nt = (int)(Cp / TB);
  float* A = (float*)workspace;
  float* Bbuf = A + Cp * Cp;
  const int blocks = num_sms() * 8;
  WOQ_CUDA(cudaMemsetAsync(info, 0, sizeof(int), st));
  flip_pad_kernel<<<blocks, 256, 0, st>>>(H, C, A, Cp);
  WOQ_LAUNCH_CHECK();
  const size_t tile_smem = 3 * (size_t)TB * SP * sizeof(float);
  WOQ_CUDA(cudaFuncSetAttribute(potrf_inv_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
  for (int k = 0; k < nt; ++k) {
    potrf_inv_tile_kernel<<<1, 256, tile_smem, st>>>(A, Cp, k, info);
    WOQ_LAUNCH_CHECK();
    const int rem = nt - k - 1;
    if (rem > 0) {
      chol_panel_kernel<<<rem, 256, 0, st>>>(A, Cp, k);
      WOQ_LAUNCH_CHECK();
      chol_trail_kernel<<<dim3(rem, rem), 256, 0, st>>>(A, Cp, k);
      WOQ_LAUNCH_CHECK();
    }
  }
  for (int s = 0; (1 << s) < nt; ++s) {
    const int half = 1 << s;
    const int nodes = (int)ceil_div(nt, 2 * half);
    trinv_level_kernel<1><<<dim3(half, half, nodes), 256, 0, st>>>(A, Bbuf, Cp, nt, s);
    WOQ_LAUNCH_CHECK();
    trinv_level_kernel<2><<<dim3(half, half, nodes), 256, 0, st>>>(A, Bbuf, Cp, nt, s);
    WOQ_LAUNCH_CHECK();
  }
  unflip_kernel<<<blocks, 256, 0, st>>>(A, Cp, U, C);
  WOQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200woq_hessian_finalize_cholinv_upper(float* Hsum, int64_t C, double nsamples, float percdamp,
                                                      uint8_t* dead_mask, float* scratch, float* U, void* workspace,
                                                      int64_t workspace_bytes, int* info, void* stream) {
  if (int rc = b200woq_hessian_finalize(Hsum, C, nsamples, percdamp, dead_mask, scratch, stream)) return rc;
  return b200woq_cholinv_upper(Hsum, C, U, workspace, workspace_bytes, info, stream);
}
