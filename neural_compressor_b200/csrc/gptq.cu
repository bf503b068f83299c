// gptq.cu -- K3: GPTQ.fasterquant column loop + lazy batch update  (gptq.py:1143-1351)
//
//   for each block [i1,i2) of `blocksize` columns:                                   gptq.py:1250
//     find_params for every group that STARTS in the block, from the global W as it is at block start
//       (the reference reads W, not the in-block working copy W1 -- gptq.py:1270)
//     for each 128-column sub-block:  column loop (quantize, err, rank-1 update)      gptq.py:1260-1299
//       then W[:, sub_end:i2] -= Err_sub @ Hinv[sub, sub_end:i2]   (in-block propagation, == the rank-1
//       updates the reference applies to the rest of W1, re-associated)
//     W[:, i2:] -= Err1 @ Hinv[i1:i2, i2:]                                           gptq.py:1304
//
// Column-loop kernel: rows are independent, so a warp owns RW rows and keeps each row's 128-column
// sub-block in registers (lane l holds columns l, l+32, l+64, l+96).  Column i is broadcast from its
// owner lane with one warp shuffle, every lane recomputes q/err (identical, cheaper than a second
// shuffle) and updates its own columns j > i with  w = fl(w - fl(err * Hinv[i,j]))  -- explicit
// __fmul_rn/__fsub_rn, never an FMA, which is the rounded-product-then-rounded-subtract the reference's
// K=1 matmul + in-place subtract performs (gptq.py:1297-1298).  Hinv's 128x128 diagonal block lives in
// shared memory for the whole sub-block.
#include <climits>

#include "common.cuh"

namespace b200woq {

constexpr int SUB = 128;  // columns per register-resident sub-block
constexpr int RW = 2;     // rows per warp (interleaved for ILP); N/2 warps keep every SM busy

// Correctly rounded fp32 division by a divisor that is reused many times (Markstein 1990): with r = RN(1/d),
// q0 = RN(a*r), rem = a - q0*d (exact in one FMA), RN(q0 + rem*r) == RN(a/d) unless d's significand is all ones
// or the operands are near the ends of the exponent range -- those cases take the IEEE division.  Three
// instructions per quotient instead of ~12, bit-identical to the `/` the reference executes.
struct RnDivisor {
  float d, r;
  bool slow;
};
__device__ __forceinline__ RnDivisor make_divisor(float d) {
  RnDivisor D;
  D.d = d;
  D.r = __frcp_rn(d);
  const uint32_t bits = __float_as_uint(d);
  const uint32_t ex = (bits >> 23) & 0xffu;
  D.slow = ((bits & 0x7fffffu) == 0x7fffffu) || ex < 32u || ex > 222u;
  return D;
}
__device__ __forceinline__ float div_rn(float a, const RnDivisor& D) {
  const uint32_t ea = (__float_as_uint(a) >> 23) & 0xffu;
  if (D.slow || (ea < 64u && a != 0.f) || ea > 190u) return __fdiv_rn(a, D.d);
  const float q0 = __fmul_rn(a, D.r);
  const float rem = __fmaf_rn(-q0, D.d, a);
  return __fmaf_rn(rem, D.r, q0);
}

struct GptqQ {
  float maxq;
  int sym;
};

// Quantizer.find_params (gptq.py:1501-1596), int dtype, perchannel, weight=True; optional mse search.
// One warp per (row, group).  Writes scale/zero at [n*G + gi].
__global__ void __launch_bounds__(256)
    gptq_find_params_kernel(const float* __restrict__ W, int64_t N, int64_t C, int64_t col0, int g, int64_t gi0,
                            int ngroups, int64_t G, float maxq, int sym, int mse, float* __restrict__ scale,
                            float* __restrict__ zero) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t task = warp; task < N * ngroups; task += nwarps) {
    const int64_t n = task / ngroups;
    const int gl = (int)(task % ngroups);
    const int64_t c0 = col0 + (int64_t)gl * g;
    const int64_t c1 = (c0 + g < C) ? c0 + g : C;
    const float* row = W + n * C;
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t c = c0 + lane; c < c1; c += 32) {
      const float v = row[c];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    float xmin = fminf(warp_min(mn), 0.f), xmax = fmaxf(warp_max(mx), 0.f);
    if (sym) {  // gptq.py:1548-1552
      xmax = fmaxf(fabsf(xmin), xmax);
      if (xmin < 0.f) xmin = -xmax;
    }
    if (xmin == 0.f && xmax == 0.f) {
      xmin = -1.f;
      xmax = 1.f;
    }
    float s = __fdiv_rn(__fsub_rn(xmax, xmin), maxq);
    float z = sym ? (maxq + 1.f) * 0.5f : rintf(__fdiv_rn(-xmin, s));
    if (mse) {  // gptq.py:1567-1584: grid=100, maxshrink=.8, norm=2.4
      float best = INFINITY;
      for (int i = 0; i < 80; ++i) {
        const float pr = (float)(1.0 - (double)i / 100.0);
        const float xmin1 = __fmul_rn(pr, xmin), xmax1 = __fmul_rn(pr, xmax);
        const float s1 = __fdiv_rn(__fsub_rn(xmax1, xmin1), maxq);
        const float z1 = sym ? z : rintf(__fdiv_rn(-xmin1, s1));
        float err = 0.f;
        for (int64_t c = c0 + lane; c < c1; c += 32) {
          const float v = row[c];
          const float q = fminf(fmaxf(__fadd_rn(rintf(__fdiv_rn(v, s1)), z1), 0.f), maxq);
          const float d = fabsf(__fsub_rn(__fmul_rn(s1, __fsub_rn(q, z1)), v));
          err += powf(d, 2.4f);
        }
        err = warp_sum(err);
        if (err < best) {
          best = err;
          s = s1;
          z = z1;
        }
      }
    }
    if (lane == 0) {
      scale[n * G + gi0 + gl] = s;
      zero[n * G + gi0 + gl] = z;
    }
  }
}

__global__ void zero_dead_columns_kernel(float* __restrict__ W, int64_t N, int64_t C, const uint8_t* __restrict__ dead) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N * C; i += (int64_t)gridDim.x * blockDim.x)
    if (dead[i % C]) W[i] = 0.f;
}

// column loop over one sub-block [c0, c0+ncols), ncols <= 128
__global__ void __launch_bounds__(256)
    gptq_subblock_kernel(float* __restrict__ W, const float* __restrict__ Hinv, int64_t N, int64_t C, int64_t c0,
                         int ncols, int g, int64_t G, float maxq, const float* __restrict__ scale,
                         const float* __restrict__ zero, uint8_t* __restrict__ codes, float* __restrict__ Q,
                         float* __restrict__ Err, int64_t err_ld, int64_t err_col0, float* __restrict__ losses) {
  extern __shared__ float hs[];  // [ncols][SUB+1] upper-triangular diagonal block of Hinv
  for (int e = threadIdx.x; e < ncols * SUB; e += blockDim.x) {
    const int r = e / SUB, c = e % SUB;
    hs[r * (SUB + 1) + c] = (c < ncols && c >= r) ? Hinv[(c0 + r) * C + c0 + c] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * RW;
  if (row0 >= N) return;

  float w[RW][4], qv[RW][4], ev[RW][4];
  uint32_t cd[RW];
  float loss[RW];
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int64_t n = row0 + r;
    cd[r] = 0;
    loss[r] = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int c = lane + 32 * s;
      w[r][s] = (n < N && c < ncols) ? W[n * C + c0 + c] : 0.f;
      qv[r][s] = 0.f;
      ev[r][s] = 0.f;
    }
  }
  float sc[RW], zr[RW];
  RnDivisor dsc[RW];
  const bool per_channel = (g <= 0);
  {  // parameters of the group that contains the first column (it may have started in an earlier sub-block)
    const int64_t gi_first = per_channel ? 0 : c0 / g;
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const int64_t n = (row0 + r < N) ? row0 + r : N - 1;
      sc[r] = scale[n * G + gi_first];
      zr[r] = zero[n * G + gi_first];
      dsc[r] = make_divisor(sc[r]);
    }
  }
  int next_group_col = per_channel ? INT_MAX : (int)(((c0 + g - 1) / g) * g - c0);  // first group start >= c0
  const bool want_loss = (losses != nullptr);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    for (int l = 0; l < 32; ++l) {
      const int i = 32 * s + l;
      if (i >= ncols) break;
      if (i == next_group_col) {  // gptq.py:1264-1272
        const int64_t gi = (c0 + i) / g;
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          const int64_t n = (row0 + r < N) ? row0 + r : N - 1;
          sc[r] = scale[n * G + gi];
          zr[r] = zero[n * G + gi];
          dsc[r] = make_divisor(sc[r]);
        }
        next_group_col += g;
      }
      const float d = hs[i * (SUB + 1) + i];
      const RnDivisor dd = make_divisor(d);
      const float inv_d2h = want_loss ? __fdividef(0.5f, d * d) : 0.f;
      float h[4];
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2) h[s2] = hs[i * (SUB + 1) + lane + 32 * s2];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const float wi = __shfl_sync(0xffffffffu, w[r][s], l);
        // Quantizer.quantize (gptq.py:1636-1637)
        const float qi = fminf(fmaxf(__fadd_rn(rintf(div_rn(wi, dsc[r])), zr[r]), 0.f), maxq);
        const float q = __fmul_rn(sc[r], __fsub_rn(qi, zr[r]));
        const float diff = __fsub_rn(wi, q);
        const float err = div_rn(diff, dd);  // gptq.py:1296
        if (lane == l) {
          qv[r][s] = q;
          ev[r][s] = err;
          cd[r] |= ((uint32_t)qi & 0xffu) << (8 * s);
          loss[r] = fmaf(diff * diff, inv_d2h, loss[r]);  // gptq.py:1294,1303 (diagnostic, not bit-pinned)
        }
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) {
          const int j = lane + 32 * s2;
          if (j > i) w[r][s2] = __fsub_rn(w[r][s2], __fmul_rn(err, h[s2]));  // gptq.py:1297-1298, no FMA
        }
      }
    }
  }
  // write back: quantised values replace the working columns (they are final), errors go to Err
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int64_t n = row0 + r;
    if (n >= N) continue;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int c = lane + 32 * s;
      if (c < ncols) {
        if (Q) Q[n * C + c0 + c] = qv[r][s];
        codes[n * C + c0 + c] = (uint8_t)((cd[r] >> (8 * s)) & 0xffu);
        Err[n * err_ld + err_col0 + c] = ev[r][s];
      }
    }
    if (losses) {
      const float tot = warp_sum(loss[r]);
      if (lane == 0) losses[n] += tot;
    }
  }
}

// W[:, j0:j1] -= Err[:, e0:e0+KK] @ Hinv[r0:r0+KK, j0:j1]   (fp32 FFMA tiles: 128 x 128, K chunks of 16)
__global__ void __launch_bounds__(256)
    gptq_lazy_update_kernel(float* __restrict__ W, const float* __restrict__ Err, const float* __restrict__ Hinv,
                            int64_t N, int64_t C, int64_t err_ld, int64_t e0, int64_t r0, int KK, int64_t j0,
                            int64_t j1) {
  __shared__ float As[16][128 + 4];  // Err^T chunk  [k][row]
  __shared__ float Bs[16][128 + 4];  // Hinv chunk   [k][col]
  const int64_t n0 = (int64_t)blockIdx.y * 128, jb = j0 + (int64_t)blockIdx.x * 128;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8] = {};
  for (int k0 = 0; k0 < KK; k0 += 16) {
    for (int e = threadIdx.x; e < 16 * 128; e += 256) {
      {  // Err tile: coalesced along k within a row is not possible (row-major [N, ld]); read 16 k per row
        const int row = e >> 4, k = e & 15;
        const int64_t n = n0 + row;
        As[k][row] = (n < N && k0 + k < KK) ? Err[n * err_ld + e0 + k0 + k] : 0.f;
      }
      {
        const int k = e >> 7, c = e & 127;
        const int64_t j = jb + c;
        Bs[k][c] = (j < j1 && k0 + k < KK) ? Hinv[(r0 + k0 + k) * C + j] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a[u] = As[k][ty * 8 + u];
        b[u] = Bs[k][tx + 16 * u];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int64_t n = n0 + ty * 8 + u;
    if (n >= N) continue;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const int64_t j = jb + tx + 16 * v;
      if (j < j1) W[n * C + j] -= acc[u][v];
    }
  }
}

}  // namespace b200woq

using namespace b200woq;

extern "C" int64_t b200woq_gptq_workspace_bytes(int64_t N, int64_t C, int blocksize) {
  const int64_t bs = blocksize <= 0 ? C : (blocksize > C ? C : blocksize);
  return N * bs * (int64_t)sizeof(float) + 256;
}

extern "C" int b200woq_gptq_fasterquant(float* W, const float* Hinv, const uint8_t* dead_mask, int64_t N, int64_t C,
                                        int blocksize, int groupsize, int bits, int sym, int flags, uint8_t* codes, float* Q,
                                        float* scale, float* zero, float* losses, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
  WOQ_CHECK_ARG(W && Hinv && codes && scale && zero && N > 0 && C > 0, "gptq_fasterquant: null pointer / empty shape");
  WOQ_CHECK_ARG(bits >= 1 && bits <= 8, "gptq_fasterquant: bits must be in [1,8]");
  WOQ_CHECK_ARG(blocksize > 0, "gptq_fasterquant: blocksize must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t bs = blocksize > C ? C : blocksize;
  if (workspace_bytes < b200woq_gptq_workspace_bytes(N, C, (int)bs) - 256 || !workspace) {
    set_error("gptq_fasterquant: workspace too small");
    return B200WOQ_EWORKSPACE;
  }
  float* Err = (float*)workspace;  // [N, bs]
  const float maxq = (float)((1 << bits) - 1);
  const bool per_channel = groupsize <= 0;
  const int g = per_channel ? (int)C : groupsize;
  const int64_t G = per_channel ? 1 : ceil_div(C, g);
  const int mse = flags & 1;
  if (losses) WOQ_CUDA(cudaMemsetAsync(losses, 0, sizeof(float) * N, st));
  const int fp_blocks = (int)std::min<int64_t>(ceil_div(N * 32, 256) * 4, (int64_t)num_sms() * 8);

  if (per_channel) {  // gptq.py:1184-1185: one find_params over the whole row, BEFORE dead columns are zeroed
    gptq_find_params_kernel<<<fp_blocks, 256, 0, st>>>(W, N, C, 0, (int)C, 0, 1, 1, maxq, sym, mse, scale, zero);
    WOQ_LAUNCH_CHECK();
  }
  if (dead_mask) {  // gptq.py:1191  W[:, dead] = 0
    zero_dead_columns_kernel<<<fp_blocks, 256, 0, st>>>(W, N, C, dead_mask);
    WOQ_LAUNCH_CHECK();
  }
  const int rows_per_cta = 8 * RW;
  for (int64_t i1 = 0; i1 < C; i1 += bs) {
    const int64_t i2 = std::min(i1 + bs, C);
    if (!per_channel) {
      // groups whose first column lies in [i1, i2)
      const int64_t gfirst = ceil_div(i1, g), glast = (i2 - 1) / g;
      if (gfirst <= glast) {
        gptq_find_params_kernel<<<fp_blocks, 256, 0, st>>>(W, N, C, gfirst * g, g, gfirst, (int)(glast - gfirst + 1), G,
                                                           maxq, sym, mse, scale, zero);
        WOQ_LAUNCH_CHECK();
      }
    }
    for (int64_t c0 = i1; c0 < i2; c0 += SUB) {
      const int ncols = (int)std::min<int64_t>(SUB, i2 - c0);
      const size_t smem = (size_t)ncols * (SUB + 1) * sizeof(float);
      if (smem > 48 * 1024)
        WOQ_CUDA(cudaFuncSetAttribute(gptq_subblock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      gptq_subblock_kernel<<<(unsigned)ceil_div(N, rows_per_cta), 256, smem, st>>>(
          W, Hinv, N, C, c0, ncols, per_channel ? 0 : g, G, maxq, scale, zero, codes, Q, Err, bs, c0 - i1, losses);
      WOQ_LAUNCH_CHECK();
      const int64_t c1 = c0 + ncols;
      if (c1 < i2) {  // in-block propagation to the rest of the block
        dim3 grid((unsigned)ceil_div(i2 - c1, 128), (unsigned)ceil_div(N, 128));
        gptq_lazy_update_kernel<<<grid, 256, 0, st>>>(W, Err, Hinv, N, C, bs, c0 - i1, c0, ncols, c1, i2);
        WOQ_LAUNCH_CHECK();
      }
    }
    if (i2 < C) {  // gptq.py:1304
      dim3 grid((unsigned)ceil_div(C - i2, 128), (unsigned)ceil_div(N, 128));
      gptq_lazy_update_kernel<<<grid, 256, 0, st>>>(W, Err, Hinv, N, C, bs, 0, i1, (int)(i2 - i1), i2, C);
      WOQ_LAUNCH_CHECK();
    }
  }
  return 0;
}
