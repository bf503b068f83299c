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

#include <cuda.h>

#include "common.cuh"
#include "rtn_math.cuh"

namespace b200woq {

constexpr int SUB = 128;  // columns per register-resident sub-block
constexpr int RW = 2;     // rows per warp: independent dependency chains interleaved in one warp
constexpr int SUB_WARPS = 8;  // warps per CTA of the column-loop kernel

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

// Quantizer.find_params with use_double_quant (gptq.py:1598-1614): the freshly computed scales of one group, taken as the
// vector [1, N] over the output rows, are fake-quantised by quant_tensor(dtype int, bits, group_size, scheme, quantile=1,
// return_int=False, full_range=False): consecutive `dq_g` rows form a group (ragged tail = its own group).
// One warp per (weight group, row chunk).
__global__ void __launch_bounds__(256)
    gptq_double_quant_kernel(float* __restrict__ scale, int64_t N, int64_t G, int64_t gi0, int ngroups, int dq_bits, int dq_g,
                             int dq_sym) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t chunks = (N + dq_g - 1) / dq_g;
  const QRange r = qrange(dq_bits, dq_sym != 0);
  for (int64_t task = warp; task < chunks * ngroups; task += nwarps) {
    const int64_t ch = task / ngroups, gi = gi0 + task % ngroups;
    const int64_t n0 = ch * dq_g, n1 = (n0 + dq_g < N) ? n0 + dq_g : N;
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t n = n0 + lane; n < n1; n += 32) {
      const float v = scale[n * G + gi];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    float s, z;
    rtn_group_params<float>(mn, mx, dq_bits, dq_sym != 0, false, 1.0f, s, z);
    for (int64_t n = n0 + lane; n < n1; n += 32) {
      float q = rtn_code<float>(scale[n * G + gi], s, z, dq_sym != 0, r);
      if (!dq_sym) q = __fsub_rn(q, z);
      scale[n * G + gi] = __fmul_rn(q, s);
    }
  }
}

__global__ void zero_dead_columns_kernel(float* __restrict__ W, int64_t N, int64_t C, const uint8_t* __restrict__ dead) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N * C; i += (int64_t)gridDim.x * blockDim.x)
    if (dead[i % C]) W[i] = 0.f;
}

// column loop over one sub-block [c0, c0+ncols), ncols <= 128.
// CTA = SUB_WARPS warps x RW rows.  Per-column constants (d = Hinv[i,i], RN(1/d), 0.5/d^2, "needs IEEE division" flag) are
// computed once per CTA into shared memory, so the serial loop body is ~30 instructions per (column, row).
// Err is written TRANSPOSED (ErrT[col_in_block][row]) so the lazy-update GEMM reads it coalesced.
__global__ void __launch_bounds__(32 * SUB_WARPS)
    gptq_subblock_kernel(float* __restrict__ W, const float* __restrict__ Hinv, int64_t N, int64_t C, int64_t c0,
                         int ncols, int g, int64_t G, float maxq, const float* __restrict__ scale,
                         const float* __restrict__ zero, uint8_t* __restrict__ codes, float* __restrict__ Q,
                         float* __restrict__ ErrT, int64_t err_col0, float* __restrict__ losses,
                         float* __restrict__ ErrT_hi, float* __restrict__ ErrT_lo) {
  extern __shared__ __align__(16) float hs[];  // [ncols][SUB] upper-triangular diagonal block of Hinv, then 4 x [SUB] column constants
  float* dcol = hs + SUB * SUB;
  float* rcol = dcol + SUB;
  float* lcol = rcol + SUB;
  uint32_t* fcol = reinterpret_cast<uint32_t*>(lcol + SUB);
  {  // a warp per row of the diagonal block, 4 consecutive columns per lane (one 16-byte load when aligned)
    const int lane_ = threadIdx.x & 31, c = lane_ * 4;
    const bool vec_ok = ((C & 3) == 0) && ((c0 & 3) == 0);
#pragma unroll 4
    for (int r = threadIdx.x >> 5; r < ncols; r += SUB_WARPS) {
      const float* src = Hinv + (c0 + r) * C + c0 + c;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vec_ok && c + 3 < ncols) {
        v = *reinterpret_cast<const float4*>(src);
      } else {
        if (c + 0 < ncols) v.x = src[0];
        if (c + 1 < ncols) v.y = src[1];
        if (c + 2 < ncols) v.z = src[2];
        if (c + 3 < ncols) v.w = src[3];
      }
      if (c + 0 < r) v.x = 0.f;
      if (c + 1 < r) v.y = 0.f;
      if (c + 2 < r) v.z = 0.f;
      if (c + 3 < r) v.w = 0.f;
      *reinterpret_cast<float4*>(hs + r * SUB + c) = v;
    }
  }
  for (int i = threadIdx.x; i < ncols; i += blockDim.x) {
    const float d = Hinv[(c0 + i) * C + c0 + i];
    const RnDivisor D = make_divisor(d);
    dcol[i] = d;
    rcol[i] = D.r;
    lcol[i] = __fdividef(0.5f, d * d);
    fcol[i] = D.slow ? 1u : 0u;
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
  float sc[RW], zr[RW], rs[RW];
  bool sslow[RW];
  const bool per_channel = (g <= 0);
  auto load_group = [&](int64_t gi) {
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const int64_t n = (row0 + r < N) ? row0 + r : N - 1;
      sc[r] = scale[n * G + gi];
      zr[r] = zero[n * G + gi];
      const RnDivisor D = make_divisor(sc[r]);
      rs[r] = D.r;
      sslow[r] = D.slow;
    }
  };
  load_group(per_channel ? 0 : c0 / g);  // the group containing the first column may have started earlier
  int next_group_col = per_channel ? INT_MAX : (int)(((c0 + g - 1) / g) * g - c0);  // first group start >= c0
#pragma unroll
  for (int s = 0; s < 4; ++s) {
#pragma unroll 1
    for (int l = 0; l < 32; ++l) {
      const int i = 32 * s + l;
      if (i >= ncols) break;
      if (i == next_group_col) {  // gptq.py:1264-1272
        load_group((c0 + i) / g);
        next_group_col += g;
      }
      const float d = dcol[i], rd = rcol[i], ld = lcol[i];
      const bool dslow = fcol[i] != 0u;
      float h[4];
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2) h[s2] = hs[i * SUB + lane + 32 * s2];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const float wi = __shfl_sync(0xffffffffu, w[r][s], l);
        // x / scale, correctly rounded (Markstein); IEEE division for the flagged divisors
        float t0;
        if (sslow[r]) {
          t0 = __fdiv_rn(wi, sc[r]);
        } else {
          const float q0 = __fmul_rn(wi, rs[r]);
          t0 = __fmaf_rn(__fmaf_rn(-q0, sc[r], wi), rs[r], q0);
        }
        // Quantizer.quantize (gptq.py:1636-1637)
        const float qi = fminf(fmaxf(__fadd_rn(rintf(t0), zr[r]), 0.f), maxq);
        const float q = __fmul_rn(sc[r], __fsub_rn(qi, zr[r]));
        const float diff = __fsub_rn(wi, q);
        float err;  // (w - q) / d   gptq.py:1296
        if (dslow) {
          err = __fdiv_rn(diff, d);
        } else {
          const float e0 = __fmul_rn(diff, rd);
          err = __fmaf_rn(__fmaf_rn(-e0, d, diff), rd, e0);
        }
        if (lane == l) {
          qv[r][s] = q;
          ev[r][s] = err;
          cd[r] |= ((uint32_t)qi & 0xffu) << (8 * s);
          loss[r] = fmaf(diff * diff, ld, loss[r]);  // gptq.py:1294,1303 (diagnostic, not bit-pinned)
        }
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) {
          const int j = lane + 32 * s2;
          if (j > i) w[r][s2] = __fsub_rn(w[r][s2], __fmul_rn(err, h[s2]));  // gptq.py:1297-1298, no FMA
        }
      }
    }
  }
  // write back
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
      }
    }
    if (losses) {
      const float tot = warp_sum(loss[r]);
      if (lane == 0) losses[n] += tot;
    }
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int c = lane + 32 * s;
    if (c < ncols) {
      float* dst = ErrT + (err_col0 + c) * N + row0;
      if (ErrT_hi) {  // TF32 split for the tensor-core lazy update (gptq_tc.cu): hi is TF32-exact, lo = e - hi
#pragma unroll
        for (int r = 0; r < RW; ++r)
          if (row0 + r < N) {
            const float hi = __uint_as_float(__float_as_uint(ev[r][s]) & 0xffffe000u);
            ErrT_hi[(err_col0 + c) * N + row0 + r] = hi;
            ErrT_lo[(err_col0 + c) * N + row0 + r] = ev[r][s] - hi;
          }
      }
      if (RW == 4 && row0 + 4 <= N && ((N & 3) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(ev[0][s], ev[1 % RW][s], ev[2 % RW][s], ev[3 % RW][s]);
      } else if (RW == 2 && row0 + 2 <= N && ((N & 1) == 0)) {
        *reinterpret_cast<float2*>(dst) = make_float2(ev[0][s], ev[1 % RW][s]);
      } else {
#pragma unroll
        for (int r = 0; r < RW; ++r)
          if (row0 + r < N) dst[r] = ev[r][s];
      }
    }
  }
}

// W[:, j0:j1] -= ErrT[e0:e0+KK, :]^T @ Hinv[r0:r0+KK, j0:j1]     exact fp32 FFMA, 128x128 tile, 8x8 per thread,
// both operands k-major -> float4 global loads, register double buffering of the next 16-deep k chunk.
template <bool VEC>
__global__ void __launch_bounds__(256, 2)
    gptq_lazy_update_kernel(float* __restrict__ W, const float* __restrict__ ErrT, const float* __restrict__ Hinv,
                            int64_t N, int64_t C, int64_t e0, int64_t r0, int KK, int64_t j0, int64_t j1) {
  __shared__ __align__(16) float As[16][128];  // ErrT chunk [k][row]
  __shared__ __align__(16) float Bs[16][128];  // Hinv chunk [k][col]
  const int64_t n0 = (int64_t)blockIdx.y * 128, jb = j0 + (int64_t)blockIdx.x * 128;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // loader mapping: 512 float4 per operand chunk, 2 per thread
  const int lk[2] = {(int)(threadIdx.x >> 5), (int)(threadIdx.x >> 5) + 8};
  const int lc = (threadIdx.x & 31) * 4;
  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = k0 + lk[u];
      ra[u] = rb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < KK) {
        const float* pa = ErrT + (e0 + k) * N + n0 + lc;
        const float* pb = Hinv + (r0 + k) * C + jb + lc;
        if (VEC && n0 + lc + 3 < N) ra[u] = *reinterpret_cast<const float4*>(pa);
        else {
          if (n0 + lc + 0 < N) ra[u].x = pa[0];
          if (n0 + lc + 1 < N) ra[u].y = pa[1];
          if (n0 + lc + 2 < N) ra[u].z = pa[2];
          if (n0 + lc + 3 < N) ra[u].w = pa[3];
        }
        if (VEC && jb + lc + 3 < j1) rb[u] = *reinterpret_cast<const float4*>(pb);
        else {
          if (jb + lc + 0 < j1) rb[u].x = pb[0];
          if (jb + lc + 1 < j1) rb[u].y = pb[1];
          if (jb + lc + 2 < j1) rb[u].z = pb[2];
          if (jb + lc + 3 < j1) rb[u].w = pb[3];
        }
      }
    }
  };
  float acc[8][8] = {};
  gload(0);
  for (int k0 = 0; k0 < KK; k0 += 16) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      *reinterpret_cast<float4*>(&As[lk[u]][lc]) = ra[u];
      *reinterpret_cast<float4*>(&Bs[lk[u]][lc]) = rb[u];
    }
    __syncthreads();
    if (k0 + 16 < KK) gload(k0 + 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
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
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int64_t n = n0 + (u < 4 ? ty * 4 + u : 64 + ty * 4 + (u - 4));
    if (n >= N) continue;
#pragma unroll
    for (int hv = 0; hv < 2; ++hv) {
      const int64_t j = jb + (hv == 0 ? tx * 4 : 64 + tx * 4);
      float* dst = W + n * C + j;
      if (VEC && j + 3 < j1) {
        float4 t = *reinterpret_cast<float4*>(dst);
        t.x -= acc[u][4 * hv + 0];
        t.y -= acc[u][4 * hv + 1];
        t.z -= acc[u][4 * hv + 2];
        t.w -= acc[u][4 * hv + 3];
        *reinterpret_cast<float4*>(dst) = t;
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (j + v < j1) dst[v] -= acc[u][4 * hv + v];
      }
    }
  }
}

}  // namespace b200woq

namespace b200woq {
// gptq_tc.cu: tensor-core (3 x TF32 split) lazy update
struct LazyTcPlan {
  CUtensorMap maps[4];
  bool ok;
};
bool lazy_tc_shape_ok(int64_t N, int64_t C);
int lazy_tc_split(const float* x, float* hi, float* lo, int64_t n, cudaStream_t st);
int lazy_tc_make_plan(LazyTcPlan* plan, const float* e_hi, const float* e_lo, int64_t err_rows, int64_t N, const float* h_hi,
                      const float* h_lo, int64_t C);
int lazy_tc_update(const LazyTcPlan* plan, float* W, int64_t N, int64_t C, int64_t e0, int64_t r0, int KK, int64_t j0,
                   int64_t j1, cudaStream_t st);
}  // namespace b200woq

using namespace b200woq;

// Q[n,c] = scale[n,g] * (code[n,c] - zero[n,g])  -- exactly the value the column loop stores (Quantizer.quantize,
// gptq.py:1636-1637), recomputed from the codes so that multi-GPU row shards only have to exchange u8 codes + params.
__global__ void gptq_rebuild_q_kernel(const uint8_t* __restrict__ codes, const float* __restrict__ scale,
                                      const float* __restrict__ zero, int64_t N, int64_t C, int g, int G,
                                      float* __restrict__ Q) {
  const int64_t total = N * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = idx / C, c = idx - n * C;
    const int64_t gi = n * G + (g > 0 ? c / g : 0);
    Q[idx] = __fmul_rn(scale[gi], __fsub_rn((float)codes[idx], zero[gi]));
  }
}

extern "C" int b200woq_gptq_rebuild_q(const uint8_t* codes, const float* scale, const float* zero, int64_t N, int64_t C,
                                      int groupsize, float* Q, void* stream) {
  WOQ_CHECK_ARG(codes && scale && zero && Q && N > 0 && C > 0, "gptq_rebuild_q: bad arguments");
  const int g = (groupsize <= 0 || groupsize > C) ? 0 : groupsize;
  const int G = g ? (int)ceil_div(C, g) : 1;
  int64_t b = ceil_div(N * C, 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  gptq_rebuild_q_kernel<<<(unsigned)(b > cap ? cap : b), 256, 0, (cudaStream_t)stream>>>(codes, scale, zero, N, C, g, G, Q);
  WOQ_LAUNCH_CHECK();
  return 0;
}

static bool lazy_tc_enabled() {
  static const int v = getenv("B200WOQ_LAZY_TC") ? atoi(getenv("B200WOQ_LAZY_TC")) : 0;
  return v != 0;
}

// [ErrT bs x N] and, for the tensor-core lazy update, [ErrT_hi][ErrT_lo] (bs x N each) + [Hinv_hi][Hinv_lo] (C x C each)
extern "C" int64_t b200woq_gptq_workspace_bytes(int64_t N, int64_t C, int blocksize) {
  const int64_t bs = blocksize <= 0 ? C : (blocksize > C ? C : blocksize);
  int64_t floats = N * bs;
  if (lazy_tc_enabled() && bs < C && lazy_tc_shape_ok(N, C)) floats += 2 * N * bs + 2 * C * C;
  return floats * (int64_t)sizeof(float) + 256;
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
  float* ErrT = (float*)workspace;  // [bs, N]  (transposed: the lazy GEMM reads it k-major)
  const bool use_tc = lazy_tc_enabled() && bs < C && lazy_tc_shape_ok(N, C);
  float *ErrT_hi = nullptr, *ErrT_lo = nullptr;
  LazyTcPlan plan;
  plan.ok = false;
  if (use_tc) {
    ErrT_hi = ErrT + N * bs;
    ErrT_lo = ErrT_hi + N * bs;
    float* H_hi = ErrT_lo + N * bs;
    float* H_lo = H_hi + C * C;
    if (int rc = lazy_tc_split(Hinv, H_hi, H_lo, C * C, st)) return rc;
    if (lazy_tc_make_plan(&plan, ErrT_hi, ErrT_lo, bs, N, H_hi, H_lo, C) != 0) {
      plan.ok = false;
      ErrT_hi = ErrT_lo = nullptr;
    }
  }
  const float maxq = (float)((1 << bits) - 1);
  const bool per_channel = groupsize <= 0;
  const int g = per_channel ? (int)C : groupsize;
  const int64_t G = per_channel ? 1 : ceil_div(C, g);
  const int mse = flags & 1;
  // flags bit 1: double quantisation of the scales; bit 2: symmetric; bits 8-15: bits; bits 16-31: group size
  const int dq = (flags >> 1) & 1, dq_sym = (flags >> 2) & 1, dq_bits = (flags >> 8) & 0xff, dq_g = (flags >> 16) & 0xffff;
  WOQ_CHECK_ARG(!dq || (dq_bits >= 1 && dq_bits <= 8 && dq_g > 0), "gptq_fasterquant: bad double-quant parameters");
  if (losses) WOQ_CUDA(cudaMemsetAsync(losses, 0, sizeof(float) * N, st));
  const int fp_blocks = (int)std::min<int64_t>(ceil_div(N * 32, 256) * 4, (int64_t)num_sms() * 8);

  if (per_channel) {  // gptq.py:1184-1185: one find_params over the whole row, BEFORE dead columns are zeroed
    gptq_find_params_kernel<<<fp_blocks, 256, 0, st>>>(W, N, C, 0, (int)C, 0, 1, 1, maxq, sym, mse, scale, zero);
    WOQ_LAUNCH_CHECK();
    if (dq) {
      gptq_double_quant_kernel<<<fp_blocks, 256, 0, st>>>(scale, N, 1, 0, 1, dq_bits, dq_g, dq_sym);
      WOQ_LAUNCH_CHECK();
    }
  }
  if (dead_mask) {  // gptq.py:1191  W[:, dead] = 0
    zero_dead_columns_kernel<<<fp_blocks, 256, 0, st>>>(W, N, C, dead_mask);
    WOQ_LAUNCH_CHECK();
  }
  const int rows_per_cta = SUB_WARPS * RW;
  const bool vec = ((N & 3) == 0) && ((C & 3) == 0);
  for (int64_t i1 = 0; i1 < C; i1 += bs) {
    const int64_t i2 = std::min(i1 + bs, C);
    if (!per_channel) {
      // groups whose first column lies in [i1, i2)
      const int64_t gfirst = ceil_div(i1, g), glast = (i2 - 1) / g;
      if (gfirst <= glast) {
        gptq_find_params_kernel<<<fp_blocks, 256, 0, st>>>(W, N, C, gfirst * g, g, gfirst, (int)(glast - gfirst + 1), G,
                                                           maxq, sym, mse, scale, zero);
        WOQ_LAUNCH_CHECK();
        if (dq) {
          gptq_double_quant_kernel<<<fp_blocks, 256, 0, st>>>(scale, N, G, gfirst, (int)(glast - gfirst + 1), dq_bits, dq_g,
                                                              dq_sym);
          WOQ_LAUNCH_CHECK();
        }
      }
    }
    for (int64_t c0 = i1; c0 < i2; c0 += SUB) {
      const int ncols = (int)std::min<int64_t>(SUB, i2 - c0);
      const size_t smem = ((size_t)SUB * SUB + 4 * SUB) * sizeof(float);
      // per-device attribute: set on every call (cheap), never cached per process
      WOQ_CUDA(cudaFuncSetAttribute(gptq_subblock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      gptq_subblock_kernel<<<(unsigned)ceil_div(N, rows_per_cta), 32 * SUB_WARPS, smem, st>>>(
          W, Hinv, N, C, c0, ncols, per_channel ? 0 : g, G, maxq, scale, zero, codes, Q, ErrT, c0 - i1, losses, ErrT_hi,
          ErrT_lo);
      WOQ_LAUNCH_CHECK();
      const int64_t c1 = c0 + ncols;
      if (c1 < i2) {  // in-block propagation to the rest of the block
        dim3 grid((unsigned)ceil_div(i2 - c1, 128), (unsigned)ceil_div(N, 128));
        if (plan.ok && (ncols % 8) == 0 && (c1 & 3) == 0) {
          if (int rc = lazy_tc_update(&plan, W, N, C, c0 - i1, c0, ncols, c1, i2, st)) return rc;
        } else if (vec && (c1 & 3) == 0)
          gptq_lazy_update_kernel<true><<<grid, 256, 0, st>>>(W, ErrT, Hinv, N, C, c0 - i1, c0, ncols, c1, i2);
        else
          gptq_lazy_update_kernel<false><<<grid, 256, 0, st>>>(W, ErrT, Hinv, N, C, c0 - i1, c0, ncols, c1, i2);
        WOQ_LAUNCH_CHECK();
      }
    }
    if (i2 < C) {  // gptq.py:1304
      dim3 grid((unsigned)ceil_div(C - i2, 128), (unsigned)ceil_div(N, 128));
      if (plan.ok && ((i2 - i1) % 8) == 0 && (i2 & 3) == 0) {
        if (int rc = lazy_tc_update(&plan, W, N, C, 0, i1, (int)(i2 - i1), i2, C, st)) return rc;
      } else if (vec && (i2 & 3) == 0)
        gptq_lazy_update_kernel<true><<<grid, 256, 0, st>>>(W, ErrT, Hinv, N, C, 0, i1, (int)(i2 - i1), i2, C);
      else
        gptq_lazy_update_kernel<false><<<grid, 256, 0, st>>>(W, ErrT, Hinv, N, C, 0, i1, (int)(i2 - i1), i2, C);
      WOQ_LAUNCH_CHECK();
    }
  }
  return 0;
}
