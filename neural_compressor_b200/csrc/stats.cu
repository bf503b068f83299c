// stats.cu -- K5/K7 statistics: AWQ weight/activation scales, search losses, SmoothQuant min/max.
//
// Reference: awq.py:131-147 _get_weight_scale, :151-154 _get_act_scale, :343-344 / :452-453 loss;
//            smooth_quant/utility.py:858-883 per-input-channel min/max hook.
// All are column reductions over a row-major [rows, K] matrix: threads sweep K (coalesced), row chunks are
// reduced in two deterministic stages (fixed order -> run-to-run reproducible, no float atomics).
#include "common.cuh"

namespace b200woq {

constexpr int kColsPerBlock = 128;
constexpr int kRowChunk = 256;

// OP: 0 = sum |x| ; 1 = max ; 2 = min ; 3 = sum |x| / gmax[row][k/g]
template <typename T, int OP>
__global__ void __launch_bounds__(kColsPerBlock)
    col_reduce_partial_kernel(const T* __restrict__ X, int64_t rows, int64_t K, int64_t ldx, const float* __restrict__ gmax,
                              int g, int64_t G, float* __restrict__ partial) {
  const int64_t k = (int64_t)blockIdx.x * kColsPerBlock + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * kRowChunk;
  const int64_t r1 = (r0 + kRowChunk < rows) ? r0 + kRowChunk : rows;
  if (k >= K) return;
  float acc = (OP == 1) ? -INFINITY : (OP == 2) ? INFINITY : 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const float v = ElemTraits<T>::load(X + r * ldx + k);
    if (OP == 0) acc += fabsf(v);
    if (OP == 1) acc = fmaxf(acc, v);
    if (OP == 2) acc = fminf(acc, v);
    if (OP == 3) acc += __fdiv_rn(fabsf(v), gmax[r * G + k / g]);
  }
  partial[(int64_t)blockIdx.y * K + k] = acc;
}

// MODE: 0 = out += sum(partials) ; 1 = out = max(out, partials) ; 2 = out = min(out, ...) ; 3 = out = sum / rows
template <int MODE>
__global__ void col_reduce_final_kernel(const float* __restrict__ partial, int64_t chunks, int64_t K, float inv_rows,
                                        float* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float acc = (MODE == 1) ? -INFINITY : (MODE == 2) ? INFINITY : 0.f;
  for (int64_t c = 0; c < chunks; ++c) {
    const float v = partial[c * K + k];
    if (MODE == 1) acc = fmaxf(acc, v);
    else if (MODE == 2) acc = fminf(acc, v);
    else acc += v;
  }
  if (MODE == 0) out[k] += acc;
  if (MODE == 1) out[k] = fmaxf(out[k], acc);
  if (MODE == 2) out[k] = fminf(out[k], acc);
  if (MODE == 3) out[k] = acc * inv_rows;
}

// per-(row, group) max |w|   (awq.py:143-144)
template <typename T>
__global__ void __launch_bounds__(256) group_absmax_kernel(const T* __restrict__ W, int64_t N, int64_t K, int g, int64_t G,
                                                          float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t task = warp; task < N * G; task += nwarps) {
    const int64_t n = task / G, gi = task % G;
    const int64_t k0 = gi * g, k1 = (k0 + g < K) ? k0 + g : K;
    float m = 0.f;
    for (int64_t k = k0 + lane; k < k1; k += 32) m = fmaxf(m, fabsf(ElemTraits<T>::load(W + n * K + k)));
    m = warp_max(m);
    if (lane == 0) out[task] = m;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) sqdiff_partial_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t count,
                                                            float* __restrict__ partial) {
  __shared__ float red[256];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    // (out1 - out2) is evaluated in the tensors' dtype, then .float() (awq.py:343-344)
    const float d = ElemTraits<T>::round(__fsub_rn(ElemTraits<T>::load(a + i), ElemTraits<T>::load(b + i)));
    s = fmaf(d, d, s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void sqdiff_final_kernel(const float* __restrict__ partial, int n, float inv_count, double* __restrict__ acc) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += (double)partial[i];
    acc[0] += (double)(float)(s * (double)inv_count);
  }
}

template <typename T, int OP, int MODE>
static int col_reduce(const void* X, int64_t rows, int64_t K, int64_t ldx, const float* gmax, int g, int64_t G, float* out,
                      cudaStream_t st) {
  const int64_t chunks = ceil_div(rows, kRowChunk);
  float* partial = nullptr;
  WOQ_CUDA(cudaMallocAsync((void**)&partial, sizeof(float) * chunks * K, st));
  dim3 grid((unsigned)ceil_div(K, kColsPerBlock), (unsigned)chunks);
  col_reduce_partial_kernel<T, OP><<<grid, kColsPerBlock, 0, st>>>((const T*)X, rows, K, ldx, gmax, g, G, partial);
  col_reduce_final_kernel<MODE><<<(unsigned)ceil_div(K, 256), 256, 0, st>>>(partial, chunks, K, 1.f / (float)rows, out);
  count_launch(2);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(partial, st);
  if (e != cudaSuccess) {
    set_error("column reduction launch failed: %s", cudaGetErrorString(e));
    return B200WOQ_ECUDA;
  }
  return 0;
}

}  // namespace b200woq

using namespace b200woq;

#define STATS_DISPATCH(dt, ...)                                   \
  switch (dt) {                                                   \
    case B200WOQ_F32: { using T = float; __VA_ARGS__; } break;    \
    case B200WOQ_F16: { using T = __half; __VA_ARGS__; } break;   \
    case B200WOQ_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break; \
    default: set_error("unsupported dtype %d", dt); return B200WOQ_EINVAL; \
  }

extern "C" int b200woq_awq_weight_scale(const void* W, int w_dtype, int64_t N, int64_t K, int group_size, float* out,
                                        void* stream) {
  WOQ_CHECK_ARG(W && out && N > 0 && K > 0, "awq_weight_scale: bad arguments");
  // awq.py:141-142: weight.view(-1, q_group_size) only when q_group_size > 0, else one group per ROW
  const int g = (group_size > 0) ? group_size : (int)K;
  WOQ_CHECK_ARG(K % g == 0, "awq_weight_scale: K must be a multiple of the group size (weight.view(-1, g))");
  const int64_t G = K / g;
  cudaStream_t st = (cudaStream_t)stream;
  float* gmax = nullptr;
  WOQ_CUDA(cudaMallocAsync((void**)&gmax, sizeof(float) * N * G, st));
  int64_t blocks = ceil_div(N * G * 32, 256);
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  int rc = 0;
  STATS_DISPATCH(w_dtype, {
    group_absmax_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const T*)W, N, K, g, G, gmax);
    count_launch(1);
    rc = col_reduce<T, 3, 3>(W, N, K, K, gmax, g, G, out, st);
  });
  cudaFreeAsync(gmax, st);
  return rc;
}

extern "C" int b200woq_abs_colsum_accumulate(const void* X, int x_dtype, int64_t T_, int64_t K, int64_t ldx, float* sum_abs,
                                             void* stream) {
  WOQ_CHECK_ARG(X && sum_abs && T_ > 0 && K > 0 && ldx >= K, "abs_colsum_accumulate: bad arguments");
  int rc = 0;
  STATS_DISPATCH(x_dtype, rc = (col_reduce<T, 0, 0>(X, T_, K, ldx, nullptr, 1, 1, sum_abs, (cudaStream_t)stream)));
  return rc;
}

extern "C" int b200woq_minmax_cols_accumulate(const void* X, int x_dtype, int64_t T_, int64_t K, int64_t ldx, float* mx,
                                              float* mn, void* stream) {
  WOQ_CHECK_ARG(X && mx && mn && T_ > 0 && K > 0 && ldx >= K, "minmax_cols_accumulate: bad arguments");
  int rc = 0;
  STATS_DISPATCH(x_dtype, {
    rc = col_reduce<T, 1, 1>(X, T_, K, ldx, nullptr, 1, 1, mx, (cudaStream_t)stream);
    if (rc == 0) rc = col_reduce<T, 2, 2>(X, T_, K, ldx, nullptr, 1, 1, mn, (cudaStream_t)stream);
  });
  return rc;
}

extern "C" int b200woq_mse_accumulate(const void* a, const void* b, int dtype, int64_t count, double* acc, void* stream) {
  WOQ_CHECK_ARG(a && b && acc && count > 0, "mse_accumulate: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = ceil_div(count, 256 * 8);
  if (blocks > 1024) blocks = 1024;
  float* partial = nullptr;
  WOQ_CUDA(cudaMallocAsync((void**)&partial, sizeof(float) * blocks, st));
  STATS_DISPATCH(dtype, (sqdiff_partial_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const T*)a, (const T*)b, count, partial)));
  sqdiff_final_kernel<<<1, 32, 0, st>>>(partial, (int)blocks, (float)(1.0 / (double)count), acc);
  count_launch(2);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(partial, st);
  if (e != cudaSuccess) {
    set_error("mse launch failed: %s", cudaGetErrorString(e));
    return B200WOQ_ECUDA;
  }
  return 0;
}
