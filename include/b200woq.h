/*
 * b200woq.h -- C ABI of libb200woq.so: the B200 (sm_100a) weight-only-quantisation hot path.
 *
 * This is the drop-in boundary for intel/neural-compressor's torch weight-only path.  The reference
 * has no FFI (it is 100 % Python, SURVEY.md "Headline facts"); every entry point below replaces one
 * Python/torch-eager function of the reference, cited as file:line relative to
 *   neural_compressor/torch/algorithms/weight_only/      (unless another path is given).
 * The binding a maintainer of the reference would add (ctypes, no torch types in any signature) is
 * shown in INTEGRATION.md and implemented in neural_compressor_b200/_lib.py.
 *
 * Conventions
 *   - plain pointers are DEVICE pointers unless a parameter is named host_*; the caller owns every
 *     buffer; the library never allocates user-visible memory (workspace sizes are queried);
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no implicit sync;
 *   - return 0 on success, negative B200WOQ_E* otherwise; b200woq_last_error() gives the text
 *     (thread-local);
 *   - matrices are row-major; a weight is [N = out_features, K = in_features] like nn.Linear.weight;
 *   - packed "optimum" layout (modules.py:236-262): qweight int32 [ceil(K/n_pack), N],
 *     qzeros int32 [ceil(K/g), ceil(N/n_pack)], scales fp16 [ceil(K/g), N], n_pack = 32 / bits;
 *     field e of a word is ((v & mask) << bits*e) (modules.py:533-543, bit_packer.py:35-278);
 *     stored zero-points are zp-1 (modules.py:363-364), symmetric codes are stored +2^(bits-1)
 *     (modules.py:329-334).
 */
#ifndef B200WOQ_H_
#define B200WOQ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200WOQ_VERSION 100

enum {
  B200WOQ_OK = 0,
  B200WOQ_EINVAL = -1,      /* bad argument (shape, dtype, alignment, unsupported combination) */
  B200WOQ_ECUDA = -2,       /* a CUDA runtime call or launch failed */
  B200WOQ_EWORKSPACE = -3,  /* workspace too small */
  B200WOQ_EUNSUPPORTED = -4 /* valid request that this build has no kernel for */
};

/* element types of floating-point inputs/outputs */
enum { B200WOQ_F32 = 0, B200WOQ_F16 = 1, B200WOQ_BF16 = 2 };

int b200woq_version(void);
const char* b200woq_last_error(void);
/* fills `out` with "sm_XY" of the current device; B200WOQ_ECUDA if there is none */
int b200woq_device_arch(char* out, int out_len);
/* number of kernels this library has launched in this process (bench.py's `gpu_launches` evidence) */
int64_t b200woq_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * K4  RTN group quantisation and bit packing
 * ---------------------------------------------------------------------------------------------- */

/* quant_tensor()'s per-group parameters (utility.py:162-244 qdq_weight_asym/_sym; :272-376 grouping
 * incl. the ragged tail group).  W [N,K] of w_dtype.  G = ceil(K/g), g = K when group_size <= 0 or
 * group_size > K.  scale[N*G], zp[N*G] are fp32 holding the exact values the reference computes in
 * W's dtype (zp is written only when sym == 0; may be NULL otherwise). */
int b200woq_rtn_params(const void* W, int w_dtype, int64_t N, int64_t K, int bits, int group_size,
                       int sym, int full_range, float quantile, float* scale, float* zp, void* stream);

/* quant_tensor(return_int=True) codes + INCWeightOnlyLinear.pack (modules.py:321-375) in one pass:
 * code = clamp(round(w/scale) (+zp), lo, hi) in W's dtype arithmetic, stored as the unsigned field
 * the reference packs (sym: +2^(bits-1)).  Writes qweight; optionally the raw stored codes
 * (uint8 [N,K], may be NULL).  Use b200woq_pack_params for scales/qzeros. */
int b200woq_rtn_quant_pack(const void* W, int w_dtype, int64_t N, int64_t K, int bits, int group_size,
                           int sym, const float* scale, const float* zp, int32_t* qweight,
                           uint8_t* codes_out, void* stream);

/* quant_tensor(return_int=False): in-place-style fake quantisation, out = scale*(q - zp)
 * (utility.py:195-198, 241-244).  `col_scale` (fp32 [K], may be NULL) implements AWQ's
 * qdq(W*s)/s (awq.py:326-335): the group statistics and rounding see W*s, the result is divided by s.
 * `out` has w_dtype and may alias W. */
int b200woq_rtn_fake_quant(const void* W, int w_dtype, int64_t N, int64_t K, int bits, int group_size,
                           int sym, int full_range, float quantile, const float* col_scale, void* out,
                           void* stream);

/* pack already-computed stored codes (uint8 [N,K], values in [0, 2^bits)) -> qweight
 * (modules.py:445-466 / bit_packer.py). Used by the GPTQ export (gptq.py:813-847). */
int b200woq_pack_codes(const uint8_t* codes, int64_t N, int64_t K, int bits, int32_t* qweight,
                       void* stream);

/* scales fp32 [N,G] -> fp16 [G,N]; zp fp32 [N,G] (or NULL => 2^(bits-1)) -> qzeros int32
 * [G, ceil(N/n_pack)] holding (zp-1)&mask (modules.py:345-371). */
int b200woq_pack_params(const float* scale, const float* zp, int64_t N, int64_t G, int bits,
                        void* scales16_out, int32_t* qzeros_out, void* stream);

/* INCWeightOnlyLinear.unpack (modules.py:377-411): stored codes uint8 [N,K] and zero-points
 * uint8 [N,G] (already +1 and wrapped). Either output may be NULL. */
int b200woq_unpack(const int32_t* qweight, const int32_t* qzeros, int64_t N, int64_t K, int64_t G,
                   int bits, uint8_t* codes_out, uint8_t* zp_out, void* stream);

/* INCWeightOnlyLinear.recover (modules.py:413-443): W_fp16[n,k] = fp16(int8(q - zp) * scale_fp16).
 * g_idx int32 [K] or NULL (=> k / group_size). */
int b200woq_dequantize(const int32_t* qweight, const int32_t* qzeros, const void* scales16,
                       const int32_t* g_idx, int64_t N, int64_t K, int bits, int group_size,
                       void* w_fp16_out, void* stream);

/* ---- 4-bit TABLE data types: nf4, fp4 (= fp4_e2m1_bnb), fp4_e2m1 (utility.py:52-103) ------------------------- */

/* One data type = its ascending levels (FLOAT_MAPPING), the integer code the reference stores for each level
 * (INT_MAPPING), the mid points between neighbours ((level[i] + level[i+1]) / 2 evaluated in double, then converted to
 * float like torch converts a Python scalar) and max(levels).  HOST struct, passed by pointer, read at call time. */
typedef struct {
  int32_t n;          /* number of levels: 16 (nf4) or 15 (fp4 variants) */
  float level[16];
  float mid[16];      /* n - 1 entries */
  int32_t code[16];   /* in [-8, 7] */
  float max_level;
} b200woq_f4_table;

/* quantize_4bit (utility.py:121-160) under quant_tensor's grouping (utility.py:272-376, ragged tail group included):
 * scale = absmax(group) * quantile / max_level, element -> nearest level by the mid-point intervals, every op rounded
 * through W's dtype like torch does.  Outputs (each may be NULL, at least one is required):
 *   codes_out int8 [N,K]   the integer codes quant_tensor(return_int=True) leaves in the tensor
 *   scale_out fp32 [N,G]   the scales (exactly the W-dtype values)
 *   fake_out  w_dtype [N,K] level * scale, i.e. quant_tensor(return_int=False); may alias W. */
int b200woq_f4_quantize(const void* W, int w_dtype, int64_t N, int64_t K, int group_size,
                        const b200woq_f4_table* host_table, float quantile, int8_t* codes_out, float* scale_out,
                        void* fake_out, void* stream);

/* INCWeightOnlyLinear.pack for the non-optimum layout with compression_dim = 1, int32 (modules.py:352-357, 445-466):
 * qweight[n, j] = OR_e (codes[n, j*n_pack + e] & mask) << bits*e, qweight int32 [N, ceil(K/n_pack)]. */
int b200woq_pack_rows(const int8_t* codes, int64_t N, int64_t K, int bits, int32_t* qweight_out, void* stream);

/* unpack + recover for a table data type (modules.py:377-443): nibble -> host_nibble_levels[nibble] (16 floats indexed
 * by the stored 4-bit field, i.e. the level of the sign-extended code, 0 for unused codes) times scales fp32 [N,G];
 * w_out fp32 [N,K]. */
int b200woq_f4_dequantize(const int32_t* qweight, const float* scales, const float* host_nibble_levels, int64_t N,
                          int64_t K, int group_size, float* w_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K6  WeightOnlyLinear.forward: fused unpack + dequant + matmul (modules.py:594-610)
 * ---------------------------------------------------------------------------------------------- */

/* bytes of scratch the call below needs for (M,N,K); 0 is possible */
int64_t b200woq_linear_workspace_bytes(int64_t M, int64_t N, int64_t K, int bits, int group_size);

/* y[M,N] = (x * input_scale)[M,K] . W^T + bias.   x_dtype / y_dtype / bias_dtype in B200WOQ_F*.
 * input_scale fp32 [K] or NULL (MulLinear, modules.py:907-949).  bias [N] or NULL.
 * workspace must be zero-initialised once by the caller (the kernels leave it zeroed).
 * flags: bit0 = force the general (CUDA-core) kernel, bit1 = programmatic dependent launch. */
int b200woq_linear_forward(const void* x, int x_dtype, int64_t M, int64_t K, int64_t N,
                           const int32_t* qweight, const int32_t* qzeros, const void* scales16,
                           const int32_t* g_idx, const void* bias, int bias_dtype,
                           const float* input_scale, void* y, int y_dtype, int bits, int group_size,
                           void* workspace, int64_t workspace_bytes, int flags, void* stream);

/* Decode-batch (M <= 4) 4-bit path on a derived, B200-native "stream layout" (woq_stream.cu): one contiguous
 * 2144-byte record per (32-column strip, group) = lane-ordered packed words + 32 fp16 scales + 32 decoded
 * zero-points, strips outermost; every warp streams its own records through a ring of cp.async.bulk copies.
 * Because strips are outermost, the layouts of several linears that read the same activation (q/k/v, gate/up)
 * concatenate into the layout of the fused [sum N, K] linear: one launch serves all of them.
 * Built once per module from the optimum-format tensors (the reference likewise caches a derived weight at first
 * forward, modules.py:603-604).  b200woq_stream_layout_bytes returns 0 when the shape is not eligible (bits != 4,
 * group % 32, K % group, N % 32).  flags bit 1 = programmatic dependent launch: the kernel prefetches the layout
 * before griddepcontrol.wait, so the flag is only legal when no in-flight kernel is still WRITING the layout (the
 * Python module clears it for the first launch after a build). */
int64_t b200woq_stream_layout_bytes(int64_t N, int64_t K, int bits, int group_size);
int b200woq_build_stream_layout(const int32_t* qweight, const int32_t* qzeros, const void* scales16, int64_t N,
                                int64_t K, int bits, int group_size, void* out, void* stream);
int b200woq_linear_forward_stream(const void* x, int x_dtype, int64_t M, int64_t K, int64_t N,
                                  const void* stream_layout, const void* bias, int bias_dtype,
                                  const float* input_scale, void* y, int y_dtype, int bits, int group_size,
                                  int flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K1-K3  GPTQ calibration (gptq.py:1089-1351)
 * ---------------------------------------------------------------------------------------------- */

/* GPTQ.add_batch (gptq.py:1111-1141) without the running-mean rescale: Hsum[C,C] (fp32, upper
 * triangle in 128x128 tiles incl. the diagonal tiles in full) += X^T X for X [T, C] of x_dtype with
 * row stride ldx elements.  fp16/bf16 inputs use one tensor-core pass (products exact in fp32);
 * fp32 inputs use the 3xTF32 split.  Call b200woq_hessian_finalize once after the last batch. */
int b200woq_hessian_accumulate(const void* X, int x_dtype, int64_t T, int64_t C, int64_t ldx,
                               float* Hsum, void* stream);

/* H = (2/nsamples) * Hsum mirrored to full symmetric (gptq.py:1136-1141 closed form); then the
 * fasterquant prologue (gptq.py:1189-1191, 1221-1227): dead = diag(H)==0 -> H[dead,dead]=1 and
 * dead_mask[c]=1 (uint8 [C]); damp = percdamp*mean(diag H) added to the diagonal.
 * `scratch` fp32 [2] device. */
int b200woq_hessian_finalize(float* H, int64_t C, double nsamples, float percdamp, uint8_t* dead_mask,
                             float* scratch, void* stream);

/* K2 -- the inverse Cholesky factor of the fasterquant prologue (gptq.py:1228-1231:
 *   H = cholesky(H); H = cholesky_inverse(H); H = cholesky(H, upper=True)): U fp32 [C,C] upper triangular with
 * U^T U = H^-1 (zeros below the diagonal), from ONE blocked Cholesky of the index-reversed matrix plus ONE blocked
 * triangular inverse (U = J inv(chol(J H J)) J), exact fp32 FFMA tiles (cholinv.cu).  H [C,C] is the full symmetric
 * damped Hessian (output of b200woq_hessian_finalize) and is left untouched.  `workspace` (16-byte aligned,
 * b200woq_cholinv_workspace_bytes(C) bytes) holds two padded C x C buffers.  `info` (device int) receives 0 on
 * success or 1 + the first (index-reversed) column whose pivot is not positive -- where the reference's
 * torch.linalg.cholesky raises; U then holds NaNs.  Stream-ordered, no host synchronisation. */
int64_t b200woq_cholinv_workspace_bytes(int64_t C);
int b200woq_cholinv_upper(const float* H, int64_t C, float* U, void* workspace, int64_t workspace_bytes, int* info,
                          void* stream);

/* b200woq_hessian_finalize followed by b200woq_cholinv_upper on the finalized matrix (SURVEY §8b
 * `hessian_finalize_cholinv_upper`): Hsum is finalized in place, U receives the factor. */
int b200woq_hessian_finalize_cholinv_upper(float* Hsum, int64_t C, double nsamples, float percdamp,
                                           uint8_t* dead_mask, float* scratch, float* U, void* workspace,
                                           int64_t workspace_bytes, int* info, void* stream);

/* Workspace of b200woq_gptq_fasterquant: the transposed error block Err^T fp32 [blocksize, N].  When the process
 * runs with B200WOQ_LAZY_TC=1 (opt-in tensor-core lazy update, gptq_tc.cu) it also holds the TF32 hi/lo splits of
 * Err^T and of Hinv: + 2*blocksize*N + 2*C*C floats. */
int64_t b200woq_gptq_workspace_bytes(int64_t N, int64_t C, int blocksize);

/* GPTQ.fasterquant column loop (gptq.py:1250-1304) + Quantizer.find_params/quantize (:1501-1637,
 * dtype int, perchannel) + the export's code extraction (utility.py:483-537), for one layer.
 *   W      fp32 [N,C]  in: weights (permuted if act_order) ; out: destroyed
 *   dead_mask uint8 [C] or NULL: columns zeroed (gptq.py:1191) after the per-channel find_params
 *   Hinv   fp32 [C,C]  upper Cholesky factor of H^-1 (gptq.py:1228-1231)
 *   codes  uint8 [N,C] stored codes (q, in [0,2^bits)) ; Q fp32 [N,C] fake-quant (may be NULL)
 *   scale/zero fp32 [N,G], G = ceil(C/groupsize) (groupsize<=0 => 1 group = per-channel)
 *   losses fp32 [N] per-row sum of (w-q)^2/d^2/2 (gptq.py:1294,1303,1318) (may be NULL)
 * blocksize must be a multiple of groupsize (or groupsize<=0) -- the find_params "stale view"
 * semantics of blocksize > groupsize are reproduced.  flags bit0: mse search; bit1: double
 * quantisation of each group's scales over the output rows (gptq.py:1598-1614) with bit2 = symmetric, bits 8-15 =
 * double_quant_bits, bits 16-31 = double_quant_group_size.
 * The lazy update W[:, i2:] -= Err1 @ Hinv[i1:i2, i2:] (gptq.py:1304) runs as exact fp32 FFMA tiles by default. */
int b200woq_gptq_fasterquant(float* W, const float* Hinv, const uint8_t* dead_mask, int64_t N, int64_t C,
                             int blocksize, int groupsize, int bits, int sym, int flags, uint8_t* codes, float* Q,
                             float* scale, float* zero, float* losses, void* workspace,
                             int64_t workspace_bytes, void* stream);

/* Fake-quant weights from codes: Q[n,c] = scale[n,g(c)] * (codes[n,c] - zero[n,g(c)]), bit-identical to the Q the
 * column loop emits (Quantizer.quantize, gptq.py:1636-1637).  Lets row-sharded ranks exchange u8 codes + params only. */
int b200woq_gptq_rebuild_q(const uint8_t* codes, const float* scale, const float* zero, int64_t N, int64_t C,
                           int groupsize, float* Q, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5  AWQ statistics (awq.py:131-154) and search losses (awq.py:343-344, 452-453)
 * ---------------------------------------------------------------------------------------------- */

/* _get_weight_scale (awq.py:131-147): out[k] = mean_n |W[n,k]| / max_{group(n,k)} |W[n,.]|.  fp32 out [K] */
int b200woq_awq_weight_scale(const void* W, int w_dtype, int64_t N, int64_t K, int group_size,
                             float* out, void* stream);

/* _get_act_scale (awq.py:151-154) accumulation: sum_abs[k] += sum_t |X[t,k]| (fp32 [K]); the caller
 * divides by the token count. */
int b200woq_abs_colsum_accumulate(const void* X, int x_dtype, int64_t T, int64_t K, int64_t ldx,
                                  float* sum_abs, void* stream);

/* loss accumulation: acc[0] (double, like the Python float the reference sums into) += float mean((a-b)^2) over `count`
 * elements, i.e. `(o1 - o2).float().pow(2).mean().item()` (awq.py:343-344). */
int b200woq_mse_accumulate(const void* a, const void* b, int dtype, int64_t count, double* acc,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7  SmoothQuant calibration statistics (smooth_quant/utility.py:858-883)
 * ---------------------------------------------------------------------------------------------- */

/* per-input-channel running max/min: mx[k] = max(mx[k], max_t X[t,k]), mn likewise. fp32 [K]. */
int b200woq_minmax_cols_accumulate(const void* X, int x_dtype, int64_t T, int64_t K, int64_t ldx,
                                   float* mx, float* mn, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7  SmoothQuant W8A8 static INT8 linear (smooth_quant/utility.py: SQLinearWrapper :2559-2662,
 *     quant_dequant_w_v1 :652-690, quant_dequant_x_v1 :726-755; the INT8 GEMM the reference delegates to IPEX)
 * ---------------------------------------------------------------------------------------------- */

/* Padded K of the int8 operands (multiple of 128: one 128-byte swizzle row per pipeline stage). */
int64_t b200woq_w8a8_padded_k(int64_t K);

/* `sq_smooth` + weight quantisation, once per layer: W' = W * smooth[k] (smooth may be NULL), per-out-channel
 * symmetric int8: w_scale[n] = max(max_k |W'[n,k]| / 127.5, eps), qweight[n,k] = clamp(rint(W'/w_scale), -128, 127)
 * written as int8 [N, padded_k(K)] (zero padded), wsum[n] = sum_k qweight[n,k] (int32). */
int b200woq_sq_smooth_quant_weight(const void* W, int w_dtype, int64_t N, int64_t K, const float* smooth,
                                   int8_t* qweight, float* w_scale, int32_t* wsum, void* stream);

/* Workspace of b200woq_w8a8_linear_forward: [u8 activation codes M x padded_k | split-K s32 sums + tile counters].
 * The part from b200woq_w8a8_workspace_zeroed_offset(M, K) on must be ZERO when first used; every call leaves it
 * zeroed.  256-byte aligned. */
int64_t b200woq_w8a8_workspace_bytes(int64_t M, int64_t N, int64_t K);
int64_t b200woq_w8a8_workspace_zeroed_offset(int64_t M, int64_t K);

/* y[M,N] = (sum_k q_x[m,k] q_w[n,k] - zp_x * wsum[n]) * x_scale * w_scale[n] + bias[n]
 * with q_x = clamp(rint(x * input_scale[k] / x_scale + x_zp), 0, 255) (static per-tensor asym uint8; input_scale
 * NULL = ones).  x_scale / x_zp are device scalars (fp32; x_zp integer-valued).  tcgen05.mma.kind::i8. */
int b200woq_w8a8_linear_forward(const void* x, int x_dtype, int64_t M, int64_t K, int64_t N, const int8_t* qweight,
                                const float* w_scale, const int32_t* wsum, const float* input_scale,
                                const float* x_scale, const float* x_zp, const void* bias, int bias_dtype, void* y,
                                int y_dtype, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200WOQ_H_ */
