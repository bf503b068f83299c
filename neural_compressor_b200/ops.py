"""Tensor-level wrappers over the C ABI (include/b200woq.h).

PyTorch is plumbing here: it owns device memory and the current stream; every computation below is a
hand-written sm_100a kernel inside libb200woq.so.  All inputs must be contiguous CUDA tensors -- there
is no CPU path and no fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from ._lib import check, dt, ptr, require_cuda, stream_ptr


def _eff_group(K: int, group_size: int) -> int:
    return K if (group_size <= 0 or group_size > K) else group_size


def n_pack(bits: int) -> int:
    return 32 // bits


# ------------------------------------------------------------------ K4: RTN + pack
def rtn_params(W: torch.Tensor, bits=4, group_size=-1, sym=False, full_range=False, quantile=1.0):
    """quant_tensor()'s per-group (scale, zp) -- fp32 [N, G] (utility.py:162-244)."""
    require_cuda(W, "W")
    N, K = W.shape
    G = math.ceil(K / _eff_group(K, group_size))
    scale = torch.empty((N, G), dtype=torch.float32, device=W.device)
    zp = None if sym else torch.empty((N, G), dtype=torch.float32, device=W.device)
    check(_lib.load().b200woq_rtn_params(ptr(W), dt(W), N, K, bits, group_size, int(sym), int(full_range),
                                         float(quantile), ptr(scale), ptr(zp), stream_ptr(W.device)), "rtn_params")
    return scale, zp


def pack_params(scale: torch.Tensor, zp: Optional[torch.Tensor], bits: int):
    """scales fp16 [G,N], qzeros int32 [G, ceil(N/n_pack)] (modules.py:345-371)."""
    require_cuda(scale, "scale")
    N, G = scale.shape
    scale = scale.float().contiguous()
    zp = None if zp is None else zp.float().contiguous()
    scales16 = torch.empty((G, N), dtype=torch.float16, device=scale.device)
    qzeros = torch.empty((G, math.ceil(N / n_pack(bits))), dtype=torch.int32, device=scale.device)
    check(_lib.load().b200woq_pack_params(ptr(scale), ptr(zp), N, G, bits, ptr(scales16), ptr(qzeros),
                                          stream_ptr(scale.device)), "pack_params")
    return scales16, qzeros


def rtn_quant_pack(W: torch.Tensor, bits=4, group_size=-1, sym=False, full_range=False, quantile=1.0,
                   return_codes=False):
    """quant_tensor(return_int=True) + INCWeightOnlyLinear.pack.  Returns dict(qweight, qzeros, scales,
    scale_f32, zp_f32[, codes])."""
    require_cuda(W, "W")
    N, K = W.shape
    scale, zp = rtn_params(W, bits, group_size, sym, full_range, quantile)
    qweight = torch.empty((math.ceil(K / n_pack(bits)), N), dtype=torch.int32, device=W.device)
    codes = torch.empty((N, K), dtype=torch.uint8, device=W.device) if return_codes else None
    check(_lib.load().b200woq_rtn_quant_pack(ptr(W), dt(W), N, K, bits, group_size, int(sym), ptr(scale), ptr(zp),
                                             ptr(qweight), ptr(codes), stream_ptr(W.device)), "rtn_quant_pack")
    scales16, qzeros = pack_params(scale, zp, bits)
    out = dict(qweight=qweight, qzeros=qzeros, scales=scales16, scale_f32=scale, zp_f32=zp)
    if return_codes:
        out["codes"] = codes
    return out


def rtn_fake_quant(W: torch.Tensor, bits=4, group_size=-1, sym=False, full_range=False, quantile=1.0,
                   col_scale: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """quant_tensor(return_int=False); with col_scale: qdq(W*s)/s (awq.py:326-335)."""
    require_cuda(W, "W")
    N, K = W.shape
    if out is None:
        out = torch.empty_like(W)
    if col_scale is not None:
        col_scale = col_scale.float().contiguous()
    check(_lib.load().b200woq_rtn_fake_quant(ptr(W), dt(W), N, K, bits, group_size, int(sym), int(full_range),
                                             float(quantile), ptr(col_scale), ptr(out), stream_ptr(W.device)),
          "rtn_fake_quant")
    return out


def pack_codes(codes: torch.Tensor, bits: int):
    require_cuda(codes, "codes")
    assert codes.dtype == torch.uint8
    N, K = codes.shape
    qweight = torch.empty((math.ceil(K / n_pack(bits)), N), dtype=torch.int32, device=codes.device)
    check(_lib.load().b200woq_pack_codes(ptr(codes), N, K, bits, ptr(qweight), stream_ptr(codes.device)), "pack_codes")
    return qweight


def unpack(qweight, qzeros, bits, in_features, out_features, n_groups):
    require_cuda(qweight, "qweight")
    if qweight.dtype != torch.int32 or qzeros.dtype != torch.int32:
        raise _lib.B200WOQError("qweight / qzeros must be int32")
    codes = torch.empty((out_features, in_features), dtype=torch.uint8, device=qweight.device)
    zps = torch.empty((out_features, n_groups), dtype=torch.uint8, device=qweight.device)
    check(_lib.load().b200woq_unpack(ptr(qweight), ptr(qzeros), out_features, in_features, n_groups, bits, ptr(codes),
                                     ptr(zps), stream_ptr(qweight.device)), "unpack")
    return codes, zps


def dequantize(qweight, qzeros, scales, bits, group_size, in_features, out_features, g_idx=None):
    """recover(): fp16 [N,K] (modules.py:413-443)."""
    require_cuda(qweight, "qweight")
    _check_packed(qweight, qzeros, scales, bits, in_features, out_features, g_idx)
    out = torch.empty((out_features, in_features), dtype=torch.float16, device=qweight.device)
    check(_lib.load().b200woq_dequantize(ptr(qweight), ptr(qzeros), ptr(scales), ptr(g_idx), out_features, in_features,
                                         bits, group_size, ptr(out), stream_ptr(qweight.device)), "dequantize")
    return out


# ------------------------------------------------------------------ 4-bit table data types (nf4 / fp4 / fp4_e2m1)
def f4_quantize(W: torch.Tensor, dtype: str, group_size=-1, quantile=1.0, want_codes=True, fake_out=None):
    """quant_tensor(dtype=nf4|fp4|...) (utility.py:121-160, 272-376).  Returns dict(codes int8 [N,K] | None,
    scale fp32 [N,G]); with `fake_out` (same dtype / shape as W, may be W) also writes the fake-quantised weight."""
    import ctypes

    from . import dtypes

    require_cuda(W, "W")
    N, K = W.shape
    G = math.ceil(K / _eff_group(K, group_size))
    codes = torch.empty((N, K), dtype=torch.int8, device=W.device) if want_codes else None
    scale = torch.empty((N, G), dtype=torch.float32, device=W.device)
    if fake_out is not None:
        require_cuda(fake_out, "fake_out")
        if fake_out.dtype != W.dtype or fake_out.shape != W.shape:
            raise _lib.B200WOQError("fake_out must match W")
    table = dtypes.table(dtype)
    check(_lib.load().b200woq_f4_quantize(ptr(W), dt(W), N, K, int(group_size), ctypes.c_void_p(ctypes.addressof(table)),
                                          float(quantile), ptr(codes), ptr(scale), ptr(fake_out), stream_ptr(W.device)),
          "f4_quantize")
    return dict(codes=codes, scale=scale)


def pack_rows(codes: torch.Tensor, bits: int):
    """Signed / unsigned int8 codes [N,K] -> int32 [N, ceil(K/n_pack)], fields along K (modules.py:352-357, 445-466)."""
    require_cuda(codes, "codes")
    if codes.dtype != torch.int8:
        raise _lib.B200WOQError("codes must be int8")
    N, K = codes.shape
    qweight = torch.empty((N, math.ceil(K / n_pack(bits))), dtype=torch.int32, device=codes.device)
    check(_lib.load().b200woq_pack_rows(ptr(codes), N, K, bits, ptr(qweight), stream_ptr(codes.device)), "pack_rows")
    return qweight


def f4_dequantize(qweight: torch.Tensor, scales: torch.Tensor, dtype: str, group_size: int, in_features: int):
    """recover() of a table-dtype module (modules.py:377-443): fp32 [N,K] = level(nibble) * scale."""
    from . import dtypes

    require_cuda(qweight, "qweight")
    require_cuda(scales, "scales")
    N = qweight.shape[0]
    G = math.ceil(in_features / _eff_group(in_features, group_size))
    if (qweight.dtype != torch.int32 or scales.dtype != torch.float32 or tuple(qweight.shape) != (N, math.ceil(in_features / 8))
            or tuple(scales.shape) != (N, G)):
        raise _lib.B200WOQError(f"f4_dequantize: qweight int32 [N, ceil(K/8)] and scales fp32 [N, G] expected, got "
                                f"{qweight.dtype} {tuple(qweight.shape)}, {scales.dtype} {tuple(scales.shape)}")
    out = torch.empty((N, in_features), dtype=torch.float32, device=qweight.device)
    levels = dtypes.nibble_levels(dtype)
    check(_lib.load().b200woq_f4_dequantize(ptr(qweight), ptr(scales), levels, N, in_features, int(group_size), ptr(out),
                                            stream_ptr(qweight.device)), "f4_dequantize")
    return out


# ------------------------------------------------------------------ K6: fused dequant GEMM
_WS_CACHE = {}


def _check_packed(qweight, qzeros, scales, bits, in_features, out_features, g_idx=None, input_scale=None, bias=None):
    """The kernels reinterpret raw pointers: a `model.to(torch.bfloat16)` that silently re-typed `scales`, a float
    `g_idx` or a mis-shaped tensor would decode garbage without any error, so every entry validates dtype, shape, device
    and contiguity of the packed tensors first."""
    np_ = n_pack(bits)

    def need(t, name, dtype, shape=None):
        if t is None:
            return
        if not t.is_cuda:
            raise _lib.B200WOQError(f"{name} must be a CUDA tensor (neural_compressor_b200 has no CPU path)")
        if t.dtype != dtype:
            raise _lib.B200WOQError(f"{name} must be {dtype}, got {t.dtype} (was the packed module cast with .to(dtype)?)")
        if not t.is_contiguous():
            raise _lib.B200WOQError(f"{name} must be contiguous")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise _lib.B200WOQError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")

    need(qweight, "qweight", torch.int32, (math.ceil(in_features / np_), out_features))
    G = scales.shape[0] if scales is not None and scales.dim() == 2 else -1
    need(scales, "scales", torch.float16, (G, out_features))
    need(qzeros, "qzeros", torch.int32, (G, math.ceil(out_features / np_)))
    need(g_idx, "g_idx", torch.int32, (in_features,))
    need(input_scale, "input_scale", torch.float32, (in_features,))
    if bias is not None:
        if bias.dtype not in _lib._DT:
            raise _lib.B200WOQError(f"bias dtype {bias.dtype} unsupported")
        need(bias, "bias", bias.dtype, (out_features,))


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (device.type, device.index)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)  # kernels keep it zeroed
        _WS_CACHE[key] = ws
    return ws


def woq_linear(x, qweight, qzeros, scales, bias, bits, group_size, in_features, out_features, g_idx=None,
               input_scale=None, out_dtype=torch.float32, flags=0, out=None):
    """INCWeightOnlyLinear.forward (modules.py:594-610) as one fused kernel."""
    require_cuda(x, "x")
    _check_packed(qweight, qzeros, scales, bits, in_features, out_features, g_idx, input_scale, bias)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, in_features)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M = x2.shape[0]
    lib = _lib.load()
    nbytes = lib.b200woq_linear_workspace_bytes(M, out_features, in_features, bits, group_size)
    ws = _workspace(x.device, nbytes)
    if out is None:
        out = torch.empty((M, out_features), dtype=out_dtype, device=x.device)
    check(lib.b200woq_linear_forward(ptr(x2), dt(x2), M, in_features, out_features, ptr(qweight), ptr(qzeros),
                                     ptr(scales), ptr(g_idx), ptr(bias), dt(bias) if bias is not None else 0,
                                     ptr(input_scale), ptr(out), dt(out), bits, group_size, ptr(ws), ws.numel(),
                                     flags, stream_ptr(x.device)), "linear_forward")
    return out.reshape(*lead, out_features)


def build_stream_layout(qweight, qzeros, scales, bits, group_size, in_features, out_features):
    """Derived B200-native layout for the small-batch 4-bit path (woq_stream.cu); None when not eligible."""
    require_cuda(qweight, "qweight")
    _check_packed(qweight, qzeros, scales, bits, in_features, out_features)
    lib = _lib.load()
    nbytes = lib.b200woq_stream_layout_bytes(out_features, in_features, bits, group_size)
    if nbytes <= 0:
        return None
    out = torch.empty(nbytes, dtype=torch.uint8, device=qweight.device)
    check(lib.b200woq_build_stream_layout(ptr(qweight), ptr(qzeros), ptr(scales), out_features, in_features, bits,
                                          group_size, ptr(out), stream_ptr(qweight.device)), "build_stream_layout")
    return out


def woq_linear_stream(x, stream_layout, bias, bits, group_size, in_features, out_features, input_scale=None,
                      out_dtype=torch.float32, flags=0, out=None):
    """INCWeightOnlyLinear.forward for M <= 4 on the stream layout (per-warp TMA bulk-copy rings)."""
    require_cuda(x, "x")
    if stream_layout.dtype != torch.uint8 or not stream_layout.is_cuda:
        raise _lib.B200WOQError("stream_layout must be the uint8 CUDA tensor returned by build_stream_layout")
    if input_scale is not None and (input_scale.dtype != torch.float32 or not input_scale.is_contiguous()):
        raise _lib.B200WOQError("input_scale must be a contiguous float32 tensor")
    lead = x.shape[:-1]
    x2 = x.reshape(-1, in_features)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M = x2.shape[0]
    if out is None:
        out = torch.empty((M, out_features), dtype=out_dtype, device=x.device)
    check(_lib.load().b200woq_linear_forward_stream(ptr(x2), dt(x2), M, in_features, out_features, ptr(stream_layout),
                                                    ptr(bias), dt(bias) if bias is not None else 0, ptr(input_scale),
                                                    ptr(out), dt(out), bits, group_size, flags, stream_ptr(x.device)),
          "linear_forward_stream")
    return out.reshape(*lead, out_features)


# ------------------------------------------------------------------ K1-K3: GPTQ
PROFILE_HOOK = None  # bench.py sets this to a list: (start_event, end_event, algorithmic_flops) per Hessian launch


def hessian_accumulate(X: torch.Tensor, Hsum: torch.Tensor):
    """Hsum += X^T X (raw fp32 sums; gptq.py:1111-1141 closed form, see hessian_finalize)."""
    require_cuda(X, "X")
    X2 = X.reshape(-1, X.shape[-1])
    if not X2.is_contiguous():
        X2 = X2.contiguous()
    T, C = X2.shape
    assert Hsum.shape == (C, C) and Hsum.dtype == torch.float32 and Hsum.is_contiguous()
    if PROFILE_HOOK is not None:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
    check(_lib.load().b200woq_hessian_accumulate(ptr(X2), dt(X2), T, C, C, ptr(Hsum), stream_ptr(X.device)),
          "hessian_accumulate")
    if PROFILE_HOOK is not None:
        e.record()
        PROFILE_HOOK.append((s, e, float(T) * C * (C + 128)))  # symmetric half incl. diagonal tiles


def hessian_finalize(Hsum: torch.Tensor, nsamples: float, percdamp: float):
    """In place: H = 2/n * Hsum (full symmetric), dead handling, damping.  Returns (H, dead_mask uint8[C])."""
    require_cuda(Hsum, "Hsum")
    C = Hsum.shape[0]
    dead = torch.empty(C, dtype=torch.uint8, device=Hsum.device)
    scratch = torch.empty(2, dtype=torch.float32, device=Hsum.device)
    check(_lib.load().b200woq_hessian_finalize(ptr(Hsum), C, float(nsamples), float(percdamp), ptr(dead), ptr(scratch),
                                               stream_ptr(Hsum.device)), "hessian_finalize")
    return Hsum, dead


_CHOL_WS = {}


def _cholinv_workspace(device, nbytes: int) -> torch.Tensor:
    """One workspace per (device, stream): concurrent factorisations on side streams must not share buffers."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _CHOL_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _CHOL_WS[key] = ws
    return ws


def release_workspaces():
    """Drop the cached K2 / dequant-GEMM workspaces (about 1 GB per stream after a C = 11008 factorisation)."""
    _CHOL_WS.clear()
    _WS_CACHE.clear()


def cholesky_inverse_upper(H: torch.Tensor, info: Optional[torch.Tensor] = None, check: bool = True) -> torch.Tensor:
    """Upper Cholesky factor U of H^-1 (U^T U = H^-1), gptq.py:1228-1231 -- K2, hand-written (cholinv.cu).

    The reference chains chol -> cholesky_inverse -> chol(upper) (4/3 C^3 flop).  U is unique, so it is computed with
    half the work from the index-reversed factorisation: with J the exchange matrix, J H J = L L^T gives
    H = (J L J)(J L J)^T with J L J upper triangular, hence U = (J L J)^-1 = J L^-1 J: one blocked Cholesky and one
    blocked triangular inverse, exact fp32 FFMA.  `info` (int32[1] device tensor) receives the factorisation status
    (0 = ok); with check=True (default) the status is read back (one host sync) and a failed factorisation raises like
    the reference's torch.linalg.cholesky; the engine passes check=False and tests all statuses once per block.
    B200WOQ_CHOLINV=torch selects the cuSOLVER/cuBLAS cross-check path (library calls, not the product path)."""
    import os

    require_cuda(H, "H")
    assert H.dtype == torch.float32 and H.dim() == 2 and H.shape[0] == H.shape[1]
    mode = os.environ.get("B200WOQ_CHOLINV", "b200")
    if mode == "chain":
        L = torch.linalg.cholesky(H)
        Hi = torch.cholesky_inverse(L)
        return torch.linalg.cholesky(Hi, upper=True).contiguous()
    if mode == "torch":
        Lf = torch.linalg.cholesky(H.flip(0, 1))
        Ut = Lf.flip(0, 1)  # upper triangular, H = Ut Ut^T
        eye = torch.eye(H.shape[0], dtype=H.dtype, device=H.device)
        return torch.linalg.solve_triangular(Ut, eye, upper=True).contiguous()
    C = H.shape[0]
    lib = _lib.load()
    nbytes = lib.b200woq_cholinv_workspace_bytes(C)
    ws = _cholinv_workspace(H.device, nbytes)
    U = torch.empty_like(H)
    if info is None:
        info = torch.zeros(1, dtype=torch.int32, device=H.device)
    check_rc = lib.b200woq_cholinv_upper(ptr(H), C, ptr(U), ptr(ws), ws.numel(), ptr(info), stream_ptr(H.device))
    _lib.check(check_rc, "cholinv_upper")
    if check:
        st = int(info.item())
        if st != 0:
            raise torch.linalg.LinAlgError(
                f"cholinv_upper: the Hessian is not positive-definite (pivot {C - st} of {C}); raise percdamp")
    return U


def gptq_fasterquant(W: torch.Tensor, Hinv: torch.Tensor, dead_mask: Optional[torch.Tensor], blocksize=128,
                     groupsize=-1, bits=4, sym=False, mse=False, want_q=True, double_quant=None):
    """GPTQ.fasterquant column loop for one layer (gptq.py:1250-1304).  W fp32 [N,C] is destroyed.

    `double_quant` = None or dict(bits, group_size, sym): fake-quantise each group's scales over the output rows
    (Quantizer.find_params with use_double_quant, gptq.py:1598-1614).

    Returns dict(codes uint8 [N,C], Q fp32 [N,C] | None, scale [N,G], zero [N,G], losses [N])."""
    require_cuda(W, "W")
    assert W.dtype == torch.float32 and Hinv.dtype == torch.float32 and Hinv.is_contiguous()
    N, C = W.shape
    G = 1 if groupsize <= 0 else math.ceil(C / groupsize)
    dev = W.device
    codes = torch.empty((N, C), dtype=torch.uint8, device=dev)
    Q = torch.empty((N, C), dtype=torch.float32, device=dev) if want_q else None
    scale = torch.empty((N, G), dtype=torch.float32, device=dev)
    zero = torch.empty((N, G), dtype=torch.float32, device=dev)
    losses = torch.empty(N, dtype=torch.float32, device=dev)
    lib = _lib.load()
    nbytes = lib.b200woq_gptq_workspace_bytes(N, C, blocksize)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    flags = 1 if mse else 0
    if double_quant:
        flags |= 2 | (4 if double_quant.get("sym") else 0) | (int(double_quant.get("bits", 8)) << 8) | \
            (int(double_quant.get("group_size", 256)) << 16)
    check(lib.b200woq_gptq_fasterquant(ptr(W), ptr(Hinv), ptr(dead_mask), N, C, blocksize, groupsize, bits, int(sym),
                                       flags, ptr(codes), ptr(Q), ptr(scale), ptr(zero), ptr(losses),
                                       ptr(ws), nbytes, stream_ptr(dev)), "gptq_fasterquant")
    return dict(codes=codes, Q=Q, scale=scale, zero=zero, losses=losses)


def gptq_rebuild_q(codes: torch.Tensor, scale: torch.Tensor, zero: torch.Tensor, groupsize: int) -> torch.Tensor:
    """Q = scale * (codes - zero), the fake-quant weights of gptq.py:1636-1637, from the u8 codes (bit-identical to the Q
    `gptq_fasterquant` emits)."""
    require_cuda(codes, "codes")
    assert codes.dtype == torch.uint8 and scale.dtype == torch.float32 and zero.dtype == torch.float32
    N, C = codes.shape
    Q = torch.empty((N, C), dtype=torch.float32, device=codes.device)
    check(_lib.load().b200woq_gptq_rebuild_q(ptr(codes), ptr(scale.contiguous()), ptr(zero.contiguous()), N, C, groupsize,
                                             ptr(Q), stream_ptr(codes.device)), "gptq_rebuild_q")
    return Q


# ------------------------------------------------------------------ K5/K7 statistics
def awq_weight_scale(W: torch.Tensor, group_size: int) -> torch.Tensor:
    require_cuda(W, "W")
    N, K = W.shape
    out = torch.empty(K, dtype=torch.float32, device=W.device)
    check(_lib.load().b200woq_awq_weight_scale(ptr(W), dt(W), N, K, group_size, ptr(out), stream_ptr(W.device)),
          "awq_weight_scale")
    return out


def abs_colsum_accumulate(X: torch.Tensor, acc: torch.Tensor):
    require_cuda(X, "X")
    X2 = X.reshape(-1, X.shape[-1])
    if not X2.is_contiguous():
        X2 = X2.contiguous()
    T, K = X2.shape
    check(_lib.load().b200woq_abs_colsum_accumulate(ptr(X2), dt(X2), T, K, K, ptr(acc), stream_ptr(X.device)),
          "abs_colsum_accumulate")
    return T


def mse_accumulate(a: torch.Tensor, b: torch.Tensor, acc: torch.Tensor):
    """acc (float64[1]) += float mean((a-b)^2)  (awq.py:343-344)."""
    require_cuda(a, "a")
    assert a.dtype == b.dtype and a.numel() == b.numel() and acc.dtype == torch.float64
    a = a.contiguous()
    b = b.contiguous()
    check(_lib.load().b200woq_mse_accumulate(ptr(a), ptr(b), dt(a), a.numel(), ptr(acc), stream_ptr(a.device)),
          "mse_accumulate")


def minmax_cols_accumulate(X: torch.Tensor, mx: torch.Tensor, mn: torch.Tensor):
    require_cuda(X, "X")
    X2 = X.reshape(-1, X.shape[-1])
    if not X2.is_contiguous():
        X2 = X2.contiguous()
    T, K = X2.shape
    check(_lib.load().b200woq_minmax_cols_accumulate(ptr(X2), dt(X2), T, K, K, ptr(mx), ptr(mn), stream_ptr(X.device)),
          "minmax_cols_accumulate")


# ------------------------------------------------------------------ K7: SmoothQuant W8A8
def sq_smooth_quant_weight(W: torch.Tensor, smooth: Optional[torch.Tensor] = None):
    """W' = W * smooth (per input channel), per-out-channel sym int8 (quant_dequant_w_v1, smooth_quant/utility.py:652-690).

    Returns dict(qweight int8 [N, Kp] zero padded to a multiple of 128, w_scale fp32 [N], wsum int32 [N])."""
    require_cuda(W, "W")
    N, K = W.shape
    lib = _lib.load()
    Kp = lib.b200woq_w8a8_padded_k(K)
    qweight = torch.empty((N, Kp), dtype=torch.int8, device=W.device)
    w_scale = torch.empty(N, dtype=torch.float32, device=W.device)
    wsum = torch.empty(N, dtype=torch.int32, device=W.device)
    if smooth is not None:
        smooth = smooth.float().contiguous()
    check(lib.b200woq_sq_smooth_quant_weight(ptr(W), dt(W), N, K, ptr(smooth), ptr(qweight), ptr(w_scale), ptr(wsum),
                                             stream_ptr(W.device)), "sq_smooth_quant_weight")
    return dict(qweight=qweight, w_scale=w_scale, wsum=wsum)


_W8A8_WS = {}


def w8a8_linear(x, qweight, w_scale, wsum, x_scale, x_zp, in_features, input_scale=None, bias=None, out_dtype=None):
    """SQLinearWrapper.forward as a static INT8 GEMM (tcgen05.mma.kind::i8) with the dequant epilogue."""
    require_cuda(x, "x")
    lead = x.shape[:-1]
    x2 = x.reshape(-1, in_features)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M, K = x2.shape
    N = qweight.shape[0]
    lib = _lib.load()
    if qweight.dtype != torch.int8 or qweight.shape[1] != lib.b200woq_w8a8_padded_k(K) or not qweight.is_contiguous():
        raise _lib.B200WOQError("qweight must be the contiguous int8 [N, padded_k(K)] tensor of sq_smooth_quant_weight")
    for t, name, dtp in ((w_scale, "w_scale", torch.float32), (wsum, "wsum", torch.int32), (x_scale, "x_scale", torch.float32),
                         (x_zp, "x_zp", torch.float32)):
        if t.dtype != dtp or not t.is_cuda or not t.is_contiguous():
            raise _lib.B200WOQError(f"{name} must be a contiguous CUDA {dtp} tensor")
    if input_scale is not None and (input_scale.dtype != torch.float32 or not input_scale.is_contiguous()):
        raise _lib.B200WOQError("input_scale must be contiguous float32")
    nbytes = lib.b200woq_w8a8_workspace_bytes(M, N, K)
    zoff = lib.b200woq_w8a8_workspace_zeroed_offset(M, K)
    # the split-K tail of the workspace is self-cleaning but must start zeroed: one cached buffer per (device, M, N, K)
    key = (x.device.index, M, N, K)
    ws = _W8A8_WS.get(key)
    if ws is None:
        if len(_W8A8_WS) > 64:
            _W8A8_WS.clear()
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=x.device)
        _W8A8_WS[key] = ws
    out_dtype = out_dtype or (x.dtype if x.dtype in _lib._DT else torch.float32)
    y = torch.empty((M, N), dtype=out_dtype, device=x.device)
    check(lib.b200woq_w8a8_linear_forward(ptr(x2), dt(x2), M, K, N, ptr(qweight), ptr(w_scale), ptr(wsum), ptr(input_scale),
                                          ptr(x_scale), ptr(x_zp), ptr(bias), dt(bias) if bias is not None else 0, ptr(y),
                                          dt(y), ptr(ws), ws.numel(), stream_ptr(x.device)), "w8a8_linear_forward")
    assert zoff <= nbytes
    return y.reshape(*lead, N)
