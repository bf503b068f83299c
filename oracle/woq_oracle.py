"""CPU oracle: a restatement of intel/neural-compressor's torch weight-only path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module;
the product package `neural_compressor_b200` never does (it fails loudly when the
CUDA library is missing instead of falling back to anything here).

Every function restates, on torch **CPU** tensors, the arithmetic of one reference
function (cited `file:line`, paths relative to
`/root/reference/neural_compressor/torch/algorithms/weight_only/`).  The reference is
pure Python + torch-eager CPU ops, so the restatement issues the *same torch CPU op
sequence per element* (same dtype promotion, same rounding points) but vectorised where
the reference loops in Python.  Parity is PINNED: `tests/test_oracle_vs_golden.py`
checks every function below against fixtures in `tests/golden/*.pt` that were produced
by running the unmodified reference in-process (`oracle/gen_golden.py`), and
`tests/test_oracle_vs_reference.py` re-checks against the live reference whenever
`/root/reference` is mounted.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import torch

CPU = torch.device("cpu")


def _cpu(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().to(CPU)


# ----------------------------------------------------------------------------------------
# RTN group quantisation  (utility.py:162-269 qdq_weight_{asym,sym,actor}; :272-436 quant_tensor)
# ----------------------------------------------------------------------------------------
def _rows_asym(w2d: torch.Tensor, bits: int, quantile: float):
    """utility.py:162-198 (qdq_weight_asym, return_int=True branch). `w2d` is consumed."""
    maxq = 2**bits - 1
    zero_f32 = torch.zeros(w2d.shape[0])  # float32 on purpose (utility.py:177): promotes fp16 rows
    lo = torch.minimum(w2d.min(1)[0], zero_f32) * quantile
    hi = torch.maximum(w2d.max(1)[0], zero_f32) * quantile
    flat = (lo == 0) & (hi == 0)
    lo[flat] = -1
    hi[flat] = +1
    scale = (hi - lo) / maxq
    zp = torch.round(-lo / scale)
    scale = scale.unsqueeze(-1)
    zp = zp.unsqueeze(-1)
    w2d.div_(scale).round_().add_(zp).clamp_(0, maxq)
    return w2d, scale, zp


def _rows_sym(w2d: torch.Tensor, bits: int, quantile: float, full_range: bool):
    """utility.py:201-244 (qdq_weight_sym, return_int=True branch). `w2d` is consumed."""
    maxq = 2 ** (bits - 1) - 1
    minq = -(2 ** (bits - 1))
    if bits == 1:
        maxq, minq = 1, 0
    hi = torch.max(w2d, 1)[0]
    lo = torch.min(w2d, 1)[0]
    flip = torch.abs(hi) > torch.abs(lo)
    amax = torch.max(torch.abs(hi), torch.abs(lo)) * quantile
    amax[amax == 0] = 1
    if full_range:
        scale = amax / (-minq)
        scale = torch.where(flip, -scale, scale)
    else:
        scale = amax / maxq
    scale = scale.unsqueeze(-1)
    w2d.div_(scale).round_().clamp_(minq, maxq)
    return w2d, scale, None


def rtn_quantize(
    weight: torch.Tensor,
    bits: int = 4,
    group_size: int = -1,
    scheme: str = "asym",
    quantile: float = 1.0,
    full_range: bool = False,
) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """`quant_tensor(..., dtype="int", return_int=True)`  (utility.py:272-376).

    Returns (codes [N,K] in weight.dtype holding integers, scale [N,G], zp [N,G] | None).
    Not in place (the reference is; callers that rely on that copy the result back).
    Group layout: reshape to [N*K/g, g] when K % g == 0 (:309-333), otherwise the aligned
    prefix and the ragged tail are quantised separately and concatenated (:334-376).
    """
    w = _cpu(weight).clone()
    n, k = w.shape
    if group_size == -1 or k < group_size:
        group_size = k
    rows = (lambda t: _rows_sym(t, bits, quantile, full_range)) if scheme == "sym" else (
        lambda t: _rows_asym(t, bits, quantile)
    )
    if k % group_size == 0:
        q, s, z = rows(w.reshape(-1, group_size))
        return q.reshape(n, k), s.reshape(n, -1), (None if z is None else z.reshape(n, -1))
    split = k // group_size * group_size
    q1, s1, z1 = rows(w[:, :split].reshape(-1, group_size))
    q2, s2, z2 = rows(w[:, split:].clone())
    q = torch.cat([q1.reshape(n, split), q2], dim=1)
    s = torch.cat([s1.reshape(n, -1), s2], dim=1)
    z = None if (z1 is None or z2 is None) else torch.cat([z1.reshape(n, -1), z2], dim=1)
    return q, s, z


def rtn_fake_quant(weight, bits=4, group_size=-1, scheme="asym", quantile=1.0, full_range=False):
    """`quant_tensor(..., return_int=False)`: de-quantised weights (utility.py:195-198, 241-244)."""
    q, s, z = rtn_quantize(weight, bits, group_size, scheme, quantile, full_range)
    n, k = q.shape
    g = k if (group_size == -1 or k < group_size) else group_size
    idx = torch.arange(k) // g
    if z is not None:
        q = q.sub_(z[:, idx])
    return q.mul_(s[:, idx])


def rtn_search_clip(weight, bits=4, group_size=32, scheme="asym", full_range=False) -> float:
    """utility.py:439-480 (`search_clip`): 40 ratios 1 - i/200, full-tensor MSE, first strict best."""
    org = _cpu(weight).clone()
    best_err, best = float("inf"), None
    for i in range(int(0.2 * 200)):
        ratio = 1 - i / 200
        fq = rtn_fake_quant(org, bits, group_size, scheme, ratio, full_range)
        loss = (org - fq).float().pow(2).mean()
        if loss < best_err:
            best_err, best = loss, ratio
    return best


# ----------------------------------------------------------------------------------------
# fake-quant fp -> integer codes with given scale (utility.py:483-537 quant_weight_w_scale)
# ----------------------------------------------------------------------------------------
def codes_from_fake_quant(q_fp, scale, zp=None, group_size=-1):
    """`quant_weight_w_scale(..., dtype="int", fp8_aware=False)`: round(Q/scale (+zp)) per group."""
    w = _cpu(q_fp).clone()
    scale = _cpu(scale)
    zp = _cpu(zp)
    if group_size == -1:
        w.div_(scale)
        if zp is not None:
            w.add_(zp)
        return w.round_()
    k = w.shape[1]
    idx = torch.clamp(torch.arange(k) // group_size, max=scale.shape[1] - 1)
    # the reference divides/rounds in-place in w's dtype and copies into a float32 buffer (:514-525)
    w.div_(scale[:, idx])
    if zp is not None:
        w.add_(zp[:, idx])
    return w.round_().to(torch.float32)


# ----------------------------------------------------------------------------------------
# packing (modules.py:321-375 pack, :445-466/528-556 pack_tensor_*, bit_packer.py:35-278)
# ----------------------------------------------------------------------------------------
def _pack_cols(v: np.ndarray, bits: int) -> np.ndarray:
    """word[:, j] = OR_e (v[:, j*n_pack+e] & mask) << (bits*e), 32-bit words (modules.py:533-543)."""
    n_pack = 32 // bits
    rows, cols = v.shape
    words = (cols + n_pack - 1) // n_pack
    padded = np.zeros((rows, words * n_pack), dtype=np.uint64)
    padded[:, :cols] = v.astype(np.int64).astype(np.uint64) & np.uint64(2**bits - 1)
    padded = padded.reshape(rows, words, n_pack)
    shifts = (np.arange(n_pack, dtype=np.uint64) * np.uint64(bits)).reshape(1, 1, n_pack)
    out = np.bitwise_or.reduce(padded << shifts, axis=2) & np.uint64(0xFFFFFFFF)
    return out.astype(np.uint32).view(np.int32)


def pack_optimum(int_weight, scales, zp, bits: int, group_size: int):
    """`INCWeightOnlyLinear.pack` in the default optimum format (modules.py:321-375).

    int_weight [N,K] integer-valued; scales [N,G]; zp [N,G] or None (sym).
    Returns qweight int32 [ceil(K/n_pack), N], qzeros int32 [G, ceil(N/n_pack)], scales fp16 [G, N].
    sym: codes += 2^(b-1), zp := 2^(b-1) (:329-334);  zp -= 1 before packing (:363-364).
    """
    w = _cpu(int_weight).to(torch.int32).numpy().astype(np.int64)
    scales = _cpu(scales)
    if zp is None:
        w = w + 2 ** (bits - 1)
        z = np.full(tuple(scales.shape), 2 ** (bits - 1), dtype=np.int64)
    else:
        z = _cpu(zp).clone().to(torch.float32).numpy().astype(np.int64)
    z = z - 1
    qweight = np.ascontiguousarray(_pack_cols(w, bits).T)  # [N, Kw] -> [Kw, N]
    qzeros = _pack_cols(np.ascontiguousarray(z.T), bits)  # [G, N] packed along N
    return (
        torch.from_numpy(qweight.copy()),
        torch.from_numpy(qzeros.copy()),
        scales.to(torch.float16).t().contiguous(),
    )


def _unpack_cols(words: np.ndarray, bits: int) -> np.ndarray:
    """modules.py:558-578: (word << (32-bits(e+1))) >> (32-bits) then & mask -> unsigned field e."""
    n_pack = 32 // bits
    u = words.view(np.uint32).astype(np.uint64)
    shifts = (np.arange(n_pack, dtype=np.uint64) * np.uint64(bits)).reshape(1, 1, n_pack)
    out = (u[:, :, None] >> shifts) & np.uint64(2**bits - 1)
    return out.reshape(words.shape[0], -1).astype(np.int16)


def unpack_optimum(qweight, qzeros, scales, bits: int, in_features: int, out_features: int):
    """`INCWeightOnlyLinear.unpack` (modules.py:377-411): codes [N,K] int16, zp [N,G] int16, scales [N,G]."""
    qw = _cpu(qweight).t().contiguous().numpy()
    codes = _unpack_cols(qw, bits)[:out_features, :in_features]
    sc = _cpu(scales).t().contiguous()
    qz = _cpu(qzeros).contiguous().numpy()
    z = _unpack_cols(qz, bits).T[: sc.shape[0], : sc.shape[1]].astype(np.int16)
    z = z + 1
    z = np.where(z > 2**bits - 1, 0, z).astype(np.int16)
    return torch.from_numpy(codes.copy()), torch.from_numpy(z.copy()), sc


def recover_fp16(qweight, qzeros, scales, bits, group_size, in_features, out_features, g_idx=None):
    """`INCWeightOnlyLinear.recover` (modules.py:413-443): fp16( int8(q - zp) * scale_fp16 )."""
    codes, z, sc = unpack_optimum(qweight, qzeros, scales, bits, in_features, out_features)
    if g_idx is None:
        gi = torch.arange(in_features) // group_size
    else:
        gi = _cpu(g_idx).long()
    return (codes - z[:, gi]).to(torch.int8) * sc[:, gi]


def woq_linear_forward(x, qweight, qzeros, scales, bias, bits, group_size, in_features, out_features, g_idx=None):
    """`INCWeightOnlyLinear.forward` on CPU (modules.py:594-610): fp32 F.linear on the recovered weight."""
    w = recover_fp16(qweight, qzeros, scales, bits, group_size, in_features, out_features, g_idx).float()
    b = None if bias is None else _cpu(bias).float()
    return torch.nn.functional.linear(_cpu(x).float(), w, b)


# ----------------------------------------------------------------------------------------
# GPTQ  (gptq.py:1089-1351 GPTQ, :1364-1644 Quantizer)
# ----------------------------------------------------------------------------------------
class GPTQGroupQuantizer:
    """gptq.py:1364-1637 `Quantizer` for dtype="int", perchannel=True, no double-quant."""

    def __init__(self, bits=4, sym=False, mse=False, norm=2.4, grid=100, maxshrink=0.8):
        self.bits, self.sym, self.mse = bits, sym, mse
        self.maxq = 2**bits - 1
        self.norm, self.grid, self.maxshrink = norm, grid, maxshrink
        self.scale = torch.zeros(1)
        self.zero = torch.zeros(1)

    def ready(self):  # gptq.py:1639-1645
        return bool(torch.all(self.scale != 0))

    def quantize(self, x, scale, zero):  # gptq.py:1626-1637
        q = torch.clamp(torch.round(x / scale) + zero, 0, self.maxq)
        return scale * (q - zero)

    def find_params(self, x):  # gptq.py:1501-1596 (weight=True)
        x = x.flatten(1)
        z = torch.zeros(x.shape[0])
        xmin = torch.minimum(x.min(1)[0], z)
        xmax = torch.maximum(x.max(1)[0], z)
        if self.sym:
            xmax = torch.maximum(torch.abs(xmin), xmax)
            neg = xmin < 0
            if torch.any(neg):
                xmin[neg] = -xmax[neg]
        flat = (xmin == 0) & (xmax == 0)
        xmin[flat] = -1
        xmax[flat] = +1
        scale = (xmax - xmin) / self.maxq
        if self.sym:
            zero = torch.full_like(scale, (self.maxq + 1) / 2)
        else:
            zero = torch.round(-xmin / scale)
        if self.mse:
            best = torch.full([x.shape[0]], float("inf"))
            for i in range(int(self.maxshrink * self.grid)):
                p = 1 - i / self.grid
                xmin1, xmax1 = p * xmin, p * xmax
                scale1 = (xmax1 - xmin1) / self.maxq
                zero1 = torch.round(-xmin1 / scale1) if not self.sym else zero
                q = self.quantize(x, scale1.unsqueeze(1), zero1.unsqueeze(1))
                err = (q - x).abs_().pow_(self.norm).sum(1)
                better = err < best
                if torch.any(better):
                    best[better] = err[better]
                    scale[better] = scale1[better]
                    zero[better] = zero1[better]
        self.scale = scale.reshape(-1, 1)
        self.zero = zero.reshape(-1, 1)


class GPTQLayerOracle:
    """gptq.py:1089-1351 for an nn.Linear layer: Hessian accumulation + `fasterquant`."""

    def __init__(self, rows: int, columns: int, bits=4, sym=False, mse=False):
        self.rows, self.columns = rows, columns
        self.H = torch.zeros((columns, columns))
        self.nsamples = 0
        self.quantizer = GPTQGroupQuantizer(bits=bits, sym=sym, mse=mse)
        self.perm = None

    def add_batch(self, inp: torch.Tensor):
        """gptq.py:1111-1141: H <- H*n/(n+b) ; H += (sqrt(2/n') X)^T (sqrt(2/n') X), b = batch dim."""
        inp = _cpu(inp)
        if inp.dim() == 2:
            inp = inp.unsqueeze(0)
        b = inp.shape[0]
        x = inp.reshape(-1, inp.shape[-1]).t()
        self.H *= self.nsamples / (self.nsamples + b)
        self.nsamples += b
        x = math.sqrt(2 / self.nsamples) * x.float()
        self.H += x.matmul(x.t())

    @staticmethod
    def cholesky_inverse_upper(H: torch.Tensor) -> torch.Tensor:
        """gptq.py:1228-1231: chol -> cholesky_inverse -> chol(upper)."""
        L = torch.linalg.cholesky(H)
        Hi = torch.cholesky_inverse(L)
        return torch.linalg.cholesky(Hi, upper=True)

    def prepare_hinv(self, W: torch.Tensor, percdamp=0.01, act_order=False):
        """gptq.py:1189-1231: dead columns, optional act_order permutation, damping, inverse factor.

        Returns (W possibly permuted & dead-zeroed, Hinv upper, perm | None)."""
        H = self.H.clone()
        dead = torch.diag(H) == 0
        H[dead, dead] = 1
        W = W.clone()
        W[:, dead] = 0
        perm = None
        if act_order:
            perm = torch.argsort(torch.diag(H), descending=True)
            W = W[:, perm]
            H = H[perm][:, perm]
        damp = percdamp * torch.mean(torch.diag(H))
        idx = torch.arange(self.columns)
        H[idx, idx] += damp
        return W, self.cholesky_inverse_upper(H), perm

    def fasterquant(self, W, blocksize=128, percdamp=0.01, groupsize=-1, act_order=False,
                    static_groups=False, hinv: Optional[torch.Tensor] = None):
        """gptq.py:1143-1351 (int dtype; no hybrid_order / fp8_aware).

        Returns dict(scale [N,G], zero [N,G], Q [N,C] fake-quant fp32 in original column order,
        losses [N,C], perm, hinv).  `hinv` may be injected to test the column loop in isolation."""
        W = _cpu(W).float().clone()
        qz = self.quantizer
        if not qz.ready():
            qz.find_params(W)
        if hinv is None:
            W, Hinv, perm = self.prepare_hinv(W, percdamp, act_order)
        else:
            Hinv, perm = _cpu(hinv), None
        groups = None
        if static_groups:  # gptq.py:1193-1200 (computed on the un-permuted W in the reference)
            raise NotImplementedError("static_groups is not restated")
        C = self.columns
        Q = torch.zeros_like(W)
        Losses = torch.zeros_like(W)
        scales: List[torch.Tensor] = []
        zeros: List[torch.Tensor] = []
        for i1 in range(0, C, blocksize):
            i2 = min(i1 + blocksize, C)
            cnt = i2 - i1
            W1 = W[:, i1:i2].clone()
            Q1 = torch.zeros_like(W1)
            Err1 = torch.zeros_like(W1)
            L1 = torch.zeros_like(W1)
            Hinv1 = Hinv[i1:i2, i1:i2]
            for i in range(cnt):
                w = W1[:, i]
                d = Hinv1[i, i]
                if groupsize != -1 and (i1 + i) % groupsize == 0:
                    # NOTE (parity trap): reads the *global* W, which only carries the
                    # lazy cross-block updates, not this block's rank-1 updates (gptq.py:1270).
                    qz.find_params(W[:, (i1 + i):(i1 + i + groupsize)])
                    scales.append(qz.scale)
                    zeros.append(qz.zero)
                q = qz.quantize(w.unsqueeze(1), qz.scale, qz.zero).flatten()
                Q1[:, i] = q
                L1[:, i] = (w - q) ** 2 / d**2
                err = (w - q) / d
                # K=1 matmul in the reference (gptq.py:1298) == rounded product, then rounded subtract
                W1[:, i:] -= err.unsqueeze(1) * Hinv1[i, i:].unsqueeze(0)
                Err1[:, i] = err
            Q[:, i1:i2] = Q1
            Losses[:, i1:i2] = L1 / 2
            W[:, i2:] -= Err1.matmul(Hinv[i1:i2, i2:])
        if perm is not None:
            Q = Q[:, torch.argsort(perm)]
        if not scales:
            scales.append(qz.scale)
            zeros.append(qz.zero)
        return dict(scale=torch.cat(scales, 1), zero=torch.cat(zeros, 1), Q=Q, losses=Losses,
                    perm=perm, hinv=Hinv)

    @staticmethod
    def export_codes(Q, scale, zero, groupsize, sym: bool, perm=None):
        """gptq.py:796-813: fake-quant Q -> integer codes as the export step computes them.

        sym passes zp=None to `quant_weight_w_scale` (codes in [-2^(b-1), 2^(b-1)-1]); with act_order
        the columns are permuted before and un-permuted after the division (:797-812)."""
        Q = _cpu(Q).clone()
        if perm is not None:
            Q = Q[:, perm]
        codes = codes_from_fake_quant(Q, scale, None if sym else zero, groupsize)
        if perm is not None:
            codes = codes[:, torch.argsort(perm)]
        return codes.to(torch.int32)


# ----------------------------------------------------------------------------------------
# AWQ  (awq.py:131-154 stats, :264-361 search_scale, :393-470 search_clip)
# ----------------------------------------------------------------------------------------
def awq_weight_scale(weight, group_size=-1):
    """awq.py:131-147: mean over rows of |W| / max_group |W|."""
    w = _cpu(weight)
    shape = w.shape
    if group_size > 0:
        w = w.reshape(-1, group_size)
    s = w.abs() / w.abs().amax(dim=1, keepdim=True)
    return s.view(shape).mean(0)


def awq_act_scale(inputs: List[torch.Tensor]):
    """awq.py:151-154: mean |x| per input channel over all cached tokens."""
    return torch.cat([_cpu(x).abs().view(-1, x.shape[-1]) for x in inputs], dim=0).mean(0)


def awq_candidate_scales(x_max, w_max, ratio: float):
    """awq.py:324-325: s = clamp(x^r / w^(1-r), 1e-4); s /= sqrt(max s * min s)."""
    s = (x_max.pow(ratio) / w_max.pow(1 - ratio)).clamp(min=1e-4).view(-1)
    return s / (s.max() * s.min()).sqrt()


def awq_scaled_fake_quant(weight, scales, group_size, scheme, full_range=False):
    """awq.py:326-335: qdq(W * s) / s.  The reference passes `data_type=`/`num_bits=` which
    `quant_tensor` ignores, so the search always fake-quantises as 4-bit int."""
    w = _cpu(weight).mul(scales.view(1, -1))
    return rtn_fake_quant(w, 4, group_size, scheme, 1.0, full_range) / scales.view(1, -1)


def awq_search_scale_module(weight, bias, inputs, group_size, scheme, full_range=False):
    """awq.py:264-361 for a single-module tuple (module_inference branch).

    Returns (best_scales [K], best_ratio, loss history)."""
    weight = _cpu(weight)
    bias = _cpu(bias)
    inputs = [_cpu(x) for x in inputs]
    w_max = awq_weight_scale(weight, group_size)
    x_max = awq_act_scale(inputs)
    lin = torch.nn.functional.linear
    org = [lin(x, weight, bias) for x in inputs]
    best_err, best_s, best_r, hist = float("inf"), None, None, []
    for i in range(20):
        ratio = i * 1 / 20
        s = awq_candidate_scales(x_max, w_max, ratio)
        wq = awq_scaled_fake_quant(weight, s, group_size, scheme, full_range)
        loss = 0.0
        for o, x in zip(org, inputs):
            loss += (o - lin(x, wq, bias)).float().pow(2).mean().item()
        hist.append(loss)
        if loss < best_err:
            best_err, best_s, best_r = loss, s, ratio
    return best_s, best_r, hist


def awq_search_clip_module(weight, bias, inputs, group_size, scheme, full_range=False):
    """awq.py:393-470: 10 ratios 1 - i/100, loss = sum_samples mean((out_fp - out_q)^2)."""
    weight = _cpu(weight)
    bias = _cpu(bias)
    inputs = [_cpu(x) for x in inputs]
    lin = torch.nn.functional.linear
    org = [lin(x, weight, bias) for x in inputs]
    best_err, best, hist = float("inf"), None, []
    for i in range(int(0.1 * 100)):
        ratio = 1 - i / 100
        wq = rtn_fake_quant(weight, 4, group_size, scheme, ratio, full_range)
        loss = 0.0
        for o, x in zip(org, inputs):
            loss += (o - lin(x, wq, bias)).float().pow(2).mean().item()
        hist.append(loss)
        if loss < best_err:
            best_err, best = loss, ratio
    return best, hist


# ----------------------------------------------------------------------------------------
# SmoothQuant scale math (smooth_quant/utility.py:605-626 cal_scale, :652-755 qdq simulation)
# PINNED (round 2): the reference module hard-imports intel_extension_for_pytorch, but with IPEX stubbed
# (oracle/ref_loader.py load_smooth_quant_utility) its torch-only parts import and run on the CPU; these functions are
# checked bit for bit against `cal_scale`, `quant_dequant_w_v1`, `quant_dequant_x_v1` and `SQLinearWrapper`'s static
# qparams in tests/test_smoothquant_transform_cpu.py (fixtures tests/golden/sq_transform.pt).  What stays unpinned is the
# int8 GEMM itself: it lives in IPEX/oneDNN, outside the tree.
# ----------------------------------------------------------------------------------------
def sq_cal_scale(input_max_abs, weights: List[torch.Tensor], alpha: float):
    """smooth_quant/utility.py:605-626."""
    w = torch.cat([_cpu(x) for x in weights], dim=0)
    weight_max = torch.clip(torch.max(torch.abs(w), dim=0)[0], 1e-5)
    input_power = torch.pow(_cpu(input_max_abs), alpha)
    weight_power = torch.pow(weight_max, 1 - alpha)
    scale = torch.clip(input_power / weight_power, min=1e-5)
    scale[input_power == 0] = 1.0
    return scale


def sq_qdq_weight_per_channel(w, bits=8):
    """smooth_quant/utility.py:652-690 (`quant_dequant_w_v1`, nn.Linear, sym): scale = max|w_n| / 127.5,
    clipped to fp32 eps; q = clamp(round(w/scale), -128, 127). Returns (qdq, q, scale[N,1])."""
    w = _cpu(w)
    eps = torch.finfo(torch.float32).eps
    q_min, q_max = -(2.0 ** (bits - 1)), 2.0 ** (bits - 1) - 1.0
    x_max = torch.max(torch.abs(w), dim=1).values
    scale = torch.clip(x_max / (float(q_max - q_min) / 2), min=eps).unsqueeze(-1)
    q = torch.round(w / scale).clamp_(q_min, q_max)
    return q * scale, q, scale


def sq_qdq_act_per_tensor(x, min_x=None, max_x=None, bits=8):
    """smooth_quant/utility.py:726-755 (`quant_dequant_x_v1`, asym uint8 per-tensor).
    Returns (qdq, q, scale, zero_point)."""
    x = _cpu(x)
    eps = torch.finfo(torch.float32).eps
    q_min, q_max = 0, 2.0**bits - 1.0
    if max_x is None or min_x is None:
        max_x, min_x = torch.max(x), torch.min(x)
    else:
        max_x, min_x = torch.max(_cpu(max_x)), torch.min(_cpu(min_x))
    scale = torch.clip((max_x - min_x) / (2**bits - 1), min=eps)
    bias = torch.round((0 - min_x) / scale)
    q = torch.round(x / scale + bias).clamp_(q_min, q_max)
    return scale * (q - bias), q, scale, bias


def sq_w8a8_linear(x, W, smooth, act_min, act_max, bias=None, qparams=None):
    """`SQLinearWrapper` + the W8A8 QDQ simulation the reference falls back to without IPEX
    (smooth_quant/utility.py:2559-2662 wrapper, :2607-2631 static activation qparams, :652-690 / :726-755 QDQ,
    `WrapperLayer.q_dq_forward` :2707-2729), in fp32 torch-CPU ops:

        W' = W * smooth ; x' = x * (1 / smooth)
        q_w, s_w = per-out-channel sym int8 of W' ; q_x = clamp(round(x' / s_x + zp_x), 0, 255) with the STATIC
        (calibrated) per-tensor range of x' ; y = (s_x * (q_x - zp_x)) @ (q_w * s_w)^T + bias

    `qparams` = (input_scale, s_x, zp_x) overrides the locally derived activation parameters (a device reciprocal may
    differ from the host's in the last bit; the GEMM check wants identical codes on both sides).

    Returns dict(y, q_w int, s_w [N,1], q_x, s_x, zp_x).  The quantisation parameters are pinned against the live
    reference; the integer GEMM of IPEX has no reference here."""
    x, W, smooth = _cpu(x).float(), _cpu(W).float(), _cpu(smooth).float()
    input_scale = 1.0 / smooth
    Ws = W * smooth.view(1, -1)
    _, q_w, s_w = sq_qdq_weight_per_channel(Ws, 8)
    eps = torch.finfo(torch.float32).eps
    mn = torch.clamp((_cpu(act_min).float() * input_scale).min(), max=0.0)
    mx = torch.clamp((_cpu(act_max).float() * input_scale).max(), min=0.0)
    s_x = torch.clip((mx - mn) / 255.0, min=eps)
    zp_x = torch.clamp(torch.round((0 - mn) / s_x), 0, 255)
    if qparams is not None:
        input_scale, s_x, zp_x = (_cpu(t).float() for t in qparams)
        s_x, zp_x = s_x.reshape(()), zp_x.reshape(())
    xs = x * input_scale
    q_x = torch.round(xs / s_x + zp_x).clamp_(0, 255)
    y = torch.nn.functional.linear(s_x * (q_x - zp_x), q_w * s_w, None if bias is None else _cpu(bias).float())
    return dict(y=y, q_w=q_w, s_w=s_w, q_x=q_x, s_x=s_x, zp_x=zp_x)
