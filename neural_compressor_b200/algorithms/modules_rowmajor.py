"""`INCWeightOnlyLinear(use_optimum_format=False)`: the reference's second on-disk layout.

Reference: neural_compressor/torch/algorithms/weight_only/modules.py
    buffers :263-314 (scales [N, G] in scale_dtype; qweight / qzeros packed along `compression_dim` into
             `compression_dtype` words), pack :321-375, unpack :377-411, recover :413-443,
    pack_tensor / unpack_tensor :445-592 (field e of a word = (v & mask) << bits*e; unpacking sign-extends a field unless
             the module has zero points), forward :594-610.
    Table data types (nf4 / fp4 / fp4_e2m1) always use this layout (:214-222): qweight int32 [N, ceil(K/8)] holding the
    signed 4-bit codes of INT_MAPPING, fp32 scales, no zero points.

Who produces it: RTN with a table dtype (algorithms/rtn.py -> float4.cu kernels: quantise, pack, recover), and direct
construction with `use_optimum_format=False` / `compression_dtype` in {int8, int16, int32, int64} / `compression_dim` in
{0, 1} (the reference's module test, test_woq_module.py:10-52).  The word packing of the general int8/16/32/64 x dim 0/1
formats is integer tensor arithmetic (shift / and / sum of disjoint fields) on whatever device the tensors live on -- a
pure storage permutation without a hot path -- while `forward` is, as in the reference, `recover()` followed by a dense
GEMM in the recovered weight's dtype.
"""
from __future__ import annotations

import math

import torch

from .. import dtypes, ops
from .modules import WeightOnlyLinear

_CBITS = {torch.int8: 8, torch.int16: 16, torch.int32: 32, torch.int64: 64}


def pack_fields(values: torch.Tensor, bits: int, cdtype: torch.dtype) -> torch.Tensor:
    """[R, C] integer-valued tensor -> [R, ceil(C / n_pack)] words of `cdtype`; field e of word j is
    `values[:, j*n_pack + e] & mask` at bit `bits*e` (modules.py:445-466, 519-534)."""
    cbits = _CBITS[cdtype]
    n_pack = cbits // bits
    rows, cols = values.shape
    words = math.ceil(cols / n_pack)
    v = values.to(torch.int64) & ((1 << bits) - 1)
    if words * n_pack != cols:
        v = torch.nn.functional.pad(v, (0, words * n_pack - cols))
    shifts = torch.arange(n_pack, device=v.device, dtype=torch.int64) * bits
    word = (v.view(rows, words, n_pack) << shifts).sum(dim=-1)     # disjoint fields: sum == or (mod 2^64)
    if cbits < 64:   # two's complement wrap into the storage type
        half = 1 << (cbits - 1)
        word = ((word + half) % (1 << cbits)) - half
    return word.to(cdtype)


def unpack_fields(words: torch.Tensor, bits: int, signed: bool) -> torch.Tensor:
    """Inverse of `pack_fields` -> int16 [R, W * n_pack]; a field is sign-extended when `signed` (a module without zero
    points stores signed codes, modules.py:477-486)."""
    cbits = _CBITS[words.dtype]
    n_pack = cbits // bits
    shifts = torch.arange(n_pack, device=words.device, dtype=torch.int64) * bits
    f = (words.to(torch.int64).unsqueeze(-1) >> shifts) & ((1 << bits) - 1)
    if signed:
        sign = 1 << (bits - 1)
        f = (f ^ sign) - sign
    return f.reshape(words.shape[0], -1).to(torch.int16)


class B200RowMajorLinear(WeightOnlyLinear):
    """`B200WeightOnlyLinear(..., use_optimum_format=False)` and every table-dtype module construct this class."""

    def __init__(self, in_features, out_features, dtype="int", bits=4, group_size=32, zp=False, bias=False,
                 scale_dtype=torch.float32, compression_dtype=torch.int32, compression_dim=1, g_idx=False,
                 device="cuda", use_optimum_format=True, **kwargs):
        super().__init__(in_features, out_features, dtype, bits, group_size, device, scale_dtype=scale_dtype)
        assert compression_dtype in _CBITS, \
            f"Only support torch.int8|16|32|64 as compressed dtype. but got {compression_dtype}"
        assert compression_dim in (0, 1), \
            "Only support 0 or 1 as compression dimension, 0 is output channel, 1 is input channel."
        self.is_table = dtypes.is_table_dtype(dtype)
        if self.is_table and bits != 4:
            raise ValueError(f"{dtype} is a 4-bit data type, got bits={bits}")
        self.use_optimum_format = False
        self.compression_dtype, self.compression_dim = compression_dtype, compression_dim
        self.compress_bits = _CBITS[compression_dtype]
        self.n_pack = self.compress_bits // bits
        self.float_type = scale_dtype
        n, k = out_features, in_features
        ng = math.ceil(k / self.group_size)
        dev = device
        self.register_buffer("scale_bf16_to_fp8", torch.zeros(1, dtype=torch.bfloat16, device=dev))
        self.register_buffer("scales", torch.zeros((n, ng), dtype=scale_dtype, device=dev))
        if compression_dim == 1:
            qw_shape, qz_shape = (n, math.ceil(k / self.n_pack)), (n, math.ceil(k / self.group_size / self.n_pack))
        else:
            qw_shape, qz_shape = (math.ceil(n / self.n_pack), k), (math.ceil(n / self.n_pack), ng)
        self.register_buffer("qweight", torch.zeros(qw_shape, dtype=compression_dtype, device=dev))
        if zp:
            self.register_buffer("qzeros", torch.zeros(qz_shape, dtype=compression_dtype, device=dev))
        if bias:
            self.register_buffer("bias", torch.zeros(n, dtype=scale_dtype, device=dev))
        else:
            self.bias = None
        if g_idx:
            self.register_buffer("g_idx", torch.zeros(k, dtype=torch.int32, device=dev))
        else:
            self.g_idx = None

    # packed words / scales are an on-disk contract: `.half()` / `.to(bfloat16)` move them but never re-type them
    def _apply(self, fn, recurse=True):
        keep = {n: b for n, b in self._buffers.items() if isinstance(b, torch.Tensor)}
        super()._apply(fn, recurse)
        for n, old in keep.items():
            new = self._buffers[n]
            if new.dtype != old.dtype:
                self._buffers[n] = old.to(new.device)
        return self

    # ------------------------------------------------------------------ pack
    def _along(self, t):
        """Orient a [N, *] tensor so that the packed axis is the last one."""
        return t if self.compression_dim == 1 else t.t()

    def pack(self, int_weight, scales, zp, bias, scale_bf16_to_fp8=None, g_idx=None, **kwargs):
        """modules.py:321-375 for the non-optimum layout: int_weight [N, K] (signed codes when zp is None), scales
        [N, G], zp [N, G] or None."""
        dev = self.qweight.device
        assert tuple(scales.shape) == tuple(self.scales.shape), \
            f"{scales.shape} != {self.scales.shape} Scale shape is mismatched."
        if zp is not None:
            assert hasattr(self, "qzeros"), "zp is not set when initializing."
        self.scales = scales.to(dev).to(self.float_type)
        if scale_bf16_to_fp8 is not None:
            self.scale_bf16_to_fp8 = scale_bf16_to_fp8.to(dev).to(self.float_type)
        if bias is not None:
            assert self.bias is not None, "bias is not set when initializing."
            self.bias = bias.detach().to(dev).to(self.float_type)
        if g_idx is not None:
            assert self.g_idx is not None, "g_idx is not set when initializing."
            g = g_idx.to(dev).to(torch.int32)
            if self.is_table:   # modules.py:340-343: table dtypes store the group of each input channel
                g = (torch.argsort(g.to(torch.int64)) // self.group_size).to(torch.int32)
            self.g_idx = g
        w = self._along(int_weight.to(dev))
        expect = tuple(self.qweight.shape)
        self.qweight = self._along(pack_fields(w, self.bits, self.compression_dtype)).contiguous()
        assert tuple(self.qweight.shape) == expect, "output channels mismatch, please check."
        if zp is not None and hasattr(self, "qzeros"):
            z = self._along(zp.to(dev))
            expect = tuple(self.qzeros.shape)
            self.qzeros = self._along(pack_fields(z, self.bits, self.compression_dtype)).contiguous()
            assert tuple(self.qzeros.shape) == expect

    def set_packed(self, qweight, scales, bias, qzeros=None):
        """Adopt tensors that already are in this module's layout (RTN's kernels write them directly)."""
        assert tuple(qweight.shape) == tuple(self.qweight.shape) and qweight.dtype == self.compression_dtype
        assert tuple(scales.shape) == tuple(self.scales.shape)
        self.qweight, self.scales = qweight, scales.to(self.float_type)
        if qzeros is not None:
            self.qzeros = qzeros
        if bias is not None:
            self.bias = bias.detach().to(qweight.device).to(self.float_type)

    # ------------------------------------------------------------------ unpack / recover / forward
    def unpack(self):
        """modules.py:377-411 -> dict(int_weight, scales [N, G], zp, g_idx, bias); table dtypes come back as levels."""
        signed = not hasattr(self, "qzeros")
        w = self._along(unpack_fields(self._along(self.qweight).contiguous(), self.bits, signed))
        w = w[: self.out_features, : self.in_features].contiguous()
        if self.is_table:
            lut = torch.tensor(list(dtypes.nibble_levels(self.dtype)), dtype=torch.float32, device=w.device)
            w = lut[(w.to(torch.int64) & 0xF)]
        zp = None
        if hasattr(self, "qzeros"):
            zp = self._along(unpack_fields(self._along(self.qzeros).contiguous(), self.bits, False))
            zp = zp[: self.scales.shape[0], : self.scales.shape[1]].contiguous()
        return dict(int_weight=w, scales=self.scales, scale_bf16_to_fp8=self.scale_bf16_to_fp8, zp=zp, g_idx=self.g_idx,
                    bias=self.bias)

    def _kernel_recover_ok(self):
        return (self.is_table and self.qweight.is_cuda and self.compression_dtype == torch.int32
                and self.compression_dim == 1 and self.float_type == torch.float32 and self.g_idx is None)

    def recover(self):
        """modules.py:413-443: W[n, k] = (code - zp) * scale in `scale_dtype` (table dtypes: level * scale)."""
        if self._kernel_recover_ok():
            return ops.f4_dequantize(self.qweight, self.scales, self.dtype, self.group_size, self.in_features)
        u = self.unpack()
        w, zp = u["int_weight"], u["zp"]
        if self.g_idx is None:
            group = torch.arange(self.in_features, device=w.device) // self.group_size
        else:
            group = self.g_idx.to(torch.int64)
        scales = self.scales[:, group]
        if zp is not None:   # `.to(torch.int8)` wraps like the reference (modules.py:435)
            w = (w.to(torch.int32) - zp.to(torch.int32)[:, group]).to(torch.int8)
        return (w * scales).to(self.float_type)

    def forward(self, input, input_scale=None):
        """modules.py:594-610: recover the weight and run a dense GEMM in its dtype.  The result is handed back in the
        activation's dtype when that is a different floating type (the reference returns the weight's dtype, which a
        half-precision model cannot consume)."""
        w = self.recover()
        x = input if input_scale is None else input * input_scale.to(input.dtype)
        y = torch.nn.functional.linear(x.to(w.dtype), w, self.bias)
        return y.to(input.dtype) if input.is_floating_point() and input.dtype != y.dtype else y

    def extra_repr(self) -> str:
        return super().extra_repr() + f", dtype={self.dtype}, compression_dtype={self.compression_dtype}, " \
                                      f"compression_dim={self.compression_dim}"
