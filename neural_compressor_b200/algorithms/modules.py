"""Packed weight-only modules.

Reference: neural_compressor/torch/algorithms/weight_only/modules.py
    WeightOnlyLinear (ABC)     :91-154
    INCWeightOnlyLinear        :157-627   (optimum format buffers :236-262, pack :321-375, unpack :377-411,
                                           recover :413-443, forward :594-610)
    MulLinear                  :907-949

`B200WeightOnlyLinear` keeps the reference's buffer names, dtypes and shapes (`qweight`, `qzeros`, `scales`,
`bias`, `g_idx`, `scale_bf16_to_fp8`) so the reference's save/load (`WOQModelLoader`) and HF-format export
consume it unchanged; pack / unpack / recover / forward are sm_100a kernels behind the C ABI.
"""
from __future__ import annotations

import math
import os
from abc import abstractmethod

import torch

from .. import ops


# Batches up to this many rows take the per-warp TMA-ring kernel on the derived stream layout (woq_stream.cu); larger
# batches amortise the weight read over more rows and use the cluster split-K kernel on the optimum tensors.
STREAM_MAX_ROWS = int(os.environ.get("B200WOQ_STREAM_MAX_ROWS", "4"))
# Above this many rows the fused dequant-GEMM kernels (tuned for the HBM-bound decode regime) hand over to
# dequantize + cuBLAS.
GEMM_MAX_ROWS = int(os.environ.get("B200WOQ_GEMM_MAX_ROWS", "128"))
STREAM_FLAGS = 2 if os.environ.get("B200WOQ_PDL", "1") != "0" else 0  # bit 1: programmatic dependent launch


class WeightOnlyLinear(torch.nn.Module):
    """modules.py:91-154."""

    def __init__(self, in_features, out_features, dtype, bits, group_size, device, scale_dtype=torch.float32, **kwargs):
        super().__init__()
        self.dtype = dtype
        self.bits = bits
        self.group_size = group_size if group_size != -1 else in_features
        self.in_features = in_features
        self.out_features = out_features
        self.device = device
        self.scale_dtype = scale_dtype

    @abstractmethod
    def pack(self, *args, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def unpack(self, *args, **kwargs):
        raise NotImplementedError

    def extra_repr(self) -> str:
        return "in_features={}, out_features={}, bits={}, group_size={}, bias={}".format(
            self.in_features, self.out_features, self.bits, self.group_size, self.bias is not None)


class B200WeightOnlyLinear(WeightOnlyLinear):
    """Drop-in for `INCWeightOnlyLinear(use_optimum_format=True)` with the kernels on the B200."""

    def __new__(cls, in_features=None, out_features=None, dtype="int", *args, **kwargs):
        """The reference's class covers two layouts behind one constructor.  The optimum format (the default, the only
        one GPTQ / AWQ / int RTN produce and the one the dequant-GEMM kernels read) is this class; the non-optimum
        layout -- `use_optimum_format=False`, which the table data types nf4 / fp4 force (modules.py:214-222) -- is
        `modules_rowmajor.B200RowMajorLinear`, returned from here so that callers keep the reference's constructor."""
        if cls is B200WeightOnlyLinear and ("int" not in str(dtype) or not kwargs.get("use_optimum_format", True)):
            from .modules_rowmajor import B200RowMajorLinear

            return B200RowMajorLinear(in_features, out_features, dtype, *args, **kwargs)
        return super().__new__(cls)

    def __init__(self, in_features, out_features, dtype="int", bits=4, group_size=32, zp=False, bias=False,
                 scale_dtype=torch.float32, compression_dtype=torch.int32, compression_dim=1, g_idx=False,
                 device="cuda", use_optimum_format=True, **kwargs):
        super().__init__(in_features, out_features, dtype, bits, group_size, device, scale_dtype=scale_dtype)
        # `compression_dtype` / `compression_dim` / `scale_dtype` are ignored in the optimum format, as in the reference
        # (modules.py:241-262 forces int32 words packed along K and fp16 scales)
        assert "int" in str(dtype) and use_optimum_format, "routed to B200RowMajorLinear by __new__"
        self.use_optimum_format = True
        self.compression_dtype = torch.int32
        self.float_type = torch.float16
        self.compress_bits = 32
        self.n_pack = 32 // bits
        ng = math.ceil(in_features / self.group_size)
        dev = device
        self.register_buffer("scale_bf16_to_fp8", torch.zeros(1, dtype=torch.bfloat16, device=dev))
        self.register_buffer("scales", torch.zeros((ng, out_features), dtype=torch.float16, device=dev))
        self.register_buffer("qweight", torch.zeros((math.ceil(in_features / self.n_pack), out_features),
                                                    dtype=torch.int32, device=dev))
        self.register_buffer("qzeros", torch.zeros((ng, math.ceil(out_features / self.n_pack)),
                                                   dtype=torch.int32, device=dev))
        self.register_buffer("bias", torch.zeros(out_features, dtype=torch.float16, device=dev))
        if g_idx:
            self.register_buffer("g_idx", torch.zeros(in_features, dtype=torch.int32, device=dev))
        else:
            self.g_idx = None

    # the packed tensors have FIXED storage dtypes (they are an on-disk contract and the kernels reinterpret raw
    # pointers): `model.to(torch.bfloat16)` / `.half()` / `.float()` must move them between devices but never re-type
    # `scales` (fp16), `bias` (fp16) or `scale_bf16_to_fp8` (bf16) -- a bf16 round trip would also destroy scale bits
    _FIXED_DTYPE = ("scales", "bias", "scale_bf16_to_fp8")

    def _apply(self, fn, recurse=True):
        keep = {n: self._buffers[n] for n in self._FIXED_DTYPE if isinstance(self._buffers.get(n), torch.Tensor)}
        super()._apply(fn, recurse)
        for n, old in keep.items():
            new = self._buffers[n]
            if new.dtype != old.dtype:
                self._buffers[n] = old.to(new.device)
        return self

    # ------------------------------------------------------------------ pack
    def pack(self, int_weight, scales, zp, bias, scale_bf16_to_fp8=None, g_idx=None, **kwargs):
        """modules.py:321-375.  int_weight [N,K] integer-valued (sym: in [-2^(b-1), 2^(b-1)-1], zp None);
        scales [N,G]; zp [N,G] or None."""
        dev = self.qweight.device
        ng = self.scales.shape[0]
        assert tuple(scales.shape) == (self.out_features, ng), \
            f"{tuple(scales.shape)} != {(self.out_features, ng)} Scale shape is mismatched."
        codes = int_weight.to(dev).to(torch.int32)
        if zp is None:
            codes = codes + 2 ** (self.bits - 1)
        codes = (codes & (2**self.bits - 1)).to(torch.uint8).contiguous()
        self.pack_stored(codes, scales.to(dev), None if zp is None else zp.to(dev), bias, g_idx)

    def pack_stored(self, stored_codes, scales, zp, bias, g_idx=None):
        """Pack codes that already are the unsigned stored fields (uint8 [N,K])."""
        self.qweight = ops.pack_codes(stored_codes, self.bits)
        self.scales, self.qzeros = ops.pack_params(scales, zp, self.bits)
        self._set_bias_gidx(bias, g_idx)

    def set_packed(self, qweight, qzeros, scales16, bias, g_idx=None):
        self.qweight, self.qzeros, self.scales = qweight, qzeros, scales16
        self._set_bias_gidx(bias, g_idx)

    def _set_bias_gidx(self, bias, g_idx):
        dev = self.qweight.device
        if bias is not None:
            self.bias = bias.detach().to(dev).to(self.float_type)
        if g_idx is not None:
            assert hasattr(self, "g_idx"), "g_idx is not set when initializing."
            perm = g_idx.to(dev).to(torch.int64)
            # optimum format stores the group of each input channel (modules.py:338-344)
            self.g_idx = (torch.argsort(perm) // self.group_size).to(torch.int32)

    # ------------------------------------------------------------------ unpack / recover / forward
    def unpack(self):
        """modules.py:377-411 -> dict(int_weight, scales [N,G], zp, g_idx, bias)."""
        ng = self.scales.shape[0]
        codes, zps = ops.unpack(self.qweight, self.qzeros, self.bits, self.in_features, self.out_features, ng)
        return dict(int_weight=codes.to(torch.int16), scales=self.scales.t().contiguous(),
                    scale_bf16_to_fp8=self.scale_bf16_to_fp8, zp=zps.to(torch.int16), g_idx=self.g_idx, bias=self.bias)

    def recover(self):
        """modules.py:413-443: fp16 [N,K] = fp16(int8(q - zp) * scale)."""
        return ops.dequantize(self.qweight, self.qzeros, self.scales, self.bits, self.group_size, self.in_features,
                              self.out_features, self.g_idx)

    def _packed_key(self):
        """Identity + version of the packed tensors: an in-place load_state_dict keeps the pointers but bumps versions."""
        return tuple(v for t in (self.qweight, self.qzeros, self.scales) for v in (t.data_ptr(), t._version))

    def _stream_layout(self):
        """Lazily derived copy of the packed tensors in the B200 stream layout (not part of the state_dict)."""
        key = self._packed_key()
        if getattr(self, "_stream_key", None) != key:
            self._stream = ops.build_stream_layout(self.qweight, self.qzeros, self.scales, self.bits, self.group_size,
                                                   self.in_features, self.out_features)
            self._stream_key = key
            # the launch right after the build must not use programmatic dependent launch: the GEMV prefetches the
            # layout before `griddepcontrol.wait`, which is only legal for data no in-flight kernel is still writing
            self._stream_fresh = True
        return self._stream

    def _stream_flags(self):
        fresh, self._stream_fresh = getattr(self, "_stream_fresh", False), False
        return 0 if fresh else STREAM_FLAGS

    def forward(self, input, input_scale=None):
        out_dtype = input.dtype if input.dtype in (torch.float16, torch.bfloat16) else torch.float32
        rows = input.numel() // self.in_features
        # The small-batch tensor-core kernels stage the activations as fp16 (the reference itself casts the input to the
        # fp16 weight's dtype on an accelerator, modules.py:606).  bf16 / fp32 activations can exceed 65504, so they take
        # the fp32-math kernel (rows <= 8) or recover() + a library GEMM in the input dtype, like the reference's CPU path.
        fp16_in = input.dtype == torch.float16
        if (rows <= STREAM_MAX_ROWS and fp16_in and self.bits == 4 and self.g_idx is None and self.qweight.is_cuda
                and os.environ.get("B200WOQ_STREAM", "1") != "0"):
            group = getattr(self, "_siblings", None)
            if group is not None and rows == 1 and input_scale is None:
                return group.forward(self, input, out_dtype)
            layout = self._stream_layout()
            if layout is not None:
                return ops.woq_linear_stream(input, layout, self.bias, self.bits, self.group_size, self.in_features,
                                             self.out_features, input_scale=input_scale, out_dtype=out_dtype,
                                             flags=self._stream_flags())
        if rows > GEMM_MAX_ROWS or (not fp16_in and rows > 8):
            # prefill / calibration batches are compute-bound: recover the fp16 weight once (K4 dequantize kernel) and
            # run a plain library GEMM in the input's dtype -- literally the reference's forward (modules.py:594-610:
            # recover() then F.linear), so calibration forwards through already-packed modules keep fp32 activations
            w = self.recover().to(input.dtype)
            x = input if input_scale is None else input * input_scale.to(input.dtype)
            return torch.nn.functional.linear(x, w, None if self.bias is None else self.bias.to(input.dtype))
        return ops.woq_linear(input, self.qweight, self.qzeros, self.scales, self.bias, self.bits, self.group_size,
                              self.in_features, self.out_features, g_idx=self.g_idx, input_scale=input_scale,
                              out_dtype=out_dtype, flags=0 if fp16_in else 1)  # bit 0: general fp32-math kernel


class SiblingGroup:
    """Linears of one parent that read the same activation (q/k/v, gate/up).  At batch 1 a dequant-GEMV over a
    4096x4096 layer is dominated by the ~2 us launch-to-launch dependency latency, not by its 8.7 MB of weights, so the
    first member called with an input runs ONE launch over the concatenated column strips of all members (the stream
    layout is strip-major, so concatenating the members' layouts is the fused layout) and the other members pick up
    their slice when they are called with the very same tensor object.  Anything else (another tensor, a modified
    tensor, batch > 1, an input scale) falls back to the member's own launch."""

    def __init__(self, members):
        self.members = list(members)
        self.offsets = []
        n = 0
        for m in self.members:
            self.offsets.append(n)
            n += m.out_features
        self.out_features = n
        self._key = None
        self._x = None

    def _fused(self):
        key = tuple(v for m in self.members for t in (m.qweight, m.qzeros, m.scales, m.bias) for v in (t.data_ptr(), t._version))
        if self._key != key:
            parts = [ops.build_stream_layout(m.qweight, m.qzeros, m.scales, m.bits, m.group_size, m.in_features,
                                             m.out_features) for m in self.members]
            if any(p is None for p in parts):
                self.layout = None
            else:
                self.layout = torch.cat(parts)
                self.bias = torch.cat([m.bias.to(torch.float16) for m in self.members])
                o = 0
                for m, p in zip(self.members, parts):  # members alias their slice: no second copy of the weights
                    m._stream = self.layout[o:o + p.numel()]
                    m._stream_key = m._packed_key()
                    m._stream_fresh = True
                    o += p.numel()
            self._key = key
            self._fresh = True
        return self.layout

    def forward(self, member, x, out_dtype):
        idx = self.members.index(member)
        if self._x is x and self._ver == x._version and self._dtype == out_dtype and idx in self._pending:
            y = self._y[..., self.offsets[idx]:self.offsets[idx] + member.out_features]
            self._pending.discard(idx)
            if not self._pending:
                self._x = self._y = None
            return y
        layout = self._fused()
        m0 = self.members[0]
        if layout is None:
            return ops.woq_linear(x, member.qweight, member.qzeros, member.scales, member.bias, member.bits,
                                  member.group_size, member.in_features, member.out_features, out_dtype=out_dtype)
        fresh, self._fresh = self._fresh, False
        y = ops.woq_linear_stream(x, layout, self.bias, m0.bits, m0.group_size, m0.in_features, self.out_features,
                                  out_dtype=out_dtype, flags=0 if fresh else STREAM_FLAGS)
        self._x, self._ver, self._dtype, self._y = x, x._version, out_dtype, y
        self._pending = set(range(len(self.members))) - {idx}
        return y[..., self.offsets[idx]:self.offsets[idx] + member.out_features]


SIBLING_PATTERNS = (("q_proj", "k_proj", "v_proj"), ("gate_proj", "up_proj"), ("w1", "w3"))


def fuse_sibling_linears(model, patterns=SIBLING_PATTERNS):
    """Attach a SiblingGroup to every parent whose children match one of `patterns` (attention q/k/v, gated-MLP
    gate/up) and are packed 4-bit modules of equal in_features / group_size without g_idx.  Returns the number of
    groups.  Disabled by B200WOQ_FUSE_SIBLINGS=0."""
    if os.environ.get("B200WOQ_FUSE_SIBLINGS", "1") == "0":
        return 0
    n = 0
    for parent in model.modules():
        kids = dict(parent.named_children())
        for pat in patterns:
            ms = [kids.get(k) for k in pat]
            if not all(isinstance(m, B200WeightOnlyLinear) for m in ms):
                continue
            m0 = ms[0]
            if any(m.bits != 4 or m.g_idx is not None or m.in_features != m0.in_features or m.group_size != m0.group_size
                   or m.out_features % 32 or not m.qweight.is_cuda for m in ms):
                continue
            group = SiblingGroup(ms)
            for m in ms:
                object.__setattr__(m, "_siblings", group)
            n += 1
    return n


class MulLinear(torch.nn.Module):
    """modules.py:907-949: linear(x * input_scale).  When the wrapped module is packed, the multiply is fused
    into the dequant-GEMM's activation staging."""

    def __init__(self, module, input_scale=None):
        super().__init__()
        if input_scale is None:
            input_scale = torch.empty(module.in_features)
        self.register_buffer("input_scale", input_scale)
        self.add_module("linear", module)

    @property
    def weight(self):
        return self.linear.weight

    @weight.setter
    def weight(self, weight):
        self.linear.weight = weight

    def forward(self, X):
        if isinstance(self.linear, B200WeightOnlyLinear):
            return self.linear(X, input_scale=self.input_scale.float())
        return self.linear(torch.mul(X, self.input_scale.to(X.dtype)))  # statistics are fp32 here; keep the model's dtype

    def _update_linear(self):
        scale = self.input_scale.view(1, self.input_scale.shape[0])
        with torch.no_grad():
            self.linear.weight /= scale

    def _recover_linear(self):
        scale = self.input_scale.view(1, self.input_scale.shape[0])
        with torch.no_grad():
            self.linear.weight *= scale
