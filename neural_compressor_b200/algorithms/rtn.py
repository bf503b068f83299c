"""RTN quantizer.  Reference: neural_compressor/torch/algorithms/weight_only/rtn.py:45-270
(RTNQuantizer.convert) and utility.py:439-480 (search_clip).

Host Python walks the modules exactly like the reference; the per-layer work -- group min/max, scale and
zero-point, rounding, bit packing -- is the K4 kernels.  The weights are quantised where the kernels live:
each Linear's weight is moved to the B200, packed there, and the packed module is placed on the model's
device (`m.to(device)` / `new_module.to(model_device)`, rtn.py:164-165, 262-264).
"""
from __future__ import annotations

import torch

from .. import dtypes, ops
from ..utils import current_device, get_model_device, logger, set_module
from .base_algorithm import Quantizer
from .modules import B200WeightOnlyLinear


def _supported_layers():
    types = [torch.nn.Linear]
    try:
        import transformers

        types.append(transformers.Conv1D)
    except Exception:  # pragma: no cover
        pass
    return tuple(types)


def search_clip(weight: torch.Tensor, bits=4, group_size=32, scheme="asym", enable_full_range=False, dtype="int") -> float:
    """utility.py:439-480: 40 ratios 1 - i/200, loss = mean((W - qdq(W))^2) over the whole tensor."""
    best_err, best = float("inf"), None
    tmp = torch.empty_like(weight)
    for i in range(int(0.2 * 200)):
        ratio = 1 - i / 200
        if dtypes.is_table_dtype(dtype):
            ops.f4_quantize(weight, dtype, group_size, ratio, want_codes=False, fake_out=tmp)
        else:
            ops.rtn_fake_quant(weight, bits, group_size, scheme == "sym", enable_full_range, ratio, out=tmp)
        loss = (weight - tmp).float().pow(2).mean()
        if loss < best_err:
            best_err, best = loss, ratio
    return best


def cast_fp8(weight: torch.Tensor, dtype: str) -> torch.Tensor:
    """utility.py:163-172 (`use_qdq=True`): round the weight through an fp8 storage type in place; no scaling, the module
    stays a Linear.  A pure dtype round trip on the device."""
    fp8 = getattr(torch, dtype.replace("fp8", "float8"))
    weight.copy_(weight.to(fp8).to(weight.dtype))
    return weight


def double_quant_scales(scale: torch.Tensor, w_dtype: torch.dtype, cfg: dict) -> torch.Tensor:
    """quant_tensor's second level (utility.py:377-434): the [N, G] scales, flattened row-major to one row, are
    fake-quantised in groups of `double_quant_group_size` (symmetric `double_quant_bits`-bit; the "asym" scheme first
    subtracts the mean of all scales and adds it back).  Arithmetic runs in the weight's dtype, like the reference's
    `scale` tensor.  `scale` is the fp32 [N, G] tensor of the K4 kernels (holding exactly the weight-dtype values)."""
    if str(cfg.get("double_quant_dtype", "int")) != "int":
        raise NotImplementedError("double quant of scales: int dtype only")
    flat = scale.to(w_dtype).reshape(1, -1).contiguous()
    asym = cfg.get("double_quant_scheme", "asym") == "asym"
    if asym:
        mean = flat.mean()
        flat.sub_(mean)
    ops.rtn_fake_quant(flat, int(cfg.get("double_quant_bits", 8)), int(cfg.get("double_quant_group_size", 256)), True, False,
                       1.0, out=flat)
    if asym:
        flat.add_(mean)
    return flat.reshape(scale.shape).float()


class RTNQuantizer(Quantizer):
    def __init__(self, quant_config=None):
        super().__init__(quant_config if quant_config is not None else {})

    def prepare(self, model, *args, **kwargs):
        """rtn.py:57-65: RTN needs no calibration."""
        return model

    @torch.no_grad()
    def convert(self, model, dtype="int", bits=4, scheme="sym", group_size=32, group_dim=1, quantile=1.0,
                use_full_range=False, use_mse_search=False, use_layer_wise=False, model_path="", quant_lm_head=False,
                *args, **kwargs):
        weight_config = self.quant_config
        device = current_device()
        model_device = get_model_device(model)
        assert isinstance(model, torch.nn.Module), "only support torch module"
        if use_layer_wise:
            logger.info("use_layer_wise is a CPU-RAM saving mode of the reference; weights are streamed to the "
                        "B200 layer by layer anyway, so it is ignored here.")
        supported = _supported_layers()
        try:
            import transformers

            conv1d = transformers.Conv1D
        except Exception:  # pragma: no cover
            conv1d = ()
        for name, m in list(model.named_modules()):
            if not isinstance(m, supported) or name not in weight_config:
                continue
            cfg = weight_config[name]
            dtype = cfg.get("dtype", "int")
            if dtype == "fp32":
                continue
            if dtype in dtypes.FP8_DTYPES:   # rtn.py:166-170: qdq cast, the module stays a Linear on the device
                m.to(device)
                cast_fp8(m.weight.data, dtype)
                continue
            bits = cfg.get("bits", 4)
            if dtype != "int" and "int" in dtype:
                bits = int(dtype.lstrip("int"))
                dtype = "int"
            group_size = cfg["group_size"]
            scheme = cfg["scheme"]
            quantile = cfg.get("quantile", 1.0)
            group_dim = cfg.get("group_dim", 1)
            use_full_range = cfg.get("use_full_range", False)
            use_mse_search = cfg.get("use_mse_search", False)
            double_quant = bool(cfg.get("use_double_quant", False))
            is_conv1d = bool(conv1d) and isinstance(m, conv1d)
            transpose = (group_dim == 0) ^ is_conv1d   # rtn.py:209-216
            if group_dim == 0 and not is_conv1d:
                # the reference cannot run this either: its scales come out as [N/g, K] and INCWeightOnlyLinear.pack
                # asserts on the shape (modules.py:345; verified on the live reference for square and non-square layers)
                raise AssertionError("group_dim=0 on nn.Linear: Scale shape is mismatched.")
            w = m.weight.detach().to(device)
            w = w.t().contiguous() if transpose else w.contiguous()   # [N = out, K = in]
            if use_mse_search:
                quantile = search_clip(w, bits, group_size, scheme, use_full_range, dtype)
            k = w.shape[1]
            if double_quant and k % (k if group_size == -1 or k < group_size else group_size) != 0:
                double_quant = False   # quant_tensor returns from its ragged-tail branch before the second level (utility.py:334-373)
            if is_conv1d:
                in_features, out_features = m.weight.shape[0], m.weight.shape[1]
            else:
                in_features, out_features = m.in_features, m.out_features
            if dtypes.is_table_dtype(dtype):
                # nf4 / fp4: nearest-level codes + absmax scales (float4.cu), packed along K into the non-optimum layout
                # the reference forces for these types (modules.py:214-222); no zero points
                r = ops.f4_quantize(w, dtype, group_size, quantile)
                scale = double_quant_scales(r["scale"], w.dtype, cfg) if double_quant else r["scale"]
                new_module = B200WeightOnlyLinear(in_features, out_features, dtype=dtype, bits=bits, group_size=group_size,
                                                  zp=False, bias=m.bias is not None, device=device)
                new_module.set_packed(ops.pack_rows(r["codes"], bits), scale, m.bias)
            else:
                r = ops.rtn_quant_pack(w, bits, group_size, scheme == "sym", use_full_range, quantile)
                new_module = B200WeightOnlyLinear(in_features, out_features, dtype=dtype, bits=bits, group_size=group_size,
                                                  zp=scheme != "sym", bias=m.bias is not None, device=device)
                if double_quant:   # the codes keep the first-level scale; only the stored scales change (utility.py:377-434)
                    scales16, _ = ops.pack_params(double_quant_scales(r["scale_f32"], w.dtype, cfg), r["zp_f32"], bits)
                    r["scales"] = scales16
                new_module.set_packed(r["qweight"], r["qzeros"], r["scales"], m.bias)
            if model_device.type == "cuda":
                new_module.to(model_device)
            if name == "":
                return new_module
            set_module(model, name, new_module)
        if model_device.type != "cuda":
            # packed modules only exist on the B200 (no CPU path): keep the rest of the model with them, as the AWQ
            # and GPTQ quantizers do, instead of returning a mixed-device model that fails at its first forward
            model = model.to(device)
        return model
