"""SmoothQuant (BASELINE configs[3]) -- calibration statistics + smoothing on the B200.

Reference: neural_compressor/torch/algorithms/smooth_quant/utility.py (Calibration :840-953, cal_scale :605-626,
SQLinearWrapper :2559-2662, quant_dequant_w_v1/x_v1 :652-755, WrapperLayer.q_dq_forward :2707-2729).
PARITY UNPINNED: the reference module hard-imports intel_extension_for_pytorch and its W8A8 GEMM lives in
IPEX/oneDNN outside the tree; this row follows the source text and the QDQ simulation only.

Per-input-channel min/max calibration (stats.cu), the alpha scale, weight smoothing + int8 quantisation
(b200woq_sq_smooth_quant_weight) and the static W8A8 forward on the tensor cores (b200woq_w8a8_linear_forward,
tcgen05.mma.kind::i8, w8a8.cu).
"""
from __future__ import annotations

import torch

from .. import ops
from ..utils import current_device, logger, set_module
from .base_algorithm import Quantizer


def cal_scale(input_max_abs, weights, alpha, weight_max_lb=1e-5):
    """smooth_quant/utility.py:605-626."""
    w = torch.cat(weights, dim=0)
    weight_max = torch.clip(torch.max(torch.abs(w), dim=0)[0], weight_max_lb)
    input_power = torch.pow(input_max_abs, alpha)
    weight_power = torch.pow(weight_max, 1 - alpha)
    scale = torch.clip(input_power / weight_power, min=1e-5)
    scale[input_power == 0] = 1.0
    return scale


class SQLinear(torch.nn.Module):
    """SQLinearWrapper (utility.py:2559-2662) as a static W8A8 module on the B200: the smoothing scale is folded into the
    weight (W' = W * s, `_scale_layer_weight`) and applied to the input as a multiply by `input_scale` = 1/s (:2599-2600);
    weights are per-out-channel symmetric int8 (`quant_dequant_w_v1`, :652-690), activations static per-tensor asymmetric
    uint8 from the calibrated range of the smoothed input (`_calculate_qparams`, :2607-2631; `quant_dequant_x_v1`, :726-755).
    forward = ONE activation-quantisation kernel + ONE tcgen05 `kind::i8` GEMM with the dequant epilogue
    `(acc - zp_x * sum_k q_w) * s_x * s_w[n] + bias` (ops.w8a8_linear) -- the integer GEMM the reference delegates to IPEX.
    `forward_qdq` evaluates the reference's pure-torch QDQ simulation (`WrapperLayer.q_dq_forward`, :2707-2729) for
    cross-checks; it is not used by the product path."""

    def __init__(self, linear: torch.nn.Linear, smooth_scale, act_min, act_max):
        super().__init__()
        self.in_features, self.out_features = linear.in_features, linear.out_features
        dev = linear.weight.device
        smooth_scale = smooth_scale.to(dev).float().contiguous()
        r = ops.sq_smooth_quant_weight(linear.weight.data.contiguous(), smooth_scale)
        self.register_buffer("qweight", r["qweight"])          # int8 [N, padded K]
        self.register_buffer("w_scale", r["w_scale"])          # fp32 [N]
        self.register_buffer("wsum", r["wsum"])                # int32 [N]
        self.register_buffer("input_scale", (1.0 / smooth_scale).float())
        eps = torch.finfo(torch.float32).eps
        # static per-tensor activation qparams from the calibrated range of the SMOOTHED input (:2607-2631)
        mn = torch.clamp((act_min.to(dev) * self.input_scale).min(), max=0.0)
        mx = torch.clamp((act_max.to(dev) * self.input_scale).max(), min=0.0)
        x_scale = torch.clip((mx - mn) / 255.0, min=eps)
        self.register_buffer("x_scale", x_scale.reshape(1).float())
        self.register_buffer("x_zp", torch.clamp(torch.round((0 - mn) / x_scale), 0, 255).reshape(1).float())
        self.bias = None if linear.bias is None else torch.nn.Parameter(linear.bias.data.clone(), requires_grad=False)

    def forward(self, x):
        return ops.w8a8_linear(x, self.qweight, self.w_scale, self.wsum, self.x_scale, self.x_zp, self.in_features,
                               input_scale=self.input_scale, bias=None if self.bias is None else self.bias.data)

    def forward_qdq(self, x):
        xs = x.float() * self.input_scale
        q = torch.round(xs / self.x_scale + self.x_zp).clamp_(0, 255)
        xq = self.x_scale * (q - self.x_zp)
        w = self.qweight[:, :self.in_features].float() * self.w_scale.view(-1, 1)
        y = torch.nn.functional.linear(xq, w, None if self.bias is None else self.bias.float())
        return y.to(x.dtype)


class SmoothQuantQuantizer(Quantizer):
    def __init__(self, quant_config=None):
        super().__init__(quant_config)

    def prepare(self, model, example_inputs=None, *args, **kwargs):
        """Register per-input-channel min/max hooks on every Linear (utility.py:858-883)."""
        self.device = current_device()
        model.to(self.device)
        self._stats, self._handles = {}, []
        for name, m in model.named_modules():
            if isinstance(m, torch.nn.Linear) and "lm_head" not in name:
                k = m.in_features
                self._stats[name] = (torch.full((k,), -float("inf"), device=self.device),
                                     torch.full((k,), float("inf"), device=self.device))

                def hook(_m, inp, _o, _n=name):
                    mx, mn = self._stats[_n]
                    ops.minmax_cols_accumulate(inp[0].detach(), mx, mn)

                self._handles.append(m.register_forward_hook(hook))
        return model

    @torch.no_grad()
    def convert(self, model, *args, **kwargs):
        for h in self._handles:
            h.remove()
        alpha = float(self.quant_config.alpha) if not isinstance(self.quant_config.alpha, str) else 0.5
        for name, m in list(model.named_modules()):
            if name not in self._stats:
                continue
            mx, mn = self._stats[name]
            if torch.isinf(mx).any():
                logger.warning(f"{name} saw no calibration data; left in fp")
                continue
            in_max_abs = torch.maximum(mx.abs(), mn.abs())
            s = cal_scale(in_max_abs, [m.weight.data.float()], alpha)
            set_module(model, name, SQLinear(m, s, mn, mx))
        return model
