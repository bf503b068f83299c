"""SmoothQuant (BASELINE configs[3]) -- calibration statistics + smoothing on the B200.

Reference: neural_compressor/torch/algorithms/smooth_quant/utility.py (Calibration :840-953, cal_scale :605-626,
SQLinearWrapper :2559-2662, quant_dequant_w_v1/x_v1 :652-755, WrapperLayer.q_dq_forward :2707-2729).
PARITY UNPINNED: the reference module hard-imports intel_extension_for_pytorch and its W8A8 GEMM lives in
IPEX/oneDNN outside the tree; this row follows the source text and the QDQ simulation only.

Round 1 implements: per-input-channel min/max calibration (kernel), the alpha scale, weight smoothing and a
`SQLinear` module that applies x*1/s and evaluates the W8A8 QDQ simulation.  The tcgen05 INT8 GEMM is the next
row (DESIGN.md "what comes next").
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
    """SQLinearWrapper semantics (utility.py:2559-2662) evaluated as the QDQ simulation (utility.py:2707-2729):
    x' = x * (1/s); xq = qdq_uint8_per_tensor(x') with static (calibrated) min/max; wq = qdq_int8_per_channel(W*s)."""

    def __init__(self, linear: torch.nn.Linear, smooth_scale, act_min, act_max):
        super().__init__()
        self.in_features, self.out_features = linear.in_features, linear.out_features
        w = linear.weight.data.float() * smooth_scale.view(1, -1)
        eps = torch.finfo(torch.float32).eps
        w_scale = torch.clip(w.abs().amax(dim=1) / 127.5, min=eps).view(-1, 1)   # utility.py:673-676
        self.register_buffer("qweight", torch.round(w / w_scale).clamp_(-128, 127).to(torch.int8))
        self.register_buffer("w_scale", w_scale)
        self.register_buffer("input_scale", (1.0 / smooth_scale).float())
        # static per-tensor activation qparams from the calibrated range of the SMOOTHED input (:2607-2631)
        mn = torch.clamp((act_min * self.input_scale).min(), max=0.0)
        mx = torch.clamp((act_max * self.input_scale).max(), min=0.0)
        x_scale = torch.clip((mx - mn) / 255.0, min=eps)
        self.register_buffer("x_scale", x_scale.reshape(1))
        self.register_buffer("x_zp", torch.round((0 - mn) / x_scale).reshape(1))
        self.bias = None if linear.bias is None else torch.nn.Parameter(linear.bias.data.clone(), requires_grad=False)

    def forward(self, x):
        xs = x.float() * self.input_scale
        q = torch.round(xs / self.x_scale + self.x_zp).clamp_(0, 255)
        xq = self.x_scale * (q - self.x_zp)
        y = torch.nn.functional.linear(xq, self.qweight.float() * self.w_scale, None if self.bias is None else self.bias.float())
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
