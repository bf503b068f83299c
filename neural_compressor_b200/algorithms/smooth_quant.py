"""SmoothQuant (BASELINE configs[3]) -- calibration statistics + smoothing on the B200.

Reference: neural_compressor/torch/algorithms/smooth_quant/utility.py (Calibration :840-953, cal_scale :605-626,
TorchSmoothQuant._cal_scales :2122-2156 / _absorb_scales :1994-2061 / _parse_absorb_to_layers :2225-2287 / transform
:2289-2432, SQLinearWrapper :2559-2662, quant_dequant_w_v1/x_v1 :652-755, WrapperLayer.q_dq_forward :2707-2729).
PARITY: the smoothing transform (calibration ranges, scale-sharing groups, `cal_scale`, folding into the producer, the
static activation qparams of `SQLinearWrapper`) and the QDQ simulation are pinned against the live reference run on the
CPU with intel_extension_for_pytorch stubbed (oracle/ref_loader.py load_smooth_quant_utility, tests/golden/
sq_transform.pt).  The reference's int8 GEMM itself lives in IPEX/oneDNN outside the tree: the W8A8 kernel
is checked against the QDQ simulation, not against IPEX.

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

    def __init__(self, linear: torch.nn.Linear, smooth_scale, act_min, act_max, folded=False):
        """`folded`: the producer of the activation already carries 1/s (`_absorb_scales`, :1994-2061), so there is no
        run-time multiply; `input_scale` is still kept -- the calibrated range belongs to the un-smoothed activation."""
        super().__init__()
        self.folded = bool(folded)
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
                               input_scale=None if self.folded else self.input_scale,
                               bias=None if self.bias is None else self.bias.data)

    def forward_qdq(self, x):
        xs = x.float() if self.folded else x.float() * self.input_scale
        q = torch.round(xs / self.x_scale + self.x_zp).clamp_(0, 255)
        xq = self.x_scale * (q - self.x_zp)
        w = self.qweight[:, :self.in_features].float() * self.w_scale.view(-1, 1)
        y = torch.nn.functional.linear(xq, w, None if self.bias is None else self.bias.float())
        return y.to(x.dtype)


def absorb_scale(layer: torch.nn.Module, scale: torch.Tensor):
    """`_absorb_scales` (:1994-2061): multiply the producer's output channels by `scale` (= 1 / smoothing scale)."""
    name = type(layer).__name__
    if isinstance(layer, (torch.nn.LayerNorm, torch.nn.BatchNorm2d, torch.nn.GroupNorm, torch.nn.InstanceNorm2d)):
        layer.weight.data.mul_(scale.to(layer.weight.dtype))
        if getattr(layer, "bias", None) is not None:
            layer.bias.data.mul_(scale.to(layer.bias.dtype))
    elif isinstance(layer, torch.nn.Linear):
        if layer.bias is not None:
            layer.bias.data.mul_(scale.to(layer.bias.dtype))
        layer.weight.data.mul_(scale.view(-1, 1).to(layer.weight.dtype))
    elif name in ("LlamaRMSNorm", "T5LayerNorm"):
        layer.weight.data.mul_(scale.to(layer.weight.dtype))
    elif getattr(getattr(layer, "weight", None), "dim", lambda: 0)() == 1 and name.endswith("Norm"):
        # a norm type admitted by absorb.get_absorb_layers(extended=True) after its numerical fold check
        layer.weight.data.mul_(scale.to(layer.weight.dtype))
        if getattr(layer, "bias", None) is not None:
            layer.bias.data.mul_(scale.to(layer.bias.dtype))
    else:
        raise NotImplementedError(f"cannot fold a smoothing scale into {name}")


def _qdq_weight(w):
    """quant_dequant_w_v1 (:652-690), nn.Linear, symmetric int8 per output channel."""
    eps = torch.finfo(torch.float32).eps
    scale = torch.clip(torch.max(torch.abs(w), dim=1).values / (255.0 / 2), min=eps).unsqueeze(-1)
    return torch.round(w / scale).clamp_(-128.0, 127.0) * scale


def _qdq_act(x, min_x, max_x):
    """quant_dequant_x_v1 (:726-755): asymmetric uint8 over the given (per-tensor) range."""
    eps = torch.finfo(torch.float32).eps
    max_x, min_x = torch.max(max_x), torch.min(min_x)
    scale = torch.clip((max_x - min_x) / 255, min=eps)
    bias = torch.round((0 - min_x) / scale)
    return scale * (torch.round(x / scale + bias).clamp_(0, 255.0) - bias)


class _QDQProbe(torch.nn.Module):
    """`WrapperLayer` (:2665-2755): a Linear that can run its W8A8 quant-dequant simulation for a given smoothing scale
    and remembers the (upstream-quantised) input and the output of its last call -- what the alpha search compares."""

    def __init__(self, layer, input_min, input_max):
        super().__init__()
        self.add_module("orig_layer", layer)
        self.quant = False
        self.q_input = self.output = None
        self.input_min, self.input_max = input_min, input_max
        self.input_scale = self.weight_scale = None

    def q_dq_forward(self, x, input_scale, weight_scale):
        dtype = x.dtype      # the simulation runs in fp32 (the reference's only mode); a half model gets its dtype back
        x = x.float()
        w = self.orig_layer.weight.detach().float()
        w = _qdq_weight(w * weight_scale if weight_scale is not None else w)
        if input_scale is None:
            x = _qdq_act(x, self.input_min, self.input_max)
        else:
            x = _qdq_act(input_scale * x, self.input_min * input_scale, self.input_max * input_scale)
        bias = self.orig_layer.bias
        return torch.nn.functional.linear(x, w, None if bias is None else bias.float()).to(dtype)

    def forward(self, x):
        if self.quant:
            self.q_input = x
            out = self.q_dq_forward(x, self.input_scale, self.weight_scale)
        else:
            out = self.orig_layer(x)
        self.output = out
        return out


def _auto_loss(output, output_q):
    """`_get_auto_loss` (:1483-1507), loss_type "abs": sum |o/max - o_q/max|^0.5 with a per-sample max."""
    if output.dim() <= 2:
        max_value = torch.max(torch.abs(output))
    else:
        output = output.reshape(output.shape[0], -1)
        output_q = output_q.reshape(output_q.shape[0], -1)
        max_value = torch.clip(torch.max(torch.abs(output), dim=-1).values.unsqueeze(-1), 1e-5)
    return torch.sum(torch.pow(torch.abs(output / max_value - output_q / max_value), 0.5))


def auto_tune_alpha(model, groups, stats, batches, alpha_min=0.0, alpha_max=1.0, alpha_step=0.1, init_alpha=0.5,
                    shared_criterion="mean", n_samples=32, **_unused):
    """`AutoAlpha._auto_tune_alpha` (:1751-1820), model-wise: per scale-sharing group the alpha whose W8A8 simulation
    stays closest to the fp output of its layers.

    groups  {key: [Linear names]}; stats {name: (max, min)} per input channel; batches: the model inputs seen during
    calibration (a list of (args, kwargs)).  Returns {key: alpha}.  Notes that matter for agreeing with the reference:
    the losses of ONE batch decide (its `loss_alphas` is re-created per batch, :1776), the running choice is refreshed every
    n_samples // 4 batches and determines the upstream quantisation the probes see, ties keep the reference's dict order
    (current alpha first, then the grid)."""
    import numpy

    digits = max(len(str(v).split(".")[1]) for v in (alpha_min, alpha_max, alpha_step))
    alpha_space = numpy.round(numpy.arange(alpha_min, alpha_max + alpha_step, alpha_step), digits).tolist()
    modules = dict(model.named_modules())
    names = [n for v in groups.values() for n in v]
    probes = {}
    for n in names:
        mx, mn = stats[n]
        probes[n] = _QDQProbe(modules[n], mn, mx)
        set_module(model, n, probes[n])

    def scales_for(alpha):
        out = {}
        for key, members in groups.items():
            a = alpha[key] if isinstance(alpha, dict) else alpha
            mx, mn = stats[members[0]]
            s = cal_scale(torch.maximum(mx.abs(), mn.abs()), [probes[n].orig_layer.weight.data.float() for n in members], a)
            inv = 1.0 / s
            inv[s == 0] = 0
            for n in members:
                out[n] = (inv.view(1, -1), s.view(1, -1))
        return out

    def apply(alpha):
        for n, (i, w) in scales_for(alpha).items():
            probes[n].input_scale, probes[n].weight_scale = i, w

    def run(batch):
        args, kwargs = batch
        model(*args, **kwargs)

    def best_of(losses):
        best = {}
        for key, members in groups.items():
            crit = "min" if len(members) == 1 else shared_criterion
            if crit == "mean":
                total = {}
                for a in losses[members[0]]:
                    total[a] = sum(losses[n][a] for n in members)
                best[key] = float(sorted(total.items(), key=lambda kv: kv[1])[0][0])
            elif crit in ("min", "max"):
                picks = [float(sorted(losses[n].items(), key=lambda kv: kv[1])[0][0]) for n in members]
                best[key] = min(picks) if crit == "min" else max(picks)
            else:
                raise NotImplementedError(crit)
        return best

    try:
        apply(init_alpha)
        best = init_alpha
        every = n_samples // 4 if n_samples >= 4 else n_samples
        seen = since = 0
        losses = None
        for batch in batches:
            per_module = best
            if isinstance(best, dict):
                per_module = dict(best)
                for key, members in groups.items():
                    for n in members:
                        per_module[n] = best[key]
            for p in probes.values():
                p.quant = False
            run(batch)
            fp_out = {n: probes[n].output for n in names}
            for p in probes.values():
                p.quant = True
            apply(per_module if not isinstance(per_module, dict) else {k: per_module[k] for k in groups})
            run(batch)
            losses = {}
            for n in names:
                cur = per_module[n] if isinstance(per_module, dict) else per_module
                losses[n] = {str(cur): _auto_loss(fp_out[n], probes[n].output)}
            for a in alpha_space:
                sc = scales_for(a)
                for n in names:
                    if str(a) in losses[n]:
                        continue
                    i, w = sc[n]
                    losses[n][str(a)] = _auto_loss(fp_out[n], probes[n].q_dq_forward(probes[n].q_input, i, w))
                # the probes keep the grid's last scales, exactly like the reference's `_update_scales_for_auto` calls
                for n, (i, w) in sc.items():
                    probes[n].input_scale, probes[n].weight_scale = i, w
            seen += 1
            since += 1
            if since // every >= 1:
                since = 0
                best = best_of(losses)
                apply(best)
            if seen >= n_samples:
                break
        return best_of(losses) if losses is not None else {k: init_alpha for k in groups}
    finally:
        for n, p in probes.items():
            set_module(model, n, p.orig_layer)


class SmoothQuantQuantizer(Quantizer):
    """Calibrate -> group -> smooth -> static W8A8 modules.

    Grouping (which Linears share one smoothing scale, and whether the scale can be folded away) follows
    `TorchSmoothQuant._parse_absorb_to_layers`.  The reference finds the structure with its torch.jit tracer, which fails
    on transformers-5 models -- every Linear then gets its own scale and `folding=True` smooths nothing; that observable
    behaviour is the default here.  `absorb_discovery="eager"` (B200WOQ_SQ_ABSORB=eager), implied by `folding=True`, uses
    algorithms/absorb.py instead: Linears reading the same tensor share a scale (insert-mul mode), and with
    `folding=True` only foldable groups are smoothed and 1/s goes into the producing norm / Linear."""

    def __init__(self, quant_config=None, absorb_discovery=None):
        super().__init__(quant_config)
        import os

        self.absorb_discovery = absorb_discovery or os.environ.get("B200WOQ_SQ_ABSORB", "off")
        assert self.absorb_discovery in ("off", "eager"), self.absorb_discovery

    def prepare(self, model, example_inputs=None, *args, **kwargs):
        """Register per-input-channel min/max hooks on every Linear (utility.py:858-883, 929-953)."""
        self.device = current_device()
        model.to(self.device)
        self.example_inputs = example_inputs
        self._stats, self._handles = {}, []
        self._batches = []           # model-level inputs of the calibration run, for alpha="auto" (build_captured_dataloader)

        def remember(_m, args, kwargs):
            if len(self._batches) < 128:
                self._batches.append((args, kwargs))

        self._handles.append(model.register_forward_pre_hook(remember, with_kwargs=True))
        for name, m in model.named_modules():
            if isinstance(m, torch.nn.Linear):
                k = m.in_features
                self._stats[name] = (torch.full((k,), -float("inf"), device=self.device),
                                     torch.full((k,), float("inf"), device=self.device))

                def hook(_m, inp, _o, _n=name):
                    mx, mn = self._stats[_n]
                    ops.minmax_cols_accumulate(inp[0].detach(), mx, mn)

                self._handles.append(m.register_forward_hook(hook))
        return model

    def _groups(self, model, folding):
        """-> ({key: [Linear names]}, folded).  key = the absorbing module when folded, else the group's first Linear."""
        calibrated = [n for n, (mx, _) in self._stats.items() if not torch.isinf(mx).any()]
        if folding or self.absorb_discovery == "eager":
            from .absorb import get_absorb_layers, get_shared_input_groups

            if folding:
                absorb_to_layer, _ = get_absorb_layers(model, self.example_inputs, supported_layers=["Linear"])
                groups = {k: [n for n in v if n in calibrated] for k, v in absorb_to_layer.items()}
                return {k: v for k, v in groups.items() if v}, True
            groups = get_shared_input_groups(model, self.example_inputs)
            groups = {k: [n for n in v if n in calibrated] for k, v in groups.items()}
            return {v[0]: v for v in groups.values() if v}, False
        return {n: [n] for n in calibrated}, False

    @torch.no_grad()
    def convert(self, model, *args, **kwargs):
        for h in self._handles:
            h.remove()
        alpha = self.quant_config.alpha
        auto = isinstance(alpha, str)
        if auto:
            assert alpha == "auto", f"alpha must be a number or 'auto', got {alpha!r}"
        else:
            alpha = float(alpha)
        folding = bool(getattr(self.quant_config, "folding", False))
        for name, (mx, _mn) in self._stats.items():
            if torch.isinf(mx).any():
                logger.warning(f"{name} saw no calibration data; left in fp")
        groups, folded = self._groups(model, folding)
        if folding and not groups:
            logger.warning("empty absorb_to_layer, smoothquant is ignored")   # utility.py:2367-2369
        modules = dict(model.named_modules())
        if auto:   # per-group alpha from the W8A8 simulation (AutoAlpha, :1232-1893; model-wise search)
            args = dict(getattr(self.quant_config, "auto_alpha_args", None) or {})
            if args.pop("do_blockwise", False):
                raise NotImplementedError("blockwise auto-alpha tuning is not built (model-wise only)")
            alpha = auto_tune_alpha(model, groups, self._stats, self._batches, **args)
            for key, a in alpha.items():
                logger.info(f"Final alpha {key}:{a}")
            self.tuned_alpha = alpha
        # all scales come from the un-modified weights (`_cal_scales`, :2122-2156) ...
        scales = {}
        for key, names in groups.items():
            mx, mn = self._stats[names[0]]                # the group shares one input (:2132-2136, 2182)
            in_max_abs = torch.maximum(mx.abs(), mn.abs())
            scales[key] = cal_scale(in_max_abs, [modules[n].weight.data.float() for n in names],
                                    alpha[key] if isinstance(alpha, dict) else alpha)
        # ... then every fold goes into the still-fp producers (a producer may itself be a smoothed Linear: fc1 takes
        # fc2's 1/s on its rows and its own s on its columns) ...
        if folded:
            for key, s in scales.items():
                absorb = 1.0 / s
                absorb[s == 0] = 0                        # :2150-2151
                absorb_scale(modules[key], absorb)
        # ... and only then the weights are smoothed and quantised
        for key, names in groups.items():
            mx, mn = self._stats[names[0]]
            for n in names:
                set_module(model, n, SQLinear(modules[n], scales[key], mn, mx, folded=folded))
        return model
