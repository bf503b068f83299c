"""Absorb-layer discovery for AWQ (and SmoothQuant-style folding) without torch.jit.

Reference: neural_compressor/torch/algorithms/weight_only/utility.py
    get_absorb_layers :657-688, GraphTrace :728-984 (get_prev_absorb_layer :822-858, skip_op_absorb_helper :860-880,
    remove_unsupported_layers :957-984).

What the reference computes: for every Linear, the module that produced its input -- looking through dtype casts and
ReLU-like activations, for which a(f(x)) = f(a x) -- provided EVERY consumer of that producer's output can take a
per-channel scale (Linear / norm / conv modules, element-wise `mul`), again looking through the same pass-through ops.
A scale `s` on the Linear's input channels can then be folded into the producer (`weight / s`) instead of a run-time
multiply.  Producers of the supported module types are kept, every other Linear is reported as not absorbable.

How: the reference walks a `torch.jit.trace` graph.  Tracing Hugging Face models stopped working with transformers 5
(GraphTrace.trace raises inside the library and the reference silently treats every Linear as not absorbable -- checked
on llama / opt / gpt-j in this image).  Here ONE eager forward is observed instead: module forward hooks record which
tensor enters and leaves every module, a `TorchDispatchMode` records which aten op consumes which tensor, and the same
rules are evaluated on that record.  Identity is Python object identity of the tensors that flow between modules, kept
alive for the duration of the trace.  Anything the record cannot classify (views, in-place edits, functional linears) is
treated as not absorbable, so the result is never more permissive than the reference's.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch.utils._python_dispatch import TorchDispatchMode
from torch.utils._pytree import tree_flatten

from ..utils import logger, move_to_device

# utility.py:733-744: module types that can take part in absorption
SUPPORTED_MODULES = ("Linear", "Conv2d", "ConvTranspose2d", "LayerNorm", "BatchNorm2d", "GroupNorm", "InstanceNorm2d",
                     "LlamaRMSNorm", "T5LayerNorm", "LPLayerNorm")
# utility.py:748: a(f(x)) = f(a x)
PASS_THROUGH_OPS = ("aten::to", "aten::_to_copy", "aten::relu", "aten::leaky_relu", "aten::hardtanh")
# utility.py:750-758 lists `aten::mul` among the ops that may consume a scaled tensor
SCALABLE_FUNCTIONAL_OPS = ("aten::mul",)


def _valid_conv(module) -> bool:
    """utility.py:897-906: grouped convolutions are excluded, except depthwise ones."""
    if not isinstance(module, torch.nn.Conv2d) or module.groups == 1:
        return True
    return module.in_channels == module.out_channels and module.groups == module.in_channels


class _Use:
    __slots__ = ("op", "module", "is_module_input", "outputs")

    def __init__(self, op, module, is_module_input, outputs):
        self.op, self.module, self.is_module_input, self.outputs = op, module, is_module_input, outputs


class EagerTrace(TorchDispatchMode):
    """Record of one forward: per tensor its producer op, its consumers and the supported module it left / entered."""

    def __init__(self, model: torch.nn.Module, supported=SUPPORTED_MODULES):
        super().__init__()
        self.names = {id(m): n for n, m in model.named_modules()}
        self.supported = tuple(supported)
        self.stack: List[torch.nn.Module] = []
        self.keep: List[torch.Tensor] = []                  # keeps ids unique for the duration of the trace
        self.uses: Dict[int, List[_Use]] = {}
        self.producer: Dict[int, Tuple[str, Optional[int]]] = {}   # tensor -> (op, first tensor input)
        self.mutated = set()
        self.module_input: Dict[int, int] = {}              # module -> tensor
        self.module_output: Dict[int, int] = {}             # tensor -> module
        self.calls: List[torch.nn.Module] = []              # supported modules in execution order
        self.by_id: Dict[int, torch.nn.Module] = {}

    # ---- module boundaries
    def _is_supported(self, m):
        return type(m).__name__ in self.supported and id(m) in self.names

    def _pre(self, module, args):
        self.stack.append(module)
        if self._is_supported(module):
            x = next((a for a in args if isinstance(a, torch.Tensor)), None)
            if x is not None:
                self.keep.append(x)
                self.module_input[id(module)] = id(x)
                self.by_id[id(module)] = module
                self.calls.append(module)

    def _post(self, module, args, output):
        if self.stack and self.stack[-1] is module:
            self.stack.pop()
        if self._is_supported(module):
            out = output[0] if isinstance(output, (tuple, list)) else output
            if isinstance(out, torch.Tensor):
                self.keep.append(out)
                self.module_output[id(out)] = id(module)

    # ---- aten ops
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        name = func._schema.name
        ins = [t for t in tree_flatten((args, kwargs or {}))[0] if isinstance(t, torch.Tensor)]
        outs = [t for t in tree_flatten(out)[0] if isinstance(t, torch.Tensor)]
        self.keep.extend(ins)
        self.keep.extend(outs)
        ctx = next((m for m in reversed(self.stack) if self._is_supported(m)), None)
        out_ids = [id(t) for t in outs]
        for t in ins:
            is_in = ctx is not None and self.module_input.get(id(ctx)) == id(t)
            self.uses.setdefault(id(t), []).append(_Use(name, ctx, is_in, out_ids))
        in_ids = {id(t) for t in ins}
        for t in outs:
            if id(t) in in_ids:          # in-place op or an op returning its input
                self.mutated.add(id(t))
            else:
                self.producer[id(t)] = (name, id(ins[0]) if ins else None)
        return out

    def run(self, model, example_inputs):
        pre = torch.nn.modules.module.register_module_forward_pre_hook(self._pre)
        post = torch.nn.modules.module.register_module_forward_hook(self._post)
        try:
            with torch.no_grad(), self:
                if isinstance(example_inputs, dict):
                    model(**example_inputs)
                elif isinstance(example_inputs, (tuple, list)):
                    model(*example_inputs)
                else:
                    model(example_inputs)
        finally:
            pre.remove()
            post.remove()
        return self

    # ---- the reference's rules on the record
    def _scalable_consumers(self, tensor_id, depth=0) -> bool:
        """Every consumer takes a per-channel scale, directly or through pass-through ops (utility.py:836-880)."""
        if depth > 16 or tensor_id in self.mutated:
            return False
        for use in self.uses.get(tensor_id, []):
            if use.module is not None:
                if use.is_module_input:
                    continue                      # a supported module reading it as its input
                return False                      # read from inside some other module's computation
            if use.op in SCALABLE_FUNCTIONAL_OPS:
                continue
            if use.op in PASS_THROUGH_OPS:
                if all(self._scalable_consumers(o, depth + 1) for o in use.outputs):
                    continue
            return False
        return True

    def producer_module(self, linear) -> Optional[torch.nn.Module]:
        """The supported module whose output feeds `linear`, or None (utility.py:822-858)."""
        t = self.module_input.get(id(linear))
        for _ in range(16):
            if t is None or t in self.mutated:
                return None
            if t in self.module_output:
                parent = self.by_id.get(self.module_output[t])
                if parent is None or parent is linear:
                    return None
                return parent if self._scalable_consumers(t) else None
            op, src = self.producer.get(t, (None, None))
            if op not in PASS_THROUGH_OPS:
                return None
            t = src
        return None


    def shared_input_root(self, linear) -> Optional[int]:
        """Identity of the tensor `linear` reads, looked through pass-through ops, when that tensor comes out of a
        supported module or an element-wise mul and all of its consumers can take a scale -- the condition under which
        the reference puts Linears that read the same tensor into one scale-sharing group
        (`get_absorb_to_layer(..., skip_unsupported_layers=False)`, smooth_quant/utility.py:2225-2246)."""
        t = self.module_input.get(id(linear))
        for _ in range(16):
            if t is None or t in self.mutated:
                return None
            op, src = self.producer.get(t, (None, None))
            if t in self.module_output or op in SCALABLE_FUNCTIONAL_OPS:
                return t if self._scalable_consumers(t) else None
            if op not in PASS_THROUGH_OPS:
                return None
            t = src
        return None


def _shrink(example_inputs, max_tokens=16):
    """The trace keeps every tensor of the observed forward alive (identity must stay unique), and the structure does
    not depend on the sequence length: [batch, seq] tensors (input_ids, attention_mask, ...) are cut to one sample of at
    most `max_tokens` tokens so that tracing a 70B model does not hold a full-length forward's activations."""
    def cut(t):
        if isinstance(t, torch.Tensor) and t.dim() == 2 and (t.shape[0] > 1 or t.shape[1] > max_tokens):
            return t[:1, :max_tokens]
        return t

    if isinstance(example_inputs, dict):
        return {k: cut(v) for k, v in example_inputs.items()}
    if isinstance(example_inputs, (list, tuple)):
        return type(example_inputs)(cut(v) for v in example_inputs)
    return cut(example_inputs)


def _trace_model(model, example_inputs):
    if example_inputs is None:
        logger.warning("No example_inputs: absorb layer detection is skipped")
        return None
    device = next(model.parameters()).device
    try:
        return EagerTrace(model).run(model, move_to_device(_shrink(example_inputs), device))
    except Exception as ex:  # pragma: no cover
        logger.warning(f"Eager trace failed ({type(ex).__name__}: {ex}), absorb layer detection is skipped")
        return None


def get_shared_input_groups(model, example_inputs) -> Dict[str, List[str]]:
    """{first Linear: [Linears reading the same tensor]} over every Linear of the model, in execution order; a Linear
    whose input cannot be shared (or an unobservable forward) forms its own group.  SmoothQuant's `insert_mul` mode gives
    each group ONE smoothing scale computed from the concatenated weights (smooth_quant/utility.py:2122-2156, 2225-2246)."""
    all_linears = [n for n, m in model.named_modules() if type(m).__name__ == "Linear"]
    trace = _trace_model(model, example_inputs)
    if trace is None:
        return {n: [n] for n in all_linears}
    groups: Dict[str, List[str]] = {}
    by_root: Dict[int, str] = {}
    seen = set()
    for m in trace.calls:
        if type(m).__name__ != "Linear" or id(m) in seen:
            continue
        seen.add(id(m))
        name = trace.names[id(m)]
        root = trace.shared_input_root(m)
        if root is not None and root in by_root:
            groups[by_root[root]].append(name)
        else:
            groups[name] = [name]
            if root is not None:
                by_root[root] = name
    for n in all_linears:       # modules the forward never reached
        if not any(n in v for v in groups.values()):
            groups[n] = [n]
    return groups


def _folds_exactly(model, example_inputs, absorber, linears, rtol=1e-4) -> bool:
    """Numerical proof that a scale can be folded into `absorber`: put a random per-channel 1/s on its weight (and
    bias), s on the input channels of the Linears it feeds, and compare the model output on the example input.  Everything
    is restored afterwards.  This is how module types outside the reference's list are admitted (see `extended`)."""
    weight = getattr(absorber, "weight", None)
    if weight is None or weight.dim() != 1 or any(l.in_features != weight.numel() for l in linears):
        return False
    device = weight.device
    inputs = move_to_device(_shrink(example_inputs), device)

    def run():
        with torch.no_grad():
            out = model(**inputs) if isinstance(inputs, dict) else (model(*inputs) if isinstance(inputs, (list, tuple)) else model(inputs))
        out = out[0] if isinstance(out, (tuple, list)) else getattr(out, "logits", out)
        return out.float()

    g = torch.Generator().manual_seed(0)
    s = (torch.rand(weight.numel(), generator=g) + 0.5).to(device)
    saved = [(p, p.detach().clone()) for p in [weight, getattr(absorber, "bias", None)] + [l.weight for l in linears] if p is not None]
    try:
        before = run()
        with torch.no_grad():
            weight.div_(s.to(weight.dtype))
            if getattr(absorber, "bias", None) is not None:
                absorber.bias.div_(s.to(absorber.bias.dtype))
            for l in linears:
                l.weight.mul_(s.view(1, -1).to(l.weight.dtype))
        after = run()
    finally:
        with torch.no_grad():
            for p, v in saved:
                p.copy_(v)
    return bool(torch.isfinite(after).all()) and float((after - before).abs().max()) <= rtol * float(before.abs().max() + 1e-12)


def get_absorb_layers(model, example_inputs, supported_layers=("Linear",), folding=False, extended=None):
    """utility.py:657-688: ({absorbing module: [absorbed Linear, ...]}, [Linear names nothing can absorb]).  When the
    forward cannot be observed every Linear is reported as not absorbable, like the reference after a failed trace.

    `extended` (default: env B200WOQ_ABSORB_EXTENDED=1): the reference admits a fixed list of module types as absorbers
    (`SUPPORTED_MODULES`: LayerNorm, LlamaRMSNorm, T5LayerNorm, ...), which leaves Mistral / Qwen2 / ... without any.  In
    extended mode every `*Norm` module with a 1-D weight is a candidate, and a candidate TYPE is admitted only if folding a
    random scale into one of its instances leaves the model output unchanged (`_folds_exactly`) -- which rejects, e.g.,
    Gemma's RMSNorm (it multiplies by 1 + weight)."""
    import os

    if extended is None:
        extended = os.environ.get("B200WOQ_ABSORB_EXTENDED", "0") == "1"
    all_linears = [n for n, m in model.named_modules() if type(m).__name__ == "Linear"]
    if example_inputs is None:
        logger.warning("No example_inputs: absorb layer detection is skipped")
        return {}, all_linears
    supported = SUPPORTED_MODULES
    if extended:
        extra = {type(m).__name__ for m in model.modules()
                 if type(m).__name__.endswith("Norm") and getattr(getattr(m, "weight", None), "dim", lambda: 0)() == 1}
        supported = tuple(SUPPORTED_MODULES) + tuple(sorted(extra - set(SUPPORTED_MODULES)))
    device = next(model.parameters()).device
    try:
        trace = EagerTrace(model, supported).run(model, move_to_device(_shrink(example_inputs), device))
    except Exception as ex:  # pragma: no cover
        logger.warning(f"Eager trace failed ({type(ex).__name__}: {ex}), absorb layer detection is skipped")
        return {}, all_linears
    absorb_to_layer: Dict[str, List[str]] = {}
    no_absorb: List[str] = []
    seen = set()
    for m in trace.calls:
        if type(m).__name__ not in supported_layers or id(m) in seen:
            continue
        seen.add(id(m))
        name = trace.names[id(m)]
        parent = trace.producer_module(m)
        if parent is None or not _valid_conv(m) or not _valid_conv(parent):
            no_absorb.append(name)
        else:
            absorb_to_layer.setdefault(trace.names[id(parent)], []).append(name)
    if extended:
        verdict = {}   # module type outside the reference's list -> does a fold preserve the function?
        mods = dict(model.named_modules())
        for key in list(absorb_to_layer):
            tname = type(mods[key]).__name__
            if tname in SUPPORTED_MODULES:
                continue
            if tname not in verdict:
                verdict[tname] = _folds_exactly(model, example_inputs, mods[key], [mods[n] for n in absorb_to_layer[key]])
                logger.info(f"absorb discovery: {tname} {'admitted' if verdict[tname] else 'rejected'} by the fold check")
            if not verdict[tname]:
                no_absorb.extend(absorb_to_layer.pop(key))
    if not absorb_to_layer:
        logger.warning("No absorb layer is detected.")
    return absorb_to_layer, no_absorb
