"""Registry + module helpers.  Reference: neural_compressor/torch/utils/utility.py:48-201,
common/utils/constants.py:55-62 (Mode)."""
import enum
import logging
import os
from typing import Callable, Dict

import torch

logger = logging.getLogger("neural_compressor_b200")
if not logger.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(asctime)s [%(levelname)s][b200woq] %(message)s"))
    logger.addHandler(_h)
    logger.setLevel(os.environ.get("LOGLEVEL", "WARNING").upper())
    logger.propagate = False


class Mode(enum.Enum):
    PREPARE = "prepare"
    CONVERT = "convert"
    QUANTIZE = "quantize"


algos_mapping: Dict[str, Callable] = {}


def register_algo(name):
    """utility.py:63-82: the plug-in seam -- entry(model, configs_mapping, mode, *args, **kwargs) -> model."""

    def decorator(fn):
        algos_mapping[name] = fn
        return fn

    return decorator


def fetch_module(model, op_name):
    module = model
    for name in op_name.split("."):
        if hasattr(module, name):
            module = getattr(module, name)
        else:
            logger.warning(f"The {op_name} is not present in the model.")
            return None
    return module


def set_module(model, op_name, new_module):
    names = op_name.split(".")
    parent = model
    for name in names[:-1]:
        parent = getattr(parent, name)
    setattr(parent, names[-1], new_module)


def get_quantizer(model, quantizer_cls, quant_config=None, *args, **kwargs):
    """utility.py:163-181: reuse the quantizer stashed on the model at PREPARE."""
    if not hasattr(model, "quantizer"):
        return quantizer_cls(quant_config=quant_config, *args, **kwargs)
    return model.quantizer


def postprocess_model(model, mode, quantizer):
    """utility.py:184-201."""
    mode = Mode(getattr(mode, "value", mode))  # the reference's own Mode enum is accepted too (same values)
    if mode == Mode.PREPARE:
        model.quantizer = quantizer
    elif mode in (Mode.CONVERT, Mode.QUANTIZE):
        if getattr(model, "quantizer", False):
            del model.quantizer


def get_model_device(model: torch.nn.Module):
    for p in model.parameters():
        return p.device
    for b in model.buffers():
        return b.device
    return torch.device("cpu")


def current_device() -> torch.device:
    """The B200 this process drives.  There is no CPU fallback: raise when CUDA is missing."""
    if not torch.cuda.is_available():
        from .._lib import B200WOQError

        raise B200WOQError("neural_compressor_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def get_block_prefix(model):
    """weight_only/utility.py:988-1006: name and length of the first nn.ModuleList."""
    for n, m in model.named_modules():
        if type(m) is torch.nn.ModuleList:
            assert len(m) > 0, "block num shouldn't be zero!"
            return n, len(m)
    raise ValueError("no nn.ModuleList of transformer blocks found in the model")


def _layer_types():
    types = [torch.nn.Linear]
    try:
        import transformers

        types.append(transformers.Conv1D)
    except Exception:  # pragma: no cover
        pass
    return tuple(types)


def find_layers(module, name=""):
    """gptq.py:109-131: quantizable leaf layers of a block, in registration order."""
    types = _layer_types()
    if isinstance(module, types):
        return {name: module}
    res = {}
    for child_name, child in module.named_children():
        res.update(find_layers(child, name + "." + child_name if name else child_name))
    return res


def move_to_device(obj, device):
    if isinstance(obj, torch.Tensor):
        return obj.to(device)
    if isinstance(obj, (list, tuple)):
        return type(obj)(move_to_device(o, device) for o in obj)
    if isinstance(obj, dict):
        return {k: move_to_device(v, device) for k, v in obj.items()}
    return obj


def dump_model_op_stats(mode, configs_mapping):
    """torch/utils/utility.py:204-254: the "Mixed Precision Statistics" table logged after convert -- per operator type how
    many operators ended in which weight format (A32W<bits>G<group_size>) and how many stayed FP32."""
    if getattr(mode, "value", mode) == Mode.PREPARE.value:
        return
    rows, formats = {}, set()
    for (_name, op_type), cfg in configs_mapping.items():
        fmt = "FP32" if getattr(cfg, "dtype", "fp32") == "fp32" else "A32W{}G{}".format(getattr(cfg, "bits", "?"),
                                                                                     getattr(cfg, "group_size", "?"))
        formats.add(fmt)
        rows.setdefault(op_type, {})
        rows[op_type][fmt] = rows[op_type].get(fmt, 0) + 1
    formats.add("FP32")
    cols = sorted(formats)
    header = ["Op Type", "Total"] + cols
    table = [header] + [[t, sum(c.values())] + [c.get(f, 0) for f in cols] for t, c in rows.items()]
    widths = [max(len(str(r[i])) for r in table) for i in range(len(header))]
    line = "+" + "+".join("-" * (w + 2) for w in widths) + "+"
    logger.info("|" + "Mixed Precision Statistics".center(len(line) - 2, "*") + "|")
    logger.info(line)
    for i, r in enumerate(table):
        logger.info("|" + "|".join(" " + str(v).center(w) + " " for v, w in zip(r, widths)) + "|")
        if i == 0:
            logger.info(line)
    logger.info(line)
