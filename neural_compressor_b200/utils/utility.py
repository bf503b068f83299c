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


# ------------------------------------------------------------------------------------------------
# small calibration helpers of weight_only/utility.py (model_forward :546, forward_wrapper :566, move_input_to_device
# :587, get_example_input :1010, CapturedDataloader :1185, convert_dtype_str2torch :1218)
# ------------------------------------------------------------------------------------------------
def move_input_to_device(input, device=torch.device("cpu")):
    """Tensors inside (nested) dicts / lists / tuples go to `device`; anything else is handed back untouched."""
    from collections import UserDict

    if isinstance(input, (dict, UserDict)):
        return {k: move_input_to_device(v, device) for k, v in input.items()}
    if isinstance(input, (list, tuple)):
        moved = [move_input_to_device(v, device) for v in input]
        return tuple(moved) if isinstance(input, tuple) else moved
    return input.to(device) if isinstance(input, torch.Tensor) else input


def forward_wrapper(model, input, device=torch.device("cpu")):
    """One forward with whatever a dataloader yields: dict -> keywords, list / tuple -> positionals (or, when the model
    does not take them that way, the sequence itself), anything else -> the single argument."""
    from collections import UserDict

    try:
        model = model.to(device)
        input = move_input_to_device(input, device)
    except Exception as e:
        logger.warning(e)
        logger.warning("Please check the input device if the error raised.")
    if isinstance(input, (dict, UserDict)):
        return model(**input)
    if isinstance(input, (list, tuple)):
        try:
            return model(*input)
        except Exception:
            return model(input)
    return model(input)


def _batches(dataloader, limit):
    """(input, label) pairs when the loader yields pairs, the raw batches otherwise -- decided like the reference does, by
    whether unpacking a batch into two works."""
    def first(n_items, pick):
        for i, batch in enumerate(dataloader):
            if limit != -1 and i >= n_items:
                break
            yield pick(batch)

    try:
        return [inp for inp in first(limit, lambda b: (lambda inp, _label: inp)(*b))]
    except Exception:
        return list(first(limit, lambda b: b))


def model_forward(model, dataloader, iters, device):
    """Run `iters` batches (-1: all) of `dataloader` through `model`; labels, when present, are dropped."""
    for inp in _batches(dataloader, iters):
        forward_wrapper(model, inp, device)


def get_example_input(dataloader, i=1):
    """The i-th input of the loader (without its label), or the last one when the loader is shorter."""
    inputs = _batches(dataloader, i + 1)
    return inputs[min(i, len(inputs) - 1)] if inputs else None


class CapturedDataloader:
    """Replays the (args, kwargs) a model was called with during `run_fn` (weight_only/utility.py:1185-1203)."""

    def __init__(self, args_list, kwargs_list):
        self.args_list, self.kwargs_list = args_list, kwargs_list

    def __iter__(self):
        for args, kwargs in zip(self.args_list, self.kwargs_list):
            if not args:
                yield kwargs
            elif not kwargs:
                yield args[0] if len(args) == 1 else args
            else:
                yield args, kwargs


def convert_dtype_str2torch(str_dtype):
    """"fp16" / "float16" / "bf16" / "bfloat16" / "fp32" / "float32" / "auto" / "int8" -> torch dtype."""
    if isinstance(str_dtype, torch.dtype) or str_dtype is None:
        return str_dtype
    table = {"int8": torch.int8, "fp32": torch.float, "float32": torch.float, "auto": torch.float, "fp16": torch.float16,
             "float16": torch.float16, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16}
    assert str_dtype in table, "Unsupported str dtype {} to torch dtype".format(str_dtype)
    return table[str_dtype]
