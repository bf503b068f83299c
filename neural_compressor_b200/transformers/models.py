"""`AutoModelForCausalLM.from_pretrained(..., quantization_config=...)` and `convert_to_quantized_model` -- the HF-style
caller of the weight-only path (neural_compressor/transformers/models/modeling_auto.py:62-90,
transformers/quantization/utils.py:337-487).  The mapping of the HF-style config onto the INC config objects follows
utils.py:350-432 field by field; calibration data comes from `config.dataset` (see utils.py here)."""
import os
import types

import torch

from ..quantization import AWQConfig as INCAWQConfig
from ..quantization import GPTQConfig as INCGPTQConfig
from ..quantization import RTNConfig as INCRTNConfig
from ..quantization import convert, load, prepare
from ..utils import logger


def _calibration_batches(config):
    """utils.py:258-334 (default_run_fn) without the hub: an iterable of input_ids tensors, or of strings."""
    ds = config.dataset
    if isinstance(ds, (str, bytes, os.PathLike)):
        raise ValueError("calibration data must be passed as an iterable of input_ids tensors (or strings + tokenizer): "
                         f"hub datasets ('{ds}') cannot be fetched on an offline B200 box")
    n = 0
    for item in ds:
        if n >= config.n_samples:
            break
        if isinstance(item, str):
            assert config.tokenizer is not None, "Please provide the tokenizer in quantization_config."
            item = config.tokenizer(item, return_tensors="pt")["input_ids"]
        if isinstance(item, dict):
            item = item["input_ids"]
        item = torch.as_tensor(item)
        if item.dim() == 1:
            item = item[None]
        yield item[:, :config.seq_len]
        n += 1


def convert_to_quantized_model(model, config, device="cuda"):
    """utils.py:337-487 for the rtn / gptq / awq methods."""
    dtype = "int4" if config.weight_dtype == "int4_fullrange" else config.weight_dtype
    method = getattr(config.quant_method, "value", config.quant_method)
    model = model.to(device)

    def exclude(qc, cls):
        for module in config.modules_to_not_convert:
            qc.set_local(".*" + module, cls(dtype="fp32"))
        return qc

    def run_fn(m):
        for ids in _calibration_batches(config):
            m(ids.to(device))

    if method == "rtn":
        qc = exclude(INCRTNConfig(dtype=dtype, bits=config.bits, use_sym=config.sym, group_size=config.group_size,
                                  use_layer_wise=bool(config.use_layer_wise), model_path=config.model_path,
                                  quant_lm_head=config.quant_lm_head), INCRTNConfig)
        logger.info(f"Do RTN algorithm with config {qc}")
        model = convert(prepare(model, qc))
    elif method == "gptq":
        model.seqlen = config.seq_len
        qc = exclude(INCGPTQConfig(dtype=dtype, bits=config.bits, use_sym=config.sym, group_size=config.group_size,
                                   use_layer_wise=bool(config.use_layer_wise), model_path=config.model_path,
                                   act_order=config.desc_act, percdamp=config.damp_percent, block_size=config.blocksize,
                                   static_groups=config.static_groups, use_mse_search=config.use_mse_search,
                                   true_sequential=config.true_sequential, quant_lm_head=config.quant_lm_head),
                     INCGPTQConfig)
        logger.info(f"Do GPTQ algorithm with config {qc}")
        model = prepare(model=model, quant_config=qc)
        run_fn(model)
        model = convert(model)
    elif method == "awq":
        qc = exclude(INCAWQConfig(dtype=dtype, bits=config.bits, use_sym=config.sym, group_size=config.group_size,
                                  use_layer_wise=bool(config.use_layer_wise), use_auto_scale=config.auto_scale,
                                  use_auto_clip=config.auto_clip, absorb_layer_dict=config.absorb_layer_dict,
                                  quant_lm_head=config.quant_lm_head), INCAWQConfig)
        logger.info(f"Do AWQ algorithm with config {qc}")
        example_inputs = torch.ones([1, 512], dtype=torch.long).to(device)
        model = prepare(model=model, quant_config=qc, example_inputs=example_inputs)
        run_fn(model)
        model = convert(model)
    else:
        raise NotImplementedError(f"quant_method {method} is outside the B200 hot path (rtn | gptq | awq)")
    return model


def _save_pretrained(self, save_directory, **kwargs):
    """modeling_auto.py:62-90 binds a save_low_bit-style method; here it is the HuggingFace layout of save_load.py."""
    from ..algorithms.save_load import save

    save(self, save_directory, format="huggingface", **kwargs)


class AutoModelForCausalLM:
    """`from_pretrained(path, quantization_config=RtnConfig|GPTQConfig|AwqConfig)` quantises on the B200 and returns
    the packed model; a directory that already holds a quantised checkpoint is loaded as such."""

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs):
        import transformers

        quantization_config = kwargs.pop("quantization_config", None)
        device = kwargs.pop("device_map", "cuda")
        device = "cuda" if device in ("auto", None) else device
        path = pretrained_model_name_or_path
        if quantization_config is None and os.path.isdir(str(path)) and (
                os.path.exists(os.path.join(path, "quantize_config.json"))):
            return load(path, format="huggingface", device=device, **kwargs)
        model = transformers.AutoModelForCausalLM.from_pretrained(path, *model_args, **kwargs)
        model.eval()
        if quantization_config is None:
            return model.to(device)
        if getattr(quantization_config, "tokenizer", None) is None and hasattr(quantization_config, "tokenizer"):
            quantization_config.tokenizer = kwargs.get("tokenizer")
        model.config.use_cache = False if getattr(quantization_config.quant_method, "value", "") != "rtn" else \
            model.config.use_cache
        model = convert_to_quantized_model(model, quantization_config, device=device)
        model.quantization_config = quantization_config
        model.save_pretrained_fp = model.save_pretrained
        model.save_pretrained = types.MethodType(_save_pretrained, model)
        return model
