"""Algorithm entries registered in `algos_mapping` -- plug-in seam #1 of the reference
(neural_compressor/torch/quantization/algorithm_entry.py: rtn_entry :63-117, gptq_entry :121-186,
awq_quantize_entry :406-487, smooth_quant_entry :336-402).  Same signature and side effects:
entry(model, configs_mapping, mode, *args, **kwargs) -> model, sets model.qconfig and binds model.save.
"""
from types import MethodType
from typing import Dict, Tuple

import torch

from ..utils import Mode, dump_model_op_stats, get_quantizer, logger, postprocess_model, register_algo
from .config import AWQ, GPTQ, RTN, SMOOTH_QUANT, AWQConfig, GPTQConfig, RTNConfig, SmoothQuantConfig


def _bind_save(model):
    """algorithm_entry.py:110-116 of the reference (model.save) + B200 inference set-up: packed q/k/v and gate/up
    siblings get one fused batch-1 launch (modules.SiblingGroup; no effect until the model is converted)."""
    from ..algorithms.modules import fuse_sibling_linears
    from ..algorithms.save_load import save

    model.save = MethodType(save, model)
    fuse_sibling_linears(model)


@register_algo(RTN)
@torch.no_grad()
def rtn_entry(model: torch.nn.Module, configs_mapping: Dict[Tuple[str, str], RTNConfig], mode: Mode = Mode.QUANTIZE,
              *args, **kwargs) -> torch.nn.Module:
    from ..algorithms.rtn import RTNQuantizer

    weight_config = {}
    quant_config = None
    for (op_name, op_type), quant_config in configs_mapping.items():
        if quant_config.name != RTN:
            continue
        weight_config[op_name] = {
            "dtype": quant_config.dtype, "bits": quant_config.bits,
            "scheme": "sym" if quant_config.use_sym else "asym", "group_size": quant_config.group_size,
            "group_dim": quant_config.group_dim, "use_full_range": quant_config.use_full_range,
            "use_mse_search": quant_config.use_mse_search, "use_double_quant": quant_config.use_double_quant,
            "double_quant_dtype": quant_config.double_quant_dtype, "double_quant_bits": quant_config.double_quant_bits,
            "double_quant_scheme": "sym" if quant_config.double_quant_use_sym else "asym",
            "double_quant_group_size": quant_config.double_quant_group_size,
        }
    if quant_config is not None:
        kwargs.update({"use_layer_wise": quant_config.use_layer_wise, "model_path": quant_config.model_path,
                       "quant_lm_head": quant_config.quant_lm_head})
    quantizer = get_quantizer(model, quantizer_cls=RTNQuantizer, quant_config=weight_config)
    model = quantizer.execute(model, mode=mode, *args, **kwargs)
    model.qconfig = configs_mapping
    _bind_save(model)
    postprocess_model(model, mode, quantizer)
    dump_model_op_stats(mode, configs_mapping)
    return model


@register_algo(GPTQ)
@torch.no_grad()
def gptq_entry(model: torch.nn.Module, configs_mapping: Dict[Tuple[str, str], GPTQConfig], mode: Mode = Mode.QUANTIZE,
               *args, **kwargs) -> torch.nn.Module:
    from ..algorithms.gptq import GPTQuantizer

    weight_config = {}
    quant_config = None
    for (op_name, op_type), quant_config in configs_mapping.items():
        if quant_config.name != GPTQ or quant_config.dtype == "fp32":
            continue
        weight_config[op_name] = {
            "dtype": quant_config.dtype, "bits": quant_config.bits, "sym": quant_config.use_sym,
            "group_size": quant_config.group_size, "mse": quant_config.use_mse_search,
            "use_double_quant": quant_config.use_double_quant, "double_quant_dtype": quant_config.double_quant_dtype,
            "double_quant_bits": quant_config.double_quant_bits, "double_quant_sym": quant_config.double_quant_use_sym,
            "double_quant_group_size": quant_config.double_quant_group_size, "act_order": quant_config.act_order,
            "hybrid_order": quant_config.hybrid_order, "fp8_aware": quant_config.fp8_aware,
            "percdamp": quant_config.percdamp, "block_size": quant_config.block_size,
            "static_groups": quant_config.static_groups, "true_sequential": quant_config.true_sequential,
        }
    if quant_config is not None:
        kwargs.update({"use_layer_wise": quant_config.use_layer_wise, "use_block_wise": quant_config.use_block_wise,
                       "model_path": quant_config.model_path, "quant_lm_head": quant_config.quant_lm_head})
    kwargs.pop("example_inputs", None)
    logger.info("lm_head in transformer model is skipped by GPTQ")
    quantizer = get_quantizer(model, quantizer_cls=GPTQuantizer, quant_config=weight_config)
    model = quantizer.execute(model, mode=mode, *args, **kwargs)
    model.qconfig = configs_mapping
    _bind_save(model)
    postprocess_model(model, mode, quantizer)
    dump_model_op_stats(mode, configs_mapping)
    return model


@register_algo(AWQ)
@torch.no_grad()
def awq_quantize_entry(model: torch.nn.Module, configs_mapping: Dict[Tuple[str, str], AWQConfig],
                       mode: Mode = Mode.QUANTIZE, *args, **kwargs) -> torch.nn.Module:
    from ..algorithms.awq import AWQQuantizer

    weight_config = {}
    use_auto_scale = use_mse_search = True
    folding = use_full_range = False
    absorb_layer_dict = {}
    for (op_name, op_type), quant_config in configs_mapping.items():
        if quant_config.name != AWQ:
            continue
        if quant_config.dtype == "fp32":
            weight_config[op_name] = {"bits": -1, "dtype": "fp32", "group_size": 128, "scheme": "asym"}
        else:
            weight_config[op_name] = {
                "dtype": quant_config.dtype, "bits": quant_config.bits, "group_size": quant_config.group_size,
                "group_dim": quant_config.group_dim, "scheme": "sym" if quant_config.use_sym else "asym",
                "use_full_range": quant_config.use_full_range, "use_mse_search": quant_config.use_mse_search,
                "use_layer_wise": quant_config.use_layer_wise, "use_double_quant": quant_config.use_double_quant,
                "double_quant_dtype": quant_config.double_quant_dtype, "double_quant_bits": quant_config.double_quant_bits,
                "double_quant_scheme": quant_config.double_quant_use_sym,
                "double_quant_group_size": quant_config.double_quant_group_size,
            }
            use_auto_scale = quant_config.use_auto_scale
            use_mse_search = quant_config.use_auto_clip  # awq clip (algorithm_entry.py:456)
            folding = quant_config.folding
            use_full_range = quant_config.use_full_range
            absorb_layer_dict = quant_config.absorb_layer_dict
    run_fn = kwargs.get("run_fn", None)
    run_args = kwargs.get("run_args", None)
    example_inputs = kwargs.get("example_inputs", None)
    assert example_inputs is not None, "Please provide example_inputs for AWQ quantization."
    quantizer = get_quantizer(model, quantizer_cls=AWQQuantizer, quant_config=weight_config,
                              absorb_layer_dict=absorb_layer_dict)
    model = quantizer.execute(model, mode=mode, bits=-1, example_inputs=example_inputs, run_fn=run_fn,
                              run_args=run_args, use_auto_scale=use_auto_scale, use_mse_search=use_mse_search,
                              folding=folding, use_full_range=use_full_range)
    model.qconfig = configs_mapping
    _bind_save(model)
    postprocess_model(model, mode, quantizer)
    dump_model_op_stats(mode, configs_mapping)
    return model


@register_algo(SMOOTH_QUANT)
@torch.no_grad()
def smooth_quant_entry(model: torch.nn.Module, configs_mapping: Dict[Tuple[str, str], SmoothQuantConfig],
                       mode: Mode = Mode.QUANTIZE, *args, **kwargs) -> torch.nn.Module:
    from ..algorithms.smooth_quant import SmoothQuantQuantizer

    quant_config = None
    for _, quant_config in configs_mapping.items():
        if quant_config.name == SMOOTH_QUANT:
            break
    run_fn = kwargs.get("run_fn", None)
    example_inputs = kwargs.get("example_inputs", None)
    assert example_inputs is not None, "Please provide example_inputs for smooth quantization."
    quantizer = get_quantizer(model, quantizer_cls=SmoothQuantQuantizer, quant_config=quant_config)
    model = quantizer.execute(model, mode=mode, run_fn=run_fn, example_inputs=example_inputs,
                              run_args=kwargs.get("run_args", None))
    model.qconfig = configs_mapping
    postprocess_model(model, mode, quantizer)
    return model
