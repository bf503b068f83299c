"""Public entry points with the reference's signatures (neural_compressor/torch/quantization/quantize.py:
quantize :138-175, prepare :179-218, convert :253-325)."""
import copy
from typing import Any, Callable

import torch

from ..utils import Mode, algos_mapping, logger
from .config import AWQConfig, BaseConfig, ComposableConfig, GPTQConfig, RTNConfig, SmoothQuantConfig

_CONFIG_BY_NAME = {c.name: c for c in (RTNConfig, GPTQConfig, AWQConfig, SmoothQuantConfig)}


def need_apply(configs_mapping, algo_name):
    return any(config.name == algo_name for config in configs_mapping.values())


def _as_config(quant_config):
    if isinstance(quant_config, dict):
        # {"rtn": {...}[, "gptq": {...}]} as produced by (Composable)Config.to_dict()   (quantize.py:81-84)
        return ComposableConfig.from_dict(quant_config, config_registry=_CONFIG_BY_NAME)
    assert isinstance(quant_config, BaseConfig), (
        f"Please pass a dict or config instance as the quantization configuration, but got {type(quant_config)}.")
    return quant_config


def preprocess_quant_config(model, quant_config, mode="prepare", example_inputs=None, run_fn=None):
    quant_config = _as_config(quant_config)
    if isinstance(quant_config, SmoothQuantConfig):
        model_info = quant_config.get_model_info(model, example_inputs)
    else:
        model_info = quant_config.get_model_info(model=model)
    if getattr(quant_config, "model_path", None) == "" and hasattr(model, "name_or_path"):
        quant_config.model_path = model.name_or_path
    return model, quant_config.to_config_mapping(model_info=model_info)


def quantize(model: torch.nn.Module, quant_config: BaseConfig, run_fn: Callable = None, run_args: Any = None,
             inplace: bool = True, example_inputs: Any = None) -> torch.nn.Module:
    q_model = model if inplace else copy.deepcopy(model)
    q_model, configs_mapping = preprocess_quant_config(q_model, quant_config, mode="quantize",
                                                       example_inputs=example_inputs, run_fn=run_fn)
    for algo_name, algo_func in algos_mapping.items():
        if need_apply(configs_mapping, algo_name):
            logger.info(f"Start to apply {algo_name} on the model.")
            q_model = algo_func(q_model, configs_mapping, run_fn=run_fn, run_args=run_args,
                                example_inputs=example_inputs, mode=Mode.QUANTIZE)
    setattr(q_model, "is_quantized", True)
    return q_model


def prepare(model: torch.nn.Module, quant_config: BaseConfig, inplace: bool = True, example_inputs: Any = None):
    prepared_model = model if inplace else copy.deepcopy(model)
    prepared_model, configs_mapping = preprocess_quant_config(prepared_model, quant_config, mode="prepare",
                                                              example_inputs=example_inputs)
    for algo_name, algo_func in algos_mapping.items():
        if need_apply(configs_mapping, algo_name):
            logger.info(f"Start to prepare model with {algo_name}.")
            prepared_model = algo_func(prepared_model, configs_mapping, example_inputs=example_inputs, mode=Mode.PREPARE)
            setattr(prepared_model, "is_prepared", True)
    setattr(prepared_model, "quant_config", quant_config)
    setattr(prepared_model, "example_inputs", example_inputs)
    return prepared_model


def convert(model: torch.nn.Module, quant_config: BaseConfig = None, inplace: bool = True, **kwargs):
    q_model = model if inplace else copy.deepcopy(model)
    assert getattr(model, "is_prepared", False) or quant_config is not None, \
        "Please pass quant_config to convert function."
    if getattr(model, "is_prepared", False):
        if quant_config is None:
            quant_config = model.quant_config
        else:
            logger.warning("quant_config will be ignored since the model has been prepared.")
            quant_config = model.quant_config
    example_inputs = model.example_inputs if getattr(model, "is_prepared", False) else None
    quant_config = _as_config(quant_config)
    if isinstance(quant_config, SmoothQuantConfig):
        model_info = quant_config.get_model_info(q_model, example_inputs)
    else:
        model_info = quant_config.get_model_info(model=q_model)
    configs_mapping = quant_config.to_config_mapping(model_info=model_info)
    for algo_name, algo_func in algos_mapping.items():
        if need_apply(configs_mapping, algo_name):
            logger.info(f"Start to convert model with {algo_name}.")
            q_model = algo_func(q_model, configs_mapping, example_inputs=example_inputs, mode=Mode.CONVERT, **kwargs)
    if hasattr(q_model, "__dict__"):
        setattr(q_model, "is_quantized", True)
    return q_model
