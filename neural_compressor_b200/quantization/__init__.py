"""Same import surface as `neural_compressor.torch.quantization`."""
from . import algorithm_entry  # registers rtn / gptq / awq / smooth_quant in algos_mapping
from .config import (AWQConfig, BaseConfig, ComposableConfig, GPTQConfig, RTNConfig, SmoothQuantConfig, get_default_awq_config,
                     get_default_double_quant_config, get_default_gptq_config, get_default_rtn_config, get_default_sq_config,
                     get_model_info)
from .quantize import convert, prepare, quantize
from ..algorithms.save_load import load, save
