"""Host-side helpers shared by the quantizers (mirrors neural_compressor/torch/utils/utility.py)."""
from .utility import (Mode, algos_mapping, register_algo, fetch_module, set_module, get_quantizer, postprocess_model,
                      get_model_device, get_block_prefix, find_layers, current_device, logger, move_to_device,
                      dump_model_op_stats)
