"""Quantization configs of the HF-style entry -- same names, arguments and defaults as
neural_compressor/transformers/utils/quantization_config.py (RtnConfig :242-295, GPTQConfig :297-385,
AwqConfig :387-455).  One deliberate difference: there is no network on a calibration box, so `dataset` may be an
iterable of `input_ids` tensors (or of strings when a tokenizer is given) besides a hub dataset name."""
import copy
from enum import Enum
from typing import Any


class QuantizationMethod(str, Enum):
    RTN = "rtn"
    GPTQ = "gptq"
    AWQ = "awq"


class _ConfigMixin:
    def _common(self, bits, group_size, compute_dtype, scale_dtype, use_layer_wise, quant_lm_head, kwargs):
        self.bits = bits
        self.group_size = group_size
        self.compute_dtype = compute_dtype
        self.scale_dtype = scale_dtype
        self.weight_dtype = "int4" if bits == 4 else "int8"
        self.use_layer_wise = use_layer_wise
        self.model_path = kwargs.get("model_path", "")
        self.quant_lm_head = quant_lm_head
        self.modules_to_not_convert = [] if quant_lm_head else kwargs.get(
            "modules_to_not_convert", ["lm_head", "transformer.output_layer", "embed_out"])
        self.device = kwargs.get("device", "auto")

    def to_dict(self):
        out = {}
        for k, v in copy.copy(self.__dict__).items():
            if k in ("tokenizer", "dataset"):
                continue
            out[k] = v.value if isinstance(v, Enum) else v
        return out

    def __repr__(self):
        return f"{type(self).__name__}({self.to_dict()})"


class RtnConfig(_ConfigMixin):
    def __init__(self, bits: int = 4, group_size: int = 32, compute_dtype: Any = None, scale_dtype: Any = None,
                 sym: bool = True, use_layer_wise: bool = None, quant_lm_head: bool = False, **kwargs):
        self.quant_method = QuantizationMethod.RTN
        self._common(bits, group_size, compute_dtype, scale_dtype, use_layer_wise, quant_lm_head, kwargs)
        self.sym = sym
        self.scheme = "sym" if sym else "asym"


class GPTQConfig(_ConfigMixin):
    def __init__(self, bits: int = 4, tokenizer: Any = None, dataset: Any = "NeelNanda/pile-10k", batch_size: int = 8,
                 group_size: int = 32, compute_dtype: Any = None, scale_dtype: Any = None, sym: bool = True,
                 blocksize: int = 128, damp_percent: float = 0.1, desc_act: bool = False, n_samples: int = 128,
                 seq_len: int = 2048, static_groups: bool = False, use_mse_search: bool = False,
                 true_sequential: bool = False, use_layer_wise: bool = None, quant_lm_head: bool = False, **kwargs):
        self.quant_method = QuantizationMethod.GPTQ
        self._common(bits, group_size, compute_dtype, scale_dtype, use_layer_wise, quant_lm_head, kwargs)
        self.tokenizer, self.dataset, self.batch_size = tokenizer, dataset, batch_size
        self.sym = sym
        self.scheme = "sym" if sym else "asym"
        self.blocksize, self.damp_percent, self.desc_act = blocksize, damp_percent, desc_act
        self.n_samples, self.seq_len = n_samples, seq_len
        self.static_groups, self.use_mse_search, self.true_sequential = static_groups, use_mse_search, true_sequential
        if self.bits not in [4, 8]:  # quantization_config.py:355-361
            raise ValueError(f"Only support quantization to [4, 8] bits but found {self.bits}")
        if not (0 < self.damp_percent < 1):
            raise ValueError("damp_percent must between 0 and 1.")


class AwqConfig(_ConfigMixin):
    def __init__(self, bits: int = 4, tokenizer: Any = None, dataset: Any = "NeelNanda/pile-10k", group_size: int = 32,
                 compute_dtype: Any = None, weight_dtype: Any = None, scale_dtype: Any = None, use_layer_wise: bool = None,
                 n_samples: int = 128, seq_len: int = 2048, auto_scale: bool = True, auto_clip: bool = True,
                 zero_point: bool = True, absorb_layer_dict: dict = {}, quant_lm_head: bool = False, backend: str = None,
                 **kwargs):
        self.quant_method = QuantizationMethod.AWQ
        self._common(bits, group_size, compute_dtype, scale_dtype, use_layer_wise, quant_lm_head, kwargs)
        self.tokenizer, self.dataset = tokenizer, dataset
        self.n_samples, self.seq_len = n_samples, seq_len
        self.auto_scale, self.auto_clip, self.zero_point = auto_scale, auto_clip, zero_point
        self.absorb_layer_dict, self.backend = absorb_layer_dict, backend
        self.scheme = "asym" if zero_point else "sym"
        self.sym = not zero_point
        self.batch_size = kwargs.pop("batch_size", 8)
