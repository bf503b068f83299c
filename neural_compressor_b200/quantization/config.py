"""Configuration surface of the weight-only path, mirroring the reference's names, argument meaning and
defaults so user scripts keep working unchanged:

    RTNConfig          neural_compressor/torch/quantization/config.py:119-187
    GPTQConfig         config.py:322-424
    AWQConfig          config.py:525-609
    SmoothQuantConfig  config.py:1485-1569
    BaseConfig         neural_compressor/common/base_config.py:190- (set_local :297-316, to_config_mapping
                       :586-617, to_dict/from_dict :350-425)

Implemented: global + local configs, regex / module-type / white-list matching, dict and json round trips, `+`
(same class: merge of the local entries; different classes: `ComposableConfig`), multi-algorithm dict configs -- the
operations of the reference's test/torch/test_config.py, checked side by side with the live reference in
tests/test_config_cpu.py.  The tuning-grid expansion of list-valued parameters (`expand`, autotune) is out of scope
(SURVEY §2.1 row 7).
"""
from __future__ import annotations

import re
from collections import OrderedDict
from typing import List, Optional, Tuple

import torch

RTN, GPTQ, AWQ, SMOOTH_QUANT = "rtn", "gptq", "awq", "smooth_quant"
LM_HEAD_NAMES = [".*lm_head", ".*output_layer", ".*embed_out"]  # torch/utils/constants.py:69
DEFAULT_WHITE_LIST = "*"


def _woq_white_list():
    types = [torch.nn.Linear]
    try:
        import transformers

        types.append(transformers.Conv1D)
    except Exception:  # pragma: no cover
        pass
    return tuple(types)


class BaseConfig:
    """Global parameters + per-operator overrides (op-name regex or module type)."""

    name = "base_config"
    params_list: List[str] = []

    def __init__(self, white_list=DEFAULT_WHITE_LIST):
        self._global_config = None
        self._local_config: "OrderedDict[Union[str, Callable], BaseConfig]" = OrderedDict()
        self._white_list = white_list

    def _post_init(self):
        # the global config is the object itself when the white list is "*" (base_config.py:253-270)
        self._global_config = self if self._white_list == DEFAULT_WHITE_LIST else None
        if self._white_list != DEFAULT_WHITE_LIST and self._white_list:
            import copy

            for op in self._white_list:
                clone = copy.copy(self)
                clone._local_config = OrderedDict()
                clone._white_list = DEFAULT_WHITE_LIST
                clone._global_config = clone
                self._local_config[op] = clone

    @property
    def global_config(self):
        return self._global_config

    @property
    def local_config(self):
        return self._local_config

    def set_local(self, operator_name_or_list, config: "BaseConfig") -> "BaseConfig":
        """base_config.py:297-316."""
        names = operator_name_or_list if isinstance(operator_name_or_list, list) else [operator_name_or_list]
        for n in names:
            self._local_config[n] = config
        return self

    # ---- dict round trip (base_config.py:350-425) ----
    def to_dict(self):
        result = {}
        global_config = {k: getattr(self, k) for k in self.params_list}
        if self._local_config:
            result["local"] = {}
            for op, cfg in self._local_config.items():
                result["local"][op] = cfg.to_dict()     # a module type stays the key it was given as (base_config.py:326-329)
            if self._global_config is not None:
                result["global"] = global_config
        else:
            result = global_config
        return result

    @classmethod
    def from_dict(cls, config_dict):
        if "global" not in config_dict and "local" not in config_dict:
            return cls(**config_dict)
        cfg = cls(**config_dict.get("global", {}))
        for op, sub in config_dict.get("local", {}).items():
            cfg.set_local(op, cls(**sub))
        return cfg

    def __repr__(self):
        return f"{self.__class__.__name__}({self.to_dict()})"

    # ---- json (base_config.py:427-452) ----
    def to_json_string(self, use_diff: bool = False) -> str:
        import json

        return json.dumps(self.to_dict(), indent=2) + "\n"

    def to_json_file(self, filename):
        with open(filename, "w", encoding="utf-8") as f:
            f.write(self.to_json_string())

    @classmethod
    def from_json_file(cls, filename):
        import json

        with open(filename, "r", encoding="utf-8") as f:
            return cls.from_dict(json.load(f))

    # ---- composition (base_config.py:454-471) ----
    def __add__(self, other: "BaseConfig") -> "BaseConfig":
        """Same class: the other config's local entries are merged into this one (its global part is dropped).
        Different classes: a `ComposableConfig` holding both."""
        if isinstance(other, type(self)):
            for op_name, config in other.local_config.items():
                self.set_local(op_name, config)
            return self
        return ComposableConfig(configs=[self, other])

    # ---- model walk + mapping ----
    @staticmethod
    def get_model_info(model: torch.nn.Module) -> List[Tuple[str, str]]:
        """config.py:252-267: every whitelisted module as (name, type name)."""
        wl = _woq_white_list()
        return [(n, type(m).__name__) for n, m in model.named_modules() if isinstance(m, wl)]

    def to_config_mapping(self, config_list=None, model_info=None) -> "OrderedDict[Tuple[str, str], BaseConfig]":
        """base_config.py:586-617: global, then op-type, then op-name regex (re.match) overrides."""
        mapping = OrderedDict()
        for config in (config_list or [self]):
            by_type, by_name = {}, {}
            for key, sub in config.local_config.items():
                if isinstance(key, str) and not _looks_like_type_name(key):
                    by_name[key] = sub
                else:
                    by_type[key if isinstance(key, str) else key.__name__] = sub
            for op_name, op_type in model_info:
                if self.global_config is not None:
                    mapping[(op_name, op_type)] = config.global_config
                if op_type in by_type:
                    mapping[(op_name, op_type)] = by_type[op_type]
                for pattern, sub in by_name.items():
                    if re.match(pattern, op_name):
                        mapping[(op_name, op_type)] = sub
        return mapping


class ComposableConfig(BaseConfig):
    """Several algorithms' configs applied to one model, built with `+` or from a multi-key dict
    (base_config.py:684-830).  The mapping takes, per member config, the entries its own `to_config_mapping` yields for
    the modules that member's `get_model_info` reports (the reference's composable mapping applies the LOCAL entries only,
    :794-816; the per-algorithm entry functions then pick the operators carrying their own config class)."""

    name = "composable_config"

    def __init__(self, configs):
        self.config_list = list(configs)

    def __add__(self, other):
        if isinstance(other, ComposableConfig):
            self.config_list.extend(other.config_list)
        else:
            self.config_list.append(other)
        return self

    def to_dict(self):
        return {config.name: config.to_dict() for config in self.config_list}

    @classmethod
    def from_dict(cls, config_dict, config_registry=None):
        registry = config_registry or {c.name: c for c in (RTNConfig, GPTQConfig, AWQConfig, SmoothQuantConfig)}
        assert len(config_dict) >= 1, "The config dict must include at least one configuration."
        config = None
        for name, value in config_dict.items():
            part = registry[name].from_dict(value)
            config = part if config is None else config + part
        return config

    def to_json_string(self, use_diff: bool = False) -> str:
        import json

        return json.dumps(self.to_dict(), indent=2) + "\n"

    def __repr__(self):
        return f"{self.__class__.__name__} {self.to_json_string()}"

    def get_model_info(self, model, *args, **kwargs):
        return {config.name: config.get_model_info(model, *args, **kwargs) for config in self.config_list}

    def to_config_mapping(self, config_list=None, model_info=None):
        mapping = OrderedDict()
        for config in self.config_list:
            info = model_info.get(config.name) if isinstance(model_info, dict) else model_info
            by_type, by_name = {}, {}
            for key, sub in config.local_config.items():
                if isinstance(key, str) and not _looks_like_type_name(key):
                    by_name[key] = sub
                else:
                    by_type[key if isinstance(key, str) else key.__name__] = sub
            for op_name, op_type in info:
                if op_type in by_type:
                    mapping[(op_name, op_type)] = by_type[op_type]
                for pattern, sub in by_name.items():
                    if re.match(pattern, op_name):
                        mapping[(op_name, op_type)] = sub
        return mapping


def get_model_info(model: torch.nn.Module, white_module_list=None) -> List[Tuple[str, str]]:
    """torch/utils/utility.py `get_model_info`: (name, type name) of every module of the listed types, first occurrence
    only (a module object reachable under two names is reported once)."""
    white = tuple(white_module_list) if white_module_list is not None else _woq_white_list()
    seen, out = set(), []
    for name, module in model.named_modules():
        if isinstance(module, white) and (name, type(module).__name__) not in seen:
            seen.add((name, type(module).__name__))
            out.append((name, type(module).__name__))
    return out


def _looks_like_type_name(key: str) -> bool:
    return key in ("Linear", "Conv1D", "Conv1d", "Conv2d", "Conv3d")


class _WOQConfig(BaseConfig):
    """Shared behaviour of the three weight-only configs: lm_head is left in fp32 unless quant_lm_head."""

    def to_config_mapping(self, config_list=None, model_info=None):
        if not self.quant_lm_head:
            kw = dict(dtype="fp32", use_layer_wise=self.use_layer_wise, model_path=self.model_path)
            if "use_block_wise" in self.params_list:
                kw["use_block_wise"] = self.use_block_wise
            self.set_local(LM_HEAD_NAMES, type(self)(**kw))
        return super().to_config_mapping(config_list, model_info)


class RTNConfig(_WOQConfig):
    """config.py:119-187."""

    name = RTN
    params_list = ["dtype", "bits", "use_sym", "group_size", "group_dim", "use_full_range", "use_mse_search",
                   "use_layer_wise", "model_path", "use_double_quant", "double_quant_dtype", "double_quant_bits",
                   "double_quant_use_sym", "double_quant_group_size", "quant_lm_head"]

    def __init__(self, dtype="int", bits=4, use_sym=True, group_size=32, group_dim=1, use_full_range=False,
                 use_mse_search=False, use_layer_wise=True, model_path="", use_double_quant=False,
                 double_quant_dtype="int", double_quant_bits=8, double_quant_use_sym=False,
                 double_quant_group_size=256, quant_lm_head=False, white_list=DEFAULT_WHITE_LIST, **kwargs):
        super().__init__(white_list=white_list)
        self.dtype, self.bits, self.use_sym, self.group_size, self.group_dim = dtype, bits, use_sym, group_size, group_dim
        self.use_full_range, self.use_mse_search = use_full_range, use_mse_search
        self.use_layer_wise, self.model_path = use_layer_wise, model_path
        self.use_double_quant, self.double_quant_dtype = use_double_quant, double_quant_dtype
        self.double_quant_bits, self.double_quant_use_sym = double_quant_bits, double_quant_use_sym
        self.double_quant_group_size, self.quant_lm_head = double_quant_group_size, quant_lm_head
        self._post_init()


class GPTQConfig(_WOQConfig):
    """config.py:322-424 (block_size default 2048 at :356)."""

    name = GPTQ
    params_list = ["dtype", "bits", "use_sym", "group_size", "use_mse_search", "use_layer_wise", "use_block_wise",
                   "model_path", "use_double_quant", "double_quant_dtype", "double_quant_bits", "double_quant_use_sym",
                   "double_quant_group_size", "quant_lm_head", "act_order", "hybrid_order", "fp8_aware", "percdamp",
                   "block_size", "static_groups", "true_sequential"]

    def __init__(self, dtype="int", bits=4, use_sym=True, group_size=32, use_mse_search=False, use_layer_wise=False,
                 use_block_wise=False, model_path="", use_double_quant=False, double_quant_dtype="int",
                 double_quant_bits=8, double_quant_use_sym=False, double_quant_group_size=256, quant_lm_head=False,
                 act_order=False, hybrid_order=False, fp8_aware=False, percdamp=0.01, block_size=2048,
                 static_groups=False, true_sequential=False, white_list=DEFAULT_WHITE_LIST, **kwargs):
        super().__init__(white_list=white_list)
        self.dtype, self.bits, self.use_sym, self.group_size = dtype, bits, use_sym, group_size
        self.use_mse_search, self.use_layer_wise, self.use_block_wise = use_mse_search, use_layer_wise, use_block_wise
        self.model_path = model_path
        self.use_double_quant, self.double_quant_dtype = use_double_quant, double_quant_dtype
        self.double_quant_bits, self.double_quant_use_sym = double_quant_bits, double_quant_use_sym
        self.double_quant_group_size, self.quant_lm_head = double_quant_group_size, quant_lm_head
        self.act_order, self.hybrid_order, self.fp8_aware = act_order, hybrid_order, fp8_aware
        self.percdamp, self.block_size = percdamp, block_size
        self.static_groups, self.true_sequential = static_groups, true_sequential
        self._post_init()


class AWQConfig(_WOQConfig):
    """config.py:525-609."""

    name = AWQ
    params_list = ["dtype", "bits", "use_sym", "group_size", "group_dim", "use_full_range", "use_mse_search",
                   "use_layer_wise", "model_path", "use_double_quant", "double_quant_dtype", "double_quant_bits",
                   "double_quant_use_sym", "double_quant_group_size", "quant_lm_head", "use_auto_scale",
                   "use_auto_clip", "folding", "absorb_layer_dict"]

    def __init__(self, dtype="int", bits=4, use_sym=True, group_size=32, group_dim=1, use_full_range=False,
                 use_mse_search=False, use_layer_wise=False, model_path="", use_double_quant=False,
                 double_quant_dtype="int", double_quant_bits=8, double_quant_use_sym=True,
                 double_quant_group_size=256, quant_lm_head=False, use_auto_scale=True, use_auto_clip=True,
                 folding=False, white_list=DEFAULT_WHITE_LIST, absorb_layer_dict: Optional[dict] = None, **kwargs):
        super().__init__(white_list=white_list)
        self.dtype, self.bits, self.use_sym, self.group_size, self.group_dim = dtype, bits, use_sym, group_size, group_dim
        self.use_full_range, self.use_mse_search = use_full_range, use_mse_search
        self.use_layer_wise, self.model_path = use_layer_wise, model_path
        self.use_double_quant, self.double_quant_dtype = use_double_quant, double_quant_dtype
        self.double_quant_bits, self.double_quant_use_sym = double_quant_bits, double_quant_use_sym
        self.double_quant_group_size, self.quant_lm_head = double_quant_group_size, quant_lm_head
        self.use_auto_scale, self.use_auto_clip, self.folding = use_auto_scale, use_auto_clip, folding
        self.absorb_layer_dict = absorb_layer_dict or {}
        self._post_init()


class SmoothQuantConfig(BaseConfig):
    """config.py:1485-1569."""

    name = SMOOTH_QUANT
    params_list = ["w_dtype", "w_sym", "w_granularity", "w_algo", "act_dtype", "act_sym", "act_granularity",
                   "act_algo", "excluded_precisions", "alpha", "folding", "scale_sharing", "auto_alpha_args"]

    def __init__(self, w_dtype="int8", w_sym=True, w_granularity="per_channel", w_algo="minmax", act_dtype="uint8",
                 act_sym=False, act_granularity="per_tensor", act_algo="minmax", excluded_precisions=None, alpha=0.5,
                 folding=False, scale_sharing=False, init_alpha=0.5, alpha_min=0.0, alpha_max=1.0, alpha_step=0.1,
                 shared_criterion="max", do_blockwise=False, auto_alpha_args=None, white_list=DEFAULT_WHITE_LIST,
                 **kwargs):
        super().__init__(white_list=white_list)
        self.w_dtype, self.w_sym, self.w_granularity, self.w_algo = w_dtype, w_sym, w_granularity, w_algo
        self.act_dtype, self.act_sym, self.act_granularity, self.act_algo = act_dtype, act_sym, act_granularity, act_algo
        self.excluded_precisions = excluded_precisions or []
        self.alpha, self.folding, self.scale_sharing = alpha, folding, scale_sharing
        self.auto_alpha_args = auto_alpha_args or dict(init_alpha=init_alpha, alpha_min=alpha_min, alpha_max=alpha_max,
                                                       alpha_step=alpha_step, shared_criterion=shared_criterion,
                                                       do_blockwise=do_blockwise)
        self._post_init()

    @staticmethod
    def get_model_info(model: torch.nn.Module, example_inputs=None):
        return [(n, type(m).__name__) for n, m in model.named_modules() if isinstance(m, torch.nn.Linear)]


def get_default_rtn_config():
    return RTNConfig()


# torch/utils/constants.py:18-42: named double-quantisation presets of RTN
DOUBLE_QUANT_CONFIGS = {
    "BNB_NF4": {"dtype": "nf4", "bits": 4, "group_size": 32, "use_double_quant": True, "double_quant_bits": 8,
                "double_quant_dtype": "int", "double_quant_use_sym": False, "double_quant_group_size": 256},
    "GGML_TYPE_Q4_K": {"dtype": "int", "bits": 4, "use_sym": False, "group_size": 32, "use_double_quant": True,
                       "double_quant_bits": 6, "double_quant_dtype": "int", "double_quant_use_sym": True,
                       "double_quant_group_size": 8},
}


def get_default_double_quant_config(type="BNB_NF4"):
    """config.py:305-317."""
    assert type in DOUBLE_QUANT_CONFIGS, "Supported double quant configs: {}".format(list(DOUBLE_QUANT_CONFIGS.keys()))
    return RTNConfig.from_dict(DOUBLE_QUANT_CONFIGS[type])


def get_default_gptq_config():
    return GPTQConfig()


def get_default_awq_config():
    return AWQConfig()


def get_default_sq_config():
    return SmoothQuantConfig()
