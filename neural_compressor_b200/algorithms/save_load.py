"""save / load of quantised models in the reference's two on-disk formats
(neural_compressor/torch/algorithms/weight_only/save_load.py:56-108 save, :111-285 load, :329-383 HF loader,
:1094-1156 change_config_to_hf_format; file names torch/utils/utility.py:56,60):

* ``default``     -- `quantized_weight.pt` (torch.save(state_dict)) + `qconfig.json`
* ``huggingface`` -- `save_pretrained` shards (safetensors) + `config.json` carrying an AutoGPTQ-style
                     `quantization_config` + `quantize_config.json`

The packed buffers keep the reference's names/dtypes/shapes, so a checkpoint written here is loadable by the
reference's WOQModelLoader (and by HF/vLLM GPTQ loaders) and vice versa.
"""
import glob
import json
import os
import re

import torch

WEIGHT_NAME = "quantized_weight.pt"
QCONFIG_NAME = "qconfig.json"
HF_QUANT_CONFIG_NAME = "quantize_config.json"
LM_HEAD_NAMES = [".*lm_head", ".*output_layer", ".*embed_out"]  # torch/utils/constants.py:69


def change_config_to_hf_format(config_mappings):
    """save_load.py:1094-1156: one AutoGPTQ-style dict for the whole model (the reference writes RTN, GPTQ, AWQ and
    TEQ checkpoints all as quant_method "gptq").  Every quantised module must share bits / group_size / sym (and the
    GPTQ-only knobs); a quantised lm_head cannot be expressed and raises ValueError."""
    out = {"bits": 4, "group_size": 128, "damp_percent": 0.01, "desc_act": True, "sym": True, "true_sequential": True,
           "model_name_or_path": None, "model_file_base_name": "model", "quant_method": "gptq"}
    seen = None
    for (name, _type), cfg in config_mappings.items():
        if any(re.match(pat, name) for pat in LM_HEAD_NAMES):
            if cfg.dtype != "fp32":
                raise ValueError(f"{name} should not be quantized if you want to save in huggingface format.")
            continue
        cur = dict(bits=cfg.bits, group_size=cfg.group_size, sym=cfg.use_sym,
                   damp_percent=getattr(cfg, "percdamp", 0), desc_act=getattr(cfg, "act_order", False),
                   true_sequential=getattr(cfg, "true_sequential", False))
        if seen is None:
            seen = cur
            continue
        for k in ("bits", "sym", "group_size"):
            assert seen[k] == cur[k], f"{k} should be the same for all modules, got {seen[k]} and {cur[k]}."
        for k, attr in (("damp_percent", "percdamp"), ("desc_act", "act_order"), ("true_sequential", "true_sequential")):
            if hasattr(cfg, attr):
                assert seen[k] == cur[k], f"{attr} should be the same for all modules, got {seen[k]} and {cur[k]}."
    if seen is not None:
        out.update(seen)
    return out


def _save_huggingface(model, output_dir, **kwargs):
    """save_load.py:72-92."""
    if not hasattr(model.config, "quantization_config"):
        model.config.quantization_config = change_config_to_hf_format(model.qconfig)
    quantization_config = model.config.quantization_config
    if not isinstance(quantization_config, dict):
        quantization_config = quantization_config.to_dict()
    # the class's method: `model.save_pretrained` may itself be bound to this function (transformers entry)
    type(model).save_pretrained(model, output_dir, max_shard_size=kwargs.get("max_shard_size", "5GB"),
                                safe_serialization=kwargs.get("safe_serialization", True))
    with open(os.path.join(output_dir, HF_QUANT_CONFIG_NAME), "w", encoding="utf-8") as f:
        json.dump(quantization_config, f, indent=2)
    if getattr(model, "generation_config", None) is not None:
        model.generation_config.save_pretrained(output_dir)
    if kwargs.get("tokenizer") is not None:
        kwargs["tokenizer"].save_pretrained(output_dir)


def _load_huggingface(model_dir, device="cuda", **kwargs):
    """save_load.py:329-383 in outline: model class + config from the directory, an uninitialised model, every Linear
    that has a `<name>.qweight` in the checkpoint becomes a packed module, then the shards are loaded."""
    from transformers import AutoConfig, AutoModelForCausalLM

    from ..utils import set_module
    from .modules import B200WeightOnlyLinear

    config = AutoConfig.from_pretrained(model_dir, trust_remote_code=kwargs.get("trust_remote_code", False))
    qc = getattr(config, "quantization_config", None)
    if qc is None:
        with open(os.path.join(model_dir, HF_QUANT_CONFIG_NAME)) as f:
            qc = json.load(f)
    if not isinstance(qc, dict):
        qc = qc.to_dict()
    if qc.get("quant_method", "gptq") not in ("gptq", "awq", "rtn", "teq"):
        raise NotImplementedError(f"quant_method {qc.get('quant_method')} (SURVEY §8 f1: AutoGPTQ layout only)")
    bits, group_size = int(qc["bits"]), int(qc["group_size"])
    # checkpoint tensors
    state = {}
    files = sorted(glob.glob(os.path.join(model_dir, "*.safetensors")))
    if files:
        from safetensors.torch import load_file

        for fn in files:
            state.update(load_file(fn, device=str(device)))
    else:
        for fn in sorted(glob.glob(os.path.join(model_dir, "pytorch_model*.bin"))):
            state.update(torch.load(fn, map_location=device, weights_only=True))
    assert state, f"no weight shards under {model_dir}"
    # an uninitialised model of the right architecture (the stored config must not re-trigger HF's own GPTQ loader)
    plain = AutoConfig.from_pretrained(model_dir, trust_remote_code=kwargs.get("trust_remote_code", False))
    if hasattr(plain, "quantization_config"):
        delattr(plain, "quantization_config")
    try:
        from transformers.initialization import no_init_weights
    except ImportError:  # older transformers
        from transformers.modeling_utils import no_init_weights
    dtype = kwargs.get("torch_dtype", None) or getattr(config, "torch_dtype", None) or getattr(config, "dtype", None)
    with no_init_weights():
        model = AutoModelForCausalLM.from_config(plain, trust_remote_code=kwargs.get("trust_remote_code", False))
    if isinstance(dtype, str):
        dtype = getattr(torch, dtype.replace("torch.", ""))
    if dtype is not None:
        model = model.to(dtype)
    model = model.to(device)
    for name, mod in list(model.named_modules()):
        if isinstance(mod, torch.nn.Linear) and (name + ".qweight") in state:
            qw = state[name + ".qweight"]
            if tuple(qw.shape) == (mod.in_features, mod.out_features // 8) and bits == 4:
                # AutoAWQ GEMM layout (packed along out_features, interleaved nibble order): repack to the optimum format
                # like the reference does at load time (transformers/quantization/utils.py:702, utility.py:1432-1459)
                from .awq_repack import repack_awq_to_optimum_format

                g_ = group_size if group_size > 0 else mod.in_features
                state[name + ".qweight"], state[name + ".qzeros"], state[name + ".scales"] = repack_awq_to_optimum_format(
                    qw, state[name + ".qzeros"], state[name + ".scales"], bits, g_)
            g = group_size if group_size > 0 else mod.in_features
            new = B200WeightOnlyLinear(mod.in_features, mod.out_features, dtype="int", bits=bits, group_size=g,
                                       zp=True, bias=(name + ".bias") in state,
                                       g_idx=(name + ".g_idx") in state, device=device)
            set_module(model, name, new)
    missing, unexpected = model.load_state_dict(state, strict=False)
    missing = [k for k in missing if not k.endswith(("scale_bf16_to_fp8", ".bias", "rotary_emb.inv_freq"))]
    assert not missing, f"checkpoint lacks {missing[:5]}"
    if getattr(config, "tie_word_embeddings", False):
        model.tie_weights()
    model.config.quantization_config = qc
    model.eval()
    from .modules import fuse_sibling_linears

    fuse_sibling_linears(model)
    return model


def save(model, output_dir="./saved_results", format="default", **kwargs):
    format = getattr(format, "value", format)
    os.makedirs(output_dir, exist_ok=True)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if format == "huggingface":
        return _save_huggingface(model, output_dir, **kwargs)
    if format != "default":
        raise ValueError(f"unknown save format {format!r} (default | huggingface)")
    torch.save(model.state_dict(), os.path.join(os.path.abspath(os.path.expanduser(output_dir)), WEIGHT_NAME))
    qcfg = {}
    for (op_name, op_type), cfg in getattr(model, "qconfig", {}).items():
        qcfg[f"('{op_name}', '{op_type}')"] = {cfg.name: {k: getattr(cfg, k) for k in cfg.params_list}}
    with open(os.path.join(output_dir, QCONFIG_NAME), "w") as f:
        json.dump(qcfg, f, indent=4)


def load(model_name_or_path, original_model=None, format="default", device="cuda", **kwargs):
    """default: rebuild packed modules on `original_model` from qconfig.json, then load the state dict.
    huggingface: everything comes from the directory (config.json + shards)."""
    format = getattr(format, "value", format)
    if format == "huggingface":
        return _load_huggingface(model_name_or_path, device=device, **kwargs)
    if format != "default":
        raise ValueError(f"unknown load format {format!r} (default | huggingface)")
    assert original_model is not None, "original_model is required for format='default'"
    from ..dtypes import is_table_dtype
    from ..utils import fetch_module, set_module
    from .modules import B200WeightOnlyLinear, MulLinear

    with open(os.path.join(model_name_or_path, QCONFIG_NAME)) as f:
        qcfg = json.load(f)
    # tensors only, never pickled code (the reference's `_load_weight_file`, save_load.py: weights_only=True)
    state = torch.load(os.path.join(model_name_or_path, WEIGHT_NAME), map_location=device, weights_only=True)
    model = original_model.to(device)
    # Which modules are packed is decided by the checkpoint (`<name>.qweight` present), not by the qconfig dtype: the
    # GPTQ engine mirrors the reference's `get_layer_config` fallback (gptq.py:367-382), which quantises in-block
    # layers even when their own entry says fp32, so the tensors are the ground truth.
    first_cfg = None
    for body in qcfg.values():
        (_algo, c), = body.items()
        if c.get("dtype") != "fp32":
            first_cfg = c
            break
    for key, body in qcfg.items():
        op_name = key.split("'")[1]
        (algo, cfg), = body.items()
        packed_name = op_name + ".qweight"
        wrapped = op_name + ".linear.qweight"
        if packed_name not in state and wrapped not in state:
            continue
        if cfg.get("dtype") == "fp32":
            assert first_cfg is not None, f"{op_name}: packed tensors in the checkpoint but no quantised config entry"
            cfg = first_cfg
        m = fetch_module(model, op_name)
        lin = m
        in_f = lin.in_features if hasattr(lin, "in_features") else lin.weight.shape[0]
        out_f = lin.out_features if hasattr(lin, "out_features") else lin.weight.shape[1]
        prefix = op_name + (".linear" if wrapped in state else "")
        bits = cfg["bits"]
        if isinstance(cfg.get("dtype"), str) and cfg["dtype"] != "int" and cfg["dtype"].startswith("int"):
            bits = int(cfg["dtype"].lstrip("int"))
        group_size = cfg["group_size"] if cfg["group_size"] > 0 else in_f
        if is_table_dtype(cfg.get("dtype", "int")):
            # nf4 / fp4 modules live in the non-optimum layout: scales [N, G], no zero points (modules.py:214-222)
            g = state[prefix + ".scales"].shape[1]
            new = B200WeightOnlyLinear(in_f, out_f, dtype=cfg["dtype"], bits=bits, group_size=group_size, zp=False,
                                       bias=(prefix + ".bias") in state, device=device)
            groups = new.scales.shape[1]
        else:
            g = state[prefix + ".scales"].shape[0]
            new = B200WeightOnlyLinear(in_f, out_f, dtype="int", bits=bits, group_size=group_size,
                                       zp=True, bias=True, g_idx=(prefix + ".g_idx") in state, device=device)
            groups = new.scales.shape[0]
        assert groups == g, f"{op_name}: {groups} groups expected, checkpoint has {g}"
        if wrapped in state:
            new = MulLinear(new, torch.empty(in_f, device=device))
        set_module(model, op_name, new)
    result = model.load_state_dict(state, strict=False)
    stray = [k for k in result.unexpected_keys if k.rsplit(".", 1)[-1] in ("qweight", "qzeros", "scales", "g_idx")]
    if stray or result.missing_keys:
        raise RuntimeError(f"quantized checkpoint does not match the rebuilt model: unexpected packed tensors {stray[:8]}, "
                           f"missing {list(result.missing_keys)[:8]}")
    from .modules import fuse_sibling_linears

    fuse_sibling_linears(model)
    return model
