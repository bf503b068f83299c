"""save / load of quantised models in the reference's default on-disk format
(neural_compressor/torch/algorithms/weight_only/save_load.py:56-108 save, :111-285 load; file names
torch/utils/utility.py:56,60): `quantized_weight.pt` (torch.save(state_dict)) + `qconfig.json`.

The packed buffers keep the reference's names/dtypes/shapes, so a checkpoint written here is loadable by the
reference's WOQModelLoader and vice versa.
"""
import json
import os

import torch

WEIGHT_NAME = "quantized_weight.pt"
QCONFIG_NAME = "qconfig.json"


def save(model, output_dir="./saved_results", format="default", **kwargs):
    if format != "default":
        raise NotImplementedError("only format='default' is implemented (SURVEY §8 f1)")
    os.makedirs(output_dir, exist_ok=True)
    torch.save(model.state_dict(), os.path.join(os.path.abspath(os.path.expanduser(output_dir)), WEIGHT_NAME))
    qcfg = {}
    for (op_name, op_type), cfg in getattr(model, "qconfig", {}).items():
        qcfg[f"('{op_name}', '{op_type}')"] = {cfg.name: {k: getattr(cfg, k) for k in cfg.params_list}}
    with open(os.path.join(output_dir, QCONFIG_NAME), "w") as f:
        json.dump(qcfg, f, indent=4)


def load(model_name_or_path, original_model=None, format="default", device="cuda", **kwargs):
    """Rebuild packed modules on `original_model` from qconfig.json, then load the state dict."""
    if format != "default":
        raise NotImplementedError("only format='default' is implemented (SURVEY §8 f1)")
    assert original_model is not None, "original_model is required for format='default'"
    from ..utils import fetch_module, set_module
    from .modules import B200WeightOnlyLinear, MulLinear

    with open(os.path.join(model_name_or_path, QCONFIG_NAME)) as f:
        qcfg = json.load(f)
    state = torch.load(os.path.join(model_name_or_path, WEIGHT_NAME), map_location=device)
    model = original_model.to(device)
    for key, body in qcfg.items():
        op_name = key.split("'")[1]
        (algo, cfg), = body.items()
        if cfg.get("dtype") == "fp32":
            continue
        packed_name = op_name + ".qweight"
        wrapped = op_name + ".linear.qweight"
        if packed_name not in state and wrapped not in state:
            continue
        m = fetch_module(model, op_name)
        lin = m
        in_f = lin.in_features if hasattr(lin, "in_features") else lin.weight.shape[0]
        out_f = lin.out_features if hasattr(lin, "out_features") else lin.weight.shape[1]
        prefix = op_name + (".linear" if wrapped in state else "")
        bits = cfg["bits"]
        g = state[prefix + ".scales"].shape[0]
        group_size = cfg["group_size"] if cfg["group_size"] > 0 else in_f
        new = B200WeightOnlyLinear(in_f, out_f, dtype="int", bits=bits, group_size=group_size,
                                   zp=True, bias=True, g_idx=(prefix + ".g_idx") in state, device=device)
        assert new.scales.shape[0] == g
        if wrapped in state:
            new = MulLinear(new, torch.empty(in_f, device=device))
        set_module(model, op_name, new)
    model.load_state_dict(state, strict=False)
    return model
