"""The configuration surface of the path (neural_compressor/common/base_config.py, torch/quantization/config.py): the
operations of the reference's own test/torch/test_config.py -- dict round trips, local overrides by name regex / module
type / white list, `+` on configs of the same and of different classes, the config mapping -- executed on this package and,
when the reference tree is present, on the live reference side by side (same parameter values, same mapping)."""
import pytest
import torch

import neural_compressor_b200.quantization as ours
from neural_compressor_b200.quantization.config import ComposableConfig, get_model_info


class Simple(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = torch.nn.Linear(8, 8)
        self.fc2 = torch.nn.Linear(8, 8)
        self.fc3 = torch.nn.Linear(8, 8)
        self.lm_head = torch.nn.Linear(8, 8)

    def forward(self, x):
        return self.lm_head(self.fc3(self.fc2(self.fc1(x))))


def reference_api():
    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        return None
    load_reference()
    import neural_compressor.torch.quantization as ref
    from neural_compressor.torch.utils import get_model_info as ref_info

    return ref, ref_info


def scenarios(api):
    """name -> config built with `api` (ours or the reference's)."""
    R, G, A = api.RTNConfig, api.GPTQConfig, api.AWQConfig
    out = {}
    out["plain"] = R(bits=4, dtype="nf4")
    c = R(bits=4, dtype="nf4")
    c.set_local("fc1", R(bits=6, dtype="int8"))
    out["local_name"] = c
    c = R(bits=4, dtype="nf4")
    c.set_local("fc1", R(bits=6, dtype="int8"))
    c.set_local("fc", R(bits=5, dtype="int8"))             # regex: matches fc1, fc2, fc3 (re.match)
    out["local_regex"] = c
    c = R(bits=4, dtype="nf4")
    c.set_local(torch.nn.Linear, R(bits=6, dtype="int8"))
    out["local_type"] = c
    out["white_list_add"] = R(bits=4, dtype="nf4") + R(bits=6, dtype="int8", white_list=["fc1"])
    out["from_dict_adv"] = R.from_dict({"global": {"dtype": "nf4", "bits": 4, "group_size": 32},
                                        "local": {"fc1": {"dtype": "int8", "bits": 4}}})
    a = R.from_dict({"dtype": "nf4", "bits": 4, "group_size": 32})
    b = R.from_dict({"global": {"bits": 8, "group_size": 32}, "local": {"fc1": {"dtype": "int8", "bits": 4}}})
    out["same_type_add"] = a + b
    out["gptq"] = G(bits=8, act_order=True)
    out["awq"] = A(bits=8, use_auto_scale=True, folding=False)
    out["quant_lm_head"] = R(bits=4, quant_lm_head=True)
    return out


def mapping_view(cfg, info):
    return {k: (v.name, v.dtype, v.bits, v.group_size) for k, v in cfg.to_config_mapping(model_info=info).items()}


def params_view(d, keys):
    if "global" in d or "local" in d:
        return {"global": {k: d["global"][k] for k in keys if k in d.get("global", {})} if "global" in d else None,
                "local": {op: {k: v[k] for k in keys if k in v} for op, v in d.get("local", {}).items()}}
    return {k: d[k] for k in keys if k in d}


def test_scenarios_match_the_reference_test_expectations():
    info = get_model_info(Simple(), white_module_list=[torch.nn.Linear])
    sc = scenarios(ours)
    m = mapping_view(sc["local_name"], info)
    assert m[("fc1", "Linear")][2] == 6 and m[("fc2", "Linear")][2] == 4                      # test_config.py:251-262
    m = mapping_view(sc["local_regex"], info)
    assert all(m[(n, "Linear")][2] == 5 for n in ("fc1", "fc2", "fc3"))                        # :263-271
    m = mapping_view(sc["local_type"], info)
    assert all(m[(n, "Linear")][2] == 6 for n in ("fc1", "fc2", "fc3"))                        # :273-286
    m = mapping_view(sc["white_list_add"], info)
    assert m[("fc1", "Linear")][2] == 6 and m[("fc2", "Linear")][2] == 4                      # :145-157
    d = sc["local_name"].to_dict()
    assert "global" in d and "local" in d                                                      # :178-184
    d = sc["same_type_add"].to_dict()
    assert d["local"]["fc1"]["dtype"] == "int8" and d["local"]["fc1"]["bits"] == 4 and d["global"]["bits"] != 8   # :186-215
    assert sc["gptq"].to_dict() == ours.GPTQConfig.from_dict({"bits": 8, "act_order": True}).to_dict()            # :288-294
    assert mapping_view(sc["plain"], info)[("lm_head", "Linear")][1] == "fp32"                 # lm_head stays fp32 ...
    assert mapping_view(sc["quant_lm_head"], info)[("lm_head", "Linear")][1] == "int"          # ... unless asked for


def test_composable_config():
    combined = ours.RTNConfig.from_dict({"dtype": "nf4", "bits": 4, "group_size": 32}) + ours.GPTQConfig(double_quant_bits=4)
    assert isinstance(combined, ComposableConfig)
    d = combined.to_dict()
    assert "rtn" in d and "gptq" in d and d["gptq"]["double_quant_bits"] == 4                  # test_config.py:217-231
    combined2 = combined + ours.GPTQConfig(double_quant_bits=4)
    combined3 = combined + combined2                                                           # :233-249
    assert len(combined3.config_list) >= 3
    again = ComposableConfig.from_dict(d)
    assert again.to_dict() == d
    assert combined.to_json_string().strip().startswith("{")
    # the reference's composable mapping applies LOCAL entries only (base_config.py:794-816): two global configs map nothing
    info = combined.get_model_info(Simple())
    assert set(info) == {"rtn", "gptq"} and combined.to_config_mapping(model_info=info) == {}
    local = ours.RTNConfig(bits=8, white_list=["fc1"]) + ours.GPTQConfig(bits=3, white_list=["fc2"])
    view = {k: (v.name, v.bits) for k, v in local.to_config_mapping(model_info=local.get_model_info(Simple())).items()}
    assert view == {("fc1", "Linear"): ("rtn", 8), ("fc2", "Linear"): ("gptq", 3)}


def test_dict_configs_and_json_round_trip(tmp_path):
    from neural_compressor_b200.quantization.quantize import _as_config

    c = _as_config({"rtn": {"dtype": "nf4", "bits": 4, "group_size": 32}})
    assert isinstance(c, ours.RTNConfig) and c.dtype == "nf4"
    c = _as_config({"rtn": {"global": {"dtype": "nf4", "bits": 4, "group_size": 32}, "local": {"fc1": {"dtype": "int8", "bits": 4}}}})
    assert c.local_config["fc1"].dtype == "int8"
    both = _as_config({"rtn": {"bits": 8}, "gptq": {"bits": 3}})
    assert isinstance(both, ComposableConfig) and [x.name for x in both.config_list] == ["rtn", "gptq"]
    f = tmp_path / "cfg.json"
    ours.GPTQConfig(bits=3, act_order=True).to_json_file(f)
    assert ours.GPTQConfig.from_json_file(f).to_dict() == ours.GPTQConfig(bits=3, act_order=True).to_dict()


def test_side_by_side_with_the_live_reference():
    live = reference_api()
    if live is None:
        pytest.skip("reference tree not present")
    ref, ref_info = live
    model = Simple()
    info_o = get_model_info(model, white_module_list=[torch.nn.Linear])
    info_r = ref_info(model, white_module_list=[torch.nn.Linear])
    assert info_o == info_r
    so, sr = scenarios(ours), scenarios(ref)
    for name in so:
        assert mapping_view(so[name], info_o) == mapping_view(sr[name], info_r), name
        keys = type(so[name]).params_list
        assert params_view(so[name].to_dict(), keys) == params_view(sr[name].to_dict(), keys), name
    co = ours.RTNConfig(bits=8, white_list=["fc1"]) + ours.GPTQConfig(bits=3, white_list=["fc2"])
    cr = ref.RTNConfig(bits=8, white_list=["fc1"]) + ref.GPTQConfig(bits=3, white_list=["fc2"])
    assert {k: (v.name, v.bits) for k, v in co.to_config_mapping(model_info=co.get_model_info(model)).items()} == \
        {k: (v.name, v.bits) for k, v in cr.to_config_mapping(model_info=cr.get_model_info(model)).items()}


def test_double_quant_presets_match_the_reference():
    for preset in ("BNB_NF4", "GGML_TYPE_Q4_K"):
        c = ours.get_default_double_quant_config(preset)
        assert c.use_double_quant and isinstance(c, ours.RTNConfig)
    assert ours.get_default_double_quant_config().dtype == "nf4"
    with pytest.raises(AssertionError):
        ours.get_default_double_quant_config("Q8")
    live = reference_api()
    if live is not None:
        ref, _ = live
        for preset in ("BNB_NF4", "GGML_TYPE_Q4_K"):
            a, b = ours.get_default_double_quant_config(preset), ref.get_default_double_quant_config(preset)
            assert all(getattr(a, k) == getattr(b, k) for k in ours.RTNConfig.params_list), preset
