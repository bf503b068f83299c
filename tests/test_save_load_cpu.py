"""On-disk formats -- host logic only, no kernels: the HuggingFace layout written by `save` and read
back by `load`, and the AutoGPTQ-style `quantization_config` (save_load.py:1094-1156 of the reference)."""
import json
import os

import pytest
import torch


def tiny():
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=128, max_position_embeddings=64, tie_word_embeddings=False)
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg).eval()


def test_change_config_to_hf_format():
    from neural_compressor_b200.algorithms.save_load import change_config_to_hf_format
    from neural_compressor_b200.quantization import GPTQConfig, RTNConfig

    g = GPTQConfig(bits=4, group_size=128, use_sym=True, act_order=True, percdamp=0.02)
    fp = GPTQConfig(dtype="fp32")
    got = change_config_to_hf_format({("model.layers.0.q_proj", "Linear"): g, ("model.layers.0.k_proj", "Linear"): g,
                                      ("lm_head", "Linear"): fp})
    assert got == {"bits": 4, "group_size": 128, "damp_percent": 0.02, "desc_act": True, "sym": True,
                   "true_sequential": False, "model_name_or_path": None, "model_file_base_name": "model",
                   "quant_method": "gptq"}
    r = RTNConfig(bits=8, group_size=32, use_sym=False)
    got = change_config_to_hf_format({("a", "Linear"): r})
    assert (got["bits"], got["group_size"], got["sym"], got["damp_percent"], got["desc_act"]) == (8, 32, False, 0, False)
    with pytest.raises(ValueError):  # a quantised lm_head cannot be written in this format
        change_config_to_hf_format({("lm_head", "Linear"): r})
    with pytest.raises(AssertionError):
        change_config_to_hf_format({("a", "Linear"): r, ("b", "Linear"): RTNConfig(bits=4, group_size=32, use_sym=False)})


def test_huggingface_format_roundtrip_cpu(tmp_path):
    from neural_compressor_b200.algorithms import save_load
    from neural_compressor_b200.algorithms.modules import B200WeightOnlyLinear
    from neural_compressor_b200.quantization import RTNConfig
    from neural_compressor_b200.utils import set_module

    m = tiny()
    cfg = RTNConfig(bits=4, group_size=32, use_sym=True)
    qconfig = {}
    gen = torch.Generator().manual_seed(1)
    for name, mod in list(m.named_modules()):
        if isinstance(mod, torch.nn.Linear) and name != "lm_head":
            new = B200WeightOnlyLinear(mod.in_features, mod.out_features, bits=4, group_size=32, zp=True, device="cpu")
            new.qweight = torch.randint(-2**31, 2**31 - 1, new.qweight.shape, generator=gen, dtype=torch.int64).to(torch.int32)
            new.qzeros = torch.randint(-2**31, 2**31 - 1, new.qzeros.shape, generator=gen, dtype=torch.int64).to(torch.int32)
            new.scales = torch.rand(new.scales.shape, generator=gen).half()
            set_module(m, name, new)
            qconfig[(name, "Linear")] = cfg
    qconfig[("lm_head", "Linear")] = RTNConfig(dtype="fp32")
    m.qconfig = qconfig
    save_load.save(m, str(tmp_path), format="huggingface")
    files = set(os.listdir(tmp_path))
    assert {"config.json", "quantize_config.json"} <= files and any(f.endswith(".safetensors") for f in files)
    qc = json.load(open(tmp_path / "quantize_config.json"))
    assert qc["quant_method"] == "gptq" and qc["bits"] == 4 and qc["group_size"] == 32 and qc["sym"] is True
    assert json.load(open(tmp_path / "config.json"))["quantization_config"]["bits"] == 4
    m2 = save_load.load(str(tmp_path), format="huggingface", device="cpu")
    s1, s2 = m.state_dict(), m2.state_dict()
    assert set(s1) == set(s2)
    for k in s1:
        assert s1[k].dtype == s2[k].dtype and torch.equal(s1[k], s2[k]), k
    assert isinstance(m2.model.layers[1].mlp.down_proj, B200WeightOnlyLinear)
    assert isinstance(m2.lm_head, torch.nn.Linear)


def test_hf_quant_config_matches_reference_fixture():
    """tests/golden/hf_quant_config.json was written by the live reference (oracle/gen_golden.py hf_config)."""
    import neural_compressor_b200.quantization as q
    from neural_compressor_b200.algorithms.save_load import change_config_to_hf_format

    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hf_quant_config.json")))
    for tag, case in cases.items():
        args = dict(case["args"])
        cfg = getattr(q, args.pop("kind"))(**args)
        mapping = {("model.layers.0.q_proj", "Linear"): cfg}
        if tag == "gptq":
            mapping[("lm_head", "Linear")] = q.GPTQConfig(dtype="fp32")
        assert change_config_to_hf_format(mapping) == case["out"], tag


def test_awq_repack_matches_reference_fixture():
    """f1: AutoAWQ GEMM layout -> optimum format (utility.py:1432-1459); the fixture was written by the live reference's
    `repack_awq_to_optimum_format` (oracle/gen_golden.py awq_repack).  Pure integer tensor ops: runs on the host."""
    import os

    import torch

    from neural_compressor_b200.algorithms.awq_repack import repack_awq_to_optimum_format

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "awq_repack.pt"))
    qw, qz, sc = repack_awq_to_optimum_format(g["awq_qweight"], g["awq_qzeros"], g["awq_scales"], 4, g["group_size"])
    assert torch.equal(qw, g["qweight"]) and torch.equal(qz, g["qzeros"]) and torch.equal(sc, g["scales"])
    assert qw.dtype == torch.int32 and tuple(qw.shape) == (g["awq_qweight"].shape[0] // 8, g["awq_qweight"].shape[1] * 8)


def _optimum_to_autoawq(qweight, qzeros):
    """Test-side inverse of the repack, written independently: optimum tensors -> AutoAWQ GEMM layout (packed along the
    output channels, nibble i of a word = column 8*w + [0, 2, 4, 6, 1, 3, 5, 7][i], zero points stored as they are)."""
    import torch

    def unpack(words, axis):
        w = words.to(torch.int64) & 0xFFFFFFFF
        f = torch.stack([(w >> (4 * e)) & 0xF for e in range(8)], dim=axis + 1)
        shape = list(words.shape)
        shape[axis] *= 8
        return f.reshape(shape)

    codes = unpack(qweight, 0)                       # [K, N]
    zeros = (unpack(qzeros, 1) + 1) & 0xF            # [G, N]   (optimum stores zp - 1)
    order = [0, 2, 4, 6, 1, 3, 5, 7]

    def pack_awq(v):                                 # [R, N] -> [R, N/8]
        v = v.reshape(v.shape[0], -1, 8)[:, :, order]
        w = sum(v[:, :, i] << (4 * i) for i in range(8))
        return torch.where(w >= 2**31, w - 2**32, w).to(torch.int32)

    return pack_awq(codes), pack_awq(zeros)


def test_autoawq_checkpoint_loads_through_the_huggingface_loader(tmp_path):
    """f1: a checkpoint in AutoAWQ's GEMM layout is repacked to the optimum format at load time, as the reference does
    (transformers/quantization/utils.py:702 -> utility.py:1432-1459).  The AutoAWQ tensors are derived here from the
    packed RTN model of the live reference (tests/golden/e2e_tiny_llama.pt) with an independently written inverse, so the
    loaded modules must hold exactly the reference's optimum tensors.  Loading touches no kernel: it runs on the host."""
    import json
    import os

    import torch
    from safetensors.torch import save_file

    from neural_compressor_b200.algorithms.modules import B200WeightOnlyLinear
    from neural_compressor_b200.quantization import load
    from tests.test_api_gpu import tiny_llama

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "e2e_tiny_llama.pt"))
    packed = g["rtn_asym"]["state"]
    m = tiny_llama(g["init_state"])
    m.save_pretrained(tmp_path, safe_serialization=True)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    for k in [k for k in packed if k.endswith(".qweight")]:
        base = k[: -len(".qweight")]
        state.pop(base + ".weight")
        aw, az = _optimum_to_autoawq(packed[base + ".qweight"], packed[base + ".qzeros"])
        assert tuple(aw.shape) == (packed[base + ".qweight"].shape[0] * 8, packed[base + ".qweight"].shape[1] // 8)
        state[base + ".qweight"], state[base + ".qzeros"], state[base + ".scales"] = aw, az, packed[base + ".scales"].clone()
    for fn in os.listdir(tmp_path):
        if fn.endswith(".safetensors"):
            os.remove(os.path.join(tmp_path, fn))
    save_file({k: v.contiguous() for k, v in state.items()}, os.path.join(tmp_path, "model.safetensors"))
    cfg_file = os.path.join(tmp_path, "config.json")
    cfg = json.load(open(cfg_file))
    cfg["quantization_config"] = {"quant_method": "awq", "bits": 4, "group_size": 32, "zero_point": True, "version": "gemm"}
    json.dump(cfg, open(cfg_file, "w"))
    loaded = load(str(tmp_path), format="huggingface", device="cpu")
    sd = loaded.state_dict()
    n = 0
    for k, ref in packed.items():
        assert torch.equal(sd[k], ref), k
        n += 1
    assert n == 42 and isinstance(loaded.model.layers[0].self_attn.q_proj, B200WeightOnlyLinear)


def test_default_format_load_reads_tensors_only(tmp_path, monkeypatch):
    """The reference's loader passes weights_only=True to torch.load (test_load.py:111-128): a checkpoint is data."""
    import json

    import torch

    from neural_compressor_b200.algorithms import save_load

    calls = []
    real = torch.load

    def spy(path, **kwargs):
        calls.append(kwargs)
        return real(path, **kwargs)

    torch.save({"weight": torch.ones(2, 2), "bias": torch.zeros(2)}, tmp_path / save_load.WEIGHT_NAME)
    json.dump({}, open(tmp_path / save_load.QCONFIG_NAME, "w"))
    monkeypatch.setattr(torch, "load", spy)
    m = save_load.load(str(tmp_path), original_model=torch.nn.Linear(2, 2), device="cpu")
    assert calls and all(c.get("weights_only") is True for c in calls)
    assert torch.equal(m.weight, torch.ones(2, 2))
