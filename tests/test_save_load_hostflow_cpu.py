"""save -> load round trips of models produced by the engines' host flows (kernels = oracle twins): GPTQ with act_order
(g_idx) in both on-disk formats, AWQ with its MulLinear wrappers in the default format.  Loading runs on the host."""
import pytest
import torch

from tests.test_awq_absorb_cpu import host_ops as awq_host_ops  # noqa: F401  (fixture)


@pytest.fixture()
def gptq_host_ops(monkeypatch):
    from tests.host_twins import install_gptq_twins
    install_gptq_twins(running_mean=True, setter=monkeypatch.setattr)
    monkeypatch.setenv("B200WOQ_CALIB_BATCH", "1")

def test_gptq_act_order_roundtrip(gptq_host_ops, golden_e2e, tmp_path):
    import neural_compressor_b200.quantization as api
    from tests.test_api_gpu import tiny_llama
    m = api.prepare(tiny_llama(golden_e2e["init_state"]), api.GPTQConfig(bits=4, group_size=32, act_order=True))
    for x in golden_e2e["ids"]:
        m(x)
    m = api.convert(m)
    for fmt in ("default", "huggingface"):
        d = tmp_path / fmt
        m.save(str(d), format=fmt)
        if fmt == "default":
            loaded = api.load(str(d), original_model=tiny_llama(golden_e2e["init_state"]), device="cpu")
        else:
            loaded = api.load(str(d), format="huggingface", device="cpu")
        a, b = m.state_dict(), loaded.state_dict()
        keys = [k for k in a if k.rsplit(".", 1)[-1] in ("qweight", "qzeros", "scales", "g_idx")]
        assert len(keys) == 56
        for k in keys:
            assert torch.equal(a[k], b[k]), (fmt, k)

def test_awq_roundtrip(awq_host_ops, golden_e2e, tmp_path):  # noqa: F811
    import neural_compressor_b200.quantization as api
    from tests.test_api_gpu import tiny_llama
    ids = golden_e2e["ids"]
    def run_fn(model):
        for x in ids:
            model(x)
    m = api.quantize(tiny_llama(golden_e2e["init_state"]), api.AWQConfig(bits=4, group_size=32, use_sym=False), run_fn=run_fn, example_inputs=ids[0])
    m.save(str(tmp_path))
    loaded = api.load(str(tmp_path), original_model=tiny_llama(golden_e2e["init_state"]), device="cpu")
    a, b = m.state_dict(), loaded.state_dict()
    keys = [k for k in a if k.rsplit(".", 1)[-1] in ("qweight", "qzeros", "scales", "input_scale")]
    assert len(keys) == 56
    for k in keys:
        assert torch.equal(a[k], b[k]), k
    assert type(loaded.model.layers[0].self_attn.q_proj).__name__ == "MulLinear"


def test_the_live_reference_loads_our_default_format_checkpoint(golden_e2e, tmp_path, monkeypatch):
    """Interop the other way round: a checkpoint written by `model.save()` here (quantized_weight.pt + qconfig.json) is
    read by the UNMODIFIED reference's `load(..., original_model=...)`, which rebuilds its own INCWeightOnlyLinear modules
    from it; the logits equal those of the model the reference quantised itself (tests/golden/e2e_tiny_llama.pt)."""
    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not present")
    load_reference()
    import neural_compressor.torch.quantization as ref

    import neural_compressor_b200.quantization as api
    from neural_compressor_b200 import ops
    from neural_compressor_b200.algorithms import rtn
    from oracle import woq_oracle as O
    from tests.test_api_gpu import tiny_llama

    def rtn_quant_pack(W, bits=4, group_size=-1, sym=False, full_range=False, quantile=1.0, return_codes=False):
        q, s, z = O.rtn_quantize(W, bits, group_size, "sym" if sym else "asym", quantile, full_range)
        qweight, qzeros, scales16 = O.pack_optimum(q, s, z, bits, group_size)
        return dict(qweight=qweight, qzeros=qzeros, scales=scales16, scale_f32=s.float(), zp_f32=None if z is None else z.float())

    monkeypatch.setattr(ops, "rtn_quant_pack", rtn_quant_pack)
    monkeypatch.setattr(rtn, "current_device", lambda: torch.device("cpu"))
    m = api.convert(api.prepare(tiny_llama(golden_e2e["init_state"]),
                                api.RTNConfig(bits=4, group_size=32, use_sym=False, use_layer_wise=False)))
    m.save(str(tmp_path))
    loaded = ref.load(str(tmp_path), original_model=tiny_llama(golden_e2e["init_state"]))
    assert type(loaded.model.layers[0].self_attn.q_proj).__name__ == "INCWeightOnlyLinear"
    with torch.no_grad():
        out = loaded(golden_e2e["probe"]).logits
    assert torch.equal(out, golden_e2e["rtn_asym"]["logits"])


def test_we_load_a_checkpoint_saved_by_the_live_reference(golden_e2e, tmp_path):
    """... and the other direction: `quantized_weight.pt` + `qconfig.json` written by the reference's own `model.save()`
    are read by `load()` here into B200WeightOnlyLinear modules holding the identical tensors."""
    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not present")
    load_reference()
    import neural_compressor.torch.quantization as ref

    import neural_compressor_b200.quantization as api
    from tests.test_api_gpu import tiny_llama

    m = ref.convert(ref.prepare(tiny_llama(golden_e2e["init_state"]),
                                ref.RTNConfig(bits=4, group_size=32, use_sym=False, use_layer_wise=False)))
    m.save(str(tmp_path))
    loaded = api.load(str(tmp_path), original_model=tiny_llama(golden_e2e["init_state"]), device="cpu")
    assert type(loaded.model.layers[0].self_attn.q_proj).__name__ == "B200WeightOnlyLinear"
    a, b = m.state_dict(), loaded.state_dict()
    keys = [k for k in a if k.rsplit(".", 1)[-1] in ("qweight", "qzeros", "scales")]
    assert len(keys) == 42 and all(torch.equal(a[k], b[k]) for k in keys)
