"""Other architectures than Llama through the GPTQ and AWQ engines (host flow on the CPU, kernels = oracle twins) against
the UNMODIFIED reference (tests/golden/families.pt): OPT (biased linears, LayerNorm, positional block kwargs), GPT-J
(parallel attention / MLP behind one LayerNorm; the architecture of the reference's own tests) and GPT-2 (Conv1D weights
stored [in, out]; the reference's GPTQ export cannot handle it -- recorded -- ours packs it)."""
import os

import pytest
import torch

from tests.test_awq_absorb_cpu import host_ops as awq_host_ops  # noqa: F401  (fixture)

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(HERE, "golden", "families.pt"))


def build(name, golden):
    from oracle.gen_golden import family_models

    m = family_models()(name)
    m.load_state_dict(golden["init"][name])
    return m


@pytest.fixture()
def gptq_host_ops(monkeypatch):
    from tests.host_twins import install_gptq_twins

    install_gptq_twins(running_mean=True, setter=monkeypatch.setattr)
    monkeypatch.setenv("B200WOQ_CALIB_BATCH", "1")


def packed(state):
    return {k: v for k, v in state.items() if k.rsplit(".", 1)[-1] in ("qweight", "qzeros", "scales", "input_scale")}


@pytest.mark.parametrize("name", ["opt", "gptj"])
def test_gptq_families(gptq_host_ops, golden, name):
    import neural_compressor_b200.quantization as api
    from tests.test_gptq_hostflow_cpu import compare

    m = api.prepare(build(name, golden), api.GPTQConfig(bits=4, group_size=32, use_sym=False, block_size=128))
    for x in golden["ids"]:
        m(x)
    m = api.convert(m)
    want = golden["cases"][f"gptq_{name}"]["state"]
    assert set(packed(m.state_dict())) == set(packed(want))
    compare(m, want)


def test_gptq_gpt2_conv1d_where_the_reference_fails(gptq_host_ops, golden):
    import neural_compressor_b200.quantization as api
    from neural_compressor_b200.algorithms.modules import B200WeightOnlyLinear

    assert golden["cases"]["gptq_gpt2"]["reference_error"].startswith("RuntimeError")
    m = api.prepare(build("gpt2", golden), api.GPTQConfig(bits=4, group_size=32, use_sym=False, block_size=128))
    for x in golden["ids"]:
        m(x)
    m = api.convert(m)
    attn = m.transformer.h[0].attn.c_attn
    assert isinstance(attn, B200WeightOnlyLinear) and (attn.in_features, attn.out_features) == (64, 192)
    assert tuple(attn.qweight.shape) == (8, 192) and tuple(attn.scales.shape) == (2, 192)


@pytest.mark.parametrize("name", ["opt", "gptj"])
def test_awq_families(awq_host_ops, golden, name):  # noqa: F811
    import neural_compressor_b200.quantization as api

    ids = golden["ids"]

    def run_fn(model):
        for x in ids:
            model(x)

    m = api.quantize(build(name, golden), api.AWQConfig(bits=4, group_size=32, use_sym=False), run_fn=run_fn, example_inputs=ids[0])
    got, want = packed(m.state_dict()), packed(golden["cases"][f"awq_{name}"]["state"])
    assert set(got) == set(want), sorted(set(got) ^ set(want))[:6]
    from tests.test_options_gpu import fields

    inexact = []
    for k, ref in want.items():
        if ref.dtype == torch.int32:
            frac = (fields(got[k], 4) != fields(ref, 4)).float().mean().item()
            if frac > 0:
                inexact.append((k, frac))
        else:
            # the activation mean |x| is accumulated in another order than the reference's cat().mean(): the smoothing
            # scales agree to fp32 rounding, and a packed fp16 scale may land one fp16 ulp away
            tol = 2e-3 if ref.dtype == torch.float16 else 2e-6
            assert torch.allclose(got[k].float(), ref.float(), rtol=tol, atol=1e-8), k
    assert len(inexact) <= len(want) // 10 and all(f <= 5e-3 for _, f in inexact), inexact
