"""AWQ absorb-layer discovery and folding, host side.

* algorithms/absorb.py (one observed eager forward) against the reference's torch.jit GraphTrace (utility.py:657-984):
  the structures recorded in tests/golden/awq_toy.pt from the live reference, and -- when the reference tree is present --
  a live comparison on four toy variants (residual consumer, cast + mul consumers, a view in between).
* The AWQ host flow with discovered tuples (awq.py:40-95 grouping, multi-module search through the block, folding the
  scale into the absorbing LayerNorm / Linear, self-absorbing leftovers as MulLinear) with the device kernels replaced by
  the oracle's CPU twins, against the packed state dict the UNMODIFIED reference produced for the same toy model.
"""
import os

import pytest
import torch

from tests.toy_models import Toy

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(HERE, "golden", "awq_toy.pt"))


def toy(golden):
    m = Toy(d=64, n=2, variant=0, vocab=64).eval()
    m.load_state_dict(golden["init_state"])
    return m


def norm(d):
    return {k: sorted(v) for k, v in d.items()}


def test_discovery_matches_recorded_graphtrace(golden):
    from neural_compressor_b200.algorithms.absorb import get_absorb_layers

    absorb, no_absorb = get_absorb_layers(toy(golden), golden["ids"][0])
    assert norm(absorb) == norm(golden["absorb_to_layer"])
    assert sorted(no_absorb) == sorted(golden["no_absorb_layers"])
    assert absorb["layers.0.ln1"] == ["layers.0.q", "layers.0.k", "layers.0.v"]   # execution order inside a tuple
    assert absorb["layers.1.fc1"] == ["layers.1.fc2"]                              # through the ReLU


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_discovery_matches_live_graphtrace(variant):
    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not present")
    load_reference()
    from neural_compressor.torch.algorithms.weight_only.utility import get_absorb_layers as ref_get

    from neural_compressor_b200.algorithms.absorb import get_absorb_layers

    torch.manual_seed(variant)
    m = Toy(variant=variant).eval()
    ids = torch.ones(2, 8, dtype=torch.long)
    ra, rn = ref_get(m, ids, supported_layers=["Linear"])
    oa, on = get_absorb_layers(m, ids)
    assert norm(oa) == norm(ra) and sorted(on) == sorted(rn)


def test_discovery_is_conservative_without_inputs_and_on_hf_models():
    from transformers import LlamaConfig, LlamaForCausalLM

    from neural_compressor_b200.algorithms.absorb import get_absorb_layers

    m = LlamaForCausalLM(LlamaConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2,
                                     num_key_value_heads=2, vocab_size=128, max_position_embeddings=64)).eval()
    absorb, no_absorb = get_absorb_layers(m, None)
    assert absorb == {} and len(no_absorb) == 8
    absorb, no_absorb = get_absorb_layers(m, torch.ones(1, 8, dtype=torch.long))
    p = "model.layers.0."
    assert absorb == {p + "input_layernorm": [p + "self_attn.q_proj", p + "self_attn.k_proj", p + "self_attn.v_proj"],
                      p + "post_attention_layernorm": [p + "mlp.gate_proj", p + "mlp.up_proj"]}
    assert set(no_absorb) == {p + "self_attn.o_proj", p + "mlp.down_proj", "lm_head"}


@pytest.fixture()
def host_ops(monkeypatch):
    """ops.* used by algorithms/awq.py and rtn.py, on the CPU through the oracle."""
    from neural_compressor_b200 import ops
    from neural_compressor_b200.algorithms import awq, rtn
    from oracle import woq_oracle as O

    def rtn_fake_quant(W, bits=4, group_size=-1, sym=False, full_range=False, quantile=1.0, col_scale=None, out=None):
        scheme = "sym" if sym else "asym"
        if col_scale is not None:
            r = O.rtn_fake_quant(W.float() * col_scale.view(1, -1), bits, group_size, scheme, quantile, full_range) / col_scale.view(1, -1)
        else:
            r = O.rtn_fake_quant(W, bits, group_size, scheme, quantile, full_range)
        return r if out is None else out.copy_(r)

    def abs_colsum_accumulate(X, acc):
        X2 = X.reshape(-1, X.shape[-1])
        acc += X2.abs().float().sum(0)
        return X2.shape[0]

    def mse_accumulate(a, b, acc):
        acc += (a - b).float().pow(2).mean().double()

    def rtn_quant_pack(W, bits=4, group_size=-1, sym=False, full_range=False, quantile=1.0, return_codes=False):
        q, s, z = O.rtn_quantize(W, bits, group_size, "sym" if sym else "asym", quantile, full_range)
        qweight, qzeros, scales16 = O.pack_optimum(q, s, z, bits, group_size)
        return dict(qweight=qweight, qzeros=qzeros, scales=scales16, scale_f32=s.float(), zp_f32=None if z is None else z.float())

    for name, fn in dict(rtn_fake_quant=rtn_fake_quant, abs_colsum_accumulate=abs_colsum_accumulate, mse_accumulate=mse_accumulate,
                         rtn_quant_pack=rtn_quant_pack, awq_weight_scale=lambda W, g: O.awq_weight_scale(W, g)).items():
        monkeypatch.setattr(ops, name, fn)
    cpu = lambda: torch.device("cpu")  # noqa: E731
    monkeypatch.setattr(awq, "current_device", cpu)
    monkeypatch.setattr(rtn, "current_device", cpu)
    monkeypatch.setenv("B200WOQ_AWQ_CHUNK_TOKENS", "0")   # sample-by-sample like the reference: same GEMM shapes on the host
    return ops


@pytest.mark.parametrize("tag", ["folding_false", "folding_true", "folding_true_sym"])
def test_awq_host_flow_with_discovered_absorption(host_ops, golden, tag, monkeypatch):
    import neural_compressor_b200.quantization as api

    kw = dict(folding_false=dict(folding=False, use_sym=False), folding_true=dict(folding=True, use_sym=False),
              folding_true_sym=dict(folding=True, use_sym=True, group_size=64))[tag]
    monkeypatch.setenv("B200WOQ_AWQ_ABSORB", "eager")
    ids = golden["ids"]

    def run_fn(model):
        for x in ids:
            model(x)

    kw.setdefault("group_size", 32)
    m = api.quantize(toy(golden), api.AWQConfig(bits=4, **kw), run_fn=run_fn, example_inputs=ids[0])
    got = {k: v for k, v in m.state_dict().items() if "bf16_to_fp8" not in k}
    want = golden["cases"][tag]["state"]
    assert set(got) == set(want), (sorted(set(got) ^ set(want)))
    for k, ref in want.items():
        if ref.dtype in (torch.int32,):
            assert torch.equal(got[k], ref), k
        else:
            assert got[k].dtype == ref.dtype and torch.allclose(got[k].float(), ref.float(), rtol=1e-6, atol=1e-7), k
    if tag == "folding_false":
        from neural_compressor_b200.algorithms.modules import MulLinear

        assert isinstance(m.layers[0].o, MulLinear) and not isinstance(m.layers[0].q, MulLinear)
    else:   # everything is folded: no run-time multiply anywhere; `o` had nothing to fold into and kept a plain RTN
        assert not any(type(x).__name__ == "MulLinear" for x in m.modules())


@pytest.mark.parametrize("arch,admitted", [("mistral", True), ("qwen2", True), ("gemma", False)])
def test_extended_absorbers_are_admitted_by_execution(arch, admitted):
    """Beyond the reference's fixed type list: a `*Norm` type is admitted iff folding a random scale into it leaves the
    model output unchanged -- MistralRMSNorm / Qwen2RMSNorm pass, GemmaRMSNorm (1 + weight) is rejected; the probe restores
    every weight."""
    from neural_compressor_b200.algorithms.absorb import get_absorb_layers
    from tests.test_arch_sweep_cpu import build

    m = build(arch).float().eval()
    before = {k: v.clone() for k, v in m.state_dict().items()}
    ids = torch.randint(0, 256, (2, 24), generator=torch.Generator().manual_seed(0))
    strict, _ = get_absorb_layers(m, ids, extended=False)
    assert strict == {}                                      # none of these norm types is in the reference's list
    found, rest = get_absorb_layers(m, ids, extended=True)
    assert all(torch.equal(v, m.state_dict()[k]) for k, v in before.items())
    if not admitted:
        assert found == {} and len(rest) == 15
        return
    p = "model.layers.0."
    assert found[p + "input_layernorm"] == [p + "self_attn.q_proj", p + "self_attn.k_proj", p + "self_attn.v_proj"]
    assert found[p + "post_attention_layernorm"] == [p + "mlp.gate_proj", p + "mlp.up_proj"]


def test_awq_folding_on_mistral_with_extended_absorbers(host_ops, monkeypatch):
    """folding=True end to end on an architecture the reference cannot fold at all: scales go into the RMSNorms, no
    MulLinear is inserted, and the 4-bit model stays close to the fp one."""
    import neural_compressor_b200.quantization as api
    from tests.test_arch_sweep_cpu import build

    monkeypatch.setenv("B200WOQ_ABSORB_EXTENDED", "1")
    g = torch.Generator().manual_seed(1234)
    ids = [torch.randint(0, 256, (1, 32), generator=g) for _ in range(8)]
    m = build("mistral").float().eval()
    with torch.no_grad():
        fp = m(ids[0]).logits.clone()
    norm0 = m.model.layers[0].input_layernorm.weight.detach().clone()

    def run_fn(model):
        for x in ids:
            model(x)

    # (sym: the reference's recover() wraps int8(q - zp) for 8-bit ASYMMETRIC codes, modules.py:435 -- reproduced by K4/K6 --
    # so only the symmetric 8-bit format can serve as a "nearly lossless" probe)
    q = api.quantize(m, api.AWQConfig(bits=8, group_size=32, use_sym=True, folding=True), run_fn=run_fn, example_inputs=ids[0])
    assert not any(type(x).__name__ == "MulLinear" for x in q.modules())
    assert not torch.equal(q.model.layers[0].input_layernorm.weight, norm0)
    packed = {n: x for n, x in q.named_modules() if type(x).__name__ == "B200WeightOnlyLinear"}
    assert len(packed) == 14       # folded q/k/v + gate/up, and o_proj / down_proj (no absorber: plain RTN) in both blocks
    # evaluate the packed model on the host: recover every weight with the oracle and run dense linears
    from neural_compressor_b200.utils import set_module
    from oracle import woq_oracle as O

    for n, x in packed.items():
        lin = torch.nn.Linear(x.in_features, x.out_features, bias=False)
        lin.weight.data = O.recover_fp16(x.qweight, x.qzeros, x.scales, x.bits, x.group_size, x.in_features, x.out_features).float()
        set_module(q, n, lin)
    with torch.no_grad():
        out = q(ids[0]).logits
    assert float((out - fp).norm() / fp.norm()) < 6e-2      # 8-bit weights (+ clip search, fp16 scales): still the same function
