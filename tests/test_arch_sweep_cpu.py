"""Architecture sweep, side by side with the LIVE reference (skipped where /root/reference is absent): tiny random-init
Mistral, Qwen2, Falcon, Bloom, GPT-NeoX, Phi, Gemma and MPT models go through the reference's GPTQ / AWQ on the CPU and
through this package's engines with the kernels swapped for their oracle twins (tests/host_twins.py).  The packed tensors
must coincide: every architecture brings its own block signature (alibi, position embeddings, parallel attention, fused
qkv, biased linears) to the calibration capture and the per-input Hessian sharing."""
import copy

import pytest
import torch

from tests.test_awq_absorb_cpu import host_ops as awq_host_ops  # noqa: F401  (fixture)
from tests.test_options_gpu import fields

ARCHS = ["mistral", "qwen2", "falcon", "bloom", "gpt_neox", "phi", "gemma", "mpt"]


def build(name):
    import transformers as T

    torch.manual_seed(0)
    v = dict(vocab_size=256)
    llama_like = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=128, **v)
    if name == "mistral":
        m = T.MistralForCausalLM(T.MistralConfig(**llama_like))
    elif name == "qwen2":
        m = T.Qwen2ForCausalLM(T.Qwen2Config(**llama_like))
    elif name == "gemma":
        m = T.GemmaForCausalLM(T.GemmaConfig(head_dim=16, **llama_like))
    elif name == "falcon":
        m = T.FalconForCausalLM(T.FalconConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, **v))
    elif name == "bloom":
        m = T.BloomForCausalLM(T.BloomConfig(hidden_size=64, n_layer=2, n_head=4, **v))
    elif name == "gpt_neox":
        m = T.GPTNeoXForCausalLM(T.GPTNeoXConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                                 num_attention_heads=4, max_position_embeddings=128, **v))
    elif name == "phi":
        m = T.PhiForCausalLM(T.PhiConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                         max_position_embeddings=128, **v))
    else:
        m = T.MptForCausalLM(T.MptConfig(d_model=64, n_heads=4, n_layers=2, max_seq_len=128, **v))
    m.eval()
    m.config.use_cache = False
    return m


@pytest.fixture(scope="module")
def ref_api():
    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not present")
    load_reference()
    import neural_compressor.torch.quantization as ref

    return ref


@pytest.fixture(scope="module")
def ids():
    g = torch.Generator().manual_seed(1234)
    return [torch.randint(0, 256, (1, 32), generator=g) for _ in range(8)]


def packed(model):
    return {k: v for k, v in model.state_dict().items() if k.rsplit(".", 1)[-1] in ("qweight", "qzeros", "scales", "input_scale")}


def assert_same(got, want):
    assert set(got) == set(want), sorted(set(got) ^ set(want))[:6]
    for k, ref in want.items():
        if ref.dtype == torch.int32:
            assert (fields(got[k], 4) != fields(ref, 4)).float().mean().item() <= 2e-3, k
        else:
            tol = 2e-3 if ref.dtype == torch.float16 else 2e-6
            assert torch.allclose(got[k].float(), ref.float(), rtol=tol, atol=1e-8), k


@pytest.mark.parametrize("arch", ARCHS)
def test_gptq_matches_live_reference(ref_api, ids, arch, monkeypatch):
    import neural_compressor_b200.quantization as ours
    from tests.host_twins import install_gptq_twins

    install_gptq_twins(running_mean=True, setter=monkeypatch.setattr)
    monkeypatch.setenv("B200WOQ_CALIB_BATCH", "1")
    base = build(arch)
    out = {}
    for tag, api in (("ref", ref_api), ("ours", ours)):
        kw = dict(bits=4, group_size=32, use_sym=False, block_size=128)
        if tag == "ref":
            kw["model_path"] = "/tmp"
        m = api.prepare(copy.deepcopy(base), api.GPTQConfig(**kw))
        for x in ids:
            m(x)
        out[tag] = packed(api.convert(m))
    assert len(out["ref"]) >= 24
    assert_same(out["ours"], out["ref"])


@pytest.mark.parametrize("arch", ["falcon", "bloom", "gpt_neox", "phi", "qwen2"])
def test_awq_matches_live_reference(ref_api, ids, arch, awq_host_ops):  # noqa: F811
    import neural_compressor_b200.quantization as ours

    def run_fn(model):
        for x in ids:
            model(x)

    base = build(arch)
    out = {}
    for tag, api in (("ref", ref_api), ("ours", ours)):
        m = api.quantize(copy.deepcopy(base), api.AWQConfig(bits=4, group_size=32, use_sym=False), run_fn=run_fn,
                         example_inputs=ids[0])
        out[tag] = packed(m)
    assert len(out["ref"]) >= 24
    assert_same(out["ours"], out["ref"])


@pytest.mark.parametrize("arch", ["falcon", "bloom", "phi", "mpt", "qwen2"])
def test_smoothquant_transform_matches_live_reference(ref_api, ids, arch, monkeypatch):
    """`TorchSmoothQuant.transform` of the live reference (IPEX stubbed) vs algorithms/smooth_quant.py (kernels = oracle
    twins): same modules smoothed, same smoothing scales, same static activation qparams."""
    from neural_compressor_b200 import ops
    from neural_compressor_b200.algorithms import smooth_quant as sq
    from oracle import woq_oracle as O
    from oracle.ref_loader import load_smooth_quant_utility
    import neural_compressor_b200.quantization as ours

    SQ = load_smooth_quant_utility()

    def minmax_cols_accumulate(X, mx, mn):
        X2 = X.reshape(-1, X.shape[-1]).float()
        mx.copy_(torch.maximum(mx, X2.max(0)[0]))
        mn.copy_(torch.minimum(mn, X2.min(0)[0]))

    def sq_smooth_quant_weight(W, smooth):
        _, q, s = O.sq_qdq_weight_per_channel(W.float() * smooth.view(1, -1))
        return dict(qweight=q.to(torch.int8), w_scale=s.flatten().float(), wsum=q.sum(1).to(torch.int32))

    monkeypatch.setattr(ops, "minmax_cols_accumulate", minmax_cols_accumulate)
    monkeypatch.setattr(ops, "sq_smooth_quant_weight", sq_smooth_quant_weight)
    monkeypatch.setattr(sq, "current_device", lambda: torch.device("cpu"))
    base = build(arch)

    def q_func(model):
        for t in ids:
            model(t)

    m = copy.deepcopy(base)
    m = SQ.TorchSmoothQuant(m, dataloader=None, example_inputs=ids[0], q_func=q_func).transform(
        alpha=0.5, folding=False, calib_iter=len(ids), op_types=[torch.nn.Linear])
    want = {n: x for n, x in m.named_modules() if type(x).__name__ == "SQLinearWrapper"}
    q = sq.SmoothQuantQuantizer(ours.SmoothQuantConfig(alpha=0.5), absorb_discovery="off")
    mo = q.prepare(copy.deepcopy(base), example_inputs=ids[0])
    with torch.no_grad():
        q_func(mo)
    got = {n: x for n, x in q.convert(mo).named_modules() if isinstance(x, sq.SQLinear)}
    assert set(got) == set(want) and len(want) >= 9
    for n, ref in want.items():
        assert torch.equal(got[n].input_scale, ref.input_scale), n
        assert float(got[n].x_scale) == float(ref.scale) and int(got[n].x_zp) == int(ref.zero_point), n


def test_gptq_quant_lm_head_vs_live_reference(ref_api, monkeypatch):
    """`GPTQConfig(quant_lm_head=True)` (gptq.py:886-1078).  All block tensors must coincide with the live reference.  The
    lm_head itself cannot: in the prepare()/convert() flow the reference feeds `range(len(self.dataloader))` = 0 batches
    to the lm_head Hessian (gptq.py:283, 935), so every column counts as dead, the weight is zeroed and the packed lm_head
    dequantises to all zeros (scales = 2/15 everywhere).  Here the cached outputs of the last block are used, as that loop
    intends, and the packed lm_head is a 4-bit image of the real weight."""
    import neural_compressor_b200.quantization as ours
    from oracle import woq_oracle as O
    from oracle.gen_golden import tiny_llama
    from tests.host_twins import install_gptq_twins

    install_gptq_twins(running_mean=True, setter=monkeypatch.setattr)
    monkeypatch.setenv("B200WOQ_CALIB_BATCH", "1")
    g = torch.Generator().manual_seed(1234)
    ids = [torch.randint(0, 512, (1, 32), generator=g) for _ in range(8)]
    base = tiny_llama()
    out = {}
    for tag, api in (("ref", ref_api), ("ours", ours)):
        kw = dict(bits=4, group_size=32, use_sym=False, block_size=128, quant_lm_head=True)
        if tag == "ref":
            kw["model_path"] = "/tmp"
        m = api.prepare(copy.deepcopy(base), api.GPTQConfig(**kw))
        for x in ids:
            m(x)
        out[tag] = packed(api.convert(m))
    assert set(out["ours"]) == set(out["ref"]) and "lm_head.qweight" in out["ref"]
    blocks = {k: v for k, v in out["ref"].items() if not k.startswith("lm_head")}
    assert_same({k: out["ours"][k] for k in blocks}, blocks)
    ref_scales = out["ref"]["lm_head.scales"].float()
    assert torch.allclose(ref_scales, torch.full_like(ref_scales, 2 / 15), rtol=1e-3)        # the degenerate all-dead result
    w = O.recover_fp16(out["ours"]["lm_head.qweight"], out["ours"]["lm_head.qzeros"], out["ours"]["lm_head.scales"], 4, 32,
                       128, 512).float()
    w0 = base.lm_head.weight.detach()
    assert float((w - w0).norm() / w0.norm()) < 0.15


def test_mixed_rtn_and_gptq_composable_config_vs_live_reference(ref_api, monkeypatch):
    """The reference's test_mixed_algos.py: `RTNConfig(white_list=[".*mlp.*"]) + GPTQConfig(white_list=[".*attn.*"])` in
    one `quantize()` call -- RTN packs the MLPs first, GPTQ then calibrates the attention linears THROUGH the packed MLPs."""
    import neural_compressor.torch.algorithms.layer_wise as LW
    import neural_compressor.torch.algorithms.layer_wise.utils as LU
    import neural_compressor_b200.quantization as ours
    from neural_compressor_b200 import ops
    from neural_compressor_b200.algorithms import rtn
    from oracle import woq_oracle as O
    from oracle.gen_golden import family_models
    from tests.host_twins import install_gptq_twins

    install_gptq_twins(running_mean=True, setter=monkeypatch.setattr)
    monkeypatch.setenv("B200WOQ_CALIB_BATCH", "1")

    def rtn_quant_pack(W, bits=4, group_size=-1, sym=False, full_range=False, quantile=1.0, return_codes=False):
        q, s, z = O.rtn_quantize(W, bits, group_size, "sym" if sym else "asym", quantile, full_range)
        qweight, qzeros, scales16 = O.pack_optimum(q, s, z, bits, group_size)
        return dict(qweight=qweight, qzeros=qzeros, scales=scales16, scale_f32=s.float(), zp_f32=None if z is None else z.float())

    def woq_linear(x, qweight, qzeros, scales, bias, bits, group_size, in_features, out_features, g_idx=None,
                   input_scale=None, out_dtype=None, flags=0):
        xx = x if input_scale is None else x * input_scale
        y = O.woq_linear_forward(xx, qweight, qzeros, scales, bias, bits, group_size, in_features, out_features, g_idx)
        return y.to(out_dtype or x.dtype)

    monkeypatch.setattr(ops, "rtn_quant_pack", rtn_quant_pack)
    monkeypatch.setattr(ops, "woq_linear", woq_linear)
    monkeypatch.setattr(rtn, "current_device", lambda: torch.device("cpu"))
    # a random-init model has no checkpoint path for the reference's layer-wise helper to resolve
    import neural_compressor.torch.algorithms.weight_only.gptq as RG

    for mod in (LU, LW, RG):
        if hasattr(mod, "get_path"):
            monkeypatch.setattr(mod, "get_path", lambda p: "/tmp")
    base = family_models()("gptj")
    tokens = torch.tensor([[10, 20, 30, 40, 50, 60, 70, 80, 90, 100, 110, 120]], dtype=torch.long)
    out = {}
    for tag, api in (("ref", ref_api), ("ours", ours)):
        cfg = api.RTNConfig(white_list=[".*mlp.*"]) + api.GPTQConfig(double_quant_bits=4, white_list=[".*attn.*"])
        m = api.quantize(copy.deepcopy(base), cfg, run_fn=lambda model: model(tokens))
        out[tag] = packed(m)
    assert len(out["ref"]) == 36
    assert_same(out["ours"], out["ref"])


def test_exotic_parameter_corners_vs_live_reference(ref_api, monkeypatch):
    """Corners of the reference's parameter matrices (test_gptq.py:106-135, test_awq.py:60-84) on the tiny llama: 2-bit
    group-8 act_order GPTQ with the default block_size 2048."""
    import neural_compressor_b200.quantization as ours
    from oracle.gen_golden import tiny_llama
    from tests.host_twins import install_gptq_twins

    install_gptq_twins(running_mean=True, setter=monkeypatch.setattr)
    monkeypatch.setenv("B200WOQ_CALIB_BATCH", "1")
    g = torch.Generator().manual_seed(1234)
    ids = [torch.randint(0, 512, (1, 32), generator=g) for _ in range(8)]
    base = tiny_llama()
    out = {}
    for tag, api in (("ref", ref_api), ("ours", ours)):
        kw = dict(bits=2, use_sym=True, group_size=8, act_order=True)
        if tag == "ref":
            kw["model_path"] = "/tmp"
        m = api.prepare(copy.deepcopy(base), api.GPTQConfig(**kw))
        for x in ids:
            m(x)
        m = api.convert(m)
        out[tag] = {k: v for k, v in m.state_dict().items() if k.rsplit(".", 1)[-1] in ("qweight", "qzeros", "scales", "g_idx")}
    assert set(out["ours"]) == set(out["ref"]) and len(out["ref"]) == 56
    for k, ref in out["ref"].items():
        if k.endswith("g_idx"):
            assert torch.equal(out["ours"][k], ref), k
        elif ref.dtype == torch.int32:
            assert (fields(out["ours"][k], 2) != fields(ref, 2)).float().mean().item() <= 2e-3, k
        else:
            assert torch.allclose(out["ours"][k].float(), ref.float(), rtol=2e-3, atol=1e-8), k
