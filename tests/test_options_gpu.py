"""Option surface of the reference's own test matrix (test_rtn.py:106-165, test_gptq.py:106-185, test_awq.py:60-84)
through the mirrored public API, against tensors produced by the UNMODIFIED reference on the CPU
(tests/golden/options_matrix.pt, oracle/gen_golden.py options).  RTN cases must be bit-exact; GPTQ / AWQ cases use the
bars of test_api_gpu.py (summation-order dependent)."""
import pytest
import torch

from tests.test_api_gpu import DEV, tiny_llama

pytestmark = pytest.mark.gpu


def fields(q, bits):
    """Unpack an int32 word tensor into its bit fields along a new last axis (n_pack = 32 // bits fields per word)."""
    q = q.cpu().to(torch.int64) & 0xFFFFFFFF
    return torch.stack([(q >> (bits * e)) & ((1 << bits) - 1) for e in range(32 // bits)], dim=-1)


def compare(model, golden_state, bits, exact):
    state = model.state_dict()
    worst = dict(code=0.0, zero=0.0, scale=0.0)
    n = 0
    for k, ref in golden_state.items():
        assert k in state, k
        got = state[k]
        assert got.shape == ref.shape and got.dtype == ref.dtype, (k, got.shape, ref.shape, got.dtype, ref.dtype)
        if k.endswith("qweight"):
            n += 1
            worst["code"] = max(worst["code"], (fields(got, bits) != fields(ref, bits)).float().mean().item())
        elif k.endswith("qzeros"):
            worst["zero"] = max(worst["zero"], (fields(got, bits) != fields(ref, bits)).float().mean().item())
        elif k.endswith("scales"):
            worst["scale"] = max(worst["scale"], (got.cpu().float() - ref.float()).abs().max().item())
        elif k.endswith("g_idx"):
            if exact:
                assert torch.equal(got.cpu(), ref), k
            else:  # act_order: argsort of diag(H); near-ties may swap neighbours -- measured, not assumed
                worst["g_idx"] = max(worst.get("g_idx", 0.0), (got.cpu() != ref).float().mean().item())
    assert n > 0
    if exact:
        assert worst == dict(code=0.0, zero=0.0, scale=0.0), worst
    return worst


def bits_of(kw):
    return int(kw["dtype"].lstrip("int")) if "dtype" in kw else kw["bits"]


def case_ids(algo):
    from oracle.gen_golden import OPTION_CASES  # the case table only (plain data; nothing from the reference is imported)

    return [t for t, a, _ in OPTION_CASES if a == algo]


@pytest.fixture(scope="module")
def api():
    import neural_compressor_b200.quantization as q

    return q


@pytest.mark.parametrize("tag", case_ids("rtn"))
def test_rtn_options_bit_exact(api, golden_e2e, golden_options, tag):
    case = golden_options["cases"][tag]
    m = tiny_llama(golden_e2e["init_state"]).to(DEV)
    m = api.convert(api.prepare(m, api.RTNConfig(use_layer_wise=False, **case["kw"])))
    compare(m, case["state"], bits_of(case["kw"]), exact=True)
    with torch.no_grad():
        logits = m(golden_e2e["probe"].to(DEV)).logits.float().cpu()
    ref = case["logits"]
    assert (logits - ref).abs().max().item() < 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("tag", case_ids("gptq"))
def test_gptq_options(api, golden_e2e, golden_options, tag, parity_log):
    case = golden_options["cases"][tag]
    m = tiny_llama(golden_e2e["init_state"]).to(DEV)
    m = api.prepare(m, api.GPTQConfig(**case["kw"]))
    for x in golden_e2e["ids"]:
        m(x.to(DEV))
    m = api.convert(m)
    from tests.conftest import parity_bound

    if case["kw"].get("act_order"):
        # the permutation is an argsort of diag(H): g_idx must equal the reference's (compare() asserts it), after which
        # the packed codes are compared position by position like every other case
        mod = m.model.layers[0].self_attn.q_proj
        assert mod.g_idx is not None and mod.g_idx.dtype == torch.int32
    worst = compare(m, case["state"], bits_of(case["kw"]), exact=False)
    parity_log(f"options/{tag}", worst)
    print(tag, worst)
    assert worst["code"] <= parity_bound(f"options/{tag}", "code", 6e-2), worst
    assert worst["scale"] <= 2e-3 and worst["zero"] <= parity_bound(f"options/{tag}", "zero", 6e-2), worst
    with torch.no_grad():
        logits = m(golden_e2e["probe"].to(DEV)).logits.float().cpu()
    ref = case["logits"]
    assert (logits - ref).abs().max().item() < 5e-2 * ref.abs().max().item()


@pytest.mark.parametrize("tag", case_ids("awq"))
def test_awq_options(api, golden_e2e, golden_options, tag, parity_log):
    case = golden_options["cases"][tag]
    ids = golden_e2e["ids"]

    def run_fn(model):
        for x in ids:
            model(x.to(DEV))

    m = tiny_llama(golden_e2e["init_state"]).to(DEV)
    m = api.quantize(m, api.AWQConfig(**case["kw"]), run_fn=run_fn, example_inputs=ids[0].to(DEV))
    from tests.conftest import parity_bound

    state = m.state_dict()
    worst_code = 0.0
    for k, ref in case["state"].items():
        assert k in state, k
        if k.endswith("input_scale"):
            rel = (state[k].cpu().float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
            assert rel < 1e-3, (k, rel)
        if k.endswith("qweight"):   # packed codes, not only logits
            worst_code = max(worst_code, (fields(state[k], 4) != fields(ref, 4)).float().mean().item())
    parity_log(f"options/{tag}", dict(code=worst_code))
    assert worst_code <= parity_bound(f"options/{tag}", "code", 3e-2), worst_code
    with torch.no_grad():
        logits = m(golden_e2e["probe"].to(DEV)).logits.float().cpu()
    ref = case["logits"]
    assert (logits - ref).abs().max().item() < 5e-2 * ref.abs().max().item()


def test_rtn_conv1d_model_bit_exact(api, golden_options):
    """GPT-2's transformers.Conv1D stores [in, out]: the packed module must come out identical (rtn.py:209-216)."""
    from transformers import GPT2Config, GPT2LMHeadModel

    g2 = GPT2LMHeadModel(GPT2Config(n_embd=64, n_layer=2, n_head=2, vocab_size=256, n_positions=64)).eval()
    g2.load_state_dict(golden_options["gpt2_init"])
    g2 = g2.to(DEV)
    g2 = api.convert(api.prepare(g2, api.RTNConfig(bits=4, group_size=32, use_sym=False, use_layer_wise=False)))
    compare(g2, golden_options["gpt2_rtn"]["state"], 4, exact=True)
    with torch.no_grad():
        logits = g2(golden_options["gpt2_probe"].to(DEV)).logits.float().cpu()
    ref = golden_options["gpt2_rtn"]["logits"]
    assert (logits - ref).abs().max().item() < 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("tag", ["gptq_double_quant", "gptq_double_quant_sym_g64"])
def test_gptq_double_quant(api, golden_e2e, golden_options_extra, tag, parity_log):
    """use_double_quant: each group's scales are fake-quantised over the output rows (gptq.py:1598-1614), against the
    packed tensors of the live reference (tests/golden/options_extra.pt)."""
    case = golden_options_extra["cases"][tag]
    m = tiny_llama(golden_e2e["init_state"]).to(DEV)
    m = api.prepare(m, api.GPTQConfig(**case["kw"]))
    for x in golden_e2e["ids"]:
        m(x.to(DEV))
    m = api.convert(m)
    worst = compare(m, case["state"], bits_of(case["kw"]), exact=False)
    parity_log(f"options/{tag}", worst)
    assert worst["code"] <= 3e-2 and worst["scale"] <= 1e-3 and worst["zero"] <= 3e-2, worst
    # the double-quantised scales take few distinct values per 256-row group: compare the layer-0 scales exactly-ish
    with torch.no_grad():
        logits = m(golden_e2e["probe"].to(DEV)).logits.float().cpu()
    ref = case["logits"]
    assert (logits - ref).abs().max().item() < 5e-2 * ref.abs().max().item()


def test_static_groups_matches_the_reference_failure(api, golden_e2e, golden_options_extra):
    """`static_groups=True` with group_size < in_features cannot run in the reference either: its fasterquant returns a
    single scale column (gptq.py:1193-1200, 1339-1341) and the export raises IndexError (recorded from the live reference
    in the fixture).  We fail loudly at the same place instead of quantising differently."""
    assert golden_options_extra["static_groups_reference"].startswith("IndexError")
    m = tiny_llama(golden_e2e["init_state"]).to(DEV)
    m = api.prepare(m, api.GPTQConfig(bits=4, group_size=32, static_groups=True))
    for x in golden_e2e["ids"][:2]:
        m(x.to(DEV))
    with pytest.raises(NotImplementedError):
        api.convert(m)
