"""Host logic of the AWQ engine on the tiny llama (algorithms/awq.py: block capture, per-module input collection, the 20
candidate scales and 10 clip ratios, self-absorbing MulLinear insertion, final RTN with the searched quantiles) with the
device kernels replaced by the oracle's CPU twins, against the packed state dicts of the UNMODIFIED reference
(tests/golden/e2e_tiny_llama.pt, options_matrix.pt).  The GPU tests compare the same fixtures with the kernels in place."""
import pytest
import torch

from tests.test_awq_absorb_cpu import host_ops  # noqa: F401  (fixture)


def run_awq(golden_e2e, kw):
    import neural_compressor_b200.quantization as api
    from tests.test_api_gpu import tiny_llama

    ids = golden_e2e["ids"]

    def run_fn(model):
        for x in ids:
            model(x)

    return api.quantize(tiny_llama(golden_e2e["init_state"]), api.AWQConfig(**kw), run_fn=run_fn, example_inputs=ids[0])


def compare(m, state):
    got = m.state_dict()
    packed = [k for k in state if k.rsplit(".", 1)[-1] in ("qweight", "qzeros", "scales", "input_scale")]
    assert len(packed) >= 14
    ours = {k for k in got if k.rsplit(".", 1)[-1] in ("qweight", "qzeros", "scales", "input_scale")}
    assert ours == set(packed), sorted(ours ^ set(packed))[:6]     # same modules packed / wrapped, no more, no fewer
    for k in packed:
        ref = state[k]
        assert k in got and got[k].shape == ref.shape and got[k].dtype == ref.dtype, k
        if ref.dtype == torch.int32:
            assert torch.equal(got[k], ref), k
        else:
            assert torch.allclose(got[k].float(), ref.float(), rtol=1e-6, atol=1e-8), k


def test_awq_e2e_host_flow(host_ops, golden_e2e):  # noqa: F811
    m = run_awq(golden_e2e, dict(bits=4, group_size=32, use_sym=False))
    compare(m, golden_e2e["awq"]["state"])
    types = {n: type(x).__name__ for n, x in m.named_modules()}
    want = golden_e2e["awq"]["module_types"]
    assert {n for n, t in types.items() if t == "MulLinear"} == {n for n, t in want.items() if t == "MulLinear"}


@pytest.mark.parametrize("tag", ["awq_sym_noclip", "awq_noscale"])
def test_awq_option_cases_host_flow(host_ops, golden_e2e, golden_options, tag):  # noqa: F811
    case = golden_options["cases"][tag]
    m = run_awq(golden_e2e, case["kw"])
    compare(m, case["state"])
