"""SmoothQuant scale-sharing groups and folding on the B200 (TorchSmoothQuant._parse_absorb_to_layers :2225-2287,
_cal_scales :2122-2156, _absorb_scales :1994-2061) against `TorchSmoothQuant.transform` of the live reference run on the
CPU with IPEX stubbed (tests/golden/sq_transform.pt).  The host flow is pinned exactly in
tests/test_smoothquant_transform_cpu.py; here the calibration / weight-quantisation / W8A8 kernels run."""
import os

import pytest
import torch

from oracle import woq_oracle as O
from tests.toy_models import Toy

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DEV = "cuda"


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(HERE, "golden", "sq_transform.pt"))


@pytest.fixture(scope="module")
def toy_data():
    return torch.load(os.path.join(HERE, "golden", "awq_toy.pt"))


def quantize_toy(toy_data, mode, **cfg):
    import neural_compressor_b200.quantization as api
    from neural_compressor_b200.algorithms.smooth_quant import SmoothQuantQuantizer

    m = Toy(d=64, n=2, variant=0, vocab=64).eval()
    m.load_state_dict(toy_data["init_state"])
    q = SmoothQuantQuantizer(api.SmoothQuantConfig(**cfg), absorb_discovery=mode)
    m = q.prepare(m.to(DEV), example_inputs=toy_data["ids"][0].to(DEV))
    with torch.no_grad():
        for t in toy_data["ids"]:
            m(t.to(DEV))
    return q.convert(m)


def codes_close(m, ref_weight, name):
    _, q_ref, s_ref = O.sq_qdq_weight_per_channel(ref_weight)
    q = m.qweight[:, :m.in_features].float().cpu()
    assert torch.allclose(m.w_scale.cpu().view(-1, 1), s_ref, rtol=1e-5, atol=0), name
    diff = (q - q_ref).abs()
    assert diff.max() <= 1 and (diff > 0).float().mean() <= 5e-3, (name, diff.max().item(), (diff > 0).float().mean().item())


def logits_close(m, toy_data, ref, tol):
    with torch.no_grad():
        y = m(toy_data["probe"].to(DEV)).float().cpu()
    assert torch.isfinite(y).all()
    assert float((y - ref).norm() / ref.norm()) < tol


def test_scale_sharing_groups(golden, toy_data):
    from neural_compressor_b200.algorithms.smooth_quant import SQLinear

    case = golden["models"]["toy_insert_mul"]
    m = quantize_toy(toy_data, "eager", alpha=0.5, folding=False)
    mods = dict(m.named_modules())
    assert {n for n, x in mods.items() if isinstance(x, SQLinear)} == set(case["wrappers"])
    for name, wr in case["wrappers"].items():
        x = mods[name]
        assert torch.allclose(x.input_scale.cpu(), wr["input_scale"], rtol=1e-5, atol=0), name
        assert abs(float(x.x_scale) - float(wr["scale"])) <= 1e-5 * float(wr["scale"]) and int(x.x_zp) == int(wr["zero_point"]), name
        codes_close(x, wr["weight"], name)
    b = m.layers[0]
    assert torch.equal(b.q.input_scale, b.k.input_scale) and torch.equal(b.q.input_scale, b.v.input_scale)
    logits_close(m, toy_data, case["logits"], 0.1)     # W8A8 against the smoothed fp model of the reference


def test_folding(golden, toy_data):
    from neural_compressor_b200.algorithms.smooth_quant import SQLinear

    case = golden["models"]["toy_folding"]
    m = quantize_toy(toy_data, "off", alpha=0.5, folding=True)
    mods = dict(m.named_modules())
    smoothed = {n for v in case["absorb_to_layer"].values() for n in v}
    assert {n for n, x in mods.items() if isinstance(x, SQLinear)} == smoothed
    st = case["state"]
    for name in ("layers.0.ln1", "layers.1.ln2", "norm"):
        assert torch.allclose(mods[name].weight.cpu(), st[name + ".weight"], rtol=1e-5, atol=0), name
    for n in smoothed:
        assert mods[n].folded
        codes_close(mods[n], st[n + ".weight"], n)
    assert isinstance(mods["layers.0.o"], torch.nn.Linear)
    logits_close(m, toy_data, case["logits"], 0.1)
