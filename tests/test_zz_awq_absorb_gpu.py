"""AWQ with discovered absorb layers on the B200 (awq.py:40-95 grouping, :264-361 multi-module scale search through the
block, :364-391 folding into the absorbing LayerNorm / Linear) against the packed model the UNMODIFIED reference produced
on the CPU for the same plain-nn.Module transformer (tests/golden/awq_toy.pt; the reference's torch.jit tracer cannot
trace transformers >= 5 models any more, tests/toy_models.py is what it still can).  The host flow is pinned exactly on
the CPU (tests/test_awq_absorb_cpu.py); here the search kernels run, so codes are compared with a measured bound."""
import os

import pytest
import torch

from tests.test_options_gpu import fields
from tests.toy_models import Toy

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DEV = "cuda"


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(HERE, "golden", "awq_toy.pt"))


@pytest.mark.parametrize("tag", ["folding_false", "folding_true", "folding_true_sym"])
def test_awq_discovered_absorption(golden, tag, parity_log, monkeypatch):
    import neural_compressor_b200.quantization as api
    from neural_compressor_b200.algorithms.modules import B200WeightOnlyLinear, MulLinear

    kw = dict(folding_false=dict(folding=False, use_sym=False, group_size=32),
              folding_true=dict(folding=True, use_sym=False, group_size=32),
              folding_true_sym=dict(folding=True, use_sym=True, group_size=64))[tag]
    monkeypatch.setenv("B200WOQ_AWQ_ABSORB", "eager")
    ids = golden["ids"]

    def run_fn(model):
        for x in ids:
            model(x.to(DEV))

    m = Toy(d=64, n=2, variant=0, vocab=64).eval()
    m.load_state_dict(golden["init_state"])
    m = api.quantize(m.to(DEV), api.AWQConfig(bits=4, **kw), run_fn=run_fn, example_inputs=ids[0].to(DEV))
    got = {k: v for k, v in m.state_dict().items() if "bf16_to_fp8" not in k}
    want = golden["cases"][tag]["state"]
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    # structure: q/k/v fold into ln1, fc1 into ln2, fc2 into fc1 (through the ReLU); `o` has no absorber
    blk = m.layers[0]
    assert isinstance(blk.q, B200WeightOnlyLinear) and isinstance(blk.fc2, B200WeightOnlyLinear)
    assert isinstance(blk.o, MulLinear) == (tag == "folding_false")
    assert not torch.equal(blk.ln1.weight.cpu(), golden["init_state"]["layers.0.ln1.weight"])
    worst_code, worst_fold = 0.0, 0.0
    for k, ref in want.items():
        if k.endswith("qweight"):
            worst_code = max(worst_code, (fields(got[k], 4) != fields(ref, 4)).float().mean().item())
        elif "ln" in k or k.endswith("input_scale"):
            den = ref.float().abs().max().clamp_min(1e-12)
            worst_fold = max(worst_fold, ((got[k].cpu().float() - ref.float()).abs().max() / den).item())
    parity_log(f"awq_toy/{tag}", dict(code=worst_code, folded_rel=worst_fold))
    print(tag, worst_code, worst_fold)
    assert worst_fold <= 1e-3, worst_fold       # same alpha chosen for every tuple, scales equal up to reduction order
    assert worst_code <= 5e-2, worst_code
    with torch.no_grad():
        logits = m(golden["probe"].to(DEV)).float().cpu()
    ref = golden["cases"][tag]["logits"]
    assert (logits - ref).abs().max().item() < 5e-2 * ref.abs().max().item()
