"""`hybrid_order=True` (gptq.py:1203-1209, 1320-1328, 1389-1474) on the B200 against the UNMODIFIED reference on the CPU
(tests/golden/gptq_hybrid.pt).  The host flow (the permutation, un-permutation and group-parameter re-ordering around the
column loop) is pinned bit for bit in tests/test_gptq_hostflow_cpu.py; here the K1/K2/K3 kernels run, so codes carry the
same summation-order sensitivity as every other GPTQ case (tests/test_options_gpu.py)."""
import os

import pytest
import torch

from tests.test_api_gpu import DEV, tiny_llama
from tests.test_options_gpu import compare

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("tag", ["hybrid_sym", "hybrid_asym_g64"])
def test_gptq_hybrid_order(golden_e2e, tag, parity_log):
    import neural_compressor_b200.quantization as api

    g = torch.load(os.path.join(HERE, "golden", "gptq_hybrid.pt"))
    case = g["cases"][tag]
    m = tiny_llama(golden_e2e["init_state"]).to(DEV)
    m = api.prepare(m, api.GPTQConfig(**case["kw"]))
    for x in golden_e2e["ids"]:
        m(x.to(DEV))
    m = api.convert(m)
    assert m.model.layers[0].self_attn.q_proj.g_idx is None
    worst = compare(m, case["state"], 4, exact=False)
    parity_log(f"gptq_hybrid/{tag}", worst)
    print(tag, worst)
    assert worst["code"] <= 6e-2 and worst["zero"] <= 6e-2 and worst["scale"] <= 2e-3, worst
    with torch.no_grad():
        logits = m(g["probe"].to(DEV)).logits.float().cpu()
    ref = case["logits"]
    assert (logits - ref).abs().max().item() < 5e-2 * ref.abs().max().item()
