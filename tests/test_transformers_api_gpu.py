"""HF-style entry: `neural_compressor_b200.transformers.AutoModelForCausalLM.from_pretrained(dir,
quantization_config=...)` must give the same packed tensors as the reference's torch API did on the CPU (fixtures of
tests/golden/e2e_tiny_llama.pt), and its `save_pretrained` / `from_pretrained` must round-trip."""
import pytest
import torch

from tests.test_api_gpu import DEV, compare_state, tiny_llama

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fp_dir(golden_e2e, tmp_path):
    m = tiny_llama(golden_e2e["init_state"])
    m.save_pretrained(str(tmp_path / "fp"))
    return str(tmp_path / "fp")


def test_rtn_from_pretrained_bit_exact(golden_e2e, fp_dir, tmp_path):
    import neural_compressor_b200.transformers as T

    m = T.AutoModelForCausalLM.from_pretrained(fp_dir, quantization_config=T.RtnConfig(bits=4, group_size=32, sym=True))
    compare_state(m, golden_e2e["rtn"]["state"], exact=True)
    assert isinstance(m.lm_head, torch.nn.Linear)  # modules_to_not_convert default (quantization_config.py:266-270)
    m.save_pretrained(str(tmp_path / "q"))
    m2 = T.AutoModelForCausalLM.from_pretrained(str(tmp_path / "q"))
    with torch.no_grad():
        a = m(golden_e2e["probe"].to(DEV)).logits
        b = m2(golden_e2e["probe"].to(DEV)).logits
    assert torch.equal(a, b)


def test_gptq_from_pretrained(golden_e2e, fp_dir):
    import neural_compressor_b200.transformers as T

    cfg = T.GPTQConfig(bits=4, group_size=32, sym=True, blocksize=128, damp_percent=0.01, dataset=golden_e2e["ids"],
                       n_samples=16, seq_len=64)
    m = T.AutoModelForCausalLM.from_pretrained(fp_dir, quantization_config=cfg)
    worst = compare_state(m, golden_e2e["gptq"]["state"], exact=False)
    assert worst["code_mismatch"] <= 2e-2 and worst["scale_diff"] <= 1e-3, worst


def test_awq_from_pretrained_and_errors(golden_e2e, fp_dir):
    import neural_compressor_b200.transformers as T

    cfg = T.AwqConfig(bits=4, group_size=32, zero_point=True, dataset=golden_e2e["ids"], n_samples=16, seq_len=64)
    m = T.AutoModelForCausalLM.from_pretrained(fp_dir, quantization_config=cfg)
    with torch.no_grad():
        logits = m(golden_e2e["probe"].to(DEV)).logits.float().cpu()
    ref = golden_e2e["awq"]["logits"]
    assert (logits - ref).abs().max().item() < 5e-2 * ref.abs().max().item()
    with pytest.raises(ValueError):  # hub dataset names cannot be resolved offline
        T.AutoModelForCausalLM.from_pretrained(fp_dir, quantization_config=T.GPTQConfig(bits=4))
    with pytest.raises(ValueError):
        T.GPTQConfig(bits=3)
