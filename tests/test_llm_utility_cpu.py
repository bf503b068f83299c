"""f4 -- the step either side of the path (llm_utility.py:17-126, examples .../weight_only/utils.py:8-198): calibration-set
selection against the reference's own DataloaderPreprocessor when the reference tree is present, the text dataloader,
the benchmark helper and the checkpoint-backed layer-sharded build.  Host logic only."""
import importlib.util
import os
import random

import pytest
import torch

from neural_compressor_b200.utils import llm_utility as U

REF_UTILS = ("/root/reference/examples/pytorch/nlp/huggingface_models/language-modeling/quantization/weight_only/utils.py")


def ragged_batches(kind, seed=0):
    g = random.Random(seed)
    out = []
    for _ in range(24):
        n = g.choice([16, 48, 64, 64, 100, 200])
        ids = torch.randint(0, 1000, (1, n), generator=torch.Generator().manual_seed(g.randrange(1 << 30)))
        if kind == "tensor":
            out.append(ids)
        elif kind == "dict":
            out.append({"input_ids": ids, "attention_mask": torch.ones_like(ids), "tag": n})
        else:
            out.append((ids, torch.ones_like(ids), torch.tensor([n])))
    return out


def flat(sample):
    if isinstance(sample, dict):
        return [sample[k] for k in sorted(sample)]
    if isinstance(sample, (list, tuple)):
        return list(sample)
    return [sample]


def same(a, b):
    fa, fb = flat(a), flat(b)
    return len(fa) == len(fb) and all(torch.equal(x, y) if isinstance(x, torch.Tensor) else x == y for x, y in zip(fa, fb))


@pytest.mark.parametrize("kind", ["tensor", "dict", "tuple"])
@pytest.mark.parametrize("use_max_length", [False, True])
def test_calibration_selection(kind, use_max_length):
    data = ragged_batches(kind)
    ours = U.DataloaderPreprocessor(data, use_max_length=use_max_length, max_seq_length=64, nsamples=10)
    got = ours.get_prepared_dataloader()
    assert len(got) == 10 and ours.is_ready
    for s in got:
        n = U._seq_len_of(s)
        assert n == 64 if use_max_length else n <= 64
    if not os.path.exists(REF_UTILS):
        pytest.skip("reference examples not present: structural checks only")
    spec = importlib.util.spec_from_file_location("ref_example_utils", REF_UTILS)
    ref_mod = importlib.util.module_from_spec(spec)
    from oracle.ref_loader import load_reference

    load_reference()   # puts `neural_compressor` (the example's imports) on the path, with the image's missing-package stubs
    spec.loader.exec_module(ref_mod)
    ref = ref_mod.DataloaderPreprocessor(ragged_batches(kind), use_max_length=use_max_length, max_seq_length=64, nsamples=10)
    want = ref.get_prepared_dataloader()
    assert len(want) == len(got) and all(same(a, b) for a, b in zip(got, want))


class ToyTokenizer:
    def __call__(self, text, max_length, padding, truncation, return_tensors):
        ids = [ord(c) % 97 + 1 for c in text][:max_length]
        mask = [1] * len(ids) + [0] * (max_length - len(ids))
        ids = ids + [0] * (max_length - len(ids))
        return {"input_ids": torch.tensor([ids]), "attention_mask": torch.tensor([mask])}


def test_default_dataloader_from_records():
    records = [{"text": "sample %d " % i * (i + 1)} for i in range(40)]
    dl = U.get_default_llm_dataloader(ToyTokenizer(), bs=4, nsamples=12, seq_len=32, seed=1, dataset=records)
    batches = list(dl)
    assert len(batches) == 3 and batches[0]["input_ids"].shape == (4, 32) and batches[0]["attention_mask"].shape == (4, 32)
    again = U.get_default_llm_dataloader(ToyTokenizer(), bs=4, nsamples=12, seq_len=32, seed=1, dataset=records)
    assert sorted(r["text"] for r in dl.dataset.records) == sorted(r["text"] for r in again.dataset.records)


def tiny_llama():
    from transformers import LlamaConfig, LlamaForCausalLM

    torch.manual_seed(0)
    cfg = LlamaConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=4, num_attention_heads=2,
                      num_key_value_heads=2, vocab_size=128, max_position_embeddings=64, tie_word_embeddings=False)
    return LlamaForCausalLM(cfg).eval()


def test_benchmark_perplexity_and_example_inputs():
    m = tiny_llama()
    r = U.llm_benchmark(m, batch_size=2, input_length=8, warmup_iters=1, total_iters=3)
    assert r["latency_s"] > 0 and r["throughput_samples_per_s"] > 0
    ppl = U.evaluate_perplexity(m, torch.randint(0, 128, (200,)), seq_len=32, batch_size=2)
    assert 50 < ppl < 400    # random-init model over a 128-token vocabulary: close to uniform
    ex = U.get_example_inputs(m, [(torch.ones(1, 4, dtype=torch.long), torch.zeros(1))])
    assert torch.equal(ex, torch.ones(1, 4, dtype=torch.long))
    ex = U.get_example_inputs(m, [{"input_ids": torch.ones(1, 4, dtype=torch.long), "label": torch.zeros(1)}])
    assert set(ex) == {"input_ids"}
    U.run_calibration(m, [torch.ones(1, 4, dtype=torch.long), {"input_ids": torch.ones(1, 4, dtype=torch.long)}])


def test_checkpoint_backed_layer_sharded_build(tmp_path):
    """Each rank materialises its own blocks plus the replicated modules straight from the safetensors files."""
    from transformers import AutoConfig, AutoModelForCausalLM

    from neural_compressor_b200.utils.sharded import build_layer_sharded, checkpoint_init, is_remote

    m = tiny_llama()
    m.save_pretrained(tmp_path, safe_serialization=True)
    cfg = AutoConfig.from_pretrained(tmp_path)
    ref_state = m.state_dict()
    for rank in (0, 1):
        sharded = build_layer_sharded(lambda: AutoModelForCausalLM.from_config(cfg), "model.layers", rank, 2, "cpu",
                                      init=checkpoint_init(str(tmp_path)))
        lo, hi = sharded._b200_shard["owned"]
        assert (lo, hi) == ((0, 2) if rank == 0 else (2, 4))
        for i, blk in enumerate(sharded.model.layers):
            assert is_remote(blk) == (not lo <= i < hi)
        for k, v in sharded.state_dict().items():
            if not v.is_meta:
                assert torch.equal(v, ref_state[k]), k
        assert not sharded.lm_head.weight.is_meta and not sharded.model.embed_tokens.weight.is_meta
