"""The step either side of the quantization path: calibration data in, latency / perplexity out.

Reference:
    neural_compressor/torch/utils/llm_utility.py
        initialize_model_and_tokenizer :17-46    update_tokenizer :49-59
        get_default_llm_dataloader     :62-104   llm_benchmark    :107-126
    examples/pytorch/nlp/huggingface_models/language-modeling/quantization/weight_only/utils.py
        DataloaderPreprocessor :8-148 (first-n / full-length calibration selection, seeded random crops)

Same names and argument meaning as the reference.  What differs is where the work runs: the model is built on the
B200 (or layer-sharded over the ranks of the job, utils/sharded.py, instead of DeepSpeed tensor parallelism) and the
benchmark is timed on the device with CUDA events around the timed iterations rather than with a host clock.
"""
from __future__ import annotations

import random

import torch

from .utility import current_device, logger, move_to_device


def update_tokenizer(model, tokenizer):
    """llm_utility.py:49-59 -- the decapoda-research special-token fix for llama / mixtral checkpoints."""
    if model.config.model_type in ("llama", "mixtral"):
        gen = model.generation_config
        gen.pad_token_id, gen.bos_token_id, gen.eos_token_id = 0, 1, 2
        tokenizer.bos_token_id, tokenizer.eos_token_id, tokenizer.pad_token_id = 1, 2, 0
    return model, tokenizer


def initialize_model_and_tokenizer(model_name_or_path, use_load=False, device=None):
    """llm_utility.py:17-46.  `use_load=True` reads an already quantized HuggingFace-format checkpoint through
    `quantization.load`; otherwise the fp checkpoint is loaded in its own dtype.  With more than one rank in the job the
    decoder blocks are partitioned over the ranks (utils/sharded.py) where the reference shards with DeepSpeed."""
    import transformers

    device = current_device() if device is None else torch.device(device)
    tokenizer = transformers.AutoTokenizer.from_pretrained(model_name_or_path)
    if use_load:
        from ..quantization import load

        model = load(model_name_or_path, format="huggingface", device=device)
        return update_tokenizer(model, tokenizer)
    config = transformers.AutoConfig.from_pretrained(model_name_or_path)
    dtype = getattr(config, "torch_dtype", None) or getattr(config, "dtype", None)
    world = torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1
    if world > 1:
        from .sharded import build_layer_sharded, checkpoint_init
        from .utility import get_block_prefix

        def factory():  # runs under torch.device("meta"): shapes only
            return transformers.AutoModelForCausalLM.from_config(config, torch_dtype=dtype)

        with torch.device("meta"):
            blocks_attr = get_block_prefix(factory())[0]
        model = build_layer_sharded(factory, blocks_attr, torch.distributed.get_rank(), world, device,
                                    init=checkpoint_init(model_name_or_path))
    else:
        model = transformers.AutoModelForCausalLM.from_pretrained(model_name_or_path, torch_dtype=dtype).to(device)
    model, tokenizer = update_tokenizer(model, tokenizer)
    return model.eval(), tokenizer


class _TokenizedTexts(torch.utils.data.Dataset):
    """Pads / truncates every text to `seq_len` tokens (llm_utility.py:82-99)."""

    def __init__(self, records, tokenizer, seq_len):
        self.records, self.tokenizer, self.seq_len = records, tokenizer, seq_len

    def __len__(self):
        return len(self.records)

    def __getitem__(self, idx):
        enc = self.tokenizer(self.records[idx]["text"], max_length=self.seq_len, padding="max_length", truncation=True,
                             return_tensors="pt")
        return {k: v.squeeze(0) for k, v in enc.items()}


def get_default_llm_dataloader(tokenizer, dataset_name="NeelNanda/pile-10k", bs=8, nsamples=128, seq_len=128, seed=42,
                               dataset=None):
    """llm_utility.py:62-104: `nsamples` shuffled texts of `dataset_name`'s train split, tokenized to `seq_len`.

    `dataset` (a sequence of {"text": ...} records, or a `datasets.Dataset`) replaces the hub download -- the GPU boxes
    have no network; with a plain sequence the shuffle uses `random.Random(seed)`."""
    if dataset is None:
        from datasets import load_dataset

        dataset = load_dataset(dataset_name, split="train")
    if hasattr(dataset, "shuffle") and hasattr(dataset, "select"):
        records = dataset.shuffle(seed=seed).select(range(min(nsamples, len(dataset))))
    else:
        order = list(range(len(dataset)))
        random.Random(seed).shuffle(order)
        records = [dataset[i] for i in order[:nsamples]]
    return torch.utils.data.DataLoader(_TokenizedTexts(records, tokenizer, seq_len), batch_size=bs, shuffle=True)


# ------------------------------------------------------------------------------------------------
# calibration-set selection (the examples' DataloaderPreprocessor)
# ------------------------------------------------------------------------------------------------
def _seq_len_of(batch):
    if isinstance(batch, (list, tuple)):
        return batch[0].shape[-1]
    if isinstance(batch, dict):
        return batch["input_ids"].shape[-1]
    return batch.shape[-1]


def _crop(batch, i, j):
    """Slice the sequence axis of every 2-D tensor of a list / dict / tensor batch to [i, j)."""
    if isinstance(batch, (list, tuple)):
        return [t[:, i:j] if isinstance(t, torch.Tensor) and t.dim() == 2 else t for t in batch]
    if isinstance(batch, dict):
        return {k: (v[:, i:j] if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    return batch[:, i:j]


class DataloaderPreprocessor:
    """Collects the first `nsamples` batches of a dataloader as the calibration set (examples utils.py:8-148).

    use_max_length=False: every batch is kept; longer ones are cropped to a random `max_seq_length` window.
    use_max_length=True (the GPTQ authors' scheme): batches shorter than `max_seq_length` are skipped so every
    calibration token is real.  The crop offsets come from `random.seed(seed)`, one `randint(0, L - max - 1)` per cropped
    batch, i.e. the same windows as the reference for the same data order."""

    def __init__(self, dataloader_original, use_max_length=False, max_seq_length=2048, nsamples=128):
        self.dataloader_original = dataloader_original
        self.use_max_length = use_max_length
        self.max_seq_length = max_seq_length
        self.nsamples = nsamples
        self.dataloader = []
        self.is_ready = False

    def get_prepared_dataloader(self):
        if not self.is_ready:
            self.prepare_dataloader()
        return self.dataloader

    def prepare_dataloader(self):
        self._select(exact=self.use_max_length)
        self.is_ready = True

    def obtain_first_n_samples(self, seed=0):
        self._select(exact=False, seed=seed)

    def obtain_first_n_samples_fulllength(self, seed=0):
        self._select(exact=True, seed=seed)

    def _select(self, exact, seed=0):
        self.dataloader.clear()
        random.seed(seed)
        limit = self.max_seq_length
        for batch in self.dataloader_original:
            if len(self.dataloader) == self.nsamples:
                logger.info(f"Successfully collect {self.nsamples} calibration samples.")
                break
            if isinstance(batch, dict) and "input_ids" not in batch:
                logger.warning("Please make sure your dict'like data contains key of 'input_ids'.")
                continue
            length = _seq_len_of(batch)
            if length > limit:
                i = random.randint(0, length - limit - 1)
                batch = _crop(batch, i, i + limit)
            elif exact and length < limit:
                continue  # too short for a full-length calibration sample
            elif isinstance(batch, tuple):
                batch = list(batch)
            self.dataloader.append(batch)
        if len(self.dataloader) < self.nsamples:
            logger.warning(f"Try to use {self.nsamples} data, but only {len(self.dataloader)} samples are available"
                           + (f" at fixed length {limit}." if exact else "."))


def get_example_inputs(model, dataloader):
    """First batch of `dataloader` on the model's device in the form `model(...)` takes it (examples utils.py:151-198):
    an (input, label) pair yields the input; a dict loses its "label" entry."""
    if dataloader is None:
        return None
    device = next(model.parameters()).device
    for batch in dataloader:
        if isinstance(batch, (list, tuple)) and len(batch) == 2:
            batch = batch[0]
        batch = move_to_device(batch, device)
        if isinstance(batch, dict):
            return {k: v for k, v in batch.items() if k != "label"}
        if isinstance(batch, (list, tuple)):
            return tuple(batch)
        return batch
    raise AssertionError("Please checkout the example_inputs format.")


def run_calibration(model, dataloader):
    """The examples' `run_fn`: one forward per calibration batch, on the model's device, without gradients."""
    device = next(model.parameters()).device
    with torch.no_grad():
        for batch in dataloader:
            batch = move_to_device(batch, device)
            if isinstance(batch, dict):
                model(**batch)
            elif isinstance(batch, (list, tuple)):
                model(*batch)
            else:
                model(batch)


# ------------------------------------------------------------------------------------------------
# evaluation
# ------------------------------------------------------------------------------------------------
def llm_benchmark(model, batch_size, input_length, warmup_iters=3, total_iters=20):
    """llm_utility.py:107-126: forward latency / throughput on an all-ones prompt.  Timed on the device (CUDA events on
    the current stream around the `total_iters - warmup_iters` timed forwards); returns the numbers it logs."""
    device = next(model.parameters()).device
    ids = torch.ones((batch_size, input_length), dtype=torch.long, device=device)
    logger.info("Batch size = {:d}".format(batch_size))
    logger.info("The length of input tokens = {:d}".format(input_length))
    timed = total_iters - warmup_iters
    assert timed > 0, "total_iters must exceed warmup_iters"
    with torch.no_grad():
        for _ in range(warmup_iters):
            model(ids)
        if device.type == "cuda":
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(device)
            start.record()
            for _ in range(timed):
                model(ids)
            end.record()
            torch.cuda.synchronize(device)
            seconds = start.elapsed_time(end) * 1e-3
        else:
            import time

            t0 = time.perf_counter()
            for _ in range(timed):
                model(ids)
            seconds = time.perf_counter() - t0
    latency = seconds / (timed * batch_size)
    throughput = (timed * batch_size) / seconds
    logger.info("Latency: {:.3f} ms".format(latency * 1e3))
    logger.info("Throughput: {:.3f} samples/sec".format(throughput))
    return {"latency_s": latency, "throughput_samples_per_s": throughput, "batch_size": batch_size,
            "input_length": input_length}


@torch.no_grad()
def evaluate_perplexity(model, token_ids, seq_len=2048, batch_size=1):
    """Perplexity of `model` over a 1-D token stream cut into `seq_len` windows (the GPTQ papers' wikitext protocol, what
    the reference's examples obtain through lm-eval's `wikitext` task): exp(mean next-token NLL)."""
    device = next(model.parameters()).device
    token_ids = token_ids.reshape(-1)
    n = token_ids.numel() // seq_len
    assert n > 0, "token stream shorter than one window"
    windows = token_ids[: n * seq_len].reshape(n, seq_len)
    nll, count = torch.zeros((), dtype=torch.float64, device=device), 0
    for s in range(0, n, batch_size):
        ids = windows[s:s + batch_size].to(device)
        logits = model(ids).logits.float()
        loss = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]), ids[:, 1:].reshape(-1),
                                                 reduction="sum")
        nll += loss.double()
        count += ids[:, 1:].numel()
    return float(torch.exp(nll / count))
