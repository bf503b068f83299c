"""Layer-sharded models for multi-GPU calibration (BASELINE configs[4]: Llama-3-70B on 8 x B200).

The reference's only placement mechanism is moving one transformer block at a time to the device and back
(`transformer_block.to(device)` / `.cpu()`, weight_only/gptq.py:617,766) with the whole model in host RAM.  On a
B200 node the model lives in HBM instead, partitioned by LAYERS: rank r owns the decoder blocks
`block_range(r)`; every other block exists on that rank only as a `meta` skeleton (shapes, no storage).  During
calibration (algorithms/gptq.py) the owner broadcasts a block's weights over NCCL/NVLink right before the block is
processed (1.7 GB per Llama-3-70B block), all ranks run the data-parallel Hessian / row-sharded column loop on it, and
the non-owners drop it again -- every GPU is busy in every phase, no GPU ever holds more than its shard plus one block.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


def block_range(rank: int, world: int, n_blocks: int):
    """Contiguous capacity partition: rank r owns blocks [lo, hi)."""
    return rank * n_blocks // world, (rank + 1) * n_blocks // world


def block_owner(idx: int, world: int, n_blocks: int) -> int:
    for r in range(world):
        lo, hi = block_range(r, world, n_blocks)
        if lo <= idx < hi:
            return r
    raise IndexError(idx)


def is_remote(module: torch.nn.Module) -> bool:
    """True when the module is a storage-less skeleton on this rank."""
    return any(p.is_meta for p in module.parameters()) or any(b.is_meta for b in module.buffers())


def _random_fill(module: torch.nn.Module, seed: int, std: float = 0.02):
    g = torch.Generator(device="cpu").manual_seed(seed)
    for name, p in module.named_parameters():
        if p.dim() <= 1 and ("norm" in name.lower() or name.endswith("layernorm.weight")):
            p.data.fill_(1.0)
        elif p.dim() <= 1:
            p.data.zero_()
        else:
            # generated on the host in slices so a 70B-scale matrix never needs an fp32 twin on the device
            flat = p.data.view(-1)
            step = 1 << 24
            for i in range(0, flat.numel(), step):
                n = min(step, flat.numel() - i)
                flat[i:i + n].copy_((torch.randn(n, generator=g) * std).to(p.dtype))


def build_layer_sharded(factory: Callable[[], torch.nn.Module], blocks_attr: str, rank: int, world: int, device,
                        seed: int = 0, init: Optional[Callable[[torch.nn.Module, int], None]] = None):
    """Build `factory()` on the meta device, then materialise on `device` everything outside the transformer stack plus
    the blocks this rank owns.  `blocks_attr` is the dotted path of the nn.ModuleList (e.g. "model.layers").  Weights are
    random-init (seeded per block, so every world size builds the same model) unless `init(module, block_idx)` fills
    them (block_idx = -1 for the non-block modules)."""
    with torch.device("meta"):
        model = factory()
    blocks = model
    for part in blocks_attr.split("."):
        blocks = getattr(blocks, part)
    n = len(blocks)
    lo, hi = block_range(rank, world, n)
    block_ids = {id(b) for b in blocks}

    def materialise(mod, idx, path):
        mod.to_empty(device=device)
        mod._b200_path = path   # qualified name, for checkpoint-backed `init` callables (checkpoint_init)
        (init or (lambda m, i: _random_fill(m, seed * 100003 + i + 1)))(mod, idx)

    # non-block modules (embeddings, final norm, head, rotary tables): replicated
    for name, child in list(model.named_children()):
        _materialise_outside(child, name, blocks_attr, block_ids, materialise)
    for i in range(lo, hi):
        materialise(blocks[i], i, f"{blocks_attr}.{i}")
    model._b200_shard = dict(rank=rank, world=world, n_blocks=n, blocks_attr=blocks_attr, owned=(lo, hi))
    return model


def checkpoint_init(model_path: str, tied: Optional[dict] = None):
    """`init` callable for `build_layer_sharded` that fills a materialised module from a HuggingFace safetensors
    checkpoint directory (single file or sharded with `model.safetensors.index.json`), reading only the tensors of that
    module -- each rank touches its own blocks plus the replicated embeddings / head, never the whole checkpoint.
    `tied` maps a parameter absent from the file to the one it is tied to (default: lm_head.weight ->
    model.embed_tokens.weight, the `tie_word_embeddings` case)."""
    import json
    import os

    from safetensors import safe_open

    tied = {"lm_head.weight": "model.embed_tokens.weight"} if tied is None else tied
    index_file = os.path.join(model_path, "model.safetensors.index.json")
    if os.path.exists(index_file):
        with open(index_file) as f:
            where = json.load(f)["weight_map"]
    else:
        single = os.path.join(model_path, "model.safetensors")
        with safe_open(single, framework="pt") as f:
            where = {k: "model.safetensors" for k in f.keys()}
    handles = {}

    def read(name):
        if name not in where:
            if name in tied and tied[name] in where:
                name = tied[name]
            else:
                raise KeyError(f"{name} is not in the checkpoint at {model_path}")
        fn = where[name]
        if fn not in handles:
            handles[fn] = safe_open(os.path.join(model_path, fn), framework="pt")
        return handles[fn].get_tensor(name)

    def init(mod, idx):
        prefix = mod._b200_path
        for name, p in mod.named_parameters():
            p.data.copy_(read(f"{prefix}.{name}"))
        persistent = {n for n, _ in mod.named_buffers()} - {
            (mn + "." if mn else "") + b for mn, m in mod.named_modules() for b in getattr(m, "_non_persistent_buffers_set", ())}
        for name, b in mod.named_buffers():
            if name in persistent:
                b.data.copy_(read(f"{prefix}.{name}"))

    return init


def _materialise_outside(module, path, blocks_attr, block_ids, materialise):
    """Materialise every sub-module that is not (inside) the transformer stack."""
    if path == blocks_attr:
        return
    if blocks_attr.startswith(path + "."):
        # the stack lives below this module: recurse, and materialise this module's OWN tensors
        for name, child in list(module.named_children()):
            _materialise_outside(child, path + "." + name, blocks_attr, block_ids, materialise)
        own = [p for p in module._parameters.values() if p is not None] + [b for b in module._buffers.values() if b is not None]
        if own:
            raise NotImplementedError(f"{path} holds tensors next to the transformer stack")
        return
    materialise(module, -1, path)
    _restore_buffers(module)


def _restore_buffers(module):
    """Non-persistent buffers computed in __init__ (rotary inv_freq tables) are lost by to_empty(): re-instantiate the
    modules that hold them (`type(m)(m.config, device=...)`, the HF rotary-embedding signature) and take their buffers;
    anything else with a non-persistent buffer fails loudly instead of running on garbage."""
    for m in module.modules():
        names = getattr(m, "_non_persistent_buffers_set", set())
        if not names:
            continue
        dev = next((b.device for b in m.buffers(recurse=False)), None)
        try:
            fresh = type(m)(m.config, device=dev)
        except Exception as ex:  # pragma: no cover
            raise NotImplementedError(f"cannot rebuild the non-persistent buffers {sorted(names)} of {type(m).__name__}") from ex
        for n in names:
            m._buffers[n] = fresh._buffers[n].to(dev)
        for attr in ("attention_scaling",):
            if hasattr(fresh, attr):
                setattr(m, attr, getattr(fresh, attr))


def fetch_block(block: torch.nn.Module, owner: int, device) -> int:
    """Owner -> everyone: the tensors of one block (one NCCL broadcast per tensor; the receivers materialise the
    skeleton first).  Returns the bytes moved per rank."""
    import torch.distributed as dist

    if is_remote(block):
        block.to_empty(device=device)
    else:
        block.to(device)
    nbytes = 0
    for t in list(block.parameters()) + list(block.buffers()):
        dist.broadcast(t.data, src=owner)
        nbytes += t.numel() * t.element_size()
    return nbytes


def release_block(block: torch.nn.Module):
    """Turn a fetched block back into a storage-less skeleton (non-owners, after the block has been processed)."""
    return block.to("meta")
