"""BASELINE configs[4]: GPTQ INT4 g128 calibration of a Llama-3-70B-shaped model, LAYER-SHARDED over the GPUs of one
box (utils/sharded.py): rank r owns a contiguous range of decoder blocks, the owner broadcasts a block's bf16 weights
(1.7 GB) over NCCL/NVLink right before it is processed, calibration sequences are sharded over the ranks, raw Hessians
are reduced onto owner ranks that factorise them and broadcast the inverse factors, the column loop is row-sharded.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_70b.py --blocks 4 [--seqs 128]

Times `--blocks` decoder blocks (device events, max over ranks), reports seconds per block, the 80-block estimate and the
NCCL bandwidths of the block broadcast / Hessian reduce / factor broadcast against the measured NVLink-5 references of
B200_PROFILING.md (770 GB/s peer copy, 725 GB/s all-reduce bus bandwidth).  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HIDDEN, INTER, HEADS, KV, VOCAB, LAYERS = 8192, 28672, 64, 8, 128256, 80
SEQ = 2048


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=4)
    ap.add_argument("--seqs", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from transformers import LlamaConfig, LlamaForCausalLM

    import neural_compressor_b200.quantization as Q
    from neural_compressor_b200.algorithms.gptq import shard_range
    from neural_compressor_b200.utils import sharded

    n_layers = max(a.blocks + a.warmup, world)
    cfg = LlamaConfig(hidden_size=HIDDEN, intermediate_size=INTER, num_hidden_layers=n_layers, num_attention_heads=HEADS,
                      num_key_value_heads=KV, vocab_size=VOCAB, max_position_embeddings=8192, tie_word_embeddings=False,
                      torch_dtype=torch.bfloat16)

    def factory():
        old = torch.get_default_dtype()
        torch.set_default_dtype(torch.bfloat16)
        try:
            return LlamaForCausalLM(cfg)
        finally:
            torch.set_default_dtype(old)

    def init(mod, idx):  # random-init on the device (seeded per block: every world size builds the same model)
        g = torch.Generator(device=dev).manual_seed(1000 + idx)
        for name, p in mod.named_parameters():
            if p.dim() <= 1:
                p.data.fill_(1.0)
            else:
                p.data.normal_(0.0, 0.02, generator=g)

    t0 = time.perf_counter()
    model = sharded.build_layer_sharded(factory, "model.layers", rank, world, dev, init=init).eval()
    model.config.use_cache = False
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    owned = model._b200_shard["owned"]
    mem_model = torch.cuda.memory_allocated(dev) / 2**30

    qc = Q.GPTQConfig(bits=4, group_size=128, use_sym=True, act_order=False, block_size=128, percdamp=0.01)
    model = Q.prepare(model, qc)
    lo, hi = shard_range(a.seqs, rank, world)
    g = torch.Generator().manual_seed(1234)
    ids = [torch.randint(0, VOCAB, (1, SEQ), generator=g) for _ in range(a.seqs)][lo:hi]
    with torch.no_grad():
        for x in ids:
            model(x.to(dev))
    eng = model.quantizer.gptq_quantizer
    eng.remove_prepare_for_calibration()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_block = []
    with torch.no_grad():
        for b in range(a.warmup + a.blocks):
            barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            eng.quantize_block(b)
            e.record()
            barrier()
            ms = torch.tensor([s.elapsed_time(e)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if b >= a.warmup:
                per_block.append(ms.item() / 1e3)
            if rank == 0:
                print(f"[70b] block {b}: {ms.item() / 1e3:.3f} s", file=sys.stderr, flush=True)
    # phase breakdown of one more block when one is left, with syncs
    nccl = {}
    if world > 1:
        def bw(fn, nbytes, reps=3):
            fn()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(reps):
                fn()
            e.record()
            torch.cuda.synchronize()
            t = torch.tensor([s.elapsed_time(e) / reps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return round(nbytes / t.item() / 1e6, 1), round(t.item(), 3)

        blk = torch.empty(1711 * 2**20 // 2, dtype=torch.bfloat16, device=dev)      # one decoder block's weights
        nccl["block_broadcast_GBs"], nccl["block_broadcast_ms"] = bw(lambda: dist.broadcast(blk, src=0), blk.numel() * 2)
        del blk
        Hb = torch.empty(INTER, INTER, dtype=torch.float32, device=dev)              # down_proj Hessian / factor (3.3 GB)
        nccl["hessian_reduce_GBs"], nccl["hessian_reduce_ms"] = bw(lambda: dist.reduce(Hb, dst=0), Hb.numel() * 4)
        nccl["factor_broadcast_GBs"], nccl["factor_broadcast_ms"] = bw(lambda: dist.broadcast(Hb, src=0), Hb.numel() * 4)
        nccl["reference"] = "B200_PROFILING.md: peer copy 770 GB/s per direction, 8-rank all-reduce bus bandwidth 725 GB/s (nominal 900)"
        del Hb
    med = sorted(per_block)[len(per_block) // 2]
    out = dict(metric="Llama-3-70B GPTQ-INT4-g128 calibration, layer-sharded", n_gpus=world, blocks_timed=len(per_block),
               sec_per_block=[round(v, 3) for v in per_block], sec_per_block_median=round(med, 3),
               est_80_blocks_sec=round(med * LAYERS, 1), calib=f"{a.seqs} seqs x {SEQ} tokens (sharded {hi - lo} per rank)",
               model_build_s=round(t_build, 1), owned_blocks=list(owned), resident_model_GiB=round(mem_model, 1),
               peak_GiB=round(torch.cuda.max_memory_allocated(dev) / 2**30, 1), nccl=nccl,
               config=dict(workload="GPTQ INT4 g128 sym block_size=128, Llama-3-70B shapes (hidden 8192, inter 28672, GQA 64/8), "
                                    f"bf16 random-init, {n_layers} decoder blocks built (capacity partition over the ranks)"))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
