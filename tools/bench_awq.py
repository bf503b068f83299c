"""AWQ calibration time on Llama-2-7B-shaped decoder blocks (BASELINE configs[2]: 128 x 512 calibration tokens,
bits 4, group 128, sym).  Quantises a `--layers`-block model through the public API and reports seconds per block.

    python tools/bench_awq.py --layers 2 --samples 128 > gpurun_out/awq.json
"""
import argparse
import json
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--samples", type=int, default=128)
    ap.add_argument("--seq", type=int, default=512)
    args = ap.parse_args()
    from transformers import LlamaConfig, LlamaForCausalLM

    import neural_compressor_b200.quantization as q
    from neural_compressor_b200 import _lib

    dev = torch.device("cuda:0")
    cfg = LlamaConfig(hidden_size=4096, intermediate_size=11008, num_hidden_layers=args.layers, num_attention_heads=32,
                      num_key_value_heads=32, vocab_size=32000, max_position_embeddings=4096, tie_word_embeddings=False)
    torch.manual_seed(0)
    with torch.device(dev):
        model = LlamaForCausalLM(cfg).half().eval()
    model.config.use_cache = False
    g = torch.Generator().manual_seed(1234)
    ids = [torch.randint(0, 32000, (1, args.seq), generator=g).to(dev) for _ in range(args.samples)]

    def run_fn(m):
        for x in ids:
            m(x)

    lib = _lib.load()
    torch.cuda.synchronize()
    n0 = lib.b200woq_launch_count()
    t0 = time.perf_counter()
    model = q.quantize(model, q.AWQConfig(bits=4, group_size=128, use_sym=True), run_fn=run_fn, example_inputs=ids[0])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = dict(metric="AWQ INT4 g128 calibration, Llama-2-7B-shaped blocks", layers=args.layers, samples=args.samples,
               seq=args.seq, seconds=round(dt, 2), seconds_per_block=round(dt / args.layers, 2),
               est_full_model_sec=round(dt / args.layers * 32, 1), gpu_launches=lib.b200woq_launch_count() - n0,
               peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 2))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
