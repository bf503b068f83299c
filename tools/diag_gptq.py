"""Diagnostic: kernel fasterquant vs oracle on the same Hinv, per column block."""
import sys, torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from neural_compressor_b200 import ops
from oracle import woq_oracle as O

dev = "cuda:0"
g = torch.Generator().manual_seed(3)
for (N, C, gs, bs, sym) in [(128, 256, 32, 128, True), (128, 256, 32, 128, False), (48, 256, 32, 128, True),
                            (128, 256, 32, 2048, True), (128, 512, 128, 128, True), (64, 384, 32, 128, True)]:
    W = torch.randn(N, C, generator=g) * 0.05
    X = [torch.randn(1, 64, C, generator=g) for _ in range(4)]
    lay = O.GPTQLayerOracle(N, C, bits=4, sym=sym)
    for x in X:
        lay.add_batch(x)
    res = lay.fasterquant(W, bs, 0.01, gs)
    exp = O.GPTQLayerOracle.export_codes(res["Q"], res["scale"], res["zero"], gs, sym)
    if sym:
        exp = exp + 8
    r = ops.gptq_fasterquant(W.clone().to(dev), res["hinv"].contiguous().to(dev), None, bs, gs, 4, sym, False)
    codes = r["codes"].cpu().to(torch.int32)
    mism = (codes != exp)
    per_blk = [mism[:, i:i + 128].float().mean().item() for i in range(0, C, 128)]
    sd = (r["scale"].cpu() - res["scale"]).abs().max(0)[0]
    print(dict(N=N, C=C, g=gs, bs=bs, sym=sym), "mismatch per 128-col block", per_blk, "scale maxdiff per group", sd.tolist())
    print("   Q maxdiff", (r["Q"].cpu() - res["Q"]).abs().max().item())
