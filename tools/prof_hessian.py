"""Tiny driver for ncu: Hessian SYRK launches at the shapes bench.py runs (8 sequences x 2048 tokens per launch)."""
import sys, torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from neural_compressor_b200 import ops
dev = torch.device("cuda:0")
for C in (4096, 11008):
    X = torch.randn(int(sys.argv[1]) if len(sys.argv) > 1 else 16384, C, device=dev, dtype=torch.float16)
    H = torch.zeros(C, C, device=dev)
    for _ in range(3):
        ops.hessian_accumulate(X, H)
    torch.cuda.synchronize()
print("done")
