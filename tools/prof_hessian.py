"""Tiny driver for ncu: Hessian SYRK launches at Llama-2-7B shapes."""
import sys, torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from neural_compressor_b200 import ops
dev = torch.device("cuda:0")
for C in (4096, 11008):
    X = torch.randn(2048, C, device=dev, dtype=torch.float16)
    H = torch.zeros(C, C, device=dev)
    for _ in range(3):
        ops.hessian_accumulate(X, H)
    torch.cuda.synchronize()
print("done")
