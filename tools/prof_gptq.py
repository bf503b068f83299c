"""Tiny driver for ncu: one GPTQ column loop (sub-block + lazy update kernels) on a 4096x4096 layer, and one RTN pack."""
import sys, torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from neural_compressor_b200 import ops
dev = torch.device("cuda:0")
N, C = 4096, 4096
X = torch.randn(8192, C, device=dev, dtype=torch.float16)
H = torch.zeros(C, C, device=dev)
ops.hessian_accumulate(X, H)
H, dead = ops.hessian_finalize(H, 4, 0.01)
Hinv = ops.cholesky_inverse_upper(H)
W = torch.randn(N, C, device=dev) * 0.02
r = ops.gptq_fasterquant(W.clone(), Hinv, dead, 128, 128, 4, True, False)
ops.rtn_quant_pack(W.half(), 4, 128, True)
torch.cuda.synchronize()
print("done", r["losses"].sum().item())
