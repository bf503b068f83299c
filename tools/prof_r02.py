"""Tiny drivers for the round-2 ncu captures (tools/profile_round2.sh): run one kernel a few times on BASELINE shapes.
    python tools/prof_r02.py k2 | woq_tc <M> | w8a8 <M> | syrk"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_compressor_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")
what = sys.argv[1]
if what == "k2":
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    X = torch.randn(2 * C, C, device=DEV)
    H = (X.t() @ X) / C
    H[torch.arange(C), torch.arange(C)] += 0.01 * torch.diag(H).mean()
    for _ in range(2):
        ops.cholesky_inverse_upper(H)
elif what == "syrk":
    for C in (4096, 11008):
        X = torch.randn(16384, C, device=DEV, dtype=torch.float16)
        H = torch.zeros(C, C, device=DEV)
        for _ in range(3):
            ops.hessian_accumulate(X, H)
elif what == "woq_tc":
    M = int(sys.argv[2])
    for N, K in ((4096, 4096), (11008, 4096), (4096, 11008)):
        W = torch.randn(N, K, device=DEV) * 0.02
        r = ops.rtn_quant_pack(W, 4, 128, True)
        x = torch.randn(M, K, device=DEV, dtype=torch.float16)
        for _ in range(3):
            ops.woq_linear(x, r["qweight"], r["qzeros"], r["scales"], None, 4, 128, K, N, out_dtype=torch.float16)
elif what == "w8a8":
    from neural_compressor_b200.algorithms.smooth_quant import SQLinear

    M = int(sys.argv[2])
    for N, K in ((4096, 4096), (16384, 4096), (4096, 16384)):
        lin = torch.nn.Linear(K, N, bias=True).to(DEV).half()
        mod = SQLinear(lin, torch.ones(K, device=DEV), -torch.ones(K, device=DEV) * 3, torch.ones(K, device=DEV) * 3)
        x = torch.randn(M, K, device=DEV, dtype=torch.float16)
        for _ in range(3):
            mod(x)
torch.cuda.synchronize()
