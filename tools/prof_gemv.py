"""Tiny driver for ncu: a few dequant-GEMV launches on Llama-2-7B shapes (rotating weight copies > L2)."""
import sys, torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from neural_compressor_b200 import ops
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
use_stream = len(sys.argv) > 3 and sys.argv[3] == "stream"
for (N, K) in ((4096, 4096), (11008, 4096), (4096, 11008)):
    packs = []
    for i in range(6 if N * K > 2e7 else 16):
        W = (torch.randn(N, K, device=dev) * 0.02)
        r = ops.rtn_quant_pack(W, 4, 128, True)
        packs.append((r["qweight"], r["qzeros"], r["scales"],
                      ops.build_stream_layout(r["qweight"], r["qzeros"], r["scales"], 4, 128, K, N)))
    x = torch.randn(M, K, device=dev, dtype=torch.float16)
    y = torch.empty(M, N, device=dev, dtype=torch.float16)
    for rep in range(2):
        for (qw, qz, sc, lay) in packs:
            if use_stream:
                ops.woq_linear_stream(x, lay, None, 4, 128, K, N, out_dtype=torch.float16, flags=flags, out=y)
            else:
                ops.woq_linear(x, qw, qz, sc, None, 4, 128, K, N, out_dtype=torch.float16, flags=flags, out=y)
    torch.cuda.synchronize()
print("done")
