"""Where does the first decoder block of a fresh process spend its time?  (block 0 took 60 s on a
fresh box.)  Times the FIRST call of every library path a calibration step touches, in a fresh process, and the second
call right after it.  Run under gpurun; prints one JSON line."""
import json
import sys
import time

import torch

t_import = time.perf_counter()
dev = torch.device("cuda:0")
res = {}


def timed(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    res[name] = dict(first_s=round(t1 - t0, 3), second_s=round(time.perf_counter() - t1, 4))
    print(name, res[name], file=sys.stderr, flush=True)


timed("cuda_init_alloc", lambda: torch.zeros(1 << 20, device=dev))
a = torch.randn(8192, 4096, device=dev, dtype=torch.float16)
b = torch.randn(4096, 4096, device=dev, dtype=torch.float16)
timed("cublas_fp16_gemm", lambda: a @ b)
timed("elementwise_silu_mul", lambda: torch.nn.functional.silu(a) * a)
q = torch.randn(8, 32, 2048, 128, device=dev, dtype=torch.float16)
from torch.nn.attention import SDPBackend, sdpa_kernel

for name, be in (("sdpa_flash", [SDPBackend.FLASH_ATTENTION]), ("sdpa_efficient", [SDPBackend.EFFICIENT_ATTENTION]),
                 ("sdpa_cudnn", [SDPBackend.CUDNN_ATTENTION])):
    def f(be=be):
        with sdpa_kernel(be):
            torch.nn.functional.scaled_dot_product_attention(q, q, q, is_causal=True)
    try:
        timed(name, f)
    except Exception as ex:
        res[name] = dict(error=str(ex)[:200])
timed("sdpa_default", lambda: torch.nn.functional.scaled_dot_product_attention(q, q, q, is_causal=True))
mask = torch.ones(2048, 2048, device=dev, dtype=torch.bool).tril()[None, None]
timed("sdpa_default_with_mask", lambda: torch.nn.functional.scaled_dot_product_attention(q, q, q, attn_mask=mask))
h = torch.randn(4096, 4096, device=dev)
h = h @ h.t() + 4096 * torch.eye(4096, device=dev)
timed("cusolver_cholesky_ex", lambda: torch.linalg.cholesky_ex(h, check_errors=False))
L = torch.linalg.cholesky(h)
timed("cublas_trsm", lambda: torch.linalg.solve_triangular(L, torch.eye(4096, device=dev), upper=False))
sys.path.insert(0, ".")
from neural_compressor_b200 import ops

x = torch.randn(8, 2048, 4096, device=dev, dtype=torch.float16)
H = torch.zeros(4096, 4096, device=dev)
timed("b200woq_hessian_accumulate", lambda: ops.hessian_accumulate(x, H))
res["total_s"] = round(time.perf_counter() - t_import, 2)
print(json.dumps(res))
