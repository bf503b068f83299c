"""Round-2 micro-benchmarks (run under gpurun): K1 SYRK time + accuracy, K2 inverse factor time + accuracy,
W8A8 GEMM rates.  Prints one JSON line per measurement.  `python tools/r02_micro.py syrk|k2|w8a8|all`"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_compressor_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def syrk():
    for C in (4096, 11008):
        T = 16384
        g = torch.Generator().manual_seed(C)
        ch = torch.exp(torch.randn(C, generator=g) * 0.8)
        X = (torch.randn(T, C, generator=g) * ch).half().to(DEV)
        H = torch.zeros(C, C, device=DEV)
        ms = timed(lambda: ops.hessian_accumulate(X, H))
        H.zero_()
        ops.hessian_accumulate(X, H)
        # accuracy on a 512 x 512 corner and a far off-diagonal block vs fp64
        Xd = X[:, :512].double()
        ref = Xd.t() @ Xd
        err = float((H[:512, :512].double() - ref).abs().max() / ref.abs().max())
        Xe = X[:, C - 512:].double()
        ref2 = Xd.t() @ Xe
        err2 = float((H[:512, C - 512:].double() - ref2).abs().max() / ref.abs().max())
        fl = T * C * (C + 128)
        print(json.dumps(dict(kernel="syrk", C=C, T=T, seg_kb=os.environ.get("B200WOQ_SYRK_SEG_KB", "default"), ms=round(ms, 4),
                              tflops_sym=round(fl / ms / 1e9, 1), rel_err_diag_block=err, rel_err_offdiag_block=err2)), flush=True)


def k2():
    for C in [int(v) for v in os.environ.get("B200WOQ_MICRO_C", "4096,11008").split(",")]:
        g = torch.Generator().manual_seed(C)
        ch = torch.exp(torch.randn(C, generator=g) * 0.8)
        X = (torch.randn(2 * C, C, generator=g) * ch).to(DEV)
        H = (X.t() @ X) * (2.0 / (2 * C))
        H[torch.arange(C), torch.arange(C)] += 0.01 * torch.diag(H).mean()
        del X
        ms = timed(lambda: ops.cholesky_inverse_upper(H, check=False), reps=3, warm=1)
        U = ops.cholesky_inverse_upper(H)
        os.environ["B200WOQ_CHOLINV"] = "torch"
        ms_lib = timed(lambda: ops.cholesky_inverse_upper(H), reps=3, warm=1)
        U_lib = ops.cholesky_inverse_upper(H)
        os.environ["B200WOQ_CHOLINV"] = "b200"
        # residual in fp64: U^T U H = I
        Ud = U.double()
        R = (Ud.t() @ Ud) @ H.double()
        res = float((R - torch.eye(C, device=DEV, dtype=torch.float64)).abs().max())
        Ul = U_lib.double()
        Rl = (Ul.t() @ Ul) @ H.double()
        res_lib = float((Rl - torch.eye(C, device=DEV, dtype=torch.float64)).abs().max())
        print(json.dumps(dict(kernel="cholinv_upper", C=C, ms=round(ms, 3), ms_cusolver_ul=round(ms_lib, 3),
                              tflops=round((2 / 3) * C**3 / ms / 1e9, 2), residual_ours=res, residual_cusolver=res_lib,
                              rel_diff_vs_cusolver=float((U - U_lib).abs().max() / U_lib.abs().max()))), flush=True)
        del H, U, U_lib, Ud, R, Ul, Rl
        torch.cuda.empty_cache()


def w8a8():
    from neural_compressor_b200.algorithms.smooth_quant import SQLinear

    for (N, K) in ((4096, 4096), (16384, 4096), (4096, 16384)):
        lin = torch.nn.Linear(K, N, bias=True).to(DEV).half()
        mod = SQLinear(lin, torch.ones(K, device=DEV), -torch.ones(K, device=DEV) * 3, torch.ones(K, device=DEV) * 3)
        for M in (1, 16, 64, 2048):
            x = torch.randn(M, K, device=DEV, dtype=torch.float16)
            ms = timed(lambda: mod(x), reps=20)
            ms_fp16 = timed(lambda: lin(x), reps=20)
            print(json.dumps(dict(kernel="w8a8_linear", N=N, K=K, M=M, ms=round(ms, 4), ms_cublas_fp16=round(ms_fp16, 4),
                                  tops=round(2 * M * N * K / ms / 1e9, 1), weight_GBs=round(N * K / ms / 1e6, 1))), flush=True)


def woq_tc():
    """tcgen05 dequant-GEMM per layer under a CUDA graph, weights rotated through > L2."""
    import math

    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    shapes = ((4096, 4096), (12288, 4096), (22016, 4096), (4096, 11008))
    if os.environ.get("B200WOQ_MICRO_SHAPES"):
        shapes = [tuple(int(v) for v in t.split("x")) for t in os.environ["B200WOQ_MICRO_SHAPES"].split(",")]
    for (N, K) in shapes:
        copies = min(48, max(4, int(math.ceil(300e6 / (N * K / 2)))))
        packs = []
        for i in range(copies):
            W = torch.randn(N, K, device=DEV) * 0.02
            r = ops.rtn_quant_pack(W, 4, 128, True)
            packs.append((r["qweight"], r["qzeros"], r["scales"]))
            del W
        for M in [int(v) for v in os.environ.get("B200WOQ_MICRO_MS", "8,16,32,64,128").split(",")]:
            x = torch.randn(M, K, device=DEV, dtype=torch.float16)
            y = torch.empty(M, N, device=DEV, dtype=torch.float16)

            def run():
                for (qw, qz, sc) in packs:
                    ops.woq_linear(x, qw, qz, sc, None, 4, 128, K, N, out_dtype=torch.float16, flags=2, out=y)
            run()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run()
            ms = timed(g.replay, reps=10) / copies
            by = N * K // 2 + 2 * N * K // 128 + N * K // 256 + 2 * M * K + 2 * M * N
            print(json.dumps(dict(kernel="woq_linear(tc)", N=N, K=K, M=M, us=round(ms * 1e3, 2), GBs=round(by / ms / 1e6, 1),
                                  hbm_frac=round(by / ms / 1e6 / peak, 3), tflops=round(2 * M * N * K / ms / 1e9, 1))), flush=True)
        del packs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("syrk", "all"):
        syrk()
    if what in ("k2", "all"):
        k2()
    if what in ("w8a8", "all"):
        w8a8()
    if what in ("woq_tc", "all"):
        woq_tc()
