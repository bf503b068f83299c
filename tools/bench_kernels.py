"""Per-kernel timings on the B200 (development tool; bench.py is the contract benchmark).

    python tools/bench_kernels.py [gemv] [hessian] [gptq] [rtn] > gpurun_out/kernels.json
"""
import json
import math
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from neural_compressor_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")
PEAK = json.load(open(__file__.rsplit("/tools/", 1)[0] + "/MEASURED_PEAKS.json")) if True else {}


def time_cuda(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters  # ms


def make_packed(N, K, g=128, bits=4, sym=True, seed=0):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    W = (torch.randn(N, K, generator=gen) * 0.02).to(DEV)
    r = ops.rtn_quant_pack(W, bits, g, sym)
    return r["qweight"], r["qzeros"], r["scales"]


def gemv_bytes(M, N, K, g=128, bits=4, xb=2, yb=2):
    return N * K * bits // 8 + 2 * N * K // g + N * K // (2 * g) * bits // 4 + xb * M * K + yb * M * N


def bench_gemvs(out):
    bench_gemv(out, stream_only=True)


def bench_gemv(out, stream_only=False):
    shapes = [("qkv/o 4096x4096", 4096, 4096), ("gate/up 11008x4096", 11008, 4096), ("down 4096x11008", 4096, 11008)]
    import os
    if os.environ.get("B200WOQ_BENCH_SHAPES"):  # e.g. "4096x128,22016x4096" (NxK)
        shapes = [(f"custom/{t}", int(t.split("x")[0]), int(t.split("x")[1]))
                  for t in os.environ["B200WOQ_BENCH_SHAPES"].split(",")]
    res = []
    for name, N, K in shapes:
        copies = min(64, max(4, int(math.ceil(400e6 / (N * K / 2)))))  # rotate over > L2-size worth of weights
        packs = [make_packed(N, K, seed=i) for i in range(copies)]
        layouts = [ops.build_stream_layout(qw, qz, sc, 4, 128, K, N) for (qw, qz, sc) in packs]
        for M in [int(v) for v in os.environ.get("B200WOQ_BENCH_MS", "1,2,4").split(",")]:
            x = torch.randn(M, K, device=DEV, dtype=torch.float16)
            y = torch.empty(M, N, device=DEV, dtype=torch.float16)
            for flags in ((2,) if stream_only else (0, 2)):
                def run_s():
                    for lay in layouts:
                        ops.woq_linear_stream(x, lay, None, 4, 128, K, N, out_dtype=torch.float16, flags=flags, out=y)
                run_s()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    run_s()
                ms = time_cuda(g.replay, iters=20) / copies
                by = gemv_bytes(M, N, K)
                res.append(dict(shape=name, M=M, pdl=bool(flags & 2), impl="stream", us=ms * 1e3, GBs=by / ms / 1e6,
                                frac_hbm=by / ms / 1e6 / PEAK["hbm_gbs"], tflops=2 * M * N * K / ms / 1e9))
                print(res[-1], file=sys.stderr)
        del layouts
        if stream_only:
            continue
        for M in (1, 2, 4, 8, 16, 32, 64):
            x = torch.randn(M, K, device=DEV, dtype=torch.float16)
            y = torch.empty(M, N, device=DEV, dtype=torch.float16)
            for flags in (0, 2):
                def run():
                    for (qw, qz, sc) in packs:
                        ops.woq_linear(x, qw, qz, sc, None, 4, 128, K, N, out_dtype=torch.float16, flags=flags, out=y)
                run()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    run()
                ms = time_cuda(g.replay, iters=20) / copies
                by = gemv_bytes(M, N, K)
                res.append(dict(shape=name, M=M, pdl=bool(flags & 2), us=ms * 1e3, GBs=by / ms / 1e6,
                                frac_hbm=by / ms / 1e6 / PEAK["hbm_gbs"], tflops=2 * M * N * K / ms / 1e9))
                print(res[-1], file=sys.stderr)
    out["gemv"] = res


def bench_hessian(out):
    res = []
    for C in (4096, 11008):
        for dt in (torch.float16, torch.bfloat16):
            T = 2048 * 8
            X = torch.randn(T, C, device=DEV, dtype=dt)
            H = torch.zeros(C, C, device=DEV)
            import os
            os.environ["B200WOQ_HESSIAN_IMPL"] = "tc"
            ms_tc = time_cuda(lambda: ops.hessian_accumulate(X, H), iters=5, warmup=2)
            os.environ["B200WOQ_HESSIAN_IMPL"] = "mma"
            ms = time_cuda(lambda: ops.hessian_accumulate(X, H), iters=5, warmup=2)
            res.append(dict(C=C, dtype=str(dt), T=T, tc_ms=ms_tc, tc_tflops_sym_half=T * C * (C + 128) / ms_tc / 1e9,
                            tc_tflops_full_equiv=2 * T * C * C / ms_tc / 1e9))
            print(res[-1], file=sys.stderr)
            fl_full = 2 * T * C * C
            nt = math.ceil(C / 128)
            fl_done = 2 * T * 128 * 128 * nt * (nt + 1) / 2
            res.append(dict(C=C, dtype=str(dt), T=T, ms=ms, tflops_algorithmic_full=fl_full / ms / 1e9,
                            tflops_executed=fl_done / ms / 1e9))
            print(res[-1], file=sys.stderr)
            ms2 = time_cuda(lambda: torch.matmul(X.t(), X), iters=5, warmup=2)
            res.append(dict(C=C, dtype=str(dt), cublas_full_gemm_ms=ms2, tflops=fl_full / ms2 / 1e9))
            print(res[-1], file=sys.stderr)
            del X, H
    out["hessian"] = res


def bench_gptq(out):
    res = []
    for N, C in ((4096, 4096), (11008, 4096), (4096, 11008)):
        X = torch.randn(4096, C, device=DEV, dtype=torch.float16)
        H = torch.zeros(C, C, device=DEV)
        ops.hessian_accumulate(X, H)
        H, dead = ops.hessian_finalize(H, 2, 0.01)
        t_ch = time_cuda(lambda: ops.cholesky_inverse_upper(H), iters=2, warmup=1)
        Hinv = ops.cholesky_inverse_upper(H)
        W0 = torch.randn(N, C, device=DEV) * 0.02
        def fq():
            ops.gptq_fasterquant(W0.clone(), Hinv, dead, 128, 128, 4, True, False)
        t_fq = time_cuda(fq, iters=2, warmup=1)
        r = ops.gptq_fasterquant(W0.clone(), Hinv, dead, 128, 128, 4, True, False)
        t_pack = time_cuda(lambda: (ops.pack_codes(r["codes"], 4), ops.pack_params(r["scale"], None, 4)), iters=5)
        res.append(dict(N=N, C=C, cholesky_chain_ms=t_ch, fasterquant_ms=t_fq, pack_ms=t_pack,
                        loss=r["losses"].sum().item()))
        print(res[-1], file=sys.stderr)
    out["gptq"] = res


def bench_rtn(out):
    res = []
    for N, K in ((4096, 4096), (11008, 4096)):
        for dt in (torch.float32, torch.float16):
            W = (torch.randn(N, K, device=DEV) * 0.02).to(dt)
            ms = time_cuda(lambda: ops.rtn_quant_pack(W, 4, 128, True), iters=10)
            by = N * K * W.element_size() + N * K // 2
            res.append(dict(N=N, K=K, dtype=str(dt), ms=ms, GBs_algorithmic=by / ms / 1e6))
            print(res[-1], file=sys.stderr)
    out["rtn"] = res


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemv", "hessian", "gptq", "rtn"]
    out = dict(peaks=PEAK)
    for w in which:
        globals()["bench_" + w](out)
    print(json.dumps(out, indent=1))
