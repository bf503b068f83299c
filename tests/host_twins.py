"""CPU twins of the device kernels the GPTQ engine calls (ops.*), built on the oracle.  TEST INFRASTRUCTURE: they let the
host logic of algorithms/gptq.py run on the CPU -- single process (tests/test_gptq_hostflow_cpu.py) and world-size-2 gloo
(tests/test_distributed_cpu.py) -- so that what is compared with the live reference's fixtures is the wiring, not the
kernels (those are compared on the B200)."""
import math

import torch

from oracle import woq_oracle as O


def gptq_twins(running_mean: bool):
    """name -> function.  `running_mean=True`: the Hessian twin follows the reference's running-mean update
    (gptq.py:1111-1141) bit for bit (single rank, one sample per call); False: raw sums + 2/n at finalize, the closed
    form the product uses and the only one that can be reduced across ranks."""
    running = {}

    def hessian_accumulate(X, Hsum):
        x = X.unsqueeze(0) if X.dim() == 2 else X
        b = x.shape[0]
        x2 = x.reshape(-1, x.shape[-1]).t().float()
        if not running_mean:
            Hsum += x2.matmul(x2.t())
            return
        st = running.setdefault(id(Hsum), [torch.zeros_like(Hsum), 0, Hsum])
        st[0] *= st[1] / (st[1] + b)
        st[1] += b
        x2 = math.sqrt(2 / st[1]) * x2
        st[0] += x2.matmul(x2.t())

    def hessian_finalize(Hsum, nsamples, percdamp):
        H = running[id(Hsum)][0].clone() if running_mean else Hsum * (2.0 / nsamples)
        dead = torch.diag(H) == 0
        H[dead, dead] = 1
        idx = torch.arange(H.shape[0])
        H[idx, idx] += percdamp * torch.mean(torch.diag(H))
        return H, dead.to(torch.uint8)

    def cholesky_inverse_upper(H, info=None, check=True):
        if info is not None:
            info.zero_()
        # contiguous like the kernel's output: torch.linalg.cholesky(upper=True) hands back a transposed view, and a
        # collective ships raw storage order
        return O.GPTQLayerOracle.cholesky_inverse_upper(H).contiguous()

    def gptq_fasterquant(W, Hinv, dead_mask, blocksize=128, groupsize=-1, bits=4, sym=False, mse=False, want_q=True,
                         double_quant=None):
        assert double_quant is None
        N, C = W.shape
        W = W.clone()
        if dead_mask is not None:
            W[:, dead_mask.bool()] = 0
        r = O.GPTQLayerOracle(N, C, bits=bits, sym=sym, mse=mse).fasterquant(W, blocksize=blocksize, groupsize=groupsize, hinv=Hinv)
        g = C if groupsize <= 0 else groupsize
        idx = torch.arange(C) // g
        codes = torch.clamp(torch.round(r["Q"] / r["scale"][:, idx]) + r["zero"][:, idx], 0, 2**bits - 1).to(torch.uint8)
        return dict(codes=codes, Q=r["Q"] if want_q else None, scale=r["scale"], zero=r["zero"], losses=r["losses"].sum(1))

    def gptq_rebuild_q(codes, scale, zero, groupsize):
        C = codes.shape[1]
        idx = torch.arange(C) // (C if groupsize <= 0 else groupsize)
        return scale[:, idx] * (codes.float() - zero[:, idx])

    def pack_codes(codes, bits):
        return O.pack_optimum(codes.float(), torch.ones(codes.shape[0], 1), torch.zeros(codes.shape[0], 1), bits, codes.shape[1])[0]

    def pack_params(scale, zp, bits):
        _, qzeros, scales16 = O.pack_optimum(torch.zeros(scale.shape[0], 1), scale, zp, bits, 1)
        return scales16, qzeros

    def dequantize(qweight, qzeros, scales, bits, group_size, in_features, out_features, g_idx=None):
        return O.recover_fp16(qweight, qzeros, scales, bits, group_size, in_features, out_features, g_idx)

    return dict(hessian_accumulate=hessian_accumulate, hessian_finalize=hessian_finalize,
                cholesky_inverse_upper=cholesky_inverse_upper, gptq_fasterquant=gptq_fasterquant,
                gptq_rebuild_q=gptq_rebuild_q, pack_codes=pack_codes, pack_params=pack_params, dequantize=dequantize)


def install_gptq_twins(running_mean: bool, setter=setattr):
    """Patch neural_compressor_b200.ops and pin the engine to the CPU.  `setter(obj, name, value)` = monkeypatch.setattr in
    a pytest process, plain setattr in a spawned worker."""
    from neural_compressor_b200 import ops
    from neural_compressor_b200.algorithms import gptq as G

    for name, fn in gptq_twins(running_mean).items():
        setter(ops, name, fn)
    setter(G, "current_device", lambda: torch.device("cpu"))
