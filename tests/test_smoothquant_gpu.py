"""SmoothQuant W8A8 (BASELINE configs[3]) on the OPT-6.7B layer shapes: q/k/v/out [4096,4096], fc1 [16384,4096],
fc2 [4096,16384]; decode batches M = 1..64 and a 2048-token prefill.

INT8 GEMM PARITY UNPINNED (IPEX absent): the reference's INT8 GEMM lives in IPEX/oneDNN outside the reference tree (SURVEY
§8c), so the checker is the oracle's restatement of the reference's own pure-torch W8A8 QDQ simulation
(`oracle.sq_w8a8_linear`; smooth_quant/utility.py:652-755, 2559-2662, 2707-2729) -- whose quantisation parameters and
helper functions ARE pinned against the live reference (tests/test_smoothquant_transform_cpu.py: `cal_scale`,
`quant_dequant_w_v1/x_v1`, `SQLinearWrapper._calculate_qparams`, `TorchSmoothQuant.transform` with IPEX stubbed).  Integer parts (weight codes, row sums) must be bit-exact; the output is compared with the
simulation (fp32 summation order differs: 1e-3) and, on a sample of rows, with exact integer arithmetic in fp64 (1e-5)."""
import pytest
import torch

from oracle import woq_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(N, K, M, seed, dtype=torch.float16, bias=True):
    g = torch.Generator().manual_seed(seed)
    W = (torch.randn(N, K, generator=g) * 0.02).to(dtype)
    b = (torch.randn(N, generator=g) * 0.1).to(dtype) if bias else None
    ch = torch.exp(torch.randn(K, generator=g) * 0.6)
    ch[torch.randint(0, K, (8,), generator=g)] *= 15.0                     # activation outlier channels
    x = (torch.randn(M, K, generator=g) * ch).to(dtype)
    calib = torch.randn(512, K, generator=g) * ch
    act_max, act_min = calib.max(0)[0], calib.min(0)[0]
    in_abs = torch.maximum(act_max.abs(), act_min.abs())
    smooth = O.sq_cal_scale(in_abs, [W.float()], 0.5)
    return W, b, x, smooth, act_min, act_max


@pytest.mark.parametrize("N,K", [(4096, 4096), (16384, 4096), (4096, 16384), (384, 200)])
def test_sq_weight_quant_bit_exact(N, K):
    from neural_compressor_b200 import ops

    W, _, _, smooth, _, _ = _case(N, K, 1, N + K)
    r = ops.sq_smooth_quant_weight(W.to(DEV), smooth.to(DEV))
    _, q_w, s_w = O.sq_qdq_weight_per_channel(W.float() * smooth.view(1, -1), 8)
    assert torch.equal(r["qweight"][:, :K].cpu().float(), q_w)
    assert int(r["qweight"][:, K:].abs().sum()) == 0
    assert torch.equal(r["w_scale"].cpu(), s_w.flatten())
    assert torch.equal(r["wsum"].cpu().long(), q_w.sum(1).long())


@pytest.mark.parametrize("N,K", [(4096, 4096), (16384, 4096), (4096, 16384), (384, 200)])
@pytest.mark.parametrize("M", [1, 4, 16, 33, 64, 2048])
def test_w8a8_linear_vs_qdq_simulation(N, K, M, parity_log):
    from neural_compressor_b200.algorithms.smooth_quant import SQLinear

    if M == 2048 and N * K > 4096 * 4096 * 2:
        M = 1024  # keep the host-side fp32 checker GEMM in seconds
    W, b, x, smooth, act_min, act_max = _case(N, K, M, N + K + M)
    lin = torch.nn.Linear(K, N, bias=True).to(DEV).half()
    with torch.no_grad():
        lin.weight.copy_(W)
        lin.bias.copy_(b)
    mod = SQLinear(lin, smooth.to(DEV), act_min.to(DEV), act_max.to(DEV))
    y = mod(x.to(DEV))
    assert y.dtype == torch.float16 and tuple(y.shape) == (M, N)
    own = O.sq_w8a8_linear(x, W, smooth, act_min, act_max, b)
    # activation qparams: equal to the host's up to the last bit of a device reciprocal
    assert abs(float(mod.x_scale) - float(own["s_x"])) <= 2e-7 * float(own["s_x"]) and float(mod.x_zp) == float(own["zp_x"])
    ref = O.sq_w8a8_linear(x, W, smooth, act_min, act_max, b, qparams=(mod.input_scale, mod.x_scale, mod.x_zp))
    rel = float((y.float().cpu() - ref["y"]).norm() / ref["y"].norm())
    # exact integer arithmetic on a sample of rows / columns (fp64)
    rows = torch.arange(0, M, max(1, M // 4))[:4]
    acc = ref["q_x"][rows].double() @ ref["q_w"].double().t()
    exact = (acc - float(ref["zp_x"]) * ref["q_w"].double().sum(1)) * float(ref["s_x"]) * ref["s_w"].double().flatten() + b.double()
    rel_exact = float((y.float().cpu()[rows].double() - exact).norm() / exact.norm())
    parity_log(f"w8a8/{N}x{K}/M{M}", dict(rel_vs_qdq_sim=rel, rel_vs_exact_int=rel_exact))
    assert rel < 1e-3, rel            # fp16 output rounding (4.9e-4 rms bound) + fp32 summation order of the simulation
    assert rel_exact < 1e-3, rel_exact
    # the module's own QDQ evaluation (torch ops on the device) agrees as well, and the kernel is deterministic
    assert float((y.float() - mod.forward_qdq(x.to(DEV)).float()).norm() / y.float().norm()) < 2e-3
    assert torch.equal(y, mod(x.to(DEV)))


def test_w8a8_fp32_output_is_integer_exact():
    """fp32 in / fp32 out: the only roundings are the dequant multiply-add, so the result matches exact integer
    arithmetic to fp32 precision."""
    from neural_compressor_b200.algorithms.smooth_quant import SQLinear

    N, K, M = 512, 1024, 7
    W, b, x, smooth, act_min, act_max = _case(N, K, M, 99, dtype=torch.float32)
    lin = torch.nn.Linear(K, N, bias=True).to(DEV)
    with torch.no_grad():
        lin.weight.copy_(W)
        lin.bias.copy_(b)
    mod = SQLinear(lin, smooth.to(DEV), act_min.to(DEV), act_max.to(DEV))
    y = mod(x.to(DEV)).cpu()
    ref = O.sq_w8a8_linear(x, W, smooth, act_min, act_max, b, qparams=(mod.input_scale, mod.x_scale, mod.x_zp))
    acc = ref["q_x"].double() @ ref["q_w"].double().t()
    exact = (acc - float(ref["zp_x"]) * ref["q_w"].double().sum(1)) * float(ref["s_x"]) * ref["s_w"].double().flatten() + b.double()
    assert float((y.double() - exact).abs().max() / exact.abs().max()) < 1e-6


def test_smoothquant_api_tiny_model(golden_e2e):
    """quantize(model, SmoothQuantConfig, run_fn, example_inputs) through the mirrored public API."""
    import neural_compressor_b200.quantization as Q
    from neural_compressor_b200.algorithms.smooth_quant import SQLinear
    from tests.test_api_gpu import tiny_llama

    m = tiny_llama(golden_e2e["init_state"]).to(DEV)
    ids = golden_e2e["ids"]
    with torch.no_grad():
        fp = m(golden_e2e["probe"].to(DEV)).logits.float().cpu()

    def run_fn(model):
        for x in ids:
            model(x.to(DEV))

    q = Q.quantize(m, Q.SmoothQuantConfig(alpha=0.5), run_fn=run_fn, example_inputs=ids[0].to(DEV))
    assert isinstance(q.model.layers[0].self_attn.q_proj, SQLinear)
    with torch.no_grad():
        logits = q(golden_e2e["probe"].to(DEV)).logits.float().cpu()
    assert torch.isfinite(logits).all()
    assert float((logits - fp).norm() / fp.norm()) < 0.1
