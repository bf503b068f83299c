"""Host logic of RTN over the reference's dtype matrix (algorithms/rtn.py: table dtypes -> row-major module, fp8 cast,
double quantisation of the scales, ragged-K rule) with the device kernels replaced by their host twins: the g++ build of
the product's own f4_math.cuh for nf4 / fp4, the oracle for int RTN.  What is checked is the wiring -- which tensor goes
into which buffer in which layout and dtype -- against the packed state dicts and logits of the UNMODIFIED reference
(tests/golden/rtn_dtypes.pt).  The kernels themselves are covered by tests/test_zz_rtn_dtypes_gpu.py on the B200."""
import os

import pytest
import torch

from tests.test_rtn_dtypes_cpu import golden, host_lib, host_quantize  # noqa: F401  (fixtures)

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture()
def host_ops(monkeypatch, host_lib):  # noqa: F811
    from neural_compressor_b200 import ops
    from neural_compressor_b200.algorithms import rtn
    from neural_compressor_b200.algorithms.modules_rowmajor import pack_fields
    from oracle import woq_oracle as O

    def f4_quantize(W, dtype, group_size=-1, quantile=1.0, want_codes=True, fake_out=None):
        codes, scale, fake = host_quantize(host_lib, W, dtype, group_size, quantile)
        if fake_out is not None:
            fake_out.copy_(fake.to(W.dtype))
        return dict(codes=codes if want_codes else None, scale=scale)

    def rtn_quant_pack(W, bits=4, group_size=-1, sym=False, full_range=False, quantile=1.0, return_codes=False):
        q, s, z = O.rtn_quantize(W, bits, group_size, "sym" if sym else "asym", quantile, full_range)
        qweight, qzeros, scales16 = O.pack_optimum(q, s, z, bits, group_size)
        return dict(qweight=qweight, qzeros=qzeros, scales=scales16, scale_f32=s.float(), zp_f32=None if z is None else z.float())

    def rtn_fake_quant(W, bits=4, group_size=-1, sym=False, full_range=False, quantile=1.0, col_scale=None, out=None):
        assert col_scale is None
        r = O.rtn_fake_quant(W, bits, group_size, "sym" if sym else "asym", quantile, full_range)
        return r if out is None else out.copy_(r)

    def pack_params(scale, zp, bits):
        _, qzeros, scales16 = O.pack_optimum(torch.zeros(scale.shape[0], 1), scale, zp, bits, 1)
        return scales16, qzeros

    for name, fn in dict(f4_quantize=f4_quantize, rtn_quant_pack=rtn_quant_pack, rtn_fake_quant=rtn_fake_quant,
                         pack_params=pack_params, pack_rows=lambda c, b: pack_fields(c, b, torch.int32)).items():
        monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(rtn, "current_device", lambda: torch.device("cpu"))
    return ops


def tiny_llama(init_state):
    from tests.test_api_gpu import tiny_llama as build

    return build(init_state)


@pytest.mark.parametrize("tag", ["rtn_nf4", "rtn_fp4", "rtn_fp4_e2m1_bnb", "rtn_fp4_e2m1", "rtn_nf4_mse", "rtn_fp8_e4m3fn",
                                 "rtn_fp8_e5m2", "rtn_int4_dq_asym", "rtn_int4_dq_sym", "rtn_nf4_dq"])
def test_rtn_dtype_host_flow(host_ops, golden, golden_e2e, tag):  # noqa: F811
    import neural_compressor_b200.quantization as api

    case = golden["models"][tag]
    m = tiny_llama(golden_e2e["init_state"])
    m = api.convert(api.prepare(m, api.RTNConfig(use_layer_wise=False, **case["kw"])))
    state = m.state_dict()
    assert len(case["state"]) >= 14
    for k, ref in case["state"].items():
        got = state[k]
        if "fp8" in tag:
            assert torch.equal(got.float(), ref.float()), k
            continue
        assert got.shape == ref.shape and got.dtype == ref.dtype, (k, got.shape, ref.shape, got.dtype, ref.dtype)
        assert torch.equal(got, ref), k
    if "int4" not in tag:   # table-dtype and fp8 models run on the host (recover + dense GEMM); packed int modules need the B200
        with torch.no_grad():
            logits = m(golden["probe"]).logits
        assert torch.allclose(logits, case["logits"], atol=1e-5, rtol=1e-5)


def test_double_quant_skipped_for_ragged_k(host_ops):
    """quant_tensor returns from its ragged-tail branch before the second level (utility.py:334-373)."""
    import neural_compressor_b200.quantization as api

    lin = torch.nn.Sequential(torch.nn.Linear(100, 16, bias=False))
    ref = torch.nn.Sequential(torch.nn.Linear(100, 16, bias=False))
    ref.load_state_dict(lin.state_dict())
    a = api.convert(api.prepare(lin, api.RTNConfig(group_size=32, use_double_quant=True, use_layer_wise=False)))
    b = api.convert(api.prepare(ref, api.RTNConfig(group_size=32, use_layer_wise=False)))
    assert torch.equal(a[0].scales, b[0].scales) and torch.equal(a[0].qweight, b[0].qweight)


def test_group_dim0_on_linear_fails_like_the_reference(host_ops):
    """`RTNConfig(group_dim=0)` on an nn.Linear dies in the reference with "Scale shape is mismatched" (modules.py:345)."""
    import neural_compressor_b200.quantization as api

    m = torch.nn.Sequential(torch.nn.Linear(64, 32))
    with pytest.raises(AssertionError, match="Scale shape is mismatched"):
        api.convert(api.prepare(m, api.RTNConfig(group_dim=0, group_size=32, use_layer_wise=False)))
