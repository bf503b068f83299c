"""Size-independent properties of the integer layout code (hypothesis): the word packers against a naive per-element
loop, pack -> unpack round trips for every field width / word type / packing axis, and the AutoAWQ repack against an
independently written inverse."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from neural_compressor_b200.algorithms.awq_repack import repack_awq_to_optimum_format
from neural_compressor_b200.algorithms.modules_rowmajor import pack_fields, unpack_fields

WORD = {8: torch.int8, 16: torch.int16, 32: torch.int32, 64: torch.int64}


def naive_pack(values, bits, cbits):
    n_pack = cbits // bits
    rows, cols = values.shape
    words = -(-cols // n_pack)
    out = torch.zeros(rows, words, dtype=torch.int64)
    for r in range(rows):
        for c in range(cols):
            field = int(values[r, c]) & ((1 << bits) - 1)
            w = int(out[r, c // n_pack]) | (field << (bits * (c % n_pack)))
            w &= (1 << cbits) - 1
            if w >= 1 << (cbits - 1):
                w -= 1 << cbits
            out[r, c // n_pack] = w
    return out.to(WORD[cbits])


@settings(max_examples=60, deadline=None)
@given(bits=st.integers(1, 8), cbits=st.sampled_from([8, 16, 32, 64]), rows=st.integers(1, 5), cols=st.integers(1, 40),
       signed=st.booleans(), seed=st.integers(0, 2**16))
def test_pack_fields_equals_the_naive_loop_and_round_trips(bits, cbits, rows, cols, signed, seed):
    g = torch.Generator().manual_seed(seed)
    lo, hi = (-(1 << (bits - 1)), 1 << (bits - 1)) if signed else (0, 1 << bits)
    v = torch.randint(lo, hi, (rows, cols), generator=g)
    packed = pack_fields(v, bits, WORD[cbits])
    assert packed.dtype == WORD[cbits] and torch.equal(packed, naive_pack(v, bits, cbits))
    back = unpack_fields(packed, bits, signed)[:, :cols]
    assert torch.equal(back.to(torch.int64), v)


@settings(max_examples=25, deadline=None)
@given(k8=st.integers(1, 6), n8=st.integers(1, 5), groups=st.sampled_from([1, 2, 4]), seed=st.integers(0, 2**16))
def test_awq_repack_inverts_the_autoawq_layout(k8, n8, groups, seed):
    from tests.test_save_load_cpu import _optimum_to_autoawq

    g = torch.Generator().manual_seed(seed)
    K, N = 8 * k8 * groups, 8 * n8
    gs = K // groups
    codes = torch.randint(0, 16, (K, N), generator=g)
    zeros = torch.randint(0, 16, (groups, N), generator=g)
    qweight = pack_fields(codes.t().contiguous(), 4, torch.int32).t().contiguous()          # optimum: packed along K
    qzeros = pack_fields((zeros - 1) & 0xF, 4, torch.int32)                                   # optimum: zp - 1 along N
    scales = torch.rand(groups, N, generator=g).half()
    aw, az = _optimum_to_autoawq(qweight, qzeros)
    rw, rz, rs = repack_awq_to_optimum_format(aw, az, scales, 4, gs)
    assert torch.equal(rw, qweight) and torch.equal(rz, qzeros) and rs is scales


def _live_utility():
    import pytest

    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not present")
    load_reference()
    import neural_compressor.torch.algorithms.weight_only.utility as U

    return U


@settings(max_examples=60, deadline=None)
@given(bits=st.integers(2, 8), sym=st.booleans(), n=st.integers(1, 6), k=st.integers(4, 70),
       gs=st.sampled_from([-1, 4, 8, 16, 32]), full_range=st.booleans(), quantile=st.sampled_from([1.0, 0.93, 0.805]),
       dtype=st.sampled_from([torch.float32, torch.float16, torch.bfloat16]), seed=st.integers(0, 2**16))
def test_oracle_rtn_equals_live_quant_tensor(bits, sym, n, k, gs, full_range, quantile, dtype, seed):
    """The oracle's restatement of `quant_tensor` on random shapes (ragged tails included), schemes, clip quantiles and
    storage dtypes -- beyond the fixed fixtures -- bit for bit against the live reference."""
    from oracle import woq_oracle as O

    U = _live_utility()
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(n, k, generator=g) * 0.05).to(dtype)
    scheme = "sym" if sym else "asym"
    q, s, z = U.quant_tensor(w.clone(), bits=bits, group_size=gs, scheme=scheme, quantile=quantile, return_int=True,
                             full_range=full_range)
    oq, os_, oz = O.rtn_quantize(w, bits, gs, scheme, quantile, full_range)
    assert torch.equal(oq.float(), q.float()) and torch.equal(os_.float(), s.float())
    assert (z is None) == (oz is None) and (z is None or torch.equal(oz.float(), z.float()))
    fq = U.quant_tensor(w.clone(), bits=bits, group_size=gs, scheme=scheme, quantile=quantile, full_range=full_range)
    assert torch.equal(O.rtn_fake_quant(w, bits, gs, scheme, quantile, full_range).float(), fq.float())


_HOST = {}


def _host_f4():
    """The g++ build of the product's f4_math.cuh (tests/host/f4_host.cpp), built once per session."""
    if "lib" not in _HOST:
        import ctypes
        import os
        import subprocess
        import tempfile

        here = os.path.dirname(os.path.abspath(__file__))
        so = os.path.join(tempfile.mkdtemp(prefix="f4host"), "f4_host.so")
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(here, "host", "f4_host.cpp")],
                       check=True)
        _HOST["lib"] = ctypes.CDLL(so)
    return _HOST["lib"]


@settings(max_examples=40, deadline=None)
@given(name=st.sampled_from(["nf4", "fp4", "fp4_e2m1_bnb", "fp4_e2m1"]), n=st.integers(1, 5), k=st.integers(2, 70),
       gs=st.sampled_from([-1, 8, 16, 32]), quantile=st.sampled_from([1.0, 0.9, 0.805]),
       dtype=st.sampled_from([torch.float32, torch.float16, torch.bfloat16]), seed=st.integers(0, 2**16))
def test_f4_kernel_math_equals_live_quantize_4bit(name, n, k, gs, quantile, dtype, seed):
    """csrc/f4_math.cuh (host build) against the live reference's `quantize_4bit` on random shapes, group sizes (ragged
    tails), clip quantiles and storage dtypes: integer codes, scales and fake-quantised values, bit for bit."""
    from tests.test_rtn_dtypes_cpu import host_quantize

    U = _live_utility()
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(n, k, generator=g) * 0.05).to(dtype)
    q, s, _ = U.quant_tensor(w.clone(), dtype=name, group_size=gs, quantile=quantile, return_int=True)
    fq = U.quant_tensor(w.clone(), dtype=name, group_size=gs, quantile=quantile)
    codes, scale, fake = host_quantize(_host_f4(), w, name, gs, quantile)
    assert torch.equal(codes.float(), q.float()) and torch.equal(scale, s.float())
    assert torch.equal(torch.nan_to_num(fake, nan=7.0), torch.nan_to_num(fq.float(), nan=7.0))


@settings(max_examples=30, deadline=None)
@given(bits=st.sampled_from([2, 3, 4, 8]), sym=st.booleans(), n=st.integers(2, 12), groups=st.integers(1, 4),
       gs=st.sampled_from([8, 16, 32]), blocksize=st.sampled_from([8, 16, 32, 2048]), act_order=st.booleans(),
       mse=st.booleans(), seed=st.integers(0, 2**16))
def test_oracle_gptq_equals_live_fasterquant(bits, sym, n, groups, gs, blocksize, act_order, mse, seed):
    """The oracle's Hessian accumulation + `fasterquant` restatement against the live reference's GPTQ class on random
    layer shapes, block sizes (smaller than, equal to and larger than the group), act_order, the mse grid, a dead input
    channel and an outlier one: fake-quantised weights, scales and zeros bit for bit."""
    import pytest

    from oracle import woq_oracle as O
    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not present")
    load_reference()
    from neural_compressor.torch.algorithms.weight_only.gptq import GPTQ

    if blocksize < gs:            # the engine's contract: a block is a multiple of the group (or the whole layer)
        blocksize = gs
    C = gs * groups
    g = torch.Generator().manual_seed(seed)
    lin = torch.nn.Linear(C, n, bias=False)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(n, C, generator=g) * 0.05)
    X = [torch.randn(1, 3 * C, C, generator=g) for _ in range(3)]
    for x in X:
        x[..., 1] *= 5.0
        if C > 8:
            x[..., 5] = 0.0
    gp = GPTQ(lin, lin.weight.data.clone(), "cpu")
    gp.quantizer.configure(dict(dtype="int", bits=bits, sym=sym, group_size=gs, mse=mse, perchannel=True,
                                use_double_quant=False, double_quant_sym=False))
    oracle = O.GPTQLayerOracle(n, C, bits=bits, sym=sym, mse=mse)
    for x in X:
        gp.add_batch(x, None)
        oracle.add_batch(x)
    assert torch.equal(oracle.H, gp.H)
    scale, _, zero, Q = gp.fasterquant(lin.weight.data.clone(), blocksize=blocksize, percdamp=0.01, groupsize=gs,
                                       act_order=act_order)
    r = oracle.fasterquant(lin.weight.data.clone(), blocksize=blocksize, percdamp=0.01, groupsize=gs, act_order=act_order)
    assert torch.equal(r["Q"], Q.float()) and torch.equal(r["scale"], scale) and torch.equal(r["zero"], zero)


@settings(max_examples=40, deadline=None)
@given(bits=st.sampled_from([2, 3, 4, 8]), sym=st.booleans(), n=st.integers(1, 20), k=st.integers(8, 80),
       gs=st.sampled_from([-1, 8, 16, 32]), seed=st.integers(0, 2**16))
def test_oracle_pack_recover_forward_equal_live_module(bits, sym, n, k, gs, seed):
    """The oracle's optimum-format pack / unpack / recover / forward against the live `INCWeightOnlyLinear` on random
    shapes (K not a multiple of n_pack or of the group, N not a multiple of n_pack), widths and schemes."""
    import pytest

    from oracle import woq_oracle as O
    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not present")
    load_reference()
    from neural_compressor.torch.algorithms.weight_only.modules import INCWeightOnlyLinear
    from neural_compressor.torch.algorithms.weight_only.utility import quant_tensor

    g = torch.Generator().manual_seed(seed)
    w = torch.randn(n, k, generator=g) * 0.05
    bias = torch.randn(n, generator=g) * 0.01
    scheme = "sym" if sym else "asym"
    q, s, z = quant_tensor(w.clone(), bits=bits, group_size=gs, scheme=scheme, return_int=True)
    group = k if gs == -1 or k < gs else gs
    m = INCWeightOnlyLinear(k, n, dtype="int", bits=bits, group_size=group, zp=z is not None, bias=True, device="cpu")
    m.pack(q.clone(), s.clone(), None if z is None else z.clone(), bias)
    qweight, qzeros, scales = O.pack_optimum(q, s, z, bits, group)
    assert torch.equal(qweight, m.qweight) and torch.equal(qzeros, m.qzeros) and torch.equal(scales, m.scales)
    rec = m.recover()
    assert torch.equal(O.recover_fp16(qweight, qzeros, scales, bits, group, k, n).float(), rec.float())
    x = torch.randn(3, k, generator=g)
    want = m(x)
    got = O.woq_linear_forward(x, qweight, qzeros, scales, m.bias, bits, group, k, n)
    assert torch.allclose(got.float(), want.float(), rtol=1e-5, atol=1e-6)
