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
