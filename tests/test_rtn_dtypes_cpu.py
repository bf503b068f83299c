"""nf4 / fp4 tables and the non-optimum module layouts, on the host.

* The per-element math of csrc/f4_math.cuh -- the header the CUDA kernels of csrc/float4.cu are built from -- is compiled
  here with g++ (tests/host/f4_host.cpp) and checked bit for bit against tensors written by the live reference's
  `quantize_4bit` / `recover` (tests/golden/rtn_dtypes.pt, oracle/gen_golden.py rtn_dtypes).
* `B200WeightOnlyLinear(use_optimum_format=False, compression_dtype=..., compression_dim=...)` packs / unpacks / recovers
  with integer tensor ops on any device: the reference's module test matrix (test_woq_module.py:10-52) runs on the CPU.
* The C ABI rejects a malformed table before touching the device.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from neural_compressor_b200 import dtypes as D

HERE = os.path.dirname(os.path.abspath(__file__))
ROUNDING = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(HERE, "golden", "rtn_dtypes.pt"))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("f4host") / "f4_host.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    os.path.join(HERE, "host", "f4_host.cpp")], check=True)
    return ctypes.CDLL(str(so))


def host_quantize(lib, W, dtype, group_size, quantile):
    N, K = W.shape
    g = K if (group_size <= 0 or group_size > K) else group_size
    G = -(-K // g)
    Wf = W.float().contiguous()
    codes, scale, fake = np.zeros((N, K), np.int8), np.zeros((N, G), np.float32), np.zeros((N, K), np.float32)
    table = D.table(dtype)
    lib.f4_quantize_host(ctypes.c_void_p(Wf.data_ptr()), ctypes.c_int64(N), ctypes.c_int64(K), ctypes.c_int(g),
                         ctypes.byref(table), ctypes.c_float(quantile), ctypes.c_int(ROUNDING[W.dtype]),
                         codes.ctypes.data_as(ctypes.c_void_p), scale.ctypes.data_as(ctypes.c_void_p),
                         fake.ctypes.data_as(ctypes.c_void_p))
    return torch.from_numpy(codes), torch.from_numpy(scale), torch.from_numpy(fake)


def test_tables_are_consistent():
    for name, levels in D.FLOAT_MAPPING.items():
        codes = D.INT_MAPPING[name]
        assert len(levels) == len(codes) and list(levels) == sorted(levels) and len(set(codes)) == len(codes)
        assert all(-8 <= c <= 7 for c in codes)
        t = D.table(name)
        assert t.n == len(levels) and t.max_level == max(levels)
        lv = list(D.nibble_levels(name))
        for level, c in zip(levels, codes):
            assert lv[c & 0xF] == np.float32(level)


def test_f4_math_header_matches_reference_quantize_4bit(host_lib, golden):
    assert len(golden["quant"]) >= 7
    for case in golden["quant"]:
        codes, scale, fake = host_quantize(host_lib, case["W"], case["dtype"], case["group_size"], case["quantile"])
        tag = (case["dtype"], case["W"].dtype, tuple(case["W"].shape), case["group_size"], case["quantile"])
        assert torch.equal(codes, case["codes"]), tag
        assert torch.equal(scale, case["scale"]), tag
        want = case["fake"].float()
        assert torch.equal(torch.nan_to_num(fake, nan=7.0), torch.nan_to_num(want, nan=7.0)), tag


def test_f4_recover_matches_reference_module(host_lib, golden):
    """codes -> row-major words (integer ops) -> f4_recover == level * scale of the reference's unpack()/recover()."""
    from neural_compressor_b200.algorithms.modules_rowmajor import pack_fields

    for case in golden["quant"]:
        codes, scale = case["codes"], case["scale"]
        N, K = codes.shape
        gs = case["group_size"]
        g = K if (gs <= 0 or gs > K) else gs
        qw = pack_fields(codes, 4, torch.int32).contiguous()
        out = np.zeros((N, K), np.float32)
        host_lib.f4_recover_host(ctypes.c_void_p(qw.data_ptr()), ctypes.c_void_p(scale.data_ptr()), D.nibble_levels(case["dtype"]),
                                 ctypes.c_int64(N), ctypes.c_int64(K), ctypes.c_int(g), out.ctypes.data_as(ctypes.c_void_p))
        lut = torch.zeros(16)
        for level, c in zip(D.FLOAT_MAPPING[case["dtype"]], D.INT_MAPPING[case["dtype"]]):
            lut[c & 0xF] = level
        want = lut[codes.long() & 0xF] * scale[:, torch.arange(K) // g]
        assert torch.equal(torch.from_numpy(out), want)


def test_rowmajor_layouts_match_reference(golden):
    from neural_compressor_b200.algorithms.modules import B200WeightOnlyLinear
    from neural_compressor_b200.algorithms.modules_rowmajor import B200RowMajorLinear

    assert len(golden["rowmajor"]) >= 18
    for c in golden["rowmajor"]:
        m = B200WeightOnlyLinear(96, 24, dtype="int", bits=c["bits"], group_size=32, zp=c["zp"] is not None, bias=True,
                                 use_optimum_format=False, compression_dtype=c["compression_dtype"],
                                 compression_dim=c["compression_dim"], device="cpu")
        assert isinstance(m, B200RowMajorLinear)
        m.pack(c["int_weight"].clone(), c["scale"].clone(), None if c["zp"] is None else c["zp"].clone(), c["bias"])
        tag = (c["bits"], c["compression_dtype"], c["compression_dim"], c["scheme"])
        assert torch.equal(m.qweight, c["qweight"]) and m.qweight.dtype == c["compression_dtype"], tag
        assert torch.equal(m.scales, c["scales"]), tag
        if c["zp"] is not None:
            assert torch.equal(m.qzeros, c["qzeros"]), tag
        u = m.unpack()
        assert torch.equal(u["int_weight"].float(), c["int_weight"].float()), tag    # test_woq_module.py:52
        if c["zp"] is not None:
            assert torch.equal(u["zp"].float(), c["zp"].float()), tag
        assert torch.equal(m.recover(), c["recover"]), tag
        assert torch.equal(m(c["x"]), c["y"]), tag
        sd = m.state_dict()
        m2 = B200WeightOnlyLinear(96, 24, dtype="int", bits=c["bits"], group_size=32, zp=c["zp"] is not None, bias=True,
                                  use_optimum_format=False, compression_dtype=c["compression_dtype"],
                                  compression_dim=c["compression_dim"], device="cpu")
        m2.load_state_dict(sd)
        assert torch.equal(m2.recover(), c["recover"]), tag


def test_table_module_half_cast_keeps_storage_types():
    from neural_compressor_b200.algorithms.modules import B200WeightOnlyLinear

    m = B200WeightOnlyLinear(64, 16, dtype="nf4", bits=4, group_size=32, bias=True, device="cpu")
    m.half()
    assert m.scales.dtype == torch.float32 and m.qweight.dtype == torch.int32 and m.bias.dtype == torch.float32
    with pytest.raises(ValueError):
        B200WeightOnlyLinear(64, 16, dtype="nf4", bits=8, device="cpu")


def test_cabi_rejects_malformed_table():
    """Argument validation happens before any CUDA call, so it can be exercised without a GPU."""
    from neural_compressor_b200 import _lib

    lib = _lib.load()
    bad = D.F4Table()
    bad.n = 99
    dummy = ctypes.c_void_p(0x1000)
    rc = lib.b200woq_f4_quantize(dummy, 0, 4, 32, 32, ctypes.c_void_p(ctypes.addressof(bad)), 1.0, dummy, dummy, None, None)
    assert rc == -1 and b"table" in lib.b200woq_last_error()
    t = D.table("nf4")
    desc = D.F4Table.from_buffer_copy(t)
    desc.level[3] = 5.0   # not ascending any more
    rc = lib.b200woq_f4_quantize(dummy, 0, 4, 32, 32, ctypes.c_void_p(ctypes.addressof(desc)), 1.0, dummy, dummy, None, None)
    assert rc == -1 and b"ascend" in lib.b200woq_last_error()
    rc = lib.b200woq_pack_rows(dummy, 4, 32, 9, dummy, None)
    assert rc == -1 and b"bits" in lib.b200woq_last_error()
