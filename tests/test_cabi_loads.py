"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/b200woq.h declares (no compute calls -- there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200woq.h")).read()
    return sorted(set(re.findall(r"\b(b200woq_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from neural_compressor_b200 import _build, _lib

    _build.build()  # nvcc cross-compiles for sm_100a without a GPU
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200woq.h but not exported"
    assert set(declared) == set(_lib.exported_symbols()), "ctypes signature table out of sync with the header"
    assert lib.b200woq_version() == 100


def test_sass_is_sm100a_and_uses_tensor_cores():
    import subprocess

    from neural_compressor_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_ops_fail_loudly_without_cuda():
    import torch

    from neural_compressor_b200 import ops
    from neural_compressor_b200._lib import B200WOQError

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(B200WOQError):
        ops.rtn_params(torch.randn(4, 8), 4, -1, True)
