"""The 4-bit TABLE data types of the reference's weight-only path (weight_only/utility.py:52-103): `nf4` (QLoRA's
normal-float levels), `fp4` = `fp4_e2m1_bnb` (bitsandbytes' FP4) and `fp4_e2m1`.  A type is its ascending list of levels
plus the signed 4-bit integer the reference stores for each level; both lists are part of the on-disk contract (a packed
nf4 checkpoint is read back through them), so the values here are the reference's, digit for digit.
"""
from __future__ import annotations

import ctypes

import numpy as np

NF4_LEVELS = (-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
              -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
              0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0)
NF4_CODES = (7, 1, 2, 3, 4, 5, 6, 0, -8, -7, -6, -5, -4, -3, -2, -1)
FP4_BNB_LEVELS = (-12.0, -8.0, -6.0, -4.0, -3.0, -2.0, -0.0625, 0.0, 0.0625, 2.0, 3.0, 4.0, 6.0, 8.0, 12.0)
FP4_BNB_CODES = (-5, -6, -3, -4, -1, -2, -7, 0, 1, 6, 7, 4, 5, 2, 3)
_THIRD, _SIXTH = 1.0 / 3.0, 1.0 / 6.0
FP4_E2M1_LEVELS = (-1.0, -2 * _THIRD, -0.5, -_THIRD, -0.25, -_SIXTH, -0.0625 * _SIXTH, 0.0, 0.0625 * _SIXTH, _SIXTH, 0.25,
                   _THIRD, 0.5, 2 * _THIRD, 1.0)
FP4_E2M1_CODES = (-1, -2, -3, -4, -5, -6, -7, 0, 1, 2, 3, 4, 5, 6, 7)

FLOAT_MAPPING = {"nf4": NF4_LEVELS, "fp4": FP4_BNB_LEVELS, "fp4_e2m1_bnb": FP4_BNB_LEVELS, "fp4_e2m1": FP4_E2M1_LEVELS}
INT_MAPPING = {"nf4": NF4_CODES, "fp4": FP4_BNB_CODES, "fp4_e2m1_bnb": FP4_BNB_CODES, "fp4_e2m1": FP4_E2M1_CODES}
FP8_DTYPES = ("fp8_e5m2", "fp8_e5m2fnuz", "fp8_e4m3fn", "fp8_e4m3fnuz")


def is_table_dtype(dtype) -> bool:
    return str(dtype) in FLOAT_MAPPING


class F4Table(ctypes.Structure):
    """`b200woq_f4_table` (include/b200woq.h)."""

    _fields_ = [("n", ctypes.c_int32), ("level", ctypes.c_float * 16), ("mid", ctypes.c_float * 16),
                ("code", ctypes.c_int32 * 16), ("max_level", ctypes.c_float)]


_TABLES = {}


def table(dtype: str) -> F4Table:
    """The ctypes table of a data type.  Mid points are evaluated in double and converted to float, as torch converts
    the Python scalar `(a[i] + a[i+1]) / 2` before comparing it with a tensor (utility.py:141)."""
    dtype = str(dtype)
    if dtype not in _TABLES:
        levels, codes = FLOAT_MAPPING[dtype], INT_MAPPING[dtype]
        t = F4Table()
        t.n = len(levels)
        for i, (lv, c) in enumerate(zip(levels, codes)):
            t.level[i] = float(np.float32(lv))
            t.code[i] = c
        for i in range(len(levels) - 1):
            t.mid[i] = float(np.float32((levels[i] + levels[i + 1]) / 2))
        t.max_level = float(np.float32(max(levels)))
        _TABLES[dtype] = t
    return _TABLES[dtype]


def nibble_levels(dtype: str):
    """16 floats indexed by the stored 4-bit field: the level of its sign-extended code, 0 for codes the type does not
    use -- what `unpack()` produces through `int2float_mapping` (modules.py:392-396)."""
    out = (ctypes.c_float * 16)()
    for lv, c in zip(FLOAT_MAPPING[str(dtype)], INT_MAPPING[str(dtype)]):
        out[c & 0xF] = float(np.float32(lv))
    return out
