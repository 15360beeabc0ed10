"""tcgen05 building blocks (descriptors, TMA swizzle, TMEM operand layouts) on the real chip."""
import ctypes as C

import pytest
import torch

from mygenerativerecommenders_b200 import _lib

pytestmark = pytest.mark.gpu

MODES = ["SS K/K N128 K128", "SS K/MN N64 K128", "TS tmem/MN N64 K128", "TS tmem/K N128 K64",
         "SS MN/MN N64 K128", "SS K/K N256 K256", "SS MN/K N128 K64"]


def test_every_operand_mode_is_exact():
    errs = (C.c_float * 16)()
    n = _lib.lib().grb_selftest_umma(errs, 16, _lib.stream_ptr(torch.device("cuda")))
    if n < 0:
        _lib.check(n)
    assert n == len(MODES)
    report = {MODES[i]: errs[i] for i in range(n)}
    print(report)
    # inputs are multiples of 1/8 in [-2, 2]: every product and partial sum is exact in fp32
    bad = {k: v for k, v in report.items() if v != 0.0}
    assert not bad, f"tcgen05 operand modes with wrong results: {bad}"
