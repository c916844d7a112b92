"""CPU: the C-ABI library loads and exports every symbol include/*.h
declare (no compute calls without a GPU); the product fails loudly without a
device."""

import ctypes
import os
import re

import pytest

from hy_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    inc = os.path.join(ROOT, "include")
    txt = "".join(open(os.path.join(inc, f)).read() for f in sorted(os.listdir(inc)) if f.endswith(".h"))
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hy_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_exported():
    lib = _cabi.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libhy_cuda.so does not export {}".format(n)
    assert sorted(_cabi.SYMBOLS) == names


def test_struct_sizes_match_header():
    from hy_b200 import decompose as D

    assert D.op_dtype.itemsize == 32 and D.term_dtype.itemsize == 24
    assert ctypes.sizeof(_cabi.dims_t) == 36
    assert ctypes.sizeof(_cabi.event_rec_t) == 24 and _cabi.event_rec_dtype.itemsize == 24
    assert ctypes.sizeof(_cabi.launch_info_t) == 36  # 9 x uint32 (incl. kernel_variant)


def test_no_cpu_fallback():
    if _cabi.device_count() > 0:
        pytest.skip("a GPU is visible")
    import numpy as np
    import hy_b200 as hy

    x, v = hy.make_vars("x", "v")
    with pytest.raises(_cabi.HyCudaError):
        hy.taylor_adaptive_batch([(x, v), (v, -x)], np.zeros((2, 4)))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "heyoka.py_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("oracle/", "").replace("the oracle", "").replace(
                    "C oracle", "").replace("CPU oracle", "") or f == "workloads.py", f
