"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ssgpu.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from statespacer_b200 import native

HEADER = os.path.join(ROOT, "include", "ssgpu.h")

pytestmark = pytest.mark.skipif(not os.path.exists(native.LIB_PATH), reason="libssgpu.so not built")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssgpu_[A-Za-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    L = native.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ssgpu.h but not exported"
    assert sorted(native.SYMBOLS) == names


def test_version_and_error_string():
    L = native.lib()
    assert L.ssgpu_version() == 100
    assert isinstance(L.ssgpu_last_error(), bytes)


def test_argument_errors_do_not_need_a_device():
    L = native.lib()
    out = C.c_double()
    y = np.zeros(4)
    rc = L.ssgpu_LogLikC(y.ctypes.data, None, 4, 1, 0, 1, None, None, None, None, 1, None, 1, None, 1, None, 1,
                         C.byref(out))
    assert rc == 1 and L.ssgpu_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    from statespacer_b200 import api, sysmat
    sm = sysmat.nile_local_level(1.0, 1.0)
    y = np.arange(10.0)[:, None]
    with pytest.raises(native.SsgpuError) as e:
        api.LogLikC(y, np.isnan(y), *sm.args())
    assert e.value.status == 2
