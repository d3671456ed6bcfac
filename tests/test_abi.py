"""The C-ABI library loads without a GPU and exports every symbol include/rtmpc.h declares."""
import ctypes
import os
import re

import pytest

import helpers as H
from rtmpc_b200 import _lib


def header_functions():
    txt = open(os.path.join(H.ROOT, "include", "rtmpc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rtmpc_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run `python __graft_entry__.py` (build()) first"


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rtmpc.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_abi_version_and_no_cpu_fallback():
    L = _lib.lib()
    assert L.rtmpc_abi_version() == _lib.ABI_VERSION == 3
    if L.rtmpc_device_count() <= 0:
        with pytest.raises(_lib.RtmpcError):
            _lib.require_cuda()
        import numpy as np
        from rtmpc_b200.qp import BatchedQP
        s = H.load("sets_di.npz")
        with pytest.raises(_lib.RtmpcError):
            BatchedQP(H.spec_tube_tracking(s), Kss=s["K"])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(H.PKG, "rtmpc_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# ", ""), fn
