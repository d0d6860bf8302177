"""The C-ABI library loads and exports every symbol include/b200vs.h declares; without a GPU the entry points
fail loudly (non-OK status + message) instead of falling back to any CPU path."""
import ctypes
import os
import re

import pytest

import b200vs

HEADER = os.path.join(b200vs.REPO_ROOT, "include", "b200vs.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200vs_[a-z_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(b200vs.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = b200vs.lib()
    for name in declared_functions():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.b200vs_version()


def test_no_silent_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b200vs.B200VSError) as e:
        b200vs.Index(b200vs.FLAT, b200vs.L2, 8)
    assert e.value.code == b200vs.EINTERNAL and e.value.msg


def test_product_library_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["nm", "-D", b200vs.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in out
    ldd = subprocess.run(["ldd", b200vs.LIB_PATH], capture_output=True, text=True).stdout
    assert "liboracle" not in ldd and "dingo_simd_ref" not in ldd
