"""CPU: the C-ABI library loads and exports every symbol include/sfdtd.h declares; the ctypes
mirror of sfdtd_args has the C layout; argument validation works without a GPU."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "sfdtd.h")


@pytest.fixture(scope="module")
def lib():
    from torch_fdtd_string_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfdtd_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert "sfdtd_forward" in names and len(names) >= 4
    for n in names:
        assert hasattr(lib, n), n
    from torch_fdtd_string_b200 import _lib
    assert sorted(_lib.EXPORTS) == names


def test_struct_layout_matches_header():
    from torch_fdtd_string_b200 import _lib
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "sz.c")
        open(src, "w").write('#include "sfdtd.h"\n#include <stdio.h>\n#include <stddef.h>\n'
                             'int main(){printf("%zu %zu %zu %zu\\n", sizeof(sfdtd_args), offsetof(sfdtd_args, state_u),'
                             ' offsetof(sfdtd_args, bow_mask), offsetof(sfdtd_args, counters));return 0;}\n')
        exe = os.path.join(d, "sz")
        subprocess.check_call(["gcc", f"-I{os.path.dirname(HDR)}", src, "-o", exe])
        size, o_su, o_bm, o_cnt = map(int, subprocess.check_output([exe]).split())
    assert ctypes.sizeof(_lib.Args) == size
    assert _lib.Args.state_u.offset == o_su
    assert _lib.Args.bow_mask.offset == o_bm
    assert _lib.Args.counters.offset == o_cnt


def test_argument_validation_without_gpu(lib):
    from torch_fdtd_string_b200 import _lib
    assert lib.sfdtd_abi_version() == _lib.SFDTD_ABI_VERSION
    assert lib.sfdtd_forward(None, None) == -1
    a = _lib.Args()
    a.abi_version = 999
    assert lib.sfdtd_forward(ctypes.byref(a), None) == -1
    assert b"abi_version" in lib.sfdtd_last_error()
    a.abi_version = _lib.SFDTD_ABI_VERSION
    a.dtype = 7
    assert lib.sfdtd_forward(ctypes.byref(a), None) == -2


def test_no_fallback_when_library_missing(monkeypatch):
    from torch_fdtd_string_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsfdtd.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "torch_fdtd_string_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert "sfdtd_oracle" not in txt and "oracle/" not in txt, f
