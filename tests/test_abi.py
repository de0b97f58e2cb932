"""The C-ABI library loads and exports every symbol include/seedvc_b200.h declares (no compute
calls: there is no GPU where the CPU suite runs)."""
import ctypes
import os
import re

import pytest

import seedvc_b200  # noqa: F401
from seedvc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "seedvc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_binding_covers_header():
    bound = set(_lib.SIGNATURES) | {"svc_last_error", "svc_version"}
    assert set(declared_symbols()) == bound


def test_load_library_and_version(lib):
    l = _lib.load_library()
    assert l.svc_version() == 101
    assert l.svc_last_error() is not None


def test_struct_layout_matches_header():
    """sizeof(svc_gemm_desc) as the compiler sees it == the ctypes mirror."""
    import subprocess
    import tempfile
    code = ('#include <stdio.h>\n#include "seedvc_b200.h"\nint main(){printf("%zu %zu %zu %zu", sizeof(svc_gemm_desc), '
            'sizeof(svc_dit_weights), sizeof(svc_dit_state), sizeof(svc_bigvgan_weights));}')
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(code)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        n = [int(v) for v in subprocess.check_output([exe]).decode().split()]
    assert n == [ctypes.sizeof(_lib.GemmDesc), ctypes.sizeof(_lib.DitWeights), ctypes.sizeof(_lib.DitState),
                 ctypes.sizeof(_lib.BigVGANWeights)]


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_lib.SvcError):
        _lib.load_library(str(tmp_path / "nope.so"))
