"""CPU: the C-ABI shared library loads and exports every symbol include/tic_b200.h declares, and the ctypes
prototypes in the Python binding agree with the header (no compute calls — there is no GPU here)."""
import ctypes
import os

import pytest

from _header_parse import parse_header

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as ge
    ge.build()
    from tic_b200 import capi
    assert os.path.exists(capi.LIB_PATH)
    return capi.LIB_PATH


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    decl = parse_header()
    assert len(decl) >= 30
    missing = [n for n in decl if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_prototypes_match_header(lib_path):
    from tic_b200 import capi
    decl = parse_header()
    assert sorted(decl) == capi.exported_names()
    for name, (codes, _ret) in decl.items():
        assert capi._SIGS[name][0] == codes, name


def test_version_and_error_string_without_gpu(lib_path):
    from tic_b200 import capi
    lib = capi.load()
    assert lib.tic_version() >= 100
    assert isinstance(capi.last_error(), str)
    assert lib.tic_itc_row_parts(1000) == 32 and lib.tic_itc_col_parts(1000) == 8   # 64-wide tiles below 2048 columns
    assert lib.tic_itc_row_parts(4096) == 32                                           # 256-wide tiles above
    assert lib.tic_ce_bidir_workspace_bytes(128) > 0
    # argument validation happens before any CUDA call, so error codes are observable on a CPU-only box
    with pytest.raises(capi.TicError):
        capi.call("tic_gemm_bf16", None, None, 8, 0, None, None, 8, 0, None, None, 8, 0, 16, 16, 16, 1.0, None, 0, 0, None)
    assert "null pointer" in capi.last_error()


def test_missing_library_fails_loudly(monkeypatch, lib_path):
    from tic_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", "/nonexistent/libtic_b200.so")
    with pytest.raises(capi.TicError):
        capi.load()
