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


def test_gemm_plan_and_eval_state_without_gpu(lib_path, monkeypatch):
    """Launch-shape selection is host arithmetic: observable without a device.  Cluster split-K is opt-in (TIC_CLUSTER_K)."""
    from tic_b200 import capi
    lib = capi.load()

    def plan(M, N, K, ns=0, acc=0):
        bn, ks, kc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        capi.call("tic_gemm_plan", M, N, K, ns, acc, ctypes.byref(bn), ctypes.byref(ks), ctypes.byref(kc))
        return bn.value, ks.value, kc.value

    monkeypatch.delenv("TIC_CLUSTER_K", raising=False)
    assert plan(256, 512, 768) == (64, 1, 1)                  # default: no cluster split-K
    assert plan(512, 768, 4096, 0, 1)[1] > 1                  # long-K weight-gradient GEMM: fp32-atomic split-K
    monkeypatch.setenv("TIC_CLUSTER_K", "96")
    assert plan(256, 512, 768) == (64, 1, 4) and plan(512, 768, 1536) == (64, 1, 2)
    assert plan(4096, 4096, 4096)[2] == 1
    assert lib.tic_eval_state_words(4) == 20 and lib.tic_eval_state_words(0) < 0
    with pytest.raises(capi.TicError):
        capi.call("tic_eval_accumulate", None, 4, None, 0, None, 8, 4, None, None, None, None, None, None)
