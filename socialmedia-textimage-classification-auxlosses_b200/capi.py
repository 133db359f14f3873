"""ctypes binding of the C ABI declared in include/tic_b200.h (libtic_b200.so, built by build.py).

There is NO fallback: if the library is missing or a call fails, a TicError is raised.  Pointers are passed as
integers (`tensor.data_ptr()`), the stream as `torch.cuda.current_stream().cuda_stream`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libtic_b200.so")


class TicError(RuntimeError):
    pass


# signature codes: p = pointer (void*), i = int, l = int64_t, f = float
_SIGS = {
    "tic_last_error_string": ("", ctypes.c_char_p),
    "tic_version": ("", ctypes.c_int),
    "tic_sm_count": ("", ctypes.c_int),
    "tic_set_pdl": ("i", ctypes.c_int),
    "tic_gemm_bf16": ("pplipplippliiiifpiip", ctypes.c_int),
    "tic_gemm_rowss_parts": ("i", ctypes.c_int),
    "tic_gemm_plan": ("iiiiippp", ctypes.c_int),
    "tic_gemm_bf16_rowss": ("pplipplippliiiifpipp", ctypes.c_int),
    "tic_gemm_bf16_simt": ("pliplipliiiifpip", ctypes.c_int),
    "tic_row_rnorm_bf16": ("ppliipplp", ctypes.c_int),
    "tic_itc_row_parts": ("i", ctypes.c_int),
    "tic_itc_q_parts": ("i", ctypes.c_int),
    "tic_itc_col_parts": ("i", ctypes.c_int),
    "tic_itc_fwd": ("pplpplppiiiiffpppplpipippiippp", ctypes.c_int),
    "tic_itc_fused_small_ok": ("ii", ctypes.c_int),
    "tic_debug_set_trace": ("p", ctypes.c_int),
    "tic_itc_fwd_bwd_small": ("pplpplppiiiifpppplpipippfplplppp", ctypes.c_int),
    "tic_itc_pick": ("pplpplppiiiiffppppp", ctypes.c_int),
    "tic_reduce_parts": ("piipp", ctypes.c_int),
    "tic_itc_lse_loss": ("pipipiiifppppp", ctypes.c_int),
    "tic_itc_lse_rows_workspace_bytes": ("i", ctypes.c_int64),
    "tic_itc_lse_rows": ("ppiipfpppppp", ctypes.c_int),
    "tic_itc_lse_rows_push": ("ppiipfppppppiilllpp", ctypes.c_int),
    "tic_peer_handle_bytes": ("", ctypes.c_int),
    "tic_peer_alloc": ("lp", ctypes.c_int),
    "tic_peer_free": ("p", ctypes.c_int),
    "tic_peer_export": ("pp", ctypes.c_int),
    "tic_peer_open": ("pp", ctypes.c_int),
    "tic_peer_close": ("p", ctypes.c_int),
    "tic_peer_exchange": ("piilpippppp", ctypes.c_int),
    "tic_peer_pull": ("piipppippppip", ctypes.c_int),
    "tic_peer_push": ("piillppippppp", ctypes.c_int),
    "tic_peer_signal": ("piilpp", ctypes.c_int),
    "tic_itc_bwd_g": ("pplpplppppiiiffplplpppipifpppiip", ctypes.c_int),
    "tic_itc_ds_operands": ("pliipppplpplp", ctypes.c_int),
    "tic_itc_grad_finalize": ("plpplppplpiiffplpplpipp", ctypes.c_int),
    "tic_ce_bidir_workspace_bytes": ("i", ctypes.c_int64),
    "tic_ce_bidir_fwd": ("plippppp", ctypes.c_int),
    "tic_ce_bidir_bwd": ("plipppplp", ctypes.c_int),
    "tic_itm_sample": ("ppiiplfpppp", ctypes.c_int),
    "tic_itm_hard_locate": ("ppiiipippppp", ctypes.c_int),
    "tic_gather_rows": ("plpllpip", ctypes.c_int),
    "tic_itm_sample_gather": ("ppiiplfppplppppp", ctypes.c_int),
    "tic_pack_cls_pairs": ("plpliipplppp", ctypes.c_int),
    "tic_unpack_cls_grad": ("plpliipplp", ctypes.c_int),
    "tic_heads_fwd_bwd": ("pliiiippppppppfffppppplplppppipppplpp", ctypes.c_int),
    "tic_fusion_pair_grad": ("ppliiippppplp", ctypes.c_int),
    "tic_heads_wgrad": ("pliiiippfppppp", ctypes.c_int),
    "tic_attn_pool_fwd": ("pllpliiiifpplplplp", ctypes.c_int),
    "tic_attn_pool_bwd": ("pllplplpliiiifpplp", ctypes.c_int),
    "tic_aspect_fwd": ("plpliippplpp", ctypes.c_int),
    "tic_aspect_bwd": ("plpliippplpplplppp", ctypes.c_int),
    "tic_gmu_gate_fwd": ("plppliipplp", ctypes.c_int),
    "tic_gmu_gate_bwd": ("plpplpliipppplplp", ctypes.c_int),
    "tic_refresh_weights": ("ippppppppppipip", ctypes.c_int),
    "tic_cast_f32_to_bf16": ("plpliip", ctypes.c_int),
    "tic_cast_bf16_to_f32": ("plpliip", ctypes.c_int),
    "tic_colsum_bf16": ("pliipp", ctypes.c_int),
    "tic_colsum_bf16_pair": ("ppliipp", ctypes.c_int),
    "tic_loss_mix": ("ppiffiipp", ctypes.c_int),
    "tic_eval_state_words": ("i", ctypes.c_int),
    "tic_eval_accumulate": ("plplpiipppppp", ctypes.c_int),
    "tic_metrics_from_confusion": ("pipp", ctypes.c_int),
}
_CT = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_int64, "f": ctypes.c_float}

_lib = None


def load(build_if_missing: bool = False):
    """Loads libtic_b200.so (once). Raises TicError if it is absent — the product path has no CPU/torch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _b
            _b.build()
        else:
            raise TicError("libtic_b200.so not found at %s — run `python __graft_entry__.py build` "
                           "(there is no fallback path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (sig, res) in _SIGS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise TicError("libtic_b200.so does not export %s (stale build?)" % name) from e
        fn.argtypes = [_CT[c] for c in sig]
        fn.restype = res
    _lib = lib
    return lib


def exported_names():
    return sorted(_SIGS)


def last_error() -> str:
    return load().tic_last_error_string().decode()


call_hook = None   # optional observer `hook(name)` of every C-ABI call (bench.py counts kernel launches with it)


def call(name: str, *args):
    """Calls an int-returning entry point and raises TicError (with the library's message) on a non-zero code."""
    lib = load()
    if call_hook is not None:
        call_hook(name)
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise TicError("%s failed (rc=%d): %s" % (name, rc, lib.tic_last_error_string().decode()))
    return rc


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
