"""ctypes binding of libmae_b200.so (include/mae_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing
(``python -m mae_b200.build`` / ``__graft_entry__.build()`` not run) importing an op raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAE_B200_LIB") or os.path.join(_PKG, "libmae_b200.so")   # env: tuning builds only

ABI_VERSION = 1

_P, _I64, _I, _F = c_void_p, c_int64, c_int, c_float

# name -> argtypes, exactly as declared in include/mae_b200.h
SIGNATURES = {
    "mae_reparam_kl_fwd": [_P, _I64, _P, _I64, _P, _I64, _I64, _I64, _P, _P, _P],
    "mae_reparam_kl_bwd": [_P, _I64, _P, _I64, _P, _P, _P, _I64, _I64, _I64, _P, _P, _I, _P],
    "mae_reparam_reg_bwd": [_P, _I64, _P, _I64, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _P, _P, _I, _P],
    "mae_mpd_prepare": [_P, _I64, _P, _I64, _I64, _I64, _P, _I64, _P, _P, _P],
    "mae_mpd_fwd": [_P, _I64, _I64, _I64, _I64, _I64, _I, _P, _P, _P],
    "mae_mpd_finalize": [_P, _I64, _P, _P],
    "mae_mpd_bwd": [_P, _I64, _P, _I64, _P, _I64, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "mae_mpd_loss_bwd": [_P, _I64, _P, _I64, _P, _I64, _I64, _I64, _I64, _I64, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P, _I, _P],
    "mae_kl_mpd_loss_fused": [_P, _I64, _P, _I64, _I64, _I64, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P],
    "mae_bernoulli_fwd": [_P, _P, _I64, _I64, _I64, _P, _P],
    "mae_bernoulli_bwd": [_P, _P, _P, _I64, _F, _I64, _I64, _I64, _P, _P],
    "mae_dmol_fwd": [_P, _P, _I64, _I64, _I64, _I64, _P, _P, _I64, _P],
    "mae_dmol_bwd": [_P, _P, _P, _I64, _F, _I64, _I64, _I64, _I64, _P, _P],
    "mae_log_prob_prior_fwd": [_P, _P, _I64, _I64, _P, _P],
    "mae_log_prob_prior_bwd": [_P, _P, _I64, _I64, _P, _P],
    "mae_log_prob_posterior_fwd": [_P, _P, _P, _I64, _P, _I64, _I64, _I64, _I64, _P, _P],
    "mae_log_prob_posterior_bwd": [_P, _P, _P, _I64, _P, _I64, _I64, _I64, _I64, _P, _P, _P, _P],
    "mae_iw_nll": [_P, _P, _P, _P, _P, _P, _I64, _P, _I64, _I64, _I64, _I64, _F, _P, _P, _P, _P, _P, _I64, _P],
    "mae_loss_assemble_fwd": [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _F, _F, _F, _F, _P, _P, _P],
    "mae_loss_assemble_bwd": [_P, _P, _P, _I64, _I64, _I64, _F, _F, _F, _F, _P, _P, _P, _P, _P],
    "mae_dmol_fwd_bwd": [_P, _P, _F, _I64, _I64, _I64, _I64, _P, _P, _P],
    "mae_scale_inplace": [_P, _I64, _P, _P],
    "mae_dmol_head": [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I, _F, _P, _P, _P, _I64, _P],
    "mae_comm_alloc": [_I64, _P],
    "mae_comm_free": [_P],
    "mae_comm_get_handle": [_P, _P],
    "mae_comm_open_handle": [_P, _P],
    "mae_comm_close_handle": [_P],
    "mae_comm_error": [_P, _P],
    "mae_allgather_rows": [_P, _I64, _P, _I64, _I64, _I64, _P, _P],
    "mae_kl_mpd_rows_fused": [_P, _P, _I64, _I64, _I64, _I64, _I64, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P, _I64, _P],
    "mae_microbench_ffma": [_I64, _I64, _P, _P, _P],
    "mae_microbench_mufu": [_I, _I64, _I64, _P, _P, _P],
}
# entry points that do not return a status code
PLAIN = {
    "mae_abi_version": (c_int, []),
    "mae_set_pdl": (c_int, [c_int]),
    "mae_set_mpd_tensor_cores": (c_int, [c_int]),
    "mae_strerror": (c_char_p, [c_int]),
    "mae_device_info": (c_int, [_P, _P, _P]),
    "mae_mpd_workspace_bytes": (c_int64, [_I64, _I64, _I]),
    "mae_mpd_workspace_bytes_rows": (c_int64, [_I64, _I64, _I64, _I64]),
    "mae_kl_mpd_loss_fused_supported": (c_int, [_I64, _I64]),
    "mae_dmol_workspace_bytes": (c_int64, [_I64, _I64]),
    "mae_iw_nll_workspace_bytes": (c_int64, [_I64, _I64]),
    "mae_comm_buffer_bytes": (c_int64, [_I64, _I64]),
    "mae_rows_comm_bytes": (c_int64, [_I64, _I64]),
    "mae_rows_scratch_bytes": (c_int64, []),
    "mae_kl_mpd_rows_fused_supported": (c_int, [_I64, _I64, _I64]),
    "mae_dmol_err_chunks": (c_int64, [_I64, _I64, _I64, _I64]),
    "mae_dmol_head_supported": (c_int, [_I64, _I64, _I64]),
    "mae_dmol_head_workspace_bytes": (c_int64, [_I64, _I64]),
}
ALL_SYMBOLS = sorted(list(SIGNATURES) + list(PLAIN))

_lib = None


class MaeLibraryError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load (once) and type the shared library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MaeLibraryError(
            "libmae_b200.so not found at %s: build it with `python -m mae_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = args
    for name, (res, args) in PLAIN.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    got = lib.mae_abi_version()
    if got != ABI_VERSION:
        raise MaeLibraryError("libmae_b200.so ABI version %d, python layer expects %d: rebuild" % (got, ABI_VERSION))
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mae_strerror(rc)
        raise MaeLibraryError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
