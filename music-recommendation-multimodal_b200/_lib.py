"""ctypes binding of libtt_b200.so — the only door to the device code.

There is deliberately no fallback: if the shared library is missing or a launcher
returns non-zero, a TTError is raised.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int32, c_uint32, c_uint64, c_void_p
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libtt_b200.so"
_lib = None


class TTError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    """Mirror of ``tt_gemm_args`` (include/tt_b200.h)."""

    _fields_ = [
        ("A", c_void_p), ("B", c_void_p),
        ("lda", c_int32), ("ldb", c_int32),
        ("a_mn", c_int32), ("b_mn", c_int32),
        ("M", c_int32), ("N", c_int32), ("K", c_int32),
        ("alpha", c_float),
        ("bias", c_void_p),
        ("relu", c_int32),
        ("drop_p", c_float),
        ("drop_seed", c_uint64),
        ("drop_site", c_uint32),
        ("gate", c_void_p),
        ("ld_gate", c_int32),
        ("gate_scale", c_float),
        ("residual", c_void_p),
        ("ld_res", c_int32),
        ("out_f32", c_void_p),
        ("ld_f32", c_int32),
        ("out_bf16", c_void_p),
        ("ld_bf16", c_int32),
        ("accumulate", c_int32),
        ("k_splits", c_int32),
        ("block_n", c_int32),
    ]


def _declare(l):
    l.tt_last_error.restype = c_char_p
    l.tt_last_error.argtypes = []
    l.tt_version.restype = c_int32
    l.tt_num_sms.restype = c_int32
    l.tt_gemm_bf16.restype = c_int32
    l.tt_gemm_bf16.argtypes = [ctypes.POINTER(GemmArgs), c_void_p]
    l.tt_attn_causal_fwd.restype = c_int32
    l.tt_attn_causal_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float,
                                     c_uint64, c_uint32, c_void_p]
    l.tt_attn_causal_bwd.restype = c_int32
    l.tt_attn_causal_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                     c_int32, c_float, c_uint64, c_uint32, c_void_p]


def lib():
    """Load (once) and return the shared library; fail loudly if it is not built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise TTError(
                f"{_LIB_PATH} not found: build it with `python __graft_entry__.py` "
                "(there is no CPU or PyTorch fallback for the device path)")
        l = ctypes.CDLL(str(_LIB_PATH))
        _declare(l)
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().tt_last_error().decode("utf-8", "replace")
        raise TTError(f"{what} failed (code {rc}): {msg}")
