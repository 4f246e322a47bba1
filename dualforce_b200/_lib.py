"""ctypes binding of ``libmova_b200.so`` (the C ABI declared in ``include/mova_b200.h``).

There is no CPU or PyTorch fallback: if the shared library is missing, or the current device is not an
sm_100 part, every entry point raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"``
(or ``make -C dualforce_b200/csrc``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmova_b200.so")
ABI_VERSION = 6

# name -> (restype, argtypes); mirrors include/mova_b200.h one to one
SIGNATURES = {
    "mova_b200_abi_version": (c_int, []),
    "mova_b200_last_error": (c_char_p, []),
    "mova_b200_debug_record": (c_void_p, []),
    "mova_b200_device_check": (c_int, [c_int]),
    "mova_b200_linear": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p,
         c_int64, c_void_p, c_float, c_int, c_void_p],
    ),
    "mova_b200_linear_ex": (
        c_int,
        [c_void_p, c_int64, c_int, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_int,
         c_int, c_int, c_void_p, c_int64, c_void_p, c_float, c_int, c_void_p],
    ),
    "mova_b200_attn_fwd": (
        c_int,
        [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64,
         c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p],
    ),
    "mova_b200_attn_fwd_variant": (
        c_int,
        [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64,
         c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p, c_void_p],
    ),
    "mova_b200_lse_merge": (
        c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mova_b200_layernorm": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
         c_void_p],
    ),
    "mova_b200_rmsnorm_rope": (
        c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p, c_int, c_void_p]),
    "mova_b200_rmsnorm_rope_seg": (
        c_int,
        [c_void_p, c_int64, c_int, c_int64, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p, c_int,
         c_void_p],
    ),
    "mova_b200_add_to_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mova_b200_patchify": (
        c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "mova_b200_unpatchify": (
        c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mova_b200_sinusoidal": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "mova_b200_cfg_euler": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_void_p]),
    "mova_b200_gemv_f32": (
        c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mova_b200_peer_alloc": (c_int, [c_int64, c_void_p, c_void_p]),
    "mova_b200_peer_open": (c_int, [c_void_p, c_void_p]),
    "mova_b200_peer_close": (c_int, [c_void_p]),
    "mova_b200_peer_free": (c_int, [c_void_p]),
    "mova_b200_peer_memops_supported": (c_int, []),
    "mova_b200_peer_push": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64, c_void_p,
                                    c_void_p]),
    "mova_b200_peer_wait": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p]),
}

EPI_BIAS, EPI_GELU_TANH, EPI_RESIDUAL = 0, 1, 2
ROPE_NONE, ROPE_INTERLEAVED, ROPE_HALF = 0, 1, 2

_lib = None
_checked_devices = set()


class MovaB200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library once; raise loudly when it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MovaB200Error(
            f"{LIB_PATH} not found: the sm_100a extension is not built. Run `make -C dualforce_b200/csrc` "
            "(or __graft_entry__.build()). dualforce_b200 has no CPU/PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header / library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.mova_b200_abi_version()
    if got != ABI_VERSION:
        raise MovaB200Error(f"libmova_b200.so ABI version {got}, python binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().mova_b200_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


LAUNCHES = 0  # kernel launches issued through the C ABI (each entry point enqueues exactly one kernel)
_TIMERS = None  # bench.py sets this to a list; ops.attention then appends (start_event, end_event, B, Sq, Skv, H, D)


def check(rc: int, what: str, launches: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += launches
    if rc != 0:
        raise MovaB200Error(f"{what} failed (code {rc}): {last_error()}")


def require_device(index: int) -> None:
    """Raise unless CUDA device ``index`` can run the sm_100a kernels."""
    if index in _checked_devices:
        return
    rc = load().mova_b200_device_check(int(index))
    if rc != 0:
        raise MovaB200Error(f"mova_b200_device_check({index}) failed (code {rc}): {last_error()}")
    _checked_devices.add(index)
