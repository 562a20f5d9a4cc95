"""ctypes binding of ``libbitdance_b200.so`` (the C ABI declared in ``include/bitdance_b200.h``).

There is no fallback: if the shared library is missing or a call fails, this module raises. PyTorch is used only
for device memory and streams; every tensor is passed as a raw device pointer.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# BD_LIB_PATH: measurement only (A/B of two builds of the library on the same box, scripts/head_ab.py)
LIB_PATH = os.environ.get("BD_LIB_PATH") or os.path.join(_HERE, "_C", "libbitdance_b200.so")


class BitDanceNativeError(RuntimeError):
    pass


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p),
        ("gate", C.c_void_p),
        ("res", C.c_void_p),
        ("out", C.c_void_p),
        ("ld_gate", C.c_int64),
        ("ld_res", C.c_int64),
        ("ld_out", C.c_int64),
        ("act", C.c_int),
        ("swiglu", C.c_int),
        ("res_f32", C.c_int),
        ("out_f32", C.c_int),
        ("res_row_mod", C.c_int),
    ]


_lib = None


def load() -> C.CDLL:
    """Load the native library (building is ``__graft_entry__.build()``'s job; never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BitDanceNativeError(
            f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU / eager fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.bd_strerror.restype = C.c_char_p
    lib.bd_strerror.argtypes = [C.c_int]
    lib.bd_gemm_workspace_bytes.restype = C.c_size_t
    lib.bd_gemm_workspace_bytes.argtypes = [C.c_int] * 5
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        lib = load()
        msg = lib.bd_strerror(status).decode()
        extra = ""
        if status == -3:
            extra = f" (cudaError {lib.bd_last_cuda_error()})"
        raise BitDanceNativeError(f"{what or 'bitdance_b200 call'} failed: {msg}{extra}")


def stream_ptr(stream: torch.cuda.Stream | None = None) -> C.c_void_p:
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def ptr(t: torch.Tensor | None) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise BitDanceNativeError("bitdance_b200 ops need CUDA tensors (there is no CPU fallback)")


_NOTED = set()


def note(msg: str):
    """One log line per distinct reason when a call leaves a persistent single-kernel path (never silently)."""
    if msg not in _NOTED:
        _NOTED.add(msg)
        import logging
        logging.getLogger("bitdance_b200").warning(msg)
