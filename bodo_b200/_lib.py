"""cffi (ABI mode) binding of libbodo_b200.so — the only door from Python into the CUDA path.

There is deliberately no fallback: if the shared library is missing or no GPU is visible, every compute
entry point raises. (The reference's GPU tests enforce the same thing with
BODO_GPU_DISABLE_CPU_FALLBACK=1, bodo/tests/conftest.py:971-980.)
"""

from __future__ import annotations

import os
import re

import cffi

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "bodo_b200.h")
LIB_PATH = os.path.join(_HERE, "libbodo_b200.so")

ffi = cffi.FFI()
_lib = None


class B200Error(RuntimeError):
    """Error raised by libbodo_b200 (message from b200_last_error())."""


def _cdef_text() -> str:
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    lines = [ln for ln in src.splitlines() if not ln.strip().startswith("#")]
    text = "\n".join(lines)
    text = text.replace('extern "C" {', "")
    # drop the closing brace of the extern "C" block (the only line that is just "}")
    text = "\n".join(ln for ln in text.splitlines() if ln.strip() != "}")
    return text


ffi.cdef(_cdef_text())


def declared_symbols() -> list[str]:
    """Every function name include/bodo_b200.h declares (used by the symbol-export test)."""
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", _cdef_text())))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} is missing: build it with `python -m bodo_b200.build` (needs nvcc). "
                "bodo_b200 has no CPU fallback for the groupby/join/shuffle path."
            )
        _lib = ffi.dlopen(LIB_PATH)
    return _lib


def check(rc, what: str = ""):
    """Raise B200Error if an int-returning entry point reported failure."""
    if rc is None or (isinstance(rc, int) and rc < 0):
        msg = ffi.string(lib().b200_last_error()).decode()
        raise B200Error(f"{what}: {msg}" if what else msg)
    return rc


def check_ptr(p, what: str = ""):
    if p == ffi.NULL:
        msg = ffi.string(lib().b200_last_error()).decode()
        raise B200Error(f"{what}: {msg}" if what else msg)
    return p


def require_gpu() -> int:
    n = lib().b200_device_count()
    if n <= 0:
        raise B200Error("no CUDA device visible: the bodo_b200 hot path is CUDA-only (no CPU fallback)")
    return n
