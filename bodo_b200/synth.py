"""Synthetic workload definition shared by bench.py, the tests and the oracle.

key = mix64(row ^ seed*PHI) % n_groups ; val = (mix64(row ^ (seed+1)*C2) % 1000) - 500 (int64) or a
uniform double in [0,1).  The device generator is b200_synth_fill (csrc/misc.cu); numpy_fill is the
bit-identical host mirror used for small cases.
"""

from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import ffi

_M = (1 << 64) - 1


def _mix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def numpy_fill(row_start: int, n: int, n_groups: int, seed: int, float_vals: bool = False):
    r = np.arange(row_start, row_start + n, dtype=np.uint64)
    sk = np.uint64((seed * 0x9E3779B97F4A7C15) & _M)
    sv = np.uint64(((seed + 1) * 0xD1B54A32D192ED03) & _M)
    keys = (_mix64(r ^ sk) % np.uint64(n_groups)).astype(np.int64)
    m = _mix64(r ^ sv)
    if float_vals:
        vals = (m >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    else:
        vals = (m % np.uint64(1000)).astype(np.int64) - 500
    return keys, vals


def device_fill(keys_t, vals_t, row_start: int, n_groups: int, seed: int, stream: int = 0):
    """Fill torch CUDA tensors (int64 keys; int64 or float64 vals) in place."""
    import torch

    ct = 6 if vals_t is not None and vals_t.dtype == torch.float64 else 4
    n = keys_t.numel() if keys_t is not None else vals_t.numel()
    _lib.check(_lib.lib().b200_synth_fill(ffi.cast("void*", keys_t.data_ptr() if keys_t is not None else 0),
                                          ffi.cast("void*", vals_t.data_ptr() if vals_t is not None else 0), row_start, n,
                                          n_groups, seed, ct, ffi.cast("void*", stream)), "b200_synth_fill")
