"""numpy twin of the counter-based generator in csrc/synth.cu (bit-identical values), plus the
feasible-by-construction instance recipe of SURVEY.md §8d.  Data generation only - no solver arithmetic."""
from __future__ import annotations

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)
_C = np.uint64(0xD1B54A32D192ED03)


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def uniform_matrix(n_row, n_col, seed, scale, row_offset=0, dtype=np.float32, rows=None, cols=None):
    """Column-major (Fortran-ordered) matrix with A[r,c] = scale * base(seed, row_offset+r, c), base in (-1, 1)."""
    with np.errstate(over="ignore"):
        r = (np.arange(n_row, dtype=np.uint64) if rows is None else np.asarray(rows, dtype=np.uint64)) + np.uint64(row_offset)
        c = np.arange(n_col, dtype=np.uint64) if cols is None else np.asarray(cols, dtype=np.uint64)
        rk = _mix64(np.uint64(seed) * _G + r)                       # per-row key
        key = rk[:, None] ^ (c[None, :] * _C)
        u24 = (_mix64(key) >> np.uint64(40)).astype(np.int64)
    base = (2 * u24 + 1).astype(np.float64) * (1.0 / 16777216.0) - 1.0
    a = np.asfortranarray(base.astype(dtype))
    a *= dtype(scale) if isinstance(dtype, type) else np.dtype(dtype).type(scale)
    return a
