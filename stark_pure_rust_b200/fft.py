"""Mirror of packages/fri/src/fft.rs entry points on the GPU backend.

Vectors are (n, 4) uint64 Montgomery arrays (see field.py); roots are canonical ints or (4,) limb
arrays.  `best_fft` / `inv_best_fft` take and return host arrays exactly like the reference's
by-value Vec<T> (fft.rs:327-379): zero padding to 2^log_order_of_root, natural order in and out.
"""
import ctypes as C

import numpy as np

from . import field
from ._lib import _ptr, default_context


def _root_limbs(root):
    if isinstance(root, (int, np.integer)):
        return field.mont_scalar(int(root))
    a = np.ascontiguousarray(root, dtype=np.uint64).reshape(4)
    return a


def _as_elems(v):
    a = np.ascontiguousarray(v, dtype=np.uint64)
    if a.size == 0:
        a = a.reshape(0, 4)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def best_fft(coefficients, root_of_unity, log_order_of_root, ctx=None, inverse=False):
    """fft.rs:327-357 (inverse=True: fft.rs:359-379).  Raises StarkB200Error(SB_ERR_ARG) where the
    reference's assert fires (len > 2^log, fft.rs:162) and SB_ERR_ROOT for a non-primitive root."""
    ctx = ctx or default_context()
    v = _as_elems(coefficients)
    n = 1 << log_order_of_root
    buf = np.zeros((max(n, v.shape[0]), 4), dtype=np.uint64)
    buf[: v.shape[0]] = v
    root = _root_limbs(root_of_unity)
    ctx.check(ctx.lib.sb_ntt(ctx.h, _ptr(buf), v.shape[0], _ptr(root), log_order_of_root, 1 if inverse else 0))
    return buf[:n]


def inv_best_fft(evaluations, root_of_unity, log_order_of_root, ctx=None):
    """fft.rs:359-379"""
    return best_fft(evaluations, root_of_unity, log_order_of_root, ctx=ctx, inverse=True)


def expand_root_of_unity(root_of_unity, order=None, ctx=None):
    """fft.rs:5-14: [1, w, w^2, ...] until the power wraps to 1.  `order` (the multiplicative order of
    w, a power of two) is derived when not given."""
    ctx = ctx or default_context()
    root = _root_limbs(root_of_unity)
    if order is None:
        w = field.from_mont(root.reshape(1, 4))[0]
        order = 1
        x = w
        while x != 1:
            x = x * x % field.P
            order *= 2
            assert order <= 1 << field.TWO_ADICITY, "root_of_unity has no power-of-two order"
    out = np.empty((order, 4), dtype=np.uint64)
    ctx.check(ctx.lib.sb_powers(ctx.h, _ptr(root), order, _ptr(out)))
    return out


def lde_batch(columns, root_big, log_s, log_ext, ctx=None):
    """the inv_best_fft -> best_fft pairs of prove.rs:100-124: columns (n_cols, col_len, 4) ->
    (n_cols, 2^(log_s+log_ext), 4)"""
    ctx = ctx or default_context()
    cols = np.ascontiguousarray(columns, dtype=np.uint64)
    assert cols.ndim == 3 and cols.shape[2] == 4
    n_cols, col_len = cols.shape[0], cols.shape[1]
    out = np.empty((n_cols, 1 << (log_s + log_ext), 4), dtype=np.uint64)
    root = _root_limbs(root_big)
    ctx.check(ctx.lib.sb_lde_batch(ctx.h, _ptr(cols), n_cols, col_len, _ptr(root), log_s, log_ext, _ptr(out)))
    return out
