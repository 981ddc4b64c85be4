"""Mirror of packages/fri/src/poly_utils.rs items that sit on the hot path."""
import numpy as np

from ._lib import _ptr, default_context


def multi_inv(values, ctx=None):
    """poly_utils.rs:38-70: element-wise inverse, zero maps to zero"""
    ctx = ctx or default_context()
    v = np.array(values, dtype=np.uint64, copy=True).reshape(-1, 4)
    ctx.check(ctx.lib.sb_batch_inverse(ctx.h, _ptr(v), v.shape[0]))
    return v
