"""The LDE -> commit -> FRI chain of mk_r1cs_proof (r1cs-stark/src/prove.rs:100-124, :235-264, :324-332, :367) with the extended
columns kept on the device(s) in coset-major layout (sb_ext_* in include/stark_b200.h).  With a multi-device Context the
columns are sharded by cosets over the GPUs; roots, openings and proofs equal the single-GPU / reference results."""
import ctypes as C

import numpy as np

from . import fri as _fri
from . import merkle as _merkle
from ._lib import _ptr, default_context


class ExtColumns:
    def __init__(self, n_cols, log_s, ctx=None):
        self.ctx = ctx or default_context()
        self.n_cols, self.log_s, self.n = n_cols, log_s, 1 << (log_s + 3)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.sb_ext_create(self.ctx.h, n_cols, log_s, 3, C.byref(h)))
        self.h = h

    def devices(self):
        return self.ctx.lib.sb_ext_devices(self.h)

    def load(self, first, cols):
        """cols: (count, col_len, 4) uint64 Montgomery (pinned or pageable host memory)"""
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        assert cols.ndim == 3 and cols.shape[2] == 4
        self.ctx.check(self.ctx.lib.sb_ext_load(self.ctx.h, self.h, first, cols.shape[0], _ptr(cols), cols.shape[1]))

    def extend(self, first=0, count=None):
        self.ctx.check(self.ctx.lib.sb_ext_extend(self.ctx.h, self.h, first, self.n_cols - first if count is None else count))

    def commit(self, col_ids):
        """-> (root bytes, tree handle); open with .open(tree, indices), release with .free_tree(tree)"""
        ids = (C.c_size_t * len(col_ids))(*col_ids)
        root, t = np.empty(32, dtype=np.uint8), C.c_void_p()
        self.ctx.check(self.ctx.lib.sb_ext_commit(self.ctx.h, self.h, ids, len(col_ids), _ptr(root), C.byref(t)))
        return root.tobytes(), t

    def open(self, tree, indices):
        """gen_proofs(indices): list of merkle.Proof (leaf bytes, nodes leaf level first)"""
        lib = self.ctx.lib
        q, lb, depth = len(indices), lib.sb_tree_leaf_bytes(tree), self.log_s + 3
        idx = (C.c_size_t * q)(*indices)
        leaves, nodes = np.empty(max(q * lb, 1), dtype=np.uint8), np.empty(max(q * depth * 32, 1), dtype=np.uint8)
        self.ctx.check(lib.sb_merkle_open(self.ctx.h, tree, idx, q, _ptr(leaves), _ptr(nodes)))
        lv, nd = leaves.tobytes(), nodes.tobytes()
        return [_merkle.Proof(lv[i * lb:(i + 1) * lb], [nd[(i * depth + l) * 32:(i * depth + l + 1) * 32] for l in range(depth)]) for i in range(q)]

    def free_tree(self, tree):
        self.ctx.lib.sb_tree_free(self.ctx.h, tree)

    def fri_prove(self, col, max_deg_plus_1, excl, tree=None, as_json=False):
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.sb_ext_fri_prove(self.ctx.h, self.h, col, tree, max_deg_plus_1, excl, C.byref(h)))
        try:
            if as_json:
                s = self.ctx.lib.sb_fri_proof_json(h)
                text = C.string_at(s).decode()
                self.ctx.lib.sb_free_string(s)
                return text
            return _fri.unpack_proof(self.ctx, h)
        finally:
            self.ctx.lib.sb_fri_proof_free(h)

    def read(self, col):
        out = np.empty((self.n, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.sb_ext_read(self.ctx.h, self.h, col, _ptr(out)))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.sb_ext_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
