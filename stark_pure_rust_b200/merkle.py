"""Mirror of packages/commitment/src/{merkle_tree.rs, merkle_proof_in_place.rs, pallarel_merkle_tree.rs}: H = BlakeDigest
(the digest the binary fixes, r1cs-stark/src/main.rs:9) or H = PoseidonDigest (commitment/src/poseidon.rs) per tree."""
import ctypes as C
from dataclasses import dataclass
from typing import List

import numpy as np

from . import utils
from ._lib import _ptr, default_context


@dataclass
class Proof:
    """merkle_tree.rs:15-18: raw leaf bytes + sibling digests, leaf level first, root excluded"""
    leaf: bytes
    nodes: List[bytes]

    def validate(self, root, index, digest="blake"):
        """merkle_tree.rs:25-43 (generic in the digest H)"""
        H = _digest(digest)
        h = H(self.leaf)
        for node in self.nodes:
            h = H(node + h) if index & 1 else H(h + node)
            index >>= 1
        return h == bytes(root)


def _digest(name):
    if name == "blake":
        return utils.blake
    if name == "poseidon":
        return utils.poseidon
    raise ValueError("digest must be 'blake' or 'poseidon'")


def verify_multi_branch(root, indices, proofs, digest="blake"):
    """merkle_tree.rs:46-58"""
    return all(p.validate(root, i, digest) for i, p in zip(indices, proofs))


def poseidon_hash_many(messages, ctx=None):
    """PoseidonDigest::hash (poseidon.rs:30-63) of equally long messages, one device thread per message"""
    ctx = ctx or default_context()
    msgs = [bytes(m) for m in messages]
    if not msgs:
        return []
    lb = len(msgs[0])
    assert all(len(m) == lb for m in msgs), "messages must have equal length"
    flat = np.frombuffer(b"".join(msgs), dtype=np.uint8)
    out = np.empty(32 * len(msgs), dtype=np.uint8)
    ctx.check(ctx.lib.sb_poseidon_hash(ctx.h, _ptr(flat) if flat.size else None, lb, len(msgs), _ptr(out)))
    o = out.tobytes()
    return [o[32 * i:32 * i + 32] for i in range(len(msgs))]


class MerkleProofInPlace:
    """merkle_proof_in_place.rs:9-51.  Leaves are byte strings of equal length (Vec<Vec<u8>>).  The tree
    is built once on the first gen_proofs/get_root after update() and kept in HBM; later gen_proofs
    calls are gathers (the reference rebuilds it every time, :106-206)."""

    def __init__(self, ctx=None, digest="blake"):
        self.ctx = ctx or default_context()
        _digest(digest)
        self.digest = digest
        self._leaves = None
        self._tree = None
        self._root = b""            # H::default(), :19
        self._width = 0

    def width(self):
        return self._width

    def get_root(self):
        return self._root

    def update(self, leaves):
        leaves = [bytes(l) for l in leaves]
        self._free()
        self._leaves = leaves
        self._width = len(leaves)

    def _build(self):
        if self._tree is not None:
            return
        leaves = self._leaves
        n = len(leaves)
        lb = len(leaves[0]) if n else 0
        assert all(len(l) == lb for l in leaves), "leaves must have equal length"
        flat = np.frombuffer(b"".join(leaves), dtype=np.uint8)
        root = np.empty(32, dtype=np.uint8)
        t = C.c_void_p()
        commit = self.ctx.lib.sb_merkle_commit if self.digest == "blake" else self.ctx.lib.sb_merkle_commit_poseidon
        self.ctx.check(commit(self.ctx.h, _ptr(flat) if flat.size else None, lb, n, _ptr(root), C.byref(t)))
        self._tree = t
        self._root = root.tobytes()

    def gen_proofs(self, indices):
        self._build()
        idx = np.asarray(list(indices), dtype=np.uint64)
        q = idx.size
        if q == 0:
            return []
        lb = int(self.ctx.lib.sb_tree_leaf_bytes(self._tree))
        depth = (self._width - 1).bit_length()
        leaves = np.empty(q * lb, dtype=np.uint8)
        nodes = np.empty(q * depth * 32, dtype=np.uint8)
        self.ctx.check(self.ctx.lib.sb_merkle_open(self.ctx.h, self._tree, idx.ctypes.data_as(C.POINTER(C.c_size_t)), q,
                                                   _ptr(leaves) if leaves.size else None, _ptr(nodes) if nodes.size else None))
        lv, nd = leaves.tobytes(), nodes.tobytes()
        return [Proof(lv[i * lb:(i + 1) * lb], [nd[(i * depth + l) * 32:(i * depth + l + 1) * 32] for l in range(depth)])
                for i in range(q)]

    def _free(self):
        if self._tree is not None:
            self.ctx.lib.sb_tree_free(self.ctx.h, self._tree)
            self._tree = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass


class ParallelMerkleTree(MerkleProofInPlace):
    """pallarel_merkle_tree.rs:13-131: same leaves, nodes, root and proofs as MerkleProofInPlace (the reference's two
    builders differ only in how the CPU work is chunked)"""
