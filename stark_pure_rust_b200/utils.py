"""Mirror of packages/fri/src/utils.rs (identical copy: packages/commitment/src/utils.rs): the
Fiat-Shamir helpers that run on the host between kernels."""
import ctypes as C

import numpy as np

from ._lib import StarkB200Error, _ptr, load


def blake(message):
    """utils.rs:5-10: Blake2s-256, unkeyed"""
    lib = load()
    m = np.frombuffer(bytes(message), dtype=np.uint8)
    out = np.empty(32, dtype=np.uint8)
    lib.sb_blake2s(_ptr(m) if m.size else None, m.size, _ptr(out))
    return out.tobytes()


def get_pseudorandom_indices(seed, modulus, count, exclude_multiples_of=0):
    """utils.rs:82-109"""
    lib = load()
    s = np.frombuffer(bytes(seed), dtype=np.uint8)
    out = np.empty(count, dtype=np.uint32)
    rc = lib.sb_pseudorandom_indices(_ptr(s) if s.size else None, s.size, modulus, count, exclude_multiples_of, _ptr(out))
    if rc != 0:
        raise StarkB200Error(rc, "get_pseudorandom_indices: the reference panics on these arguments")
    return [int(x) for x in out]
