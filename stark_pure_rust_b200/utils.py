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


def poseidon(message):
    """commitment/src/poseidon.rs:30-63 `PoseidonDigest::hash` (host side; the batched device version is
    `merkle.poseidon_hash_many`).  ValueError where the reference panics: empty or > 64 bytes, or a 32-byte chunk
    that is not a canonical BLS12-381 scalar."""
    lib = load()
    m = np.frombuffer(bytes(message), dtype=np.uint8)
    out = np.empty(32, dtype=np.uint8)
    if lib.sb_poseidon_hash_host(_ptr(m) if m.size else None, m.size, _ptr(out)) != 0:
        raise ValueError("PoseidonDigest::hash: message must be 1..64 bytes of canonical 32-byte scalars")
    return out.tobytes()


def get_pseudorandom_indices(seed, modulus, count, exclude_multiples_of=0, ctx=None):
    """utils.rs:82-109.  With a context whose extended domain is enabled (sb_set_extended_domain) moduli >= 2^24
    are accepted; the reference asserts modulus < 2^24 (utils.rs:88)."""
    lib = load()
    s = np.frombuffer(bytes(seed), dtype=np.uint8)
    out = np.empty(count, dtype=np.uint32)
    if ctx is not None:
        rc = lib.sb_pseudorandom_indices_ctx(ctx.h, _ptr(s) if s.size else None, s.size, modulus, count, exclude_multiples_of, _ptr(out))
    else:
        rc = lib.sb_pseudorandom_indices(_ptr(s) if s.size else None, s.size, modulus, count, exclude_multiples_of, _ptr(out))
    if rc != 0:
        raise StarkB200Error(rc, "get_pseudorandom_indices: the reference panics on these arguments")
    return [int(x) for x in out]
