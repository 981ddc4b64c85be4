"""BN254 scalar field element codecs (packages/ff_utils/src/fp.rs).

In memory an element is the reference's `Fp([u64; 4])`: Montgomery form (x * 2^256 mod p), four
little-endian u64 limbs.  Vectors are numpy arrays of shape (n, 4), dtype uint64.
"""
import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # fp.rs:9
GENERATOR = 7                                                                         # fp.rs:10
R = 1 << 256
R_INV = pow(R, -1, P)
TWO_ADICITY = 28


def root_of_unity(log_n):
    """7^((p-1)/2^log_n) as a canonical int (prove.rs:71-82)"""
    assert 0 <= log_n <= TWO_ADICITY
    return pow(GENERATOR, (P - 1) >> log_n, P)


def to_mont(values):
    """canonical ints -> (n, 4) uint64 Montgomery limbs"""
    buf = b"".join(((int(v) % P) * R % P).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype="<u8").reshape(-1, 4).copy()


def from_mont(arr):
    """(n, 4) uint64 Montgomery limbs -> list of canonical ints"""
    b = np.ascontiguousarray(arr, dtype="<u8").tobytes()
    return [int.from_bytes(b[i:i + 32], "little") * R_INV % P for i in range(0, len(b), 32)]


def mont_scalar(x):
    return to_mont([x])[0].copy()


def to_bytes_le(x):
    """fp.rs:39-43"""
    return (int(x) % P).to_bytes(32, "little")


def from_bytes_le(b):
    """fp.rs:74-76: integer value of any length, reduced mod p"""
    return int.from_bytes(bytes(b), "little") % P
