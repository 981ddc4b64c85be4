"""stark_pure_rust_b200 -- B200 (sm_100a) backend for the stark-pure-rust hot path.

Host-side mirror of the reference's crate APIs on top of the C ABI in include/stark_b200.h
(libstark_b200.so, built in-tree by stark_pure_rust_b200.build):

    fft        best_fft / inv_best_fft / expand_root_of_unity   (packages/fri/src/fft.rs)
    merkle     MerkleProofInPlace, Proof, verify_multi_branch    (packages/commitment/src/*.rs)
    fri        prove_low_degree, FriProof layout                 (packages/fri/src/fri.rs)
    poly_utils multi_inv                                         (packages/fri/src/poly_utils.rs)
    prove      mk_r1cs_proof (device-resident prover)               (packages/r1cs-stark/src/prove.rs)
    ext        the LDE -> commit -> FRI chain on device-resident, coset-major columns, one or several GPUs
    utils      blake, get_pseudorandom_indices                   (packages/fri/src/utils.rs)
    field      Fp <-> Montgomery-limb conversions                (packages/ff_utils/src/fp.rs)

There is no CPU fallback: importing works anywhere, but creating a Context without the built
library or without a B200 raises.
"""
from . import _lib  # noqa: F401
from ._lib import Context, StarkB200Error, default_context, library_path  # noqa: F401
from . import field, fft, merkle, fri, poly_utils, prove, utils, ext  # noqa: F401

__all__ = ["Context", "StarkB200Error", "default_context", "library_path", "field", "fft", "merkle", "fri",
           "poly_utils", "utils", "prove", "ext"]
