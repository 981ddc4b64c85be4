"""Mirror of r1cs-stark/src/prove.rs::mk_r1cs_proof on the GPU backend (device-resident pipeline)."""
import ctypes as C

import numpy as np

from ._lib import _ptr, default_context


class SbTrace(C.Structure):
    """sb_trace (include/stark_b200.h)"""
    _fields_ = [("original_steps", C.c_size_t),
                ("witness_trace", C.c_void_p), ("computational_trace", C.c_void_p), ("coefficients", C.c_void_p),
                ("flag0", C.c_void_p), ("flag1", C.c_void_p), ("flag2", C.c_void_p),
                ("permuted_indices", C.c_void_p),
                ("n_public", C.c_size_t), ("public_wires", C.c_void_p),
                ("n_pfi", C.c_size_t), ("pfi_k", C.c_void_p), ("pfi_w", C.c_void_p)]


def mk_r1cs_proof(witness_trace, computational_trace, public_wires, public_first_indices, permuted_indices, coefficients,
                  flag0, flag1, flag2, ctx=None, return_stages=False):
    """prove.rs:14-26 argument order (n_constraints / n_wires only feed an assert there and are dropped).
    Vectors are (n, 4) uint64 Montgomery arrays; returns the serde_json text of StarkProof (run.rs:549)."""
    ctx = ctx or default_context()
    keep = []

    def fp(a):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        keep.append(a)
        return a

    def sz(a):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
        keep.append(a)
        return a

    w, c, k, f0, f1, f2, pw = fp(witness_trace), fp(computational_trace), fp(coefficients), fp(flag0), fp(flag1), fp(flag2), fp(public_wires)
    os_ = k.shape[0]
    assert w.shape[0] == os_ and c.shape[0] == os_                       # prove.rs:34-35
    perm = sz(permuted_indices)
    pfi = list(public_first_indices)
    pk, pwi = sz([a for a, _ in pfi]), sz([b for _, b in pfi])
    t = SbTrace(os_, w.ctypes.data, c.ctypes.data, k.ctypes.data, f0.ctypes.data, f1.ctypes.data, f2.ctypes.data,
                perm.ctypes.data, pw.shape[0], pw.ctypes.data if pw.size else None, len(pfi),
                pk.ctypes.data if pk.size else None, pwi.ctypes.data if pwi.size else None)
    h = C.c_void_p()
    ctx.check(ctx.lib.sb_prove_r1cs(ctx.h, C.byref(t), C.byref(h)))
    try:
        n = C.c_size_t()
        s = ctx.lib.sb_stark_proof_json(h, C.byref(n))
        text = C.string_at(s, n.value).decode()
        ctx.lib.sb_free_string(s)
        if return_stages:
            ms = (C.c_double * 5)()
            ctx.lib.sb_stark_proof_stage_ms(h, ms)
            return text, list(ms)
        return text
    finally:
        ctx.lib.sb_stark_proof_free(h)


def prove_with_file_path(r1cs_path, wtns_path, proof_path, ctx=None):
    """run.rs:528-554.  Returns the stage times in ms: [LDE, m_tree, FRI, rest, prove wall, front end, JSON]."""
    ctx = ctx or default_context()
    ms = (C.c_double * 7)()
    ctx.check(ctx.lib.sb_prove_files(ctx.h, str(r1cs_path).encode(), str(wtns_path).encode(),
                                     str(proof_path).encode() if proof_path else None, ms))
    return list(ms)


def verify_with_file_path(r1cs_path, wtns_path, proof_path, ctx=None):
    """run.rs:556-590.  Returns [front end + parse ms, verify ms]; raises StarkB200Error(SB_ERR_VERIFY) when the proof is
    rejected (the reference panics)."""
    ctx = ctx or default_context()
    ms = (C.c_double * 2)()
    ctx.check(ctx.lib.sb_verify_files(ctx.h, str(r1cs_path).encode(), str(wtns_path).encode(), str(proof_path).encode(), ms))
    return list(ms)


def trace_from_files(r1cs_path, wtns_path):
    """the host front end alone (run.rs:109-308, :390-419 and the two readers): the arguments of mk_r1cs_proof as numpy
    copies.  Needs no GPU."""
    from ._lib import StarkB200Error, load
    lib = load()
    h, view = C.c_void_p(), C.c_void_p()
    rc = lib.sb_trace_from_files(str(r1cs_path).encode(), str(wtns_path).encode(), C.byref(h), C.byref(view))
    if rc != 0:
        raise StarkB200Error(rc, "cannot build the trace of %s / %s" % (r1cs_path, wtns_path))
    try:
        t = C.cast(view, C.POINTER(SbTrace)).contents
        n = t.original_steps

        def fp(p, cnt):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(cnt, 4)).copy() if cnt else np.zeros((0, 4), dtype=np.uint64)

        def sz(p, cnt):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_size_t)), shape=(cnt,)).copy() if cnt else np.zeros(0, dtype=np.uint64)

        return {"original_steps": n, "witness_trace": fp(t.witness_trace, n), "computational_trace": fp(t.computational_trace, n),
                "coefficients": fp(t.coefficients, n), "flag0": fp(t.flag0, n), "flag1": fp(t.flag1, n), "flag2": fp(t.flag2, n),
                "permuted_indices": sz(t.permuted_indices, n), "public_wires": fp(t.public_wires, t.n_public),
                "pfi_k": sz(t.pfi_k, t.n_pfi), "pfi_w": sz(t.pfi_w, t.n_pfi)}
    finally:
        lib.sb_host_trace_free(h)
