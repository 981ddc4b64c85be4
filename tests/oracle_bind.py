"""ctypes binding of oracle/liboracle.so (the CPU restatement of the reference) -- TEST INFRASTRUCTURE.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")


class FpT(C.Structure):
    _fields_ = [("l", C.c_uint64 * 4)]


class Buf(C.Structure):
    _fields_ = [("p", C.c_void_p), ("len", C.c_size_t), ("cap", C.c_size_t)]


class Branch(C.Structure):
    _fields_ = [("leaf", C.POINTER(C.c_uint8)), ("leaf_bytes", C.c_size_t), ("nodes", C.POINTER(C.c_uint8)), ("depth", C.c_size_t)]


class FriLayer(C.Structure):
    _fields_ = [("is_last", C.c_int), ("root2", C.c_uint8 * 32),
                ("column_branches", C.POINTER(Branch)), ("n_column", C.c_size_t),
                ("poly_branches", C.POINTER(Branch)), ("n_poly", C.c_size_t),
                ("last", C.POINTER(C.c_uint8)), ("n_last", C.c_size_t)]


class FriProof(C.Structure):
    _fields_ = [("layers", C.POINTER(FriLayer)), ("n_layers", C.c_size_t)]


class Trace(C.Structure):
    """orc_trace (oracle.h): run.rs:109-308, 390-419"""
    _fields_ = [("original_steps", C.c_size_t),
                ("witness_trace", C.c_void_p), ("computational_trace", C.c_void_p), ("coefficients", C.c_void_p),
                ("flag0", C.c_void_p), ("flag1", C.c_void_p), ("flag2", C.c_void_p),
                ("permuted_indices", C.c_void_p),
                ("n_public", C.c_size_t), ("public_wires", C.c_void_p),
                ("n_pfi", C.c_size_t), ("pfi_k", C.c_void_p), ("pfi_w", C.c_void_p),
                ("n_constraints", C.c_size_t), ("n_wires", C.c_size_t)]


_lib = None


def build():
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    if not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, sz, u32 = C.c_void_p, C.c_size_t, C.c_uint32
        L.orc_best_fft.argtypes = [vp, sz, vp, u32, C.c_uint]
        L.orc_inv_best_fft.argtypes = [vp, sz, vp, u32, C.c_uint]
        L.orc_serial_fft.argtypes = [vp, vp, u32]
        L.orc_expand_root_of_unity.argtypes = [vp, sz, vp]
        L.orc_expand_root_of_unity.restype = sz
        L.orc_multi_inv.argtypes = [vp, vp, sz]
        L.orc_blake2s.argtypes = [vp, vp, sz]
        L.orc_get_pseudorandom_indices.argtypes = [vp, vp, sz, u32, sz, u32]
        L.orc_merkle_gen_proofs.argtypes = [vp, sz, sz, vp, sz, vp, vp]
        L.orc_merkle_validate.argtypes = [vp, sz, vp, sz, vp, sz]
        L.orc_merkle_validate.restype = C.c_int
        L.orc_poseidon_hash.argtypes = [vp, vp, sz]
        L.orc_poseidon_hash.restype = C.c_int
        L.orc_poseidon_merkle_gen_proofs.argtypes = [vp, sz, sz, vp, sz, vp, vp]
        L.orc_poseidon_merkle_gen_proofs.restype = C.c_int
        L.orc_poseidon_merkle_validate.argtypes = [vp, sz, vp, sz, vp, sz]
        L.orc_poseidon_merkle_validate.restype = C.c_int
        L.orc_prove_low_degree.argtypes = [C.POINTER(FriProof), vp, sz, vp, sz, u32]
        L.orc_verify_low_degree_proof.argtypes = [vp, vp, C.POINTER(FriProof), sz, u32]
        L.orc_verify_low_degree_proof.restype = C.c_int
        L.orc_fri_proof_free.argtypes = [C.POINTER(FriProof)]
        L.orc_fri_proof_json.argtypes = [C.POINTER(Buf), C.POINTER(FriProof)]
        L.orc_buf_free.argtypes = [C.POINTER(Buf)]
        L.orc_prove_files.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint, C.c_int, C.POINTER(C.c_double)]
        L.orc_prove_files.restype = C.c_int
        L.orc_trace_from_files.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Trace)]
        L.orc_trace_from_files.restype = C.c_int
        L.orc_trace_free.argtypes = [C.POINTER(Trace)]
        L.fp_root_of_unity.argtypes = [vp, u32]
        L.fp_mul.argtypes = [vp, vp, vp]
        L.fp_to_bytes_le.argtypes = [vp, vp]
        L.fp_from_bytes_le.argtypes = [vp, vp, sz]
        L.orc_fp_to_bytes_le_vec.argtypes = [vp, vp, sz]
        _lib = L
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def cpus():
    return os.cpu_count() or 1


def root_of_unity(log_n):
    r = np.zeros(4, dtype=np.uint64)
    lib().fp_root_of_unity(_p(r), log_n)
    return r


def best_fft(vals, root, log_n, inverse=False, n_cpus=None):
    """fft.rs:327-379 on (len, 4) uint64 Montgomery limbs"""
    n = 1 << log_n
    v = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
    buf = np.zeros((n, 4), dtype=np.uint64)
    buf[: v.shape[0]] = v
    root = np.ascontiguousarray(root, dtype=np.uint64)
    fn = lib().orc_inv_best_fft if inverse else lib().orc_best_fft
    fn(_p(buf), v.shape[0], _p(root), log_n, n_cpus or cpus())
    return buf


def expand_root_of_unity(root, cap):
    out = np.zeros((cap, 4), dtype=np.uint64)
    root = np.ascontiguousarray(root, dtype=np.uint64)
    order = lib().orc_expand_root_of_unity(_p(out), cap, _p(root))
    return out[: min(order, cap)], order


def multi_inv(vals):
    v = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(v)
    lib().orc_multi_inv(_p(out), _p(v), v.shape[0])
    return out


def blake(msg):
    m = np.frombuffer(bytes(msg), dtype=np.uint8).copy()
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_blake2s(_p(out), _p(m) if m.size else None, m.size)
    return out.tobytes()


def get_pseudorandom_indices(seed, modulus, count, excl=0):
    s = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
    out = np.zeros(count, dtype=np.uint32)
    lib().orc_get_pseudorandom_indices(_p(out), _p(s), s.size, modulus, count, excl)
    return [int(x) for x in out]


def merkle_gen_proofs(leaves_flat, leaf_bytes, n, indices):
    """returns (root bytes, nodes array (n_idx, depth, 32))"""
    lv = np.frombuffer(bytes(leaves_flat), dtype=np.uint8).copy()
    idx = np.asarray(list(indices), dtype=np.uint64)
    depth = (n - 1).bit_length()
    root = np.zeros(32, dtype=np.uint8)
    nodes = np.zeros((max(idx.size, 1), max(depth, 1), 32), dtype=np.uint8)
    lib().orc_merkle_gen_proofs(_p(lv), leaf_bytes, n, _p(idx) if idx.size else None, idx.size, _p(root), _p(nodes) if idx.size else None)
    return root.tobytes(), nodes[: idx.size, :depth].copy()


def poseidon(msg):
    """poseidon.rs:30-63; raises ValueError where the reference panics"""
    m = np.frombuffer(bytes(msg), dtype=np.uint8).copy()
    out = np.zeros(32, dtype=np.uint8)
    if lib().orc_poseidon_hash(_p(out), _p(m) if m.size else None, m.size):
        raise ValueError("message cannot be hashed (length 0 or > 64, or a chunk is not a canonical scalar)")
    return out.tobytes()


def poseidon_merkle_gen_proofs(leaves_flat, leaf_bytes, n, indices):
    """ParallelMerkleTree<_, PoseidonDigest>: (root bytes, nodes array (n_idx, depth, 32))"""
    lv = np.frombuffer(bytes(leaves_flat), dtype=np.uint8).copy()
    idx = np.asarray(list(indices), dtype=np.uint64)
    depth = (n - 1).bit_length()
    root = np.zeros(32, dtype=np.uint8)
    nodes = np.zeros((max(idx.size, 1), max(depth, 1), 32), dtype=np.uint8)
    if lib().orc_poseidon_merkle_gen_proofs(_p(lv), leaf_bytes, n, _p(idx), idx.size, _p(root), _p(nodes)):
        raise ValueError("a leaf cannot be hashed")
    return root.tobytes(), nodes[: idx.size, :depth].copy()


def poseidon_merkle_validate(root, index, leaf, nodes):
    lf = np.frombuffer(bytes(leaf), dtype=np.uint8).copy()
    nd = np.ascontiguousarray(nodes, dtype=np.uint8).reshape(-1, 32)
    r = np.frombuffer(bytes(root), dtype=np.uint8).copy()
    return bool(lib().orc_poseidon_merkle_validate(_p(r), index, _p(lf), lf.size, _p(nd) if nd.size else None, nd.shape[0]))


def fp_to_bytes_le(vals):
    """fp.rs:39-43 for every element: (n, 32) uint8"""
    v = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((v.shape[0], 32), dtype=np.uint8)
    lib().orc_fp_to_bytes_le_vec(_p(out), _p(v), v.shape[0])
    return out


fp_to_bytes_le_fast = fp_to_bytes_le


def prove_low_degree_json(vals, root, max_deg_plus_1, excl, verify=True):
    """fri.rs:46-224 -> serde_json text of Vec<FriProof>; also returns whether the restated verifier accepts"""
    v = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
    root = np.ascontiguousarray(root, dtype=np.uint64)
    pr = FriProof()
    L = lib()
    L.orc_prove_low_degree(C.byref(pr), _p(v), v.shape[0], _p(root), max_deg_plus_1, excl)
    b = Buf()
    L.orc_fri_proof_json(C.byref(b), C.byref(pr))
    text = C.string_at(b.p, b.len).decode()
    L.orc_buf_free(C.byref(b))
    if not verify:
        L.orc_fri_proof_free(C.byref(pr))
        return text, None
    # verifier needs the root of the values tree
    leaves = fp_to_bytes_le(v).tobytes()
    mroot, _ = merkle_gen_proofs(leaves, 32, v.shape[0], [])
    mr = np.frombuffer(mroot, dtype=np.uint8).copy()
    ok = L.orc_verify_low_degree_proof(_p(mr), _p(root), C.byref(pr), max_deg_plus_1, excl)
    L.orc_fri_proof_free(C.byref(pr))
    return text, bool(ok)


def prove_files(r1cs, wtns, out_path, n_cpus=None, verify=True):
    t = C.c_double()
    rc = lib().orc_prove_files(r1cs.encode(), wtns.encode(), out_path.encode(), n_cpus or cpus(), 1 if verify else 0, C.byref(t))
    return rc, t.value


def sha256_file(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def trace_from_files(r1cs, wtns):
    """the arguments of mk_r1cs_proof for a circuit, as numpy copies (dict)"""
    t = Trace()
    rc = lib().orc_trace_from_files(r1cs.encode(), wtns.encode(), C.byref(t))
    assert rc == 0, rc
    os_ = t.original_steps

    def fp_arr(p, n):
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n, 4)).copy() if n else np.zeros((0, 4), dtype=np.uint64)

    def sz_arr(p, n):
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_size_t)), shape=(n,)).copy() if n else np.zeros(0, dtype=np.uint64)

    out = {"original_steps": os_,
           "witness_trace": fp_arr(t.witness_trace, os_), "computational_trace": fp_arr(t.computational_trace, os_),
           "coefficients": fp_arr(t.coefficients, os_), "flag0": fp_arr(t.flag0, os_), "flag1": fp_arr(t.flag1, os_),
           "flag2": fp_arr(t.flag2, os_), "permuted_indices": sz_arr(t.permuted_indices, os_),
           "public_wires": fp_arr(t.public_wires, t.n_public),
           "pfi_k": sz_arr(t.pfi_k, t.n_pfi), "pfi_w": sz_arr(t.pfi_w, t.n_pfi)}
    lib().orc_trace_free(C.byref(t))
    return out
