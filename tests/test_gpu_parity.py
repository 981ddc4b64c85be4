"""GPU parity tests: the CUDA path (through the C ABI, via the Python mirror of the reference API)
against the CPU oracle on the same seeded inputs.  Bit-exact everywhere (integer / byte work)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from conftest import P, ROOT, random_elems

pytestmark = pytest.mark.gpu


def edge_vectors(n):
    import stark_pure_rust_b200 as sb
    zero = np.zeros((n, 4), dtype=np.uint64)
    d0 = zero.copy(); d0[0] = sb.field.mont_scalar(1)
    dl = zero.copy(); dl[n - 1] = sb.field.mont_scalar(1)
    pm1 = np.tile(sb.field.mont_scalar(P - 1), (n, 1))
    return {"zero": zero, "delta0": d0, "delta_last": dl, "all_p_minus_1": pm1}


def test_library_is_the_cuda_build(ctx):
    import stark_pure_rust_b200 as sb
    assert sb.library_path().endswith("libstark_b200.so")
    assert ctx.launch_count() >= 0


@pytest.mark.parametrize("log_n", [0, 1, 5, 10, 13])
def test_expand_root_of_unity(ctx, oracle, log_n):
    import stark_pure_rust_b200 as sb
    w = oracle.root_of_unity(log_n)
    got = sb.fft.expand_root_of_unity(w, order=1 << log_n, ctx=ctx)
    want, order = oracle.expand_root_of_unity(w, 1 << log_n)
    assert order == 1 << log_n
    assert np.array_equal(got, want)


@pytest.mark.parametrize("log_n", list(range(0, 15)) + [16, 17, 18, 20])
def test_best_fft_matches_oracle(ctx, oracle, log_n):
    import stark_pure_rust_b200 as sb
    n = 1 << log_n
    w = oracle.root_of_unity(log_n)
    v = random_elems(n, 0xB200 + log_n)
    got = sb.fft.best_fft(v, w, log_n, ctx=ctx)
    want = oracle.best_fft(v, w, log_n)
    assert np.array_equal(got, want), "forward NTT differs at %s" % np.argwhere((got != want).any(axis=1))[:4].ravel()
    got_i = sb.fft.inv_best_fft(v, w, log_n, ctx=ctx)
    want_i = oracle.best_fft(v, w, log_n, inverse=True)
    assert np.array_equal(got_i, want_i)
    # round trip
    assert np.array_equal(sb.fft.inv_best_fft(got, w, log_n, ctx=ctx), v)


@pytest.mark.parametrize("log_n,len_in", [(3, 0), (3, 5), (9, 1), (9, 300), (12, 2049), (16, 8192), (16, 6684)])
def test_best_fft_zero_padding(ctx, oracle, log_n, len_in):
    import stark_pure_rust_b200 as sb
    w = oracle.root_of_unity(log_n)
    v = random_elems(max(len_in, 1), 77 + len_in)[:len_in]
    assert np.array_equal(sb.fft.best_fft(v, w, log_n, ctx=ctx), oracle.best_fft(v, w, log_n))
    assert np.array_equal(sb.fft.inv_best_fft(v, w, log_n, ctx=ctx), oracle.best_fft(v, w, log_n, inverse=True))


@pytest.mark.parametrize("log_n", [4, 9, 12])
def test_best_fft_edge_vectors(ctx, oracle, log_n):
    import stark_pure_rust_b200 as sb
    w = oracle.root_of_unity(log_n)
    for name, v in edge_vectors(1 << log_n).items():
        assert np.array_equal(sb.fft.best_fft(v, w, log_n, ctx=ctx), oracle.best_fft(v, w, log_n)), name
        assert np.array_equal(sb.fft.inv_best_fft(v, w, log_n, ctx=ctx), oracle.best_fft(v, w, log_n, inverse=True)), name


def test_best_fft_other_roots(ctx, oracle):
    """any primitive 2^k-th root works (the reference takes the root as an argument): w^3, w^-1"""
    import stark_pure_rust_b200 as sb
    log_n = 10
    w = sb.field.from_mont(oracle.root_of_unity(log_n).reshape(1, 4))[0]
    v = random_elems(1 << log_n, 5)
    for root in (pow(w, 3, P), pow(w, -1, P), pow(w, 2**log_n - 5, P)):
        rl = sb.field.mont_scalar(root)
        assert np.array_equal(sb.fft.best_fft(v, rl, log_n, ctx=ctx), oracle.best_fft(v, rl, log_n))


def test_best_fft_errors(ctx, oracle):
    import stark_pure_rust_b200 as sb
    w = oracle.root_of_unity(4)
    with pytest.raises(sb.StarkB200Error) as e:        # fft.rs:162 assert
        sb.fft.best_fft(random_elems(17, 1), w, 4, ctx=ctx)
    assert e.value.code == -3
    with pytest.raises(sb.StarkB200Error) as e:        # not a primitive 2^5-th root
        sb.fft.best_fft(random_elems(8, 1), w, 5, ctx=ctx)
    assert e.value.code == -4


@pytest.mark.parametrize("log_s,n_cols,col_len", [(4, 1, 16), (4, 3, 15), (10, 10, 1024), (13, 6, 6684), (15, 2, 1 << 15), (15, 5, 30000)])
def test_lde_batch(ctx, oracle, log_s, n_cols, col_len):
    """prove.rs:100-124: best_fft(inv_best_fft(col, g1, log_s), g2, log_s + 3); out[8j] == col[j]"""
    import stark_pure_rust_b200 as sb
    g2 = oracle.root_of_unity(log_s + 3)
    g1 = oracle.root_of_unity(log_s)
    cols = random_elems(n_cols * col_len, 900 + log_s).reshape(n_cols, col_len, 4)
    got = sb.fft.lde_batch(cols, g2, log_s, 3, ctx=ctx)
    for c in range(n_cols):
        coef = oracle.best_fft(cols[c], g1, log_s, inverse=True)
        want = oracle.best_fft(coef, g2, log_s + 3)
        assert np.array_equal(got[c], want), c
        assert np.array_equal(got[c][::8][:col_len], cols[c])


def test_multi_inv(ctx, oracle):
    import stark_pure_rust_b200 as sb
    for n in (1, 7, 1000, 50000):
        v = random_elems(n, 31 + n)
        v[::5] = 0                       # zero passthrough (poly_utils.rs:38-70)
        assert np.array_equal(sb.poly_utils.multi_inv(v, ctx=ctx), oracle.multi_inv(v)), n


def test_multi_inv_two_level(ctx, oracle):
    """n above 148*16*128*8 takes the two-level path (slice products inverted by a nested batch inverse);
    a run of zeros covers a whole slice whose product stays 1"""
    import stark_pure_rust_b200 as sb
    n = 148 * 16 * 128 * 8 + 12345
    v = random_elems(n, 977)
    v[::7] = 0
    v[5::148 * 16 * 128] = 0             # every element of slice 5
    got = sb.poly_utils.multi_inv(v, ctx=ctx)
    assert np.array_equal(got, oracle.multi_inv(v))


# ---- Merkle ---------------------------------------------------------------------------------
KAT16 = ["7fffffff", "80000000", "00000003", "00000000", "7ffffffe", "80000001", "00000004", "00000001",
         "7ffffffd", "80000002", "00000005", "00000002", "7ffffffc", "80000003", "00000006", "00000003"]


def test_merkle_reference_kats(ctx):
    """commitment/src/pallarel_merkle_tree.rs:133-216"""
    import stark_pure_rust_b200 as sb
    t = sb.merkle.MerkleProofInPlace(ctx)
    t.update([bytes.fromhex(x) for x in KAT16])
    pr = t.gen_proofs([2])[0]
    assert t.get_root().hex() == "9f04496db6a8c505e88a7db289161a540a0cb953ef81c9b86103f0d6d12e8e15"
    assert pr.leaf == bytes.fromhex("00000003")
    assert [x.hex() for x in pr.nodes] == [
        "4cd90cc0d54239ee5b3fd9989b4ef4cbebbbdd08410758cbd2d291fa364c82d5",
        "2e3d3579213e0a992d60b503f1d8fe331b8bd548e227e8dbd741ca1752077b84",
        "9a8c87bb98f1b2e0f7036a27a343dc8fd649bedc737093c2080a34c6b9f6f375",
        "ef459d75e20ce2f3fc4378ff20fe2d594fbcf16cccd986c2e0d3df41bd3bbe44"]
    assert pr.validate(t.get_root(), 2)
    t2 = sb.merkle.MerkleProofInPlace(ctx)
    t2.update([bytes.fromhex("7fffffff")] * 4096)
    prs = t2.gen_proofs([2, 7, 13])
    assert t2.get_root().hex() == "a0d91c3115f9e4d9f142e7cb2f413c10f0f2f9f65d9f918b80f852f9ebc06ebc"
    assert prs[0].nodes[0].hex() == "b72b5371ceffa4e01aa1849cdb8705406e14791db359f826bc01a392ed26b6b9"
    assert sb.merkle.verify_multi_branch(t2.get_root(), [2, 7, 13], prs)


@pytest.mark.parametrize("n,leaf_bytes", [(1, 32), (2, 32), (4, 40), (8, 1), (16, 4), (32, 33), (64, 64), (128, 65),
                                          (256, 256), (1024, 40), (4096, 32), (1 << 15, 32), (1 << 13, 256), (512, 0)])
def test_merkle_bytes_matches_oracle(ctx, oracle, n, leaf_bytes):
    """merkle_proof_in_place.rs:106-206 incl. caller-order / duplicate indices (:199-205, test :228)"""
    import stark_pure_rust_b200 as sb
    rng = np.random.default_rng(n * 1000 + leaf_bytes)
    flat = rng.integers(0, 256, size=n * leaf_bytes, dtype=np.uint8).tobytes()
    idx = [int(x) for x in rng.integers(0, n, size=11)] + [0, n - 1, n // 2, n // 2]
    t = sb.merkle.MerkleProofInPlace(ctx)
    t.update([flat[i * leaf_bytes:(i + 1) * leaf_bytes] for i in range(n)])
    prs = t.gen_proofs(idx)
    root, nodes = oracle.merkle_gen_proofs(flat, leaf_bytes, n, idx)
    assert t.get_root() == root
    assert t.width() == n
    for q, i in enumerate(idx):
        assert prs[q].leaf == flat[i * leaf_bytes:(i + 1) * leaf_bytes]
        assert b"".join(prs[q].nodes) == nodes[q].tobytes()
        assert prs[q].validate(root, i)


def test_merkle_in_place_reference_test(ctx, oracle):
    """merkle_proof_in_place.rs:209-261: 16 leaves, indices [10,4,6,3,6,8]"""
    import stark_pure_rust_b200 as sb
    leaves = [bytes.fromhex("%08x" % i) for i in range(16)]
    idx = [10, 4, 6, 3, 6, 8]
    t = sb.merkle.MerkleProofInPlace(ctx)
    t.update(leaves)
    prs = t.gen_proofs(idx)
    root, nodes = oracle.merkle_gen_proofs(b"".join(leaves), 4, 16, idx)
    assert t.get_root() == root
    assert [b"".join(p.nodes) for p in prs] == [nodes[q].tobytes() for q in range(len(idx))]


# ---- the alternative digest (commitment/src/poseidon.rs) ------------------------------------------------
def test_poseidon_reference_kats(ctx):
    """poseidon.rs:66-106 and pallarel_merkle_tree.rs:219-253 on the device"""
    import json
    import stark_pure_rust_b200 as sb
    k = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))["reference_kats"]["poseidon"]
    msgs = [bytes(range(int(n))) + bytes(64 - int(n)) for n in k["digest"]]
    assert [d.hex() for d in sb.merkle.poseidon_hash_many(msgs, ctx)] == list(k["digest"].values())
    t = sb.merkle.ParallelMerkleTree(ctx, digest="poseidon")
    t.update([bytes.fromhex("7fffffff")] * 4096)
    prs = t.gen_proofs([2, 7, 13])
    assert t.get_root().hex() == k["merkle4096"]["root"]
    assert prs[0].leaf == bytes.fromhex("7fffffff")
    assert prs[0].nodes[0].hex() == k["merkle4096"]["first_node_of_2"]
    assert sb.merkle.verify_multi_branch(t.get_root(), [2, 7, 13], prs, "poseidon")
    assert not sb.merkle.verify_multi_branch(t.get_root(), [2, 7, 13], prs)             # not a Blake tree


def _canonical_messages(rng, n, ln):
    a = rng.integers(0, 256, size=(n, ln), dtype=np.uint8)
    for top in (31, 63):
        if top < ln:
            a[:, top] &= 0x3f                     # every 32-byte chunk below the BLS12-381 scalar modulus
    return a


@pytest.mark.parametrize("ln", [1, 3, 4, 31, 32, 33, 36, 40, 63, 64])
def test_poseidon_hash_matches_oracle(ctx, oracle, ln):
    import stark_pure_rust_b200 as sb
    rng = np.random.default_rng(700 + ln)
    a = _canonical_messages(rng, 300, ln)
    r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    if ln >= 32:                                  # the largest scalar, zero, and limbs of all ones below the top
        a[0, :32] = np.frombuffer((r - 1).to_bytes(32, "little"), dtype=np.uint8)
        a[1, :32] = 0
        a[2, :31] = 0xff
    if ln == 64:
        a[3, 32:] = np.frombuffer((r - 1).to_bytes(32, "little"), dtype=np.uint8)
        a[4, 32:63] = 0xff
    got = sb.merkle.poseidon_hash_many([row.tobytes() for row in a], ctx)
    assert got == [oracle.poseidon(row.tobytes()) for row in a]
    assert got[:8] == [sb.utils.poseidon(row.tobytes()) for row in a[:8]]


def test_poseidon_rejects_what_the_reference_panics_on(ctx):
    import stark_pure_rust_b200 as sb
    r_le = (0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001).to_bytes(32, "little")
    for bad in ([bytes(65)], [b""], [bytes(32), r_le], [bytes(32) + r_le], [b"\xff" * 64], [bytes(31) + b"\x80"]):
        with pytest.raises(sb.StarkB200Error):    # poseidon.rs:33 assert, :48 unwrap
            sb.merkle.poseidon_hash_many(bad, ctx)
    t = sb.merkle.MerkleProofInPlace(ctx, digest="poseidon")
    t.update([bytes(32), r_le])
    with pytest.raises(sb.StarkB200Error):
        t.gen_proofs([0])
    t.update([bytes(66)] * 4)
    with pytest.raises(sb.StarkB200Error):
        t.gen_proofs([0])
    t.update([b"abcd"] * 12)
    with pytest.raises(sb.StarkB200Error):
        t.gen_proofs([0])
    assert sb.merkle.poseidon_hash_many([], ctx) == []


@pytest.mark.parametrize("n,leaf_bytes", [(1, 32), (2, 4), (8, 64), (16, 40), (256, 33), (1024, 32), (1 << 13, 64), (1 << 15, 32)])
def test_poseidon_merkle_matches_oracle(ctx, oracle, n, leaf_bytes):
    """ParallelMerkleTree<Vec<u8>, PoseidonDigest>: root, caller-order / duplicate openings, branch checks"""
    import stark_pure_rust_b200 as sb
    rng = np.random.default_rng(n * 77 + leaf_bytes)
    flat = _canonical_messages(rng, n, leaf_bytes).tobytes()
    idx = [int(x) for x in rng.integers(0, n, size=9)] + [0, n - 1, n // 2, n // 2]
    t = sb.merkle.MerkleProofInPlace(ctx, digest="poseidon")
    t.update([flat[i * leaf_bytes:(i + 1) * leaf_bytes] for i in range(n)])
    prs = t.gen_proofs(idx)
    root, nodes = oracle.poseidon_merkle_gen_proofs(flat, leaf_bytes, n, idx)
    assert t.get_root() == root and t.width() == n
    for q, i in enumerate(idx):
        assert prs[q].leaf == flat[i * leaf_bytes:(i + 1) * leaf_bytes]
        assert b"".join(prs[q].nodes) == nodes[q].tobytes()
        assert oracle.poseidon_merkle_validate(root, i, prs[q].leaf, nodes[q])
    assert sb.merkle.verify_multi_branch(root, idx[:3], prs[:3], "poseidon")


def test_merkle_not_power_of_two(ctx):
    import stark_pure_rust_b200 as sb
    t = sb.merkle.MerkleProofInPlace(ctx)
    t.update([b"abcd"] * 12)
    with pytest.raises(sb.StarkB200Error):            # merkle_proof_in_place.rs:113 assert
        t.gen_proofs([0])


@pytest.mark.parametrize("n,nc", [(8, 1), (1 << 12, 1), (1 << 10, 8), (1 << 14, 8), (64, 3), (1 << 16, 1), (256, 5), (2, 2)])
def test_merkle_cols_dev(ctx, oracle, n, nc):
    """column-backed leaves: to_bytes_le(col_0[i]) || ... (prove.rs:235-258, :324-327)"""
    import stark_pure_rust_b200 as sb
    cols = random_elems(n * nc, 4242 + n + nc).reshape(nc, n, 4)
    dptrs = [ctx.to_device(cols[k]) for k in range(nc)]
    arr = (C.c_void_p * nc)(*dptrs)
    root = np.empty(32, dtype=np.uint8)
    tree = C.c_void_p()
    ctx.check(ctx.lib.sb_merkle_commit_cols_dev(ctx.h, arr, nc, n, root.ctypes.data_as(C.c_void_p), C.byref(tree)))
    by = np.concatenate([oracle.fp_to_bytes_le(cols[k]).reshape(n, 1, 32) for k in range(nc)], axis=1)   # (n, nc, 32)
    flat = by.tobytes()
    idx = [0, n - 1, n // 3, n // 3]
    want_root, want_nodes = oracle.merkle_gen_proofs(flat, 32 * nc, n, idx)
    assert root.tobytes() == want_root
    depth = (n - 1).bit_length()
    leaves = np.empty(len(idx) * 32 * nc, dtype=np.uint8)
    nodes = np.empty(len(idx) * depth * 32, dtype=np.uint8)
    ia = np.asarray(idx, dtype=np.uint64)
    ctx.check(ctx.lib.sb_merkle_open(ctx.h, tree, ia.ctypes.data_as(C.POINTER(C.c_size_t)), len(idx),
                                     leaves.ctypes.data_as(C.c_void_p), nodes.ctypes.data_as(C.c_void_p) if depth else None))
    assert nodes.tobytes() == want_nodes.tobytes()
    assert leaves.tobytes() == b"".join(flat[i * 32 * nc:(i + 1) * 32 * nc] for i in idx)
    ctx.lib.sb_tree_free(ctx.h, tree)
    for d in dptrs:
        ctx.free(d)


# ---- FRI --------------------------------------------------------------------------------------
@pytest.mark.parametrize("log_n,deg_log", [(7, 4), (9, 6), (10, 7), (12, 9), (14, 11), (16, 13)])
def test_prove_low_degree_matches_oracle(ctx, oracle, log_n, deg_log):
    """fri.rs:46-224 called the way prove.rs:367 does: values = evaluations of a polynomial of degree
    < 2^deg_log on the 2^log_n domain, bound 2^(log_n-2), exclude multiples of 8.  Compared as serde
    JSON text (roots, 40 + 160 openings per layer, last layer) byte for byte."""
    import stark_pure_rust_b200 as sb
    w = oracle.root_of_unity(log_n)
    coeffs = random_elems(1 << deg_log, 600 + log_n)
    values = oracle.best_fft(coeffs, w, log_n)
    want, ok = oracle.prove_low_degree_json(values, w, (1 << log_n) // 4, 8)
    assert ok, "oracle verifier rejects its own proof"
    got = sb.fri.prove_low_degree(values, w, (1 << log_n) // 4, 8, ctx=ctx, as_json=True)
    assert hashlib.sha256(got.encode()).hexdigest() == hashlib.sha256(want.encode()).hexdigest()
    assert got == want


def test_prove_low_degree_structure(ctx, oracle):
    import stark_pure_rust_b200 as sb
    log_n = 10
    w = oracle.root_of_unity(log_n)
    values = oracle.best_fft(random_elems(64, 3), w, log_n)
    proof = sb.fri.prove_low_degree(values, w, 256, 8, ctx=ctx)
    # 1024 -> 256 -> 64 -> 16 (bound 256 -> 64 -> 16): two Middle layers then Last of 64 values
    assert [list(l)[0] for l in proof] == ["Middle", "Middle", "Last"]
    assert len(proof[0]["Middle"]["column_branches"]) == 40 and len(proof[0]["Middle"]["poly_branches"]) == 160
    assert len(proof[-1]["Last"]["last"]) == 64
    # column openings verify against root2, and the first layer's root2 is the next layer's values root
    ys = sb.utils.get_pseudorandom_indices(proof[0]["Middle"]["root2"], 256, 40, 8)
    assert all(y % 8 != 0 for y in ys)
    assert sb.merkle.verify_multi_branch(proof[0]["Middle"]["root2"], ys, proof[0]["Middle"]["column_branches"])


# ---- whole prover ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["compute", "poseidon3_test"])
def test_mk_r1cs_proof_bit_identical(ctx, oracle, name, tmp_path):
    """prove.rs:14-378 device resident: proof.json byte-identical to the oracle's (and to the committed golden
    hash, which the survey's independent model also reproduced)"""
    import json
    import os
    import stark_pure_rust_b200 as sb
    from conftest import ROOT
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    tr = oracle.trace_from_files(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"))
    text = sb.prove.mk_r1cs_proof(tr["witness_trace"], tr["computational_trace"], tr["public_wires"],
                                  list(zip(tr["pfi_k"].tolist(), tr["pfi_w"].tolist())), tr["permuted_indices"], tr["coefficients"],
                                  tr["flag0"], tr["flag1"], tr["flag2"], ctx=ctx)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))["proofs"][name]
    out = str(tmp_path / "p.json")
    rc, _ = oracle.prove_files(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), out)
    assert rc == 0
    want = open(out).read()
    if text != want:
        a, b = json.loads(text), json.loads(want)
        for key in ("a_root", "m_root", "l_root"):
            assert a[key] == b[key], key
    assert text == want
    assert len(text) == gold["proof_json_bytes"] and hashlib.sha256(text.encode()).hexdigest() == gold["proof_json_sha256"]


def test_mk_r1cs_proof_rejects_bad_witness(ctx, oracle):
    """the reference panics in calc_d1 when the trace does not satisfy the constraints (utils.rs:379-384)"""
    import os
    import stark_pure_rust_b200 as sb
    from conftest import ROOT
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    tr = oracle.trace_from_files(os.path.join(d, "compute.r1cs"), os.path.join(d, "compute.wtns"))
    bad = tr["computational_trace"].copy()
    bad[3] = sb.field.mont_scalar(12345)
    with pytest.raises(sb.StarkB200Error) as e:
        sb.prove.mk_r1cs_proof(tr["witness_trace"], bad, tr["public_wires"], list(zip(tr["pfi_k"].tolist(), tr["pfi_w"].tolist())),
                               tr["permuted_indices"], tr["coefficients"], tr["flag0"], tr["flag1"], tr["flag2"], ctx=ctx)
    assert e.value.code == -3


@pytest.mark.parametrize("name", ["compute", "poseidon3_test"])
def test_prove_with_file_path_and_cli(ctx, oracle, name, tmp_path):
    """run.rs:528-554 + main.rs:4-11: product-side parsers / trace arrangement / JSON writer, no oracle on the path"""
    import json
    import os
    import subprocess
    import stark_pure_rust_b200 as sb
    from conftest import ROOT
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))["proofs"][name]
    out = str(tmp_path / "proof.json")
    ms = sb.prove.prove_with_file_path(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), out, ctx=ctx)
    assert oracle.sha256_file(out) == gold["proof_json_sha256"] and os.path.getsize(out) == gold["proof_json_bytes"]
    assert ms[4] > 0
    out2 = str(tmp_path / "proof_cli.json")
    r = subprocess.run([os.path.join(ROOT, "stark_pure_rust_b200", "r1cs-stark"), os.path.join(d, name + ".r1cs"),
                        os.path.join(d, name + ".wtns"), out2], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert oracle.sha256_file(out2) == gold["proof_json_sha256"]


def test_prove_with_file_path_errors(ctx, tmp_path):
    import os
    import stark_pure_rust_b200 as sb
    from conftest import ROOT
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    with pytest.raises(sb.StarkB200Error):
        sb.prove.prove_with_file_path(os.path.join(d, "missing.r1cs"), os.path.join(d, "compute.wtns"), None, ctx=ctx)
    bad = tmp_path / "bad.r1cs"
    bad.write_bytes(b"nope" + bytes(100))
    with pytest.raises(sb.StarkB200Error):
        sb.prove.prove_with_file_path(bad, os.path.join(d, "compute.wtns"), None, ctx=ctx)
    # witness of another circuit: the constraint check of the reference fires (utils.rs:379-418)
    with pytest.raises(sb.StarkB200Error):
        sb.prove.prove_with_file_path(os.path.join(d, "poseidon3_test.r1cs"), os.path.join(d, "compute.wtns"), None, ctx=ctx)


@pytest.mark.parametrize("n_constraints,avg_terms,seed,n_pub", [(40, 2.0, 3, 2), (300, 3.0, 1, 2), (1200, 4.0, 7, 2), (400, 3.0, 11, 60)])
def test_prove_synthetic_circuits(ctx, oracle, n_constraints, avg_terms, seed, n_pub, tmp_path):
    """seeded synthetic circuits (tools/gen_r1cs.py: linear and multi-term-C constraints, padding rows, public
    inputs): proof.json of the CUDA pipeline == the oracle's, and the oracle's restated verifier accepts it.
    n_pub = 2 takes the Horner evaluation of the boundary polynomials, n_pub = 60 the NTT of their coefficients."""
    import os
    import sys
    import stark_pure_rust_b200 as sb
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_r1cs
    prefix = str(tmp_path / "syn")
    wit, cons = gen_r1cs.generate(n_constraints, avg_terms, n_pub, seed)
    gen_r1cs.write_files(prefix, wit, cons, n_pub)
    want = str(tmp_path / "want.json")
    rc, _ = oracle.prove_files(prefix + ".r1cs", prefix + ".wtns", want)
    assert rc == 0                      # 0 = proved and verified by the oracle
    got = str(tmp_path / "got.json")
    sb.prove.prove_with_file_path(prefix + ".r1cs", prefix + ".wtns", got, ctx=ctx)
    assert oracle.sha256_file(got) == oracle.sha256_file(want)


# ---- sharded commitment -------------------------------------------------------------------
def _run_sharded_check(world, log_n, n_cols):
    import os
    import subprocess
    import sys
    from conftest import ROOT
    script = os.path.join(ROOT, "tools", "sharded_check.py")
    if world == 1:
        cmd = [sys.executable, script, str(log_n), str(n_cols)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(29600 + world), script, str(log_n), str(n_cols)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout and "SHARDED_FRI_OK" in r.stdout and "SHARDED_NTT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_sharded_commit_world1():
    """the sharded code path with one rank (no collective): same root and openings as the plain column tree"""
    _run_sharded_check(1, 14, 5)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_commit_multi_gpu(world):
    """column-sharded LDE -> NCCL send/recv to row shards -> subtree commit -> all_gather of roots, against one
    single-GPU tree over all columns.  Needs `world` GPUs on the box (gpurun --gpus N); skipped otherwise."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _run_sharded_check(world, 16, 9)


# ---- BASELINE.json full sizes: size-independent properties (the oracle takes minutes at these sizes) ------------
def _fri_verify_python(proof, values_root, root_int, n, max_deg_plus_1, excl):
    """fri.rs:244-404 restated with Python big ints on the GPU proof: every opening validates against its root, the
    sampled column values equal the degree-<4 interpolant of the four row values at special_x, and the last layer is a
    polynomial of degree < its bound."""
    import stark_pure_rust_b200 as sb
    from stark_pure_rust_b200 import field
    Pm = field.P
    w, bound, root = root_int, max_deg_plus_1, values_root
    for layer in proof[:-1]:
        m = layer["Middle"]
        q = n // 4
        special_x = int.from_bytes(root, "little") % Pm                       # fri.rs:275
        ys = sb.utils.get_pseudorandom_indices(m["root2"], q, 40, excl)
        assert sb.merkle.verify_multi_branch(m["root2"], ys, m["column_branches"])
        pos = [y + q * j for y in ys for j in range(4)]
        assert sb.merkle.verify_multi_branch(root, pos, m["poly_branches"])
        iota = pow(w, q, Pm)
        for k, y in enumerate(ys):
            x = pow(w, y, Pm)
            xs = [x * pow(iota, j, Pm) % Pm for j in range(4)]
            vs = [int.from_bytes(m["poly_branches"][4 * k + j].leaf, "little") for j in range(4)]
            acc = 0                                                          # Lagrange through the four points
            for i in range(4):
                num, den = 1, 1
                for j in range(4):
                    if i != j:
                        num = num * (special_x - xs[j]) % Pm
                        den = den * (xs[i] - xs[j]) % Pm
                acc = (acc + vs[i] * num * pow(den, -1, Pm)) % Pm
            assert acc == int.from_bytes(m["column_branches"][k].leaf, "little")
        root, w, n, bound = m["root2"], pow(w, 4, Pm), q, bound // 4
    last = [int.from_bytes(b, "little") for b in proof[-1]["Last"]["last"]]
    assert len(last) == n
    # degree check of the last layer by evaluating the interpolant of the first `bound` points at the others is O(n^2)
    # with n <= 64: Lagrange at a few extra points
    xs = [pow(w, i, Pm) for i in range(n)]
    for probe in range(bound, min(n, bound + 3)):
        acc = 0
        for i in range(bound):
            num, den = 1, 1
            for j in range(bound):
                if i != j:
                    num = num * (xs[probe] - xs[j]) % Pm
                    den = den * (xs[i] - xs[j]) % Pm
            acc = (acc + last[i] * num * pow(den, -1, Pm)) % Pm
        assert acc == last[probe]
    return True


def test_full_size_2_24_properties(ctx):
    """BASELINE.json configs[1] top of the sweep (one 2^24 column): LDE restricted to the original domain returns the
    input, INTT(NTT) round trip with the zero upper 7/8, Merkle checksum-of-checksums and openings, FRI self-consistency."""
    import stark_pure_rust_b200 as sb
    from stark_pure_rust_b200 import field
    from stark_pure_rust_b200._lib import _ptr
    L, log_s = 24, 21
    N, S = 1 << L, 1 << log_s
    g2i = field.root_of_unity(L)
    col = random_elems(S, 2424)
    ext = sb.fft.lde_batch(col.reshape(1, S, 4), g2i, log_s, 3, ctx=ctx)[0]
    assert np.array_equal(ext[::8], col)                                        # the subgroup LDE keeps the trace values
    coef = sb.fft.inv_best_fft(ext, g2i, L, ctx=ctx)
    assert not coef[S:].any()                                                   # degree < S
    assert np.array_equal(sb.fft.best_fft(coef[:S], pow(g2i, 8, field.P), log_s, ctx=ctx), col)
    # Merkle: the root of the whole tree is H(root(first half) || root(second half)); openings validate on the host
    leaves = np.ascontiguousarray(ext).view(np.uint8).reshape(N, 32)            # any 32-byte leaves will do
    def commit(buf, n):
        root, t = np.empty(32, dtype=np.uint8), C.c_void_p()
        ctx.check(ctx.lib.sb_merkle_commit(ctx.h, _ptr(buf), 32, n, _ptr(root), C.byref(t)))
        return root.tobytes(), t
    root, tree = commit(leaves, N)
    r0, t0 = commit(leaves[: N // 2], N // 2)
    r1, t1 = commit(leaves[N // 2:], N // 2)
    assert sb.utils.blake(r0 + r1) == root
    idx = np.array([0, N - 1, N // 2, 12345678, 12345678], dtype=np.uint64)
    lv, nd = np.empty(idx.size * 32, dtype=np.uint8), np.empty(idx.size * L * 32, dtype=np.uint8)
    ctx.check(ctx.lib.sb_merkle_open(ctx.h, tree, idx.ctypes.data_as(C.POINTER(C.c_size_t)), idx.size, _ptr(lv), _ptr(nd)))
    for q, i in enumerate(idx):
        pr = sb.merkle.Proof(lv[32 * q:32 * q + 32].tobytes(), [nd[(q * L + l) * 32:(q * L + l + 1) * 32].tobytes() for l in range(L)])
        assert pr.leaf == leaves[int(i)].tobytes() and pr.validate(root, int(i))
    for t in (tree, t0, t1):
        ctx.lib.sb_tree_free(ctx.h, t)
    # FRI on the extended column (degree < S <= N/4)
    proof = sb.fri.prove_low_degree(ext, g2i, N // 4, 8, ctx=ctx)
    assert [list(l)[0] for l in proof] == ["Middle"] * 9 + ["Last"]
    # root of the values tree: leaves are to_bytes_le of the values
    vals_tree = sb.merkle.MerkleProofInPlace(ctx)
    h, rootv, tv = C.c_void_p(), np.empty(32, dtype=np.uint8), C.c_void_p()
    d = ctx.to_device(ext)
    ptrs = (C.c_void_p * 1)(d)
    ctx.check(ctx.lib.sb_merkle_commit_cols_dev(ctx.h, ptrs, 1, N, _ptr(rootv), C.byref(tv)))
    ctx.lib.sb_tree_free(ctx.h, tv)
    ctx.free(d)
    assert _fri_verify_python(proof, rootv.tobytes(), g2i, N, N // 4, 8)


def test_full_size_2_24_golden(ctx):
    """Byte-level parity at the headline size (BASELINE.json configs[1], N = 2^24): sha256 of best_fft / inv_best_fft at 2^22
    and 2^24, of every column of the 8-column LDE 2^21 -> 2^24, the 8-column Merkle root + openings, the 1-column root and
    the serde JSON of prove_low_degree, against digests the CPU oracle produced once (tests/golden/make_golden_large.py)."""
    import json
    import os
    import stark_pure_rust_b200 as sb
    from stark_pure_rust_b200 import field
    from stark_pure_rust_b200._lib import _ptr
    from conftest import ROOT
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors_large.json")))
    sha = lambda b: hashlib.sha256(bytes(b)).hexdigest()
    for e in gold["ntt"]:
        k = e["log_n"]
        v = random_elems(1 << k, e["seed"])
        w = field.root_of_unity(k)
        assert sha(sb.fft.best_fft(v, w, k, ctx=ctx).tobytes()) == e["fwd_sha256"], "best_fft 2^%d" % k
        assert sha(sb.fft.inv_best_fft(v, w, k, ctx=ctx).tobytes()) == e["inv_sha256"], "inv_best_fft 2^%d" % k
    log_s, L, nc = gold["log_s"], gold["log_n"], gold["n_cols"]
    S, N = 1 << log_s, 1 << L
    g2 = field.mont_scalar(field.root_of_unity(L))
    cols = random_elems(nc * S, gold["lde"]["seed"]).reshape(nc, S, 4)
    d_cols = ctx.to_device(cols)
    d_ext = ctx.alloc(nc * N * 32)
    ctx.check(ctx.lib.sb_lde_batch_dev(ctx.h, C.c_void_p(d_cols), nc, S, S, _ptr(g2), log_s, L - log_s, C.c_void_p(d_ext)))
    one = np.empty((N, 4), dtype=np.uint64)
    for c in range(nc):
        ctx.d2h(one, d_ext + c * N * 32)
        assert sha(one.tobytes()) == gold["lde"]["col_sha256"][c], "LDE column %d" % c
    ptrs = (C.c_void_p * nc)(*[d_ext + c * N * 32 for c in range(nc)])
    root, tree = np.empty(32, dtype=np.uint8), C.c_void_p()
    ctx.check(ctx.lib.sb_merkle_commit_cols_dev(ctx.h, ptrs, nc, N, _ptr(root), C.byref(tree)))
    assert root.tobytes().hex() == gold["merkle8"]["root"]
    idx = np.array(gold["merkle8"]["open"], dtype=np.uint64)
    lv, nd = np.empty(idx.size * 32 * nc, dtype=np.uint8), np.empty(idx.size * L * 32, dtype=np.uint8)
    ctx.check(ctx.lib.sb_merkle_open(ctx.h, tree, idx.ctypes.data_as(C.POINTER(C.c_size_t)), idx.size, _ptr(lv), _ptr(nd)))
    assert sha(nd.tobytes()) == gold["merkle8"]["nodes_sha256"] and sha(lv.tobytes()) == gold["merkle8"]["leaves_sha256"]
    ctx.lib.sb_tree_free(ctx.h, tree)
    p1 = (C.c_void_p * 1)(d_ext + (nc - 1) * N * 32)
    ctx.check(ctx.lib.sb_merkle_commit_cols_dev(ctx.h, p1, 1, N, _ptr(root), C.byref(tree)))
    assert root.tobytes().hex() == gold["merkle1"]["root"]
    pr = C.c_void_p()
    ctx.check(ctx.lib.sb_fri_prove_dev(ctx.h, C.c_void_p(p1[0]), N, _ptr(g2), N // 4, 8, tree, C.byref(pr)))
    txt = ctx.lib.sb_fri_proof_json(pr)
    text = C.string_at(txt)
    ctx.lib.sb_free_string(txt)
    ctx.lib.sb_fri_proof_free(pr)
    ctx.lib.sb_tree_free(ctx.h, tree)
    ctx.free(d_cols)
    ctx.free(d_ext)
    assert len(text) == gold["fri"]["json_len"] and sha(text) == gold["fri"]["json_sha256"]


@pytest.mark.parametrize("name", ["compute", "poseidon3_test"])
def test_product_verifier_on_oracle_proof_files(ctx, oracle, name, tmp_path):
    """the product verifier (sb_verify_files: verify.rs:13-258) is fed whole proof.json files written by the ORACLE's prover,
    and rejects them after a one-byte change"""
    import os
    import stark_pure_rust_b200 as sb
    from conftest import ROOT
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    r1cs, wtns, out = os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), str(tmp_path / "oracle_proof.json")
    rc, _ = oracle.prove_files(r1cs, wtns, out, verify=False)
    assert rc == 0
    sb.prove.verify_with_file_path(r1cs, wtns, out, ctx=ctx)
    text = open(out).read()
    k = text.index('"l_root":[') + len('"l_root":[')
    digit = text[k]
    bad = text[:k] + ("1" if digit != "1" else "2") + text[k + 1:]
    open(out, "w").write(bad)
    with pytest.raises(sb.StarkB200Error):
        sb.prove.verify_with_file_path(r1cs, wtns, out, ctx=ctx)


def test_synthetic_sha256_scale_prove_matches_golden(ctx, tmp_path):
    """BASELINE.json configs[3] stand-in (tools/gen_r1cs.py 30000 constraints, avg 8 terms, seed 1: 955086 steps, precision
    2^23): proof.json hash equals the CPU oracle's, recorded in tests/golden/vectors.json (the oracle needs ~70 s on 16
    cores / 3 min on 8 for this circuit, so it is not re-run here)."""
    import json
    import os
    import sys
    import stark_pure_rust_b200 as sb
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_r1cs
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))["proofs"]["synthetic_30000_8_2_1"]
    prefix = str(tmp_path / "syn")
    wit, cons = gen_r1cs.generate(30000, 8.0, 2, 1)
    info = gen_r1cs.write_files(prefix, wit, cons, 2)
    assert info["original_steps"] == gold["original_steps"]
    out = str(tmp_path / "proof.json")
    sb.prove.prove_with_file_path(prefix + ".r1cs", prefix + ".wtns", out, ctx=ctx)
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == gold["proof_json_sha256"]
    assert os.path.getsize(out) == gold["proof_json_bytes"]


# ---- verifier (verify.rs:13-258, fri.rs:226-404): SURVEY.md 8f next-3 ------------------------------------------
def _prove_files(ctx, name, tmp_path):
    import os
    import stark_pure_rust_b200 as sb
    from conftest import ROOT
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    out = str(tmp_path / (name + ".proof.json"))
    sb.prove.prove_with_file_path(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), out, ctx=ctx)
    return os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), out


@pytest.mark.parametrize("name", ["compute", "poseidon3_test"])
def test_verifier_accepts_and_rejects(ctx, name, tmp_path):
    """the device-assisted verifier accepts the prover's proof.json and rejects every single-byte tampering tried:
    roots, a main-branch leaf, a Merkle sibling, an FRI column value, an FRI root2, a last-layer value"""
    import json
    import stark_pure_rust_b200 as sb
    r1cs, wtns, out = _prove_files(ctx, name, tmp_path)
    ms = sb.prove.verify_with_file_path(r1cs, wtns, out, ctx=ctx)
    assert ms[1] > 0
    proof = json.load(open(out))

    def tampered(mutate):
        p = json.loads(json.dumps(proof))
        mutate(p)
        path = str(tmp_path / "bad.json")
        json.dump(p, open(path, "w"), separators=(",", ":"))
        with pytest.raises(sb.StarkB200Error) as e:
            sb.prove.verify_with_file_path(r1cs, wtns, path, ctx=ctx)
        assert e.value.code == -6, e.value
        return str(e.value)

    def flip(lst, i=0):
        lst[i] = (lst[i] + 1) % 256

    assert "m_root" in tampered(lambda p: flip(p["m_root"]))
    tampered(lambda p: flip(p["l_root"]))
    assert "Q3" in tampered(lambda p: flip(p["a_root"]))
    assert "main branch" in tampered(lambda p: flip(p["main_branches"][5]["leaf"], 40))
    assert "main branch" in tampered(lambda p: flip(p["main_branches"][0]["nodes"][2], 7))
    assert "linear combination branch" in tampered(lambda p: flip(p["linear_comb_branches"][3]["leaf"], 1))
    assert "FRI" in tampered(lambda p: flip(p["fri_proof"][0]["Middle"]["column_branches"][2]["leaf"], 3))
    assert "FRI" in tampered(lambda p: flip(p["fri_proof"][0]["Middle"]["root2"], 9))
    assert "FRI" in tampered(lambda p: flip(p["fri_proof"][-1]["Last"]["last"][5], 0))
    # a consistent-looking change that only the constraint equations catch: swap d1 and d2 inside a leaf of position 0
    def swap_fields(p):
        leaf = p["main_branches"][0]["leaf"]
        leaf[96:128], leaf[128:160] = leaf[128:160], leaf[96:128]
    assert "main branch" in tampered(swap_fields)


def test_verifier_rejects_other_circuit_and_public_inputs(ctx, oracle, tmp_path):
    """a valid proof of one circuit does not verify against another circuit's r1cs or against changed public wires"""
    import os
    import sys
    import stark_pure_rust_b200 as sb
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_r1cs
    pa, pb = str(tmp_path / "a"), str(tmp_path / "b")
    wit, cons = gen_r1cs.generate(300, 3.0, 2, 1)
    gen_r1cs.write_files(pa, wit, cons, 2)
    out = str(tmp_path / "a.json")
    sb.prove.prove_with_file_path(pa + ".r1cs", pa + ".wtns", out, ctx=ctx)
    sb.prove.verify_with_file_path(pa + ".r1cs", pa + ".wtns", out, ctx=ctx)
    # same structure, different public input value: the boundary check S(x) - I2(x) = Zb2(x) B2(x) fails
    wit2 = list(wit)
    wit2[1] = (wit2[1] + 1) % gen_r1cs.P
    gen_r1cs.write_files(pb, wit2, cons, 2)
    with pytest.raises(sb.StarkB200Error) as e:
        sb.prove.verify_with_file_path(pa + ".r1cs", pb + ".wtns", out, ctx=ctx)
    assert e.value.code == -6 and "I2" in str(e.value)
    # different circuit of the same size class
    wit3, cons3 = gen_r1cs.generate(300, 3.0, 2, 2)
    gen_r1cs.write_files(pb, wit3, cons3, 2)
    with pytest.raises(sb.StarkB200Error) as e:
        sb.prove.verify_with_file_path(pb + ".r1cs", pb + ".wtns", out, ctx=ctx)
    assert e.value.code in (-6, -3)


@pytest.mark.parametrize("name", ["bits", "pedersen_test"])
def test_other_bundled_circuits_match_golden(ctx, name, tmp_path):
    """the reference's two other bundled circuits (r1cs-stark/tests: bits has 1062 public wires -> the NTT path of the boundary
    polynomials and an O(n^2) host interpolation; pedersen_test has precision 2^18): proof.json hash equals the oracle's
    (tests/golden/vectors.json; the same hashes as SURVEY.md Appendix C's independent model), and the verifier accepts it"""
    import json
    import os
    import stark_pure_rust_b200 as sb
    from conftest import ROOT
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))["proofs"][name]
    out = str(tmp_path / "proof.json")
    sb.prove.prove_with_file_path(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), out, ctx=ctx)
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == gold["proof_json_sha256"]
    assert os.path.getsize(out) == gold["proof_json_bytes"]
    sb.prove.verify_with_file_path(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), out, ctx=ctx)
