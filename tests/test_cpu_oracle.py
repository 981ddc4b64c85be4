"""CPU suite (-m "not gpu"): pins the oracle against the reference's own KATs and the committed golden
vectors, checks the host-side pieces of the product (Blake2s, sampler, C ABI surface, generated field
code) and the no-fallback rule.  No CUDA compute happens here."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import P, ROOT, random_elems

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
KAT16 = ["7fffffff", "80000000", "00000003", "00000000", "7ffffffe", "80000001", "00000004", "00000001",
         "7ffffffd", "80000002", "00000005", "00000002", "7ffffffc", "80000003", "00000006", "00000003"]


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


# ---- oracle vs the reference's KATs ----------------------------------------------------------
def test_oracle_blake_kat(oracle):
    k = GOLD["reference_kats"]["blake"]
    d = oracle.blake(b"hello world")
    assert d.hex() == k["hello world"]
    assert oracle.blake(d).hex() == k["blake(blake(hello world))"]
    for m in (b"", b"a" * 63, b"a" * 64, b"a" * 65, b"a" * 128, bytes(range(256))):
        assert oracle.blake(m) == hashlib.blake2s(m).digest()


def test_oracle_sampler_kat(oracle):
    k = GOLD["reference_kats"]["sampler"]
    assert oracle.get_pseudorandom_indices(oracle.blake(b"hello world"), 7, 5, 0) == k["hello world,7,5,0"]
    assert oracle.get_pseudorandom_indices(oracle.blake(b"hello another world"), 7, 20, 0) == k["hello another world,7,20,0"]
    idx = oracle.get_pseudorandom_indices(oracle.blake(b"x"), 1 << 13, 40, 8)
    assert all(i % 8 != 0 and 0 < i < (1 << 13) for i in idx)


def test_oracle_merkle_kats(oracle):
    k = GOLD["reference_kats"]
    root, nodes = oracle.merkle_gen_proofs(b"".join(bytes.fromhex(x) for x in KAT16), 4, 16, [2])
    assert root.hex() == k["merkle16"]["root"]
    assert [nodes[0][l].tobytes().hex() for l in range(4)] == k["merkle16"]["nodes_of_2"]
    root, nodes = oracle.merkle_gen_proofs(bytes.fromhex("7fffffff") * 4096, 4, 4096, [2, 7, 13])
    assert root.hex() == k["merkle4096"]["root"]
    assert nodes[0][0].tobytes().hex() == k["merkle4096"]["first_node_of_2"]


def test_oracle_field_codecs(oracle):
    """ff_utils/src/fp.rs:28-68: 31 -> big/little endian bytes"""
    from stark_pure_rust_b200 import field
    b = oracle.fp_to_bytes_le(field.to_mont([31, 1, P - 1]))
    assert b[0].tobytes() == bytes([31] + [0] * 31)
    assert b[2].tobytes() == (P - 1).to_bytes(32, "little")
    assert field.from_bytes_le(b"\xff" * 32) == (2**256 - 1) % P


# ---- oracle vs committed golden vectors ----------------------------------------------------------
@pytest.mark.parametrize("g", GOLD["ntt"], ids=lambda g: "2^%d/%d" % (g["log_n"], g["len_in"]))
def test_oracle_ntt_golden(oracle, g):
    v = random_elems(max(g["len_in"], 1), g["seed"])[: g["len_in"]]
    w = oracle.root_of_unity(g["log_n"])
    assert sha(oracle.best_fft(v, w, g["log_n"]).tobytes()) == g["fwd_sha256"]
    assert sha(oracle.best_fft(v, w, g["log_n"], inverse=True).tobytes()) == g["inv_sha256"]
    # results do not depend on the thread split (fft.rs:342-347)
    assert sha(oracle.best_fft(v, w, g["log_n"], n_cpus=1).tobytes()) == g["fwd_sha256"]


def test_oracle_ntt_known_answer(oracle):
    from stark_pure_rust_b200 import field
    g = GOLD["ntt_1_to_8"]
    w8 = oracle.root_of_unity(3)
    assert hex(field.from_mont(w8.reshape(1, 4))[0]) == g["root"]
    out = field.from_mont(oracle.best_fft(field.to_mont(range(1, 9)), w8, 3))
    assert [hex(x) for x in out] == g["out"]
    assert out[0] == 36          # sum 1..8
    w = field.from_mont(w8.reshape(1, 4))[0]
    assert out == [sum((j + 1) * pow(w, j * k, P) for j in range(8)) % P for k in range(8)]


@pytest.mark.parametrize("g", GOLD["merkle_cols"], ids=lambda g: "%dx%d" % (g["n"], g["nc"]))
def test_oracle_merkle_golden(oracle, g):
    n, nc = g["n"], g["nc"]
    cols = random_elems(n * nc, g["seed"]).reshape(nc, n, 4)
    by = np.concatenate([oracle.fp_to_bytes_le(cols[k]).reshape(n, 1, 32) for k in range(nc)], axis=1).tobytes()
    root, nodes = oracle.merkle_gen_proofs(by, 32 * nc, n, [0, n - 1, n // 3])
    assert root.hex() == g["root"] and sha(nodes.tobytes()) == g["nodes_sha256"]


@pytest.mark.parametrize("g", GOLD["fri"], ids=lambda g: "2^%d" % g["log_n"])
def test_oracle_fri_golden(oracle, g):
    w = oracle.root_of_unity(g["log_n"])
    values = oracle.best_fft(random_elems(1 << g["deg_log"], g["seed"]), w, g["log_n"])
    text, ok = oracle.prove_low_degree_json(values, w, (1 << g["log_n"]) // 4, 8)
    assert ok and len(text) == g["json_len"] and sha(text.encode()) == g["json_sha256"]


@pytest.mark.parametrize("name", ["compute", "poseidon3_test"])
def test_oracle_proof_golden(oracle, name, tmp_path):
    """whole-proof hashes: two independent restatements (survey model, C oracle) agree on these
    and the restated verifier (verify.rs) accepts"""
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    out = str(tmp_path / "proof.json")
    rc, _ = oracle.prove_files(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), out)
    assert rc == 0
    assert oracle.sha256_file(out) == GOLD["proofs"][name]["proof_json_sha256"]
    assert os.path.getsize(out) == GOLD["proofs"][name]["proof_json_bytes"]


def test_oracle_poseidon_kats(oracle):
    """the alternative digest against every vector the reference holds for it, in both restatements"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import py_model as pm
    k = GOLD["reference_kats"]["poseidon"]
    for n, want in k["digest"].items():
        msg = bytes(range(int(n))) + bytes(64 - int(n))
        assert oracle.poseidon(msg).hex() == want and pm.poseidon_digest(msg).hex() == want
    root, nodes = oracle.poseidon_merkle_gen_proofs(bytes.fromhex("7fffffff") * 4096, 4, 4096, [2, 7, 13])
    assert root.hex() == k["merkle4096"]["root"]
    assert nodes[0][0].tobytes().hex() == k["merkle4096"]["first_node_of_2"]
    for j, i in enumerate([2, 7, 13]):                       # verify_multi_branch, pallarel_merkle_tree.rs:248
        assert oracle.poseidon_merkle_validate(root, i, bytes.fromhex("7fffffff"), nodes[j])
    assert not oracle.poseidon_merkle_validate(root, 3, bytes.fromhex("7ffffffe"), nodes[0])
    rng = np.random.default_rng(11)
    for ln in (1, 4, 31, 32, 33, 40, 64):                    # the two restatements on random messages of every shape
        m = bytearray(rng.integers(0, 256, ln, dtype=np.uint8).tobytes())
        for top in (31, 63):
            if top < ln:
                m[top] &= 0x3f                               # keep every chunk below the modulus
        assert oracle.poseidon(bytes(m)) == pm.poseidon_digest(bytes(m))
    r_le = pm.BLS_R.to_bytes(32, "little")
    for bad in (b"", bytes(65), r_le, bytes(32) + r_le, b"\xff" * 32):   # where the reference panics (poseidon.rs:33, :48)
        with pytest.raises(ValueError):
            oracle.poseidon(bad)
    assert oracle.poseidon((pm.BLS_R - 1).to_bytes(32, "little")) == pm.poseidon_digest((pm.BLS_R - 1).to_bytes(32, "little"))


def test_oracle_matches_python_model(oracle):
    """the second, independent restatement (oracle/py_model.py) on a small case of every stage"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import py_model as pm
    from stark_pure_rust_b200 import field
    log_n = 7
    w = oracle.root_of_unity(log_n)
    wi = field.from_mont(w.reshape(1, 4))[0]
    v = random_elems(16, 99)
    vals = oracle.best_fft(v, w, log_n)
    assert field.from_mont(vals) == pm.best_fft(field.from_mont(v), wi, log_n)
    text, ok = oracle.prove_low_degree_json(vals, w, 32, 8)
    assert ok and text == pm.fri_json(pm.prove_low_degree(field.from_mont(vals), wi, 32, 8))
    z = random_elems(50, 5); z[::7] = 0
    assert field.from_mont(oracle.multi_inv(z)) == pm.multi_inv(field.from_mont(z))


# ---- product, host side -----------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    from stark_pure_rust_b200 import _lib
    L = _lib.load()
    syms = _lib.header_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert set(L._proto) == set(syms)


def test_public_header_is_plain_c():
    """the drop-in boundary is a C ABI: include/stark_b200.h must compile as C (what a bindgen / cgo / ctypes user feeds it to)"""
    hdr = os.path.join(ROOT, "include", "stark_b200.h")
    for lang, cc in (("c", "gcc"), ("c++", "g++")):
        r = subprocess.run([cc, "-fsyntax-only", "-Wall", "-Werror", "-x", lang, hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_host_blake_and_sampler_match_reference_kats():
    import stark_pure_rust_b200 as sb
    k = GOLD["reference_kats"]
    d = sb.utils.blake(b"hello world")
    assert d.hex() == k["blake"]["hello world"]
    assert sb.utils.blake(d).hex() == k["blake"]["blake(blake(hello world))"]
    for m in (b"", b"a" * 63, b"a" * 64, b"a" * 65, b"a" * 129):
        assert sb.utils.blake(m) == hashlib.blake2s(m).digest()
    assert sb.utils.get_pseudorandom_indices(d, 7, 5, 0) == k["sampler"]["hello world,7,5,0"]
    assert sb.utils.get_pseudorandom_indices(sb.utils.blake(b"hello another world"), 7, 20, 0) == k["sampler"]["hello another world,7,20,0"]


def test_host_poseidon_matches_reference_kats_and_oracle(oracle):
    """the product's host digest (Proof::validate with H = PoseidonDigest): poseidon.rs:66-106, pallarel_merkle_tree.rs:235-246"""
    import stark_pure_rust_b200 as sb
    k = GOLD["reference_kats"]["poseidon"]
    for n, want in k["digest"].items():
        assert sb.utils.poseidon(bytes(range(int(n))) + bytes(64 - int(n))).hex() == want
    h = sb.utils.poseidon(bytes.fromhex("7fffffff"))
    assert h.hex() == k["merkle4096"]["first_node_of_2"]
    for _ in range(12):
        h = sb.utils.poseidon(h + h)
    assert h.hex() == k["merkle4096"]["root"]
    rng = np.random.default_rng(5)
    for ln in list(range(1, 65)):
        m = bytearray(rng.integers(0, 256, ln, dtype=np.uint8).tobytes())
        for top in (31, 63):
            if top < ln:
                m[top] &= 0x3f
        assert sb.utils.poseidon(bytes(m)) == oracle.poseidon(bytes(m))
    r_le = bytes.fromhex("01000000fffffffffe5bfeff02a4bd5305d8a10908d83933487d9d2953a7ed73")
    assert sb.utils.poseidon((int.from_bytes(r_le, "little") - 1).to_bytes(32, "little")) == oracle.poseidon((int.from_bytes(r_le, "little") - 1).to_bytes(32, "little"))
    for bad in (b"", bytes(65), r_le, bytes(32) + r_le, b"\xff" * 64):
        with pytest.raises(ValueError):
            sb.utils.poseidon(bad)
    root, nodes = oracle.poseidon_merkle_gen_proofs(bytes(range(64)) * 2, 16, 8, [5])
    pr = sb.merkle.Proof(bytes(range(64))[16:32], [nodes[0][l].tobytes() for l in range(3)])
    assert pr.validate(root, 5, "poseidon") and not pr.validate(root, 4, "poseidon") and not pr.validate(root, 5)


def test_host_sampler_matches_oracle(oracle):
    import stark_pure_rust_b200 as sb
    for seed, mod, cnt, ex in [(b"a" * 32, 1 << 13, 40, 8), (b"b" * 32, (1 << 24) - 1, 80, 8), (b"c" * 32, 128, 24, 0), (b"d" * 40, 100, 9, 3)]:
        assert sb.utils.get_pseudorandom_indices(seed, mod, cnt, ex) == oracle.get_pseudorandom_indices(seed, mod, cnt, ex)
    with pytest.raises(sb.StarkB200Error):     # fri/src/utils.rs:88 assert
        sb.utils.get_pseudorandom_indices(b"a" * 32, 1 << 24, 4, 0)


def test_merkle_proof_validate_host(oracle):
    """Proof::validate (merkle_tree.rs:25-43) against oracle openings"""
    import stark_pure_rust_b200 as sb
    leaves = [bytes([i]) * 5 for i in range(32)]
    root, nodes = oracle.merkle_gen_proofs(b"".join(leaves), 5, 32, [3, 30])
    for q, i in enumerate([3, 30]):
        pr = sb.merkle.Proof(leaves[i], [nodes[q][l].tobytes() for l in range(5)])
        assert pr.validate(root, i) and not pr.validate(root, i ^ 1)


def test_no_cpu_fallback():
    """without a B200 the product refuses to run instead of computing on the host"""
    import stark_pure_rust_b200 as sb
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(sb.StarkB200Error) as e:
        sb.Context(0)
    assert e.value.code == -1
    # and nothing in the product imports the oracle
    pkg = os.path.join(ROOT, "stark_pure_rust_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert "liboracle" not in text and "oracle_bind" not in text and "py_model" not in text, f


def test_generated_field_code_is_current_and_emulator_passes():
    for gen in ("gen_fp.py", "gen_fp_bls.py"):         # the prover's field, the Poseidon digest's field
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", gen), "--check"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr


def test_ntt_plan_model():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ntt_plan_model.py")], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr


def test_bench_reference_arm_runs():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-log-n", "12"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"


def test_proof_json_reader_round_trip(oracle, tmp_path):
    """sb_stark_proof_from_json (serde_json::from_reader::<StarkProof>, run.rs:578) followed by sb_stark_proof_json
    reproduces the oracle's proof.json byte for byte; malformed text is refused.  Host only: no GPU needed."""
    import ctypes as C
    from stark_pure_rust_b200 import _lib
    L = _lib.load()
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    for name in ("compute", "poseidon3_test"):            # one thread; pieces formatted on several threads (2.5 MB)
        path = str(tmp_path / (name + ".json"))
        rc, _ = oracle.prove_files(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), path)
        assert rc == 0
        text = open(path, "rb").read()
        assert hashlib.sha256(text).hexdigest() == GOLD["proofs"][name]["proof_json_sha256"]
        h = C.c_void_p()
        assert L.sb_stark_proof_from_json(text, len(text), C.byref(h)) == 0
        for _ in range(3):
            n = C.c_size_t()
            s = L.sb_stark_proof_json(h, C.byref(n))
            try:
                assert C.string_at(s, n.value) == text
            finally:
                L.sb_free_string(s)
        L.sb_stark_proof_free(h)
    for bad in (text[:-1], text.replace(b'"l_root"', b'"x_root"'), text.replace(b"[", b"[256,", 1), b"{}", b""):
        h = C.c_void_p()
        assert L.sb_stark_proof_from_json(bad, len(bad), C.byref(h)) == -3


@pytest.mark.parametrize("name", ["compute", "poseidon3_test", "bits", "pedersen_test"])
def test_host_front_end_matches_oracle(oracle, name):
    """the product's parsers + trace arrangement (csrc/frontend.cu, threaded, in-place r1cs decode) against the oracle's
    restatement of run.rs:109-452 on the bundled circuits: every array of the mk_r1cs_proof arguments is identical"""
    import stark_pure_rust_b200 as sb
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    got = sb.prove.trace_from_files(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"))
    want = oracle.trace_from_files(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"))
    assert got["original_steps"] == want["original_steps"]
    for k in ("witness_trace", "computational_trace", "coefficients", "flag0", "flag1", "flag2", "permuted_indices", "public_wires", "pfi_k", "pfi_w"):
        assert np.array_equal(got[k], want[k]), k


def test_host_front_end_synthetic_and_errors(oracle, tmp_path):
    """seeded synthetic circuits (multi-term C, linear constraints, padding rows, 60 public inputs) and malformed files"""
    import stark_pure_rust_b200 as sb
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_r1cs
    for n_constraints, avg_terms, seed, n_pub in ((40, 2.0, 3, 2), (1200, 4.0, 7, 2), (400, 3.0, 11, 60)):
        prefix = str(tmp_path / ("syn%d" % seed))
        wit, cons = gen_r1cs.generate(n_constraints, avg_terms, n_pub, seed)
        gen_r1cs.write_files(prefix, wit, cons, n_pub)
        got = sb.prove.trace_from_files(prefix + ".r1cs", prefix + ".wtns")
        want = oracle.trace_from_files(prefix + ".r1cs", prefix + ".wtns")
        for k in want:
            assert np.array_equal(got[k], want[k]), (seed, k)
    bad = tmp_path / "bad.r1cs"
    bad.write_bytes(b"r1cs" + bytes(40))
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    with pytest.raises(sb.StarkB200Error):
        sb.prove.trace_from_files(bad, os.path.join(d, "compute.wtns"))
    with pytest.raises(sb.StarkB200Error):
        sb.prove.trace_from_files(os.path.join(d, "compute.r1cs"), bad)
    trunc = tmp_path / "trunc.r1cs"
    trunc.write_bytes(open(os.path.join(d, "poseidon3_test.r1cs"), "rb").read()[:5000])
    with pytest.raises(sb.StarkB200Error):
        sb.prove.trace_from_files(trunc, os.path.join(d, "poseidon3_test.wtns"))


def test_product_fri_verifier_on_oracle_proofs(oracle):
    """sb_fri_verify_json (the product's restatement of fri.rs:226-404, host only) accepts the oracle prover's proofs of
    low-degree vectors, and rejects tampered proofs, a wrong Merkle root and a tighter degree bound (the reference prover
    itself panics on a vector that is not of low degree, so that case cannot be produced)"""
    import json
    import stark_pure_rust_b200 as sb
    from stark_pure_rust_b200 import field
    for log_n, deg_log in ((7, 4), (10, 7), (12, 9)):
        n = 1 << log_n
        w = field.root_of_unity(log_n)
        wl = field.mont_scalar(w)
        values = oracle.best_fft(random_elems(1 << deg_log, 600 + log_n), wl, log_n)
        text, ok = oracle.prove_low_degree_json(values, wl, n // 4, 8)
        assert ok
        root, _ = oracle.merkle_gen_proofs(oracle.fp_to_bytes_le(values).tobytes(), 32, n, [])
        assert sb.fri.verify_low_degree_proof(root, w, text, n // 4, 8, n)
        proof = json.loads(text)

        def rejected(p, mroot=root, bound=n // 4):
            with pytest.raises(sb.StarkB200Error) as e:
                sb.fri.verify_low_degree_proof(mroot, w, json.dumps(p, separators=(",", ":")), bound, 8, n)
            return e.value.code == -6

        assert rejected(proof, mroot=bytes(32))
        if len(proof) > 1:
            p = json.loads(text); p[0]["Middle"]["column_branches"][0]["leaf"][0] ^= 1
            assert rejected(p)
            p = json.loads(text); p[0]["Middle"]["poly_branches"][7]["nodes"][0][3] ^= 1
            assert rejected(p)
            p = json.loads(text); p[0]["Middle"]["root2"][31] ^= 0x80
            assert rejected(p)
        p = json.loads(text); p[-1]["Last"]["last"][3][0] ^= 1
        assert rejected(p)
        p = json.loads(text); p[-1]["Last"]["last"].pop()
        assert rejected(p)
        # the same proof does not pass for a tighter degree bound
        assert rejected(proof, bound=n // 16)
