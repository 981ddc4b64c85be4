"""ONE job over several devices behind the C ABI (sb_init_multi, sb_ext_*, sb_prove_r1cs): every result must
equal the single-GPU / oracle result bit for bit.

The sharded code path is selected by the number of devices in the context, not by the number of physical GPUs: a context may
list the same ordinal several times (logical devices on one GPU), so the whole partitioning / peer-store / subtree / top-of-
tree / opening logic for g = 2, 4, 8 is exercised on a single-GPU box; the *_real_gpus tests repeat it on distinct GPUs when
the box has them."""
import ctypes as C
import hashlib
import json
import os
import threading

import numpy as np
import pytest

from conftest import ROOT, random_elems

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module", params=[1, 2, 4, 8])
def mctx(request):
    import stark_pure_rust_b200 as sb
    c = sb.Context(devices=[0] * request.param)
    yield c
    c.close()


def _ext_chain(ctx, oracle, log_s, n_cols, col_len, seed):
    import stark_pure_rust_b200 as sb
    from stark_pure_rust_b200 import field
    S, N = 1 << log_s, 8 << log_s
    g2, g1 = oracle.root_of_unity(log_s + 3), oracle.root_of_unity(log_s)
    cols = random_elems(n_cols * col_len, seed).reshape(n_cols, col_len, 4)
    e = sb.ext.ExtColumns(n_cols, log_s, ctx=ctx)
    e.load(0, cols)
    e.extend()
    want = []
    for c in range(n_cols):
        padded = np.zeros((S, 4), dtype=np.uint64)
        padded[:col_len] = cols[c]
        want.append(oracle.best_fft(oracle.best_fft(padded, g1, log_s, inverse=True), g2, log_s + 3))
        assert np.array_equal(e.read(c), want[c]), "LDE column %d" % c
    # prove.rs:235-264: tree over the rows of up to eight columns
    k = min(8, n_cols)
    ids = list(range(k))
    rows = np.concatenate([oracle.fp_to_bytes_le(want[c]).reshape(N, 1, 32) for c in ids], axis=1).tobytes()
    idx = [0, N - 1, 5, 5, N // 2 + 3, N // 8 - 1, N // 8, 7 * N // 8 + 1] + [int(x) for x in np.random.default_rng(seed).integers(0, N, 24)]
    root_want, nodes_want = oracle.merkle_gen_proofs(rows, 32 * k, N, idx)
    root, tree = e.commit(ids)
    assert root == root_want
    proofs = e.open(tree, idx)
    for q, i in enumerate(idx):
        assert proofs[q].leaf == rows[i * 32 * k:(i + 1) * 32 * k], "leaf %d" % i
        assert b"".join(proofs[q].nodes) == nodes_want[q].tobytes(), "path %d" % i
    e.free_tree(tree)
    # prove.rs:324-332 + :367: one-column tree and the low-degree proof on it, with and without the committed tree
    last = n_cols - 1
    root1, tree1 = e.commit([last])
    r1_want, _ = oracle.merkle_gen_proofs(oracle.fp_to_bytes_le(want[last]).tobytes(), 32, N, [])
    assert root1 == r1_want
    text_want, ok = oracle.prove_low_degree_json(want[last], g2, N // 4, 8)
    assert ok
    assert e.fri_prove(last, N // 4, 8, tree=tree1, as_json=True) == text_want
    assert e.fri_prove(last, N // 4, 8, tree=None, as_json=True) == text_want
    e.free_tree(tree1)
    e.close()


@pytest.mark.parametrize("log_s,n_cols,col_len", [(10, 3, 1024), (11, 9, 2000), (13, 8, 6684), (5, 2, 32), (3, 1, 8), (17, 2, 100000)])
def test_ext_chain_matches_oracle(mctx, oracle, log_s, n_cols, col_len):
    _ext_chain(mctx, oracle, log_s, n_cols, col_len, 9000 + log_s + n_cols)


def test_ext_chain_2_24_golden(mctx):
    """the headline size on g devices: 8-column LDE 2^21 -> 2^24, both trees and the FRI proof against the oracle's digests
    (tests/golden/vectors_large.json).  At this size the first FRI layer stays sharded (column next to the data, column tree as
    per-device subtrees with the staged digest exchange); the second one is gathered to the primary device."""
    import stark_pure_rust_b200 as sb
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors_large.json")))
    log_s, nc = gold["log_s"], gold["n_cols"]
    S, N = 1 << log_s, 8 << log_s
    e = sb.ext.ExtColumns(nc, log_s, ctx=mctx)
    e.load(0, random_elems(nc * S, gold["lde"]["seed"]).reshape(nc, S, 4))
    e.extend()
    assert hashlib.sha256(e.read(nc - 1).tobytes()).hexdigest() == gold["lde"]["col_sha256"][nc - 1]
    root8, t8 = e.commit(list(range(nc)))
    assert root8.hex() == gold["merkle8"]["root"]
    proofs = e.open(t8, gold["merkle8"]["open"])
    assert hashlib.sha256(b"".join(p.leaf for p in proofs)).hexdigest() == gold["merkle8"]["leaves_sha256"]
    assert hashlib.sha256(b"".join(b"".join(p.nodes) for p in proofs)).hexdigest() == gold["merkle8"]["nodes_sha256"]
    e.free_tree(t8)
    root1, t1 = e.commit([nc - 1])
    assert root1.hex() == gold["merkle1"]["root"]
    text = e.fri_prove(nc - 1, N // 4, 8, tree=t1, as_json=True)
    e.free_tree(t1)
    e.close()
    assert len(text) == gold["fri"]["json_len"] and hashlib.sha256(text.encode()).hexdigest() == gold["fri"]["json_sha256"]


def test_ext_direct_fri_and_errors(mctx, oracle):
    """max_deg_plus_1 <= 16: the proof is the values themselves (fri.rs:88-112), gathered in natural order"""
    import stark_pure_rust_b200 as sb
    log_s = 10
    N = 8 << log_s
    g2 = oracle.root_of_unity(log_s + 3)
    # a column of degree < 16: the S-point evaluations of 16 random coefficients
    col = oracle.best_fft(random_elems(16, 31), oracle.root_of_unity(log_s), log_s).reshape(1, 1 << log_s, 4)
    e = sb.ext.ExtColumns(1, log_s, ctx=mctx)
    e.load(0, col)
    e.extend()
    vals = e.read(0)
    text_want, ok = oracle.prove_low_degree_json(vals, g2, 16, 8)
    assert e.fri_prove(0, 16, 8, as_json=True) == text_want
    with pytest.raises(sb.StarkB200Error):
        e.commit([3])
    with pytest.raises(sb.StarkB200Error):
        e.load(0, np.zeros((2, 8, 4), dtype=np.uint64))
    e.close()


@pytest.mark.parametrize("name", ["compute", "poseidon3_test", "pedersen_test"])
def test_prove_sharded_matches_golden(mctx, name, tmp_path):
    """mk_r1cs_proof over g (logical) devices: proof.json byte-identical to the oracle's (golden hashes)"""
    import stark_pure_rust_b200 as sb
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))["proofs"][name]
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    out = str(tmp_path / "proof.json")
    sb.prove.prove_with_file_path(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), out, ctx=mctx)
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == gold["proof_json_sha256"]
    sb.prove.verify_with_file_path(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), out, ctx=mctx)


def test_prove_sharded_many_public_wires(mctx, oracle, tmp_path):
    """the coset-transform path of the boundary polynomials (more than 23 public wires) on every device count"""
    import sys
    import stark_pure_rust_b200 as sb
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_r1cs
    prefix = str(tmp_path / "syn")
    wit, cons = gen_r1cs.generate(1500, 3.0, 60, 5)
    gen_r1cs.write_files(prefix, wit, cons, 60)
    out, want = str(tmp_path / "proof.json"), str(tmp_path / "oracle.json")
    rc, _ = oracle.prove_files(prefix + ".r1cs", prefix + ".wtns", want, verify=False)
    assert rc == 0
    sb.prove.prove_with_file_path(prefix + ".r1cs", prefix + ".wtns", out, ctx=mctx)
    assert open(out, "rb").read() == open(want, "rb").read()


def test_prove_sharded_rejects_bad_witness(mctx, tmp_path):
    import stark_pure_rust_b200 as sb
    d = os.path.join(ROOT, "tests", "golden", "circuits")
    w = bytearray(open(os.path.join(d, "poseidon3_test.wtns"), "rb").read())
    w[-40] ^= 1
    bad = str(tmp_path / "bad.wtns")
    open(bad, "wb").write(bytes(w))
    with pytest.raises(sb.StarkB200Error):
        sb.prove.prove_with_file_path(os.path.join(d, "poseidon3_test.r1cs"), bad, None, ctx=mctx)


# ---- distinct GPUs (run with gpurun --gpus N) ---------------------------------------------------------------------------
@pytest.mark.parametrize("g", [2, 4, 8])
def test_ext_chain_real_gpus(oracle, g):
    if n_gpus() < g:
        pytest.skip("needs %d GPUs" % g)
    import stark_pure_rust_b200 as sb
    ctx = sb.Context(devices=list(range(g)))
    try:
        _ext_chain(ctx, oracle, 13, 9, 6684, 77)
        _ext_chain(ctx, oracle, 10, 8, 1024, 78)
        _ext_chain(ctx, oracle, 17, 2, 1 << 17, 79)        # three FRI layers stay sharded
    finally:
        ctx.close()


# ---- ONE transform over several devices (sb_ntt on a multi-device context, sb_ntt_multi_dev) ----------
def _ntt_multi_checks(ctx, ctx1, oracle, log_ns):
    """best_fft / inv_best_fft through the multi-device context against the single-device result (and the oracle at 2^20)"""
    import stark_pure_rust_b200 as sb
    for log_n in log_ns:
        n = 1 << log_n
        w = oracle.root_of_unity(log_n)
        v = random_elems(n, 40 + log_n)
        got = sb.fft.best_fft(v, w, log_n, ctx=ctx)
        want = sb.fft.best_fft(v, w, log_n, ctx=ctx1)
        assert np.array_equal(got, want), "forward 2^%d" % log_n
        if log_n == 20:
            assert np.array_equal(want, oracle.best_fft(v, w, log_n))
        back = sb.fft.inv_best_fft(got, w, log_n, ctx=ctx)
        assert np.array_equal(back, v), "round trip 2^%d" % log_n
        short = v[: n // 2 + 12345]                                     # zero padding, fft.rs:335-338
        assert np.array_equal(sb.fft.best_fft(short, w, log_n, ctx=ctx), sb.fft.best_fft(short, w, log_n, ctx=ctx1))
        assert np.array_equal(sb.fft.best_fft(v[:3], w, log_n, ctx=ctx), sb.fft.best_fft(v[:3], w, log_n, ctx=ctx1))


def test_ntt_multi_logical(mctx, ctx, oracle):
    g = mctx.lib.sb_device_count(mctx.h)
    _ntt_multi_checks(mctx, ctx, oracle, [20, 21, 22] if g > 1 else [20])


def test_ntt_multi_with_a_larger_table_cached_on_one_device(oracle, ctx):
    """the primary serves the root from a strided view of a larger table it already holds, the other devices build the
    small table: every device must index its own table geometry"""
    import stark_pure_rust_b200 as sb
    m = sb.Context(devices=[0] * 4)
    try:
        w23 = oracle.root_of_unity(23)
        cols = random_elems(2 * (1 << 20), 5).reshape(2, 1 << 20, 4)
        sb.fft.lde_batch(cols, w23, 20, 3, ctx=m)                       # primary only: caches the 2^23 table of w23 there
        w20 = oracle.root_of_unity(20)                                  # = w23^8
        v = random_elems(1 << 20, 6)
        assert np.array_equal(sb.fft.best_fft(v, w20, 20, ctx=m), sb.fft.best_fft(v, w20, 20, ctx=ctx))
        assert np.array_equal(sb.fft.inv_best_fft(v, w20, 20, ctx=m), sb.fft.inv_best_fft(v, w20, 20, ctx=ctx))
    finally:
        m.close()


def test_ntt_multi_dev_slabs(oracle, ctx):
    """sb_ntt_multi_dev on slabs the caller placed (sb_dev_alloc_on), and its argument checks"""
    import stark_pure_rust_b200 as sb
    from stark_pure_rust_b200._lib import _ptr
    g, log_n = 4, 20
    n = 1 << log_n
    m = sb.Context(devices=[0] * g)
    try:
        w = oracle.root_of_unity(log_n)
        v = random_elems(n, 7)
        slabs = (C.c_void_p * g)()
        for d in range(g):
            p = C.c_void_p()
            m.check(m.lib.sb_dev_alloc_on(m.h, d, (n // g) * 32, C.byref(p)))
            slabs[d] = p
            part = np.ascontiguousarray(v[d * (n // g):(d + 1) * (n // g)])
            m.check(m.lib.sb_h2d(m.h, p, _ptr(part), part.nbytes))
        root = np.ascontiguousarray(w, dtype=np.uint64).reshape(4)
        m.check(m.lib.sb_ntt_multi_dev(m.h, slabs, _ptr(root), log_n, 0))
        out = np.empty_like(v)
        for d in range(g):
            part = np.empty((n // g, 4), dtype=np.uint64)
            m.check(m.lib.sb_d2h(m.h, _ptr(part), slabs[d], part.nbytes))
            out[d * (n // g):(d + 1) * (n // g)] = part
        assert np.array_equal(out, sb.fft.best_fft(v, w, log_n, ctx=ctx))
        with pytest.raises(sb.StarkB200Error):
            m.check(m.lib.sb_ntt_multi_dev(m.h, slabs, _ptr(root), 19, 0))          # below the multi-device range
        with pytest.raises(sb.StarkB200Error):
            m.check(m.lib.sb_ntt_multi_dev(ctx.h, slabs, _ptr(root), log_n, 0))     # single-device context
        with pytest.raises(sb.StarkB200Error):
            m.check(m.lib.sb_dev_alloc_on(m.h, g, 32, C.byref(C.c_void_p())))
        for d in range(g):
            m.lib.sb_dev_free(m.h, slabs[d])
    finally:
        m.close()


@pytest.mark.parametrize("g", [2, 4, 8])
def test_ntt_multi_real_gpus(g, ctx, oracle):
    if n_gpus() < g:
        pytest.skip("needs %d GPUs" % g)
    import stark_pure_rust_b200 as sb
    m = sb.Context(devices=list(range(g)))
    try:
        _ntt_multi_checks(m, ctx, oracle, [20, 23])
    finally:
        m.close()


@pytest.mark.parametrize("g", [2, 4, 8])
def test_prove_real_gpus(g, tmp_path):
    if n_gpus() < g:
        pytest.skip("needs %d GPUs" % g)
    import stark_pure_rust_b200 as sb
    ctx = sb.Context(devices=list(range(g)))
    try:
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))["proofs"]
        d = os.path.join(ROOT, "tests", "golden", "circuits")
        out = str(tmp_path / "proof.json")
        for name in ("poseidon3_test", "pedersen_test", "bits"):
            sb.prove.prove_with_file_path(os.path.join(d, name + ".r1cs"), os.path.join(d, name + ".wtns"), out, ctx=ctx)
            assert hashlib.sha256(open(out, "rb").read()).hexdigest() == gold[name]["proof_json_sha256"], name
    finally:
        ctx.close()


def test_context_used_from_another_thread():
    """the current CUDA device is per host thread; every entry point must select the context's device itself.
    A context on the last GPU is created here and used from a fresh thread (whose current device is 0)."""
    import stark_pure_rust_b200 as sb
    from stark_pure_rust_b200 import field
    dev = n_gpus() - 1
    ctx = sb.Context(dev)
    v = random_elems(1 << 12, 5)
    w = field.root_of_unity(12)
    res = {}

    def work():
        try:
            import torch
            if dev:
                torch.cuda.set_device(0)
            y = sb.fft.best_fft(v, w, 12, ctx=ctx)
            res["ok"] = np.array_equal(sb.fft.inv_best_fft(y, w, 12, ctx=ctx), v)
        except Exception as ex:      # noqa: BLE001
            res["err"] = repr(ex)

    th = threading.Thread(target=work)
    th.start()
    th.join()
    ctx.close()
    assert res.get("ok"), res
