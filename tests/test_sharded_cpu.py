"""world_size-2 (and 4) gloo tests of the sharded commitment's host logic (stark_pure_rust_b200/sharded.py):
column ownership, the column->row exchange plan, subtree-root gather, top-of-tree combination and opening-path
assembly.  The field / hash work is done by a CHECKER backend built on the CPU oracle (tests may use the oracle;
the product's only backend is CudaBackend), so the result is compared with the single-process oracle tree."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def fri_layers_from_json(text):
    """serde text of Vec<FriProof> -> the layout of stark_pure_rust_b200.fri.unpack_proof"""
    import json
    from stark_pure_rust_b200.merkle import Proof
    out = []
    for layer in json.loads(text):
        if "Last" in layer:
            out.append({"Last": {"last": [bytes(v) for v in layer["Last"]["last"]]}})
        else:
            m = layer["Middle"]
            conv = lambda bs: [Proof(bytes(b["leaf"]), [bytes(x) for x in b["nodes"]]) for b in bs]
            out.append({"Middle": {"root2": bytes(m["root2"]), "column_branches": conv(m["column_branches"]),
                                   "poly_branches": conv(m["poly_branches"])}})
    return out


class OracleBackend:
    """same duck type as sharded.CudaBackend, CPU tensors + oracle arithmetic"""

    def __init__(self):
        import torch
        import oracle_bind as ob
        self.torch, self.ob = torch, ob

    def empty(self, *shape):
        return self.torch.zeros(*shape, dtype=self.torch.int64)

    def from_numpy(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy())

    def _np(self, t):
        return t.contiguous().numpy().view(np.uint64)

    def lde(self, cols, root_big, log_s, log_ext):
        ob = self.ob
        g1 = np.array(root_big, dtype=np.uint64)
        g_small = ob.root_of_unity(log_s)
        out = self.empty(cols.shape[0], 1 << (log_s + log_ext), 4)
        for k in range(cols.shape[0]):
            coef = ob.best_fft(self._np(cols[k]), g_small, log_s, inverse=True)
            out[k] = self.from_numpy(ob.best_fft(coef, g1, log_s + log_ext))
        return out

    def _leaves(self, cols):
        n = cols[0].shape[0]
        return np.concatenate([self.ob.fp_to_bytes_le(self._np(c)).reshape(n, 1, 32) for c in cols], axis=1).tobytes()

    def commit_cols(self, cols):
        n = cols[0].shape[0]
        leaves = self._leaves(cols)
        root, _ = self.ob.merkle_gen_proofs(leaves, 32 * len(cols), n, [])
        return root, (leaves, n, 32 * len(cols))

    def open(self, tree, idx):
        leaves, n, lb = tree
        _, nodes = self.ob.merkle_gen_proofs(leaves, lb, n, idx)
        return [(leaves[i * lb:(i + 1) * lb], [nodes[q, l].tobytes() for l in range(nodes.shape[1])]) for q, i in enumerate(idx)]

    def free(self, tree):
        pass

    def ntt_batch(self, x, root_int, log_n, inverse=False):
        from stark_pure_rust_b200 import field
        root = field.mont_scalar(root_int)
        out = self.empty(*x.shape)
        for k in range(x.shape[0]):
            out[k] = self.from_numpy(self.ob.best_fft(self._np(x[k]), root, log_n, inverse=inverse))
        return out

    def twiddle_mul(self, x, row0, root_int, log_n, inverse=False):
        from stark_pure_rust_b200 import field
        w = pow(root_int, -1, field.P) if inverse else root_int
        rows, cols = x.shape[0], x.shape[1]
        vals = field.from_mont(self._np(x).reshape(-1, 4))
        out = [v * pow(w, ((row0 + i // cols) * (i % cols)) % (1 << log_n), field.P) % field.P for i, v in enumerate(vals)]
        x.copy_(self.from_numpy(field.to_mont(out)).reshape(rows, cols, 4))

    def fri_fold(self, vals, root_limbs, values_root):
        """fri.rs:135-164 with Python big ints: Lagrange through the four row points, evaluated at special_x"""
        from stark_pure_rust_b200 import field
        Pm = field.P
        v = field.from_mont(self._np(vals))
        n, q = len(v), len(v) // 4
        w = field.from_mont(np.asarray(root_limbs, dtype=np.uint64).reshape(1, 4))[0]
        sx = int.from_bytes(values_root, "little") % Pm
        iota = pow(w, q, Pm)
        out = []
        for i in range(q):
            xs = [pow(w, i, Pm) * pow(iota, j, Pm) % Pm for j in range(4)]
            acc = 0
            for a in range(4):
                num = den = 1
                for b in range(4):
                    if a != b:
                        num = num * (sx - xs[b]) % Pm
                        den = den * (xs[a] - xs[b]) % Pm
                acc = (acc + v[i + q * a] * num * pow(den, -1, Pm)) % Pm
            out.append(acc)
        return self.from_numpy(field.to_mont(out))

    def fri_rest(self, col, root_limbs, max_deg_plus_1, excl, tree):
        text, _ = self.ob.prove_low_degree_json(self._np(col), np.asarray(root_limbs, dtype=np.uint64), max_deg_plus_1, excl, verify=False)
        return fri_layers_from_json(text)

    def root_tensor(self, root):
        return self.torch.frombuffer(bytearray(root), dtype=self.torch.uint8).clone()

    def bytes_tensor(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint8).copy())

    def leaf_bytes(self, tree):
        return tree[2]


def _worker(rank, world, port, n_cols, log_s, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import random_elems
        import oracle_bind as ob
        from stark_pure_rust_b200 import sharded
        be = OracleBackend()
        S, N = 1 << log_s, 1 << (log_s + 3)
        g2 = ob.root_of_unity(log_s + 3)
        cols = random_elems(n_cols * S, 4242).reshape(n_cols, S, 4)          # every rank derives the same inputs
        mine = sharded.owned_columns(n_cols, world, rank)
        ext = be.lde(be.from_numpy(cols[mine]) if mine else be.empty(0, S, 4), g2, log_s, 3)
        sc = sharded.ShardedCommitter(be, dist)
        rows = sc.exchange({c: ext[k] for k, c in enumerate(mine)}, n_cols, N)
        tree = sc.commit_rows(rows, list(range(n_cols)), N)
        idx = [0, N - 1, N // 2, 5, 5, N // world, N // world - 1, 17]
        proofs = tree.gen_proofs(idx, dist)
        if rank == 0:
            # single-process reference: the whole tree in one oracle call
            full = be.lde(be.from_numpy(cols), g2, log_s, 3)
            leaves = be._leaves([full[c] for c in range(n_cols)])
            root, nodes = ob.merkle_gen_proofs(leaves, 32 * n_cols, N, idx)
            lb = 32 * n_cols
            ok = tree.get_root() == root
            for q, i in enumerate(idx):
                leaf, path = proofs[q]
                ok = ok and leaf == leaves[i * lb:(i + 1) * lb] and path == [nodes[q, l].tobytes() for l in range(nodes.shape[1])]
            ret.put(bool(ok))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _ntt_worker(rank, world, port, log_n, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import random_elems
        import oracle_bind as ob
        from stark_pure_rust_b200 import field, sharded
        be = OracleBackend()
        n = 1 << log_n
        x = random_elems(n, 99 + log_n)
        w = field.root_of_unity(log_n)
        lo, hi = sharded.row_range(n, world, rank)
        ok = True
        for inverse in (False, True):
            got = sharded.distributed_ntt(be, be.from_numpy(x[lo:hi]), w, log_n, inverse, dist)
            want = ob.best_fft(x, field.mont_scalar(w), log_n, inverse=inverse)
            ok = ok and np.array_equal(be._np(got), want[lo:hi])
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if rank == 0:
            ret.put(all(flags))
    finally:
        dist.destroy_process_group()


def _fri_worker(rank, world, port, log_n, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import random_elems
        import oracle_bind as ob
        from stark_pure_rust_b200 import field, sharded
        be = OracleBackend()
        n = 1 << log_n
        w = field.root_of_unity(log_n)
        coeffs = random_elems(n // 8, 1234)
        values = ob.best_fft(coeffs, field.mont_scalar(w), log_n)          # degree < n/8 <= n/4
        lo, hi = sharded.row_range(n, world, rank)
        rows = be.from_numpy(values).reshape(1, n, 4)[:, lo:hi]
        sc = sharded.ShardedCommitter(be, dist)
        tree = sc.commit_rows(rows, [0], n)
        owner = world - 1
        vals = be.from_numpy(values) if rank == owner else None
        proof = sharded.prove_low_degree_sharded(be, tree, vals, owner, w, n, n // 4, 8, dist, replicate=True)
        if rank == 0:
            text, ok = ob.prove_low_degree_json(values, field.mont_scalar(w), n // 4, 8)
            ret.put(bool(ok) and proof == fri_layers_from_json(text))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_fri_matches_single_process(world):
    """prove_low_degree_sharded (layer 0 on the row-sharded values tree, the rest on the owner) under gloo: the assembled
    proof equals the oracle's prove_low_degree of the whole vector"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 33500 + (os.getpid() % 2000) + world * 13
    procs = [ctx.Process(target=_fri_worker, args=(r, world, port, 9, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) is True


@pytest.mark.parametrize("world,log_n", [(2, 6), (2, 9), (4, 8)])
def test_distributed_ntt_matches_single_process(world, log_n):
    """four-step transform over `world` ranks (three all_to_all exchanges): every rank's slab of the result equals the
    oracle's best_fft / inv_best_fft of the whole vector"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world * 11 + log_n
    procs = [ctx.Process(target=_ntt_worker, args=(r, world, port, log_n, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) is True


@pytest.mark.parametrize("world,n_cols", [(2, 3), (2, 8), (4, 5)])
def test_sharded_commit_matches_single_process(world, n_cols):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world * 7 + n_cols
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cols, 5, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) is True


def test_partitioning_and_top_of_tree():
    from stark_pure_rust_b200 import sharded, utils
    assert sharded.owned_columns(10, 4, 1) == [1, 5, 9]
    assert sharded.row_range(1 << 10, 4, 3) == (768, 1024)
    with pytest.raises(ValueError):
        sharded.row_range(1 << 10, 3, 0)
    roots = [bytes([i]) * 32 for i in range(4)]
    root, levels = sharded.combine_roots(roots)
    l01, l23 = utils.blake(roots[0] + roots[1]), utils.blake(roots[2] + roots[3])
    assert root == utils.blake(l01 + l23)
    assert sharded.top_path(levels, 2) == [roots[3], l01]
    assert sharded.combine_roots([roots[0]])[0] == roots[0]
