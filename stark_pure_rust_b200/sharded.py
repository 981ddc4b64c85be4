"""One commitment sharded over the GPUs of a node.

The reference is single-process; what it computes in `mk_r1cs_proof` shards like this:

  1. the columns' low-degree extensions are independent (prove.rs:100-124)      -> column c on rank c % world
  2. a Merkle leaf is one ROW of all committed columns (prove.rs:235-258)       -> one exchange: every rank sends
     the row range [j N/g, (j+1) N/g) of each column it extended to rank j (grouped NCCL send/recv over NVLink)
  3. rank r hashes the subtree over its contiguous row range                    -> all_gather of g 32-byte roots,
     the top log2(g) levels are finished on every rank (identical result, no broadcast needed)
  4. an opening of leaf i = path inside the owner's subtree || path through the top levels

One process per GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the only transport.  All
field / hash work goes through a backend object: `CudaBackend` (the C ABI of libstark_b200.so on torch CUDA
tensors) is the only backend in this package -- there is no CPU fallback; the CPU tests inject their own
checker backend to exercise the partitioning / exchange / path-assembly logic with gloo.
"""
import ctypes as C

import numpy as np

from . import utils


# ---- partitioning ------------------------------------------------------------------------------------
def owned_columns(n_cols, world, rank):
    """columns extended by `rank` (round robin: column c lives on rank c % world)"""
    return list(range(rank, n_cols, world))


def row_range(n, world, rank):
    """contiguous row range hashed by `rank`; world must divide n (both are powers of two)"""
    if world <= 0 or world & (world - 1) or n % world:
        raise ValueError("world size %d must be a power of two dividing %d" % (world, n))
    q = n // world
    return rank * q, (rank + 1) * q


def combine_roots(roots):
    """top of the tree over the subtree roots (merkle_proof_in_place.rs:78-98: parent = H(left || right)).
    Returns (root, levels) with levels[0] = the subtree roots, levels[-1] = [root]."""
    g = len(roots)
    if g == 0 or g & (g - 1):
        raise ValueError("number of subtrees must be a power of two")
    levels = [[bytes(r) for r in roots]]
    while len(levels[-1]) > 1:
        cur = levels[-1]
        levels.append([utils.blake(cur[2 * i] + cur[2 * i + 1]) for i in range(len(cur) // 2)])
    return levels[-1][0], levels


def top_path(levels, rank):
    """sibling digests from the subtree root of `rank` up to (excluding) the root (merkle_tree.rs:25-43 order)"""
    out, i = [], rank
    for lv in levels[:-1]:
        out.append(lv[i ^ 1])
        i >>= 1
    return out


# ---- backends ----------------------------------------------------------------------------------------
class CudaBackend:
    """field / hash work on torch CUDA tensors through the C ABI.  Tensors are int64 views of the (n, 4) u64
    Montgomery limbs; the library runs on torch's current stream so NCCL and kernels stay ordered."""

    def __init__(self, ctx, device):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.device = device
        ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)

    def empty(self, *shape):
        return self.torch.empty(*shape, dtype=self.torch.int64, device=self.device)

    def from_numpy(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(self.device)

    def lde(self, cols, root_big_limbs, log_s, log_ext):
        """(k, col_len, 4) -> (k, 2^(log_s+log_ext), 4): the inv_best_fft -> best_fft pairs of prove.rs:100-124"""
        k, col_len = cols.shape[0], cols.shape[1]
        out = self.empty(k, 1 << (log_s + log_ext), 4)
        if k:
            root = np.ascontiguousarray(root_big_limbs, dtype=np.uint64)
            self.ctx.check(self.ctx.lib.sb_lde_batch_dev(self.ctx.h, C.c_void_p(cols.data_ptr()), k, col_len, col_len,
                                                         C.c_void_p(root.ctypes.data), log_s, log_ext, C.c_void_p(out.data_ptr())))
        return out

    def commit_cols(self, cols):
        """tree over leaves = rows of the given (n, 4) column tensors (to_bytes_le of each, concatenated)"""
        n = cols[0].shape[0]
        ptrs = (C.c_void_p * len(cols))(*[c.data_ptr() for c in cols])
        root = np.empty(32, dtype=np.uint8)
        t = C.c_void_p()
        self.ctx.check(self.ctx.lib.sb_merkle_commit_cols_dev(self.ctx.h, ptrs, len(cols), n, C.c_void_p(root.ctypes.data), C.byref(t)))
        return root.tobytes(), (t, n, 32 * len(cols), cols)       # the columns must outlive the tree

    def open(self, tree, idx):
        t, n, lb, _ = tree
        q = len(idx)
        depth = n.bit_length() - 1
        leaves = np.empty(q * lb, dtype=np.uint8)
        nodes = np.empty(max(q * depth * 32, 1), dtype=np.uint8)
        if q:
            ia = np.asarray(idx, dtype=np.uint64)
            self.ctx.check(self.ctx.lib.sb_merkle_open(self.ctx.h, t, ia.ctypes.data_as(C.POINTER(C.c_size_t)), q,
                                                       C.c_void_p(leaves.ctypes.data), C.c_void_p(nodes.ctypes.data) if depth else None))
        lv, nd = leaves.tobytes(), nodes.tobytes()
        return [(lv[i * lb:(i + 1) * lb], [nd[(i * depth + l) * 32:(i * depth + l + 1) * 32] for l in range(depth)]) for i in range(q)]

    def free(self, tree):
        self.ctx.lib.sb_tree_free(self.ctx.h, tree[0])

    def fri_fold(self, vals, root_limbs, values_root):
        """fri.rs:135-164: (n, 4) values -> (n/4, 4) column at special_x = int_LE(values_root) mod p"""
        n = vals.shape[0]
        col = self.empty(n // 4, 4)
        root = np.ascontiguousarray(root_limbs, dtype=np.uint64)
        vr = np.frombuffer(bytes(values_root), dtype=np.uint8).copy()
        self.ctx.check(self.ctx.lib.sb_fri_fold_dev(self.ctx.h, C.c_void_p(vals.data_ptr()), n, C.c_void_p(root.ctypes.data),
                                                    C.c_void_p(vr.ctypes.data), C.c_void_p(col.data_ptr())))
        return col

    def fri_rest(self, col, root_limbs, max_deg_plus_1, excl, tree):
        """prove_low_degree on the folded column with its already committed tree -> list of layers (fri.py layout)"""
        from . import fri
        root = np.ascontiguousarray(root_limbs, dtype=np.uint64)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.sb_fri_prove_dev(self.ctx.h, C.c_void_p(col.data_ptr()), col.shape[0], C.c_void_p(root.ctypes.data),
                                                     max_deg_plus_1, excl, tree[0], C.byref(h)))
        try:
            return fri.unpack_proof(self.ctx, h)
        finally:
            self.ctx.lib.sb_fri_proof_free(h)

    def ntt_batch(self, x, root_int, log_n, inverse=False):
        """(polys, 2^log_n, 4) contiguous -> same shape: best_fft / inv_best_fft of every row (fft.rs:327-379)"""
        from . import field
        polys, n = x.shape[0], 1 << log_n
        out = self.empty(polys, n, 4)
        root = field.mont_scalar(root_int)
        self.ctx.check(self.ctx.lib.sb_ntt_dev(self.ctx.h, C.c_void_p(x.data_ptr()), n, n, C.c_void_p(out.data_ptr()), n, polys,
                                               C.c_void_p(root.ctypes.data), log_n, 1 if inverse else 0))
        return out

    def twiddle_mul(self, x, row0, root_int, log_n, inverse=False):
        """x[r][c] *= root^((row0 + r) c) in place (inverse: root^-1)"""
        from . import field
        root = field.mont_scalar(root_int)
        self.ctx.check(self.ctx.lib.sb_twiddle_mul_dev(self.ctx.h, C.c_void_p(x.data_ptr()), x.shape[0], x.shape[1], row0,
                                                       C.c_void_p(root.ctypes.data), log_n, 1 if inverse else 0))

    def root_tensor(self, root):
        return self.torch.frombuffer(bytearray(root), dtype=self.torch.uint8).to(self.device)

    def bytes_tensor(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint8)).to(self.device)

    def leaf_bytes(self, tree):
        return tree[2]


# ---- the sharded commitment ----------------------------------------------------------------------------
class ShardedTree:
    def __init__(self, backend, local_tree, n, world, rank, roots):
        self.backend, self.local_tree = backend, local_tree
        self.n, self.world, self.rank = n, world, rank
        self.root, self.levels = combine_roots(roots)

    def get_root(self):
        return self.root

    def width(self):
        return self.n

    def owner(self, index):
        return index // (self.n // self.world)

    def gen_proofs_local(self, indices):
        """openings of the indices this rank owns: {position in `indices`: (leaf, nodes)}; nodes are the full path,
        leaf level first, root excluded -- what Proof::validate (merkle_tree.rs:25-43) consumes"""
        lo, _ = row_range(self.n, self.world, self.rank)
        for i in indices:
            if not 0 <= i < self.n:
                raise ValueError("leaf index %d out of range (width %d)" % (i, self.n))
        mine = [(p, i) for p, i in enumerate(indices) if self.owner(i) == self.rank]
        got = self.backend.open(self.local_tree, [i - lo for _, i in mine])
        top = top_path(self.levels, self.rank)
        return {p: (leaf, nodes + top) for (p, _), (leaf, nodes) in zip(mine, got)}

    def gen_proofs(self, indices, dist=None, group=None):
        """every rank calls this with the same indices (they come out of the Fiat-Shamir sampler, which every rank
        runs on the same root) and gets all openings, in caller order, duplicates allowed.  Each opening is owned by
        exactly one rank, so the exchange is one all_reduce(SUM) of a byte matrix that is zero outside the caller's rows
        (a few hundred KB; no pickling on the path)."""
        local = self.gen_proofs_local(indices)
        if self.world == 1 or dist is None:
            return [local[p] for p in range(len(indices))]
        lb = self.backend.leaf_bytes(self.local_tree)
        depth = self.n.bit_length() - 1
        row = lb + 32 * depth
        buf = np.zeros((len(indices), row), dtype=np.uint8)
        for p, (leaf, nodes) in local.items():
            buf[p, :lb] = np.frombuffer(leaf, dtype=np.uint8)
            buf[p, lb:] = np.frombuffer(b"".join(nodes), dtype=np.uint8)
        t = self.backend.bytes_tensor(buf)
        dist.all_reduce(t, group=group)
        raw = t.cpu().numpy().tobytes()
        out = []
        for p in range(len(indices)):
            r = raw[p * row:(p + 1) * row]
            out.append((r[:lb], [r[lb + 32 * l:lb + 32 * (l + 1)] for l in range(depth)]))
        return out

    def free(self):
        if self.local_tree is not None:
            self.backend.free(self.local_tree)
            self.local_tree = None


class ShardedCommitter:
    """exchange + subtree commit + root gather for one set of columns of n rows"""

    def __init__(self, backend, dist=None, group=None):
        self.backend, self.dist, self.group = backend, dist, group
        self.world = dist.get_world_size(group) if dist is not None else 1
        self.rank = dist.get_rank(group) if dist is not None else 0

    def exchange(self, ext_local, n_cols, n):
        """ext_local: {column id: (n, 4) tensor} for owned_columns(n_cols, world, rank)  ->  (n_cols, n/world, 4)
        tensor holding this rank's row range of EVERY column.  One grouped send/recv per (column, peer) pair;
        both sides walk the columns in ascending order so the pairs match."""
        lo, hi = row_range(n, self.world, self.rank)
        rows = self.backend.empty(n_cols, hi - lo, 4)
        for c, t in ext_local.items():
            rows[c].copy_(t[lo:hi])
        if self.world == 1:
            return rows
        dist = self.dist
        ops = []
        for c in range(n_cols):
            owner = c % self.world
            if owner == self.rank:
                for j in range(self.world):
                    if j != self.rank:
                        jl, jh = row_range(n, self.world, j)
                        ops.append(dist.P2POp(dist.isend, ext_local[c][jl:jh], self._global(j), group=self.group))
            else:
                ops.append(dist.P2POp(dist.irecv, rows[c], self._global(owner), group=self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return rows

    def _global(self, r):
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def commit_rows(self, rows, col_ids, n):
        """subtree over this rank's rows of the given columns + all_gather of the subtree roots"""
        root, tree = self.backend.commit_cols([rows[c] for c in col_ids])
        if self.world == 1:
            roots = [root]
        else:
            mine = self.backend.root_tensor(root)
            allr = self.backend.torch.empty(self.world * 32, dtype=mine.dtype, device=mine.device)
            self.dist.all_gather_into_tensor(allr, mine, group=self.group)
            b = allr.cpu().numpy().tobytes()
            roots = [b[32 * i:32 * (i + 1)] for i in range(self.world)]
        return ShardedTree(self.backend, tree, n, self.world, self.rank, roots)


def prove_low_degree_sharded(backend, values_tree, vals, owner, root_int, n, max_deg_plus_1, excl, dist=None, group=None,
                             replicate=False):
    """fri.rs:46-224 when the values tree of layer 0 is a ShardedTree (the prover's l_tree, prove.rs:324-332, built
    over row shards) and the values themselves sit whole on rank `owner` (the rank that extended / combined that
    column).  Layer 0: the owner folds and commits the column, broadcasts the 32-byte root2, every rank derives the
    same query positions (fri.rs:181-204) and contributes the openings of the rows it owns; layers >= 1 hold <= n/4
    values and run on the owner alone.  Returns the proof (fri.py layout) on the owner and None
    elsewhere; replicate=True broadcasts it (pickled) to every rank."""
    from . import field, merkle as mk
    world, rank = values_tree.world, values_tree.rank
    if max_deg_plus_1 <= 16:
        raise ValueError("a direct (Last-only) proof has no layer to shard")      # fri.rs:88 MIN_DEG_DIRECT_CHECKING
    q = n // 4
    w_limbs = field.mont_scalar(root_int)
    w4_limbs = field.mont_scalar(pow(root_int, 4, field.P))
    rest, col_tree, column_branches = None, None, None
    if rank == owner:
        col = backend.fri_fold(vals, w_limbs, values_tree.get_root())
        root2, col_tree = backend.commit_cols([col])
    else:
        root2 = bytes(32)
    if world > 1:
        t = backend.root_tensor(root2)
        dist.broadcast(t, owner if group is None else dist.get_global_rank(group, owner), group=group)
        root2 = t.cpu().numpy().tobytes()
    ys = utils.get_pseudorandom_indices(root2, q, 40, excl, ctx=getattr(backend, "ctx", None))                      # fri.rs:181-190
    positions = [y + q * j for y in ys for j in range(4)]                        # fri.rs:193-204
    poly = values_tree.gen_proofs(positions, dist if world > 1 else None, group)
    layer0 = None
    if rank == owner:
        cb = backend.open(col_tree, ys)
        layer0 = {"Middle": {"root2": root2,
                             "column_branches": [mk.Proof(leaf, nodes) for leaf, nodes in cb],
                             "poly_branches": [mk.Proof(leaf, nodes) for leaf, nodes in poly]}}
        rest = backend.fri_rest(col, w4_limbs, max_deg_plus_1 // 4, excl, col_tree)
        backend.free(col_tree)
        proof = [layer0] + rest
    else:
        proof = None
    if replicate and world > 1 and dist is not None:
        box = [proof]
        dist.broadcast_object_list(box, owner if group is None else dist.get_global_rank(group, owner), group=group)
        proof = box[0]
    return proof


# ---- one transform over several GPUs ---------------------------------------------------
def distributed_ntt(backend, x_local, root_int, log_n, inverse=False, dist=None, group=None):
    """best_fft / inv_best_fft (fft.rs:327-379) of ONE 2^log_n-point vector spread over the ranks in natural order:
    rank r holds elements [r n/g, (r+1) n/g) on entry and the same range of the result on return.  Four-step
    decomposition n = n1 n2 (j = j1 n2 + j2, k = k1 + n1 k2):
        all_to_all (row slabs -> column slabs), n1-point transforms along j1, twiddle w^(j2 k1),
        all_to_all (-> k1 slabs), n2-point transforms along j2, all_to_all (-> natural order).
    The local transforms are the batched single-GPU kernels; the transposes in between are torch copies; the three
    exchanges are NCCL all_to_all_single calls of n/g elements each.  The inverse works the same way with w^-1 (the two
    local inverses contribute 1/n1 and 1/n2)."""
    from . import field
    g = dist.get_world_size(group) if dist is not None else 1
    rank = dist.get_rank(group) if dist is not None else 0
    n = 1 << log_n
    if g == 1:
        return backend.ntt_batch(x_local.reshape(1, n, 4), root_int, log_n, inverse).reshape(n, 4)
    log_n1 = (log_n + 1) // 2
    log_n2 = log_n - log_n1
    n1, n2 = 1 << log_n1, 1 << log_n2
    if g & (g - 1) or n1 % g or n2 % g:
        raise ValueError("world size %d must be a power of two dividing both factors of 2^%d" % (g, log_n))
    if x_local.shape[0] != n // g:
        raise ValueError("rank slab must hold n / world = %d elements" % (n // g))
    a, b = n1 // g, n2 // g

    def exchange(t):
        out = backend.empty(*t.shape)
        dist.all_to_all_single(out, t, group=group)
        return out

    # rows j1 in [rank a, (rank+1) a) -> columns j2 in [rank b, (rank+1) b), all j1
    m1 = exchange(x_local.reshape(a, g, b, 4).permute(1, 0, 2, 3).contiguous())          # [rho][j1'][j2'] = [j1][j2']
    col = m1.reshape(n1, b, 4).transpose(0, 1).contiguous()                              # [j2'][j1]
    col = backend.ntt_batch(col, pow(root_int, n2, field.P), log_n1, inverse)            # [j2'][k1]
    backend.twiddle_mul(col, rank * b, root_int, log_n, inverse)                         # * w^(j2 k1)
    # columns -> k1 slabs, all j2
    m2 = exchange(col.reshape(b, g, a, 4).permute(1, 0, 2, 3).contiguous())              # [sigma][j2'][k1'] = [j2][k1']
    row = m2.reshape(n2, a, 4).transpose(0, 1).contiguous()                              # [k1'][j2]
    row = backend.ntt_batch(row, pow(root_int, n1, field.P), log_n2, inverse)            # [k1'][k2] = X[k1 + n1 k2]
    # k1 slabs -> natural order: rank u owns k2 in [u b, (u+1) b)
    m3 = exchange(row.reshape(a, g, b, 4).permute(1, 0, 2, 3).contiguous())              # [tau][k1'][k2'] = [k1][k2']
    return m3.reshape(n1, b, 4).transpose(0, 1).contiguous().reshape(n // g, 4)          # [k2'][k1]
