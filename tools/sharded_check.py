#!/usr/bin/env python3
"""Parity check of the sharded commitment on real GPUs: run under torchrun (or plain python for world = 1).
Every rank extends its columns, exchanges to row shards, commits its subtree; rank 0 additionally commits ALL columns
on its own GPU in one tree and compares root and openings.  Prints "SHARDED_OK ..." on success.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py 16 5
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import stark_pure_rust_b200 as sb
    from stark_pure_rust_b200 import field, sharded, merkle
    from conftest import random_elems
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    Cn = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = sb.Context(local_rank)
    be = sharded.CudaBackend(ctx, dev)
    log_s, N, S = L - 3, 1 << L, 1 << (L - 3)
    g2 = field.mont_scalar(field.root_of_unity(L))
    cols = random_elems(Cn * S, 777).reshape(Cn, S, 4)            # same on every rank
    mine = sharded.owned_columns(Cn, world, rank)
    ext = be.lde(be.from_numpy(cols[mine]) if mine else be.empty(0, S, 4), g2, log_s, 3)
    sc = sharded.ShardedCommitter(be, dist if world > 1 else None)
    rows = sc.exchange({c: ext[k] for k, c in enumerate(mine)}, Cn, N)
    k_tree = min(8, Cn)
    tree = sc.commit_rows(rows, list(range(k_tree)), N)
    rng = np.random.default_rng(5)
    idx = [0, N - 1, (N // world) % N, N // world - 1, 3, 3] + [int(x) for x in rng.integers(0, N, size=40)]
    proofs = tree.gen_proofs(idx, dist if world > 1 else None)
    ok = True
    if rank == 0:
        full = be.lde(be.from_numpy(cols), g2, log_s, 3)
        root1, t1 = be.commit_cols([full[c] for c in range(k_tree)])
        want = be.open(t1, idx)
        ok = root1 == tree.get_root() and all(a == b for a, b in zip(proofs, want))
        ok = ok and all(merkle.Proof(leaf, nodes).validate(root1, i) for (leaf, nodes), i in zip(proofs, idx))
        be.free(t1)
        print(("SHARDED_OK" if ok else "SHARDED_MISMATCH"), "world", world, "L", L, "cols", Cn, "root", tree.get_root().hex(), flush=True)
    # FRI with the row-sharded values tree of the last column (the prover's l_tree) against the single-GPU prover
    lc = Cn - 1
    owner = lc % world
    ltree = sc.commit_rows(rows, [lc], N)
    vals = ext[mine.index(lc)] if rank == owner else None
    g2i = field.root_of_unity(L)
    proof = sharded.prove_low_degree_sharded(be, ltree, vals, owner, g2i, N, N // 4, 8, dist if world > 1 else None, replicate=True)
    if rank == 0:
        want = sb.fri.prove_low_degree(full[lc].cpu().numpy().view(np.uint64), g2i, N // 4, 8, ctx=ctx)
        same = len(proof) == len(want)
        for a, b in zip(proof, want):
            same = same and a == b
        print(("SHARDED_FRI_OK" if same else "SHARDED_FRI_MISMATCH"), "layers", len(proof), flush=True)
        ok = ok and same
    # one transform over all ranks (four-step, three all_to_all exchanges) against the single-GPU transform
    x = random_elems(N, 31337)
    lo, hi = sharded.row_range(N, world, rank)
    ntt_ok = True
    for inverse in (False, True):
        got = sharded.distributed_ntt(be, be.from_numpy(x[lo:hi]), g2i, L, inverse, dist if world > 1 else None)
        want = sb.fft.best_fft(x, g2i, L, ctx=ctx, inverse=inverse)
        ntt_ok = ntt_ok and np.array_equal(got.cpu().numpy().view(np.uint64), want[lo:hi])
    if world > 1:
        flag = torch.tensor([1 if ntt_ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ntt_ok = bool(flag.item())
    if rank == 0:
        print(("SHARDED_NTT_OK" if ntt_ok else "SHARDED_NTT_MISMATCH"), flush=True)
    ok = ok and ntt_ok
    ltree.free()
    tree.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
