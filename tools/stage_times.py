#!/usr/bin/env python3
"""Per-stage wall-clock vs CUDA-event vs per-kernel-family time of the bench step (where do the gaps come from?)."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import stark_pure_rust_b200 as sb
from stark_pure_rust_b200 import field
from stark_pure_rust_b200._lib import _ptr
from conftest import random_elems

KINDS = ["ntt_pass", "merkle_leaves", "merkle_nodes", "fri_fold", "open", "other"]


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    Cn = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    ctx = sb.Context(0)
    lib = ctx.lib
    log_s, N, S = L - 3, 1 << L, 1 << (L - 3)
    g2 = field.mont_scalar(field.root_of_unity(L))
    cols = random_elems(Cn * S, 0xB200).reshape(Cn, S, 4)
    d_cols = ctx.to_device(cols)
    d_out = ctx.alloc(Cn * N * 32)
    k_tree = min(8, Cn)
    col_ptrs8 = (C.c_void_p * k_tree)(*[d_out + i * N * 32 for i in range(k_tree)])
    col_ptr1 = (C.c_void_p * 1)(d_out + (Cn - 1) * N * 32)
    root = np.empty(32, dtype=np.uint8)
    state = {}

    def lde():
        ctx.check(lib.sb_lde_batch_dev(ctx.h, C.c_void_p(d_cols), Cn, S, S, _ptr(g2), log_s, 3, C.c_void_p(d_out)))

    def commit8():
        t = C.c_void_p()
        ctx.check(lib.sb_merkle_commit_cols_dev(ctx.h, col_ptrs8, k_tree, N, _ptr(root), C.byref(t)))
        lib.sb_tree_free(ctx.h, t)

    def commit1():
        t = C.c_void_p()
        ctx.check(lib.sb_merkle_commit_cols_dev(ctx.h, col_ptr1, 1, N, _ptr(root), C.byref(t)))
        state["tl"] = t

    def fri():
        pr = C.c_void_p()
        ctx.check(lib.sb_fri_prove_dev(ctx.h, C.c_void_p(col_ptr1[0]), N, _ptr(g2), N // 4, 8, state["tl"], C.byref(pr)))
        lib.sb_fri_proof_free(pr)
        lib.sb_tree_free(ctx.h, state["tl"])

    stages = [("lde", lde), ("commit8", commit8), ("commit1", commit1), ("fri", fri)]
    for _ in range(2):
        for _, f in stages:
            f()
    ctx.sync()
    for rep in range(2):
        for name, f in stages:
            ctx.sync()
            ctx.profile(True)
            t0 = time.perf_counter()
            ctx.timer_start()
            f()
            ev = ctx.timer_stop()
            wall = (time.perf_counter() - t0) * 1e3
            prof = {k: ctx.profile_read(i) for i, k in enumerate(KINDS)}
            ctx.profile(False)
            ksum = sum(v[1] for v in prof.values())
            print("L=%d %-8s wall %8.3f ms  events %8.3f ms  kernels %8.3f ms  %s" % (
                L, name, wall, ev, ksum, {k: (v[0], round(v[1], 3)) for k, v in prof.items() if v[0]}))
    # same without per-launch events
    for name, f in stages:
        ctx.sync()
        t0 = time.perf_counter()
        ctx.timer_start()
        f()
        ev = ctx.timer_stop()
        wall = (time.perf_counter() - t0) * 1e3
        print("L=%d %-8s (no profile) wall %8.3f ms  events %8.3f ms" % (L, name, wall, ev))


if __name__ == "__main__":
    main()
