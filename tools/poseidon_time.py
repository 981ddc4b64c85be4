"""Times the Poseidon-digest tree (sb_merkle_commit_poseidon, host leaves in, root out) on cuda:0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, ctypes as C
import stark_pure_rust_b200 as sb
from stark_pure_rust_b200._lib import _ptr
ctx = sb.Context(0)
for logn in (16, 20):
    n = 1 << logn
    a = np.random.default_rng(1).integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 31] &= 0x3f
    root = np.empty(32, dtype=np.uint8)
    for rep in range(3):
        t = C.c_void_p()
        t0 = time.perf_counter()
        ctx.check(ctx.lib.sb_merkle_commit_poseidon(ctx.h, _ptr(a), 32, n, _ptr(root), C.byref(t)))
        dt = time.perf_counter() - t0
        ctx.lib.sb_tree_free(ctx.h, t)
    print("poseidon tree 2^%d leaves: %.2f ms (%.3g hashes/s)" % (logn, dt * 1e3, (2 * n - 1) / dt))
