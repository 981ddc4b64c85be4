import sys, os, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import stark_pure_rust_b200 as sb
from conftest import random_elems
g = int(sys.argv[1]); log_s = int(sys.argv[2])
ctx = sb.Context(devices=[0] * g if os.environ.get("SB_LOGICAL") else list(range(g)))
ctx.check(ctx.lib.sb_set_extended_domain(ctx.h, 1))
nc = 8
S = 1 << log_s
base = random_elems(1 << 18, 3)
cols = np.stack([np.roll(np.tile(base, (S >> 18, 1)), 7 * c, axis=0) for c in range(nc)])
e = sb.ext.ExtColumns(nc, log_s, ctx=ctx)
e.load(0, cols)
for rep in range(3):
    t0 = time.perf_counter(); e.extend(); t1 = time.perf_counter()
    print("rep", rep, "extend %.2f ms" % ((t1 - t0) * 1e3), file=sys.stderr)
    r, t = e.commit(list(range(nc))); t2 = time.perf_counter()
    print("commit8 %.2f ms" % ((t2 - t1) * 1e3), file=sys.stderr)
    r1, tl = e.commit([nc - 1]); t3 = time.perf_counter()
    print("commit1 %.2f ms" % ((t3 - t2) * 1e3), file=sys.stderr)
    e.fri_prove(nc - 1, (8 * S) // 4, 8, tree=tl, as_json=True); t4 = time.perf_counter()
    print("fri %.2f ms" % ((t4 - t3) * 1e3), file=sys.stderr)
    e.free_tree(t); e.free_tree(tl)
