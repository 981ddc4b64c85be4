import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, random
import stark_pure_rust_b200 as sb
from stark_pure_rust_b200 import field
from stark_pure_rust_b200._lib import _ptr
P = field.P; R = 1 << 256
ctx = sb.default_context()
def raw(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype="<u8").reshape(-1, 4).copy()
def unraw(a):
    b = a.tobytes(); return [int.from_bytes(b[i:i+32], "little") for i in range(0, len(b), 32)]
def op(o, A, B):
    a, b = raw(A), raw(B); out = np.zeros_like(a)
    ctx.check(ctx.lib.sb_fp_vec_op(ctx.h, o, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
    return unraw(out)
rnd = random.Random(1)
n = 64
A = [rnd.randrange(P) for _ in range(n)]; B = [rnd.randrange(P) for _ in range(n)]
A[0], B[0] = R % P, 5; A[1], B[1] = 7, 1; A[2], B[2] = 1, 1; A[3], B[3] = 1 << 32, 1; A[4], B[4] = 1, 1 << 32; A[5], B[5] = (1<<64), (1<<64)
A[6], B[6] = 1 << 32, 1 << 32
Rinv = pow(R, -1, P)
got = op(0, A, B)
bad = [i for i in range(n) if got[i] % P != A[i] * B[i] * Rinv % P]
print("mul bad", len(bad), bad[:10])
for i in bad[:6]:
    print(" a", hex(A[i]), "b", hex(B[i]), "\n  got", hex(got[i]), "\n  want", hex(A[i] * B[i] * Rinv % P))
for name, o, f in (("add", 1, lambda a, b: (a + b) % P), ("sub", 2, lambda a, b: (a - b) % P), ("sub_lazy", 3, lambda a, b: (a - b) % P),
                   ("canon", 4, lambda a, b: a % P), ("half", 5, lambda a, b: a * pow(2, -1, P) % P)):
    g = op(o, A, B)
    bad = [i for i in range(n) if g[i] % P != f(A[i], B[i])]
    print(name, "bad", len(bad), bad[:5])
g = op(11, A, B)
bad = [i for i in range(n) if g[i] % P != A[i] * A[i] * Rinv % P]
print("sqr bad", len(bad), bad[:5])
B2 = [i % 7 for i in range(n)]
g = op(12, A, B2)
def rep(a, k):
    for _ in range(k): a = a * a * Rinv % P
    return a
bad = [i for i in range(n) if g[i] != rep(A[i], B2[i])]
print("repeated sqr bad", len(bad), bad[:5])
w = field.root_of_unity(2)
got = field.from_mont(sb.fft.expand_root_of_unity(w, order=4, ctx=ctx))
print([hex(x) for x in got]); print([hex(pow(w, i, P)) for i in range(4)])
