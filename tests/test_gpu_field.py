"""the device field library (csrc/fp.cuh, fp_gen.cuh) driven directly through the unit-test hook
sb_fp_vec_op, every op against Python big-int ground truth, bit-exact.

The reference's arithmetic is ff_derive-generated code (ff_utils/src/fp.rs:8-12); results in F_p are unique, so the
ground truth is plain integer arithmetic mod p.  The device library works with lazy ranges ([0,2p) between butterflies,
[0,4p) feeding a product's first operand), so the operands here are adversarial for exactly those contracts: 0, 1, p-1,
p, p+1, 2p-1, 2p, 4p-1, 2^256-p-1, limbs of 0xffffffff / 0, and values straddling every conditional-subtract boundary.
Lazy results are compared with the exact integer the algorithm defines (not only modulo p), so a wrong carry shows up
even when the residue happens to survive.
"""
import random

import numpy as np
import pytest

from conftest import P

pytestmark = pytest.mark.gpu

R = 1 << 256
M32 = (1 << 32) - 1
NINV = (-pow(P, -1, R)) % R          # -p^-1 mod 2^256
RINV = pow(R, -1, P)


def raw(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype="<u8").reshape(-1, 4).copy()


def unraw(a):
    b = a.tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def run(ctx, op, A, B):
    from stark_pure_rust_b200._lib import _ptr
    a, b = raw(A), raw(B)
    out = np.zeros_like(a)
    ctx.check(ctx.lib.sb_fp_vec_op(ctx.h, op, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
    return unraw(out)


def redc(a, b):
    """the integer a Montgomery product without final subtraction yields: (ab + m p) / 2^256, m = -ab p^-1 mod 2^256"""
    t = a * b
    m = (t * NINV) % R
    assert (t + m * P) % R == 0
    return (t + m * P) >> 256


def limb_patterns():
    out = []
    for mask in range(256):            # every combination of all-ones / all-zero 32-bit limbs
        out.append(sum((M32 if (mask >> j) & 1 else 0) << (32 * j) for j in range(8)))
    for j in range(8):                 # a single 1 / a single 0xffffffff / everything but one limb
        out += [1 << (32 * j), M32 << (32 * j), (R - 1) ^ (M32 << (32 * j)), (1 << (32 * j)) - 1, (1 << (32 * j + 31))]
    return out


def edges():
    e = [0, 1, 2, P - 2, P - 1, P, P + 1, 2 * P - 2, 2 * P - 1, 2 * P, 2 * P + 1, 3 * P - 1, 3 * P, 3 * P + 1, 4 * P - 2, 4 * P - 1,
         R - P - 2, R - P - 1, (P - 1) // 2, (P + 1) // 2, R % P, (R * R) % P, 1 << 255, (1 << 254) - 1, 1 << 254]
    return e


def in_range(vals, bound):
    return [v for v in vals if 0 <= v < bound]


def pairs(As, Bs, rnd, extra):
    cases = [(a, b) for a in As for b in Bs]
    cases += extra
    rnd.shuffle(cases)
    return [c[0] for c in cases], [c[1] for c in cases]


def test_fp_mul_lazy_exact(ctx):
    """op 0 / 11: fp_mul(a, b) for a < 2^256 - p (the contract; covers [0,4p)), b < 2p: exactly (ab + mp) / 2^256"""
    rnd = random.Random(0xA1)
    As = in_range(edges() + limb_patterns(), R - P)
    Bs = in_range(edges() + limb_patterns(), 2 * P)
    extra = [(rnd.randrange(4 * P), rnd.randrange(2 * P)) for _ in range(20000)]
    extra += [(rnd.randrange(R - P), rnd.randrange(2 * P)) for _ in range(20000)]
    A, B = pairs(As, Bs, rnd, extra)
    got = run(ctx, 0, A, B)
    for a, b, g in zip(A, B, got):
        want = redc(a, b)
        assert g == want, "fp_mul(%x, %x) = %x, want %x" % (a, b, g, want)
        if a * b < P * R:
            assert g < 2 * P
    # squares of values in [0, 2p)
    A2 = in_range(edges() + limb_patterns(), 2 * P) + [rnd.randrange(2 * P) for _ in range(5000)]
    got = run(ctx, 11, A2, A2)
    for a, g in zip(A2, got):
        assert g == redc(a, a)
    # canonical product (op 10) on canonical operands equals a b R^-1 mod p
    A3 = [rnd.randrange(P) for _ in range(5000)] + [0, 1, P - 1]
    B3 = [rnd.randrange(P) for _ in range(5000)] + [P - 1, P - 1, P - 1]
    got = run(ctx, 10, A3, B3)
    for a, b, g in zip(A3, B3, got):
        assert g == a * b * RINV % P


def test_fp_add_sub_ranges(ctx):
    """ops 1, 2, 3, 9: add / sub keep [0,2p); sub_lazy is a + 2p - b exactly; reduce_2p folds [0,4p) into [0,2p)"""
    rnd = random.Random(0xA2)
    E = in_range(edges() + limb_patterns(), 2 * P)
    extra = [(rnd.randrange(2 * P), rnd.randrange(2 * P)) for _ in range(20000)]
    # straddle the conditional-subtract boundary a + b = 2p and the borrow boundary a = b
    for _ in range(2000):
        a = rnd.randrange(2 * P)
        for d in (-2, -1, 0, 1, 2):
            b = 2 * P - a + d
            if 0 <= b < 2 * P:
                extra.append((a, b))
            b = a + d
            if 0 <= b < 2 * P:
                extra.append((a, b))
    A, B = pairs(E, E, rnd, extra)
    add, sub, lazy = run(ctx, 1, A, B), run(ctx, 2, A, B), run(ctx, 3, A, B)
    for a, b, g1, g2, g3 in zip(A, B, add, sub, lazy):
        s = a + b
        assert g1 == (s - 2 * P if s >= 2 * P else s), "fp_add(%x, %x) = %x" % (a, b, g1)
        assert g2 == (a - b if a >= b else a - b + 2 * P), "fp_sub(%x, %x) = %x" % (a, b, g2)
        assert g3 == a + 2 * P - b, "fp_sub_lazy(%x, %x) = %x" % (a, b, g3)
    X = in_range(edges() + limb_patterns(), 4 * P) + [rnd.randrange(4 * P) for _ in range(20000)]
    X += [2 * P + d for d in range(-3, 4)]
    got = run(ctx, 9, X, X)
    for x, g in zip(X, got):
        assert g == (x - 2 * P if x >= 2 * P else x), "fp_reduce_2p(%x) = %x" % (x, g)


def test_fp_canon_half_codecs(ctx):
    """ops 4, 5, 6, 7: canonicalise [0,2p) -> [0,p); exact halving; from / to Montgomery (fp.rs:39-43, :74-76 codecs)"""
    rnd = random.Random(0xA3)
    X = in_range(edges() + limb_patterns(), 2 * P) + [rnd.randrange(2 * P) for _ in range(20000)] + [P + d for d in range(-3, 4)]
    canon, half = run(ctx, 4, X, X), run(ctx, 5, X, X)
    for x, g, h in zip(X, canon, half):
        assert g == (x - P if x >= P else x), "fp_canon(%x) = %x" % (x, g)
        assert h == ((x + P) >> 1 if x & 1 else x >> 1), "fp_half(%x) = %x" % (x, h)
        assert h < (3 * P + 1) // 2 + 1 and (2 * h - x) % P == 0
    # from_mont (dedicated reduction): any a < 2^256 -> a R^-1 mod p, canonical
    Y = in_range(edges() + limb_patterns(), R) + [rnd.randrange(R) for _ in range(20000)] + [R - 1 - d for d in range(8)]
    got = run(ctx, 6, Y, Y)
    for y, g in zip(Y, got):
        assert g == y * RINV % P, "fp_from_mont(%x) = %x" % (y, g)
    # to_mont: any a < 2^256 - p -> a R mod p, canonical
    Y = in_range(edges() + limb_patterns(), R - P) + [rnd.randrange(R - P) for _ in range(20000)]
    got = run(ctx, 7, Y, Y)
    for y, g in zip(Y, got):
        assert g == y * R % P, "fp_to_mont(%x) = %x" % (y, g)


def test_fp_inverse_and_chains(ctx):
    """op 8: Fermat inverse (the batch inverse's per-slice inversion); op 12: chains of canonical squarings"""
    rnd = random.Random(0xA4)
    X = [1, 2, P - 1, P - 2, R % P, (R * R) % P, (P - 1) // 2] + [rnd.randrange(1, P) for _ in range(500)]
    got = run(ctx, 8, X, X)
    for x, g in zip(X, got):
        # Montgomery inverse: g = x^-1 in the Montgomery domain, i.e. g * x * R^-1 = R (mod p)
        assert g < P and g * x % P == R * R % P, "fp_inv(%x) = %x" % (x, g)
    assert run(ctx, 8, [0], [0]) == [0]
    K = [i % 9 for i in range(len(X))]
    got = run(ctx, 12, X, K)
    for x, k, g in zip(X, K, got):
        v = x
        for _ in range(k):
            v = v * v * RINV % P
        assert g == v
