#!/usr/bin/env python3
"""Generator + emulator for the BN254-Fr PTX carry-chain blocks (csrc/fp_gen.cuh).

Every multi-limb primitive of the device field library is described ONCE as a list of PTX
instructions on symbolic registers.  This script (1) executes that list in Python with exact PTX
carry-flag semantics against big-int ground truth (random + edge operands), and only then
(2) prints the very same list as CUDA inline-asm statements.  The carry flag never has to survive
between two asm statements: each chain lives inside one statement, and the emulator poisons CC
at statement boundaries to prove it.

Run:  python tools/gen_fp.py            (self-test, then rewrite csrc/fp_gen.cuh)
      python tools/gen_fp.py --check    (self-test + verify the committed header is current)

Element representation: 8 x u32 little-endian limbs, Montgomery form with R = 2^256, which is
bit-identical to the reference's Fp([u64;4]) (ff_utils/src/fp.rs:8-12, ff_derive 0.10.0).
Lazy-reduction contract (p < 2^254, so 4p < 2^256):
    mul(a, b)   needs a < 2^256 - p (~4.29p) and a*b < p*R  ->  result < 2p
    values travelling between butterflies live in [0, 2p); canonical (< p) only at the ABI edge.
"""
import os
import random
import sys

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = 1 << 256
M32 = 0xFFFFFFFF
NINV32 = (-pow(P, -1, 1 << 32)) % (1 << 32)
assert NINV32 == 0xEFFFFFFF


def limbs(x, n=8):
    return [(x >> (32 * i)) & M32 for i in range(n)]


PL = limbs(P)
P2L = limbs(2 * P)

# --------------------------------------------------------------------------------------------
# tiny PTX model
# --------------------------------------------------------------------------------------------
# operand: ("r", name) register | ("i", value) immediate
def reg(name, idx=None):
    return ("r", name if idx is None else f"{name}{idx}")


def imm(v):
    return ("i", v & M32)


class Stmt:
    """one asm statement = list of (op, dst, srcs)"""

    def __init__(self):
        self.ins = []
        self.temps = []

    def add(self, op, dst, *srcs):
        self.ins.append((op, dst, srcs))

    def temp(self, name):
        self.temps.append(name)
        return ("r", name)


def emulate(stmts, env):
    """env: dict name -> u32.  Executes statements; CC is poisoned between statements."""
    for st in stmts:
        cc = None
        local = {}

        def rd(o):
            if o[0] == "i":
                return o[1]
            if o[1] in local:
                return local[o[1]]
            return env[o[1]]

        def wr(o, v):
            assert o[0] == "r"
            if o[1] in st.temps:
                local[o[1]] = v & M32
            else:
                env[o[1]] = v & M32

        for op, dst, srcs in st.ins:
            v = [rd(s) for s in srcs]
            base = op.split(".")
            kind = base[0]
            uses_c = kind.endswith("c") and kind not in ("sub",)  # addc/subc/madc
            sets_c = "cc" in base
            if uses_c:
                assert cc is not None, f"{op}: carry-in used but CC undefined in this statement"
            cin = cc if uses_c else 0
            if kind in ("mul",):
                prod = v[0] * v[1]
                res = prod & M32 if "lo" in base else prod >> 32
                assert not sets_c
                wr(dst, res)
            elif kind in ("mad", "madc"):
                prod = v[0] * v[1]
                part = prod & M32 if "lo" in base else prod >> 32
                tot = part + v[2] + cin
                wr(dst, tot)
                if sets_c:
                    cc = tot >> 32
                else:
                    # a dropped carry must be provably zero in our usage
                    assert tot >> 32 == 0, f"{op}: carry dropped but non-zero"
            elif kind in ("shl", "shr"):
                assert srcs[1][0] == "i" and 0 < v[1] < 32 and not sets_c
                wr(dst, (v[0] << v[1]) if kind == "shl" else (v[0] >> v[1]))
            elif kind in ("add", "addc"):
                tot = v[0] + v[1] + cin
                wr(dst, tot)
                if sets_c:
                    cc = tot >> 32
            elif kind in ("sub", "subc"):
                tot = v[0] - v[1] - cin
                wr(dst, tot)
                if sets_c:
                    cc = 1 if tot < 0 else 0
            else:
                raise ValueError(op)
    return env


def emit(stmts, indent="    "):
    """CUDA inline asm text for the statements.  Registers named X<k> map to C expression X[k];
    plain names map to themselves.  Every register that is written is a "+r" operand, every
    other one an "r" operand."""
    out = []
    for st in stmts:
        written, read = [], []
        for op, dst, srcs in st.ins:
            for s in srcs:
                if s[0] == "r" and s[1] not in st.temps and s[1] not in read and s[1] not in written:
                    read.append(s[1])
            if dst[1] not in st.temps and dst[1] not in written:
                written.append(dst[1])
        read = [r for r in read if r not in written]
        order = written + read
        num = {name: i for i, name in enumerate(order)}

        def o(x):
            if x[0] == "i":
                return "0x%08x" % x[1]
            if x[1] in st.temps:
                return x[1]
            return "%%%d" % num[x[1]]

        lines = []
        if st.temps:
            lines.append(".reg .u32 " + ", ".join(st.temps) + ";")
        for op, dst, srcs in st.ins:
            lines.append(f"{op} {o(dst)}, " + ", ".join(o(s) for s in srcs) + ";")

        def cexpr(name):
            head = name.rstrip("0123456789")
            tail = name[len(head):]
            return f"{head}[{tail}]" if tail else head

        body = ("\\n\\t".join(lines))
        text = indent + 'asm("{\\n\\t' + body + '\\n\\t}"\n'
        text += indent + "    : " + ", ".join(f'"+r"({cexpr(w)})' for w in written) + "\n"
        text += indent + "    : " + ", ".join(f'"r"({cexpr(r_)})' for r_ in read) + ");\n"
        out.append(text)
    return "".join(out)


# --------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------
def chain(st, first_op, rest_op, last_op, dsts, a_list, b_list, c_list):
    """generic helper: emits first_op for element 0, rest_op for the middle, last_op for the end"""
    n = len(dsts)
    for i in range(n):
        op = first_op if i == 0 else (last_op if i == n - 1 else rest_op)
        st.add(op, dsts[i], a_list[i], b_list[i], *([c_list[i]] if c_list else []))


assert PL[0] == 0xF0000001 and NINV32 == (-(2**28 + 1)) % 2**32


def mont_m(st, m, x0):
    """m = x0 * (-p^-1) mod 2^32.  (-p^-1 = -(2^28 + 1), so shifts would do, but m heads the dependency chain of the whole
    row and one IMAD has a shorter latency than shift + add + negate: measured slower on B200.)"""
    st.add("mul.lo.u32", m, x0, imm(NINV32))


def mp_rows(st, X, Y, m):
    """T += m * p with T = X (even-aligned words 0..7) + Y (odd-aligned: Y[j] is word j+1).
    Odd chain first, then even chain, whose carry out of word 7 lands in word 8 = Y[7].
    (p0 = 2^32 - 2^28 + 1 would allow replacing the product m * p0 by shifts: word 0 becomes 0 by construction and word 1
    receives m - (m >> 4) + borrow(m << 28, m).  Tried on B200: the extra ALU instructions cost more than the IMAD.WIDE they
    save -- the LDE went from 34.2 to 34.8 ms -- so the plain product stays.)"""
    # odd limbs p1,p3,p5,p7 -> Y0..Y7
    for k, j in enumerate((1, 3, 5, 7)):
        lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
        hi = "madc.hi.cc.u32" if k < 3 else "madc.hi.u32"
        st.add(lo, reg(Y, 2 * k), m, imm(PL[j]), reg(Y, 2 * k))
        st.add(hi, reg(Y, 2 * k + 1), m, imm(PL[j]), reg(Y, 2 * k + 1))
    for k, j in enumerate((0, 2, 4, 6)):
        lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
        st.add(lo, reg(X, 2 * k), m, imm(PL[j]), reg(X, 2 * k))
        st.add("madc.hi.cc.u32", reg(X, 2 * k + 1), m, imm(PL[j]), reg(X, 2 * k + 1))
    st.add("addc.u32", reg(Y, 7), reg(Y, 7), imm(0))


def gen_mul():
    """Montgomery product, operand scanning over b, even/odd accumulator split so that every
    (mad.lo.cc, madc.hi.cc) pair is one 64-bit IMAD.WIDE with carry in SASS.
    Row 0's plain products are written in C++ (no carries); stmt list starts at its reduction."""
    stmts = []
    # ---- row 0 reduction (E = a_even * b0, O = a_odd * b0 already set by the caller) ----
    st = Stmt()
    m = st.temp("m")
    mont_m(st, m, reg("E", 0))
    mp_rows(st, "E", "O", m)
    stmts.append(st)
    # ---- rows 1..7 ----
    for i in range(1, 8):
        X, Y = ("O", "E") if i % 2 == 1 else ("E", "O")   # X becomes even-aligned, Y odd-aligned
        st = Stmt()
        m = st.temp("m")
        bi = reg("b", i)
        # shift right by one word: leftover word Y1 joins word 0; Y'[j] = Y[j+2] + a_odd*bi
        st.add("add.cc.u32", reg(X, 0), reg(X, 0), reg(Y, 1))
        for k, j in enumerate((1, 3, 5, 7)):
            c_lo = reg(Y, 2 * k + 2) if 2 * k + 2 < 8 else imm(0)
            c_hi = reg(Y, 2 * k + 3) if 2 * k + 3 < 8 else imm(0)
            st.add("madc.lo.cc.u32", reg(Y, 2 * k), reg("a", j), bi, c_lo)
            st.add("madc.hi.cc.u32" if k < 3 else "madc.hi.u32", reg(Y, 2 * k + 1), reg("a", j), bi, c_hi)
        for k, j in enumerate((0, 2, 4, 6)):
            lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
            st.add(lo, reg(X, 2 * k), reg("a", j), bi, reg(X, 2 * k))
            st.add("madc.hi.cc.u32", reg(X, 2 * k + 1), reg("a", j), bi, reg(X, 2 * k + 1))
        st.add("addc.u32", reg(Y, 7), reg(Y, 7), imm(0))
        mont_m(st, m, reg(X, 0))
        mp_rows(st, X, Y, m)
        stmts.append(st)
    # ---- final shift: r = (X >> 32) + Y with X = O, Y = E after row 7 ----
    st = Stmt()
    for j in range(8):
        op = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < 7 else "addc.u32")
        src = reg("O", j + 1) if j < 7 else imm(0)
        st.add(op, reg("E", j), reg("E", j), src)
    stmts.append(st)
    return stmts


def gen_redc():
    """Montgomery reduction alone: r = a / 2^256 mod p (<= p) for any a < 2^256 -- gen_mul's rows without the a*b products
    (E = a even-aligned, O = 0 on entry): 8 + 64 wide multiplies instead of mont_mul(a, 1)'s 8 + 128."""
    stmts = []
    st = Stmt()
    m = st.temp("m")
    mont_m(st, m, reg("E", 0))
    mp_rows(st, "E", "O", m)
    stmts.append(st)
    for i in range(1, 8):
        X, Y = ("O", "E") if i % 2 == 1 else ("E", "O")
        st = Stmt()
        m = st.temp("m")
        st.add("add.cc.u32", reg(X, 0), reg(X, 0), reg(Y, 1))       # leftover word joins word 0
        for j in range(8):                                          # Y' = Y >> 64 with the carry rippling through
            src = reg(Y, j + 2) if j + 2 < 8 else imm(0)
            st.add("addc.cc.u32" if j < 7 else "addc.u32", reg(Y, j), src, imm(0))
        mont_m(st, m, reg(X, 0))
        mp_rows(st, X, Y, m)
        stmts.append(st)
    st = Stmt()
    for j in range(8):
        op = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < 7 else "addc.u32")
        src = reg("O", j + 1) if j < 7 else imm(0)
        st.add(op, reg("E", j), reg("E", j), src)
    stmts.append(st)
    return stmts


def run_redc(a):
    env = {}
    for j in range(8):
        env[f"E{j}"] = limbs(a)[j]
        env[f"O{j}"] = 0
    emulate(REDC, env)
    return sum(env[f"E{j}"] << (32 * j) for j in range(8))


def run_mul(a, b):
    al, bl = limbs(a), limbs(b)
    env = {}
    for j in range(8):
        env[f"a{j}"] = al[j]
        env[f"b{j}"] = bl[j]
        env[f"r{j}"] = 0
    for k, j in enumerate((0, 2, 4, 6)):
        pr = al[j] * bl[0]
        env[f"E{2*k}"], env[f"E{2*k+1}"] = pr & M32, pr >> 32
    for k, j in enumerate((1, 3, 5, 7)):
        pr = al[j] * bl[0]
        env[f"O{2*k}"], env[f"O{2*k+1}"] = pr & M32, pr >> 32
    emulate(MUL, env)
    return sum(env[f"E{j}"] << (32 * j) for j in range(8))


def gen_add():
    """r += b (caller guarantees no overflow of 2^256)"""
    st = Stmt()
    for j in range(8):
        op = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < 7 else "addc.u32")
        st.add(op, reg("r", j), reg("r", j), reg("b", j))
    return [st]


def gen_addk(KL):
    """r += K (constant)"""
    st = Stmt()
    for j in range(8):
        op = "add.cc.u32" if j == 0 else ("addc.cc.u32" if j < 7 else "addc.u32")
        st.add(op, reg("r", j), reg("r", j), imm(KL[j]))
    return [st]


def gen_sub(with_bw):
    """r -= b mod 2^256; optionally bw = 0xffffffff when the true difference is negative"""
    st = Stmt()
    for j in range(8):
        op = "sub.cc.u32" if j == 0 else ("subc.cc.u32" if (j < 7 or with_bw) else "subc.u32")
        st.add(op, reg("r", j), reg("r", j), reg("b", j))
    if with_bw:
        st.add("subc.u32", reg("bw"), imm(0), imm(0))
    return [st]


def gen_subk_borrow(KL):
    """r -= K mod 2^256, bw = 0xffffffff when r < K"""
    st = Stmt()
    for j in range(8):
        op = "sub.cc.u32" if j == 0 else "subc.cc.u32"
        st.add(op, reg("r", j), reg("r", j), imm(KL[j]))
    st.add("subc.u32", reg("bw"), imm(0), imm(0))
    return [st]


MUL = gen_mul()
REDC = gen_redc()
ADD = gen_add()
ADD_P = gen_addk(PL)
ADD_2P = gen_addk(P2L)
SUB = gen_sub(False)
SUB_BW = gen_sub(True)
SUBK_P = gen_subk_borrow(PL)
SUBK_2P = gen_subk_borrow(P2L)


def val(env, name):
    return sum(env[f"{name}{j}"] << (32 * j) for j in range(8))


def selftest(iters=3000):
    rnd = random.Random(0xB200)
    edge = [0, 1, P - 1, P, P + 1, 2 * P - 1, 2 * P, 4 * P - 1, R - P - 1, (1 << 255), M32, R - 1]
    Rinv = pow(R, -1, P)
    # --- mul ---
    cases = []
    for a in edge:
        for b in edge:
            cases.append((a, b))
    for _ in range(iters):
        cases.append((rnd.randrange(4 * P), rnd.randrange(P)))
        cases.append((rnd.randrange(2 * P), rnd.randrange(2 * P)))
        cases.append((rnd.randrange(R - P), rnd.randrange(R)))
    n_ok = 0
    for a, b in cases:
        if a >= R - P:
            continue  # outside the contract
        got = run_mul(a, b)
        assert got % P == a * b * Rinv % P, (hex(a), hex(b))
        assert got < a * b // R + P + 1
        if a * b < P * R:
            assert got < 2 * P
        n_ok += 1
    # --- redc through mul by the raw integer 1 ---
    for a in edge + [rnd.randrange(R - P) for _ in range(iters)]:
        if a >= R - P:
            continue
        got = run_mul(a, 1)
        assert got % P == a * Rinv % P and got <= P, hex(a)
    # --- the dedicated reduction: any a < 2^256 ---
    for a in edge + [rnd.randrange(R) for _ in range(iters)] + [R - 1 - rnd.randrange(1 << 40) for _ in range(50)]:
        got = run_redc(a)
        assert got % P == a * Rinv % P and got <= P, hex(a)
        assert got == (a + ((a * ((-pow(P, -1, R)) % R)) % R) * P) >> 256
    # --- add / sub ---
    for _ in range(iters):
        a, b = rnd.randrange(2 * P), rnd.randrange(2 * P)
        def fresh(x):
            env = {f"r{j}": limbs(x)[j] for j in range(8)}
            env.update({f"b{j}": limbs(b)[j] for j in range(8)})
            env["bw"] = 0
            return env
        env = emulate(ADD, fresh(a)); assert val(env, "r") == a + b
        env = emulate(ADD_P, fresh(a)); assert val(env, "r") == a + P
        env = emulate(ADD_2P, fresh(a)); assert val(env, "r") == a + 2 * P
        env = emulate(SUB, fresh(a + 2 * P)); assert val(env, "r") == a + 2 * P - b
        env = emulate(SUB_BW, fresh(a))
        assert val(env, "r") == (a - b) % R and env["bw"] == (M32 if a < b else 0)
        for K, prog in ((P, SUBK_P), (2 * P, SUBK_2P)):
            env = emulate(prog, fresh(a))
            assert val(env, "r") == (a - K) % R and env["bw"] == (M32 if a < K else 0)
    return n_ok


HEADER = '''// GENERATED by tools/gen_fp.py -- do not edit.  BN254-Fr multi-limb carry chains for sm_100a.
// Each asm statement below was executed instruction by instruction in the generator's PTX
// emulator against big-integer ground truth before being printed.  The carry flag never crosses
// an asm-statement boundary.
#pragma once
#include <stdint.h>

namespace fpgen {

'''


def generate():
    s = HEADER
    s += "// r = a*b/2^256 mod p (lazy: r < a*b/R + p; < 2p when a*b < p*R).  Requires a < 2^256 - p.\n"
    s += "__device__ __forceinline__ void mont_mul(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {\n"
    s += "    uint32_t E[8], O[8];\n"
    s += "#pragma unroll\n    for (int k = 0; k < 4; k++) {      // 64-bit products: one IMAD.WIDE each\n"
    s += "        const unsigned long long e = (unsigned long long)a[2 * k] * b[0], o = (unsigned long long)a[2 * k + 1] * b[0];\n"
    s += "        E[2 * k] = (uint32_t)e;      E[2 * k + 1] = (uint32_t)(e >> 32);\n"
    s += "        O[2 * k] = (uint32_t)o;      O[2 * k + 1] = (uint32_t)(o >> 32);\n"
    s += "    }\n"
    s += emit(MUL)
    s += "#pragma unroll\n    for (int k = 0; k < 8; k++) r[k] = E[k];\n"
    s += "}\n\n"
    s += "// r = a/2^256 mod p, r <= p, for any a < 2^256 (Montgomery reduction without a product)\n"
    s += "__device__ __forceinline__ void redc(uint32_t (&r)[8], const uint32_t (&a)[8]) {\n"
    s += "    uint32_t E[8], O[8];\n"
    s += "#pragma unroll\n    for (int k = 0; k < 8; k++) { E[k] = a[k]; O[k] = 0; }\n"
    s += emit(REDC)
    s += "#pragma unroll\n    for (int k = 0; k < 8; k++) r[k] = E[k];\n"
    s += "}\n\n"
    def fn(sig, doc, prog):
        return f"// {doc}\n__device__ __forceinline__ void {sig} {{\n" + emit(prog) + "}\n\n"
    s += fn("add_ip(uint32_t (&r)[8], const uint32_t (&b)[8])", "r += b (caller guarantees no overflow of 2^256)", ADD)
    s += fn("add_p_ip(uint32_t (&r)[8])", "r += p", ADD_P)
    s += fn("add_2p_ip(uint32_t (&r)[8])", "r += 2p", ADD_2P)
    s += fn("sub_ip(uint32_t (&r)[8], const uint32_t (&b)[8])", "r -= b (caller guarantees r >= b)", SUB)
    s += fn("sub_bw_ip(uint32_t (&r)[8], uint32_t &bw, const uint32_t (&b)[8])", "r -= b mod 2^256; bw = ~0 when r < b (bw must be initialised)", SUB_BW)
    s += fn("sub_p_bw_ip(uint32_t (&r)[8], uint32_t &bw)", "r -= p mod 2^256; bw = ~0 when r < p", SUBK_P)
    s += fn("sub_2p_bw_ip(uint32_t (&r)[8], uint32_t &bw)", "r -= 2p mod 2^256; bw = ~0 when r < 2p", SUBK_2P)
    s += "} // namespace fpgen\n"
    return s


if __name__ == "__main__":
    n = selftest()
    print(f"emulator self-test passed ({n} mul cases)")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "stark_pure_rust_b200", "csrc", "fp_gen.cuh")
    text = generate()
    if "--check" in sys.argv:
        assert open(path).read() == text, "fp_gen.cuh is stale: rerun tools/gen_fp.py"
        print("fp_gen.cuh is current")
    else:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        open(path, "w").write(text)
        print("wrote", os.path.normpath(path))
