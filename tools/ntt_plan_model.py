"""Index-level model of the multi-pass GPU NTT (csrc/ntt.cu) in Python ints.

Checks the decomposition n = n1*...*nm used by the CUDA pass kernel: per pass a DIF sub-NTT of
size 2^b over the stride-`inner` dimension done in radix-8 register rounds, the inter-pass twiddle
w_{n'}^(k*c), in-place storage, and the digit-reversed store of the last pass.  Run directly.
"""
import random, sys
sys.path.insert(0, "oracle")
from py_model import P, ntt, root_of_unity

def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1); x >>= 1
    return r

def rounds_for(b):
    """DIF register rounds from the top bit down: list of (s, r): the thread holds idx = base + e*2^s,
    e<8 (or fewer when b<3) and applies the r stages for bits s+r-1 .. s"""
    out, hi = [], b
    while hi > 0:
        r = min(3, hi)
        s_bits = hi - r          # lowest processed bit
        s = min(s_bits, max(b - 3, 0)) if b >= 3 else 0
        out.append((s_bits, r, s))
        hi -= r
    return out

def sub_ntt_dif(x, b, w_sub):
    """in-place DIF over len 2^b producing bit-reversed order, organised as the kernel's rounds"""
    n = 1 << b
    W = [pow(w_sub, e, P) for e in range(max(n // 2, 1))]
    for (lo, r, s) in rounds_for(b):
        # thread elements: idx = base + e * 2^s where bits [s, s+3) vary (or [0,b) if b<3)
        span = min(3, b)
        for base in range(n):
            if (base >> s) & ((1 << span) - 1):
                continue
            regs = [x[base + (e << s)] for e in range(1 << span)]
            # stages for bits lo+r-1 down to lo ; register bit = bit - s
            for bit in range(lo + r - 1, lo - 1, -1):
                rb = bit - s
                half = 1 << rb
                L = 1 << (bit + 1)             # current block length
                for e in range(1 << span):
                    if e & half:
                        continue
                    idx = base + (e << s)
                    i = idx & (L // 2 - 1)      # position inside half block
                    tw = W[i * (n // L)]
                    a, c = regs[e], regs[e + half]
                    regs[e] = (a + c) % P
                    regs[e + half] = (a - c) * tw % P
            for e in range(1 << span):
                x[base + (e << s)] = regs[e]
    return x

def digitrev(o, bits_list):
    """o has digits (k1 most significant ... k_{m-1}); returns k1 + n1*k2 + ..."""
    ks, rem = [], o
    for b in reversed(bits_list):
        ks.append(rem & ((1 << b) - 1)); rem >>= b
    ks.reverse()          # ks[0] = k1
    out, shift = 0, 0
    for k, b in zip(ks, bits_list):
        out |= k << shift; shift += b
    return out

def gpu_ntt_model(v, w, bits):
    n = len(v); k = sum(bits); assert n == 1 << k
    data = list(v); out = [None] * n
    outer = 1
    for p, b in enumerate(bits):
        n_p = 1 << b
        inner = n // (outer * n_p)
        w_block = pow(w, outer, P)            # root of the block DFT of size n' = n_p*inner
        w_sub = pow(w_block, inner, P)        # root of the size-n_p sub DFT
        last = p == len(bits) - 1
        for o in range(outer):
            for c in range(inner):
                col = [data[o * n_p * inner + j * inner + c] for j in range(n_p)]
                sub_ntt_dif(col, b, w_sub)
                for pos in range(n_p):
                    kk = bitrev(pos, b)
                    val = col[pos]
                    if not last:
                        val = val * pow(w_block, kk * c, P) % P
                        data[o * n_p * inner + kk * inner + c] = val
                    else:
                        out[digitrev(o, bits[:-1]) + outer * kk] = val
        outer *= n_p
    return out

# ---- ONE transform over g devices (ntt_multi in csrc/api.cu, ntt_pass_kernel<.., DIST>) --------------------------------
def plan_bits(log_n, maxb=8):
    m = (log_n + maxb - 1) // maxb
    base, rem = divmod(log_n, m)
    return [base + (1 if i < rem else 0) for i in range(m)]

def log_tile_for(b):
    return 11 if b >= 8 else 10

def digitrev_inv(d, prev_bits):
    """ntt_digitrev_inv of ntt.cuh: block index of the inputs of output column d in the last pass"""
    o = 0
    for b in prev_bits:
        o = (o << b) | (d & ((1 << b) - 1)); d >>= b
    return o

def check_distributed_plan(log_n, g):
    """index model of the tile -> device assignment: every tile of every pass is taken by exactly one device, the loads of
    the passes after the first and the stores of the passes before the last stay inside the device's own slab"""
    lg = g.bit_length() - 1
    n, bits = 1 << log_n, plan_bits(log_n)
    m = len(bits)
    slab = log_n - lg
    log_cc = [log_tile_for(b) - b for b in bits]
    assert m >= 2 and all(6 <= b <= 8 for b in bits) and bits[0] >= lg + log_cc[-1], (log_n, g, bits)
    log_outer = 0
    remote = {"load": 0, "store": 0}
    for p, b in enumerate(bits):
        first, last = p == 0, p == m - 1
        log_inner = log_n - log_outer - b
        log_tiles = log_n - b - log_cc[p]
        lo = bits[0] - lg - log_cc[p] if last else log_tiles - lg
        seen = set()
        for d in range(g):
            for i in range(0, 1 << (log_tiles - lg), max(1, (1 << (log_tiles - lg)) // 97)):      # a sample of the device's CTAs
                tile = (((i >> lo) << lg | d) << lo) | (i & ((1 << lo) - 1))
                assert tile < (1 << log_tiles) and tile not in seen
                seen.add(tile)
                for j in (0, (1 << log_cc[p]) - 1):
                    gl = (tile << log_cc[p]) + j
                    for r in (0, (1 << b) - 1):
                        if not last:
                            o, c = gl >> log_inner, gl & ((1 << log_inner) - 1)
                            e_load = (o << (b + log_inner)) + (r << log_inner) + c
                            e_store = e_load                          # in place (row kk instead of r: same index set)
                        else:
                            e_load = (digitrev_inv(gl, bits[:-1]) << b) + r
                            e_store = gl + (r << log_outer)
                        assert e_load < n and e_store < n
                        if first:
                            remote["load"] += (e_load >> slab) != d
                        else:
                            assert (e_load >> slab) == d, ("load not local", log_n, g, p, tile)
                        if not first and not last:
                            assert (e_store >> slab) == d, ("store not local", log_n, g, p, tile)
                        else:
                            remote["store"] += (e_store >> slab) != d
        # the bijection i -> tile covers all tiles of the pass: each (hi, lo) pair once per device
        assert (1 << (log_tiles - lg)) * g == 1 << log_tiles
        log_outer += b
    assert remote["load"] and remote["store"]              # the exchanges are there (first-pass loads, first / last pass stores)
    return bits


if __name__ == "__main__":
    random.seed(1)
    for log_n in range(20, 29):
        for g in (2, 4, 8):
            check_distributed_plan(log_n, g)
    print("ok distributed plans 2^20..2^28 over 2, 4, 8 devices")
    for bits in [[3], [4], [5], [1], [2], [3, 3], [4, 3], [5, 4], [3, 4, 3], [6, 5], [3, 3, 3, 3], [7, 3], [2, 3]]:
        k = sum(bits); n = 1 << k
        w = root_of_unity(k)
        v = [random.randrange(P) for _ in range(n)]
        assert gpu_ntt_model(v, w, bits) == ntt(v, w), bits
        print("ok", bits)
