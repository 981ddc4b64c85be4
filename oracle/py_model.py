"""py_model.py — independent big-int restatement of the stark-pure-rust prover (TEST INFRASTRUCTURE).

A second, deliberately naive restatement (Python ints in canonical form, hashlib.blake2s) of the
same reference semantics the C oracle (oracle/*.c) follows.  Its only job is to pin the C oracle:
the two must agree byte-for-byte on every intermediate and on proof.json.  Nothing in the product
imports this file.  Reference citations are relative to /root/reference/packages/.

    field / codecs   ff_utils/src/fp.rs:8-12, 35-44, 70-77
    NTT              fri/src/fft.rs:150-193, 284-293, 327-379   (natural-order DFT)
    poly utils       fri/src/poly_utils.rs:38-70, 93-102, 409-511
    sampler          fri/src/utils.rs:82-109
    Merkle           commitment/src/merkle_proof_in_place.rs:54-206, merkle_tree.rs:15-58
    FRI              fri/src/fri.rs:46-224
    prover           r1cs-stark/src/prove.rs:14-378, utils.rs:14-524, run.rs:109-452
    parsers          circom2bellman_core/src/reader.rs:4-89, r1cs-stark/src/reader.rs:7-42
"""
import hashlib
import struct

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
GEN = 7
R = 1 << 256
EXT = 8
SPOT = 80


def blake(b):
    return hashlib.blake2s(bytes(b)).digest()


def to_le(x):
    return int(x % P).to_bytes(32, "little")


def from_le(b):
    return int.from_bytes(bytes(b), "little") % P


def from_be(b):
    return int.from_bytes(bytes(b), "big") % P


def to_mont_limbs(x):
    """canonical int -> the reference's in-memory Fp([u64;4]) (Montgomery, R = 2^256)"""
    v = (x % P) * R % P
    return [(v >> (64 * i)) & (2**64 - 1) for i in range(4)]


def from_mont_limbs(l):
    v = sum(int(l[i]) << (64 * i) for i in range(4))
    return v * pow(R, -1, P) % P


def root_of_unity(log_n):
    return pow(GEN, (P - 1) >> log_n, P)


def expand_root_of_unity(w):
    out = [1]
    c = w % P
    while c != 1:
        out.append(c)
        c = c * w % P
    return out


def ntt(vals, w):
    """natural-order DFT out[k] = sum_j v[j] w^(jk), iterative radix-2 (fft.rs:150-193)"""
    n = len(vals)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    a = list(vals)
    for k in range(n):
        rk = int(format(k, "0%db" % log_n)[::-1], 2) if log_n else 0
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    m = 1
    while m < n:
        w_m = pow(w, n // (2 * m), P)
        for k in range(0, n, 2 * m):
            ww = 1
            for j in range(m):
                t = a[k + j + m] * ww % P
                a[k + j + m] = (a[k + j] - t) % P
                a[k + j] = (a[k + j] + t) % P
                ww = ww * w_m % P
        m *= 2
    return a


def best_fft(vals, w, log_n):
    v = list(vals) + [0] * ((1 << log_n) - len(vals))
    assert len(v) == 1 << log_n
    return ntt(v, w)


def inv_best_fft(vals, w, log_n):
    n = 1 << log_n
    v = list(vals) + [0] * (n - len(vals))
    assert len(v) == n
    inv_n = pow(n, -1, P)
    return [x * inv_n % P for x in ntt(v, pow(w, -1, P))]


def multi_inv(vals):
    partials = [1]
    for v in vals:
        partials.append(partials[-1] * (v if v else 1) % P)
    inv = pow(partials[-1], -1, P)
    out = [0] * len(vals)
    for i in range(len(vals) - 1, -1, -1):
        out[i] = partials[i] * inv % P if vals[i] else 0
        inv = inv * (vals[i] if vals[i] else 1) % P
    return out


def eval_poly_at(poly, x):
    y, pw = 0, 1
    for c in poly:
        y = (y + pw * c) % P
        pw = pw * x % P
    return y


def lagrange_interp(xs, ys):
    """unique interpolant, low degree first (poly_utils.rs:409-439)"""
    n = len(xs)
    root = [1]
    for x in xs:  # multiply by (X - x); root is low-degree-first here
        root = [(-x * root[0]) % P] + [(root[i - 1] - x * root[i]) % P for i in range(1, len(root))] + [root[-1]]
    out = [0] * n
    for i in range(n):
        # synthetic division by (X - xs[i])
        q = [0] * n
        carry = root[n]
        q[n - 1] = carry
        for d in range(n - 2, -1, -1):
            carry = (root[d + 1] + xs[i] * carry) % P
            q[d] = carry
        denom = eval_poly_at(q, xs[i])
        s = ys[i] * pow(denom, -1, P) % P
        for j in range(n):
            out[j] = (out[j] + q[j] * s) % P
    return out


def interp4_eval(xs4, ys4, x):
    """value at x of the degree<4 interpolant through 4 points (what multi_interp_4+eval_quartic yield)"""
    return eval_poly_at(lagrange_interp(xs4, ys4), x)


def get_pseudorandom_indices(seed, modulus, count, excl=0):
    assert modulus < 2**24
    data = bytes(seed)
    while len(data) < 4 * count:
        data += blake(data[-32:])
    vals = [int.from_bytes(data[i : i + 4], "big") for i in range(0, 4 * count, 4)]
    if excl == 0:
        return [v % modulus for v in vals]
    real = (modulus * (excl - 1) // excl) & 0xFFFFFFFF
    out = []
    for v in vals:
        t = v % real
        out.append(t + 1 + t // (excl - 1))
    return out


def merkle_levels(leaves):
    lv = [[blake(l) for l in leaves]]
    while len(lv[-1]) > 1:
        p = lv[-1]
        lv.append([blake(p[i] + p[i + 1]) for i in range(0, len(p), 2)])
    return lv


def merkle_root(leaves):
    return merkle_levels(leaves)[-1][0]


def merkle_proofs(leaves, indices):
    lv = merkle_levels(leaves)
    out = []
    for idx in indices:
        nodes, i = [], idx
        for d in range(len(lv) - 1):
            nodes.append(lv[d][i ^ 1])
            i >>= 1
        out.append((bytes(leaves[idx]), nodes))
    return lv[-1][0], out


def validate(root, index, leaf, nodes):
    cur = blake(leaf)
    for nd in nodes:
        cur = blake(cur + nd) if index % 2 == 0 else blake(nd + cur)
        index //= 2
    return cur == root


def fri_fold(values, w, special_x):
    n = len(values)
    q = n // 4
    xs = expand_root_of_unity(w)
    assert len(xs) == n
    return [interp4_eval([xs[i + q * j] for j in range(4)], [values[i + q * j] for j in range(4)], special_x) for i in range(q)]


def prove_low_degree(values, w, max_deg_plus_1, excl):
    """returns list of ('Middle', root2, column_branches, poly_branches) / ('Last', [bytes])"""
    out = []
    values = list(values)
    while True:
        if max_deg_plus_1 <= 16:
            out.append(("Last", [to_le(v) for v in values]))
            return out
        n = len(values)
        enc = [to_le(v) for v in values]
        m_root = merkle_root(enc)
        special_x = from_le(m_root)
        column = fri_fold(values, w, special_x)
        enc_col = [to_le(v) for v in column]
        m2_root, _ = merkle_proofs(enc_col, [])
        ys = get_pseudorandom_indices(m2_root, n // 4, 40, excl)
        _, col_br = merkle_proofs(enc_col, ys)
        pos = [y + (n // 4) * j for y in ys for j in range(4)]
        _, poly_br = merkle_proofs(enc, pos)
        out.append(("Middle", m2_root, col_br, poly_br))
        values, w, max_deg_plus_1 = column, pow(w, 4, P), max_deg_plus_1 // 4


# ---- parsers -----------------------------------------------------------------------------------
def read_r1cs(b):
    assert b[:4] == b"r1cs"
    ver, nsec = struct.unpack_from("<II", b, 4)
    assert ver == 1 and nsec == 3
    off = 12
    st, _ = struct.unpack_from("<IQ", b, off)
    assert st == 1
    off += 12
    (fs,) = struct.unpack_from("<I", b, off)
    off += 4
    prime = int.from_bytes(b[off : off + 32], "little")
    off += 32
    n_wires, n_pub_out, n_pub_in, n_priv = struct.unpack_from("<IIII", b, off)
    off += 16
    (n_labels,) = struct.unpack_from("<Q", b, off)
    off += 8
    (n_cons,) = struct.unpack_from("<I", b, off)
    off += 4
    st, _ = struct.unpack_from("<IQ", b, off)
    assert st == 2
    off += 12
    cons = []
    for _ in range(n_cons):
        fac = []
        for _ in range(3):
            (n,) = struct.unpack_from("<I", b, off)
            off += 4
            items = []
            for _ in range(n):
                (wid,) = struct.unpack_from("<I", b, off)
                off += 4
                items.append((wid, int.from_bytes(b[off : off + 32], "little") % P))
                off += 32
            fac.append(items)
        cons.append(fac)
    return dict(prime=prime, n_wires=n_wires, n_pub_out=n_pub_out, n_pub_in=n_pub_in, n_constraints=n_cons, constraints=cons)


def read_witness(b):
    assert struct.unpack_from("<I", b, 0)[0] == 1936618615
    off = 4 + 5 * 4
    (fs,) = struct.unpack_from("<I", b, off)
    off += 4 + fs
    (n_wires,) = struct.unpack_from("<I", b, off)
    off += 16
    return [int.from_bytes(b[off + i * fs : off + (i + 1) * fs], "little") % P for i in range(n_wires)]


def build_trace(r1cs, witness):
    n_wires = r1cs["n_wires"]
    lists = {k: dict(wit=[], coef=[], trace=[]) for k in range(3)}
    uses = [[] for _ in range(n_wires)]
    last = []
    acc = 0
    for fac in r1cs["constraints"]:
        n = max(len(f) for f in fac)
        for k in range(3):
            t = 0
            for i in range(n):
                if i < len(fac[k]):
                    w, c = fac[k][i]
                    t = (t + c * witness[w]) % P
                else:
                    w, c = n_wires - 1, 0
                uses[w].append((k, len(lists[k]["coef"])))
                lists[k]["wit"].append(witness[w])
                lists[k]["coef"].append(c)
                lists[k]["trace"].append(t)
        acc += n
        last.append(acc - 1)
    a = acc
    tr = dict(
        witness_trace=sum((lists[k]["wit"] for k in range(3)), []),
        computational_trace=sum((lists[k]["trace"] for k in range(3)), []),
        coefficients=sum((lists[k]["coef"] for k in range(3)), []),
    )
    os_ = 3 * a
    flag0, flag1, flag2 = [1] * os_, [1] * os_, [0] * os_
    for l in last:
        f = (l + 1) % a
        flag1[f] = flag1[f + a] = flag1[f + 2 * a] = 0
        flag2[l] = 1
    perm = [0] * os_
    for vs in uses:
        if not vs:
            continue
        old = a * vs[-1][0] + vs[-1][1]
        for k, v in vs:
            w = a * k + v
            perm[w] = old
            old = w
    n_pub = 1 + r1cs["n_pub_in"] + r1cs["n_pub_out"]
    pfi = [(w, a * uses[w][0][0] + uses[w][0][1]) for w in range(n_pub) if uses[w]]
    tr.update(flag0=flag0, flag1=flag1, flag2=flag2, permuted_indices=perm, public_wires=witness[:n_pub], pfi=pfi, original_steps=os_)
    return tr


def log2_ceil_quirk(v):
    l = 1
    while v > 1:
        v //= 2
        l += 1
    return l


def mk_r1cs_proof(tr, taps=None):
    os_ = tr["original_steps"]
    log_steps = log2_ceil_quirk(os_ - 1)
    S = max(1 << log_steps, 8)
    N = S * EXT
    log_n = log_steps + 3
    perm = list(tr["permuted_indices"]) + list(range(os_, S))
    pad = lambda v: list(v) + [0] * (S - len(v))
    coeffs, wit, comp = pad(tr["coefficients"]), pad(tr["witness_trace"]), pad(tr["computational_trace"])
    g2 = pow(GEN, (P - 1) // N, P)
    xs = expand_root_of_unity(g2)
    sk = N // S
    g1 = xs[sk]
    lde = lambda col: best_fft(inv_best_fft(col, g1, log_steps), g2, log_n)
    k_ev, f0, f1, f2 = lde(coeffs), lde(tr["flag0"]), lde(tr["flag1"]), lde(tr["flag2"])
    s_ev, p_ev = lde(wit), lde(comp)
    z_ev = best_fft([P - 1] + [0] * (S - 1) + [1], g2, log_n)
    o3 = os_ // 3
    q1 = [f0[j] * (p_ev[j] - f1[j] * p_ev[(j - sk) % N] - k_ev[j] * s_ev[j]) % P for j in range(N)]
    q2 = [f2[j] * (p_ev[(j + 2 * o3 * sk) % N] - p_ev[j] * p_ev[(j + o3 * sk) % N]) % P for j in range(N)]
    idx_ev, pidx_ev = lde(list(range(S))), lde(perm)
    a_root = merkle_root([struct.pack("<Q", perm[j]) + to_le(wit[j]) for j in range(S)])
    rnd = get_pseudorandom_indices(a_root, N, 24, 0)
    r = [from_le(b"".join(struct.pack(">I", v) for v in rnd[8 * i : 8 * i + 8])) for i in range(3)]
    nm, dn, an, ad = [], [], 1, 1
    for j in range(S):
        an = an * (r[0] + r[1] * idx_ev[j * sk] + r[2] * wit[j]) % P
        ad = ad * (r[0] + r[1] * pidx_ev[j * sk] + r[2] * wit[j]) % P
        nm.append(an)
        dn.append(ad)
    a_mini = [x * y % P for x, y in zip(nm, multi_inv(dn))]
    a_ev = lde(a_mini)
    q3 = [
        (a_ev[j] * (r[0] + r[1] * pidx_ev[j] + r[2] * s_ev[j]) - a_ev[(j - sk) % N] * (r[0] + r[1] * idx_ev[j] + r[2] * s_ev[j])) % P
        for j in range(N)
    ]
    inv_z = multi_inv(z_ev)
    for j in range(N):
        if inv_z[j] == 0:
            assert q1[j] == 0 and q2[j] == 0 and q3[j] == 0, j
    d1 = [a * b % P for a, b in zip(q1, inv_z)]
    d2 = [a * b % P for a, b in zip(q2, inv_z)]
    d3 = [a * b % P for a, b in zip(q3, inv_z)]
    xv = [xs[sk * w] for _, w in tr["pfi"]]
    yv = [tr["public_wires"][k] for k, _ in tr["pfi"]]
    interp2 = lagrange_interp(xv, yv)
    i2 = [eval_poly_at(interp2, x) for x in xs]
    zb2 = []
    for x in xs:
        acc = 1
        for xw in xv:
            acc = acc * (x - xw) % P
        zb2.append(acc)
    x_last = xs[N - sk]
    zb3 = [(x - x_last) % P for x in xs]
    b2 = [(s - i) * z % P for s, i, z in zip(s_ev, i2, multi_inv(zb2))]
    b3 = [(a - 1) * z % P for a, z in zip(a_ev, multi_inv(zb3))]
    cols = [p_ev, a_ev, s_ev, d1, d2, d3, b2, b3]
    m_leaves = [b"".join(to_le(c[j]) for c in cols) for j in range(N)]
    m_root = merkle_root(m_leaves)
    k = [1] + [from_be(blake(m_root + bytes([i]))) for i in range(1, 11)]
    gs = xs[S]
    l_ev, pw = [], 1
    for j in range(N):
        l_ev.append(
            (k[0] * d1[j] + k[1] * d2[j] + k[2] * d3[j] + k[3] * p_ev[j] + k[4] * p_ev[j] * pw + k[5] * b2[j] + k[6] * b2[j] * pw
             + k[7] * b3[j] + k[8] * b3[j] * pw + k[9] * a_ev[j] + k[10] * s_ev[j]) % P
        )
        pw = pw * gs % P
    l_leaves = [to_le(v) for v in l_ev]
    l_root = merkle_root(l_leaves)
    positions = get_pseudorandom_indices(l_root, N, SPOT, sk)
    _, lc_br = merkle_proofs(l_leaves, positions)
    aug = []
    for j in positions:
        aug += [j, (j + N - sk) % N, (j + o3 * sk) % N, (j + o3 * 2 * sk) % N]
    _, main_br = merkle_proofs(m_leaves, aug)
    fri = prove_low_degree(l_ev, g2, N // 4, sk)
    if taps is not None:
        taps.update(S=S, N=N, g2=g2, lde=[k_ev, f0, f1, f2, s_ev, p_ev, idx_ev, pidx_ev, a_ev], tree_cols=cols, l_evals=l_ev, r=r, k=k, positions=positions)
    return dict(m_root=m_root, l_root=l_root, a_root=a_root, main_branches=main_br, linear_comb_branches=lc_br, fri_proof=fri)


# ---- serde_json::to_string layout (run.rs:549; utils.rs:122-130; fri.rs:16-26; merkle_tree.rs:14-18)
def _b(x):
    return "[" + ",".join(str(v) for v in x) + "]"


def _branches(brs):
    return "[" + ",".join('{"leaf":%s,"nodes":[%s]}' % (_b(leaf), ",".join(_b(n) for n in nodes)) for leaf, nodes in brs) + "]"


def fri_json(fri):
    parts = []
    for layer in fri:
        if layer[0] == "Last":
            parts.append('{"Last":{"last":[%s]}}' % ",".join(_b(v) for v in layer[1]))
        else:
            _, root2, cb, pb = layer
            parts.append('{"Middle":{"root2":%s,"column_branches":%s,"poly_branches":%s}}' % (_b(root2), _branches(cb), _branches(pb)))
    return "[" + ",".join(parts) + "]"


def proof_json(pr):
    return '{"m_root":%s,"l_root":%s,"a_root":%s,"main_branches":%s,"linear_comb_branches":%s,"fri_proof":%s}' % (
        _b(pr["m_root"]), _b(pr["l_root"]), _b(pr["a_root"]), _branches(pr["main_branches"]), _branches(pr["linear_comb_branches"]), fri_json(pr["fri_proof"]))


def prove_files(r1cs_path, wtns_path, taps=None):
    r1cs = read_r1cs(open(r1cs_path, "rb").read())
    assert r1cs["prime"] == P
    wit = read_witness(open(wtns_path, "rb").read())
    assert wit[0] == 1
    return proof_json(mk_r1cs_proof(build_trace(r1cs, wit), taps))


if __name__ == "__main__":
    import sys

    js = prove_files(sys.argv[1], sys.argv[2])
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(js)
    print(hashlib.sha256(js.encode()).hexdigest(), len(js))


# ---- the alternative digest: neptune 5.1.0 Poseidon, arity 2, over the BLS12-381 scalar field ----------------------------
# (commitment/src/poseidon.rs:30-63; second, independent restatement next to oracle/poseidon.c -- big-int arithmetic)
BLS_R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def _grain_bits(field, sbox, n_bits, t, rf, rp):
    st = []
    for v, w in ((field, 2), (sbox, 4), (n_bits, 12), (t, 12), (rf, 10), (rp, 10), ((1 << 30) - 1, 30)):
        st += [(v >> i) & 1 for i in range(w - 1, -1, -1)]

    def step():
        b = st[62] ^ st[51] ^ st[38] ^ st[23] ^ st[13] ^ st[0]
        st.pop(0)
        st.append(b)
        return b
    for _ in range(160):
        step()
    while True:
        a, b = step(), step()
        if a:
            yield b


_POSEIDON = {}


def poseidon_params(arity=2, rf=8, rp=55):
    if arity not in _POSEIDON:
        t = arity + 1
        bits = _grain_bits(1, 1, 255, t, rf, rp)
        rc = []
        while len(rc) < t * (rf + rp):
            v = 0
            for _ in range(255):
                v = (v << 1) | next(bits)
            if v < BLS_R:
                rc.append(v)
        mds = [[pow(i + t + j, -1, BLS_R) for j in range(t)] for i in range(t)]
        _POSEIDON[arity] = (rc, mds)
    return _POSEIDON[arity]


def poseidon_digest(msg, rf=8, rp=55):
    """PoseidonDigest::hash: 1..64 bytes -> 32 bytes"""
    if not 0 < len(msg) <= 64:
        raise ValueError("message length")
    pad = bytes(msg) + bytes((-len(msg)) % 32)
    ins = [int.from_bytes(pad[i:i + 32], "little") for i in range(0, len(pad), 32)]
    if any(x >= BLS_R for x in ins):
        raise ValueError("chunk is not a canonical scalar")
    rc, mds = poseidon_params(2, rf, rp)
    st = [3] + ins + [0] * (2 - len(ins))
    for r in range(rf + rp):
        st = [(st[i] + rc[3 * r + i]) % BLS_R for i in range(3)]
        if r < rf // 2 or r >= rf // 2 + rp:
            st = [pow(x, 5, BLS_R) for x in st]
        else:
            st[0] = pow(st[0], 5, BLS_R)
        st = [sum(st[i] * mds[i][j] for i in range(3)) % BLS_R for j in range(3)]
    return st[1].to_bytes(32, "little")
