/*
 * poseidon.c — the reference's alternative digest, restating
 * /root/reference/packages/commitment/src/poseidon.rs:30-63 (`PoseidonDigest::hash`) and the tree it is
 * used in (commitment/src/pallarel_merkle_tree.rs:219-253, same shape as merkle.c).
 * ORACLE / TEST INFRASTRUCTURE — see oracle.h.
 *
 * The algorithm lives in third-party crates that are not under /root/reference:
 *   neptune 5.1.0 (commitment/Cargo.toml:10)  Poseidon, arity 2 (width t = 3), Strength::Standard,
 *                                             HashType::MerkleTree, HashMode::Correct
 *   blstrs 0.4.1  (commitment/Cargo.toml:13)  the BLS12-381 scalar field
 * Restated from the published construction (Poseidon paper, Filecoin's Poseidon specification):
 *   R_F = 8 full and R_P = 55 partial rounds, S-box x^5;
 *   round constants: Grain LFSR in self-shrinking mode, 80-bit seed = field 1 (2 bits), S-box tag 1 (4 bits -- neptune passes 1),
 *       field size 255 (12), t (12), R_F (10), R_P (10), thirty 1 bits; 160 bits discarded; 255 bits per constant,
 *       big-endian, values >= r rejected;
 *   MDS matrix M[i][j] = 1 / (i + t + j)  (Cauchy matrix with x_i = i, y_j = t + j; symmetric);
 *   state = (2^arity - 1, m_0, m_1); every round: add the round's t constants, x^5 on all elements (full round) or on
 *       element 0 (partial round), multiply by M; digest = element 1 as 32 little-endian bytes.
 * The message (1..64 bytes) is zero-padded to a multiple of 32 bytes; every 32-byte chunk must be a canonical scalar
 * (`Fr::from_bytes_le(..).unwrap()`, poseidon.rs:38-48): a chunk >= r makes the reference panic, here it is an error.
 *
 * Pinned by the reference's own vectors: the four digests of poseidon.rs:66-106 and the root / first sibling of
 * pallarel_merkle_tree.rs:235-246 (tests/test_cpu_oracle.py).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } sc_t; /* Montgomery form, R = 2^256 */

static const uint64_t SC_R[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
static uint64_t SC_NINV;  /* -r^-1 mod 2^64 */
static sc_t SC_R2;        /* 2^512 mod r */

static int sc_geq_r(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > SC_R[i]) return 1;
        if (a[i] < SC_R[i]) return 0;
    }
    return 1;
}
static void sc_sub_r(uint64_t a[4]) {
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - SC_R[i] - bw;
        a[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
}
static void sc_add(sc_t *r, const sc_t *a, const sc_t *b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a->l[i] + b->l[i];
        r->l[i] = (uint64_t)c;
        c >>= 64;
    }
    if (c || sc_geq_r(r->l)) sc_sub_r(r->l); /* a + b < 2r < 2^256, so c == 0 */
}
static void sc_mul(sc_t *r, const sc_t *a, const sc_t *b) {
    uint64_t t[6] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        const uint64_t m = t[0] * SC_NINV;
        c = (u128)m * SC_R[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * SC_R[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    memcpy(r->l, t, 32);
    if (t[4] || sc_geq_r(r->l)) sc_sub_r(r->l);
}
static void sc_from_canonical(sc_t *r, const uint64_t v[4]) {
    sc_t a;
    memcpy(a.l, v, 32);
    sc_mul(r, &a, &SC_R2);
}
static void sc_to_canonical(uint64_t v[4], const sc_t *a) {
    sc_t one = {{1, 0, 0, 0}}, r;
    sc_mul(&r, a, &one);
    memcpy(v, r.l, 32);
}
static void sc_pow(sc_t *r, const sc_t *a, const uint64_t e[4]) {
    sc_t acc, one_c = {{1, 0, 0, 0}};
    sc_from_canonical(&acc, one_c.l);
    for (int i = 255; i >= 0; i--) {
        sc_mul(&acc, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) sc_mul(&acc, &acc, a);
    }
    *r = acc;
}

/* ---- parameters of the arity-2 instance ---------------------------------------------------------------------------- */
#define POS_T 3
#define POS_RF 8
#define POS_RP 55
static sc_t POS_RC[POS_T * (POS_RF + POS_RP)], POS_MDS[POS_T][POS_T];
static int pos_ready;

static uint8_t grain[80];
static int grain_pos; /* state = grain[(grain_pos + i) % 80] */
static int grain_next(void) {
    static const int taps[6] = {62, 51, 38, 23, 13, 0};
    int b = 0;
    for (int k = 0; k < 6; k++) b ^= grain[(grain_pos + taps[k]) % 80];
    grain[grain_pos] = (uint8_t)b; /* the oldest bit leaves, the new one takes the last place */
    grain_pos = (grain_pos + 1) % 80;
    return b;
}
static int grain_bit(void) { /* self-shrinking: of every pair of bits the second is kept when the first is 1 */
    for (;;) {
        const int a = grain_next(), b = grain_next();
        if (a) return b;
    }
}
static void grain_seed(unsigned field, unsigned sbox, unsigned n_bits, unsigned t, unsigned rf, unsigned rp) {
    const unsigned val[7] = {field, sbox, n_bits, t, rf, rp, 0x3fffffffu}, width[7] = {2, 4, 12, 12, 10, 10, 30};
    int k = 0;
    for (int f = 0; f < 7; f++)
        for (int i = (int)width[f] - 1; i >= 0; i--) grain[k++] = (uint8_t)((val[f] >> i) & 1);
    grain_pos = 0;
    for (int i = 0; i < 160; i++) grain_next();
}

static void pos_init(void) {
    if (pos_ready) return;
    uint64_t x = 1; /* Newton iteration for r^-1 mod 2^64 */
    for (int i = 0; i < 6; i++) x *= 2 - SC_R[0] * x;
    SC_NINV = (uint64_t)0 - x;
    sc_t v = {{1, 0, 0, 0}}; /* 2^512 mod r by doubling */
    for (int i = 0; i < 512; i++) {
        const int top = (int)(v.l[3] >> 63);
        for (int j = 3; j > 0; j--) v.l[j] = (v.l[j] << 1) | (v.l[j - 1] >> 63);
        v.l[0] <<= 1;
        if (top || sc_geq_r(v.l)) sc_sub_r(v.l);
    }
    SC_R2 = v;
    grain_seed(1, 1, 255, POS_T, POS_RF, POS_RP);
    for (int n = 0; n < POS_T * (POS_RF + POS_RP);) {
        uint64_t c[4] = {0, 0, 0, 0};
        for (int i = 254; i >= 0; i--) c[i / 64] |= (uint64_t)grain_bit() << (i % 64);
        if (sc_geq_r(c)) continue;
        sc_from_canonical(&POS_RC[n++], c);
    }
    uint64_t e[4]; /* r - 2 */
    memcpy(e, SC_R, 32);
    e[0] -= 2;
    for (int i = 0; i < POS_T; i++)
        for (int j = 0; j < POS_T; j++) {
            const uint64_t d[4] = {(uint64_t)(i + POS_T + j), 0, 0, 0};
            sc_t dm;
            sc_from_canonical(&dm, d);
            sc_pow(&POS_MDS[i][j], &dm, e);
        }
    pos_ready = 1;
}

static void pos_sbox(sc_t *x) {
    sc_t x2, x4;
    sc_mul(&x2, x, x);
    sc_mul(&x4, &x2, &x2);
    sc_mul(x, &x4, x);
}

/* poseidon.rs:30-63.  Returns 0, or -1 where the reference panics (length 0 or > 64, chunk not a canonical scalar). */
int orc_poseidon_hash(uint8_t out[32], const uint8_t *msg, size_t len) {
    if (len == 0 || len > 64) return -1; /* :33 assert, and (len - 1) underflows for an empty message */
    pos_init();
    uint8_t padded[64] = {0};
    memcpy(padded, msg, len);
    const size_t n_in = (len + 31) / 32;
    sc_t st[POS_T];
    const uint64_t tag[4] = {(1u << 2) - 1, 0, 0, 0}; /* HashType::MerkleTree: 2^arity - 1 */
    sc_from_canonical(&st[0], tag);
    for (size_t k = 0; k < 2; k++) {
        uint64_t c[4] = {0, 0, 0, 0};
        if (k < n_in) {
            memcpy(c, padded + 32 * k, 32); /* little-endian host */
            if (sc_geq_r(c)) return -1;
        }
        sc_from_canonical(&st[1 + k], c);
    }
    const sc_t *rc = POS_RC;
    for (int r = 0; r < POS_RF + POS_RP; r++) {
        for (int i = 0; i < POS_T; i++) sc_add(&st[i], &st[i], rc++);
        if (r < POS_RF / 2 || r >= POS_RF / 2 + POS_RP) {
            for (int i = 0; i < POS_T; i++) pos_sbox(&st[i]);
        } else {
            pos_sbox(&st[0]);
        }
        sc_t nx[POS_T];
        for (int j = 0; j < POS_T; j++) {
            sc_t acc, term;
            sc_mul(&acc, &st[0], &POS_MDS[0][j]);
            for (int i = 1; i < POS_T; i++) {
                sc_mul(&term, &st[i], &POS_MDS[i][j]);
                sc_add(&acc, &acc, &term);
            }
            nx[j] = acc;
        }
        memcpy(st, nx, sizeof nx);
    }
    uint64_t d[4];
    sc_to_canonical(d, &st[1]);
    memcpy(out, d, 32);
    return 0;
}

/* ParallelMerkleTree<Vec<u8>, PoseidonDigest> (pallarel_merkle_tree.rs:219-253): merkle.c with the other digest.
 * Returns -1 if a leaf cannot be hashed. */
int orc_poseidon_merkle_gen_proofs(const uint8_t *leaves, size_t leaf_bytes, size_t n, const size_t *indices, size_t n_idx,
                                   uint8_t root[32], uint8_t *nodes_out) {
    if (n == 0 || (n & (n - 1))) abort();
    uint8_t *cur = (uint8_t *)malloc(n * 32);
    for (size_t i = 0; i < n; i++)
        if (orc_poseidon_hash(cur + 32 * i, leaves + i * leaf_bytes, leaf_bytes)) {
            free(cur);
            return -1;
        }
    size_t depth = 0;
    while (((size_t)1 << depth) < n) depth++;
    for (size_t lg = 0; ((size_t)1 << lg) < n; lg++) {
        const size_t interval = (size_t)1 << lg;
        for (size_t i = 0; i < n_idx; i++) {
            const size_t twin = ((indices[i] >> lg) ^ 1) << lg;
            memcpy(nodes_out + (i * depth + lg) * 32, cur + twin * 32, 32);
        }
        for (size_t base = 0; base < n; base += 2 * interval) {
            uint8_t msg[64];
            memcpy(msg, cur + base * 32, 32);
            memcpy(msg + 32, cur + (base + interval) * 32, 32);
            orc_poseidon_hash(cur + base * 32, msg, 64);
        }
    }
    memcpy(root, cur, 32);
    free(cur);
    return 0;
}

/* merkle_tree.rs:25-43 with the Poseidon digest */
int orc_poseidon_merkle_validate(const uint8_t root[32], size_t index, const uint8_t *leaf, size_t leaf_bytes, const uint8_t *nodes,
                                 size_t depth) {
    uint8_t cur[32], msg[64];
    if (orc_poseidon_hash(cur, leaf, leaf_bytes)) return 0;
    for (size_t d = 0; d < depth; d++, index /= 2) {
        memcpy(msg + (index % 2 ? 32 : 0), cur, 32);
        memcpy(msg + (index % 2 ? 0 : 32), nodes + d * 32, 32);
        if (orc_poseidon_hash(cur, msg, 64)) return 0;
    }
    return memcmp(cur, root, 32) == 0;
}
