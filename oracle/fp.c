/*
 * fp.c — BN254 scalar field, restating what `#[derive(PrimeField)]` generates for
 * /root/reference/packages/ff_utils/src/fp.rs:8-12 (ff_derive 0.10.0: Montgomery form,
 * R = 2^(64*4), little-endian limbs) plus the byte codecs of fp.rs:35-44 and :70-77.
 * ORACLE / TEST INFRASTRUCTURE — see oracle.h.
 */
#include "oracle.h"
#include <string.h>

typedef unsigned __int128 u128;

const uint64_t FP_P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL,
                          0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t P_INV = 0xc2e1f593efffffffULL; /* -p^{-1} mod 2^64 */

const fp_t FP_ZERO = {{0, 0, 0, 0}};
const fp_t FP_ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL,
                      0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
const fp_t FP_R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL,
                     0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};

static inline int geq_p(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > FP_P[i]) return 1;
        if (a[i] < FP_P[i]) return 0;
    }
    return 1;
}

static inline void sub_p(uint64_t a[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - FP_P[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
}

void fp_add(fp_t *r, const fp_t *a, const fp_t *b) {
    uint64_t t[4];
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a->l[i] + b->l[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    /* a,b < p < 2^254 so no carry out of limb 3 */
    if (geq_p(t)) sub_p(t);
    memcpy(r->l, t, sizeof t);
}

void fp_sub(fp_t *r, const fp_t *a, const fp_t *b) {
    uint64_t t[4];
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - borrow;
        t[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    if (borrow) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)t[i] + FP_P[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(r->l, t, sizeof t);
}

void fp_neg(fp_t *r, const fp_t *a) { fp_sub(r, &FP_ZERO, a); }

/* Montgomery product r = a*b/R mod p, laid out like ff_derive's generated code: full 4x4
 * schoolbook product into 8 limbs, then four reduction rounds (montgomery_reduce).  Inputs may be
 * any value < 2^256 as long as one of them is < p (needed by fp_from_bytes_le); the result is
 * fully reduced. */
static inline uint64_t mac(uint64_t acc, uint64_t b, uint64_t c, uint64_t *carry) {
    u128 t = (u128)b * c + acc + *carry;
    *carry = (uint64_t)(t >> 64);
    return (uint64_t)t;
}
static inline uint64_t adc(uint64_t a, uint64_t b, uint64_t *carry) {
    u128 t = (u128)a + b + *carry;
    *carry = (uint64_t)(t >> 64);
    return (uint64_t)t;
}

void fp_mul(fp_t *r, const fp_t *a, const fp_t *b) {
    const uint64_t a0 = a->l[0], a1 = a->l[1], a2 = a->l[2], a3 = a->l[3];
    const uint64_t b0 = b->l[0], b1 = b->l[1], b2 = b->l[2], b3 = b->l[3];
    uint64_t c = 0, r0, r1, r2, r3, r4, r5, r6, r7;
    r0 = mac(0, a0, b0, &c); r1 = mac(0, a0, b1, &c); r2 = mac(0, a0, b2, &c); r3 = mac(0, a0, b3, &c); r4 = c;
    c = 0;
    r1 = mac(r1, a1, b0, &c); r2 = mac(r2, a1, b1, &c); r3 = mac(r3, a1, b2, &c); r4 = mac(r4, a1, b3, &c); r5 = c;
    c = 0;
    r2 = mac(r2, a2, b0, &c); r3 = mac(r3, a2, b1, &c); r4 = mac(r4, a2, b2, &c); r5 = mac(r5, a2, b3, &c); r6 = c;
    c = 0;
    r3 = mac(r3, a3, b0, &c); r4 = mac(r4, a3, b1, &c); r5 = mac(r5, a3, b2, &c); r6 = mac(r6, a3, b3, &c); r7 = c;
    /* reduction */
    uint64_t k, c2;
    k = r0 * P_INV; c = 0;
    (void)mac(r0, k, FP_P[0], &c); r1 = mac(r1, k, FP_P[1], &c); r2 = mac(r2, k, FP_P[2], &c); r3 = mac(r3, k, FP_P[3], &c);
    c2 = 0; r4 = adc(r4, c, &c2);
    k = r1 * P_INV; c = 0;
    (void)mac(r1, k, FP_P[0], &c); r2 = mac(r2, k, FP_P[1], &c); r3 = mac(r3, k, FP_P[2], &c); r4 = mac(r4, k, FP_P[3], &c);
    r5 = adc(r5, c, &c2);
    k = r2 * P_INV; c = 0;
    (void)mac(r2, k, FP_P[0], &c); r3 = mac(r3, k, FP_P[1], &c); r4 = mac(r4, k, FP_P[2], &c); r5 = mac(r5, k, FP_P[3], &c);
    r6 = adc(r6, c, &c2);
    k = r3 * P_INV; c = 0;
    (void)mac(r3, k, FP_P[0], &c); r4 = mac(r4, k, FP_P[1], &c); r5 = mac(r5, k, FP_P[2], &c); r6 = mac(r6, k, FP_P[3], &c);
    r7 = adc(r7, c, &c2);
    uint64_t t[4] = {r4, r5, r6, r7};
    if (c2 || geq_p(t)) sub_p(t);
    memcpy(r->l, t, 4 * sizeof(uint64_t));
}

void fp_sqr(fp_t *r, const fp_t *a) { fp_mul(r, a, a); }

/* ff::Field::pow_vartime: square-and-multiply from the most significant bit of the
 * little-endian u64 exponent limbs */
void fp_pow_limbs(fp_t *r, const fp_t *a, const uint64_t *e, size_t n_limbs) {
    fp_t res = FP_ONE, base = *a;
    for (size_t li = n_limbs; li-- > 0;) {
        for (int bit = 63; bit >= 0; bit--) {
            fp_sqr(&res, &res);
            if ((e[li] >> bit) & 1) fp_mul(&res, &res, &base);
        }
    }
    *r = res;
}

void fp_pow_u64(fp_t *r, const fp_t *a, uint64_t e) { fp_pow_limbs(r, a, &e, 1); }

void fp_inv(fp_t *r, const fp_t *a) {
    /* a^(p-2) */
    uint64_t e[4] = {FP_P[0] - 2, FP_P[1], FP_P[2], FP_P[3]};
    fp_pow_limbs(r, a, e, 4);
}

int fp_eq(const fp_t *a, const fp_t *b) { return memcmp(a->l, b->l, 32) == 0; }
int fp_is_zero(const fp_t *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }

void fp_from_u64(fp_t *r, uint64_t v) {
    fp_t t = {{v, 0, 0, 0}};
    fp_mul(r, &t, &FP_R2);
}

/* integer (<2^256, little-endian limbs) -> Montgomery, reducing mod p */
static void from_raw256(fp_t *r, const uint64_t raw[4]) {
    fp_t t;
    memcpy(t.l, raw, 32);
    fp_mul(r, &t, &FP_R2); /* raw * R^2 / R = raw * R mod p, valid for raw < 2^256 */
}

/* fp.rs:74-76: BigUint::from_bytes_le(value) -> decimal -> Fp::from_str, which multiplies by ten
 * and adds digit by digit IN THE FIELD, i.e. the integer is reduced mod p whatever its length. */
void fp_from_bytes_le(fp_t *r, const uint8_t *b, size_t len) {
    if (len <= 32) {
        uint64_t raw[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < len; i++) raw[i / 8] |= (uint64_t)b[i] << (8 * (i % 8));
        from_raw256(r, raw);
        return;
    }
    /* generic length: Horner over 32-byte digits, most significant first; 2^256 = R = FP_ONE's
     * canonical value, and mont(2^256) = R*R mod p = FP_R2 */
    fp_t acc = FP_ZERO;
    size_t pos = len;
    while (pos > 0) {
        size_t take = pos % 32 ? pos % 32 : 32;
        fp_t digit;
        fp_from_bytes_le(&digit, b + pos - take, take);
        fp_mul(&acc, &acc, &FP_R2);
        fp_add(&acc, &acc, &digit);
        pos -= take;
    }
    *r = acc;
}

void fp_from_bytes_be(fp_t *r, const uint8_t *b, size_t len) {
    uint8_t tmp[64];
    if (len <= sizeof tmp) {
        for (size_t i = 0; i < len; i++) tmp[i] = b[len - 1 - i];
        fp_from_bytes_le(r, tmp, len);
        return;
    }
    fp_t acc = FP_ZERO, c256;
    fp_from_u64(&c256, 256);
    for (size_t i = 0; i < len; i++) {
        fp_t d;
        fp_from_u64(&d, b[i]);
        fp_mul(&acc, &acc, &c256);
        fp_add(&acc, &acc, &d);
    }
    *r = acc;
}

static void to_raw(uint64_t raw[4], const fp_t *a) {
    fp_t one = {{1, 0, 0, 0}}, t;
    fp_mul(&t, a, &one); /* a / R */
    memcpy(raw, t.l, 32);
}

void fp_to_bytes_le(uint8_t out[32], const fp_t *a) {
    uint64_t raw[4];
    to_raw(raw, a);
    for (int i = 0; i < 32; i++) out[i] = (uint8_t)(raw[i / 8] >> (8 * (i % 8)));
}

void fp_to_bytes_be(uint8_t out[32], const fp_t *a) {
    uint8_t le[32];
    fp_to_bytes_le(le, a);
    for (int i = 0; i < 32; i++) out[i] = le[31 - i];
}

void fp_multiplicative_generator(fp_t *r) { fp_from_u64(r, 7); }

/* prove.rs:71-82: g2 = 7^((p-1)/precision) */
void fp_root_of_unity(fp_t *r, uint32_t log_n) {
    uint64_t e[4] = {FP_P[0] - 1, FP_P[1], FP_P[2], FP_P[3]};
    /* shift right by log_n (p-1 = 2^28 * odd, log_n <= 28) */
    for (uint32_t s = 0; s < log_n; s++) {
        for (int i = 0; i < 4; i++) {
            e[i] >>= 1;
            if (i < 3) e[i] |= e[i + 1] << 63;
        }
    }
    fp_t g;
    fp_multiplicative_generator(&g);
    fp_pow_limbs(r, &g, e, 4);
}

/* batch form of fp_to_bytes_le for the ctypes harness (tests / bench): out = n x 32 bytes */
void orc_fp_to_bytes_le_vec(uint8_t *out, const fp_t *a, size_t n) {
    for (size_t i = 0; i < n; i++) fp_to_bytes_le(out + 32 * i, &a[i]);
}
