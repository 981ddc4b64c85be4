// hostfp.h -- scalar BN254-Fr arithmetic on the host for the handful of values that sit between
// kernels: roots of unity and their powers, n^-1, FRI's special_x = from_bytes_le(root) (fri.rs:135).
// Same representation as the device: Montgomery, R = 2^256, 4 x u64 little-endian limbs
// (ff_utils/src/fp.rs:8-12).  Never used for vector work.
#pragma once
#include <stdint.h>
#include <string.h>

namespace hfp {

typedef unsigned __int128 u128;

struct el {
    uint64_t l[4];
};

static const uint64_t PMOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t NINV64 = 0xc2e1f593efffffffULL;
static const el ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const el R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const el ZERO = {{0, 0, 0, 0}};

inline bool geq_p(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] != PMOD[i]) return a[i] > PMOD[i];
    }
    return true;
}
inline void sub_p(uint64_t a[4]) {
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - PMOD[i] - bw;
        a[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
}
inline bool eq(const el &a, const el &b) { return memcmp(a.l, b.l, 32) == 0; }
inline bool is_zero(const el &a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }

// Montgomery product; a*b < p * 2^256 must hold (true when either operand is < p); result < p
inline el mul(const el &a, const el &b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a.l[j] * b.l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * NINV64;
        c = (u128)m * PMOD[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * PMOD[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
        t[5] = 0;
    }
    el r = {{t[0], t[1], t[2], t[3]}};
    while (t[4] || geq_p(r.l)) {      // at most a couple of rounds
        u128 bw = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)r.l[i] - PMOD[i] - bw;
            r.l[i] = (uint64_t)d;
            bw = (d >> 64) & 1;
        }
        t[4] -= (uint64_t)bw;
    }
    return r;
}
inline el sqr(const el &a) { return mul(a, a); }
inline el add(const el &a, const el &b) {
    el r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.l[i] + b.l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_p(r.l)) sub_p(r.l);
    return r;
}
inline el neg(const el &a) {
    if (is_zero(a)) return a;
    el r;
    u128 bw = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)PMOD[i] - a.l[i] - bw;
        r.l[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
    return r;
}
inline el pow_limbs(const el &a, const uint64_t *e, int n_limbs) {
    el r = ONE;
    for (int i = n_limbs * 64 - 1; i >= 0; i--) {
        r = sqr(r);
        if ((e[i / 64] >> (i % 64)) & 1) r = mul(r, a);
    }
    return r;
}
inline el pow_u64(const el &a, uint64_t e) { return pow_limbs(a, &e, 1); }
inline el inv(const el &a) {           // Fermat; a != 0
    uint64_t e[4] = {PMOD[0] - 2, PMOD[1], PMOD[2], PMOD[3]};
    return pow_limbs(a, e, 4);
}
inline el from_u64(uint64_t v) {
    el t = {{v, 0, 0, 0}};
    return mul(t, R2);
}
// integer value of 32 little-endian bytes, reduced mod p, in Montgomery form (fp.rs:74-76)
inline el from_bytes_le32(const uint8_t b[32]) {
    el t;
    memcpy(t.l, b, 32);
    while (geq_p(t.l)) sub_p(t.l);     // 2^256 < 6p
    return mul(t, R2);
}
// canonical little-endian bytes (fp.rs:39-43)
inline void to_bytes_le(uint8_t out[32], const el &a) {
    el one_raw = {{1, 0, 0, 0}};
    el c = mul(a, one_raw);
    memcpy(out, c.l, 32);
}
inline el from_limbs(const uint64_t l[4]) {
    el r;
    memcpy(r.l, l, 32);
    return r;
}

} // namespace hfp
