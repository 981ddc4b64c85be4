// fp.cuh -- BN254 scalar field (the reference's `Fp`, ff_utils/src/fp.rs:8-12) on sm_100a.
//
// Memory format = the reference's in-memory Fp([u64;4]): Montgomery form (x * 2^256 mod p),
// little-endian limbs, 32 bytes, here viewed as 8 x u32 / 2 x uint4.
// Register format = 8 x u32; arithmetic = 32-bit IMAD carry chains (fp_gen.cuh, generated and
// emulator-checked by tools/gen_fp.py).
//
// Lazy reduction: p < 2^254 so 4p < 2^256.  Values between butterflies live in [0, 2p).
//   fp_mul(a,b): a < 4p (strictly: a < 2^256 - p), a*b < p*2^256   -> result in [0, 2p)
//   fp_add    : [0,2p) + [0,2p) -> [0,2p)   (one conditional subtraction of 2p)
//   fp_sub    : [0,2p) - [0,2p) -> [0,2p)   (borrow-masked add of 2p)
//   fp_sub_lazy: a + 2p - b in [0,4p)       (feeds fp_mul's first operand only)
//   fp_canon  : [0,2p) -> [0,p)             (ABI edge, hashing, equality)
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "fp_gen.cuh"
#include "params.h"


#define FP_P_LIMBS  {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}
#define FP_2P_LIMBS {0xe0000002u, 0x87c3eb27u, 0xf372e122u, 0x5067d090u, 0x0302b0bau, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u}
// R mod p : Montgomery form of 1
#define FP_ONE_LIMBS {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}
// R^2 mod p
#define FP_R2_LIMBS {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u}

__device__ __forceinline__ fp fp_zero() {
    fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
    return r;
}
__device__ __forceinline__ fp fp_one() {
    const uint32_t k[8] = FP_ONE_LIMBS;
    fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = k[i];
    return r;
}

__device__ __forceinline__ fp fp_from_u4(uint4 lo, uint4 hi) {
    fp r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
    r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}
__device__ __forceinline__ uint4 fp_lo(const fp &a) { return make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]); }
__device__ __forceinline__ uint4 fp_hi(const fp &a) { return make_uint4(a.l[4], a.l[5], a.l[6], a.l[7]); }

// 32-byte element at index i of a global array (two 128-bit accesses)
__device__ __forceinline__ fp fp_ldg(const uint4 *base, size_t i) {
    uint4 lo = base[2 * i], hi = base[2 * i + 1];
    return fp_from_u4(lo, hi);
}
// read-only path (twiddle tables)
__device__ __forceinline__ fp fp_ldg_ro(const uint4 *base, size_t i) {
    uint4 lo = __ldg(base + 2 * i), hi = __ldg(base + 2 * i + 1);
    return fp_from_u4(lo, hi);
}
__device__ __forceinline__ void fp_stg(uint4 *base, size_t i, const fp &a) {
    base[2 * i] = fp_lo(a);
    base[2 * i + 1] = fp_hi(a);
}

__device__ __forceinline__ fp fp_mul(const fp &a, const fp &b) {
    fp r;
    fpgen::mont_mul(r.l, a.l, b.l);
    return r;
}

// conditional subtraction of 2p: [0,4p) -> [0,2p)
__device__ __forceinline__ fp fp_reduce_2p(const fp &a) {
    fp t = a;
    uint32_t bw = 0;
    fpgen::sub_2p_bw_ip(t.l, bw);
    fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = bw ? a.l[i] : t.l[i];
    return r;
}
// conditional subtraction of p: [0,2p) -> [0,p)
__device__ __forceinline__ fp fp_canon(const fp &a) {
    fp t = a;
    uint32_t bw = 0;
    fpgen::sub_p_bw_ip(t.l, bw);
    fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = bw ? a.l[i] : t.l[i];
    return r;
}
__device__ __forceinline__ fp fp_add(const fp &a, const fp &b) {
    fp t = a;
    fpgen::add_ip(t.l, b.l);
    return fp_reduce_2p(t);
}
// a - b for a,b in [0,2p): result in [0,2p)
__device__ __forceinline__ fp fp_sub(const fp &a, const fp &b) {
    fp t = a;
    uint32_t bw = 0;
    fpgen::sub_bw_ip(t.l, bw, b.l);
    const uint32_t k[8] = FP_2P_LIMBS;
    fp m;
#pragma unroll
    for (int i = 0; i < 8; i++) m.l[i] = k[i] & bw;
    fpgen::add_ip(t.l, m.l);   // carry out of 2^256 cancels the earlier borrow
    return t;
}
// a + 2p - b in [0,4p): only valid as the FIRST operand of fp_mul
__device__ __forceinline__ fp fp_sub_lazy(const fp &a, const fp &b) {
    fp t = a;
    fpgen::add_2p_ip(t.l);
    fpgen::sub_ip(t.l, b.l);
    return t;
}
__device__ __forceinline__ fp fp_neg(const fp &a) {   // [0,2p] -> [0,2p]
    const uint32_t k[8] = FP_2P_LIMBS;
    fp t;
#pragma unroll
    for (int i = 0; i < 8; i++) t.l[i] = k[i];
    fpgen::sub_ip(t.l, a.l);
    return t;
}
// exact halving mod p of a value in [0,2p): (a + (a odd ? p : 0)) >> 1  -> [0, 1.5p)
__device__ __forceinline__ fp fp_half(const fp &a) {
    const uint32_t k[8] = FP_P_LIMBS;
    uint32_t mask = 0u - (a.l[0] & 1u);
    fp m, t = a;
#pragma unroll
    for (int i = 0; i < 8; i++) m.l[i] = k[i] & mask;
    fpgen::add_ip(t.l, m.l);
    fp r;
#pragma unroll
    for (int i = 0; i < 7; i++) r.l[i] = __funnelshift_r(t.l[i], t.l[i + 1], 1);
    r.l[7] = t.l[7] >> 1;
    return r;
}
// Montgomery -> canonical integer (< p): a / R mod p, for any a < 2^256 (the dedicated reduction: half the wide multiplies
// of a product by the raw integer 1; the leaf kernels run one per committed element)
__device__ __forceinline__ fp fp_from_mont(const fp &a) {
    fp r;
    fpgen::redc(r.l, a.l);              // <= p
    return fp_canon(r);
}
// canonical integer (any value < 2^256 - p) -> Montgomery, canonical representative
__device__ __forceinline__ fp fp_to_mont(const fp &a) {
    const uint32_t k[8] = FP_R2_LIMBS;
    fp r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.l[i] = k[i];
    return fp_canon(fp_mul(a, r2));
}
// 32 little-endian bytes (any value < 2^256 ~ 5.3 p) -> field element: int_LE mod p in Montgomery form (fp.rs:70-77 from_bytes_le)
__device__ __forceinline__ fp fp_from_le256(const uint4 *d) {
    const uint4 lo = d[0], hi = d[1];
    fp v;
    v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w;
    v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
    v = fp_reduce_2p(fp_reduce_2p(v));      // < 2p after two conditional subtractions of 2p (5.3 p -> 3.3 p -> < 2p)
    return fp_to_mont(fp_canon(v));
}
__device__ __forceinline__ bool fp_is_zero_canon(const fp &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.l[i];
    return o == 0;
}
__device__ __forceinline__ bool fp_eq_canon(const fp &a, const fp &b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.l[i] ^ b.l[i];
    return o == 0;
}
