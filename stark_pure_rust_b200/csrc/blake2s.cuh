// blake2s.cuh -- Blake2s-256 (RFC 7693), unkeyed, default parameters: the reference's only digest
// (commitment/src/utils.rs:5-10 `Blake2s::new(); update; finalize`, selected at
// r1cs-stark/src/main.rs:9).  One thread computes one hash; the state lives in registers.
// Host + device: the host build of the same code serves the Fiat-Shamir sampler.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define B2S_HD __host__ __device__ __forceinline__
#else
#define B2S_HD inline
#endif

namespace b2s {

B2S_HD uint32_t rotr(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
    if (n == 16) return __byte_perm(x, 0, 0x1032);
    if (n == 8) return __byte_perm(x, 0, 0x0321);
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

#define B2S_IV0 0x6A09E667u
#define B2S_IV1 0xBB67AE85u
#define B2S_IV2 0x3C6EF372u
#define B2S_IV3 0xA54FF53Au
#define B2S_IV4 0x510E527Fu
#define B2S_IV5 0x9B05688Cu
#define B2S_IV6 0x1F83D9ABu
#define B2S_IV7 0x5BE0CD19u

// h = IV ^ parameter block (digest_length 32, key 0, fanout 1, depth 1)
B2S_HD void init(uint32_t (&h)[8]) {
    h[0] = B2S_IV0 ^ 0x01010020u; h[1] = B2S_IV1; h[2] = B2S_IV2; h[3] = B2S_IV3;
    h[4] = B2S_IV4; h[5] = B2S_IV5; h[6] = B2S_IV6; h[7] = B2S_IV7;
}

// Pipe balance on the device: 10 of the 12 operations of a G run on the ALU pipe (LOP3 / SHF / PRMT / IADD3), the compiler
// issues the two two-input adds as IMAD on the FMA pipe.  (Tried: moving one / both three-input adds over as well, written as
// multiply-adds by gridDim.z so that ptxas cannot fold them back into IADD3 -- ALU 9 / 8, FMA 4 / 6 operations per G.  Slower on
// B200: the 8-column tree of 2^24 leaves 5.12 -> 5.51 / 5.66 ms, the 1-column tree 1.78 -> 1.91 / 1.96 ms.)
#define B2S_G(a, b, c, d, x, y)            \
    a = a + b + (x); d = rotr(d ^ a, 16);  \
    c = c + d;       b = rotr(b ^ c, 12);  \
    a = a + b + (y); d = rotr(d ^ a, 8);   \
    c = c + d;       b = rotr(b ^ c, 7);

#define B2S_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    B2S_G(v0, v4, v8, v12, m[s0], m[s1])    \
    B2S_G(v1, v5, v9, v13, m[s2], m[s3])    \
    B2S_G(v2, v6, v10, v14, m[s4], m[s5])   \
    B2S_G(v3, v7, v11, v15, m[s6], m[s7])   \
    B2S_G(v0, v5, v10, v15, m[s8], m[s9])   \
    B2S_G(v1, v6, v11, v12, m[s10], m[s11]) \
    B2S_G(v2, v7, v8, v13, m[s12], m[s13])  \
    B2S_G(v3, v4, v9, v14, m[s14], m[s15])

// one compression: h <- F(h, m, t, last).  t = bytes hashed so far including this block (< 2^32).
B2S_HD void compress(uint32_t (&h)[8], const uint32_t (&m)[16], uint32_t t, bool last) {
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = B2S_IV0, v9 = B2S_IV1, v10 = B2S_IV2, v11 = B2S_IV3;
    uint32_t v12 = B2S_IV4 ^ t, v13 = B2S_IV5, v14 = last ? ~B2S_IV6 : B2S_IV6, v15 = B2S_IV7;
    B2S_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2S_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    B2S_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    B2S_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    B2S_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    B2S_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    B2S_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    B2S_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    B2S_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    B2S_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    h[0] ^= v0 ^ v8;  h[1] ^= v1 ^ v9;  h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

// digest of a 64-byte message given as 16 little-endian words (Merkle node: left || right)
B2S_HD void hash64(uint32_t (&out)[8], const uint32_t (&m)[16]) {
    init(out);
    compress(out, m, 64, true);
}

// digest of a 32-byte message (field-element leaf)
B2S_HD void hash32(uint32_t (&out)[8], const uint32_t (&w)[8]) {
    uint32_t m[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 8; i++) { m[i] = w[i]; m[i + 8] = 0; }
    init(out);
    compress(out, m, 32, true);
}

// host: digest of an arbitrary byte string
inline void hash_bytes(uint8_t out[32], const uint8_t *msg, size_t len) {
    uint32_t h[8];
    init(h);
    size_t off = 0;
    uint32_t m[16];
    while (len - off > 64) {
        memcpy(m, msg + off, 64);
        off += 64;
        compress(h, m, (uint32_t)off, false);
    }
    uint8_t blk[64];
    memset(blk, 0, 64);
    if (len > off) memcpy(blk, msg + off, len - off);
    memcpy(m, blk, 64);
    compress(h, m, (uint32_t)len, true);
    memcpy(out, h, 32);
}

} // namespace b2s
