/*
 * blake2s.c — Blake2s-256, unkeyed, default parameters (RFC 7693), restating what
 * /root/reference/packages/commitment/src/utils.rs:5-10 (and the identical
 * fri/src/utils.rs:5-10) obtain from the `blake2` 0.9.1 crate: Blake2s::new / update / finalize.
 * Pinned by the reference KAT at commitment/src/utils.rs:13-24.
 * ORACLE / TEST INFRASTRUCTURE — see oracle.h.
 */
#include "oracle.h"
#include <string.h>

static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                               0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};

static const uint8_t SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
    {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4},
    {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13},
    {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11},
    {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5},
    {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

#define G(a, b, c, d, x, y)      \
    do {                         \
        a = a + b + (x);         \
        d = rotr(d ^ a, 16);     \
        c = c + d;               \
        b = rotr(b ^ c, 12);     \
        a = a + b + (y);         \
        d = rotr(d ^ a, 8);      \
        c = c + d;               \
        b = rotr(b ^ c, 7);      \
    } while (0)

static void compress(uint32_t h[8], const uint8_t block[64], uint64_t t, int last) {
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; i++)
        m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) |
               ((uint32_t)block[4 * i + 2] << 16) | ((uint32_t)block[4 * i + 3] << 24);
    for (int i = 0; i < 8; i++) {
        v[i] = h[i];
        v[i + 8] = IV[i];
    }
    v[12] ^= (uint32_t)t;
    v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] = ~v[14];
    for (int r = 0; r < 10; r++) {
        const uint8_t *s = SIGMA[r];
        G(v[0], v[4], v[8], v[12], m[s[0]], m[s[1]]);
        G(v[1], v[5], v[9], v[13], m[s[2]], m[s[3]]);
        G(v[2], v[6], v[10], v[14], m[s[4]], m[s[5]]);
        G(v[3], v[7], v[11], v[15], m[s[6]], m[s[7]]);
        G(v[0], v[5], v[10], v[15], m[s[8]], m[s[9]]);
        G(v[1], v[6], v[11], v[12], m[s[10]], m[s[11]]);
        G(v[2], v[7], v[8], v[13], m[s[12]], m[s[13]]);
        G(v[3], v[4], v[9], v[14], m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

void orc_blake2s(uint8_t out[32], const uint8_t *msg, size_t len) {
    uint32_t h[8];
    memcpy(h, IV, sizeof h);
    h[0] ^= 0x01010020u; /* digest_length 32, key_length 0, fanout 1, depth 1 */
    uint64_t t = 0;
    /* all blocks but the last */
    while (len > 64) {
        t += 64;
        compress(h, msg, t, 0);
        msg += 64;
        len -= 64;
    }
    uint8_t block[64];
    memset(block, 0, 64);
    if (len) memcpy(block, msg, len);
    t += len;
    compress(h, block, t, 1);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)h[i];
        out[4 * i + 1] = (uint8_t)(h[i] >> 8);
        out[4 * i + 2] = (uint8_t)(h[i] >> 16);
        out[4 * i + 3] = (uint8_t)(h[i] >> 24);
    }
}
