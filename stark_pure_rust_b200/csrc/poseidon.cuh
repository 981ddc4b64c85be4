// poseidon.cuh -- the reference's alternative digest on the device: `PoseidonDigest::hash` (commitment/src/poseidon.rs:30-63),
// i.e. neptune 5.1.0 Poseidon with arity 2 (width 3, R_F = 8, R_P = 55, S-box x^5, HashType::MerkleTree, HashMode::Correct)
// over the BLS12-381 scalar field (blstrs 0.4.1), and the kernels of the tree it is used in
// (ParallelMerkleTree<Vec<u8>, PoseidonDigest>, commitment/src/pallarel_merkle_tree.rs:219-253).
//
// One thread computes one digest.  The field is NOT the prover's field: r = 0x73eda753...00000001 has 255 bits, so 4r > 2^256 and
// the lazy [0, 2p) ranges of fp_gen.cuh do not carry over; elements here are 8 x u32 limbs in Montgomery form (R = 2^256), always
// fully reduced; the product is the generated carry-chain code of tools/gen_fp_bls.py (r = 1 mod 2^32, so the Montgomery factor
// of a row is just the negated low word).
// The 189 round constants and the 9 matrix entries (Montgomery form, made on the host: poseidon.cu) are staged in shared memory;
// every thread reads the same word at the same time (broadcast).
#pragma once
#include <stdint.h>

#include "fp_bls_gen.cuh"

#define POS_T 3
#define POS_RF 8
#define POS_RP 55
#define POS_N_RC (POS_T * (POS_RF + POS_RP))
#define POS_CONST_WORDS ((POS_N_RC + POS_T * POS_T + 1) * 8)      // round constants, matrix, R^2 mod r
#define POS_THREADS 128

namespace bls {

__device__ __constant__ uint32_t MOD[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};

// t - r if t >= r (t < 2r)
__device__ __forceinline__ void reduce_once(uint32_t (&t)[8]) {
    uint32_t d[8], bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = t[i];
    blsgen::sub_r_bw_ip(d, bw);
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = bw ? t[i] : d[i];
}
__device__ __forceinline__ bool is_canonical(const uint32_t (&a)[8]) {
    for (int i = 7; i >= 0; i--) {
        if (a[i] < MOD[i]) return true;
        if (a[i] > MOD[i]) return false;
    }
    return false;
}
__device__ __forceinline__ void add(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t *b) {
    uint32_t bb[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        bb[i] = b[i];
        r[i] = a[i];
    }
    blsgen::add_ip(r, bb);          // a + b < 2r < 2^256: no carry out
    reduce_once(r);
}
// Montgomery product a b / 2^256 mod r, operands and result < r: the generated IMAD.WIDE carry chains (fp_bls_gen.cuh, the
// same statement lists as the prover's field) and one conditional subtraction.  (First version: a CIOS loop in plain C with
// 64-bit accumulators -- 2.7e7 digests/s on B200; the chains: see DESIGN.md 4.8.)
__device__ __forceinline__ void mul(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t *b) {
    uint32_t aa[8], bb[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        aa[i] = a[i];
        bb[i] = b[i];
    }
    blsgen::mont_mul(r, aa, bb);    // < 2r
    reduce_once(r);
}
__device__ __forceinline__ void sbox(uint32_t (&x)[8]) {
    uint32_t x2[8], x4[8];
    mul(x2, x, x);
    mul(x4, x2, x2);
    mul(x, x4, x);
}

} // namespace bls

// constants in shared memory: rc[POS_N_RC][8], mds[9][8] (row-major, symmetric), r2[8]
__device__ __forceinline__ void poseidon_stage_consts(uint32_t *sc, const uint32_t *consts) {
    for (int i = threadIdx.x; i < POS_CONST_WORDS; i += blockDim.x) sc[i] = consts[i];
    __syncthreads();
}

// digest of the two canonical scalars a, b (a missing second scalar is 0): canonical little-endian limbs
__device__ __noinline__ void poseidon_hash2(uint32_t (&out)[8], const uint32_t (&a)[8], const uint32_t (&b)[8], const uint32_t *sc) {
    const uint32_t *rc = sc, *mds = sc + POS_N_RC * 8, *r2 = mds + POS_T * POS_T * 8;
    uint32_t st[POS_T][8];
    {
        uint32_t tag[8] = {(1u << 2) - 1, 0, 0, 0, 0, 0, 0, 0};          // HashType::MerkleTree: 2^arity - 1
        bls::mul(st[0], tag, r2);
        bls::mul(st[1], a, r2);
        bls::mul(st[2], b, r2);
    }
#pragma unroll 1
    for (int r = 0; r < POS_RF + POS_RP; r++) {
#pragma unroll
        for (int i = 0; i < POS_T; i++) bls::add(st[i], st[i], rc + (r * POS_T + i) * 8);
        bls::sbox(st[0]);
        if (r < POS_RF / 2 || r >= POS_RF / 2 + POS_RP) {
            bls::sbox(st[1]);
            bls::sbox(st[2]);
        }
        uint32_t nx[POS_T][8];
#pragma unroll
        for (int j = 0; j < POS_T; j++) {
            uint32_t term[8];
            bls::mul(nx[j], st[0], mds + j * 8);                          // M is symmetric: M[i][j] = M[j][i]
            bls::mul(term, st[1], mds + (POS_T + j) * 8);
            bls::add(nx[j], nx[j], term);
            bls::mul(term, st[2], mds + (2 * POS_T + j) * 8);
            bls::add(nx[j], nx[j], term);
        }
#pragma unroll
        for (int j = 0; j < POS_T; j++)
#pragma unroll
            for (int w = 0; w < 8; w++) st[j][w] = nx[j][w];
    }
    uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    bls::mul(out, st[1], one);                                            // out of Montgomery form
}

struct PoseidonLeavesParams {
    const uint8_t *msgs;           // n messages of msg_bytes each (1..64)
    uint4 *out;                    // n digests
    unsigned long long n;
    uint32_t msg_bytes;
    const uint32_t *consts;
    int *err;                      // set to 1 when a 32-byte chunk is not a canonical scalar (the reference panics, poseidon.rs:48)
};

__global__ void __launch_bounds__(POS_THREADS) poseidon_leaves_kernel(PoseidonLeavesParams P) {
    __shared__ uint32_t sc[POS_CONST_WORDS];
    poseidon_stage_consts(sc, P.consts);
    const size_t i = (size_t)blockIdx.x * POS_THREADS + threadIdx.x;
    if (i >= P.n) return;
    uint32_t a[8], b[8];
#pragma unroll
    for (int w = 0; w < 8; w++) a[w] = b[w] = 0;
    const uint8_t *m = P.msgs + i * P.msg_bytes;
    if (P.msg_bytes % 4 == 0) {                                            // word loads (messages start at a multiple of 4 bytes)
        const uint32_t *mw = (const uint32_t *)m;
        for (uint32_t w = 0; w < P.msg_bytes / 4; w++) {
            const uint32_t v = mw[w];
            if (w < 8) a[w] = v;
            else b[w - 8] = v;
        }
    } else {
        for (uint32_t k = 0; k < P.msg_bytes; k++) {
            const uint32_t v = (uint32_t)m[k] << (8 * (k & 3));
            if (k < 32) a[k >> 2] |= v;
            else b[(k - 32) >> 2] |= v;
        }
    }
    if (!bls::is_canonical(a) || !bls::is_canonical(b)) {
        *P.err = 1;
        return;
    }
    uint32_t d[8];
    poseidon_hash2(d, a, b, sc);
    P.out[2 * i] = make_uint4(d[0], d[1], d[2], d[3]);
    P.out[2 * i + 1] = make_uint4(d[4], d[5], d[6], d[7]);
}

// level `level` -> `level + 1` of a tree's node array (merkle_level_off layout): parent = H(left || right)
__global__ void __launch_bounds__(POS_THREADS) poseidon_nodes_kernel(uint4 *nodes, unsigned long long n, uint32_t level, const uint32_t *consts) {
    __shared__ uint32_t sc[POS_CONST_WORDS];
    poseidon_stage_consts(sc, consts);
    const size_t i = (size_t)blockIdx.x * POS_THREADS + threadIdx.x;
    const size_t width = n >> (level + 1);
    if (i >= width) return;
    const uint4 *src = nodes + 2 * (merkle_level_off(n, level) + 2 * i);
    const uint4 a0 = src[0], a1 = src[1], b0 = src[2], b1 = src[3];
    const uint32_t a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint32_t d[8];
    poseidon_hash2(d, a, b, sc);                                          // digests are canonical by construction
    uint4 *dst = nodes + 2 * (merkle_level_off(n, level + 1) + i);
    dst[0] = make_uint4(d[0], d[1], d[2], d[3]);
    dst[1] = make_uint4(d[4], d[5], d[6], d[7]);
}
