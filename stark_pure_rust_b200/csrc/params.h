// params.h -- plain-data kernel parameter blocks shared by the host API (api.cu) and the kernel
// translation units.  No device code here.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

// BN254 Fr element in registers / kernel parameters: 8 x u32 little-endian limbs, Montgomery form
struct fp {
    uint32_t l[8];
};

// tile of a pass of width B: 1024 elements (32 KiB of shared memory, 128 threads, 4 CTAs per SM) for B <= 7, 2048 elements
// (256 threads, 2 CTAs per SM) for B = 8 so that a tile row stays 8 contiguous elements (256 B).  Measured: the 7-bit passes
// of the LDE gain 2 % from the smaller tile, the 8-bit passes of a plain 2^24-point transform lose 7 % with it.
#define NTT_LOG_TILE_FOR(B) ((B) >= 8 ? 11 : 10)
#define NTT_MAX_PASSES 6

struct NttPassParams {
    const uint4 *src;              // pass input  (first pass: caller's vectors, len_in each)
    uint4 *dst;                    // pass output
    const uint4 *tw;               // table T[e] = w_T^e, e < 2^tw_log_n
    unsigned long long src_stride; // elements between consecutive polynomials in src / dst
    unsigned long long dst_stride;
    unsigned long long len_in;     // first pass: elements >= len_in read as zero
    unsigned long long n_cols_total; // batch * (n >> B)
    uint32_t log_n;
    uint32_t log_outer;            // bits already transformed by earlier passes
    uint32_t log_inner;            // log_n - log_outer - B
    uint32_t first, last, inverse;
    uint32_t tw_log_n;             // log2 of the table length
    uint32_t tw_log_stride;        // w = w_T^(2^tw_log_stride)
    uint32_t n_polys;              // > 0: CTAs are ordered tile-major / polynomial-minor so that the CTAs that
                                   // need the same inter-pass twiddles run together and share them in L2
    uint32_t n_prev;               // last pass: widths of the earlier passes, in order
    uint32_t prev_bits[NTT_MAX_PASSES];
    uint32_t n_inv[8];             // Montgomery form of n^-1 (inverse transform, last pass)
    // coset mode (low-degree extension by 2^coset_log, lde_dev): the batch holds (column, r) pairs, r = 1 .. 2^coset_log - 1
    // (polynomial id = column * coset_m1 + r - 1).  The first pass reads coefficient j of `column` and scales it by
    // W^(j r) (W = the extended domain's root, w = W^(2^coset_log)); the last pass writes output k to element
    // k * 2^coset_log + r of `column`.  coset_m1 == 0: plain transform.
    uint32_t coset_m1, coset_log;
    // last pass of the coset transforms, cluster variant (cluster == 1): the launch groups the coset_m1 = 7 CTAs that hold the same
    // tile of the seven cosets of one column into one thread-block cluster; after the butterflies they read each other's
    // shared memory (DSMEM) so that every CTA writes whole 2^coset_log-element groups out[8k .. 8k+7] -- 256 B contiguous
    // instead of one 32-byte element per 256 B -- and coset 0 (the input column itself) is filled in by the writer.
    uint32_t cluster;
    const uint4 *c0_src;           // the LDE's input columns (coset 0)
    unsigned long long c0_stride, c0_len;
};

struct MerkleColsParams {
    const uint4 *cols[8];
    uint4 *nodes;
    unsigned long long n;          // leaves (power of two)
    uint32_t nc;                   // columns per leaf, 1..8
};

struct MerkleBytesParams {
    const uint8_t *leaves;
    uint4 *nodes;
    unsigned long long n;
    uint32_t leaf_bytes;
};

// digest offset of level l (l = 0: leaf hashes) inside a tree's node array of 2n - 1 digests
__host__ __device__ inline size_t merkle_level_off(size_t n, uint32_t level) {
    return 2 * n - ((2 * n) >> level);
}

struct FriFoldParams {
    const uint4 *vals;             // n values, Montgomery
    uint4 *col;                    // n/4 outputs, Montgomery canonical
    const uint4 *tw;               // T[e] = w_T^e
    unsigned long long n;
    uint32_t tw_log_n, tw_log_stride;   // layer root w = w_T^(2^tw_log_stride)
    uint32_t special_x[8];         // Montgomery
};
