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
    // coset mode (low-degree extension by 2^coset_log): the batch holds (column, coset) pairs, polynomial id =
    // column * coset_cnt + (coset - coset_r0) for cosets coset_r0 .. coset_r0 + coset_cnt - 1.  The first pass reads
    // coefficient j of `column` and scales it by W^(j * coset) (W = the extended domain's root, w = W^(2^coset_log)).
    // coset_cnt == 0: plain transform.  The last pass stores according to coset_store:
    //   NTT_STORE_PLAIN       polynomial-major like a plain transform: output k of (column, coset) at
    //                         dst[polynomial * dst_stride + k] -- the coset-major layout of the prover (ext.cu)
    //   NTT_STORE_INTERLEAVED natural order of the extended domain: element k * 2^coset_log + coset of `column`
    uint32_t coset_cnt, coset_r0, coset_log, coset_store;
    uint32_t coset_dst_cpd, coset_dst_r0;   // NTT_STORE_PLAIN: (column, coset) is stored as polynomial column * coset_dst_cpd + coset - coset_dst_r0
    // ONE transform spread over 2^dist_log_g devices (ntt_pass_kernel<.., DIST = true>, ntt_multi in api.cu): the vector lives in
    // natural-order slabs of 2^dist_log_slab elements, element e on device e >> dist_log_slab; src_tab / dst_tab hold every
    // device's slab (peer pointers).  This launch runs on device dist_dev and takes the tiles whose index carries dist_dev in
    // the bits [dist_blk_lo, dist_blk_lo + dist_log_g): CTA i works on tile ((i >> lo) << (lo + log_g)) | dev << lo | (i & (2^lo - 1)).
    const uint4 *src_tab[8];
    uint4 *dst_tab[8];
    uint32_t dist_log_slab, dist_log_g, dist_dev, dist_blk_lo;
};
#define NTT_STORE_PLAIN 0
#define NTT_STORE_INTERLEAVED 1

#define SB_MAX_DEV 8

struct MerkleColsParams {
    const uint4 *cols[8];
    uint4 *nodes;
    unsigned long long n;          // leaves (power of two)
    uint32_t nc;                   // columns per leaf, 1..8
    uint32_t coset_log_s;          // != 0 (openings only): coset-major columns, leaf i = element (i & 7) << coset_log_s | i >> 3
};

// Leaf hashing of coset-major columns on ONE device of a g-device context (merkle_leaves_ext_kernel): the device holds
// cosets r0 .. r0 + cpd - 1 (r0 = d * cpd, cpd = 8 / g = 2^lv) of every column as arrays of S = 2^log_s values; leaf
// 8 k + r of the tree is the row (col_0[r][k], ..., col_{nc-1}[r][k]).  Thread k hashes its cpd adjacent leaves and reduces
// them lv levels.  Destinations: `single` != NULL -> a standard tree array over N = 8 S leaves (all levels there: one
// device); otherwise levels < lv go to this device's `low` array and the level-lv digest of step k goes to this device's
// `stage[k]`: the host then moves the k-range of every node-range owner to that owner with one contiguous peer copy per
// (source, owner) pair, and the owner interleaves the g sources into level 0 of its subtree (node k g + d).
// (First version: the kernel stored each digest straight into the owner's subtree array -- 32-byte peer stores at a stride of
// 32 g bytes.  At g = 8 that ran at ~25 GB/s over NVLink: the 1-column tree of 2^26 leaves took 10.2 ms on 8 GPUs against
// 6.9 ms on one.)
struct ExtLeavesParams {
    const uint4 *cols[8];
    uint4 *low;
    uint4 *stage;
    uint4 *single;
    uint32_t nc, log_s, cpd, lv, d, g;
};
// offset (in digests) of level l < lv inside a device's `low` array: levels 0 .. l-1 hold S * (cpd >> l') digests each
__host__ __device__ inline size_t ext_low_off(uint32_t log_s, uint32_t cpd, uint32_t l) {
    size_t off = 0;
    for (uint32_t i = 0; i < l; i++) off += ((size_t)(cpd >> i)) << log_s;
    return off;
}

// Openings of a sharded tree, gathered by one kernel on the primary device through peer pointers
struct ExtOpenParams {
    const uint4 *low[SB_MAX_DEV];
    const uint4 *sub[SB_MAX_DEV];
    const uint4 *cols[SB_MAX_DEV][8];
    uint32_t nc, log_s, cpd, lv, g;
};

struct MerkleBytesParams {
    const uint8_t *leaves;
    uint4 *nodes;
    unsigned long long n;
    uint32_t leaf_bytes;
    unsigned long long first, count;   // this launch hashes leaves [first, first + count) (chunks of a pipelined upload; first is a multiple of 1024)
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
    const uint4 *special_root;     // != NULL: special_x is derived on the device from this 32-byte digest (int_LE mod p, fri.rs:135)
                                   // -- the root of the previous layer's column tree, still in its node array: the host never
                                   // waits for it between layers (merkle_leaves_fold_kernel only)
    // coset-major input (merkle_leaves_fold_ext_kernel): vals = this device's cosets r0 .. r0 + cpd - 1 of the layer's
    // values, 2^log_s each (n = 8 * 2^log_s); row i = 8 k + r reads vals[(r - r0) * S + k + j S / 4], j < 4.  Row i of the
    // folded column is position 8 k + r of the next layer's domain, i.e. the same coset: col_local (this device's cosets of
    // the column, S / 4 values each, when the next layer stays sharded) and / or col (natural order on the primary device,
    // when the next layer runs there alone).  The column's tree goes to `nodes` (standard layout on the primary, kernel
    // argument) or, when that is NULL, to low / stage like ExtLeavesParams.
    uint4 *col_local;
    uint4 *low;
    uint4 *stage;
    uint32_t log_s, cpd, lv, d, g;
};
