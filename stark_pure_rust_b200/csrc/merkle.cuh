// merkle.cuh -- Blake2s Merkle commitment kernels.
//
// Replaces commitment/src/merkle_proof_in_place.rs: leaf hashing (:128-131), the in-place binary
// reduction parent = H(left || right) (:78-98) and sibling collection (:78-82, :191-197).  Unlike
// the reference (which rebuilds the whole tree on every gen_proofs call) every level is stored
// once: digests of level l (l = 0: leaf hashes) start at digest offset 2n - (2n >> l), so an
// opening is a gather of log2(n) siblings.
//
// One thread = one hash, state in registers.  Each thread owns 2^LV adjacent inputs and reduces
// them LV levels up in registers, so all 32 lanes stay busy on every level of a launch and a
// whole tree takes ceil(log2(n) / 3) launches.  Bound: 32-bit ALU pipe (~1000 IADD3/LOP3/SHF per
// compression); HBM traffic is the leaf bytes in plus 64 B per leaf of stored levels.
#pragma once
#include "blake2s.cuh"
#include "fp.cuh"
#include "fri_fold.cuh"
#include "params.h"

struct digest_t {
    uint32_t w[8];
};

__device__ __forceinline__ void digest_store(uint4 *nodes, size_t idx, const uint32_t (&d)[8]) {
    nodes[2 * idx] = make_uint4(d[0], d[1], d[2], d[3]);
    nodes[2 * idx + 1] = make_uint4(d[4], d[5], d[6], d[7]);
}
__device__ __forceinline__ void digest_load(uint32_t (&d)[8], const uint4 *nodes, size_t idx) {
    uint4 a = nodes[2 * idx], b = nodes[2 * idx + 1];
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
    d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}


// Digest scratch in shared memory: digest i, word w of reducing thread t lives at sd[(i * 8 + w) * MERKLE_PITCH + t]
// (word-interleaved by thread: the lanes of a warp stay in distinct banks).  Keeping the 2^LV digests there instead of
// in a register array lets the loops below stay rolled, so the kernel holds ONE copy of the leaf compression
// and ONE of the node compression (~40 KB of SASS) instead of 2^LV + 2^LV - 1 inlined copies (290 KB, which
// missed the instruction cache on every iteration: ncu showed 3.3 "no instruction" stall cycles per issue).
//
// Work assignment inside a CTA (128 threads, 128 << LV consecutive inputs): in the PRODUCE phase thread t hashes inputs
// base + i * 128 + t, i < 2^LV, so the lanes of a warp read consecutive leaves / rows (coalesced 32-byte elements; the
// first version gave thread t the 2^LV adjacent inputs it later reduces, a 256-byte lane stride that fetched every
// 128-byte line twice: DRAM read 1.75x the column bytes in ncu) and hands the digest to the thread that reduces it
// (input l of the CTA belongs to reducer l >> LV, slot l & (2^LV - 1)); after a barrier thread t reduces its 2^LV
// adjacent digests LV levels up as before.  The hand-over stores are 2-way bank conflicted (pitch 129), 8 words per hash.
#define MERKLE_THREADS 128
#define MERKLE_PITCH 129
#define MERKLE_SMEM_WORDS(LV) ((1 << (LV)) * 8 * MERKLE_PITCH)
__device__ __forceinline__ void sd_put(uint32_t *sd, int i, const uint32_t (&d)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++) sd[(i * 8 + k) * MERKLE_PITCH] = d[k];
}
// digest of input l (0 <= l < 128 << LV) of this CTA -> its reducer's slot
template <int LV>
__device__ __forceinline__ void sd_hand_over(uint32_t *sd_all, uint32_t l, const uint32_t (&d)[8]) {
    sd_put(sd_all + (l >> LV), (int)(l & ((1u << LV) - 1)), d);
}

// reduce the 2^LV digests of this thread LV levels up, storing every intermediate level.
// digest i is node (first + i) of level `level`; n = number of leaves of the tree.
template <int LV>
__device__ __forceinline__ void merkle_reduce_smem(uint32_t *sd, uint4 *nodes, size_t n, uint32_t level, size_t first) {
#pragma unroll 1
    for (int s = 0; s < LV; s++) {
        const int cnt = 1 << (LV - 1 - s);      // nodes produced at this step
        const size_t off = merkle_level_off(n, level + s + 1), base = first >> (s + 1);
#pragma unroll 1
        for (int i = 0; i < cnt; i++) {
            uint32_t m[16], o[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                m[k] = sd[((2 * i) * 8 + k) * MERKLE_PITCH];
                m[k + 8] = sd[((2 * i + 1) * 8 + k) * MERKLE_PITCH];
            }
            b2s::hash64(o, m);
            sd_put(sd, i, o);                   // i <= 2i: never overwrites an unread child
            digest_store(nodes, off + base + i, o);
        }
    }
}

// ---- leaves = rows of NC field columns (Montgomery in memory) -------------------------------
// leaf i = to_bytes_le(col_0[i]) || ... || to_bytes_le(col_{NC-1}[i])   (prove.rs:235-258 for the
// 8-column m_tree, prove.rs:324-327 / fri.rs:120-123,165-168 for single-column trees).

__device__ __forceinline__ void merkle_leaf_from_cols(uint32_t (&out)[8], const MerkleColsParams &P, size_t i) {
    b2s::init(out);
    uint32_t m[16];
#pragma unroll 1
    for (uint32_t k = 0; k < P.nc; k += 2) {
        fp v = fp_from_mont(fp_ldg(P.cols[k], i));
#pragma unroll
        for (int w = 0; w < 8; w++) m[w] = v.l[w];
        const bool pair = k + 1 < P.nc;
        if (pair) {
            fp u = fp_from_mont(fp_ldg(P.cols[k + 1], i));
#pragma unroll
            for (int w = 0; w < 8; w++) m[8 + w] = u.l[w];
        } else {
#pragma unroll
            for (int w = 0; w < 8; w++) m[8 + w] = 0;
        }
        const uint32_t done = pair ? k + 2 : k + 1;
        b2s::compress(out, m, 32u * done, done == P.nc);
    }
}

template <int LV>
__global__ void __launch_bounds__(MERKLE_THREADS, 4) merkle_leaves_cols_kernel(const __grid_constant__ MerkleColsParams P) {
    __shared__ uint32_t sd_all[MERKLE_SMEM_WORDS(LV)];
    const size_t base = ((size_t)blockIdx.x * MERKLE_THREADS) << LV;
#pragma unroll 1
    for (int i = 0; i < (1 << LV); i++) {
        const uint32_t l = i * MERKLE_THREADS + threadIdx.x;
        if (base + l < P.n) {
            uint32_t h[8];
            merkle_leaf_from_cols(h, P, base + l);
            digest_store(P.nodes, base + l, h);
            sd_hand_over<LV>(sd_all, l, h);
        }
    }
    __syncthreads();
    const size_t first = base + ((size_t)threadIdx.x << LV);
    if (first < P.n) merkle_reduce_smem<LV>(sd_all + threadIdx.x, P.nodes, P.n, 0, first);
}

// ---- FRI: fold fused with the next layer's leaf hashing (fri.rs:141-172) ----------------------------------------
// leaf i of the column tree = to_bytes_le(column[i]) where column[i] is folded from four values of the current layer;
// the column is also written out (it is the next layer's input and the opened leaves are re-read from it).
template <int LV>
__global__ void __launch_bounds__(MERKLE_THREADS, 4) merkle_leaves_fold_kernel(const __grid_constant__ FriFoldParams F, uint4 *nodes) {
    __shared__ uint32_t sd_all[MERKLE_SMEM_WORDS(LV)];
    const size_t q = F.n >> 2;
    const size_t base = ((size_t)blockIdx.x * MERKLE_THREADS) << LV;
    fp sx;
    if (F.special_root) {
        sx = fp_from_le256(F.special_root);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) sx.l[k] = F.special_x[k];
    }
    const unsigned long long nT = 1ull << F.tw_log_n;
    const fp iota_inv = fp_ldg_ro(F.tw, (nT - ((unsigned long long)q << F.tw_log_stride)) & (nT - 1));
#pragma unroll 1
    for (int i = 0; i < (1 << LV); i++) {
        const uint32_t l = i * MERKLE_THREADS + threadIdx.x;
        if (base + l < q) {
            fp r = fri_fold_row(F, base + l, sx, iota_inv);
            fp_stg(F.col, base + l, r);
            fp c = fp_from_mont(r);
            uint32_t h[8];
            b2s::hash32(h, c.l);
            digest_store(nodes, base + l, h);
            sd_hand_over<LV>(sd_all, l, h);
        }
    }
    __syncthreads();
    const size_t first = base + ((size_t)threadIdx.x << LV);
    if (first < q) merkle_reduce_smem<LV>(sd_all + threadIdx.x, nodes, q, 0, first);
}

// ---- leaves = n byte strings of leaf_bytes each (caller's Vec<Vec<u8>>, flattened) ----------

__device__ __forceinline__ void merkle_leaf_from_bytes(uint32_t (&out)[8], const uint8_t *p, uint32_t len) {
    b2s::init(out);
    uint32_t m[16];
    uint32_t off = 0;
    const bool aligned = ((((size_t)p) | len) & 3) == 0;
    while (true) {
        const uint32_t rem = len - off, take = rem > 64 ? 64 : rem;
        if (aligned) {
            const uint32_t *q = (const uint32_t *)(p + off);
#pragma unroll
            for (int w = 0; w < 16; w++) m[w] = (4u * w < take) ? q[w] : 0u;
        } else {
#pragma unroll
            for (int w = 0; w < 16; w++) {
                uint32_t x = 0;
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (4u * w + b < take) x |= (uint32_t)p[off + 4 * w + b] << (8 * b);
                m[w] = x;
            }
        }
        off += take;
        const bool last = off == len;
        b2s::compress(out, m, off, last);
        if (last) break;
    }
}

template <int LV>
__global__ void __launch_bounds__(MERKLE_THREADS, 4) merkle_leaves_bytes_kernel(const __grid_constant__ MerkleBytesParams P) {
    __shared__ uint32_t sd_all[MERKLE_SMEM_WORDS(LV)];
    const size_t base = P.first + (((size_t)blockIdx.x * MERKLE_THREADS) << LV), end = P.first + P.count;
#pragma unroll 1
    for (int i = 0; i < (1 << LV); i++) {
        const uint32_t l = i * MERKLE_THREADS + threadIdx.x;
        if (base + l < end) {
            uint32_t h[8];
            merkle_leaf_from_bytes(h, P.leaves + (base + l) * (size_t)P.leaf_bytes, P.leaf_bytes);
            digest_store(P.nodes, base + l, h);
            sd_hand_over<LV>(sd_all, l, h);
        }
    }
    __syncthreads();
    const size_t first = base + ((size_t)threadIdx.x << LV);
    if (first < end) merkle_reduce_smem<LV>(sd_all + threadIdx.x, P.nodes, P.n, 0, first);
}

// ---- inner levels: each thread lifts 2^LV nodes of level `level` LV levels up -----------------
template <int LV>
__global__ void __launch_bounds__(MERKLE_THREADS, 4) merkle_nodes_kernel(uint4 *nodes, unsigned long long n, uint32_t level) {
    __shared__ uint32_t sd_all[MERKLE_SMEM_WORDS(LV)];
    const size_t base = ((size_t)blockIdx.x * MERKLE_THREADS) << LV;
    const size_t width = n >> level;
    const size_t off = merkle_level_off(n, level);
#pragma unroll
    for (int i = 0; i < (1 << LV); i++) {
        const uint32_t l = i * MERKLE_THREADS + threadIdx.x;
        if (base + l < width) {
            uint32_t d[8];
            digest_load(d, nodes, off + base + l);
            sd_hand_over<LV>(sd_all, l, d);
        }
    }
    __syncthreads();
    const size_t first = base + ((size_t)threadIdx.x << LV);
    if (first < width) merkle_reduce_smem<LV>(sd_all + threadIdx.x, nodes, n, level, first);
}

// ---- coset-major columns over g devices (ExtLeavesParams): thread k hashes the cpd leaves 8k + r0 .. 8k + r0 + cpd - 1 ----
// Lanes read consecutive k of one coset array: coalesced by construction.  The level-lv digest of a k is a peer store into
// the subtree array of the device that owns its node range (32 bytes per k over NVLink), or into `single`.
template <int LV>
__device__ __forceinline__ void merkle_ext_reduce(uint32_t *sd, const ExtLeavesParams &P, size_t k, uint4 *single, size_t n_single) {
    // digests 0 .. 2^LV - 1 of this thread sit in its smem slots; they are nodes (k cpd g + d cpd + i) of level 0 (cpd = 2^LV)
    const size_t first0 = (k * P.g + P.d) << LV;                  // level-0 index of this thread's first leaf
#pragma unroll 1
    for (int s = 0; s < LV; s++) {
        const int cnt = 1 << (LV - 1 - s);
#pragma unroll 1
        for (int i = 0; i < cnt; i++) {
            uint32_t m[16], o[8];
#pragma unroll
            for (int w = 0; w < 8; w++) {
                m[w] = sd[((2 * i) * 8 + w) * MERKLE_PITCH];
                m[w + 8] = sd[((2 * i + 1) * 8 + w) * MERKLE_PITCH];
            }
            b2s::hash64(o, m);
            sd_put(sd, i, o);
            const uint32_t l = s + 1;                             // level of the node just made
            if (single) {
                digest_store(single, merkle_level_off(n_single, l) + (first0 >> l) + i, o);
            } else if (l < (uint32_t)LV) {
                digest_store(P.low, ext_low_off(P.log_s, P.cpd, l) + k * (P.cpd >> l) + i, o);
            } else {                                              // l == LV: node k g + d of level lv, staged for its range owner
                digest_store(P.stage, k, o);
            }
        }
    }
}

// level-0 digest of leaf 8 k + r0 + i of this thread
template <int LV>
__device__ __forceinline__ void merkle_ext_store_leaf(const ExtLeavesParams &P, uint4 *single, size_t k, int i, const uint32_t (&h)[8]) {
    if (single) {
        digest_store(single, ((k * P.g + P.d) << LV) + i, h);
    } else if (LV > 0) {
        digest_store(P.low, k * P.cpd + i, h);
    } else {
        digest_store(P.stage, k, h);
    }
}

template <int LV>
__global__ void __launch_bounds__(MERKLE_THREADS, 4) merkle_leaves_ext_kernel(const __grid_constant__ ExtLeavesParams P) {
    __shared__ uint32_t sd_all[MERKLE_SMEM_WORDS(LV)];
    uint32_t *sd = sd_all + threadIdx.x;
    const size_t S = (size_t)1 << P.log_s, n_single = S << 3;
    const size_t k = (size_t)blockIdx.x * MERKLE_THREADS + threadIdx.x;
    if (k >= S) return;
    MerkleColsParams C;
    C.nc = P.nc;
#pragma unroll 1
    for (int i = 0; i < (1 << LV); i++) {
#pragma unroll
        for (int c = 0; c < 8; c++) C.cols[c] = P.cols[c] + 2 * ((size_t)i << P.log_s);     // coset r0 + i of every column
        uint32_t h[8];
        merkle_leaf_from_cols(h, C, k);
        merkle_ext_store_leaf<LV>(P, P.single, k, i, h);
        sd_put(sd, i, h);
    }
    merkle_ext_reduce<LV>(sd, P, k, P.single, n_single);
}

// FRI layer over coset-major values: fold (fri.rs:141-164) + leaf hashing of the folded column.  The column is written
// coset-major next to the data (col_local: the next layer stays sharded) and / or in natural order on the primary device
// (col: the next layer runs there alone); its tree goes to `nodes` (standard layout on the primary) or to the shards.
template <int LV>
__global__ void __launch_bounds__(MERKLE_THREADS, 4) merkle_leaves_fold_ext_kernel(const __grid_constant__ FriFoldParams F, uint4 *nodes) {
    __shared__ uint32_t sd_all[MERKLE_SMEM_WORDS(LV)];
    uint32_t *sd = sd_all + threadIdx.x;
    const size_t S = (size_t)1 << F.log_s, S4 = S >> 2, q = F.n >> 2;      // q = 8 S4 leaves in the column tree
    const size_t k = (size_t)blockIdx.x * MERKLE_THREADS + threadIdx.x;
    if (k >= S4) return;
    fp sx;
#pragma unroll
    for (int w = 0; w < 8; w++) sx.l[w] = F.special_x[w];
    const unsigned long long nT = 1ull << F.tw_log_n;
    const fp iota_inv = fp_ldg_ro(F.tw, (nT - ((unsigned long long)q << F.tw_log_stride)) & (nT - 1));
    ExtLeavesParams E;
    E.log_s = F.log_s - 2;
    E.cpd = F.cpd; E.lv = F.lv; E.d = F.d; E.g = F.g;
    E.low = F.low;
    E.stage = F.stage;
#pragma unroll 1
    for (int i = 0; i < (1 << LV); i++) {
        const uint4 *v = F.vals + 2 * (((size_t)i << F.log_s) + k);
        const size_t row = ((k * F.g + F.d) << LV) + i;                    // = 8 k + r0 + i
        fp y0 = fp_from_u4(v[0], v[1]), y1 = fp_from_u4(v[2 * S4], v[2 * S4 + 1]), y2 = fp_from_u4(v[4 * S4], v[4 * S4 + 1]),
           y3 = fp_from_u4(v[6 * S4], v[6 * S4 + 1]);
        fp r = fri_fold_vals(F, row, y0, y1, y2, y3, sx, iota_inv);
        if (F.col) fp_stg(F.col, row, r);
        if (F.col_local) fp_stg(F.col_local, ((size_t)i << (F.log_s - 2)) + k, r);
        fp c = fp_from_mont(r);
        uint32_t h[8];
        b2s::hash32(h, c.l);
        merkle_ext_store_leaf<LV>(E, nodes, k, i, h);
        sd_put(sd, i, h);
    }
    merkle_ext_reduce<LV>(sd, E, k, nodes, q);
}

// Level 0 .. log2 g of a device's subtree from the staged digests of the g sources: thread k reads the g digests of step k
// (recv[d * per + k], coalesced over k for every d), stores them as nodes k g .. k g + g - 1 of level 0 (g contiguous digests
// per thread) and reduces them log2 g levels.  (First version: a separate interleave pass whose 32-byte writes of one 128-byte
// line came from threads far apart in the grid -- 2.5 ms for 2^23 digests; L2 evicted the partial lines.)
template <int LG>
__global__ void __launch_bounds__(MERKLE_THREADS, 4) merkle_gather_reduce_kernel(const uint4 *recv, uint4 *sub, unsigned long long per) {
    __shared__ uint32_t sd_all[MERKLE_SMEM_WORDS(LG)];
    uint32_t *sd = sd_all + threadIdx.x;
    const size_t k = (size_t)blockIdx.x * MERKLE_THREADS + threadIdx.x;
    if (k >= per) return;
#pragma unroll
    for (int d = 0; d < (1 << LG); d++) {
        uint32_t h[8];
        digest_load(h, recv, (size_t)d * per + k);
        digest_store(sub, (k << LG) + d, h);
        sd_put(sd, d, h);
    }
    merkle_reduce_smem<LG>(sd, sub, per << LG, 0, k << LG);
}
// fold alone on coset-major values (the layer after it runs on the primary device, which hashes the gathered column itself)
__global__ void __launch_bounds__(128) fri_fold_ext_kernel(const __grid_constant__ FriFoldParams F) {
    const size_t S = (size_t)1 << F.log_s, S4 = S >> 2, q = F.n >> 2;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)F.cpd * S4) return;
    const size_t i = t / S4, k = t - i * S4;
    fp sx;
#pragma unroll
    for (int w = 0; w < 8; w++) sx.l[w] = F.special_x[w];
    const unsigned long long nT = 1ull << F.tw_log_n;
    const fp iota_inv = fp_ldg_ro(F.tw, (nT - ((unsigned long long)q << F.tw_log_stride)) & (nT - 1));
    const uint4 *v = F.vals + 2 * ((i << F.log_s) + k);
    const size_t row = (k << 3) + (size_t)F.d * F.cpd + i;
    fp y0 = fp_from_u4(v[0], v[1]), y1 = fp_from_u4(v[2 * S4], v[2 * S4 + 1]), y2 = fp_from_u4(v[4 * S4], v[4 * S4 + 1]),
       y3 = fp_from_u4(v[6 * S4], v[6 * S4 + 1]);
    fp_stg(F.col_local, t, fri_fold_vals(F, row, y0, y1, y2, y3, sx, iota_inv));
}

// sibling digests of the levels held in the shards (levels 0 .. lv + log_s - 1), leaf level first:
// out[q * (lv + log_s) + l]; the top log2 g levels come from the host copy
__global__ void merkle_open_ext_kernel(const __grid_constant__ ExtOpenParams P, const unsigned long long *idx, uint32_t n_idx, uint4 *out) {
    const uint32_t depth = P.lv + P.log_s;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_idx * depth) return;
    const uint32_t q = t / depth, l = t % depth;
    const unsigned long long i = idx[q];
    const uint4 *src;
    size_t at;
    if (l < P.lv) {
        const uint32_t r = (uint32_t)(i & 7), d = r / P.cpd, rl = r % P.cpd;
        src = P.low[d];
        at = ext_low_off(P.log_s, P.cpd, l) + (i >> 3) * (P.cpd >> l) + ((rl >> l) ^ 1);
    } else {
        const uint32_t sl = l - P.lv;
        const unsigned long long m = (i >> l) ^ 1, per = (1ull << P.log_s) >> sl;      // nodes of this level per device
        src = P.sub[m / per];
        at = merkle_level_off((size_t)1 << P.log_s, sl) + (m % per);
    }
    out[2 * (size_t)t] = src[2 * at];
    out[2 * (size_t)t + 1] = src[2 * at + 1];
}
// leaf bytes: out[q * nc + c] = to_bytes_le(col_c[idx[q]]) read from the device that holds coset idx[q] & 7
__global__ void merkle_open_leaves_ext_kernel(const __grid_constant__ ExtOpenParams P, const unsigned long long *idx, uint32_t n_idx, uint4 *out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_idx * P.nc) return;
    const uint32_t q = t / P.nc, c = t % P.nc;
    const unsigned long long i = idx[q];
    const uint32_t r = (uint32_t)(i & 7), d = r / P.cpd, rl = r % P.cpd;
    fp v = fp_from_mont(fp_ldg(P.cols[d][c], ((size_t)rl << P.log_s) + (i >> 3)));
    fp_stg(out, t, v);
}
// one coset-major column -> natural order (canonical Montgomery), gathered through peer pointers: out[8 k + r] = col_d[rl][k]
__global__ void ext_to_natural_kernel(const __grid_constant__ ExtOpenParams P, uint4 *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)8 << P.log_s)) return;
    const uint32_t r = (uint32_t)(i & 7), d = r / P.cpd, rl = r % P.cpd;
    fp_stg(out, i, fp_canon(fp_ldg(P.cols[d][0], ((size_t)rl << P.log_s) + (i >> 3))));
}

// ---- openings: sibling digests leaf level first, root excluded (merkle_tree.rs:25-43) --------
// out[q * depth + l] = level-l node (idx[q] >> l) ^ 1
__global__ void merkle_open_kernel(const uint4 *nodes, unsigned long long n, uint32_t depth,
                                   const unsigned long long *idx, uint32_t n_idx, uint4 *out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_idx * depth) return;
    const uint32_t q = t / depth, l = t % depth;
    const size_t node = merkle_level_off(n, l) + ((idx[q] >> l) ^ 1);
    out[2 * (size_t)t] = nodes[2 * node];
    out[2 * (size_t)t + 1] = nodes[2 * node + 1];
}

// leaf bytes of column-backed trees for the opened indices: out[q] = concat_k to_bytes_le(col_k[idx[q]])
__global__ void merkle_open_leaves_cols_kernel(MerkleColsParams P, const unsigned long long *idx, uint32_t n_idx, uint4 *out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nc = P.nc;
    if (t >= n_idx * nc) return;
    const uint32_t q = t / nc, k = t % nc;
    size_t at = idx[q];
    if (P.coset_log_s) at = ((at & 7) << P.coset_log_s) | (at >> 3);
    fp v = fp_from_mont(fp_ldg(P.cols[k], at));
    fp_stg(out, t, v);
}
