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


// Per-thread digest scratch in shared memory: digest i, word w of thread t lives at sd[(i * 8 + w) * 128 + t]
// (word-interleaved by thread: every lane stays in its own bank).  Keeping the 2^LV digests there instead of
// in a register array lets the loops below stay rolled, so the kernel holds ONE copy of the leaf compression
// and ONE of the node compression (~40 KB of SASS) instead of 2^LV + 2^LV - 1 inlined copies (290 KB, which
// missed the instruction cache on every iteration: ncu showed 3.3 "no instruction" stall cycles per issue).
#define MERKLE_THREADS 128
__device__ __forceinline__ void sd_put(uint32_t *sd, int i, const uint32_t (&d)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++) sd[(i * 8 + k) * MERKLE_THREADS] = d[k];
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
                m[k] = sd[((2 * i) * 8 + k) * MERKLE_THREADS];
                m[k + 8] = sd[((2 * i + 1) * 8 + k) * MERKLE_THREADS];
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
    __shared__ uint32_t sd_all[(1 << LV) * 8 * MERKLE_THREADS];
    uint32_t *sd = sd_all + threadIdx.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t first = t << LV;
    if (first >= P.n) return;
#pragma unroll 1
    for (int i = 0; i < (1 << LV); i++) {
        uint32_t h[8];
        merkle_leaf_from_cols(h, P, first + i);
        digest_store(P.nodes, first + i, h);
        sd_put(sd, i, h);
    }
    merkle_reduce_smem<LV>(sd, P.nodes, P.n, 0, first);
}

// ---- FRI: fold fused with the next layer's leaf hashing (fri.rs:141-172) ----------------------------------------
// leaf i of the column tree = to_bytes_le(column[i]) where column[i] is folded from four values of the current layer;
// the column is also written out (it is the next layer's input and the opened leaves are re-read from it).
template <int LV>
__global__ void __launch_bounds__(MERKLE_THREADS, 4) merkle_leaves_fold_kernel(const __grid_constant__ FriFoldParams F, uint4 *nodes) {
    __shared__ uint32_t sd_all[(1 << LV) * 8 * MERKLE_THREADS];
    uint32_t *sd = sd_all + threadIdx.x;
    const size_t q = F.n >> 2;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t first = t << LV;
    if (first >= q) return;
    fp sx;
#pragma unroll
    for (int k = 0; k < 8; k++) sx.l[k] = F.special_x[k];
    const unsigned long long nT = 1ull << F.tw_log_n;
    const fp iota_inv = fp_ldg_ro(F.tw, (nT - ((unsigned long long)q << F.tw_log_stride)) & (nT - 1));
#pragma unroll 1
    for (int i = 0; i < (1 << LV); i++) {
        fp r = fri_fold_row(F, first + i, sx, iota_inv);
        fp_stg(F.col, first + i, r);
        fp c = fp_from_mont(r);
        uint32_t h[8];
        b2s::hash32(h, c.l);
        digest_store(nodes, first + i, h);
        sd_put(sd, i, h);
    }
    merkle_reduce_smem<LV>(sd, nodes, q, 0, first);
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
    __shared__ uint32_t sd_all[(1 << LV) * 8 * MERKLE_THREADS];
    uint32_t *sd = sd_all + threadIdx.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t first = t << LV;
    if (first >= P.n) return;
#pragma unroll 1
    for (int i = 0; i < (1 << LV); i++) {
        uint32_t h[8];
        merkle_leaf_from_bytes(h, P.leaves + (first + i) * (size_t)P.leaf_bytes, P.leaf_bytes);
        digest_store(P.nodes, first + i, h);
        sd_put(sd, i, h);
    }
    merkle_reduce_smem<LV>(sd, P.nodes, P.n, 0, first);
}

// ---- inner levels: each thread lifts 2^LV nodes of level `level` LV levels up -----------------
template <int LV>
__global__ void __launch_bounds__(MERKLE_THREADS, 4) merkle_nodes_kernel(uint4 *nodes, unsigned long long n, uint32_t level) {
    __shared__ uint32_t sd_all[(1 << LV) * 8 * MERKLE_THREADS];
    uint32_t *sd = sd_all + threadIdx.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t first = t << LV;
    const size_t width = n >> level;
    if (first >= width) return;
    const size_t off = merkle_level_off(n, level);
#pragma unroll
    for (int i = 0; i < (1 << LV); i++) {
        uint32_t d[8];
        digest_load(d, nodes, off + first + i);
        sd_put(sd, i, d);
    }
    merkle_reduce_smem<LV>(sd, nodes, n, level, first);
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
    fp v = fp_from_mont(fp_ldg(P.cols[k], idx[q]));
    fp_stg(out, t, v);
}
