// merkle.cu -- launchers of the Merkle kernels (merkle.cuh)
#include "kernels.h"
#include "merkle.cuh"

static_assert(MERKLE_THREADS == 128, "launchers assume 128-thread CTAs");
static unsigned blocks128(size_t threads) { return (unsigned)((threads + 127) / 128); }

int merkle_launch_leaves_cols(cudaStream_t s, uint32_t lv, const MerkleColsParams &P) {
    const unsigned b = blocks128(P.n >> lv);
    switch (lv) {
        case 0: merkle_leaves_cols_kernel<0><<<b, 128, 0, s>>>(P); break;
        case 1: merkle_leaves_cols_kernel<1><<<b, 128, 0, s>>>(P); break;
        case 2: merkle_leaves_cols_kernel<2><<<b, 128, 0, s>>>(P); break;
        default: merkle_leaves_cols_kernel<3><<<b, 128, 0, s>>>(P); break;
    }
    return 1;
}
int merkle_launch_leaves_fold(cudaStream_t s, uint32_t lv, const FriFoldParams &F, uint4 *nodes) {
    const unsigned b = blocks128((F.n >> 2) >> lv);
    switch (lv) {
        case 0: merkle_leaves_fold_kernel<0><<<b, 128, 0, s>>>(F, nodes); break;
        case 1: merkle_leaves_fold_kernel<1><<<b, 128, 0, s>>>(F, nodes); break;
        case 2: merkle_leaves_fold_kernel<2><<<b, 128, 0, s>>>(F, nodes); break;
        default: merkle_leaves_fold_kernel<3><<<b, 128, 0, s>>>(F, nodes); break;
    }
    return 1;
}
int merkle_launch_leaves_bytes(cudaStream_t s, uint32_t lv, const MerkleBytesParams &P) {
    const unsigned b = blocks128((P.count + ((size_t)1 << lv) - 1) >> lv);
    switch (lv) {
        case 0: merkle_leaves_bytes_kernel<0><<<b, 128, 0, s>>>(P); break;
        case 1: merkle_leaves_bytes_kernel<1><<<b, 128, 0, s>>>(P); break;
        case 2: merkle_leaves_bytes_kernel<2><<<b, 128, 0, s>>>(P); break;
        default: merkle_leaves_bytes_kernel<3><<<b, 128, 0, s>>>(P); break;
    }
    return 1;
}
int merkle_launch_nodes(cudaStream_t s, uint32_t lv, uint4 *nodes, unsigned long long n, uint32_t level) {
    const unsigned b = blocks128((n >> level) >> lv);
    switch (lv) {
        case 1: merkle_nodes_kernel<1><<<b, 128, 0, s>>>(nodes, n, level); break;
        case 2: merkle_nodes_kernel<2><<<b, 128, 0, s>>>(nodes, n, level); break;
        default: merkle_nodes_kernel<3><<<b, 128, 0, s>>>(nodes, n, level); break;
    }
    return 1;
}
int merkle_launch_open(cudaStream_t s, const uint4 *nodes, unsigned long long n, uint32_t depth,
                       const unsigned long long *idx, uint32_t n_idx, uint4 *out) {
    merkle_open_kernel<<<blocks128((size_t)n_idx * depth), 128, 0, s>>>(nodes, n, depth, idx, n_idx, out);
    return 1;
}
int merkle_launch_open_leaves_cols(cudaStream_t s, const MerkleColsParams &P, const unsigned long long *idx,
                                   uint32_t n_idx, uint4 *out) {
    merkle_open_leaves_cols_kernel<<<blocks128((size_t)n_idx * P.nc), 128, 0, s>>>(P, idx, n_idx, out);
    return 1;
}

int merkle_launch_leaves_ext(cudaStream_t s, const ExtLeavesParams &P) {
    const unsigned b = blocks128((size_t)1 << P.log_s);
    switch (P.lv) {
        case 0: merkle_leaves_ext_kernel<0><<<b, 128, 0, s>>>(P); break;
        case 1: merkle_leaves_ext_kernel<1><<<b, 128, 0, s>>>(P); break;
        case 2: merkle_leaves_ext_kernel<2><<<b, 128, 0, s>>>(P); break;
        default: merkle_leaves_ext_kernel<3><<<b, 128, 0, s>>>(P); break;
    }
    return 1;
}
int merkle_launch_leaves_fold_ext(cudaStream_t s, const FriFoldParams &F, uint4 *nodes) {
    const unsigned b = blocks128(((size_t)1 << F.log_s) >> 2);
    switch (F.lv) {
        case 0: merkle_leaves_fold_ext_kernel<0><<<b, 128, 0, s>>>(F, nodes); break;
        case 1: merkle_leaves_fold_ext_kernel<1><<<b, 128, 0, s>>>(F, nodes); break;
        case 2: merkle_leaves_fold_ext_kernel<2><<<b, 128, 0, s>>>(F, nodes); break;
        default: merkle_leaves_fold_ext_kernel<3><<<b, 128, 0, s>>>(F, nodes); break;
    }
    return 1;
}
int merkle_launch_open_ext(cudaStream_t s, const ExtOpenParams &P, const unsigned long long *idx, uint32_t n_idx, uint4 *nodes_out,
                           uint4 *leaves_out) {
    int n = 0;
    if (nodes_out && P.lv + P.log_s) {
        merkle_open_ext_kernel<<<blocks128((size_t)n_idx * (P.lv + P.log_s)), 128, 0, s>>>(P, idx, n_idx, nodes_out);
        n++;
    }
    if (leaves_out) {
        merkle_open_leaves_ext_kernel<<<blocks128((size_t)n_idx * P.nc), 128, 0, s>>>(P, idx, n_idx, leaves_out);
        n++;
    }
    return n;
}
// levels 0 .. log2 g of a subtree from the digests staged by source device; g = 2, 4 or 8
int merkle_launch_gather_reduce(cudaStream_t s, const uint4 *recv, uint4 *sub, unsigned long long per, uint32_t g) {
    const unsigned b = blocks128(per);
    switch (g) {
        case 2: merkle_gather_reduce_kernel<1><<<b, 128, 0, s>>>(recv, sub, per); break;
        case 4: merkle_gather_reduce_kernel<2><<<b, 128, 0, s>>>(recv, sub, per); break;
        default: merkle_gather_reduce_kernel<3><<<b, 128, 0, s>>>(recv, sub, per); break;
    }
    return 1;
}
int fri_launch_fold_ext(cudaStream_t s, const FriFoldParams &F) {
    const size_t total = (size_t)F.cpd << (F.log_s - 2);
    fri_fold_ext_kernel<<<blocks128(total), 128, 0, s>>>(F);
    return 1;
}
int ext_launch_to_natural(cudaStream_t s, const ExtOpenParams &P, uint4 *out) {
    ext_to_natural_kernel<<<(unsigned)((((size_t)8 << P.log_s) + 255) / 256), 256, 0, s>>>(P, out);
    return 1;
}

__global__ void gather_leaf_bytes_kernel(const uint8_t *leaves, size_t leaf_bytes, const unsigned long long *idx, uint32_t n_idx,
                                         uint8_t *out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_idx * leaf_bytes) return;
    const size_t q = t / leaf_bytes, b = t % leaf_bytes;
    out[t] = leaves[idx[q] * leaf_bytes + b];
}
int merkle_launch_gather_bytes(cudaStream_t s, const uint8_t *leaves, size_t leaf_bytes, const unsigned long long *idx,
                               uint32_t n_idx, uint8_t *out) {
    const size_t tot = (size_t)n_idx * leaf_bytes;
    gather_leaf_bytes_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(leaves, leaf_bytes, idx, n_idx, out);
    return 1;
}
