// kernels.h -- host-callable launchers of the sm_100a kernels; one translation unit per kernel family
// (ntt_b*.cu, merkle.cu, fri.cu) so that the library builds in parallel.  Every launcher enqueues on
// `stream` and returns the number of kernels it launched.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct NttPassParams;
struct MerkleColsParams;
struct MerkleBytesParams;
struct FriFoldParams;
struct ExtLeavesParams;
struct ExtOpenParams;
struct fp;

// ntt_b*.cu
cudaError_t ntt_set_attrs();
int ntt_launch_pass(cudaStream_t stream, uint32_t bits, const NttPassParams &P);
int lde_launch_coset0(cudaStream_t stream, const uint4 *in, unsigned long long col_len, unsigned long long in_stride, uint4 *out,
                      unsigned long long out_stride, unsigned long long s_len, uint32_t log_ext, unsigned long long n_cols);
// merkle.cu
int merkle_launch_leaves_cols(cudaStream_t stream, uint32_t lv, const MerkleColsParams &P);
int merkle_launch_leaves_fold(cudaStream_t stream, uint32_t lv, const FriFoldParams &F, uint4 *nodes);
int merkle_launch_leaves_bytes(cudaStream_t stream, uint32_t lv, const MerkleBytesParams &P);
int merkle_launch_nodes(cudaStream_t stream, uint32_t lv, uint4 *nodes, unsigned long long n, uint32_t level);
int merkle_launch_open(cudaStream_t stream, const uint4 *nodes, unsigned long long n, uint32_t depth,
                       const unsigned long long *idx, uint32_t n_idx, uint4 *out);
int merkle_launch_open_leaves_cols(cudaStream_t stream, const MerkleColsParams &P, const unsigned long long *idx,
                                   uint32_t n_idx, uint4 *out);
int merkle_launch_leaves_ext(cudaStream_t stream, const ExtLeavesParams &P);
int merkle_launch_leaves_fold_ext(cudaStream_t stream, const FriFoldParams &F, uint4 *nodes);
int merkle_launch_open_ext(cudaStream_t stream, const ExtOpenParams &P, const unsigned long long *idx, uint32_t n_idx, uint4 *nodes_out,
                           uint4 *leaves_out);
int merkle_launch_gather_reduce(cudaStream_t stream, const uint4 *recv, uint4 *sub, unsigned long long per, uint32_t g);
int fri_launch_fold_ext(cudaStream_t stream, const FriFoldParams &F);
int ext_launch_to_natural(cudaStream_t stream, const ExtOpenParams &P, uint4 *out);
int merkle_launch_gather_bytes(cudaStream_t stream, const uint8_t *leaves, size_t leaf_bytes,
                               const unsigned long long *idx, uint32_t n_idx, uint8_t *out);
// fri.cu
int fri_launch_fold(cudaStream_t stream, const FriFoldParams &P);
int powers_launch_seed(cudaStream_t stream, uint4 *T, unsigned long long count, const fp &w);
int powers_launch_double(cudaStream_t stream, uint4 *T, unsigned long long cur, unsigned long long n_total, const fp &wcur);
int fp_launch_to_bytes(cudaStream_t stream, const uint4 *in, uint4 *out, unsigned long long n);
unsigned long long batch_inverse_scratch_elems(unsigned long long n);
int batch_inverse_launch(cudaStream_t stream, uint4 *vals, uint4 *scratch, unsigned long long n);
int fp_launch_vec_op(cudaStream_t stream, int op, const uint4 *a, const uint4 *b, uint4 *out, unsigned long long n);
double pipe_probe_launch(cudaStream_t stream, int mode, uint4 *out, unsigned blocks, uint32_t iters, const fp &seed);
int twiddle_mul_launch(cudaStream_t stream, uint4 *vals, unsigned long long rows, unsigned long long cols, unsigned long long row0,
                       const uint4 *tw, uint32_t tw_log_n, uint32_t tw_log_stride, uint32_t log_n, int inverse);
