// ntt.cu -- dispatch over the pass widths compiled in ntt_b*.cu
#include "kernels.h"
#include "ntt.cuh"
cudaError_t ntt_set_attrs_b05();
cudaError_t ntt_set_attrs_b6();
cudaError_t ntt_set_attrs_b7();
cudaError_t ntt_set_attrs_b8();
int ntt_launch_pass_b05(cudaStream_t s, uint32_t bits, const NttPassParams &P);
int ntt_launch_pass_b6(cudaStream_t s, const NttPassParams &P);
int ntt_launch_pass_b7(cudaStream_t s, const NttPassParams &P);
int ntt_launch_pass_b8(cudaStream_t s, const NttPassParams &P);

cudaError_t ntt_set_attrs() {
    cudaError_t e = ntt_set_attrs_b05();
    if (e == cudaSuccess) e = ntt_set_attrs_b6();
    if (e == cudaSuccess) e = ntt_set_attrs_b7();
    if (e == cudaSuccess) e = ntt_set_attrs_b8();
    return e;
}
int ntt_launch_pass(cudaStream_t s, uint32_t bits, const NttPassParams &P) {
    if (bits <= 5) return ntt_launch_pass_b05(s, bits, P);
    if (bits == 6) return ntt_launch_pass_b6(s, P);
    if (bits == 7) return ntt_launch_pass_b7(s, P);
    if (bits == 8) return ntt_launch_pass_b8(s, P);
    return -1;
}

// coset 0 of a low-degree extension is the input itself: out[col][k << log_ext] = in[col][k] (zero beyond col_len)
__global__ void lde_coset0_kernel(const uint4 *in, unsigned long long col_len, unsigned long long in_stride, uint4 *out,
                                  unsigned long long out_stride, unsigned long long s_len, uint32_t log_ext, unsigned long long total) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const unsigned long long col = t / s_len, k = t % s_len;
    fp v = fp_zero();
    if (k < col_len) v = fp_canon(fp_ldg(in, col * in_stride + k));
    fp_stg(out, col * out_stride + (k << log_ext), v);
}

int lde_launch_coset0(cudaStream_t s, const uint4 *in, unsigned long long col_len, unsigned long long in_stride, uint4 *out,
                      unsigned long long out_stride, unsigned long long s_len, uint32_t log_ext, unsigned long long n_cols) {
    const unsigned long long total = n_cols * s_len;
    lde_coset0_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, col_len, in_stride, out, out_stride, s_len, log_ext, total);
    return 1;
}
