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
