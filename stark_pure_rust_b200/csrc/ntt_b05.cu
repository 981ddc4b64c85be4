#include "ntt_inst.cuh"
cudaError_t ntt_set_attrs_b05() {
    cudaError_t e = ntt_set_attr_one<0>();
    if (e == cudaSuccess) e = ntt_set_attr_one<1>();
    if (e == cudaSuccess) e = ntt_set_attr_one<2>();
    if (e == cudaSuccess) e = ntt_set_attr_one<3>();
    if (e == cudaSuccess) e = ntt_set_attr_one<4>();
    if (e == cudaSuccess) e = ntt_set_attr_one<5>();
    return e;
}
int ntt_launch_pass_b05(cudaStream_t s, uint32_t bits, const NttPassParams &P) {
    switch (bits) {
        case 0: return ntt_launch_one<0>(s, P);
        case 1: return ntt_launch_one<1>(s, P);
        case 2: return ntt_launch_one<2>(s, P);
        case 3: return ntt_launch_one<3>(s, P);
        case 4: return ntt_launch_one<4>(s, P);
        case 5: return ntt_launch_one<5>(s, P);
    }
    return -1;
}
