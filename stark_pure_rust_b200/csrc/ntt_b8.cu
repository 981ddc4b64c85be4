#include "ntt_inst.cuh"
cudaError_t ntt_set_attrs_b8() { return ntt_set_attr_one<8>(); }
int ntt_launch_pass_b8(cudaStream_t s, const NttPassParams &P) { return ntt_launch_one<8>(s, P); }
