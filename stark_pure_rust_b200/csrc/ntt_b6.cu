#include "ntt_inst.cuh"
cudaError_t ntt_set_attrs_b6() { return ntt_set_attr_one<6>(); }
int ntt_launch_pass_b6(cudaStream_t s, const NttPassParams &P) { return ntt_launch_one<6>(s, P); }
