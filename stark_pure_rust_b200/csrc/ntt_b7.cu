#include "ntt_inst.cuh"
cudaError_t ntt_set_attrs_b7() { return ntt_set_attr_one<7>(); }
int ntt_launch_pass_b7(cudaStream_t s, const NttPassParams &P) { return ntt_launch_one<7>(s, P); }
