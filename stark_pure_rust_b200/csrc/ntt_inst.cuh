// ntt_inst.cuh -- included by ntt_b*.cu: each translation unit instantiates the pass kernel for the
// widths B in [NTT_B_LO, NTT_B_HI] (separate files only to build them in parallel).
#include "kernels.h"
#include "ntt.cuh"

template <int B>
static size_t ntt_smem_bytes() {
    constexpr int R = 1 << B, CC = (1 << NTT_LOG_TILE_FOR(B)) >> B, PITCH = CC + 1;
    return (size_t)(2 * R * PITCH + 2 * (R / 2 > 0 ? R / 2 : 1)) * sizeof(uint4);
}
template <int B>
static cudaError_t ntt_set_attr_one() {
    return cudaFuncSetAttribute(ntt_pass_kernel<B, NTT_LOG_TILE_FOR(B)>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)ntt_smem_bytes<B>());
}
template <int B>
static int ntt_launch_one(cudaStream_t stream, const NttPassParams &P) {
    constexpr int TILE = 1 << NTT_LOG_TILE_FOR(B), CC = TILE >> B;
    unsigned long long grid = (P.n_cols_total + CC - 1) / CC;
    if (P.cluster) {          // the coset_m1 consecutive CTAs of one (column, tile) form a cluster
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid, 1, 1);
        cfg.blockDim = dim3(TILE / 8, 1, 1);
        cfg.dynamicSmemBytes = ntt_smem_bytes<B>();
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = P.coset_m1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, ntt_pass_kernel<B, NTT_LOG_TILE_FOR(B)>, P);
        return 1;
    }
    ntt_pass_kernel<B, NTT_LOG_TILE_FOR(B)><<<(unsigned)grid, TILE / 8, ntt_smem_bytes<B>(), stream>>>(P);
    return 1;
}
