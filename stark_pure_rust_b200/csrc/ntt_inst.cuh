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
    if constexpr (B >= 6) {          // the widths a transform of 2^20 points or more is made of (ntt_multi)
        cudaError_t e = cudaFuncSetAttribute(ntt_pass_kernel<B, NTT_LOG_TILE_FOR(B), true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes<B>());
        if (e != cudaSuccess) return e;
    }
    return cudaFuncSetAttribute(ntt_pass_kernel<B, NTT_LOG_TILE_FOR(B)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes<B>());
}
template <int B>
static int ntt_launch_one(cudaStream_t stream, const NttPassParams &P) {
    constexpr int TILE = 1 << NTT_LOG_TILE_FOR(B), CC = TILE >> B;
    unsigned long long grid = (P.n_cols_total + CC - 1) / CC;
    if constexpr (B >= 6) {
        if (P.dist_log_g) {          // this device's share of the tiles
            ntt_pass_kernel<B, NTT_LOG_TILE_FOR(B), true><<<(unsigned)(grid >> P.dist_log_g), TILE / 8, ntt_smem_bytes<B>(), stream>>>(P);
            return 1;
        }
    }
    ntt_pass_kernel<B, NTT_LOG_TILE_FOR(B)><<<(unsigned)grid, TILE / 8, ntt_smem_bytes<B>(), stream>>>(P);
    return 1;
}
