// fri_fold.cuh -- one row of the FRI 4-to-1 fold (closed form of fri.rs:141-164, see fri.cuh); shared by the plain fold
// kernel (fri.cuh) and the fold fused with the next layer's leaf hashing (merkle.cuh).
#pragma once
#include "fp.cuh"
#include "params.h"

// y0..y3 = values[i + j q], j < 4 (q = n / 4): the degree-<4 interpolant through (x iota^j, y_j) evaluated at special_x
__device__ __forceinline__ fp fri_fold_vals(const FriFoldParams &P, size_t i, const fp &y0, const fp &y1, const fp &y2, const fp &y3,
                                            const fp &sx, const fp &iota_inv) {
    const unsigned long long nT = 1ull << P.tw_log_n;
    fp xinv = fp_ldg_ro(P.tw, (nT - ((unsigned long long)i << P.tw_log_stride)) & (nT - 1));
    fp z = fp_mul(sx, xinv);
    fp a = fp_add(y0, y2), b = fp_sub(y0, y2), c = fp_add(y1, y3);
    fp d = fp_mul(fp_sub_lazy(y1, y3), iota_inv);
    fp s0 = fp_add(a, c), s2 = fp_sub(a, c), s1 = fp_add(b, d), s3 = fp_sub(b, d);
    fp r = fp_add(fp_mul(s3, z), s2);
    r = fp_add(fp_mul(r, z), s1);
    r = fp_add(fp_mul(r, z), s0);
    r = fp_canon(r);
    r = fp_half(r);                 // [0,1.5p)
    r = fp_half(r);                 // [0,1.25p)
    return fp_canon(r);
}

__device__ __forceinline__ fp fri_fold_row(const FriFoldParams &P, size_t i, const fp &sx, const fp &iota_inv) {
    const size_t q = P.n >> 2;
    fp y0 = fp_ldg(P.vals, i), y1 = fp_ldg(P.vals, i + q), y2 = fp_ldg(P.vals, i + 2 * q), y3 = fp_ldg(P.vals, i + 3 * q);
    return fri_fold_vals(P, i, y0, y1, y2, y3, sx, iota_inv);
}
