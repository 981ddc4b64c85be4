// ntt.cuh -- multi-pass radix-2^B NTT over BN254 Fr for sm_100a.
//
// Replaces fri/src/fft.rs: serial_fft (:150-193), parallel_fft (:195-251), inv_*_fft (:284-309),
// best_fft / inv_best_fft (:327-379).  Same result: out[k] = sum_j v[j] * w^(jk), natural order
// in and out, zero padding to 2^log_n, inverse = forward with w^-1 then * n^-1.
//
// Decomposition (index model: tools/ntt_plan_model.py): log_n = b_1 + ... + b_m, decimation in
// frequency from the top bits down.  Pass p works on sub-transforms of 2^{b_p} points taken with
// stride `inner` (the not-yet-transformed low bits), entirely inside one CTA:
//     global -> shared tile (2^b rows x CC contiguous columns, 32-B elements split in two uint4
//     planes, row pitch CC+1 so that row-wise and column-wise warp accesses are conflict free)
//     -> register radix-8 rounds (3 butterfly stages per shared-memory round trip)
//     -> multiply by the inter-pass twiddle w^(outer*k*c) -> global (same tile, natural k order).
// The last pass (inner == 1) gathers CC sub-transforms whose OUTPUT positions are adjacent and
// writes the digit-reversed result, i.e. the transpose happens through shared memory and both
// the loads (2^b x 32 B rows) and the stores (CC x 32 B rows) are coalesced.  Bit reversal never
// touches HBM on its own.  All twiddles come from one table T[e] = w^e (the `xs` array the
// prover needs anyway, prove.rs:84); the inverse transform indexes it backwards.
//
// Bound: 32-bit integer pipe (IMAD carry chains); HBM traffic is 64 B per element per pass.
#pragma once
#include "fp.cuh"
#include "params.h"

// elements in flight per thread in the tile load / store loops (8 elements per thread and tile).  Measured on B200 (LDE
// 2^21 -> 2^24 x 10, ntt_pass ms per step): load/store unroll 2/2 33.27, 4/2 33.77, 8/1 34.16, 8/2 34.69, 4/4 35.31, 8/4 36.75 --
// 2/1 32.88, 1/2 32.89, 1/1 32.72 -- less unrolling wins (the kernel is ~110 KB of SASS: instruction supply matters more than
// loads in flight).
#ifndef NTT_LOAD_UNROLL
#define NTT_LOAD_UNROLL 1
#endif
#ifndef NTT_STORE_UNROLL
#define NTT_STORE_UNROLL 1
#endif


// table index of w^e (e < n) for the forward / inverse transform
__device__ __forceinline__ size_t ntt_tw_index(const NttPassParams &P, unsigned long long e) {
    unsigned long long nT = 1ull << P.tw_log_n;
    unsigned long long i = e << P.tw_log_stride;
    return P.inverse ? ((nT - i) & (nT - 1)) : i;
}

// one radix-2^Q register round over rows {base + e*2^S}; DIF, highest bit first.
// x[e] in [0,2p).  wlo/whi: shared planes of W[i] = w^(i * n / 2^B), i < 2^(B-1).
template <int B, int S, int Q>
__device__ __forceinline__ void ntt_round_regs(fp (&x)[1 << Q], uint32_t low, const uint4 *wlo, const uint4 *whi) {
#pragma unroll
    for (int tb = Q - 1; tb >= 0; tb--) {
        const int t = S + tb;               // bit of the row index this stage pairs on
#pragma unroll
        for (int e = 0; e < (1 << Q); e++) {
            if (e & (1 << tb)) continue;
            const int e_low = e & ((1 << tb) - 1);
            fp a = x[e], c = x[e | (1 << tb)];
            x[e] = fp_add(a, c);
            if (S == 0 && e_low == 0) {
                x[e | (1 << tb)] = fp_sub(a, c);            // twiddle w^0
            } else {
                uint32_t expo = (low | ((uint32_t)e_low << S)) << (B - 1 - t);
                fp w = fp_from_u4(wlo[expo], whi[expo]);
                x[e | (1 << tb)] = fp_mul(fp_sub_lazy(a, c), w);
            }
        }
    }
}

// (Tried: the S > 0 radix-8 round as one rolled loop over its three stages with a constant-geometry register renumbering -- 4
// inlined products instead of 12, 64 KB of SASS instead of 91 KB -- and the product as a real function call (47 KB).  Both
// slower on B200: 34.9 ms of ntt_pass per step against 32.7: the moves / spills / call overhead cost more than the smaller
// instruction footprint gains.)
template <int B, int S, int Q, int LOG_TILE>
__device__ __forceinline__ void ntt_round(uint4 *slo, uint4 *shi, const uint4 *wlo, const uint4 *whi) {
    constexpr int TILE = 1 << LOG_TILE, NT = TILE / 8, CC = TILE >> B, PITCH = CC + 1;
    constexpr int GROUPS = TILE >> Q;
#pragma unroll 1
    for (int g = threadIdx.x; g < GROUPS; g += NT) {
        const int j = g % CC;
        const uint32_t rb = g / CC;                          // row index with the Q round bits removed
        const uint32_t low = rb & ((1u << S) - 1), high = rb >> S;
        const uint32_t base = (high << (S + Q)) | low;
        fp x[1 << Q];
#pragma unroll
        for (int e = 0; e < (1 << Q); e++) {
            const int idx = (base + (e << S)) * PITCH + j;
            x[e] = fp_from_u4(slo[idx], shi[idx]);
        }
        ntt_round_regs<B, S, Q>(x, low, wlo, whi);
#pragma unroll
        for (int e = 0; e < (1 << Q); e++) {
            const int idx = (base + (e << S)) * PITCH + j;
            slo[idx] = fp_lo(x[e]);
            shi[idx] = fp_hi(x[e]);
        }
    }
}

// all rounds of a 2^B sub-transform: first round takes B mod 3 bits (if any), the rest 3 each (radix 8 in registers between
// shared-memory round trips).  Measured alternatives on B200 (LDE 2^21 -> 2^24 x 10, ms of ntt_pass per step): radix 8 33.8,
// radix 4 34.0 (96 registers, 5 CTAs per SM), radix 2 35.9, radix 2 rolled into one loop 36.5 (64 registers, 6 CTAs per SM):
// fewer shared-memory round trips beat more resident warps.
template <int B, int HI, int LOG_TILE>
__device__ __forceinline__ void ntt_rounds(uint4 *slo, uint4 *shi, const uint4 *wlo, const uint4 *whi) {
    if constexpr (HI > 0) {
        constexpr int Q = (HI % 3) ? (HI % 3) : 3;
        ntt_round<B, HI - Q, Q, LOG_TILE>(slo, shi, wlo, whi);
        __syncthreads();
        ntt_rounds<B, HI - Q, LOG_TILE>(slo, shi, wlo, whi);
    }
}

// last pass: column g (output position d = g) -> block index o of the input, i.e. the inverse of
// digitrev(o) = k_1 + n_1 k_2 + ... for o = (k_1, ..., k_{m-1}) with k_1 most significant.
__device__ __forceinline__ unsigned long long ntt_digitrev_inv(const NttPassParams &P, unsigned long long d) {
    unsigned long long o = 0;
    for (uint32_t i = 0; i < P.n_prev; i++) {
        uint32_t b = P.prev_bits[i];
        o = (o << b) | (d & ((1ull << b) - 1));
        d >>= b;
    }
    return o;
}

// DIST: one transform over several devices (see NttPassParams::src_tab): the tile index comes from the device's share of the
// tiles, loads and stores go through the slab tables -- remote reads in the first pass (the exchange "rows -> columns" of a
// four-step transform), remote writes in the first and in the last pass (the exchanges "-> k1 slabs" and "-> natural order"),
// always in runs of CC x 32 contiguous bytes.  Plain single-vector transforms only (no batch, no cosets, no zero padding).
template <int B, int LOG_TILE, bool DIST = false>
__global__ void __launch_bounds__((1 << LOG_TILE) / 8, LOG_TILE >= 11 ? 2 : 4) ntt_pass_kernel(const __grid_constant__ NttPassParams P) {
    constexpr int TILE = 1 << LOG_TILE, NT = TILE / 8, R = 1 << B, CC = TILE >> B, PITCH = CC + 1;
    extern __shared__ uint4 smem[];
    uint4 *slo = smem, *shi = smem + R * PITCH;
    uint4 *wlo = shi + R * PITCH, *whi = wlo + (R / 2 > 0 ? R / 2 : 1);

    const uint32_t log_cpp = P.log_n - B;                     // log2(columns per polynomial)
    unsigned long long blk_col0 = (unsigned long long)blockIdx.x * CC;
    if constexpr (DIST) {
        const unsigned long long i = blockIdx.x, lo = P.dist_blk_lo;
        blk_col0 = ((((i >> lo) << P.dist_log_g) | P.dist_dev) << lo | (i & ((1ull << lo) - 1))) * CC;
    } else if (P.n_polys) {   // tile-major order (host guarantees CC divides the columns of one polynomial)
        const unsigned long long poly = blockIdx.x % P.n_polys, tile = blockIdx.x / P.n_polys;
        blk_col0 = (poly << log_cpp) + tile * CC;
    }

    // sub-transform twiddles W[i] = w^(i * n/R)
    for (int i = threadIdx.x; i < R / 2; i += NT) {
        size_t ti = ntt_tw_index(P, (unsigned long long)i << (P.log_n - B));
        wlo[i] = __ldg(P.tw + 2 * ti);
        whi[i] = __ldg(P.tw + 2 * ti + 1);
    }

    // ---- load tile ----
    constexpr int LOAD_UNROLL = NTT_LOAD_UNROLL, STORE_UNROLL = NTT_STORE_UNROLL;
#pragma unroll LOAD_UNROLL
    for (int idx = threadIdx.x; idx < TILE; idx += NT) {
        int r, j;
        if (!P.last) { j = idx % CC; r = idx / CC; } else { r = idx % R; j = idx / R; }
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        {
            const unsigned long long g = blk_col0 + j;
            if (g < P.n_cols_total) {
                const unsigned long long poly = g >> log_cpp, gl = g & ((1ull << log_cpp) - 1);
                unsigned long long e;
                if (!P.last) {
                    const unsigned long long o = gl >> P.log_inner, c = gl & ((1ull << P.log_inner) - 1);
                    e = (o << (B + P.log_inner)) + ((unsigned long long)r << P.log_inner) + c;
                } else {
                    e = (ntt_digitrev_inv(P, gl) << B) + r;
                }
                if (!P.first || e < P.len_in) {
                    // coset mode, first pass: coefficient e of column poly / coset_cnt (scaled onto its coset below)
                    const unsigned long long sp = (P.first && P.coset_cnt) ? poly / P.coset_cnt : poly;
                    const uint4 *s = DIST ? P.src_tab[e >> P.dist_log_slab] + 2 * (e & ((1ull << P.dist_log_slab) - 1))
                                          : P.src + 2 * (sp * P.src_stride + e);
                    lo = s[0];
                    hi = s[1];
                }
            }
        }
        slo[r * PITCH + j] = lo;
        shi[r * PITCH + j] = hi;
    }
    __syncthreads();
    if (P.first && P.coset_cnt) {
        // scale onto the coset: c_e * W^(e rr) (a rolled loop with one product: keeps the unrolled load loop free of arithmetic)
        const unsigned long long nT = 1ull << P.tw_log_n;
#pragma unroll 1
        for (int idx = threadIdx.x; idx < TILE; idx += NT) {
            int r, j;
            if (!P.last) { j = idx % CC; r = idx / CC; } else { r = idx % R; j = idx / R; }
            const unsigned long long g = blk_col0 + j;
            if (g >= P.n_cols_total) continue;
            const unsigned long long poly = g >> log_cpp, gl = g & ((1ull << log_cpp) - 1);
            const unsigned long long rr = poly % P.coset_cnt + P.coset_r0;
            if (!rr) continue;
            unsigned long long e;
            if (!P.last) {
                const unsigned long long o = gl >> P.log_inner, c = gl & ((1ull << P.log_inner) - 1);
                e = (o << (B + P.log_inner)) + ((unsigned long long)r << P.log_inner) + c;
            } else {
                e = (ntt_digitrev_inv(P, gl) << B) + r;
            }
            const unsigned long long ti = ((e * rr) << (P.tw_log_stride - P.coset_log)) & (nT - 1);
            fp v = fp_mul(fp_from_u4(slo[r * PITCH + j], shi[r * PITCH + j]), fp_ldg_ro(P.tw, ti));
            slo[r * PITCH + j] = fp_lo(v);
            shi[r * PITCH + j] = fp_hi(v);
        }
        __syncthreads();
    }

    // ---- butterflies ----
    ntt_rounds<B, B, LOG_TILE>(slo, shi, wlo, whi);

    // ---- store tile (row r of the tile holds output k = bitrev_B(r)) ----
    fp ninv;
#pragma unroll
    for (int i = 0; i < 8; i++) ninv.l[i] = P.n_inv[i];
#pragma unroll STORE_UNROLL
    for (int idx = threadIdx.x; idx < TILE; idx += NT) {
        const int j = idx % CC;
        const uint32_t kk = idx / CC;
        const uint32_t r = B ? (__brev(kk) >> (32 - B)) : 0;
        const unsigned long long g = blk_col0 + j;
        if (g >= P.n_cols_total) continue;
        fp v = fp_from_u4(slo[r * PITCH + j], shi[r * PITCH + j]);
        const unsigned long long poly = g >> log_cpp, gl = g & ((1ull << log_cpp) - 1);
        unsigned long long e;
        fp w = ninv;                          // one product per element: the inter-pass twiddle, or n^-1 at the end of an inverse
        if (!P.last) {
            const unsigned long long o = gl >> P.log_inner, c = gl & ((1ull << P.log_inner) - 1);
            e = (o << (B + P.log_inner)) + ((unsigned long long)kk << P.log_inner) + c;
            const unsigned long long ex = ((unsigned long long)kk * c) << P.log_outer;   // < n
            w = fp_ldg_ro(P.tw, ntt_tw_index(P, ex));
        } else {
            e = gl + ((unsigned long long)kk << P.log_outer);
        }
        if (!P.last || P.inverse) v = fp_mul(v, w);
        if (P.last) v = fp_canon(v);
        if constexpr (DIST) {
            fp_stg(P.dst_tab[e >> P.dist_log_slab], e & ((1ull << P.dist_log_slab) - 1), v);
        } else if (P.last && P.coset_store == NTT_STORE_INTERLEAVED) {
            const unsigned long long col = poly / P.coset_cnt, rr = poly % P.coset_cnt + P.coset_r0;
            fp_stg(P.dst, col * P.dst_stride + (e << P.coset_log) + rr, v);
        } else {
            unsigned long long dp = poly;
            if (P.last && P.coset_cnt) dp = (poly / P.coset_cnt) * P.coset_dst_cpd + (poly % P.coset_cnt + P.coset_r0 - P.coset_dst_r0);
            fp_stg(P.dst, dp * P.dst_stride + e, v);
        }
    }
}
