// pointwise.cuh -- the element-wise stage of mk_r1cs_proof (r1cs-stark/src/prove.rs:133-232, :287-322)
// between the NTT and Merkle kernels: constraint quotients, accumulator prefix products, boundary
// quotients and the 11-term linear combination, all on device-resident columns (Montgomery in memory).
// Formulas: SURVEY.md A.9; indices are mod N, sk = 8 (EXTENSION_FACTOR, utils.rs:134).
//
// inv_z (prove.rs:203, utils.rs:173-178): z[j] = g2^(jS) - 1 takes 8 values because g2^S is a primitive
// 8th root of unity, so multi_inv(z) is a table of 8 inverses with entry 0 equal to 0 (0 -> 0 rule,
// poly_utils.rs:38-70); likewise X = (g2^S)^j in the linear combination (prove.rs:287-322).
#pragma once
#include "fp.cuh"
#include "params.h"

// Layout: every N-point column is coset-major and sharded by cosets (sb_ext, internal.h).  A kernel launch covers ONE device's
// cosets r0 .. r0 + cpd - 1: thread t < cpd * S handles coset r = r0 + (t >> log_s) at step k = t & (S - 1), i.e. domain
// position j = 8 k + r, and element t of every column pointer (which points at that device's block of the column).  The
// prover's shifts are whole steps: position j - 8 is (r, k - 1), j + 8 o3 is (r, k + o3), all mod S inside the same coset.
struct PwView {
    uint32_t log_s, r0;
    unsigned long long n;            // cpd << log_s
};
#define PW_DECODE(V)                                                          \
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;           \
    if (t >= (V).n) return;                                                   \
    const size_t S_ = (size_t)1 << (V).log_s, k_ = t & (S_ - 1), cb_ = t - k_; \
    const uint32_t r_ = (V).r0 + (uint32_t)(t >> (V).log_s);                  \
    const size_t j_ = (k_ << 3) + r_;                                         \
    (void)j_; (void)cb_;
#define PW_SHIFT(steps) (cb_ + ((k_ + (steps)) & (S_ - 1)))

struct PwConsts {
    uint32_t inv_z8[8][8];     // multi_inv(z)[j] for j mod 8
    uint32_t pw8[8][8];        // (g2^S)^(j mod 8)
    uint32_t r[3][8];          // accumulator challenges (utils.rs:272-290)
    uint32_t k[11][8];         // linear-combination challenges (prove.rs:274-283)
    uint32_t kx[3][8][8];      // per coset r: k3 + k4 X_r, k5 + k6 X_r, k7 + k8 X_r with X_r = (g2^S)^r (the p, b2, b3 terms of l)
    uint32_t one[8];
    uint32_t x_last[8];        // xs[N - sk] (utils.rs:459)
};

__device__ __forceinline__ fp pw_const(const uint32_t (&c)[8]) {
    fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = c[i];
    return r;
}
__device__ __forceinline__ bool pw_is_zero(const fp &a) { return fp_is_zero_canon(fp_canon(a)); }

// out[i] = Fp::from(v[i]) (v == NULL: v[i] = i): the index columns of prove.rs:160-167
__global__ void pw_u64_to_fp_kernel(const unsigned long long *v, uint4 *out, unsigned long long n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long x = v ? v[i] : (unsigned long long)i;
    fp a = fp_zero();
    a.l[0] = (uint32_t)x;
    a.l[1] = (uint32_t)(x >> 32);
    fp_stg(out, i, fp_to_mont(a));
}

// flag vectors generated on the device (FlagSpec, internal.h; run.rs:283-308): out[i] = value
__global__ void pw_fill_kernel(uint4 *out, unsigned long long n, const fp value) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fp_stg(out, i, value);
}
// f1[k] = f1[k + a] = f1[k + 2a] = 0 with k = (l + 1) mod a, f2[l] = 1 for the last row l of every constraint (either may be NULL)
__global__ void pw_flags_scatter_kernel(const unsigned long long *last_rows, unsigned long long n_last, unsigned long long a, uint4 *f1, uint4 *f2,
                                        const fp one) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_last) return;
    const unsigned long long l = last_rows[c], k = (l + 1) % a;
    if (f1) {
        const fp z = fp_zero();
        fp_stg(f1, k, z);
        fp_stg(f1, k + a, z);
        fp_stg(f1, k + 2 * a, z);
    }
    if (f2) fp_stg(f2, l, one);
}

// identity padding of the copy permutation beyond the original steps (prove.rs:55-56): perm[i] = i for os <= i < n
__global__ void pw_perm_pad_kernel(unsigned long long *perm, unsigned long long os, unsigned long long n) {
    const size_t i = os + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = i;
}

// accumulator-tree leaves (utils.rs:250-270): u64_LE(permuted_index[j]) || to_bytes_le(witness_trace[j]), 40 B
__global__ void pw_a_leaves_kernel(const unsigned long long *perm, const uint4 *wit, uint32_t *leaves, unsigned long long n) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    fp w = fp_from_mont(fp_ldg(wit, j));
    uint32_t *o = leaves + 10 * j;
    o[0] = (uint32_t)perm[j];
    o[1] = (uint32_t)(perm[j] >> 32);
#pragma unroll
    for (int i = 0; i < 8; i++) o[2 + i] = w.l[i];
}

// d1 = q1 * inv_z, d2 = q2 * inv_z  (utils.rs:181-248, :379-418)
//   q1[j] = f0[j] * (p[j] - f1[j] p[j - sk] - k[j] s[j]);  q2[j] = f2[j] * (p[j + 2 o3 sk] - p[j] p[j + o3 sk])
struct PwQ12Params {
    const uint4 *k, *f0, *f1, *f2, *s, *p;
    uint4 *d1, *d2;
    PwView v;
    unsigned long long o3;           // original_steps / 3
    int *err;                        // set when q != 0 where inv_z == 0 (reference: assert, utils.rs:379-418)
};
__global__ void __launch_bounds__(128) pw_q12_kernel(const __grid_constant__ PwQ12Params P, const __grid_constant__ PwConsts Cst) {
    PW_DECODE(P.v)
    fp pj = fp_ldg(P.p, t);
    fp t1 = fp_mul(fp_ldg(P.f1, t), fp_ldg(P.p, PW_SHIFT(S_ - 1)));
    fp t2 = fp_mul(fp_ldg(P.k, t), fp_ldg(P.s, t));
    fp q1 = fp_mul(fp_sub(fp_sub(pj, t1), t2), fp_ldg(P.f0, t));
    fp t3 = fp_mul(pj, fp_ldg(P.p, PW_SHIFT(P.o3)));
    fp q2 = fp_mul(fp_sub(fp_ldg(P.p, PW_SHIFT(2 * P.o3)), t3), fp_ldg(P.f2, t));
    fp iz = pw_const(Cst.inv_z8[r_]);
    if (r_ == 0 && !(pw_is_zero(q1) && pw_is_zero(q2))) atomicExch(P.err, 1);
    fp_stg(P.d1, t, fp_canon(fp_mul(q1, iz)));
    fp_stg(P.d2, t, fp_canon(fp_mul(q2, iz)));
}

// accumulator factors (utils.rs:293-339): nmr_j = r0 + r1*j + r2*w_j, dnm_j = r0 + r1*perm[j] + r2*w_j
// (the reference reads j and perm[j] back from the index LDEs at j*sk, where they equal the inputs)
__global__ void __launch_bounds__(128) pw_acc_terms_kernel(const unsigned long long *perm, const uint4 *wit, uint4 *nmr, uint4 *dnm,
                                                          unsigned long long n, const __grid_constant__ PwConsts Cst) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    fp r0 = pw_const(Cst.r[0]), r1 = pw_const(Cst.r[1]), r2 = pw_const(Cst.r[2]);
    fp a = fp_zero(), b = fp_zero();
    a.l[0] = (uint32_t)j; a.l[1] = (uint32_t)((unsigned long long)j >> 32);
    b.l[0] = (uint32_t)perm[j]; b.l[1] = (uint32_t)(perm[j] >> 32);
    fp t2 = fp_mul(r2, fp_ldg(wit, j));
    fp vn = fp_add(fp_add(r0, fp_mul(r1, fp_to_mont(a))), t2);
    fp vd = fp_add(fp_add(r0, fp_mul(r1, fp_to_mont(b))), t2);
    fp_stg(nmr, j, fp_canon(vn));
    fp_stg(dnm, j, fp_canon(vd));
}

// ---- inclusive prefix product over n elements in three launches ------------------------------------
// chunk c = elements [c*len, (c+1)*len): (1) per-chunk products, (2) exclusive scan of the chunk products
// in one CTA, (3) per-chunk rescan seeded with the chunk's prefix.
__global__ void __launch_bounds__(128) scan_chunk_reduce_kernel(const uint4 *in, uint4 *partial, unsigned long long n, unsigned long long len) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = c * len;
    if (lo >= n) return;
    const size_t hi = lo + len < n ? lo + len : n;
    fp acc = fp_ldg(in, lo);
    for (size_t i = lo + 1; i < hi; i++) acc = fp_mul(acc, fp_ldg(in, i));
    fp_stg(partial, c, fp_canon(acc));
}
// in place: partial[c] <- product of partial[0..c) (exclusive), n_chunks <= 1024 * per
__global__ void __launch_bounds__(1024) scan_partials_kernel(uint4 *partial, unsigned long long n_chunks) {
    __shared__ uint4 s[2048];
    const unsigned t = threadIdx.x;
    const size_t per = (n_chunks + 1023) / 1024;
    const size_t lo = t * per, hi = lo + per < n_chunks ? lo + per : n_chunks;
    fp acc = fp_one();
    for (size_t i = lo; i < hi; i++) acc = fp_mul(acc, fp_ldg(partial, i));
    acc = fp_canon(acc);
    s[2 * t] = fp_lo(acc); s[2 * t + 1] = fp_hi(acc);
    __syncthreads();
    for (unsigned d = 1; d < 1024; d <<= 1) {        // Hillis-Steele inclusive scan of the 1024 thread totals
        fp v = fp_from_u4(s[2 * t], s[2 * t + 1]);
        fp u = v;
        if (t >= d) u = fp_canon(fp_mul(fp_from_u4(s[2 * (t - d)], s[2 * (t - d) + 1]), v));
        __syncthreads();
        s[2 * t] = fp_lo(u); s[2 * t + 1] = fp_hi(u);
        __syncthreads();
    }
    fp pre = t ? fp_from_u4(s[2 * (t - 1)], s[2 * (t - 1) + 1]) : fp_one();
    for (size_t i = lo; i < hi; i++) {
        fp v = fp_ldg(partial, i);
        fp_stg(partial, i, fp_canon(pre));
        pre = fp_mul(pre, v);
    }
}
__global__ void __launch_bounds__(128) scan_chunk_apply_kernel(const uint4 *in, const uint4 *partial_excl, uint4 *out, unsigned long long n,
                                                              unsigned long long len) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = c * len;
    if (lo >= n) return;
    const size_t hi = lo + len < n ? lo + len : n;
    fp acc = fp_ldg(partial_excl, c);
    for (size_t i = lo; i < hi; i++) {
        acc = fp_mul(acc, fp_ldg(in, i));
        fp_stg(out, i, fp_canon(acc));
    }
}

// out[i] = a[i] * b[i]
__global__ void __launch_bounds__(128) pw_mul_kernel(const uint4 *a, const uint4 *b, uint4 *out, unsigned long long n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_stg(out, i, fp_canon(fp_mul(fp_ldg(a, i), fp_ldg(b, i))));
}

// d3 = q3 * inv_z (utils.rs:344-376): q3[j] = a[j] (r0 + r1 pidx[j] + r2 s[j]) - a[j - sk] (r0 + r1 idx[j] + r2 s[j])
struct PwQ3Params {
    const uint4 *a, *s, *idx, *pidx;
    uint4 *d3;
    PwView v;
    int *err;
};
__global__ void __launch_bounds__(128) pw_q3_kernel(const __grid_constant__ PwQ3Params P, const __grid_constant__ PwConsts Cst) {
    PW_DECODE(P.v)
    fp r0 = pw_const(Cst.r[0]), r1 = pw_const(Cst.r[1]), r2 = pw_const(Cst.r[2]);
    fp t2 = fp_mul(r2, fp_ldg(P.s, t));
    fp vn = fp_add(fp_add(r0, fp_mul(r1, fp_ldg(P.idx, t))), t2);
    fp vd = fp_add(fp_add(r0, fp_mul(r1, fp_ldg(P.pidx, t))), t2);
    fp u = fp_mul(fp_ldg(P.a, t), vd);
    fp v = fp_mul(fp_ldg(P.a, PW_SHIFT(S_ - 1)), vn);
    fp q3 = fp_sub(u, v);
    if (r_ == 0 && !pw_is_zero(q3)) atomicExch(P.err, 1);
    fp_stg(P.d3, t, fp_canon(fp_mul(q3, pw_const(Cst.inv_z8[r_]))));
}

// zb3[j] = xs[j] - xs[N - sk] (utils.rs:458-474), written behind zb2 so that one batch inverse serves both
__global__ void __launch_bounds__(128) pw_zb3_kernel(const uint4 *xs, uint4 *zb3, PwView V, const __grid_constant__ PwConsts Cst) {
    PW_DECODE(V)
    fp_stg(zb3, t, fp_canon(fp_sub(fp_ldg(xs, j_), pw_const(Cst.x_last))));
}

// out[j] = sum_k coef[k] xs[j]^k (Horner): the boundary polynomials i2 / zb2 when only a few public wires are in use
// (prove.rs:216-224 evaluates them point by point as well; for many public wires the coset transforms of the same
// coefficients are used instead -- identical field elements either way)
__global__ void __launch_bounds__(128) pw_poly_eval_kernel(const uint4 *xs, const uint4 *coef, uint32_t n_coef, uint4 *out, PwView V) {
    PW_DECODE(V)
    const fp x = fp_ldg(xs, j_);
    fp acc = fp_ldg_ro(coef, n_coef - 1);
    for (uint32_t k = n_coef - 1; k-- > 0;) acc = fp_add(fp_mul(acc, x), fp_ldg_ro(coef, k));
    fp_stg(out, t, fp_canon(acc));
}

// b2 = (s - i2) * inv(zb2), b3 = (a - 1) * inv(zb3)  (utils.rs:477-524); in place over the inverse arrays
struct PwB23Params {
    const uint4 *s, *a, *i2;     // i2 == NULL: the interpolant is the zero polynomial (no public wire in use)
    uint4 *inv_zb2, *inv_zb3;    // in: inverses, out: b2, b3
    PwView v;
    int *err;
};
__global__ void __launch_bounds__(128) pw_b23_kernel(const __grid_constant__ PwB23Params P, const __grid_constant__ PwConsts Cst) {
    PW_DECODE(P.v)
    fp i2 = P.i2 ? fp_ldg(P.i2, t) : fp_zero();
    fp df2 = fp_sub(fp_ldg(P.s, t), i2);
    fp inv2 = fp_ldg(P.inv_zb2, t);
    if (pw_is_zero(inv2) && !pw_is_zero(df2)) atomicExch(P.err, 2);     // utils.rs:489
    fp_stg(P.inv_zb2, t, fp_canon(fp_mul(df2, inv2)));
    fp df3 = fp_sub(fp_ldg(P.a, t), pw_const(Cst.one));
    fp inv3 = fp_ldg(P.inv_zb3, t);
    if (pw_is_zero(inv3) && !pw_is_zero(df3)) atomicExch(P.err, 3);     // utils.rs:514
    fp_stg(P.inv_zb3, t, fp_canon(fp_mul(df3, inv3)));
}

// l[j] = k0 d1 + k1 d2 + k2 d3 + k3 p + k4 p X + k5 b2 + k6 b2 X + k7 b3 + k8 b3 X + k9 a + k10 s,  X = (g2^S)^j
// (prove.rs:287-322)
struct PwLParams {
    const uint4 *d1, *d2, *d3, *p, *b2, *b3, *a, *s;
    uint4 *l;
    PwView v;
};
__global__ void __launch_bounds__(128) pw_l_kernel(const __grid_constant__ PwLParams P, const __grid_constant__ PwConsts Cst) {
    PW_DECODE(P.v)
    // k3 p + k4 p X = (k3 + k4 X) p, and X takes one value per coset: 8 products per row instead of 14, same field element
    fp acc = fp_mul(fp_ldg(P.d1, t), pw_const(Cst.k[0]));
    acc = fp_add(acc, fp_mul(fp_ldg(P.d2, t), pw_const(Cst.k[1])));
    acc = fp_add(acc, fp_mul(fp_ldg(P.d3, t), pw_const(Cst.k[2])));
    acc = fp_add(acc, fp_mul(fp_ldg(P.p, t), pw_const(Cst.kx[0][r_])));
    acc = fp_add(acc, fp_mul(fp_ldg(P.b2, t), pw_const(Cst.kx[1][r_])));
    acc = fp_add(acc, fp_mul(fp_ldg(P.b3, t), pw_const(Cst.kx[2][r_])));
    acc = fp_add(acc, fp_mul(fp_ldg(P.a, t), pw_const(Cst.k[9])));
    acc = fp_add(acc, fp_mul(fp_ldg(P.s, t), pw_const(Cst.k[10])));
    fp_stg(P.l, t, fp_canon(acc));
}

// ---- interpolation through n public points on the device (poly_utils.rs:409-439 lagrange_interp, :362-373 zpoly) ----------
// The reference's O(n^2) scalar code; n = number of public wires in use (1062 for the bundled `bits` circuit, where the
// threaded host version took 30 of the prover's 36 ms).  Same field elements (the interpolant is unique), computed as
//     Z(X) = prod (X - x_i)                                       interp_zpoly_kernel (one CTA, coefficients in shared memory)
//     s_i  = y_i / Z'(x_i),  Z'(x_i) = prod_{j != i} (x_i - x_j)  interp_denoms_kernel + the batch inverse + a product
//     P[t] = sum_i s_i x_i^t                                      interp_powers_kernel + interp_powersum_kernel
//     I[k] = sum_{m > k} Z[m] P[m - k - 1]                        interp_conv_kernel
// because Z(X) / (X - x_i) = sum_k X^k sum_{m > k} Z[m] x_i^(m - k - 1).

// Z's n + 1 coefficients, lowest first: n steps c <- c * (X - x_j), every thread owning a strided share of the coefficients
__global__ void __launch_bounds__(1024) interp_zpoly_kernel(const uint4 *xs, uint32_t n, uint4 *zcoef) {
    extern __shared__ uint4 zs[];                 // two buffers of n + 1 elements
    uint4 *cur = zs, *nxt = zs + 2 * (n + 1);
    for (uint32_t k = threadIdx.x; k <= n; k += blockDim.x) {
        fp v = k == 0 ? fp_one() : fp_zero();
        cur[2 * k] = fp_lo(v);
        cur[2 * k + 1] = fp_hi(v);
    }
    __syncthreads();
    for (uint32_t j = 0; j < n; j++) {            // degree j -> j + 1
        const fp x = fp_ldg_ro(xs, j);
        for (uint32_t k = threadIdx.x; k <= j + 1; k += blockDim.x) {
            fp lower = k ? fp_from_u4(cur[2 * (k - 1)], cur[2 * (k - 1) + 1]) : fp_zero();
            fp v = lower;
            if (k <= j) v = fp_sub(lower, fp_mul(fp_from_u4(cur[2 * k], cur[2 * k + 1]), x));
            nxt[2 * k] = fp_lo(v);
            nxt[2 * k + 1] = fp_hi(v);
        }
        __syncthreads();
        uint4 *t = cur; cur = nxt; nxt = t;
    }
    for (uint32_t k = threadIdx.x; k <= n; k += blockDim.x) fp_stg(zcoef, k, fp_canon(fp_from_u4(cur[2 * k], cur[2 * k + 1])));
}
// d[i] = prod_{j != i} (x_i - x_j)
__global__ void __launch_bounds__(128) interp_denoms_kernel(const uint4 *xs, uint32_t n, uint4 *d) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fp xi = fp_ldg_ro(xs, i);
    fp acc = fp_one();
    for (uint32_t j = 0; j < n; j++)
        if (j != i) acc = fp_mul(fp_sub_lazy(xi, fp_ldg_ro(xs, j)), acc);
    fp_stg(d, i, fp_canon(acc));
}
// tab[t * n + i] = s_i x_i^t, t < n
__global__ void __launch_bounds__(128) interp_powers_kernel(const uint4 *xs, const uint4 *s, uint32_t n, uint4 *tab) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fp x = fp_ldg_ro(xs, i);
    fp v = fp_ldg(s, i);
    for (uint32_t t = 0; t < n; t++) {
        fp_stg(tab, (size_t)t * n + i, v);
        v = fp_mul(v, x);
    }
}
// P[t] = sum_i tab[t * n + i]: one CTA per t
__global__ void __launch_bounds__(128) interp_powersum_kernel(const uint4 *tab, uint32_t n, uint4 *P) {
    __shared__ uint4 red[256];
    const uint32_t t = blockIdx.x;
    fp acc = fp_zero();
    for (uint32_t i = threadIdx.x; i < n; i += 128) acc = fp_add(acc, fp_ldg(tab, (size_t)t * n + i));
    red[2 * threadIdx.x] = fp_lo(acc);
    red[2 * threadIdx.x + 1] = fp_hi(acc);
    __syncthreads();
    for (uint32_t w = 64; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            fp a = fp_from_u4(red[2 * threadIdx.x], red[2 * threadIdx.x + 1]);
            fp b = fp_from_u4(red[2 * (threadIdx.x + w)], red[2 * (threadIdx.x + w) + 1]);
            a = fp_add(a, b);
            red[2 * threadIdx.x] = fp_lo(a);
            red[2 * threadIdx.x + 1] = fp_hi(a);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) fp_stg(P, t, fp_canon(fp_from_u4(red[0], red[1])));
}
// out[k] = sum_{m = k + 1}^{n} Z[m] P[m - k - 1], k < n
__global__ void __launch_bounds__(128) interp_conv_kernel(const uint4 *z, const uint4 *P, uint32_t n, uint4 *out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    fp acc = fp_zero();
    for (uint32_t m = k + 1; m <= n; m++) acc = fp_add(acc, fp_mul(fp_ldg_ro(z, m), fp_ldg_ro(P, m - k - 1)));
    fp_stg(out, k, fp_canon(acc));
}
