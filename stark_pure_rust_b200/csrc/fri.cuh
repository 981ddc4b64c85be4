// fri.cuh -- FRI 4-to-1 fold and small vector helpers.
//
// Replaces the per-layer arithmetic of fri/src/fri.rs:141-164: for every row i < q = n/4 the
// reference Lagrange-interpolates the cubic through (x*iota^j, values[i + q*j]), j = 0..3
// (poly_utils.rs:449-511 multi_interp_4, one batch inverse over n denominators) and evaluates it
// at special_x (poly_utils.rs:442-446 eval_quartic).  The interpolant is unique, so the closed
// form below yields the same field element: with x = w^i, iota = w^q (primitive 4th root),
//     s_k = sum_j y_j iota^(-jk)            (inverse DFT-4 without the 1/4)
//     column[i] = 1/4 * sum_k s_k z^k ,     z = special_x * x^-1 ,   x^-1 = w^(n-i) from the table.
// 5 modmuls per row (z, iota^-1, three Horner steps); the 1/4 is two exact halvings.
#pragma once
#include "fp.cuh"
#include "fri_fold.cuh"
#include "params.h"


__global__ void __launch_bounds__(128) fri_fold_kernel(const __grid_constant__ FriFoldParams P) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t q = P.n >> 2;
    if (i >= q) return;
    fp sx;
#pragma unroll
    for (int k = 0; k < 8; k++) sx.l[k] = P.special_x[k];
    const unsigned long long nT = 1ull << P.tw_log_n;
    // iota^-1 = w^(-q) = w^(3q)
    fp iota_inv = fp_ldg_ro(P.tw, (nT - ((unsigned long long)q << P.tw_log_stride)) & (nT - 1));
    fp r = fri_fold_row(P, i, sx, iota_inv);
    fp_stg(P.col, i, r);
}

// ---- table of powers T[i] = w^i by doubling: T[cur + j] = T[j] * w^cur ------------------------
__global__ void powers_double_kernel(uint4 *T, unsigned long long cur, unsigned long long n_total,
                                     const fp wcur) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cur || cur + j >= n_total) return;
    fp v = fp_ldg(T, j);
    fp_stg(T, cur + j, fp_canon(fp_mul(v, wcur)));
}

// first `count` (<= 1024) powers in one CTA: log-step doubling in shared memory
__global__ void powers_seed_kernel(uint4 *T, unsigned long long count, const fp w_in) {
    __shared__ uint4 s[2048];
    const int t = threadIdx.x;
    if (t == 0) {
        fp one = fp_one();
        s[0] = fp_lo(one); s[1] = fp_hi(one);
    }
    __syncthreads();
    fp wc;                          // w^cur (plain copy: a __grid_constant__ source was miscompiled here)
#pragma unroll
    for (int k = 0; k < 8; k++) wc.l[k] = w_in.l[k];
    for (unsigned cur = 1; cur < count; cur <<= 1) {
        if ((unsigned)t < cur && cur + t < count) {
            fp v = fp_from_u4(s[2 * t], s[2 * t + 1]);
            fp r = fp_canon(fp_mul(v, wc));
            s[2 * (cur + t)] = fp_lo(r);
            s[2 * (cur + t) + 1] = fp_hi(r);
        }
        wc = fp_canon(fp_mul(wc, wc));
        __syncthreads();
    }
    for (unsigned i = t; i < count; i += blockDim.x) {
        T[2 * i] = s[2 * i];
        T[2 * i + 1] = s[2 * i + 1];
    }
}

// Montgomery -> canonical little-endian bytes (to_bytes_le, fp.rs:39-43) for n elements
__global__ void fp_to_bytes_kernel(const uint4 *in, uint4 *out, unsigned long long n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_stg(out, i, fp_from_mont(fp_ldg(in, i)));
}

// ---- multi_inv (poly_utils.rs:38-70): element-wise inverse, 0 -> 0 ----------------------------
// Thread t owns elements t, t+T, t+2T, ... (coalesced), runs Montgomery's trick over them with a
// zero-skipping prefix product and one Fermat inversion per thread.
__device__ __forceinline__ fp fp_inv_fermat(const fp &a) {
    // a^(p-2), square-and-multiply over the fixed exponent, MSB first
    const uint32_t e[8] = {0xeffffffFu, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    fp r = fp_one();
    bool started = false;
    for (int i = 7; i >= 0; i--) {
        for (int b = 31; b >= 0; b--) {
            if (started) r = fp_mul(r, r);
            if ((e[i] >> b) & 1) {
                r = started ? fp_mul(r, a) : a;
                started = true;
            }
        }
    }
    return r;
}

__global__ void __launch_bounds__(128) batch_inverse_kernel(uint4 *vals, uint4 *scratch, unsigned long long n) {
    const size_t T = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    // forward: scratch[i] = product of the non-zero elements before i (within this thread's slice)
    fp acc = fp_one();
    for (size_t i = t; i < n; i += T) {
        fp_stg(scratch, i, acc);
        fp v = fp_canon(fp_ldg(vals, i));
        if (!fp_is_zero_canon(v)) acc = fp_mul(acc, v);
    }
    fp inv = fp_inv_fermat(fp_canon(acc));
    // backward
    size_t cnt = (n - t + T - 1) / T;
    for (size_t k = cnt; k-- > 0;) {
        const size_t i = t + k * T;
        fp v = fp_canon(fp_ldg(vals, i));
        if (!fp_is_zero_canon(v)) {
            fp pre = fp_ldg(scratch, i);
            fp_stg(vals, i, fp_canon(fp_mul(inv, pre)));
            inv = fp_mul(inv, v);
        }
    }
}

// Two-level variant for large n: the per-thread Fermat inversion (~380 products) would dominate once a thread owns
// only a few dozen elements, so the slice products are themselves batch-inverted (recursively, by the kernel above)
// and the cost per element drops to the 3 products of Montgomery's trick.
//   fwd : scratch[i] = product of the non-zero elements before i in the slice; tot[t] = product of the whole slice
//   bwd : tot[t] holds the inverse of the slice product on entry
__global__ void __launch_bounds__(128) batch_inverse_fwd_kernel(const uint4 *vals, uint4 *scratch, uint4 *tot, unsigned long long n) {
    const size_t T = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    fp acc = fp_one();
    for (size_t i = t; i < n; i += T) {
        fp_stg(scratch, i, acc);
        fp v = fp_canon(fp_ldg(vals, i));
        if (!fp_is_zero_canon(v)) acc = fp_mul(acc, v);
    }
    fp_stg(tot, t, fp_canon(acc));      // never zero
}
__global__ void __launch_bounds__(128) batch_inverse_bwd_kernel(uint4 *vals, const uint4 *scratch, const uint4 *tot, unsigned long long n) {
    const size_t T = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    fp inv = fp_ldg(tot, t);
    size_t cnt = (n - t + T - 1) / T;
    for (size_t k = cnt; k-- > 0;) {
        const size_t i = t + k * T;
        fp v = fp_canon(fp_ldg(vals, i));
        if (!fp_is_zero_canon(v)) {
            fp pre = fp_ldg(scratch, i);
            fp_stg(vals, i, fp_canon(fp_mul(inv, pre)));
            inv = fp_mul(inv, v);
        }
    }
}

// ---- element-wise field ops on vectors (unit tests of fp.cuh through the C ABI) -----------------
// op: 0 mul (raw, lazy result), 1 add, 2 sub, 3 sub_lazy, 4 canon(a), 5 half(a), 6 from_mont(a),
//     7 to_mont(a), 8 inverse(a), 9 reduce_2p(a), 10 canon(mul)
__global__ void fp_vec_op_kernel(int op, const uint4 *a, const uint4 *b, uint4 *out, unsigned long long n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp x = fp_ldg(a, i), y = fp_ldg(b, i), r;
    switch (op) {
        case 0: r = fp_mul(x, y); break;
        case 1: r = fp_add(x, y); break;
        case 2: r = fp_sub(x, y); break;
        case 3: r = fp_sub_lazy(x, y); break;
        case 4: r = fp_canon(x); break;
        case 5: r = fp_half(x); break;
        case 6: r = fp_from_mont(x); break;
        case 7: r = fp_to_mont(x); break;
        case 8: r = fp_canon(fp_inv_fermat(x)); break;
        case 9: r = fp_reduce_2p(x); break;
        case 10: r = fp_canon(fp_mul(x, y)); break;
        case 11: r = fp_mul(x, x); break;
        default: {
            r = x;
            for (uint32_t k = 0; k < y.l[0]; k++) r = fp_canon(fp_mul(r, r));
            break;
        }
    }
    fp_stg(out, i, r);
}

// ---- four-step twiddles (distributed transform, sharded.py): vals[r][c] *= w^((row0 + r) * c), w = T[1 << stride] ----
__global__ void twiddle_mul_kernel(uint4 *vals, unsigned long long rows, unsigned long long cols, unsigned long long row0,
                                   const uint4 *tw, uint32_t tw_log_n, uint32_t tw_log_stride, uint32_t log_n, int inverse) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * cols) return;
    const unsigned long long r = t / cols, c = t % cols;
    const unsigned long long n_mask = (1ull << log_n) - 1, nT = 1ull << tw_log_n;
    const unsigned long long e = ((row0 + r) * c) & n_mask;
    unsigned long long ti = e << tw_log_stride;
    if (inverse) ti = (nT - ti) & (nT - 1);
    fp v = fp_ldg(vals, t);
    fp_stg(vals, t, fp_canon(fp_mul(v, fp_ldg_ro(tw, ti))));
}

// ---- issue-rate probes (bench.py's integer roofline): no memory traffic inside the loop ----------------------
// mode 0: independent chains of Montgomery products (what every arithmetic kernel here is made of)
// mode 1: independent chains of IMAD.WIDE.U32 (the instruction a Montgomery product is made of: 128 per product)
template <int MODE>
__global__ void __launch_bounds__(256) pipe_probe_kernel(uint4 *out, uint32_t iters, const fp seed) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 0) {
        fp x[2], m;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            m.l[k] = seed.l[k];
            x[0].l[k] = seed.l[k] ^ (t * 2654435761u);
            x[1].l[k] = seed.l[(k + 3) & 7] + t;
        }
        x[0].l[7] &= 0x0fffffffu; x[1].l[7] &= 0x0fffffffu;
#pragma unroll 1
        for (uint32_t i = 0; i < iters; i++) {
            x[0] = fp_mul(x[0], m);
            x[1] = fp_mul(x[1], m);
        }
        fp r = fp_add(x[0], x[1]);
        fp_stg(out, t, r);
    } else {
        unsigned long long w[8];
        const uint32_t m = seed.l[0] | 1u;
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = ((unsigned long long)(t + k) << 32) | seed.l[k];
#pragma unroll 1
        for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; xor.b32 lo, lo, hi; mul.wide.u32 %0, lo, %1;}" : "+l"(w[k]) : "r"(m));
            }
        }
        unsigned long long acc = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) acc += w[k];
        out[t] = make_uint4((uint32_t)acc, (uint32_t)(acc >> 32), 0, 0);
    }
}
