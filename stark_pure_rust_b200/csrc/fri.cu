// fri.cu -- launchers of the FRI fold / powers / batch-inverse kernels (fri.cuh)
#include "kernels.h"
#include "fri.cuh"

// Batch inverse by levels: n elements are cut into T slices (one thread each: Montgomery's trick inside the slice, 3 products
// per element), the T slice products are inverted by the next level, and only the last level (<= BINV_DIRECT slices) pays
// a Fermat inversion (~380 products) per thread.  The first version inverted up to 303 104 slice products directly: 1.9 ms
// for the prover's 2^20-element accumulator denominators; with levels the same call is a few hundred microseconds.
static const unsigned BINV_MAX_THREADS = 148 * 16 * 128;
static const unsigned long long BINV_PER_SLICE = 32, BINV_DIRECT = 2048;
static unsigned long long binv_slices(unsigned long long n) {
    unsigned long long t = (n + BINV_PER_SLICE - 1) / BINV_PER_SLICE;
    if (t > BINV_MAX_THREADS) t = BINV_MAX_THREADS;
    return (t + 127) / 128 * 128;          // whole CTAs: the forward kernel writes a total for every launched thread
}

int fri_launch_fold(cudaStream_t s, const FriFoldParams &P) {
    const size_t q = P.n >> 2;
    fri_fold_kernel<<<(unsigned)((q + 127) / 128), 128, 0, s>>>(P);
    return 1;
}
int powers_launch_seed(cudaStream_t s, uint4 *T, unsigned long long count, const fp &w) {
    powers_seed_kernel<<<1, 1024, 0, s>>>(T, count, w);
    return 1;
}
int powers_launch_double(cudaStream_t s, uint4 *T, unsigned long long cur, unsigned long long n_total, const fp &wcur) {
    unsigned long long cnt = (n_total - cur) < cur ? (n_total - cur) : cur;
    powers_double_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(T, cur, n_total, wcur);
    return 1;
}
int fp_launch_to_bytes(cudaStream_t s, const uint4 *in, uint4 *out, unsigned long long n) {
    fp_to_bytes_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(in, out, n);
    return 1;
}
// scratch must hold batch_inverse_scratch_elems(n) elements: per level the prefix products (n_l) and the slice totals (T_l)
unsigned long long batch_inverse_scratch_elems(unsigned long long n) {
    unsigned long long tot = 0;
    while (n > BINV_DIRECT) {
        const unsigned long long t = binv_slices(n);
        tot += n + t;
        n = t;
    }
    return tot + n;
}
int batch_inverse_launch(cudaStream_t s, uint4 *vals, uint4 *scratch, unsigned long long n) {
    if (n == 0) return 0;
    if (n <= BINV_DIRECT) {
        const unsigned long long threads = (n + 7) / 8;       // a few elements per Fermat inversion
        batch_inverse_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(vals, scratch, n);
        return 1;
    }
    const unsigned long long t = binv_slices(n);
    uint4 *pre = scratch, *tot = scratch + 2 * n, *rest = tot + 2 * t;
    batch_inverse_fwd_kernel<<<(unsigned)(t / 128), 128, 0, s>>>(vals, pre, tot, n);
    const int inner = batch_inverse_launch(s, tot, rest, t);
    batch_inverse_bwd_kernel<<<(unsigned)(t / 128), 128, 0, s>>>(vals, pre, tot, n);
    return inner + 2;
}
int fp_launch_vec_op(cudaStream_t s, int op, const uint4 *a, const uint4 *b, uint4 *out, unsigned long long n) {
    fp_vec_op_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(op, a, b, out, n);
    return 1;
}

// returns the number of probe operations (Montgomery products / IMAD.WIDE) the launch performs in total
double pipe_probe_launch(cudaStream_t s, int mode, uint4 *out, unsigned blocks, uint32_t iters, const fp &seed) {
    if (mode == 0) {
        pipe_probe_kernel<0><<<blocks, 256, 0, s>>>(out, iters, seed);
        return 2.0 * iters * blocks * 256.0;
    }
    pipe_probe_kernel<1><<<blocks, 256, 0, s>>>(out, iters, seed);
    return 32.0 * iters * blocks * 256.0;
}

int twiddle_mul_launch(cudaStream_t s, uint4 *vals, unsigned long long rows, unsigned long long cols, unsigned long long row0,
                       const uint4 *tw, uint32_t tw_log_n, uint32_t tw_log_stride, uint32_t log_n, int inverse) {
    const unsigned long long total = rows * cols;
    if (total == 0) return 0;
    twiddle_mul_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(vals, rows, cols, row0, tw, tw_log_n, tw_log_stride, log_n, inverse);
    return 1;
}
