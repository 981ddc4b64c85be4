// fri.cu -- launchers of the FRI fold / powers / batch-inverse kernels (fri.cuh)
#include "kernels.h"
#include "fri.cuh"

// slices of the two-level batch inverse: 148 SMs x 16 CTAs x 128 threads
static const unsigned BINV_THREADS = 148 * 16 * 128;
static const unsigned long long BINV_TWO_LEVEL_MIN = (unsigned long long)BINV_THREADS * 8;

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
// scratch must hold batch_inverse_scratch_elems(n) elements
unsigned long long batch_inverse_scratch_elems(unsigned long long n) {
    return n + (n > BINV_TWO_LEVEL_MIN ? 2 * (unsigned long long)BINV_THREADS : 0);
}
int batch_inverse_launch(cudaStream_t s, uint4 *vals, uint4 *scratch, unsigned long long n) {
    if (n <= BINV_TWO_LEVEL_MIN) {
        size_t threads = n < (size_t)BINV_THREADS ? n : (size_t)BINV_THREADS;
        batch_inverse_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(vals, scratch, n);
        return 1;
    }
    uint4 *tot = scratch + 2 * n, *tot_scratch = tot + 2 * (size_t)BINV_THREADS;
    batch_inverse_fwd_kernel<<<BINV_THREADS / 128, 128, 0, s>>>(vals, scratch, tot, n);
    batch_inverse_kernel<<<(BINV_THREADS / 64 + 127) / 128, 128, 0, s>>>(tot, tot_scratch, BINV_THREADS);
    batch_inverse_bwd_kernel<<<BINV_THREADS / 128, 128, 0, s>>>(vals, scratch, tot, n);
    return 3;
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
