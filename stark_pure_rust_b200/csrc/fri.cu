// fri.cu -- launchers of the FRI fold / powers / batch-inverse kernels (fri.cuh)
#include "kernels.h"
#include "fri.cuh"

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
int batch_inverse_launch(cudaStream_t s, uint4 *vals, uint4 *scratch, unsigned long long n) {
    size_t threads = n < (size_t)148 * 16 * 128 ? n : (size_t)148 * 16 * 128;
    batch_inverse_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(vals, scratch, n);
    return 1;
}
int fp_launch_vec_op(cudaStream_t s, int op, const uint4 *a, const uint4 *b, uint4 *out, unsigned long long n) {
    fp_vec_op_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(op, a, b, out, n);
    return 1;
}
