/*
 * fft.c — radix-2 NTT restating /root/reference/packages/fri/src/fft.rs.
 * ORACLE / TEST INFRASTRUCTURE — see oracle.h.
 *
 *   expand_root_of_unity  fft.rs:5-14
 *   serial_fft            fft.rs:150-193  (bit-reversal, then log_n DIT stages, running twiddle)
 *   parallel_fft          fft.rs:195-251  (2^log_cpus pre-twiddled sub-FFTs on threads + un-shuffle)
 *   inv_*_fft             fft.rs:284-309  (forward with root^-1, then * n^-1)
 *   best_fft/inv_best_fft fft.rs:327-379  (zero-pad, dispatch on Worker::cpus)
 */
#include "oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

size_t orc_expand_root_of_unity(fp_t *out, size_t cap, const fp_t *root) {
    size_t n = 0;
    fp_t cur = *root;
    if (cap) out[0] = FP_ONE;
    n = 1;
    while (!fp_eq(&cur, &FP_ONE)) {
        if (n < cap) out[n] = cur;
        n++;
        fp_mul(&cur, &cur, root);
    }
    return n;
}

static inline uint32_t bit_reverse(uint32_t n, uint32_t l) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < l; i++) {
        r = (r << 1) | (n & 1);
        n >>= 1;
    }
    return r;
}

void orc_serial_fft(fp_t *values, const fp_t *root, uint32_t log_n) {
    const uint32_t n = 1u << log_n;
    for (uint32_t k = 0; k < n; k++) {
        uint32_t rk = bit_reverse(k, log_n);
        if (k < rk) {
            fp_t tmp = values[rk];
            values[rk] = values[k];
            values[k] = tmp;
        }
    }
    uint32_t m = 1;
    for (uint32_t s = 0; s < log_n; s++) {
        fp_t w_m;
        fp_pow_u64(&w_m, root, n / (2 * m));
        for (uint32_t k = 0; k < n; k += 2 * m) {
            fp_t w = FP_ONE;
            for (uint32_t j = 0; j < m; j++) {
                fp_t t, tmp;
                fp_mul(&t, &values[k + j + m], &w);
                fp_sub(&tmp, &values[k + j], &t);
                values[k + j + m] = tmp;
                fp_add(&values[k + j], &values[k + j], &t);
                fp_mul(&w, &w, &w_m);
            }
        }
        m *= 2;
    }
}

typedef struct {
    const fp_t *values;
    fp_t *tmp;
    const fp_t *root;
    fp_t new_omega;
    uint32_t log_n, log_cpus, log_new_n, j;
} sub_fft_job;

static void *sub_fft_thread(void *arg) {
    sub_fft_job *job = (sub_fft_job *)arg;
    const uint32_t num_cpus = 1u << job->log_cpus;
    const uint32_t new_n = 1u << job->log_new_n;
    const uint32_t mask = (1u << job->log_n) - 1;
    fp_t omega_j, omega_step, elt = FP_ONE;
    fp_pow_u64(&omega_j, job->root, job->j);
    fp_pow_u64(&omega_step, job->root, (uint64_t)job->j << job->log_new_n);
    for (uint32_t i = 0; i < new_n; i++) {
        for (uint32_t s = 0; s < num_cpus; s++) {
            uint32_t idx = (i + (s << job->log_new_n)) & mask;
            fp_t t;
            fp_mul(&t, &job->values[idx], &elt);
            fp_add(&job->tmp[i], &job->tmp[i], &t);
            fp_mul(&elt, &elt, &omega_step);
        }
        fp_mul(&elt, &elt, &omega_j);
    }
    orc_serial_fft(job->tmp, &job->new_omega, job->log_new_n);
    return NULL;
}

void orc_parallel_fft(fp_t *values, const fp_t *root, uint32_t log_n, uint32_t log_cpus) {
    if (log_n < log_cpus) abort(); /* fft.rs:202 assert */
    const uint32_t num_cpus = 1u << log_cpus;
    const uint32_t log_new_n = log_n - log_cpus;
    const size_t new_n = (size_t)1 << log_new_n;
    fp_t *tmp = (fp_t *)calloc((size_t)num_cpus * new_n, sizeof(fp_t));
    sub_fft_job *jobs = (sub_fft_job *)malloc(num_cpus * sizeof(sub_fft_job));
    pthread_t *th = (pthread_t *)malloc(num_cpus * sizeof(pthread_t));
    fp_t new_omega;
    fp_pow_u64(&new_omega, root, num_cpus);
    for (uint32_t j = 0; j < num_cpus; j++) {
        jobs[j].values = values;
        jobs[j].tmp = tmp + (size_t)j * new_n;
        jobs[j].root = root;
        jobs[j].new_omega = new_omega;
        jobs[j].log_n = log_n;
        jobs[j].log_cpus = log_cpus;
        jobs[j].log_new_n = log_new_n;
        jobs[j].j = j;
        pthread_create(&th[j], NULL, sub_fft_thread, &jobs[j]);
    }
    for (uint32_t j = 0; j < num_cpus; j++) pthread_join(th[j], NULL);
    /* fft.rs:237-250 un-shuffle (chunked over threads in the reference; a memory-bound copy) */
    const size_t n = (size_t)1 << log_n;
    const size_t mask = num_cpus - 1;
    for (size_t idx = 0; idx < n; idx++) values[idx] = tmp[(idx & mask) * new_n + (idx >> log_cpus)];
    free(th);
    free(jobs);
    free(tmp);
}

static uint32_t log2_floor_u(unsigned v) {
    uint32_t l = 0;
    while (v > 1) {
        v >>= 1;
        l++;
    }
    return l;
}

static void pad(fp_t *values, size_t len_in, uint32_t log_n) {
    size_t n = (size_t)1 << log_n;
    if (len_in > n) abort(); /* serial_fft assert_eq fft.rs:162 */
    if (len_in < n) memset(values + len_in, 0, (n - len_in) * sizeof(fp_t));
}

void orc_best_fft(fp_t *values, size_t len_in, const fp_t *root, uint32_t log_n, unsigned cpus) {
    pad(values, len_in, log_n);
    size_t n = (size_t)1 << log_n;
    if (cpus <= 1 || n <= cpus)
        orc_serial_fft(values, root, log_n);
    else
        orc_parallel_fft(values, root, log_n, log2_floor_u(cpus)); /* multicore.rs:47 log_num_cpus */
}

void orc_inv_best_fft(fp_t *values, size_t len_in, const fp_t *root, uint32_t log_n, unsigned cpus) {
    pad(values, len_in, log_n);
    size_t n = (size_t)1 << log_n;
    fp_t m, inv_len, inv_root;
    fp_from_u64(&m, (uint64_t)n);
    fp_inv(&inv_len, &m);
    fp_inv(&inv_root, root);
    if (cpus <= 1 || n <= cpus)
        orc_serial_fft(values, &inv_root, log_n);
    else
        orc_parallel_fft(values, &inv_root, log_n, log2_floor_u(cpus));
    for (size_t i = 0; i < n; i++) fp_mul(&values[i], &values[i], &inv_len);
}
