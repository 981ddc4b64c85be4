/*
 * poly.c — restates the parts of /root/reference/packages/fri/src/poly_utils.rs that the prover
 * and the FRI fold use, and the index sampler of fri/src/utils.rs:82-109.
 * ORACLE / TEST INFRASTRUCTURE — see oracle.h.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* poly_utils.rs:38-70 — batch inverse where zeros pass through as zeros */
void orc_multi_inv(fp_t *out, const fp_t *values, size_t n) {
    fp_t *partials = (fp_t *)malloc((n + 1) * sizeof(fp_t));
    partials[0] = FP_ONE;
    for (size_t i = 0; i < n; i++) {
        const fp_t *f = fp_is_zero(&values[i]) ? &FP_ONE : &values[i];
        fp_mul(&partials[i + 1], &partials[i], f);
    }
    fp_t inv;
    fp_inv(&inv, &partials[n]);
    for (size_t i = n; i-- > 0;) {
        if (!fp_is_zero(&values[i])) {
            fp_t o;
            fp_mul(&o, &partials[i], &inv);
            fp_mul(&inv, &inv, &values[i]);
            out[i] = o;
        } else {
            out[i] = FP_ZERO;
        }
    }
    free(partials);
}

/* poly_utils.rs:93-102 */
void orc_eval_poly_at(fp_t *r, const fp_t *poly, size_t n, const fp_t *x) {
    fp_t y = FP_ZERO, pw = FP_ONE, t;
    for (size_t i = 0; i < n; i++) {
        fp_mul(&t, &pw, &poly[i]);
        fp_add(&y, &y, &t);
        fp_mul(&pw, &pw, x);
    }
    *r = y;
}

/* poly_utils.rs:362-373 — coefficients (low degree first) of prod (X - xs[i]) */
void orc_zpoly(fp_t *out, const fp_t *xs, size_t n) {
    /* root is built highest-degree-first then reversed */
    fp_t *root = (fp_t *)malloc((n + 1) * sizeof(fp_t));
    size_t len = 1;
    root[0] = FP_ONE;
    for (size_t i = 0; i < n; i++) {
        root[len++] = FP_ZERO;
        for (size_t j = i + 1; j-- > 0;) {
            fp_t t;
            fp_mul(&t, &root[j], &xs[i]);
            fp_sub(&root[j + 1], &root[j + 1], &t);
        }
    }
    for (size_t i = 0; i <= n; i++) out[i] = root[n - i];
    free(root);
}

/* poly_utils.rs:235-262 specialised to a monic linear divisor [ -x, 1 ] (the only way
 * lagrange_interp :419 calls it): synthetic division, quotient has n coefficients */
static void div_by_linear(fp_t *q, const fp_t *a, size_t a_len, const fp_t *x) {
    /* a has a_len = n+1 coefficients; q[n-1] = a[n]; q[d] = a[d+1] + x*q[d+1] */
    size_t n = a_len - 1;
    fp_t carry = a[n];
    q[n - 1] = carry;
    for (size_t d = n - 1; d-- > 0;) {
        fp_t t;
        fp_mul(&t, &carry, x);
        fp_add(&carry, &a[d + 1], &t);
        q[d] = carry;
    }
}

/* poly_utils.rs:409-439 */
void orc_lagrange_interp(fp_t *out, const fp_t *xs, const fp_t *ys, size_t n) {
    fp_t *root = (fp_t *)malloc((n + 1) * sizeof(fp_t));
    fp_t *nums = (fp_t *)malloc(n * n * sizeof(fp_t));
    fp_t *denoms = (fp_t *)malloc(n * sizeof(fp_t));
    fp_t *inv_denoms = (fp_t *)malloc(n * sizeof(fp_t));
    orc_zpoly(root, xs, n);
    for (size_t i = 0; i < n; i++) {
        div_by_linear(nums + i * n, root, n + 1, &xs[i]);
        orc_eval_poly_at(&denoms[i], nums + i * n, n, &xs[i]);
    }
    orc_multi_inv(inv_denoms, denoms, n);
    for (size_t j = 0; j < n; j++) out[j] = FP_ZERO;
    for (size_t i = 0; i < n; i++) {
        fp_t yslice;
        fp_mul(&yslice, &ys[i], &inv_denoms[i]);
        for (size_t j = 0; j < n; j++) {
            if (!fp_is_zero(&nums[i * n + j]) && !fp_is_zero(&ys[i])) {
                fp_t t;
                fp_mul(&t, &nums[i * n + j], &yslice);
                fp_add(&out[j], &out[j], &t);
            }
        }
    }
    free(root);
    free(nums);
    free(denoms);
    free(inv_denoms);
}

/* poly_utils.rs:442-446 */
void orc_eval_quartic(fp_t *r, const fp_t p[4], const fp_t *x) {
    fp_t xsq, xcb, t, acc;
    fp_mul(&xsq, x, x);
    fp_mul(&xcb, &xsq, x);
    acc = p[0];
    fp_mul(&t, &p[1], x);
    fp_add(&acc, &acc, &t);
    fp_mul(&t, &p[2], &xsq);
    fp_add(&acc, &acc, &t);
    fp_mul(&t, &p[3], &xcb);
    fp_add(&acc, &acc, &t);
    *r = acc;
}

/* poly_utils.rs:449-511 */
void orc_multi_interp_4(fp_t *out, const fp_t *xsets, const fp_t *ysets, size_t rows) {
    fp_t *eqs = (fp_t *)malloc(rows * 16 * sizeof(fp_t));
    fp_t *inv_targets = (fp_t *)malloc(rows * 4 * sizeof(fp_t));
    fp_t *inv_alls = (fp_t *)malloc(rows * 4 * sizeof(fp_t));
    for (size_t key = 0; key < rows; key++) {
        const fp_t *xs = xsets + key * 4;
        fp_t x01, x02, x03, x12, x13, x23, t;
        fp_mul(&x01, &xs[0], &xs[1]);
        fp_mul(&x02, &xs[0], &xs[2]);
        fp_mul(&x03, &xs[0], &xs[3]);
        fp_mul(&x12, &xs[1], &xs[2]);
        fp_mul(&x13, &xs[1], &xs[3]);
        fp_mul(&x23, &xs[2], &xs[3]);
        fp_t *eq0 = eqs + key * 16, *eq1 = eq0 + 4, *eq2 = eq0 + 8, *eq3 = eq0 + 12;
        /* eq0 */
        fp_mul(&t, &x12, &xs[3]);
        fp_neg(&eq0[0], &t);
        fp_add(&t, &x12, &x13);
        fp_add(&eq0[1], &t, &x23);
        fp_sub(&t, &FP_ZERO, &xs[1]);
        fp_sub(&t, &t, &xs[2]);
        fp_sub(&eq0[2], &t, &xs[3]);
        eq0[3] = FP_ONE;
        /* eq1 */
        fp_mul(&t, &x02, &xs[3]);
        fp_neg(&eq1[0], &t);
        fp_add(&t, &x02, &x03);
        fp_add(&eq1[1], &t, &x23);
        fp_sub(&t, &FP_ZERO, &xs[0]);
        fp_sub(&t, &t, &xs[2]);
        fp_sub(&eq1[2], &t, &xs[3]);
        eq1[3] = FP_ONE;
        /* eq2 */
        fp_mul(&t, &x01, &xs[3]);
        fp_neg(&eq2[0], &t);
        fp_add(&t, &x01, &x03);
        fp_add(&eq2[1], &t, &x13);
        fp_sub(&t, &FP_ZERO, &xs[0]);
        fp_sub(&t, &t, &xs[1]);
        fp_sub(&eq2[2], &t, &xs[3]);
        eq2[3] = FP_ONE;
        /* eq3 */
        fp_mul(&t, &x01, &xs[2]);
        fp_neg(&eq3[0], &t);
        fp_add(&t, &x01, &x02);
        fp_add(&eq3[1], &t, &x12);
        fp_sub(&t, &FP_ZERO, &xs[0]);
        fp_sub(&t, &t, &xs[1]);
        fp_sub(&eq3[2], &t, &xs[2]);
        eq3[3] = FP_ONE;
        orc_eval_quartic(&inv_targets[key * 4 + 0], eq0, &xs[0]);
        orc_eval_quartic(&inv_targets[key * 4 + 1], eq1, &xs[1]);
        orc_eval_quartic(&inv_targets[key * 4 + 2], eq2, &xs[2]);
        orc_eval_quartic(&inv_targets[key * 4 + 3], eq3, &xs[3]);
    }
    orc_multi_inv(inv_alls, inv_targets, rows * 4);
    for (size_t i = 0; i < rows; i++) {
        const fp_t *ys = ysets + i * 4;
        const fp_t *eq = eqs + i * 16;
        fp_t inv_y[4];
        for (int j = 0; j < 4; j++) fp_mul(&inv_y[j], &ys[j], &inv_alls[i * 4 + j]);
        for (int c = 0; c < 4; c++) {
            fp_t acc = FP_ZERO, t;
            for (int j = 0; j < 4; j++) {
                fp_mul(&t, &eq[j * 4 + c], &inv_y[j]);
                fp_add(&acc, &acc, &t);
            }
            out[i * 4 + c] = acc;
        }
    }
    free(eqs);
    free(inv_targets);
    free(inv_alls);
}

/* fri/src/utils.rs:82-109 (identical copy in commitment/src/utils.rs:82-109) */
void orc_get_pseudorandom_indices(uint32_t *out, const uint8_t *seed, size_t seed_len,
                                  uint32_t modulus, size_t count, uint32_t excl) {
    if (!(modulus < (1u << 24))) abort(); /* utils.rs:88 assert */
    size_t cap = seed_len + 4 * count + 64;
    uint8_t *data = (uint8_t *)malloc(cap);
    size_t len = seed_len;
    memcpy(data, seed, seed_len);
    while (len < 4 * count) {
        orc_blake2s(data + len, data + len - 32, 32);
        len += 32;
    }
    uint32_t real_modulus = excl ? modulus * (excl - 1) / excl : modulus;
    for (size_t i = 0; i < count; i++) {
        const uint8_t *d = data + 4 * i;
        uint32_t v = ((uint32_t)d[0] << 24) | ((uint32_t)d[1] << 16) | ((uint32_t)d[2] << 8) | d[3];
        if (excl == 0) {
            out[i] = v % modulus;
        } else {
            uint32_t t = v % real_modulus;
            out[i] = t + 1 + t / (excl - 1);
        }
    }
    free(data);
}
