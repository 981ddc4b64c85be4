/*
 * fri.c — FRI low-degree prover and verifier restating
 * /root/reference/packages/fri/src/fri.rs:14-404.
 * ORACLE / TEST INFRASTRUCTURE — see oracle.h.
 */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MIN_DEG_DIRECT_CHECKING 16 /* fri.rs:14 */

/* ---- byte buffer --------------------------------------------------------------------------- */
static void buf_reserve(orc_buf *b, size_t extra) {
    if (b->len + extra + 1 > b->cap) {
        size_t cap = b->cap ? b->cap : 1 << 16;
        while (b->len + extra + 1 > cap) cap *= 2;
        b->p = (char *)realloc(b->p, cap);
        b->cap = cap;
    }
}
static void buf_puts(orc_buf *b, const char *s) {
    size_t l = strlen(s);
    buf_reserve(b, l);
    memcpy(b->p + b->len, s, l);
    b->len += l;
    b->p[b->len] = 0;
}
static void buf_bytes(orc_buf *b, const uint8_t *d, size_t n) { /* serde: Vec<u8> -> [1,2,3] */
    buf_reserve(b, 4 * n + 2);
    char *p = b->p + b->len;
    *p++ = '[';
    for (size_t i = 0; i < n; i++) {
        if (i) *p++ = ',';
        unsigned v = d[i];
        if (v >= 100) *p++ = (char)('0' + v / 100);
        if (v >= 10) *p++ = (char)('0' + (v / 10) % 10);
        *p++ = (char)('0' + v % 10);
    }
    *p++ = ']';
    b->len = (size_t)(p - b->p);
    b->p[b->len] = 0;
}
void orc_buf_free(orc_buf *b) {
    free(b->p);
    b->p = NULL;
    b->len = b->cap = 0;
}

/* merkle_tree.rs:14-18 Proof{leaf,nodes} */
static void branch_json(orc_buf *b, const orc_branch *br) {
    buf_puts(b, "{\"leaf\":");
    buf_bytes(b, br->leaf, br->leaf_bytes);
    buf_puts(b, ",\"nodes\":[");
    for (size_t d = 0; d < br->depth; d++) {
        if (d) buf_puts(b, ",");
        buf_bytes(b, br->nodes + 32 * d, 32);
    }
    buf_puts(b, "]}");
}
static void branches_json(orc_buf *b, const orc_branch *br, size_t n) {
    buf_puts(b, "[");
    for (size_t i = 0; i < n; i++) {
        if (i) buf_puts(b, ",");
        branch_json(b, &br[i]);
    }
    buf_puts(b, "]");
}

/* fri.rs:16-26: externally tagged enum */
void orc_fri_proof_json(orc_buf *b, const orc_fri_proof *p) {
    buf_puts(b, "[");
    for (size_t i = 0; i < p->n_layers; i++) {
        const orc_fri_layer *l = &p->layers[i];
        if (i) buf_puts(b, ",");
        if (l->is_last) {
            buf_puts(b, "{\"Last\":{\"last\":[");
            for (size_t j = 0; j < l->n_last; j++) {
                if (j) buf_puts(b, ",");
                buf_bytes(b, l->last + 32 * j, 32);
            }
            buf_puts(b, "]}}");
        } else {
            buf_puts(b, "{\"Middle\":{\"root2\":");
            buf_bytes(b, l->root2, 32);
            buf_puts(b, ",\"column_branches\":");
            branches_json(b, l->column_branches, l->n_column);
            buf_puts(b, ",\"poly_branches\":");
            branches_json(b, l->poly_branches, l->n_poly);
            buf_puts(b, "}}");
        }
    }
    buf_puts(b, "]");
}

/* shared with stark.c */
void orc__buf_puts(orc_buf *b, const char *s) { buf_puts(b, s); }
void orc__buf_bytes(orc_buf *b, const uint8_t *d, size_t n) { buf_bytes(b, d, n); }
void orc__branches_json(orc_buf *b, const orc_branch *br, size_t n) { branches_json(b, br, n); }

/* build Proof{leaf,nodes} records for indices on a tree over `leaves` (gen_proofs) */
orc_branch *orc__gen_branches(const uint8_t *leaves, size_t leaf_bytes, size_t n,
                              const size_t *idx, size_t n_idx, uint8_t root[32]) {
    size_t depth = 0;
    while (((size_t)1 << depth) < n) depth++;
    uint8_t *nodes = (uint8_t *)malloc(n_idx * depth * 32 + 1);
    orc_merkle_gen_proofs(leaves, leaf_bytes, n, idx, n_idx, root, nodes);
    orc_branch *br = (orc_branch *)calloc(n_idx ? n_idx : 1, sizeof(orc_branch));
    for (size_t i = 0; i < n_idx; i++) {
        br[i].leaf_bytes = leaf_bytes;
        br[i].leaf = (uint8_t *)malloc(leaf_bytes);
        memcpy(br[i].leaf, leaves + idx[i] * leaf_bytes, leaf_bytes);
        br[i].depth = depth;
        br[i].nodes = (uint8_t *)malloc(depth * 32 + 1);
        memcpy(br[i].nodes, nodes + i * depth * 32, depth * 32);
    }
    free(nodes);
    return br;
}

void orc__branches_free(orc_branch *br, size_t n) {
    if (!br) return;
    for (size_t i = 0; i < n; i++) {
        free(br[i].leaf);
        free(br[i].nodes);
    }
    free(br);
}

/* fri.rs:88-112: the direct check of the last layer (debug_assert in the prover, assert in the
 * verifier).  Returns 1 when every remaining point lies on the interpolant. */
static int direct_check(const fp_t *values, size_t n, const fp_t *xs, size_t max_deg_plus_1, uint32_t excl) {
    size_t *pts = (size_t *)malloc(n * sizeof(size_t));
    size_t n_pts = 0;
    for (size_t x = 0; x < n; x++)
        if (excl == 0 || x % excl != 0) pts[n_pts++] = x;
    if (n_pts < max_deg_plus_1) abort(); /* split_off would panic */
    fp_t *xv = (fp_t *)malloc(max_deg_plus_1 * sizeof(fp_t));
    fp_t *yv = (fp_t *)malloc(max_deg_plus_1 * sizeof(fp_t));
    fp_t *poly = (fp_t *)malloc(max_deg_plus_1 * sizeof(fp_t));
    for (size_t i = 0; i < max_deg_plus_1; i++) {
        xv[i] = xs[pts[i]];
        yv[i] = values[pts[i]];
    }
    orc_lagrange_interp(poly, xv, yv, max_deg_plus_1);
    int ok = 1;
    for (size_t i = max_deg_plus_1; i < n_pts; i++) {
        fp_t e;
        orc_eval_poly_at(&e, poly, max_deg_plus_1, &xs[pts[i]]);
        if (!fp_eq(&e, &values[pts[i]])) ok = 0;
    }
    free(pts);
    free(xv);
    free(yv);
    free(poly);
    return ok;
}

void orc_prove_low_degree(orc_fri_proof *out, const fp_t *values_in, size_t n, const fp_t *root_in,
                          size_t max_deg_plus_1, uint32_t excl) {
    out->layers = NULL;
    out->n_layers = 0;
    fp_t root = *root_in;
    fp_t *values = (fp_t *)malloc(n * sizeof(fp_t));
    memcpy(values, values_in, n * sizeof(fp_t));
    for (;;) {
        out->layers = (orc_fri_layer *)realloc(out->layers, (out->n_layers + 1) * sizeof(orc_fri_layer));
        orc_fri_layer *L = &out->layers[out->n_layers++];
        memset(L, 0, sizeof *L);
        /* :84 xs = expand_root_of_unity(root) */
        fp_t *xs = (fp_t *)malloc(n * sizeof(fp_t));
        size_t order = orc_expand_root_of_unity(xs, n, &root);
        if (order != n) abort();
        if (max_deg_plus_1 <= MIN_DEG_DIRECT_CHECKING) { /* :88-112 */
            if (!direct_check(values, n, xs, max_deg_plus_1, excl)) {
                fprintf(stderr, "oracle: FRI direct check failed (fri.rs:105)\n");
                abort();
            }
            L->is_last = 1;
            L->n_last = n;
            L->last = (uint8_t *)malloc(n * 32);
            for (size_t i = 0; i < n; i++) fp_to_bytes_le(L->last + 32 * i, &values[i]);
            free(xs);
            break;
        }
        /* :120-131 m_tree over to_bytes_le(values) */
        uint8_t *enc = (uint8_t *)malloc(n * 32);
        for (size_t i = 0; i < n; i++) fp_to_bytes_le(enc + 32 * i, &values[i]);
        uint8_t m_root[32];
        orc_merkle_gen_proofs(enc, 32, n, NULL, 0, m_root, NULL);
        /* :135 special_x */
        fp_t special_x;
        fp_from_bytes_le(&special_x, m_root, 32);
        /* :141-164 column */
        size_t q = n / 4;
        fp_t *xsets = (fp_t *)malloc(n * sizeof(fp_t));
        fp_t *ysets = (fp_t *)malloc(n * sizeof(fp_t));
        for (size_t i = 0; i < q; i++)
            for (size_t j = 0; j < 4; j++) {
                xsets[i * 4 + j] = xs[i + q * j];
                ysets[i * 4 + j] = values[i + q * j];
            }
        fp_t *x_polys = (fp_t *)malloc(n * sizeof(fp_t));
        orc_multi_interp_4(x_polys, xsets, ysets, q);
        fp_t *column = (fp_t *)malloc(q * sizeof(fp_t));
        for (size_t i = 0; i < q; i++) orc_eval_quartic(&column[i], x_polys + 4 * i, &special_x);
        /* :165-172 m2_tree */
        uint8_t *enc_col = (uint8_t *)malloc(q * 32);
        for (size_t i = 0; i < q; i++) fp_to_bytes_le(enc_col + 32 * i, &column[i]);
        uint8_t m2_root[32];
        orc_merkle_gen_proofs(enc_col, 32, q, NULL, 0, m2_root, NULL);
        /* :181-190 */
        uint32_t ys32[40];
        orc_get_pseudorandom_indices(ys32, m2_root, 32, (uint32_t)q, 40, excl);
        size_t ys[40], poly_pos[160];
        for (int i = 0; i < 40; i++) ys[i] = ys32[i];
        uint8_t tmp_root[32];
        L->n_column = 40;
        L->column_branches = orc__gen_branches(enc_col, 32, q, ys, 40, tmp_root);
        /* :193-205 */
        for (int i = 0; i < 40; i++)
            for (int j = 0; j < 4; j++) poly_pos[i * 4 + j] = ys[i] + q * (size_t)j;
        L->n_poly = 160;
        L->poly_branches = orc__gen_branches(enc, 32, n, poly_pos, 160, tmp_root);
        memcpy(L->root2, m2_root, 32);
        /* :215-223 recurse */
        free(values);
        values = column;
        n = q;
        fp_t r2, r4;
        fp_mul(&r2, &root, &root);
        fp_mul(&r4, &r2, &r2);
        root = r4;
        max_deg_plus_1 /= 4;
        free(xs);
        free(enc);
        free(xsets);
        free(ysets);
        free(x_polys);
        free(enc_col);
    }
    free(values);
}

void orc_fri_proof_free(orc_fri_proof *p) {
    for (size_t i = 0; i < p->n_layers; i++) {
        orc_fri_layer *l = &p->layers[i];
        orc__branches_free(l->column_branches, l->n_column);
        orc__branches_free(l->poly_branches, l->n_poly);
        free(l->last);
    }
    free(p->layers);
    p->layers = NULL;
    p->n_layers = 0;
}

/* merkle_tree.rs:46-58 verify_multi_branch */
static int verify_multi_branch(const uint8_t root[32], const size_t *idx, const orc_branch *br, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (!orc_merkle_validate(root, idx[i], br[i].leaf, br[i].leaf_bytes, br[i].nodes, br[i].depth)) return 0;
    return 1;
}
int orc__verify_multi_branch(const uint8_t root[32], const size_t *idx, const orc_branch *br, size_t n) {
    return verify_multi_branch(root, idx, br, n);
}

/* fri.rs:244-404 */
int orc_verify_low_degree_proof(const uint8_t merkle_root_in[32], const fp_t *root_in,
                                const orc_fri_proof *proof, size_t max_deg_plus_1, uint32_t excl) {
    uint8_t merkle_root[32];
    memcpy(merkle_root, merkle_root_in, 32);
    fp_t root = *root_in;
    /* :253-258 */
    fp_t test_val = root;
    uint64_t rou_deg = 1;
    while (!fp_eq(&test_val, &FP_ONE)) {
        rou_deg *= 2;
        fp_mul(&test_val, &test_val, &test_val);
    }
    if (proof->n_layers == 0) return 0;
    for (size_t li = 0; li + 1 < proof->n_layers; li++) {
        const orc_fri_layer *L = &proof->layers[li];
        if (L->is_last) return 0;
        /* :262-267.  The reference computes these once before the loop; recomputing per layer
         * gives the same values because (root^4)^((deg/4)/4) == root^(deg/4). */
        fp_t qr[4];
        qr[0] = FP_ONE;
        fp_pow_u64(&qr[1], &root, rou_deg / 4);
        fp_pow_u64(&qr[2], &root, rou_deg / 2);
        fp_pow_u64(&qr[3], &root, rou_deg * 3 / 4);
        fp_t special_x;
        fp_from_bytes_le(&special_x, merkle_root, 32); /* :284 */
        uint32_t ys32[40];
        orc_get_pseudorandom_indices(ys32, L->root2, 32, (uint32_t)(rou_deg / 4), 40, excl);
        size_t ys[40], poly_pos[160];
        for (int i = 0; i < 40; i++) {
            ys[i] = ys32[i];
            for (int j = 0; j < 4; j++) poly_pos[i * 4 + j] = (size_t)j * (size_t)(rou_deg / 4) + ys[i];
        }
        if (L->n_column != 40 || L->n_poly != 160) return 0;
        if (!verify_multi_branch(L->root2, ys, L->column_branches, 40)) return 0;
        if (!verify_multi_branch(merkle_root, poly_pos, L->poly_branches, 160)) return 0;
        fp_t x_coords[160], rows[160], polys[160];
        for (int i = 0; i < 40; i++) {
            fp_t x1;
            fp_pow_u64(&x1, &root, ys[i]);
            for (int j = 0; j < 4; j++) {
                fp_mul(&x_coords[i * 4 + j], &qr[j], &x1);
                fp_from_bytes_le(&rows[i * 4 + j], L->poly_branches[i * 4 + j].leaf, L->poly_branches[i * 4 + j].leaf_bytes);
            }
        }
        orc_multi_interp_4(polys, x_coords, rows, 40);
        for (int i = 0; i < 40; i++) {
            fp_t e, c;
            orc_eval_quartic(&e, polys + 4 * i, &special_x);
            fp_from_bytes_le(&c, L->column_branches[i].leaf, L->column_branches[i].leaf_bytes);
            if (!fp_eq(&e, &c)) return 0;
        }
        memcpy(merkle_root, L->root2, 32);
        fp_t r2;
        fp_mul(&r2, &root, &root);
        fp_mul(&root, &r2, &r2);
        max_deg_plus_1 /= 4;
        rou_deg /= 4;
    }
    if (!(max_deg_plus_1 >= MIN_DEG_DIRECT_CHECKING / 2)) return 0; /* :353-356 */
    const orc_fri_layer *Last = &proof->layers[proof->n_layers - 1];
    if (!Last->is_last) return 0;
    size_t n = Last->n_last;
    if (!(n > max_deg_plus_1)) return 0;
    fp_t *dec = (fp_t *)malloc(n * sizeof(fp_t));
    for (size_t i = 0; i < n; i++) fp_from_bytes_le(&dec[i], Last->last + 32 * i, 32);
    uint8_t m_root[32];
    orc_merkle_gen_proofs(Last->last, 32, n, NULL, 0, m_root, NULL);
    int ok = memcmp(m_root, merkle_root, 32) == 0; /* :381 */
    fp_t *xs = (fp_t *)malloc(n * sizeof(fp_t));
    size_t order = orc_expand_root_of_unity(xs, n, &root);
    if (order != n) ok = 0;
    if (ok) ok = direct_check(dec, n, xs, max_deg_plus_1, excl);
    free(dec);
    free(xs);
    return ok;
}
