// ext.cu -- the LDE -> commit -> FRI chain of mk_r1cs_proof (r1cs-stark/src/prove.rs:100-124, :235-264, :324-332, :367) on
// columns kept in coset-major layout and spread over the devices of a context (sb_init_multi); see sb_ext in internal.h.
//
// What moves between devices (one process, peer access over NVLink / NVSwitch):
//   * the S coefficients of a column, once, from the device that ran its inverse transform to the others (peer copies
//     pulled on the receiving device's stream, ordered behind the owner's transform by an event);
//   * 32-byte digests: a device hashes the leaves of its cosets and stores the level-lv digest directly into the subtree
//     array of the device that owns that node range (peer stores from inside the hashing kernel);
//   * the folded FRI column of layer 0 and its leaf digests, stored directly into the primary device's memory by the fold
//     kernels; layers >= 1 hold <= N/4 values and run on the primary alone;
//   * g subtree roots and the openings, to the host.
// With one device the same code runs with cpd = 8 and a standard single-array tree.
#include "internal.h"

#include <algorithm>

int sync_all(sb_ctx *root) {
    sb_ctx *ctx = root;
    for (sb_ctx *c : root->dev) {
        DevGuard g(c);
        CU(cudaStreamSynchronize(c->stream));
    }
    return SB_OK;
}

void ext_free(sb_ext *e) {
    if (!e) return;
    for (int d = 0; d < e->g; d++) {
        sb_ctx *c = e->root->dev[d];
        DevGuard g(c);
        if (e->buf[d]) cudaFreeAsync(e->buf[d], c->stream);
        if (e->coef[d]) cudaFreeAsync(e->coef[d], c->stream);
        if (e->in[d]) cudaFreeAsync(e->in[d], c->stream);
    }
    delete e;
}

int ext_create(sb_ctx *root, size_t n_cols, size_t n_lde, uint32_t log_s, sb_ext **out) {
    sb_ctx *ctx = root;
    if (log_s + 3 > 28) return fail(ctx, SB_ERR_ARG, "extended domain 2^%u exceeds two-adicity 28", log_s + 3);
    if (log_s < 3) return fail(ctx, SB_ERR_ARG, "columns shorter than 8 values are not supported");
    sb_ext *e = new sb_ext();
    e->root = root;
    e->g = root->n_dev();
    // tiny domains gain nothing from several devices and would leave a device without a whole node range
    if (((size_t)1 << log_s) < 1024) e->g = 1;
    e->cpd = 8 / (uint32_t)e->g;
    e->lv = ilog2(e->cpd);
    e->log_s = log_s;
    e->S = (size_t)1 << log_s;
    e->N = e->S << 3;
    e->n_cols = n_cols;
    e->n_lde = n_lde < n_cols ? n_lde : n_cols;
    e->owner.resize(n_cols);
    for (size_t c = 0; c < n_cols; c++) e->owner[c] = (int)(c % (size_t)e->g);
    {   // g2 = 7^((p - 1) / N)  (prove.rs:71-82)
        uint64_t ex[4] = {hfp::PMOD[0] - 1, hfp::PMOD[1], hfp::PMOD[2], hfp::PMOD[3]};
        for (uint32_t i = 0; i < log_s + 3; i++) {
            for (int k = 0; k < 3; k++) ex[k] = (ex[k] >> 1) | (ex[k + 1] << 63);
            ex[3] >>= 1;
        }
        e->g2 = hfp::pow_limbs(hfp::from_u64(7), ex, 4);
    }
    for (int d = 0; d < e->g; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard g(c);
        const size_t nb = n_cols * e->cpd * e->S * 32, nc = e->n_lde * e->S * 32;
        cudaError_t e1 = cudaMallocAsync((void **)&e->buf[d], nb ? nb : 16, c->stream);
        cudaError_t e2 = e1 == cudaSuccess ? cudaMallocAsync((void **)&e->coef[d], nc ? nc : 16, c->stream) : e1;
        cudaError_t e3 = e2 == cudaSuccess ? cudaMallocAsync((void **)&e->in[d], nc ? nc : 16, c->stream) : e2;
        if (e3 == cudaSuccess) e3 = cudaMemsetAsync(e->in[d], 0, nc ? nc : 16, c->stream);      // zero padding of short columns
        if (e3 != cudaSuccess) {
            ext_free(e);
            return fail(ctx, SB_ERR_OOM, "device %d: allocation of the extended columns failed: %s", c->device, cudaGetErrorString(e3));
        }
    }
    *out = e;
    return SB_OK;
}

// low-degree extension of columns [first, first + count): inverse transform on the owning device, coefficients to every
// device, then each device's cosets.  Inputs are e->input(c) (S values, zero padded, canonical).  Asynchronous: the work is
// queued on the devices' streams, ordered by events; the caller synchronises (sync_all) when it needs the host to see it.
int ext_extend(sb_ext *e, size_t first, size_t count) {
    sb_ctx *root = e->root, *ctx = root;
    if (first + count > e->n_lde) return fail(ctx, SB_ERR_ARG, "column range out of bounds");
    if (count == 0) return SB_OK;
    const int g = e->g;
    const size_t S = e->S;
    const uint32_t log_n = e->log_s + 3;
    // per device: the extended domain's table (W = g2); the S-point transforms use it with stride 8
    const uint4 *tw[SB_MAX_DEV];
    uint32_t tw_log_n[SB_MAX_DEV], tw_stride[SB_MAX_DEV];
    for (int d = 0; d < g; d++) {
        DevGuard dg(root->dev[d]);
        int rc = get_table(root->dev[d], e->g2, log_n, &tw[d], &tw_log_n[d], &tw_stride[d]);
        if (rc != SB_OK) {
            if (root->dev[d] != root) fail(ctx, rc, "%s", root->dev[d]->err);
            return rc;
        }
    }
    // 1. inverse transforms on the owners: runs of columns with a constant stride go in one batch
    std::vector<cudaEvent_t> done(g, nullptr);
    for (int o = 0; o < g; o++) {
        std::vector<size_t> mine;
        for (size_t c = first; c < first + count; c++)
            if (e->owner[c] == o) mine.push_back(c);
        if (mine.empty()) continue;
        sb_ctx *c = root->dev[o];
        DevGuard dg(c);
        size_t i = 0;
        while (i < mine.size()) {
            size_t j = i + 1, stride = j < mine.size() ? mine[j] - mine[i] : 1;
            while (j < mine.size() && mine[j] - mine[j - 1] == stride) j++;
            int rc = ntt_dev_tw(c, e->in[o] + 2 * mine[i] * S, S, stride * S, e->coef[o] + 2 * mine[i] * S, stride * S, j - i, e->log_s, 1, tw[o],
                                tw_log_n[o], tw_stride[o] + 3, nullptr);
            if (rc != SB_OK) {
                if (c != root) fail(ctx, rc, "%s", c->err);
                return rc;
            }
            i = j;
        }
        if (g > 1) {
            CU(cudaEventCreateWithFlags(&done[o], cudaEventDisableTiming));
            CU(cudaEventRecord(done[o], c->stream));
        }
    }
    // 2. every other device pulls the coefficients (peer copy on its own stream, behind the owner's event)
    if (g > 1) {
        for (int d = 0; d < g; d++) {
            sb_ctx *c = root->dev[d];
            DevGuard dg(c);
            for (int o = 0; o < g; o++)
                if (o != d && done[o]) CU(cudaStreamWaitEvent(c->stream, done[o], 0));
            for (size_t col = first; col < first + count; col++) {
                const int o = e->owner[col];
                if (o == d) continue;
                CU(cudaMemcpyAsync(e->coef[d] + 2 * col * S, e->coef[o] + 2 * col * S, S * 32, cudaMemcpyDefault, c->stream));
            }
        }
        for (int o = 0; o < g; o++)
            if (done[o]) cudaEventDestroy(done[o]);      // released once the recorded work has completed
    }
    // 3. each device's cosets of every column: S-point transforms of the coefficients scaled by W^(j r)
    for (int d = 0; d < g; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard dg(c);
        CosetSpec cs;
        cs.log_ext = 3;
        cs.store = NTT_STORE_PLAIN;
        cs.dst_cpd = e->cpd;
        cs.dst_r0 = (uint32_t)d * e->cpd;
        if (g == 1) {
            // coset 0 is the input column itself (same field elements): a strided copy instead of a transform
            cs.r0 = 1;
            cs.cnt = 7;
            CU(cudaMemcpy2DAsync(e->col(0, first), 8 * S * 32, e->in[0] + 2 * first * S, S * 32, S * 32, count, cudaMemcpyDeviceToDevice, c->stream));
        } else {
            cs.r0 = (uint32_t)d * e->cpd;
            cs.cnt = e->cpd;
        }
        int rc = ntt_dev_tw(c, e->coef[d] + 2 * first * S, S, S, e->col(d, first), S, count, e->log_s, 0, tw[d], tw_log_n[d], tw_stride[d] + 3, &cs);
        if (rc != SB_OK) {
            if (c != root) fail(ctx, rc, "%s", c->err);
            return rc;
        }
    }
    return SB_OK;
}

// levels above level 0 of a standard tree array of n_leaves leaf digests
static int nodes_build(sb_ctx *ctx, uint4 *nodes, size_t n_leaves) {
    const uint32_t depth = ilog2(n_leaves);
    uint32_t level = 0;
    while (level < depth) {
        const uint32_t lv = depth - level < 3 ? depth - level : 3;
        KLAUNCH(SB_KIND_MERKLE_NODES, merkle_launch_nodes(ctx->stream, lv, nodes, n_leaves, level));
        level += lv;
    }
    CU(cudaGetLastError());
    return SB_OK;
}

// MerkleProofInPlace over the rows of the given columns (prove.rs:235-264 / :324-332): leaf 8 k + r =
// to_bytes_le(col_0[r][k]) || ... .  Synchronous (the root is on the host when it returns).
int ext_commit(const sb_ext *e, const size_t *col_ids, size_t n_ids, sb_tree **tree) {
    sb_ctx *root = e->root, *ctx = root;
    if (n_ids < 1 || n_ids > 8) return fail(ctx, SB_ERR_ARG, "1..8 columns per leaf supported, got %zu", n_ids);
    for (size_t i = 0; i < n_ids; i++)
        if (col_ids[i] >= e->n_cols) return fail(ctx, SB_ERR_ARG, "column %zu out of range", col_ids[i]);
    const int g = e->g;
    const size_t S = e->S;
    ExtLeavesParams P;
    memset(&P, 0, sizeof P);
    P.nc = (uint32_t)n_ids;
    P.log_s = e->log_s;
    P.cpd = e->cpd;
    P.lv = e->lv;
    P.g = (uint32_t)g;
    if (g == 1) {
        sb_tree *t = nullptr;
        TRY(tree_new(ctx, e->N, 32 * n_ids, &t));
        t->n_cols = (int)n_ids;
        t->coset_log_s = e->log_s;
        for (size_t i = 0; i < n_ids; i++) t->cols[i] = P.cols[i] = e->col(0, col_ids[i]);
        P.single = t->d_nodes;
        KLAUNCH(SB_KIND_MERKLE_LEAVES, merkle_launch_leaves_ext(ctx->stream, P));
        int rc = merkle_finish(ctx, t, 3, true);
        if (rc != SB_OK) {
            free_tree(t);
            return rc;
        }
        *tree = t;
        return SB_OK;
    }
    sb_tree *t = new sb_tree();
    t->n = e->N;
    t->depth = e->log_s + 3;
    t->leaf_bytes = 32 * n_ids;
    t->n_cols = (int)n_ids;
    t->stream = root->stream;
    t->device = root->device;
    TreeShards *sh = t->sh = new TreeShards();
    sh->g = g;
    sh->lv = e->lv;
    sh->log_s = e->log_s;
    sh->cpd = e->cpd;
    const size_t low_digests = ext_low_off(e->log_s, e->cpd, e->lv);
    cudaEvent_t hashed[SB_MAX_DEV] = {0};
    int rc = SB_OK;
    for (int d = 0; d < g && rc == SB_OK; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard dg(c);
        sh->streams[d] = c->stream;
        sh->devices[d] = c->device;
        for (size_t i = 0; i < n_ids; i++) sh->cols[d][i] = e->col(d, col_ids[i]);
        cudaError_t e1 = low_digests ? cudaMallocAsync((void **)&sh->low[d], low_digests * 32, c->stream) : cudaSuccess;
        cudaError_t e2 = e1 == cudaSuccess ? cudaMallocAsync((void **)&sh->sub[d], (2 * S - 1) * 32, c->stream) : e1;
        if (e2 != cudaSuccess) rc = fail(ctx, SB_ERR_OOM, "device %d: tree allocation failed: %s", c->device, cudaGetErrorString(e2));
    }
    // every subtree array must exist before a peer stores into it
    cudaEvent_t ready[SB_MAX_DEV] = {0};
    for (int d = 0; d < g && rc == SB_OK; d++) {
        DevGuard dg(root->dev[d]);
        if (cudaEventCreateWithFlags(&ready[d], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(ready[d], root->dev[d]->stream) != cudaSuccess)
            rc = fail(ctx, SB_ERR_CUDA, "event: %s", cudaGetErrorString(cudaGetLastError()));
    }
    for (int d = 0; d < g && rc == SB_OK; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard dg(c);
        for (int o = 0; o < g; o++)
            if (o != d) cudaStreamWaitEvent(c->stream, ready[o], 0);
        ExtLeavesParams Q = P;
        Q.d = (uint32_t)d;
        for (size_t i = 0; i < n_ids; i++) Q.cols[i] = sh->cols[d][i];
        Q.low = sh->low[d];
        for (int o = 0; o < g; o++) Q.sub[o] = sh->sub[o];
        {
            sb_ctx *ctx = c;       // launch accounting on the device that runs the kernel
            KLAUNCH(SB_KIND_MERKLE_LEAVES, merkle_launch_leaves_ext(c->stream, Q));
        }
        if (cudaEventCreateWithFlags(&hashed[d], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(hashed[d], c->stream) != cudaSuccess)
            rc = fail(ctx, SB_ERR_CUDA, "event: %s", cudaGetErrorString(cudaGetLastError()));
    }
    // subtrees: a device waits for every device's leaf kernel (they all store into its level 0), then reduces
    std::vector<uint8_t> roots((size_t)g * 32);
    for (int d = 0; d < g && rc == SB_OK; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard dg(c);
        for (int o = 0; o < g; o++)
            if (o != d) cudaStreamWaitEvent(c->stream, hashed[o], 0);
        rc = nodes_build(c, sh->sub[d], S);
        if (rc != SB_OK && c != root) fail(ctx, rc, "%s", c->err);
        if (rc == SB_OK && cudaMemcpyAsync(&roots[32 * d], (const uint8_t *)sh->sub[d] + (2 * S - 2) * 32, 32, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
            rc = fail(ctx, SB_ERR_CUDA, "D2H subtree root: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc == SB_OK) rc = sync_all(root);
    for (int d = 0; d < g; d++) {
        if (ready[d]) cudaEventDestroy(ready[d]);
        if (hashed[d]) cudaEventDestroy(hashed[d]);
    }
    if (rc != SB_OK) {
        free_tree(t);
        return rc;
    }
    // top of the tree on the host: g subtree roots -> root (merkle_proof_in_place.rs:78-98: parent = H(left || right))
    sh->top.assign((size_t)(2 * g - 1) * 32, 0);
    memcpy(sh->top.data(), roots.data(), (size_t)g * 32);
    for (uint32_t l = 0; ((size_t)g >> l) > 1; l++) {
        const size_t off = merkle_level_off((size_t)g, l), offn = merkle_level_off((size_t)g, l + 1);
        for (size_t i = 0; i < ((size_t)g >> (l + 1)); i++) b2s::hash_bytes(&sh->top[(offn + i) * 32], &sh->top[(off + 2 * i) * 32], 64);
    }
    memcpy(t->root, &sh->top[(size_t)(2 * g - 2) * 32], 32);
    *tree = t;
    return SB_OK;
}

static void ext_open_params(const sb_ext *e, size_t col, ExtOpenParams &P) {
    memset(&P, 0, sizeof P);
    for (int d = 0; d < e->g; d++) P.cols[d][0] = e->col(d, col);
    P.nc = 1; P.log_s = e->log_s; P.cpd = e->cpd; P.lv = e->lv; P.g = (uint32_t)e->g;
}

// one column in natural order on the primary device (d_out: N elements)
int ext_to_natural(const sb_ext *e, size_t col, uint4 *d_out) {
    sb_ctx *ctx = e->root;
    ExtOpenParams P;
    ext_open_params(e, col, P);
    KLAUNCH(SB_KIND_OTHER, ext_launch_to_natural(ctx->stream, P, d_out));
    CU(cudaGetLastError());
    return SB_OK;
}

// prove_low_degree (fri/src/fri.rs:46-224) on one coset-major column.  Layer 0 (fri.rs:120-213) runs where the values
// are: every device folds the rows of its cosets (the four values of a row share a coset: a quarter turn is S/4 steps of
// 8), hashes the folded leaves and stores column + digests straight into the primary device's memory; the primary
// finishes the column tree, the openings of the (sharded) values tree are gathered through peer pointers, and layers >= 1
// run on the primary alone with the column tree already built.
int ext_fri_prove(const sb_ext *e, size_t col, const sb_tree *values_tree, size_t max_deg_plus_1, uint32_t excl, sb_fri_proof **out) {
    sb_ctx *root = e->root, *ctx = root;
    if (col >= e->n_cols) return fail(ctx, SB_ERR_ARG, "column %zu out of range", col);
    const size_t N = e->N, q = N / 4;
    const int g = e->g;
    if (max_deg_plus_1 <= FRI_MIN_DEG_DIRECT) {
        // fri.rs:88-112: the values themselves are the proof; gather them in natural order
        DevBuf nat(ctx);
        TRY(nat.alloc(N * 32));
        TRY(sync_all(root));
        TRY(ext_to_natural(e, col, (uint4 *)nat.p));
        return fri_prove_dev(ctx, (const uint4 *)nat.p, N, e->g2, max_deg_plus_1, excl, nullptr, out);
    }
    if (q >= (1u << 24) && !(ctx->extended_domain && q <= (1u << 28))) return fail(ctx, SB_ERR_ARG, "FRI layer of %zu values unsupported", N);
    sb_tree *own_tree = nullptr;
    if (!values_tree) {
        size_t id = col;
        TRY(ext_commit(e, &id, 1, &own_tree));
        values_tree = own_tree;
    }
    struct Guard {
        sb_tree *a = nullptr, *b = nullptr;
        void *col = nullptr;
        cudaStream_t s;
        ~Guard() {
            free_tree(a);
            free_tree(b);
            if (col) cudaFreeAsync(col, s);
        }
    } guard;
    guard.a = own_tree;
    guard.s = ctx->stream;
    FriLayer L;
    memcpy(L.values_root, values_tree->root, 32);
    const hfp::el special_x = hfp::from_bytes_le32(values_tree->root);        // fri.rs:135
    CU(cudaMallocAsync(&guard.col, q * 32, ctx->stream));
    sb_tree *t2 = nullptr;
    TRY(tree_new(ctx, q, 32, &t2));
    guard.b = t2;
    t2->n_cols = 1;
    t2->cols[0] = (const uint4 *)guard.col;
    cudaEvent_t ready = nullptr, folded[SB_MAX_DEV] = {0};
    if (g > 1) {
        CU(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        CU(cudaEventRecord(ready, ctx->stream));
    }
    int rc = SB_OK;
    for (int d = 0; d < g && rc == SB_OK; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard dg(c);
        const uint4 *tw;
        uint32_t tw_log_n, tw_stride;
        rc = get_table(c, e->g2, e->log_s + 3, &tw, &tw_log_n, &tw_stride);
        if (rc != SB_OK) {
            if (c != root) fail(ctx, rc, "%s", c->err);
            break;
        }
        if (d > 0) cudaStreamWaitEvent(c->stream, ready, 0);
        FriFoldParams F;
        memset(&F, 0, sizeof F);
        F.vals = e->col(d, col);
        F.col = (uint4 *)guard.col;
        F.tw = tw;
        F.n = N;
        F.tw_log_n = tw_log_n;
        F.tw_log_stride = tw_stride;
        memcpy(F.special_x, special_x.l, 32);
        F.log_s = e->log_s; F.cpd = e->cpd; F.lv = e->lv; F.d = (uint32_t)d; F.g = (uint32_t)g;
        {
            sb_ctx *ctx = c;
            KLAUNCH(SB_KIND_FRI_FOLD, merkle_launch_leaves_fold_ext(c->stream, F, t2->d_nodes));
        }
        if (d > 0) {
            if (cudaEventCreateWithFlags(&folded[d], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(folded[d], c->stream) != cudaSuccess)
                rc = fail(ctx, SB_ERR_CUDA, "event: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    for (int d = 1; d < g; d++)
        if (folded[d]) {
            cudaStreamWaitEvent(ctx->stream, folded[d], 0);
            cudaEventDestroy(folded[d]);
        }
    if (ready) cudaEventDestroy(ready);
    if (rc != SB_OK) return rc;
    TRY(merkle_finish(ctx, t2, e->lv, true));
    memcpy(L.root2, t2->root, 32);
    // fri.rs:181-204
    uint32_t ys[FRI_QUERIES];
    if (pseudorandom_indices(t2->root, 32, (uint32_t)q, FRI_QUERIES, excl, ys, ctx->extended_domain) != SB_OK)
        return fail(ctx, SB_ERR_ARG, "sampler: column length %zu out of range", q);
    std::vector<size_t> yi(FRI_QUERIES), pp(4 * FRI_QUERIES);
    for (size_t i = 0; i < FRI_QUERIES; i++) {
        yi[i] = ys[i];
        for (size_t j = 0; j < 4; j++) pp[4 * i + j] = ys[i] + q * j;
    }
    L.n_column = FRI_QUERIES;
    L.depth_column = t2->depth;
    L.column_leaves.resize(FRI_QUERIES * 32);
    L.column_nodes.resize(FRI_QUERIES * t2->depth * 32);
    TRY(sb_merkle_open(ctx, t2, yi.data(), FRI_QUERIES, L.column_leaves.data(), L.column_nodes.data()));
    L.n_poly = 4 * FRI_QUERIES;
    L.depth_poly = values_tree->depth;
    L.poly_leaves.resize(L.n_poly * 32);
    L.poly_nodes.resize(L.n_poly * values_tree->depth * 32);
    TRY(sb_merkle_open(ctx, values_tree, pp.data(), L.n_poly, L.poly_leaves.data(), L.poly_nodes.data()));
    // fri.rs:215-223: the rest on the primary device, natural order, column tree already committed
    sb_fri_proof *rest = nullptr;
    const hfp::el w4 = hfp::sqr(hfp::sqr(e->g2));
    TRY(fri_prove_dev(ctx, (const uint4 *)guard.col, q, w4, max_deg_plus_1 / 4, excl, t2, &rest));
    rest->layers.insert(rest->layers.begin(), std::move(L));
    *out = rest;
    return SB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------------
extern "C" int sb_ext_create(sb_ctx *ctx, size_t n_cols, uint32_t log_s, uint32_t log_ext, sb_ext **out) {
    return guarded(ctx, [&]() -> int {
        if (!ctx || !out) return SB_ERR_ARG;
        if (log_ext != 3) return fail(ctx, SB_ERR_ARG, "the coset-major layout is built for the reference's extension factor 8 (utils.rs:134)");
        if (n_cols == 0 || n_cols > 64) return fail(ctx, SB_ERR_ARG, "1..64 columns");
        TRY(ext_create(ctx, n_cols, n_cols, log_s, out));
        return sync_all(ctx);
    });
}
extern "C" void sb_ext_free(sb_ctx *ctx, sb_ext *e) {
    DevGuard g(ctx);
    ext_free(e);
}
extern "C" int sb_ext_devices(const sb_ext *e) { return e ? e->g : 0; }
extern "C" int sb_ext_load(sb_ctx *ctx, sb_ext *e, size_t first, size_t count, const uint64_t *cols, size_t col_len) {
    return guarded(ctx, [&]() -> int {
        if (!ctx || !e || (!cols && count * col_len)) return SB_ERR_ARG;
        if (first + count > e->n_lde) return fail(ctx, SB_ERR_ARG, "column range out of bounds");
        if (col_len > e->S) return fail(ctx, SB_ERR_ARG, "column of %zu elements does not fit 2^%u", col_len, e->log_s);
        for (size_t c = first; c < first + count; c++) {
            sb_ctx *o = ctx->dev[e->owner[c]];
            DevGuard dg(o);
            if (col_len < e->S) CU(cudaMemsetAsync(e->input(c) + 2 * col_len, 0, (e->S - col_len) * 32, o->stream));
            CU(cudaMemcpyAsync(e->input(c), cols + 4 * (c - first) * col_len, col_len * 32, cudaMemcpyHostToDevice, o->stream));
        }
        return sync_all(ctx);
    });
}
extern "C" int sb_ext_extend(sb_ctx *ctx, sb_ext *e, size_t first, size_t count) {
    return guarded(ctx, [&]() -> int {
        if (!ctx || !e) return SB_ERR_ARG;
        TRY(ext_extend(e, first, count));
        return sync_all(ctx);
    });
}
extern "C" int sb_ext_commit(sb_ctx *ctx, const sb_ext *e, const size_t *col_ids, size_t n_ids, uint8_t root[32], sb_tree **tree) {
    return guarded(ctx, [&]() -> int {
        if (!ctx || !e || !col_ids || !tree) return SB_ERR_ARG;
        TRY(ext_commit(e, col_ids, n_ids, tree));
        if (root) memcpy(root, (*tree)->root, 32);
        return SB_OK;
    });
}
extern "C" int sb_ext_fri_prove(sb_ctx *ctx, const sb_ext *e, size_t col, const sb_tree *values_tree, size_t max_deg_plus_1, uint32_t excl,
                                sb_fri_proof **out) {
    return guarded(ctx, [&]() -> int {
        if (!ctx || !e || !out) return SB_ERR_ARG;
        if (values_tree && (values_tree->n != e->N || values_tree->leaf_bytes != 32)) return fail(ctx, SB_ERR_ARG, "values_tree does not match the column");
        return ext_fri_prove(e, col, values_tree, max_deg_plus_1, excl, out);
    });
}
extern "C" int sb_ext_read(sb_ctx *ctx, const sb_ext *e, size_t col, uint64_t *out) {
    return guarded(ctx, [&]() -> int {
        if (!ctx || !e || !out) return SB_ERR_ARG;
        if (col >= e->n_cols) return fail(ctx, SB_ERR_ARG, "column %zu out of range", col);
        DevBuf nat(ctx);
        TRY(nat.alloc(e->N * 32));
        TRY(sync_all(ctx));
        TRY(ext_to_natural(e, col, (uint4 *)nat.p));
        CU(cudaMemcpyAsync(out, nat.p, e->N * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return SB_OK;
    });
}
