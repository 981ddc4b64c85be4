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

// SB_TRACE=1: phase times of the sharded tree builder on stderr (every device waited for between phases)
static bool trace_on() {
    static const bool on = getenv("SB_TRACE") != nullptr;
    return on;
}
static double trace_now() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
int sync_all(sb_ctx *root);
struct TracePhase {
    sb_ctx *root;
    double t0;
    explicit TracePhase(sb_ctx *r) : root(r), t0(0) {
        if (trace_on()) {
            sync_all(root);
            t0 = trace_now();
        }
    }
    void lap(const char *what) {
        if (!trace_on()) return;
        sync_all(root);
        const double t = trace_now();
        fprintf(stderr, "[sb_trace] %-28s %8.3f ms\n", what, t - t0);
        t0 = t;
    }
};

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
        blk_release(c, e->buf[d]);
        blk_release(c, e->coef[d]);
        blk_release(c, e->in[d]);
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
        int rc = blk_alloc(c, nb, (void **)&e->buf[d]);
        if (rc == SB_OK) rc = blk_alloc(c, nc, (void **)&e->coef[d]);
        if (rc == SB_OK) rc = blk_alloc(c, nc, (void **)&e->in[d]);
        if (rc == SB_OK && cudaMemsetAsync(e->in[d], 0, nc ? nc : 16, c->stream) != cudaSuccess)      // zero padding of short columns
            rc = fail(c, SB_ERR_CUDA, "memset: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc != SB_OK) {
            if (c != root) fail(ctx, rc, "%s", c->err);
            ext_free(e);
            return rc;
        }
    }
    *out = e;
    return SB_OK;
}

// low-degree extension of columns [first, first + count): inverse transform on the owning device, coefficients to every
// device, then each device's cosets.  Inputs are e->input(c) (S values, zero padded, canonical).  Asynchronous: the work is
// queued on the devices' streams, ordered by events; the caller synchronises (sync_all) when it needs the host to see it.
int ext_extend(sb_ext *e, size_t first, size_t count) {
    NvtxRange nvtx("ext_extend: INTT + coefficient all-gather + coset transforms");
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
    // 2. + 3. per device: the cosets of its own columns first, then -- in ring order, so that every source is read by one
    // device at a time -- the columns of owner d+1, d+2, ...: their coefficients are pulled by peer copies on the device's
    // copy stream (behind the owner's event) while the transforms of the previous group run on the compute stream.
    auto transform = [&](int d, const std::vector<size_t> &cols) -> int {
        // cosets of the given columns (a run with constant stride) on device d: S-point transforms of the coefficients scaled by W^(j r)
        sb_ctx *c = root->dev[d];
        size_t i = 0;
        while (i < cols.size()) {
            size_t j = i + 1, stride = j < cols.size() ? cols[j] - cols[i] : 1;
            while (j < cols.size() && cols[j] - cols[j - 1] == stride) j++;
            CosetSpec cs;
            cs.log_ext = 3;
            cs.store = NTT_STORE_PLAIN;
            cs.dst_cpd = e->cpd * (uint32_t)stride;           // consecutive polynomials of the batch are `stride` columns apart
            cs.dst_r0 = (uint32_t)d * e->cpd;
            if (g == 1) {
                // coset 0 is the input column itself (same field elements): a strided copy instead of a transform
                cs.r0 = 1;
                cs.cnt = 7;
                CU(cudaMemcpy2DAsync(e->col(0, cols[i]), stride * 8 * S * 32, e->in[0] + 2 * cols[i] * S, stride * S * 32, S * 32, j - i,
                                     cudaMemcpyDeviceToDevice, c->stream));
            } else {
                cs.r0 = (uint32_t)d * e->cpd;
                cs.cnt = e->cpd;
            }
            int rc = ntt_dev_tw(c, e->coef[d] + 2 * cols[i] * S, S, stride * S, e->col(d, cols[i]), S, j - i, e->log_s, 0, tw[d], tw_log_n[d],
                                tw_stride[d] + 3, &cs);
            if (rc != SB_OK) {
                if (c != root) fail(ctx, rc, "%s", c->err);
                return rc;
            }
            i = j;
        }
        return SB_OK;
    };
    std::vector<cudaEvent_t> pulled;
    int rc = SB_OK;
    for (int d = 0; d < g && rc == SB_OK; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard dg(c);
        if (g > 1 && !c->h2d_stream) CU(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
        for (int k = 0; k < g && rc == SB_OK; k++) {
            const int o = (d + k) % g;
            std::vector<size_t> cols;
            for (size_t col = first; col < first + count; col++)
                if (e->owner[col] == o) cols.push_back(col);
            if (cols.empty()) continue;
            if (o != d) {
                CU(cudaStreamWaitEvent(c->h2d_stream, done[o], 0));
                for (size_t col : cols) CU(cudaMemcpyAsync(e->coef[d] + 2 * col * S, e->coef[o] + 2 * col * S, S * 32, cudaMemcpyDefault, c->h2d_stream));
                cudaEvent_t ev;
                CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                pulled.push_back(ev);
                CU(cudaEventRecord(ev, c->h2d_stream));
                CU(cudaStreamWaitEvent(c->stream, ev, 0));
            }
            rc = transform(d, cols);
        }
    }
    for (int o = 0; o < g; o++)
        if (done[o]) cudaEventDestroy(done[o]);      // released once the recorded work has completed
    for (auto ev : pulled) cudaEventDestroy(ev);
    return rc;
}

// levels above `level` of a standard tree array of n_leaves leaf digests
static int nodes_build(sb_ctx *ctx, uint4 *nodes, size_t n_leaves, uint32_t level) {
    const uint32_t depth = ilog2(n_leaves);
    while (level < depth) {
        const uint32_t lv = depth - level < 3 ? depth - level : 3;
        KLAUNCH(SB_KIND_MERKLE_NODES, merkle_launch_nodes(ctx->stream, lv, nodes, n_leaves, level));
        level += lv;
    }
    CU(cudaGetLastError());
    return SB_OK;
}

// ---- trees whose levels are spread over the devices (TreeShards) -----------------------------------------------------
// shards_begin allocates a device's low levels and subtree array on every device and makes every device's stream wait
// for all of those allocations (peers store into them); the caller then launches its leaf kernels, records one event per
// device, and shards_finish builds the subtrees, fetches their roots and finishes the top on the host.
struct ShardGeom {
    int g;
    uint32_t log_s, cpd, lv;       // leaves = 8 << log_s; device d hashes cosets d * cpd .. of 2^log_s leaves each
};
static int shards_begin(sb_ctx *root, const ShardGeom &G, size_t leaf_bytes, int n_cols, sb_tree **out) {
    sb_ctx *ctx = root;
    TracePhase tp(root);
    struct LapAtExit {
        TracePhase &t;
        ~LapAtExit() { t.lap("shards_begin (allocations)"); }
    } lap_at_exit{tp};
    const size_t S = (size_t)1 << G.log_s;
    sb_tree *t = new sb_tree();
    t->n = S << 3;
    t->depth = G.log_s + 3;
    t->leaf_bytes = leaf_bytes;
    t->n_cols = n_cols;
    t->stream = root->stream;
    t->device = root->device;
    t->owner = root;
    TreeShards *sh = t->sh = new TreeShards();
    sh->g = G.g;
    sh->lv = G.lv;
    sh->log_s = G.log_s;
    sh->cpd = G.cpd;
    const size_t low_digests = ext_low_off(G.log_s, G.cpd, G.lv);
    int rc = SB_OK;
    cudaEvent_t ready[SB_MAX_DEV] = {0};
    for (int d = 0; d < G.g && rc == SB_OK; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard dg(c);
        sh->ctxs[d] = c;
        struct { uint4 **p; size_t bytes; const char *what; } want[4] = {{&sh->low[d], low_digests * 32, "low levels"}, {&sh->sub[d], (2 * S - 1) * 32, "subtree"},
                                                                        {&sh->stage[d], S * 32, "digest staging"}, {&sh->recv[d], S * 32, "digest receive"}};
        for (auto &w : want) {
            if (!w.bytes) continue;
            rc = blk_alloc(c, w.bytes, (void **)w.p);
            if (rc != SB_OK) {
                fail(ctx, rc, "%s (%s)", c->err, w.what);
                break;
            }
        }
        if (rc != SB_OK) break;
        cudaError_t e2 = cudaEventCreateWithFlags(&ready[d], cudaEventDisableTiming);
        if (e2 == cudaSuccess) e2 = cudaEventRecord(ready[d], c->stream);
        if (e2 != cudaSuccess) rc = fail(ctx, SB_ERR_CUDA, "device %d: event: %s", c->device, cudaGetErrorString(e2));
    }
    for (int d = 0; d < G.g && rc == SB_OK; d++) {
        DevGuard dg(root->dev[d]);
        for (int o = 0; o < G.g; o++)
            if (o != d) cudaStreamWaitEvent(root->dev[d]->stream, ready[o], 0);
    }
    for (int d = 0; d < G.g; d++)
        if (ready[d]) cudaEventDestroy(ready[d]);
    if (rc != SB_OK) {
        free_tree(t);
        return rc;
    }
    *out = t;
    return SB_OK;
}
// Called after every device has queued its leaf kernel (which leaves the level-lv digests in stage[d]).  Every device pushes the
// k-range of each node-range owner to that owner with one contiguous copy (peer DMA at full NVLink rate), every owner interleaves
// the g sources into level 0 of its subtree and reduces it; the subtree roots go to the host.
static int shards_finish(sb_ctx *root, sb_tree *t) {
    sb_ctx *ctx = root;
    TreeShards *sh = t->sh;
    const int g = sh->g;
    const size_t S = (size_t)1 << sh->log_s, per = S / g;
    std::vector<uint8_t> roots((size_t)g * 32);
    cudaEvent_t pushed[SB_MAX_DEV] = {0};
    int rc = SB_OK;
    TracePhase tp(root);
    tp.lap("leaf kernels");
    for (int d = 0; d < g && rc == SB_OK; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard dg(c);
        for (int k = 0; k < g && rc == SB_OK; k++) {        // ring order: every owner receives from one source at a time
            const int j = (d + k) % g;
            if (cudaMemcpyAsync(sh->recv[j] + 2 * ((size_t)d * per), sh->stage[d] + 2 * ((size_t)j * per), per * 32, cudaMemcpyDefault, c->stream) != cudaSuccess)
                rc = fail(ctx, SB_ERR_CUDA, "digest exchange %d -> %d: %s", d, j, cudaGetErrorString(cudaGetLastError()));
        }
        if (rc == SB_OK && (cudaEventCreateWithFlags(&pushed[d], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(pushed[d], c->stream) != cudaSuccess))
            rc = fail(ctx, SB_ERR_CUDA, "event: %s", cudaGetErrorString(cudaGetLastError()));
    }
    tp.lap("digest exchange");
    for (int d = 0; d < g && rc == SB_OK; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard dg(c);
        for (int o = 0; o < g; o++)
            if (o != d) cudaStreamWaitEvent(c->stream, pushed[o], 0);
        {
            sb_ctx *ctx = c;
            KLAUNCH(SB_KIND_MERKLE_NODES, merkle_launch_gather_reduce(c->stream, sh->recv[d], sh->sub[d], per, (uint32_t)g));
        }
        rc = nodes_build(c, sh->sub[d], S, ilog2((size_t)g));
        if (rc != SB_OK && c != root) fail(ctx, rc, "%s", c->err);
    }
    // the roots in a second loop: a copy into pageable host memory blocks the host until that device is done, so issuing it
    // inside the loop above ran the devices' subtrees one after the other (measured: 2.0 ms instead of 0.5 on 4 GPUs)
    for (int d = 0; d < g && rc == SB_OK; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard dg(c);
        if (cudaMemcpyAsync(&roots[32 * d], (const uint8_t *)sh->sub[d] + (2 * S - 2) * 32, 32, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
            rc = fail(ctx, SB_ERR_CUDA, "D2H subtree root: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc == SB_OK) rc = sync_all(root);
    tp.lap("gather + subtrees");
    for (int d = 0; d < g; d++) {
        if (pushed[d]) cudaEventDestroy(pushed[d]);
        DevGuard dg(root->dev[d]);                    // the staging buffers are only needed while the tree is built
        blk_release(root->dev[d], sh->stage[d]);
        blk_release(root->dev[d], sh->recv[d]);
        sh->stage[d] = sh->recv[d] = nullptr;
    }
    if (rc != SB_OK) return rc;
    // top of the tree on the host    // top of the tree on the host: g subtree roots -> root (merkle_proof_in_place.rs:78-98: parent = H(left || right))
    sh->top.assign((size_t)(2 * g - 1) * 32, 0);
    memcpy(sh->top.data(), roots.data(), (size_t)g * 32);
    for (uint32_t l = 0; ((size_t)g >> l) > 1; l++) {
        const size_t off = merkle_level_off((size_t)g, l), offn = merkle_level_off((size_t)g, l + 1);
        for (size_t i = 0; i < ((size_t)g >> (l + 1)); i++) b2s::hash_bytes(&sh->top[(offn + i) * 32], &sh->top[(off + 2 * i) * 32], 64);
    }
    memcpy(t->root, &sh->top[(size_t)(2 * g - 2) * 32], 32);
    return SB_OK;
}

// MerkleProofInPlace over the rows of the given columns (prove.rs:235-264 / :324-332): leaf 8 k + r =
// to_bytes_le(col_0[r][k]) || ... .  Synchronous (the root is on the host when it returns).
int ext_commit(const sb_ext *e, const size_t *col_ids, size_t n_ids, sb_tree **tree) {
    NvtxRange nvtx("ext_commit: leaf hashing + subtrees + top");
    sb_ctx *root = e->root, *ctx = root;
    if (n_ids < 1 || n_ids > 8) return fail(ctx, SB_ERR_ARG, "1..8 columns per leaf supported, got %zu", n_ids);
    for (size_t i = 0; i < n_ids; i++)
        if (col_ids[i] >= e->n_cols) return fail(ctx, SB_ERR_ARG, "column %zu out of range", col_ids[i]);
    const int g = e->g;
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
    sb_tree *t = nullptr;
    const ShardGeom G{g, e->log_s, e->cpd, e->lv};
    TRY(shards_begin(root, G, 32 * n_ids, (int)n_ids, &t));
    TreeShards *sh = t->sh;
    int rc = SB_OK;
    for (int d = 0; d < g; d++) {
        sb_ctx *c = root->dev[d];
        DevGuard dg(c);
        ExtLeavesParams Q = P;
        Q.d = (uint32_t)d;
        for (size_t i = 0; i < n_ids; i++) Q.cols[i] = sh->cols[d][i] = e->col(d, col_ids[i]);
        Q.low = sh->low[d];
        Q.stage = sh->stage[d];
        sb_ctx *ctx = c;       // launch accounting on the device that runs the kernel
        KLAUNCH(SB_KIND_MERKLE_LEAVES, merkle_launch_leaves_ext(c->stream, Q));
    }
    rc = shards_finish(root, t);
    if (rc != SB_OK) {
        free_tree(t);
        return rc;
    }
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

// prove_low_degree (fri/src/fri.rs:46-224) on one coset-major column.  A layer (fri.rs:120-213) runs where its values
// are: every device folds the rows of its cosets (the four values of a row share a coset: a quarter turn is S/4 steps of
// 8) and hashes the folded leaves.  Row 8 k + r of the folded column is position 8 k + r of the next layer's domain, so
// the column is coset-sharded exactly like its input: while a layer is large it stays on the devices (column next to the
// data, its tree as per-device subtrees); the first small layer is written in natural order, with its leaf digests,
// straight into the primary device's memory, and the remaining layers run there alone.  Openings of
// sharded trees are gathered by the primary through peer pointers.
static const uint32_t FRI_SHARD_MIN_LOG_S = 19;       // a layer stays sharded while every coset of the NEXT layer has >= 2^19 values
// (i.e. the next layer holds >= 2^22 values: below that a layer is a few hundred microseconds on one GPU and the per-device launches and
// synchronisations of a sharded layer cost more than they save -- measured: FRI of a 2^23 proof 2.4 ms with layer 0 alone sharded, 4.4 ms
// on 8 GPUs with three sharded layers)

int ext_fri_prove(const sb_ext *e, size_t col, const sb_tree *values_tree, size_t max_deg_plus_1, uint32_t excl, sb_fri_proof **out) {
    NvtxRange nvtx("ext_fri_prove");
    sb_ctx *root = e->root, *ctx = root;
    if (col >= e->n_cols) return fail(ctx, SB_ERR_ARG, "column %zu out of range", col);
    const int g = e->g;
    if (max_deg_plus_1 <= FRI_MIN_DEG_DIRECT) {
        // fri.rs:88-112: the values themselves are the proof; gather them in natural order
        DevBuf nat(ctx);
        TRY(nat.alloc(e->N * 32));
        TRY(sync_all(root));
        TRY(ext_to_natural(e, col, (uint4 *)nat.p));
        return fri_prove_dev(ctx, (const uint4 *)nat.p, e->N, e->g2, max_deg_plus_1, excl, nullptr, out);
    }
    // everything this call owns: trees, per-device column buffers of the sharded layers, the natural-order column
    struct Owned {
        sb_ctx *root;
        std::vector<sb_tree *> trees;
        std::vector<std::pair<int, void *>> bufs;       // (device index, pointer)
        sb_fri_proof *proof = nullptr;
        ~Owned() {
            for (auto t : trees) free_tree(t);
            for (auto &b : bufs) {
                DevGuard dg(root->dev[b.first]);
                blk_release(root->dev[b.first], b.second);
            }
            delete proof;
        }
    } own;
    own.root = root;
    own.proof = new sb_fri_proof();
    if (!values_tree) {
        sb_tree *t = nullptr;
        size_t id = col;
        TRY(ext_commit(e, &id, 1, &t));
        own.trees.push_back(t);
        values_tree = t;
    }
    const uint4 *cur[SB_MAX_DEV];
    for (int d = 0; d < g; d++) cur[d] = e->col(d, col);
    uint32_t cur_log_s = e->log_s, layer = 0;
    const sb_tree *cur_tree = values_tree;
    size_t bound = max_deg_plus_1;
    hfp::el w = e->g2;
    while (true) {
        const size_t n = (size_t)8 << cur_log_s, q = n / 4, S4 = (size_t)1 << (cur_log_s - 2);
        if (q >= (1u << 24) && !(ctx->extended_domain && q <= (1u << 28))) return fail(ctx, SB_ERR_ARG, "FRI layer of %zu values unsupported", n);
        const bool next_sharded = g > 1 && cur_log_s >= FRI_SHARD_MIN_LOG_S + 2 && bound / 4 > FRI_MIN_DEG_DIRECT;
        FriLayer L;
        memcpy(L.values_root, cur_tree->root, 32);
        const hfp::el special_x = hfp::from_bytes_le32(cur_tree->root);        // fri.rs:135
        // where the folded column and its tree go:
        //   next_sharded     column next to the data (coset-major), tree as per-device subtrees (fused fold + leaf hashing)
        //   g == 1           natural-order column + standard tree on the device (fused fold + leaf hashing)
        //   otherwise        the devices fold into local coset arrays, the primary pulls them with one contiguous peer copy per
        //                    device, reorders to natural order and hashes the (small) column itself
        sb_tree *t2 = nullptr;
        void *col_nat = nullptr;
        uint4 *nxt[SB_MAX_DEV] = {0};
        if (next_sharded) {
            const ShardGeom G{g, cur_log_s - 2, e->cpd, e->lv};
            TRY(shards_begin(root, G, 32, 1, &t2));
            own.trees.push_back(t2);
        } else {
            TRY(blk_alloc(ctx, q * 32, &col_nat));
            own.bufs.push_back({0, col_nat});
        }
        if (g > 1) {
            for (int d = 0; d < g; d++) {
                sb_ctx *c = root->dev[d];
                DevGuard dg(c);
                void *p = nullptr;
                if (blk_alloc(c, ((size_t)e->cpd * S4) * 32, &p) != SB_OK) return fail(ctx, SB_ERR_OOM, "%s", c->err);
                own.bufs.push_back({d, p});
                nxt[d] = (uint4 *)p;
                if (next_sharded) t2->sh->cols[d][0] = nxt[d];
            }
        } else {
            TRY(tree_new(ctx, q, 32, &t2));
            own.trees.push_back(t2);
            t2->n_cols = 1;
            t2->cols[0] = (const uint4 *)col_nat;
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
            FriFoldParams F;
            memset(&F, 0, sizeof F);
            F.vals = cur[d];
            F.tw = tw;
            F.n = n;
            F.tw_log_n = tw_log_n;
            F.tw_log_stride = tw_stride + 2 * layer;           // the layer's root is g2^(4^layer)
            memcpy(F.special_x, special_x.l, 32);
            F.log_s = cur_log_s; F.cpd = e->cpd; F.lv = e->lv; F.d = (uint32_t)d; F.g = (uint32_t)g;
            F.col_local = nxt[d];
            sb_ctx *ctx = c;
            if (next_sharded) {
                F.low = t2->sh->low[d];
                F.stage = t2->sh->stage[d];
                KLAUNCH(SB_KIND_FRI_FOLD, merkle_launch_leaves_fold_ext(c->stream, F, nullptr));
            } else if (g == 1) {
                F.col = (uint4 *)col_nat;
                KLAUNCH(SB_KIND_FRI_FOLD, merkle_launch_leaves_fold_ext(c->stream, F, t2->d_nodes));
            } else {
                KLAUNCH(SB_KIND_FRI_FOLD, fri_launch_fold_ext(c->stream, F));
            }
        }
        if (rc != SB_OK) return rc;
        if (next_sharded) {
            TRY(shards_finish(root, t2));
        } else if (g == 1) {
            TRY(merkle_finish(ctx, t2, e->lv, true));
        } else {
            // gather: coset arrays -> one coset-major buffer on the primary -> natural order -> standard tree
            DevBuf cm(ctx);                    // (block cache: peer-mapped pool memory is expensive to hand out)
            TRY(cm.alloc(q * 32));
            cudaEvent_t ready, folded[SB_MAX_DEV] = {0};
            CU(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
            CU(cudaEventRecord(ready, ctx->stream));
            for (int d = 0; d < g && rc == SB_OK; d++) {
                sb_ctx *c = root->dev[d];
                DevGuard dg(c);
                if (d > 0) cudaStreamWaitEvent(c->stream, ready, 0);
                if (cudaMemcpyAsync((uint4 *)cm.p + 2 * ((size_t)d * e->cpd * S4), nxt[d], (size_t)e->cpd * S4 * 32, cudaMemcpyDefault, c->stream) != cudaSuccess)
                    rc = fail(ctx, SB_ERR_CUDA, "column gather from device %d: %s", d, cudaGetErrorString(cudaGetLastError()));
                if (d > 0 && rc == SB_OK && (cudaEventCreateWithFlags(&folded[d], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(folded[d], c->stream) != cudaSuccess))
                    rc = fail(ctx, SB_ERR_CUDA, "event: %s", cudaGetErrorString(cudaGetLastError()));
            }
            cudaEventDestroy(ready);
            for (int d = 1; d < g; d++)
                if (folded[d]) {
                    cudaStreamWaitEvent(ctx->stream, folded[d], 0);
                    cudaEventDestroy(folded[d]);
                }
            if (rc != SB_OK) return rc;
            ExtOpenParams P;
            memset(&P, 0, sizeof P);
            P.cols[0][0] = (const uint4 *)cm.p;
            P.nc = 1; P.log_s = cur_log_s - 2; P.cpd = 8; P.lv = 3; P.g = 1;
            KLAUNCH(SB_KIND_OTHER, ext_launch_to_natural(ctx->stream, P, (uint4 *)col_nat));
            const uint4 *cols1[1] = {(const uint4 *)col_nat};
            TRY(commit_cols(ctx, cols1, 1, q, &t2));
            own.trees.push_back(t2);
        }
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
        L.n_poly = 4 * FRI_QUERIES;
        L.depth_poly = cur_tree->depth;
        L.poly_leaves.resize(L.n_poly * 32);
        L.poly_nodes.resize(L.n_poly * cur_tree->depth * 32);
        {
            const OpenReq reqs[2] = {{t2, yi.data(), FRI_QUERIES, L.column_leaves.data(), L.column_nodes.data()},
                                     {cur_tree, pp.data(), L.n_poly, L.poly_leaves.data(), L.poly_nodes.data()}};
            TRY(merkle_open_many(ctx, reqs, 2));
        }
        own.proof->layers.push_back(std::move(L));
        // fri.rs:215-223
        w = hfp::sqr(hfp::sqr(w));
        bound /= 4;
        layer++;
        if (!next_sharded) {
            // the rest on the primary device: natural order, column tree already committed
            sb_fri_proof *rest = nullptr;
            TRY(fri_prove_dev(ctx, (const uint4 *)col_nat, q, w, bound, excl, t2, &rest));
            for (auto &l : rest->layers) own.proof->layers.push_back(std::move(l));
            delete rest;
            break;
        }
        for (int d = 0; d < g; d++) cur[d] = nxt[d];
        cur_log_s -= 2;
        cur_tree = t2;
    }
    *out = own.proof;
    own.proof = nullptr;
    return SB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------------
extern "C" int sb_ext_create(sb_ctx *ctx, size_t n_cols, uint32_t log_s, uint32_t log_ext, sb_ext **out) {
    return guarded(ctx, __func__, [&]() -> int {
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
    return guarded(ctx, __func__, [&]() -> int {
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
    return guarded(ctx, __func__, [&]() -> int {
        if (!ctx || !e) return SB_ERR_ARG;
        TRY(ext_extend(e, first, count));
        return sync_all(ctx);
    });
}
extern "C" int sb_ext_commit(sb_ctx *ctx, const sb_ext *e, const size_t *col_ids, size_t n_ids, uint8_t root[32], sb_tree **tree) {
    return guarded(ctx, __func__, [&]() -> int {
        if (!ctx || !e || !col_ids || !tree) return SB_ERR_ARG;
        TRY(ext_commit(e, col_ids, n_ids, tree));
        if (root) memcpy(root, (*tree)->root, 32);
        return SB_OK;
    });
}
extern "C" int sb_ext_fri_prove(sb_ctx *ctx, const sb_ext *e, size_t col, const sb_tree *values_tree, size_t max_deg_plus_1, uint32_t excl,
                                sb_fri_proof **out) {
    return guarded(ctx, __func__, [&]() -> int {
        if (!ctx || !e || !out) return SB_ERR_ARG;
        if (values_tree && (values_tree->n != e->N || values_tree->leaf_bytes != 32)) return fail(ctx, SB_ERR_ARG, "values_tree does not match the column");
        return ext_fri_prove(e, col, values_tree, max_deg_plus_1, excl, out);
    });
}
extern "C" int sb_ext_read(sb_ctx *ctx, const sb_ext *e, size_t col, uint64_t *out) {
    return guarded(ctx, __func__, [&]() -> int {
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
