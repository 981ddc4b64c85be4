// internal.h -- definitions shared by the host translation units of libstark_b200.so (api.cu, prover.cu).
// Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/stark_b200.h"
#include "blake2s.cuh"
#include "hostfp.h"
#include "kernels.h"
#include "params.h"

// NVTX range over a stage (visible in Nsight Systems / ncu --nvtx): header-only NVTX v3, a no-op
// without an attached tool
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

// consecutive stages of one call: next() closes the previous range; the destructor closes the last one on every exit path
struct NvtxStages {
    bool open = false;
    void next(const char *name) {
        if (open) nvtxRangePop();
        nvtxRangePushA(name);
        open = true;
    }
    ~NvtxStages() {
        if (open) nvtxRangePop();
    }
};

struct TwTable {
    hfp::el root;
    uint32_t log_n;
    uint4 *d;
    uint64_t last_use;
};

struct sb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t launches = 0;
    std::vector<TwTable> tables;
    uint64_t table_clock = 0;
    size_t table_cache_bytes = (size_t)8 << 30;   // soft cap of the twiddle-table cache (SB_TABLE_CACHE_BYTES)
    char err[512] = {0};
    bool extended_domain = false;      // sb_set_extended_domain: FRI layers beyond the reference sampler's 2^24 limit
    // copy streams of the host-buffer entry points (sb_lde_batch pipelines upload / transform / download), created on demand
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    // pinned host staging arena (front end -> sb_prove_r1cs uploads); grows on demand, freed in sb_destroy
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    void *pinned2 = nullptr;          // small pinned scratch for host scalars / coefficient vectors sent to several devices
    size_t pinned2_bytes = 0;
    uint32_t *poseidon_consts = nullptr;   // device copy of the Poseidon parameters (poseidon.cu), made on first use
    // optional per-kernel-family timing (sb_profile): CUDA events around every launch
    bool prof = false;
    struct ProfRec { int kind; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[SB_KIND_COUNT] = {0};
    uint64_t prof_n[SB_KIND_COUNT] = {0};
    // Block cache for the large per-call arrays of the resident pipeline (extended columns, sharded trees, FRI layers).  Sizes repeat
    // from proof to proof, so a freed block is kept and handed out again instead of going back to the stream-ordered pool: with
    // peer access enabled, allocating / freeing pool memory cost ~0.1 ms per call and device on an 8-GPU context (measured: a
    // sharded tree of 2^26 leaves spent 7 of its 9 ms there).  All users enqueue on the device's one compute stream, so reuse is
    // stream-ordered like the pool's.
    struct Block { void *p; size_t bytes; };
    std::vector<Block> blk_free, blk_live;
    size_t blk_free_bytes = 0;
    // multi-device contexts (sb_init_multi): the primary context is dev[0] and owns one sub-context per further device;
    // every sub-context is a full context (own stream, table cache, profile counters) whose `primary` points back.
    // Plain sb_init contexts have dev = {this}.  Two entries may name the same physical GPU (logical devices: the
    // sharded code paths then run unchanged on a single-GPU box).
    std::vector<sb_ctx *> dev;
    sb_ctx *primary = nullptr;
    int n_dev() const { return dev.empty() ? 1 : (int)dev.size(); }
};

static cudaEvent_t prof_event(sb_ctx *ctx) {
    cudaEvent_t e;
    if (!ctx->prof_pool.empty()) {
        e = ctx->prof_pool.back();
        ctx->prof_pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
static void prof_begin(sb_ctx *ctx, int kind) {
    if (!ctx->prof) return;
    sb_ctx::ProfRec r{kind, prof_event(ctx), prof_event(ctx)};
    cudaEventRecord(r.a, ctx->stream);
    ctx->prof_recs.push_back(r);
}
static void prof_end(sb_ctx *ctx) {
    if (!ctx->prof) return;
    cudaEventRecord(ctx->prof_recs.back().b, ctx->stream);
}
static void prof_collect(sb_ctx *ctx) {
    for (auto &r : ctx->prof_recs) {
        float ms = 0;
        cudaEventSynchronize(r.b);
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            ctx->prof_ms[r.kind] += ms;
            ctx->prof_n[r.kind]++;
        }
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof_recs.clear();
}
#define KLAUNCH(kind, expr)          \
    do {                             \
        prof_begin(ctx, kind);       \
        ctx->launches += (expr);     \
        prof_end(ctx);               \
    } while (0)

static int fail(sb_ctx *ctx, int code, const char *fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof ctx->err, fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CU(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? SB_ERR_OOM : SB_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                \
    } while (0)
// Every extern "C" entry point runs its body through guarded(): (1) the context's device becomes the calling thread's
// current device for the duration of the call (the current device is per host thread: a context for device != 0 used from
// a worker thread, two contexts in one process, or torch switching devices would otherwise launch on the wrong GPU) and is
// restored afterwards; (2) no C++ exception crosses the C ABI (std::bad_alloc from a vector sized by untrusted input
// becomes SB_ERR_OOM, anything else SB_ERR_ARG).
struct DevGuard {
    int prev = -1;
    bool switched = false;
    explicit DevGuard(const sb_ctx *ctx) {
        if (!ctx) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device) switched = cudaSetDevice(ctx->device) == cudaSuccess;
    }
    explicit DevGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DevGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DevGuard(const DevGuard &) = delete;
    DevGuard &operator=(const DevGuard &) = delete;
};
template <class F>
static int guarded(sb_ctx *ctx, const char *entry, F &&body) {
    DevGuard g(ctx);
    int rc;
    try {
        rc = body();
    } catch (const std::bad_alloc &) {
        rc = fail(ctx, SB_ERR_OOM, "host allocation failed");
    } catch (const std::exception &e) {
        rc = fail(ctx, SB_ERR_ARG, "exception: %s", e.what());
    } catch (...) {
        rc = fail(ctx, SB_ERR_ARG, "unknown exception");
    }
    // A CUDA call whose status was dropped leaves its error pending and the NEXT entry point reports it against the wrong call:
    // surface it here (SB_DEBUG_ERRORS=1 names the entry point on stderr) and do not let it leak out of a successful call.
    if (ctx) {
        const cudaError_t pending = cudaGetLastError();
        if (pending != cudaSuccess) {
            static const bool verbose = getenv("SB_DEBUG_ERRORS") != nullptr;
            if (verbose) fprintf(stderr, "stark_b200: %s left a pending CUDA error: %s (rc %d)\n", entry, cudaGetErrorString(pending), rc);
            if (rc == SB_OK) rc = fail(ctx, SB_ERR_CUDA, "%s: a CUDA call failed: %s", entry, cudaGetErrorString(pending));
        }
    }
    return rc;
}

// SB_DEBUG_ERRORS=1: report (and clear) a pending CUDA error at a named point
static inline void dbg_check(const char *where) {
    static const bool verbose = getenv("SB_DEBUG_ERRORS") != nullptr;
    if (!verbose) return;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fprintf(stderr, "stark_b200: pending CUDA error at %s: %s\n", where, cudaGetErrorString(e));
}

#define TRY(expr)                \
    do {                         \
        int rc_ = (expr);        \
        if (rc_ != SB_OK) return rc_; \
    } while (0)

// Tree over coset-major columns spread over g devices (ext.cu): device d hashes the leaves of its cosets and keeps the
// lv = log2(cosets per device) lowest levels of those; the level-lv digests are written straight into the memory of the
// device that owns their node range, which builds the subtree over its S level-lv digests; the top log2 g levels are
// finished on the host from the g subtree roots.
struct TreeShards {
    int g = 1;
    uint32_t lv = 0, log_s = 0, cpd = 8;
    uint4 *low[SB_MAX_DEV] = {0};          // per device: levels 0 .. lv-1 of its own leaves (NULL when lv == 0)
    uint4 *sub[SB_MAX_DEV] = {0};          // per device: 2S - 1 digests, standard layout over its S level-lv digests
    uint4 *stage[SB_MAX_DEV] = {0};        // build only: the level-lv digests a device produced (S, by step k)
    uint4 *recv[SB_MAX_DEV] = {0};         // build only: the digests of a device's node range, by source device (g x S/g)
    const uint4 *cols[SB_MAX_DEV][8] = {{0}};   // per device: (column, coset 0, k = 0) of the committed columns
    sb_ctx *ctxs[SB_MAX_DEV] = {0};        // the devices' contexts (block cache, stream, ordinal)
    std::vector<uint8_t> top;              // host: (2g - 1) digests over the subtree roots, standard layout
};

// cached device blocks (see sb_ctx::blk_free): blk_alloc never returns a block more than twice the requested size
static int blk_alloc(sb_ctx *ctx, size_t bytes, void **out) {
    if (!bytes) bytes = 16;
    size_t best = (size_t)-1;
    for (size_t i = 0; i < ctx->blk_free.size(); i++)
        if (ctx->blk_free[i].bytes >= bytes && ctx->blk_free[i].bytes <= 2 * bytes && (best == (size_t)-1 || ctx->blk_free[i].bytes < ctx->blk_free[best].bytes)) best = i;
    if (best != (size_t)-1) {
        sb_ctx::Block b = ctx->blk_free[best];
        ctx->blk_free.erase(ctx->blk_free.begin() + best);
        ctx->blk_free_bytes -= b.bytes;
        ctx->blk_live.push_back(b);
        *out = b.p;
        return SB_OK;
    }
    void *p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, bytes, ctx->stream);
    if (e == cudaErrorMemoryAllocation) {          // give the cache back, let pending frees land, try once more
        cudaGetLastError();
        for (auto &b : ctx->blk_free) cudaFreeAsync(b.p, ctx->stream);
        ctx->blk_free.clear();
        ctx->blk_free_bytes = 0;
        cudaDeviceSynchronize();
        e = cudaMallocAsync(&p, bytes, ctx->stream);
    }
    if (e != cudaSuccess) {
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        cudaGetLastError();
        *out = nullptr;
        return fail(ctx, e == cudaErrorMemoryAllocation ? SB_ERR_OOM : SB_ERR_CUDA, "device %d: cudaMallocAsync(%zu bytes) failed: %s (free %zu of %zu MiB)", ctx->device, bytes,
                    cudaGetErrorString(e), fr >> 20, tot >> 20);
    }
    ctx->blk_live.push_back({p, bytes});
    *out = p;
    return SB_OK;
}
static void blk_release(sb_ctx *ctx, void *p) {
    if (!p) return;
    for (size_t i = 0; i < ctx->blk_live.size(); i++)
        if (ctx->blk_live[i].p == p) {
            sb_ctx::Block b = ctx->blk_live[i];
            ctx->blk_live.erase(ctx->blk_live.begin() + i);
            ctx->blk_free.push_back(b);
            ctx->blk_free_bytes += b.bytes;
            // soft cap: keep at most 48 GiB parked; the oldest blocks go back to the pool
            while (ctx->blk_free_bytes > ((size_t)48 << 30) && !ctx->blk_free.empty()) {
                cudaFreeAsync(ctx->blk_free.front().p, ctx->stream);
                ctx->blk_free_bytes -= ctx->blk_free.front().bytes;
                ctx->blk_free.erase(ctx->blk_free.begin());
            }
            return;
        }
    cudaFreeAsync(p, ctx->stream);       // not one of ours
}

struct DevBuf {   // stream-ordered scratch: the block cache from 1 MiB on, the pool below
    sb_ctx *ctx;
    void *p = nullptr;
    bool cached = false;
    explicit DevBuf(sb_ctx *c) : ctx(c) {}
    int alloc(size_t bytes) {
        if (bytes >= ((size_t)1 << 20)) {
            cached = true;
            return blk_alloc(ctx, bytes, &p);
        }
        cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 16, ctx->stream);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(ctx, e == cudaErrorMemoryAllocation ? SB_ERR_OOM : SB_ERR_CUDA, "cudaMallocAsync(%zu): %s", bytes,
                        cudaGetErrorString(e));
        }
        return SB_OK;
    }
    ~DevBuf() {
        if (!p) return;
        if (cached) blk_release(ctx, p);
        else cudaFreeAsync(p, ctx->stream);
    }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

struct sb_tree {
    size_t n = 0;
    uint32_t depth = 0;
    size_t leaf_bytes = 0;
    uint4 *d_nodes = nullptr;      // 2n - 1 digests, level l at merkle_level_off(n, l)
    uint8_t *d_leaves = nullptr;   // owned copy of byte leaves (NULL for column-backed trees)
    int n_cols = 0;
    const uint4 *cols[8] = {0};
    uint32_t coset_log_s = 0;      // != 0: the columns are coset-major (leaf i = element (i & 7) * 2^coset_log_s + (i >> 3))
    uint8_t root[32] = {0};
    cudaStream_t stream = nullptr; // allocations are stream-ordered (pool) on the owning context's stream
    int device = 0;
    sb_ctx *owner = nullptr;       // context whose block cache d_nodes came from
    TreeShards *sh = nullptr;      // != NULL: the levels live in the shards (d_nodes == NULL)
};

static bool is_pow2(size_t n) { return n && !(n & (n - 1)); }
static uint32_t ilog2(size_t n) {
    uint32_t l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}

struct FriLayer {
    bool is_last = false;
    uint8_t values_root[32] = {0};      // root of the tree over this layer's values (tap)
    uint8_t root2[32] = {0};
    size_t n_column = 0, depth_column = 0, n_poly = 0, depth_poly = 0;
    std::vector<uint8_t> column_leaves, column_nodes, poly_leaves, poly_nodes;
    std::vector<uint8_t> last;           // n_last * 32
};
struct sb_fri_proof {
    std::vector<FriLayer> layers;
};

static const size_t FRI_MIN_DEG_DIRECT = 16;   // fri.rs:14
static const size_t FRI_QUERIES = 40;          // fri.rs:184

static fp to_dev_fp(const hfp::el &a) {
    fp r;
    memcpy(r.l, a.l, 32);
    return r;
}

// ---- functions defined in api.cu ---------------------------------------------------------------
int pseudorandom_indices(const uint8_t *seed, size_t seed_len, uint32_t modulus, size_t count, uint32_t excl, uint32_t *out,
                         bool extended);
// returns a pinned host buffer of at least `bytes` owned by the context (contents are scratch), NULL on failure
void *pinned_arena(sb_ctx *ctx, size_t bytes);
void *pinned_scratch(sb_ctx *ctx, size_t bytes);
// dense: the caller indexes the table directly (the prover's `xs`), so a strided view of a larger cached table will not do
int get_table(sb_ctx *ctx, const hfp::el &w, uint32_t log_n, const uint4 **tw, uint32_t *tw_log_n, uint32_t *log_stride,
              bool dense = false);
// coset transforms of a low-degree extension by 2^log_ext (NttPassParams): cosets r0 .. r0 + cnt - 1 of every polynomial
struct CosetSpec {
    uint32_t log_ext = 3, r0 = 1, cnt = 7;
    uint32_t store = NTT_STORE_PLAIN;
    uint32_t dst_cpd = 0, dst_r0 = 0;  // NTT_STORE_PLAIN: (column, coset) lands at polynomial slot column * dst_cpd + coset - dst_r0 (0: dst_cpd = cnt, dst_r0 = r0)
};
int ntt_dev_tw(sb_ctx *ctx, const uint4 *d_src, size_t len_in, size_t src_stride, uint4 *d_dst, size_t dst_stride, size_t n_polys,
               uint32_t log_n, int inverse, const uint4 *tw, uint32_t tw_log_n, uint32_t log_stride, const CosetSpec *cs);
int ntt_dev(sb_ctx *ctx, const uint4 *d_src, size_t len_in, size_t src_stride, uint4 *d_dst, size_t dst_stride, size_t n_polys,
            const hfp::el &root, uint32_t log_n, int inverse);
int lde_dev(sb_ctx *ctx, const uint4 *d_cols, size_t n_cols, size_t col_len, size_t col_stride, const hfp::el &root_big,
            uint32_t log_s, uint32_t log_ext, uint4 *d_out);
int ctx_create(int device, sb_ctx **out);
void free_tree(sb_tree *t);
int tree_new(sb_ctx *ctx, size_t n, size_t leaf_bytes, sb_tree **out);
int merkle_finish(sb_ctx *ctx, sb_tree *t, uint32_t level, bool fetch_root);
int commit_bytes_owned(sb_ctx *ctx, uint8_t *d_leaves, size_t leaf_bytes, size_t n, sb_tree **tree);
int commit_cols(sb_ctx *ctx, const uint4 *const *d_cols, size_t n_cols, size_t n, sb_tree **tree);
int fri_prove_dev(sb_ctx *ctx, const uint4 *d_vals, size_t n, const hfp::el &root, size_t max_deg_plus_1, uint32_t excl,
                  const sb_tree *values_tree, sb_fri_proof **out);
// ---- ext.cu: columns of the extended domain in coset-major layout, spread over the context's devices --------------------
// Device d of g holds cosets r0 = d * cpd .. r0 + cpd - 1 (cpd = 8 / g) of EVERY column: element (c, r, k) -- the value at
// position 8 k + r of column c -- lives at buf[d][((c * cpd + r - r0) << log_s) + k].  Every shift the prover applies is a
// multiple of 8 (prove.rs: previous step, +o3 steps, the FRI fold's quarter turns), so pointwise kernels, leaf hashing and
// the first FRI fold never leave a device; only coefficients (once per column) and 32-byte digests cross NVLink.
struct sb_ext {
    sb_ctx *root = nullptr;
    int g = 1;
    uint32_t log_s = 0, cpd = 8, lv = 3;
    size_t S = 0, N = 0, n_cols = 0, n_lde = 0;       // the first n_lde columns have input / coefficient staging (ext_extend)
    hfp::el g2;                                 // root of unity of the extended domain (order N)
    uint4 *buf[SB_MAX_DEV] = {0};               // n_cols * cpd * S elements
    uint4 *coef[SB_MAX_DEV] = {0};              // n_lde * S coefficients (every device holds every column's)
    uint4 *in[SB_MAX_DEV] = {0};                // n_lde * S input staging; column c is read on device owner[c]
    std::vector<int> owner;
    uint4 *col(int d, size_t c) const { return buf[d] + 2 * ((c * cpd) << log_s); }
    uint4 *input(size_t c) const { return in[owner[c]] + 2 * (c << log_s); }
};
int ext_create(sb_ctx *root, size_t n_cols, size_t n_lde, uint32_t log_s, sb_ext **out);
void ext_free(sb_ext *e);
int ext_extend(sb_ext *e, size_t first, size_t count);
int ext_commit(const sb_ext *e, const size_t *col_ids, size_t n_ids, sb_tree **tree);
int ext_fri_prove(const sb_ext *e, size_t col, const sb_tree *values_tree, size_t max_deg_plus_1, uint32_t excl, sb_fri_proof **out);
int ext_to_natural(const sb_ext *e, size_t col, uint4 *d_out);
int sync_all(sb_ctx *root);

// calc_flags (run.rs:283-308) in compressed form: the three flag vectors are all-one / all-one / all-zero except at one position per
// constraint (its last row l, and (l + 1) mod a in each third), so sb_prove_files hands the prover the list of those rows and the
// vectors are generated on the device instead of being written (92 MB at 955 086 steps) and uploaded
struct FlagSpec {
    const unsigned long long *last_rows;   // l = last row of every constraint that has rows
    size_t n_last;
    size_t a;                              // rows of one third (original_steps / 3)
};
int prove_r1cs_impl(sb_ctx *ctx, const sb_trace *t, const FlagSpec *flags, sb_stark_proof **out);

struct OpenReq {
    const sb_tree *t;
    const size_t *idx;
    size_t n_idx;
    uint8_t *leaves_out, *nodes_out;      // either may be NULL
};
int merkle_open_many(sb_ctx *ctx, const OpenReq *reqs, int n_req);
void json_bytes(std::string &s, const uint8_t *b, size_t n);
char *json_bytes_raw(char *w, const uint8_t *b, size_t n);
char *json_branches_raw(char *w, const uint8_t *leaves, size_t leaf_bytes, const uint8_t *nodes, size_t depth, size_t q0, size_t q1);
size_t json_branches_bound(size_t leaf_bytes, size_t depth, size_t count);
void json_branches(std::string &s, const uint8_t *leaves, size_t leaf_bytes, const uint8_t *nodes, size_t depth, size_t count);
void fri_proof_json_into(std::string &s, const sb_fri_proof *p);
void fri_layer_json_into(std::string &s, const FriLayer &L);
