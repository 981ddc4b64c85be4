// api.cu -- host side of libstark_b200.so: the C ABI of include/stark_b200.h on top of the sm_100a
// kernels (ntt.cuh, merkle.cuh, fri.cuh).  No CPU fallback anywhere: every entry point that does
// vector work needs a live context, and sb_init fails without a compute-capability-10 device.
#include "internal.h"

#include <algorithm>

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
int ctx_create(int device, sb_ctx **out) {
    if (!out) return SB_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return SB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SB_ERR_NO_DEVICE;
    if (prop.major != 10) return SB_ERR_NO_DEVICE;   // kernels are built for sm_100a only
    DevGuard g(device);
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != device) return SB_ERR_NO_DEVICE;
    sb_ctx *ctx = new sb_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return SB_ERR_NO_DEVICE;
    }
    ctx->own_stream = true;
    if (const char *e = getenv("SB_TABLE_CACHE_BYTES")) {
        const unsigned long long v = strtoull(e, nullptr, 10);
        if (v) ctx->table_cache_bytes = (size_t)v;
    }
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    // keep freed scratch cached in the stream-ordered pool
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaError_t e = ntt_set_attrs();
    if (e != cudaSuccess) {
        fprintf(stderr, "stark_b200: kernel image not loadable on this device: %s\n", cudaGetErrorString(e));
        sb_destroy(ctx);
        return SB_ERR_NO_DEVICE;
    }
    ctx->dev.push_back(ctx);
    *out = ctx;
    return SB_OK;
}

extern "C" int sb_init(int device, sb_ctx **out) {
    try {
        return ctx_create(device, out);
    } catch (...) {
        return SB_ERR_OOM;
    }
}

// One context over several GPUs of the node (one process, peer access over NVLink / NVSwitch).  devices[0] is the primary
// device: every single-device entry point runs there; sb_prove_r1cs / sb_prove_files and the sb_ext_* calls spread their
// work over all of them.  n_devices must be 1, 2, 4 or 8.  The same ordinal may appear more than once (logical devices
// sharing one GPU: same code path, no speed-up -- that is how the sharded logic is tested on a single-GPU box).
extern "C" int sb_init_multi(const int *devices, int n_devices, sb_ctx **out) {
    if (!out || !devices) return SB_ERR_ARG;
    *out = nullptr;
    if (n_devices != 1 && n_devices != 2 && n_devices != 4 && n_devices != 8) return SB_ERR_ARG;
    try {
        sb_ctx *root = nullptr;
        int rc = ctx_create(devices[0], &root);
        if (rc != SB_OK) return rc;
        for (int i = 1; i < n_devices; i++) {
            sb_ctx *c = nullptr;
            rc = ctx_create(devices[i], &c);
            if (rc != SB_OK) {
                sb_destroy(root);
                return rc;
            }
            c->primary = root;
            root->dev.push_back(c);
        }
        for (int i = 0; i < n_devices; i++)          // peer access between every pair of distinct GPUs
            for (int j = 0; j < n_devices; j++) {
                if (devices[i] == devices[j]) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can) {
                    fail(root, SB_ERR_NO_DEVICE, "no peer access from device %d to device %d", devices[i], devices[j]);
                    fprintf(stderr, "stark_b200: %s\n", root->err);
                    sb_destroy(root);
                    return SB_ERR_NO_DEVICE;
                }
                DevGuard g(devices[i]);
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    sb_destroy(root);
                    return SB_ERR_NO_DEVICE;
                }
                cudaGetLastError();
                // the stream-ordered pool (cudaMallocAsync: every column, tree and scratch buffer) has its own access list:
                // device i may read and write device j's pool
                cudaMemPool_t pool;
                cudaMemAccessDesc desc = {};
                desc.location.type = cudaMemLocationTypeDevice;
                desc.location.id = devices[i];
                desc.flags = cudaMemAccessFlagsProtReadWrite;
                if (cudaDeviceGetDefaultMemPool(&pool, devices[j]) != cudaSuccess || cudaMemPoolSetAccess(pool, &desc, 1) != cudaSuccess) {
                    fail(root, SB_ERR_NO_DEVICE, "cannot open the memory pool of device %d to device %d: %s", devices[j], devices[i],
                         cudaGetErrorString(cudaGetLastError()));
                    fprintf(stderr, "stark_b200: %s\n", root->err);
                    sb_destroy(root);
                    return SB_ERR_NO_DEVICE;
                }
            }
        *out = root;
        return SB_OK;
    } catch (...) {
        return SB_ERR_OOM;
    }
}
extern "C" int sb_device_count(const sb_ctx *ctx) { return ctx ? ctx->n_dev() : 0; }

extern "C" void sb_destroy(sb_ctx *ctx) {
    if (!ctx) return;
    for (size_t i = 1; i < ctx->dev.size(); i++) {
        ctx->dev[i]->dev.clear();
        sb_destroy(ctx->dev[i]);
    }
    DevGuard g(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (auto &b : ctx->blk_free) cudaFreeAsync(b.p, ctx->stream);
    for (auto &b : ctx->blk_live) cudaFreeAsync(b.p, ctx->stream);      // leaked by the caller (trees / columns never freed)
    ctx->blk_free.clear();
    ctx->blk_live.clear();
    cudaStreamSynchronize(ctx->stream);
    for (auto &t : ctx->tables) cudaFree(t.d);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->pinned2) cudaFreeHost(ctx->pinned2);
    if (ctx->poseidon_consts) cudaFree(ctx->poseidon_consts);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    prof_collect(ctx);
    for (auto e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    {   // hand the cached scratch of the stream-ordered pool back to the driver (another process may want this GPU's memory)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    delete ctx;
}

void *pinned_arena(sb_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pinned_bytes) return ctx->pinned;
    // uploads from the old arena ran on the devices' copy streams as well: wait for everything before it goes away (syncing the
    // compute stream alone left a pending "invalid argument" behind the next cudaMallocHost and a crash later on)
    for (sb_ctx *c : ctx->dev) {
        DevGuard g(c);
        cudaDeviceSynchronize();
    }
    static const bool verbose = getenv("SB_DEBUG_ERRORS") != nullptr;
    if (ctx->pinned) {
        const cudaError_t e = cudaFreeHost(ctx->pinned);
        if (verbose) fprintf(stderr, "stark_b200: pinned_arena: cudaFreeHost(%p) -> %s\n", ctx->pinned, cudaGetErrorString(e));
        cudaGetLastError();
    }
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    const size_t want = bytes + bytes / 8;
    const cudaError_t e = cudaMallocHost(&ctx->pinned, want);
    if (verbose) fprintf(stderr, "stark_b200: pinned_arena: cudaMallocHost(%zu) -> %s, %p, pending: %s\n", want, cudaGetErrorString(e), ctx->pinned, cudaGetErrorString(cudaPeekAtLastError()));
    if (e != cudaSuccess) {
        cudaGetLastError();
        ctx->pinned = nullptr;
        return nullptr;
    }
    ctx->pinned_bytes = want;
    return ctx->pinned;
}

void *pinned_scratch(sb_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->pinned2_bytes) return ctx->pinned2;
    for (sb_ctx *c : ctx->dev) {                                       // an upload from the old buffer may still be in flight
        DevGuard g(c);
        cudaDeviceSynchronize();
    }
    if (ctx->pinned2) cudaFreeHost(ctx->pinned2);
    ctx->pinned2 = nullptr;
    ctx->pinned2_bytes = 0;
    const size_t want = bytes < ((size_t)1 << 16) ? ((size_t)1 << 16) : 2 * bytes;
    if (cudaMallocHost(&ctx->pinned2, want) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    ctx->pinned2_bytes = want;
    return ctx->pinned2;
}

extern "C" const char *sb_last_error(const sb_ctx *ctx) { return ctx ? ctx->err : "no context"; }

extern "C" int sb_set_stream(sb_ctx *ctx, void *s) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx) return SB_ERR_ARG;
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)s;
    ctx->own_stream = false;
    return SB_OK;
    });
}
extern "C" int sb_sync(sb_ctx *ctx) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx) return SB_ERR_ARG;
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 1; i < ctx->dev.size(); i++) CU(cudaStreamSynchronize(ctx->dev[i]->stream));
    return SB_OK;
    });
}
extern "C" int sb_timer_start(sb_ctx *ctx) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx) return SB_ERR_ARG;
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    return SB_OK;
    });
}
extern "C" int sb_timer_stop(sb_ctx *ctx, float *ms) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !ms) return SB_ERR_ARG;
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaEventSynchronize(ctx->ev1));
    CU(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return SB_OK;
    });
}
extern "C" uint64_t sb_launch_count(const sb_ctx *ctx) {
    if (!ctx) return 0;
    uint64_t n = ctx->launches;
    for (size_t i = 1; i < ctx->dev.size(); i++) n += ctx->dev[i]->launches;
    return n;
}
extern "C" int sb_profile(sb_ctx *ctx, int enable) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx) return SB_ERR_ARG;
    for (sb_ctx *c : ctx->dev) {
        DevGuard g(c);
        prof_collect(c);
        c->prof = enable != 0;
        for (int k = 0; k < SB_KIND_COUNT; k++) c->prof_ms[k] = 0, c->prof_n[k] = 0;
    }
    return SB_OK;
    });
}
// multi-device contexts: launches and milliseconds are summed over the devices
extern "C" int sb_profile_read(sb_ctx *ctx, int kind, uint64_t *launches, double *total_ms) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || kind < 0 || kind >= SB_KIND_COUNT) return SB_ERR_ARG;
    uint64_t n = 0;
    double ms = 0;
    for (sb_ctx *c : ctx->dev) {
        DevGuard g(c);
        CU(cudaStreamSynchronize(c->stream));
        prof_collect(c);
        n += c->prof_n[kind];
        ms += c->prof_ms[kind];
    }
    if (launches) *launches = n;
    if (total_ms) *total_ms = ms;
    return SB_OK;
    });
}

extern "C" int sb_dev_alloc(sb_ctx *ctx, size_t bytes, void **p) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !p) return SB_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMalloc(p, bytes ? bytes : 16));
    return SB_OK;
    });
}
extern "C" int sb_dev_free(sb_ctx *ctx, void *p) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx) return SB_ERR_ARG;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(p));
    return SB_OK;
    });
}
extern "C" int sb_h2d(sb_ctx *ctx, void *d, const void *s, size_t bytes) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx) return SB_ERR_ARG;
    CU(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
    });
}
extern "C" int sb_d2h(sb_ctx *ctx, void *d, const void *s, size_t bytes) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx) return SB_ERR_ARG;
    CU(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
    });
}
extern "C" int sb_host_alloc_pinned(sb_ctx *ctx, size_t bytes, void **p) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !p) return SB_ERR_ARG;
    CU(cudaMallocHost(p, bytes ? bytes : 16));
    return SB_OK;
    });
}
extern "C" int sb_host_free_pinned(sb_ctx *ctx, void *p) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx) return SB_ERR_ARG;
    CU(cudaFreeHost(p));
    return SB_OK;
    });
}

// ------------------------------------------------------------------------------------------------
// twiddle tables
// ------------------------------------------------------------------------------------------------
static int powers_into(sb_ctx *ctx, const hfp::el &root, size_t n, uint4 *d_out) {
    if (n == 0) return SB_OK;
    dbg_check("powers_into entry");
    size_t seed = n < 1024 ? n : 1024;
    KLAUNCH(SB_KIND_OTHER, powers_launch_seed(ctx->stream, d_out, seed, to_dev_fp(root)));
    hfp::el wc = hfp::pow_u64(root, 1024);
    for (size_t cur = 1024; cur < n; cur <<= 1) {
        KLAUNCH(SB_KIND_OTHER, powers_launch_double(ctx->stream, d_out, cur, n, to_dev_fp(wc)));
        wc = hfp::sqr(wc);
    }
    CU(cudaGetLastError());
    return SB_OK;
}

// w must be a primitive 2^log_n-th root of unity (log_n = 0: w = 1)
static int check_root(sb_ctx *ctx, const hfp::el &w, uint32_t log_n) {
    if (log_n > 28) return fail(ctx, SB_ERR_ARG, "log_n %u exceeds the field's two-adicity 28", log_n);
    if (log_n == 0) return hfp::eq(w, hfp::ONE) ? SB_OK : fail(ctx, SB_ERR_ROOT, "root of order 1 must be 1");
    hfp::el t = w;
    for (uint32_t i = 0; i + 1 < log_n; i++) t = hfp::sqr(t);
    if (!hfp::eq(t, hfp::neg(hfp::ONE))) return fail(ctx, SB_ERR_ROOT, "root_of_unity^(2^%u) != -1: not a primitive 2^%u-th root", log_n - 1, log_n);
    return SB_OK;
}

// finds or builds a table T with w = T-root^(2^log_stride)
int get_table(sb_ctx *ctx, const hfp::el &w, uint32_t log_n, const uint4 **tw, uint32_t *tw_log_n, uint32_t *log_stride,
              bool dense) {
    TRY(check_root(ctx, w, log_n));
    for (auto &t : ctx->tables) {
        if (t.log_n < log_n || (dense && t.log_n != log_n)) continue;
        hfp::el r = t.root;
        for (uint32_t i = 0; i < t.log_n - log_n; i++) r = hfp::sqr(r);
        if (hfp::eq(r, w)) {
            t.last_use = ++ctx->table_clock;
            *tw = t.d;
            *tw_log_n = t.log_n;
            *log_stride = t.log_n - log_n;
            return SB_OK;
        }
    }
    // soft cap: drop least-recently-used tables, but never one of the four most recently used (a single API call
    // works with at most three: the extended domain's, a strided view of it, and the prover's dense copy)
    {
        size_t total = (size_t)32 << log_n;
        for (auto &t : ctx->tables) total += (size_t)32 << t.log_n;
        while (total > ctx->table_cache_bytes && ctx->tables.size() > 4) {
            size_t lru = 0;
            for (size_t i = 1; i < ctx->tables.size(); i++)
                if (ctx->tables[i].last_use < ctx->tables[lru].last_use) lru = i;
            size_t newer = 0;
            for (auto &t : ctx->tables) newer += t.last_use > ctx->tables[lru].last_use;
            if (newer < 4) break;
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->tables[lru].d);
            total -= (size_t)32 << ctx->tables[lru].log_n;
            ctx->tables.erase(ctx->tables.begin() + lru);
        }
    }
    dbg_check("get_table before cudaMalloc");
    TwTable t;
    t.root = w;
    t.log_n = log_n;
    t.last_use = ++ctx->table_clock;
    CU(cudaMalloc(&t.d, ((size_t)32) << log_n));
    int rc = powers_into(ctx, w, (size_t)1 << log_n, t.d);
    if (rc != SB_OK) {
        cudaFree(t.d);
        return rc;
    }
    ctx->tables.push_back(t);
    *tw = t.d;
    *tw_log_n = log_n;
    *log_stride = 0;
    return SB_OK;
}

// ------------------------------------------------------------------------------------------------
// NTT
// ------------------------------------------------------------------------------------------------
static int plan_bits(uint32_t log_n, uint32_t *bits) {
    if (log_n == 0) {
        bits[0] = 0;
        return 1;
    }
    const uint32_t maxb = 8;                    // 2^8 rows x 4 contiguous columns (128 B) at the widest
    if (const char *e = getenv("SB_NTT_PLAN")) {   // experiment hook: "21:8,8,5" forces the passes of 2^21-point transforms
        unsigned ln = 0, b[NTT_MAX_PASSES] = {0};
        int cnt = sscanf(e, "%u:%u,%u,%u,%u,%u,%u", &ln, &b[0], &b[1], &b[2], &b[3], &b[4], &b[5]) - 1;
        unsigned sum = 0;
        for (int i = 0; i < cnt; i++) sum += b[i];
        if (cnt > 0 && ln == log_n && sum == log_n) {
            for (int i = 0; i < cnt; i++) bits[i] = b[i];
            return cnt;
        }
    }
    int m = (int)((log_n + maxb - 1) / maxb);
    uint32_t base = log_n / m, rem = log_n % m;
    for (int i = 0; i < m; i++) bits[i] = base + ((uint32_t)i < rem ? 1 : 0);
    return m;
}

// cs != NULL: the coset transforms of a low-degree extension (see NttPassParams); the transform's root is then
// W^(2^log_ext) and the table must be the extended domain's (tw of W).
int ntt_dev_tw(sb_ctx *ctx, const uint4 *d_src, size_t len_in, size_t src_stride, uint4 *d_dst, size_t dst_stride,
               size_t n_polys, uint32_t log_n, int inverse, const uint4 *tw, uint32_t tw_log_n, uint32_t log_stride,
               const CosetSpec *cs) {
    const size_t n = (size_t)1 << log_n;
    if (len_in > n) return fail(ctx, SB_ERR_ARG, "vector of %zu elements does not fit a 2^%u transform", len_in, log_n);
    if (n_polys == 0) return SB_OK;
    const uint32_t coset_cnt = cs ? cs->cnt : 0;
    if (cs && (cs->cnt == 0 || cs->r0 + cs->cnt > (1u << cs->log_ext))) return fail(ctx, SB_ERR_ARG, "internal: coset range");
    const size_t n_batch = coset_cnt ? n_polys * coset_cnt : n_polys;       // transforms in flight
    uint32_t bits[NTT_MAX_PASSES];
    const int m = plan_bits(log_n, bits);
    const uint32_t store = cs ? cs->store : NTT_STORE_PLAIN;
    DevBuf work(ctx);
    if (m > 1) TRY(work.alloc(n_batch * n * 32));
    hfp::el ninv = hfp::inv(hfp::from_u64((uint64_t)n));
    uint32_t log_outer = 0;
    for (int p = 0; p < m; p++) {
        NttPassParams P;
        memset(&P, 0, sizeof P);
        const bool first = p == 0, last = p == m - 1;
        P.src = first ? d_src : (const uint4 *)work.p;
        P.dst = last ? d_dst : (uint4 *)work.p;
        P.src_stride = first ? src_stride : n;
        P.dst_stride = last ? dst_stride : n;
        P.tw = tw;
        P.len_in = len_in;
        P.n_cols_total = (unsigned long long)n_batch << (log_n - bits[p]);
        P.log_n = log_n;
        P.log_outer = log_outer;
        P.log_inner = log_n - log_outer - bits[p];
        P.first = first;
        P.last = last;
        P.inverse = inverse ? 1 : 0;
        P.tw_log_n = tw_log_n;
        P.tw_log_stride = log_stride;
        if (cs) {
            P.coset_cnt = cs->cnt;
            P.coset_r0 = cs->r0;
            P.coset_log = cs->log_ext;
            P.coset_store = store;
            P.coset_dst_cpd = cs->dst_cpd ? cs->dst_cpd : cs->cnt;
            P.coset_dst_r0 = cs->dst_cpd ? cs->dst_r0 : cs->r0;
        }
        {   // interleave polynomials when a tile never straddles two of them (the interleaved coset store also in the last
            // pass: the CTAs that fill the same output lines then run together)
            const unsigned long long cpp = 1ull << (log_n - bits[p]), cc = (1ull << NTT_LOG_TILE_FOR(bits[p])) >> bits[p];
            const bool il_last = last && store == NTT_STORE_INTERLEAVED;
            P.n_polys = (n_batch > 1 && (!last || il_last) && cpp % cc == 0 && n_batch < (1u << 20)) ? (uint32_t)n_batch : 0;
        }
        P.n_prev = (uint32_t)p;
        for (int i = 0; i < p; i++) P.prev_bits[i] = bits[i];
        memcpy(P.n_inv, ninv.l, 32);
        if (bits[p] > 8) return fail(ctx, SB_ERR_ARG, "internal: pass width %u", bits[p]);
        KLAUNCH(SB_KIND_NTT_PASS, ntt_launch_pass(ctx->stream, bits[p], P));
        log_outer += bits[p];
    }
    CU(cudaGetLastError());
    return SB_OK;
}

int ntt_dev(sb_ctx *ctx, const uint4 *d_src, size_t len_in, size_t src_stride, uint4 *d_dst, size_t dst_stride,
                   size_t n_polys, const hfp::el &root, uint32_t log_n, int inverse) {
    const size_t n = (size_t)1 << log_n;
    if (len_in > n) return fail(ctx, SB_ERR_ARG, "vector of %zu elements does not fit a 2^%u transform", len_in, log_n);
    if (n_polys == 0) return SB_OK;
    const uint4 *tw;
    uint32_t tw_log_n, log_stride;
    TRY(get_table(ctx, root, log_n, &tw, &tw_log_n, &log_stride));
    return ntt_dev_tw(ctx, d_src, len_in, src_stride, d_dst, dst_stride, n_polys, log_n, inverse, tw, tw_log_n, log_stride, nullptr);
}

extern "C" int sb_ntt_dev(sb_ctx *ctx, const uint64_t *d_src, size_t len_in, size_t src_stride, uint64_t *d_dst,
                          size_t dst_stride, size_t n_polys, const uint64_t root[4], uint32_t log_n, int inverse) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !d_src || !d_dst || !root) return SB_ERR_ARG;
    if (d_src == d_dst) return fail(ctx, SB_ERR_ARG, "sb_ntt_dev is out of place");
    return ntt_dev(ctx, (const uint4 *)d_src, len_in, src_stride, (uint4 *)d_dst, dst_stride, n_polys, hfp::from_limbs(root),
                   log_n, inverse);
    });
}

// ---- ONE transform over the devices of a multi-device context -----------------------------------------------
// Device d holds elements [d n/g, (d+1) n/g) of the vector in slabs[d] (natural order) on entry and the same range of the result
// on return.  The passes are those of the single-device transform on the virtual array "element e lives on device
// e >> log2(n/g)"; what a four-step formulation does as three separate all-to-all exchanges with transposes in between happens
// inside the pass kernels' own loads and stores over peer memory (ntt_pass_kernel<.., DIST>):
//   pass 0        every device takes a contiguous share of the tiles; a tile's rows are spread over all devices -> remote reads
//                 (exchange 1), and its outputs k1 belong to the device that owns the k1 range -> remote writes (exchange 2)
//   passes 1..m-2 work inside one k1 range: local, in place
//   last pass     device d takes the tiles whose inputs it holds; outputs go to their natural-order owner -> remote writes (3)
// always in runs of CC x 32 contiguous bytes (256-512 B).  Cross-device barriers (events) sit before pass 0, after it and at the end.
static const uint32_t NTT_MULTI_MIN_LOG_N = 20;

static int cross_barrier(sb_ctx *ctx) {
    const int g = ctx->n_dev();
    cudaEvent_t ev[SB_MAX_DEV] = {0};
    int rc = SB_OK;
    for (int d = 0; d < g && rc == SB_OK; d++) {
        DevGuard dg(ctx->dev[d]);
        if (cudaEventCreateWithFlags(&ev[d], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(ev[d], ctx->dev[d]->stream) != cudaSuccess)
            rc = fail(ctx, SB_ERR_CUDA, "barrier event on device %d: %s", d, cudaGetErrorString(cudaGetLastError()));
    }
    for (int d = 0; d < g && rc == SB_OK; d++) {
        DevGuard dg(ctx->dev[d]);
        for (int o = 0; o < g; o++)
            if (o != d && cudaStreamWaitEvent(ctx->dev[d]->stream, ev[o], 0) != cudaSuccess)
                rc = fail(ctx, SB_ERR_CUDA, "barrier wait on device %d: %s", d, cudaGetErrorString(cudaGetLastError()));
    }
    for (int d = 0; d < g; d++)
        if (ev[d]) cudaEventDestroy(ev[d]);           // released once the waits have passed
    return rc;
}

static int launch_pass_on(sb_ctx *ctx, uint32_t bits, const NttPassParams &P) {
    KLAUNCH(SB_KIND_NTT_PASS, ntt_launch_pass(ctx->stream, bits, P));
    CU(cudaGetLastError());
    return SB_OK;
}

int ntt_multi(sb_ctx *ctx, uint4 *const *slabs, const hfp::el &root, uint32_t log_n, int inverse) {
    NvtxRange nvtx("ntt_multi");
    const int g = ctx->n_dev();
    const uint32_t lg = ilog2((size_t)g);
    if (g < 2 || (1 << lg) != g) return fail(ctx, SB_ERR_ARG, "a transform over several devices needs 2, 4 or 8 of them, got %d", g);
    if (log_n < NTT_MULTI_MIN_LOG_N || log_n > 28) return fail(ctx, SB_ERR_ARG, "transforms over several devices: 2^%u .. 2^28 points, got 2^%u", NTT_MULTI_MIN_LOG_N, log_n);
    const size_t n = (size_t)1 << log_n;
    uint32_t bits[NTT_MAX_PASSES];
    const int m = plan_bits(log_n, bits);
    auto log_cc = [&](int p) { return (uint32_t)NTT_LOG_TILE_FOR(bits[p]) - bits[p]; };
    for (int p = 0; p < m; p++)
        if (bits[p] < 6 || bits[p] > 8) return fail(ctx, SB_ERR_ARG, "internal: pass width %u has no multi-device kernel", bits[p]);
    if (m < 2 || bits[0] < lg + log_cc(m - 1)) return fail(ctx, SB_ERR_ARG, "internal: pass plan of 2^%u does not split over %d devices", log_n, g);
    const uint4 *tw[SB_MAX_DEV];
    uint4 *work[SB_MAX_DEV] = {0};
    uint32_t tw_log_n[SB_MAX_DEV] = {0}, log_stride[SB_MAX_DEV] = {0};   // per device: a device may serve the root from a larger cached table
    struct Release {
        sb_ctx *ctx;
        uint4 **w;
        ~Release() {
            for (int d = 0; d < ctx->n_dev(); d++)
                if (w[d]) {
                    DevGuard dg(ctx->dev[d]);
                    blk_release(ctx->dev[d], w[d]);       // stream ordered: the device's queued passes still own the block
                }
        }
    } release{ctx, work};
    for (int d = 0; d < g; d++) {
        sb_ctx *c = ctx->dev[d];
        DevGuard dg(c);
        int rc = get_table(c, root, log_n, &tw[d], &tw_log_n[d], &log_stride[d]);
        if (rc == SB_OK) rc = blk_alloc(c, (n >> lg) * 32, (void **)&work[d]);
        if (rc != SB_OK) {
            if (c != ctx) fail(ctx, rc, "%s", c->err);
            return rc;
        }
    }
    const hfp::el ninv = hfp::inv(hfp::from_u64((uint64_t)n));
    TRY(cross_barrier(ctx));                       // every slab is complete before a peer reads it
    uint32_t log_outer = 0;
    for (int p = 0; p < m; p++) {
        const bool first = p == 0, last = p == m - 1;
        const uint32_t log_tiles = log_n - bits[p] - log_cc(p);
        for (int d = 0; d < g; d++) {
            sb_ctx *c = ctx->dev[d];
            DevGuard dg(c);
            NttPassParams P;
            memset(&P, 0, sizeof P);
            for (int o = 0; o < g; o++) {
                P.src_tab[o] = first ? slabs[o] : work[o];
                P.dst_tab[o] = last ? slabs[o] : work[o];
            }
            P.tw = tw[d];
            P.len_in = n;
            P.n_cols_total = (unsigned long long)n >> bits[p];
            P.log_n = log_n;
            P.log_outer = log_outer;
            P.log_inner = log_n - log_outer - bits[p];
            P.first = first;
            P.last = last;
            P.inverse = inverse ? 1 : 0;
            P.tw_log_n = tw_log_n[d];
            P.tw_log_stride = log_stride[d];
            P.n_prev = (uint32_t)p;
            for (int i = 0; i < p; i++) P.prev_bits[i] = bits[i];
            memcpy(P.n_inv, ninv.l, 32);
            P.dist_log_slab = log_n - lg;
            P.dist_log_g = lg;
            P.dist_dev = (uint32_t)d;
            // last pass: the inputs of output column gl sit at digitrev_inv(gl), whose top digit is gl's low bits[0] bits -- the
            // owner is bits [bits[0] - lg, bits[0]) of gl; earlier passes: the owner is the top of the tile index
            P.dist_blk_lo = last ? bits[0] - lg - log_cc(p) : log_tiles - lg;
            int rc = launch_pass_on(c, bits[p], P);
            if (rc != SB_OK) {
                if (c != ctx) fail(ctx, rc, "%s", c->err);
                return rc;
            }
        }
        if (first || last) TRY(cross_barrier(ctx));
        log_outer += bits[p];
    }
    return SB_OK;
}

extern "C" int sb_ntt_multi_dev(sb_ctx *ctx, uint64_t *const *d_slabs, const uint64_t root[4], uint32_t log_n, int inverse) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !d_slabs || !root) return SB_ERR_ARG;
    for (int d = 0; d < ctx->n_dev(); d++)
        if (!d_slabs[d]) return fail(ctx, SB_ERR_ARG, "slab %d is NULL", d);
    return ntt_multi(ctx, (uint4 *const *)d_slabs, hfp::from_limbs(root), log_n, inverse);
    });
}

extern "C" int sb_dev_alloc_on(sb_ctx *ctx, int dev_index, size_t bytes, void **p) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !p) return SB_ERR_ARG;
    if (dev_index < 0 || dev_index >= ctx->n_dev()) return fail(ctx, SB_ERR_ARG, "device index %d outside the context's %d devices", dev_index, ctx->n_dev());
    DevGuard dg(ctx->dev[dev_index]);
    CU(cudaMalloc(p, bytes ? bytes : 16));
    return SB_OK;
    });
}

// four-step twiddle step of a transform split over several GPUs (sharded.py::distributed_ntt)
extern "C" int sb_twiddle_mul_dev(sb_ctx *ctx, uint64_t *d_vals, size_t rows, size_t cols, size_t row0, const uint64_t root[4],
                                  uint32_t log_n, int inverse) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !d_vals || !root) return SB_ERR_ARG;
    if (log_n > 28) return fail(ctx, SB_ERR_ARG, "log_n %u exceeds the field's two-adicity 28", log_n);
    const uint4 *tw;
    uint32_t tw_log_n, log_stride;
    TRY(get_table(ctx, hfp::from_limbs(root), log_n, &tw, &tw_log_n, &log_stride));
    KLAUNCH(SB_KIND_OTHER, twiddle_mul_launch(ctx->stream, (uint4 *)d_vals, rows, cols, row0, tw, tw_log_n, log_stride, log_n, inverse ? 1 : 0));
    CU(cudaGetLastError());
    return SB_OK;
    });
}

extern "C" int sb_ntt(sb_ctx *ctx, uint64_t *vals, size_t len_in, const uint64_t root[4], uint32_t log_n, int inverse) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !vals || !root) return SB_ERR_ARG;
    if (log_n > 28) return fail(ctx, SB_ERR_ARG, "log_n %u exceeds the field's two-adicity 28", log_n);
    const size_t n = (size_t)1 << log_n;
    if (len_in > n) return fail(ctx, SB_ERR_ARG, "vector of %zu elements does not fit a 2^%u transform", len_in, log_n);
    if (ctx->n_dev() > 1 && log_n >= NTT_MULTI_MIN_LOG_N) {
        // several devices: slab d of the vector goes to device d (the uploads and downloads of the slabs run on the devices'
        // own PCIe links side by side when `vals` is pinned), one transform over all of them (ntt_multi)
        const int g = ctx->n_dev();
        const size_t slab = n / g;
        uint4 *slabs[SB_MAX_DEV] = {0};
        struct Release {
            sb_ctx *ctx;
            uint4 **s;
            ~Release() {
                for (int d = 0; d < ctx->n_dev(); d++)
                    if (s[d]) {
                        DevGuard dg(ctx->dev[d]);
                        blk_release(ctx->dev[d], s[d]);
                    }
            }
        } release{ctx, slabs};
        for (int d = 0; d < g; d++) {
            sb_ctx *c = ctx->dev[d];
            DevGuard dg(c);
            int rc = blk_alloc(c, slab * 32, (void **)&slabs[d]);
            if (rc != SB_OK) return c != ctx ? fail(ctx, rc, "%s", c->err) : rc;
            const size_t lo = (size_t)d * slab, have = len_in > lo ? std::min(len_in - lo, slab) : 0;
            if (have) CU(cudaMemcpyAsync(slabs[d], vals + 4 * lo, have * 32, cudaMemcpyHostToDevice, c->stream));
            if (have < slab) CU(cudaMemsetAsync((uint8_t *)slabs[d] + have * 32, 0, (slab - have) * 32, c->stream));   // fft.rs:335-338
        }
        TRY(ntt_multi(ctx, slabs, hfp::from_limbs(root), log_n, inverse));
        for (int d = 0; d < g; d++) {
            sb_ctx *c = ctx->dev[d];
            DevGuard dg(c);
            CU(cudaMemcpyAsync(vals + 4 * (size_t)d * slab, slabs[d], slab * 32, cudaMemcpyDeviceToHost, c->stream));
        }
        for (int d = 0; d < g; d++) {
            DevGuard dg(ctx->dev[d]);
            CU(cudaStreamSynchronize(ctx->dev[d]->stream));
        }
        return SB_OK;
    }
    DevBuf a(ctx), b(ctx);
    TRY(a.alloc((len_in ? len_in : 1) * 32));
    TRY(b.alloc(n * 32));
    CU(cudaMemcpyAsync(a.p, vals, len_in * 32, cudaMemcpyHostToDevice, ctx->stream));
    TRY(ntt_dev(ctx, (const uint4 *)a.p, len_in, n, (uint4 *)b.p, n, 1, hfp::from_limbs(root), log_n, inverse));
    CU(cudaMemcpyAsync(vals, b.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
    });
}

// Low-degree extension (prove.rs:100-124: best_fft(inv_best_fft(col, W^E, log_s), W, log_s + log_ext), E = 2^log_ext).
// The zero-padded forward transform is computed coset by coset: out[E k + r] = sum_j (c_j W^(j r)) (W^E)^(j k), i.e. E - 1
// transforms of S points on pre-scaled coefficients; coset r = 0 is the input column itself (same field elements, the
// arithmetic is exact), so it is copied.  This skips the three degenerate butterfly stages and one eighth of the rest.
int lde_dev(sb_ctx *ctx, const uint4 *d_cols, size_t n_cols, size_t col_len, size_t col_stride,
                   const hfp::el &root_big, uint32_t log_s, uint32_t log_ext, uint4 *d_out) {
    NvtxRange nvtx("lde_dev");
    if (log_s + log_ext > 28) return fail(ctx, SB_ERR_ARG, "extended domain 2^%u exceeds two-adicity 28", log_s + log_ext);
    const size_t S = (size_t)1 << log_s, N = S << log_ext;
    if (col_len > S) return fail(ctx, SB_ERR_ARG, "column of %zu elements does not fit 2^%u", col_len, log_s);
    if (n_cols == 0) return SB_OK;
    // the extended domain's table serves both transforms: W^E with stride E, and the coset scaling W^(j r)
    const uint4 *tw;
    uint32_t tw_log_n, log_stride;
    TRY(get_table(ctx, root_big, log_s + log_ext, &tw, &tw_log_n, &log_stride));
    if (log_ext == 0 || log_ext > 8) {
        hfp::el root_small = root_big;
        for (uint32_t i = 0; i < log_ext; i++) root_small = hfp::sqr(root_small);
        DevBuf coef(ctx);
        TRY(coef.alloc(n_cols * S * 32));
        TRY(ntt_dev(ctx, d_cols, col_len, col_stride, (uint4 *)coef.p, S, n_cols, root_small, log_s, 1));
        return ntt_dev(ctx, (const uint4 *)coef.p, S, S, d_out, N, n_cols, root_big, log_s + log_ext, 0);
    }
    DevBuf coef(ctx);
    TRY(coef.alloc(n_cols * S * 32));
    TRY(ntt_dev_tw(ctx, d_cols, col_len, col_stride, (uint4 *)coef.p, S, n_cols, log_s, 1, tw, tw_log_n, log_stride + log_ext, nullptr));
    // (Tried: one CTA holding the same tile of all eight cosets so that every output row is 256 contiguous bytes and the copy
    // kernel disappears.  Measured slower on B200, LDE 2^21 -> 2^24 x 10: 35.8 against 33.5 ms -- the eight-column tile carries
    // seven transforms, so every butterfly round runs with 1/8 of its lanes idle, which costs more than the stores gain.)
    CosetSpec cs;
    cs.log_ext = log_ext;
    cs.r0 = 1;
    cs.cnt = (1u << log_ext) - 1;
    cs.store = NTT_STORE_INTERLEAVED;
    KLAUNCH(SB_KIND_OTHER, lde_launch_coset0(ctx->stream, d_cols, col_len, col_stride, d_out, N, S, log_ext, n_cols));
    return ntt_dev_tw(ctx, (const uint4 *)coef.p, S, S, d_out, N, n_cols, log_s, 0, tw, tw_log_n, log_stride + log_ext, &cs);
}

extern "C" int sb_lde_batch_dev(sb_ctx *ctx, const uint64_t *d_cols, size_t n_cols, size_t col_len, size_t col_stride,
                                const uint64_t root_big[4], uint32_t log_s, uint32_t log_ext, uint64_t *d_out) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !d_cols || !d_out || !root_big) return SB_ERR_ARG;
    return lde_dev(ctx, (const uint4 *)d_cols, n_cols, col_len, col_stride, hfp::from_limbs(root_big), log_s, log_ext,
                   (uint4 *)d_out);
    });
}

extern "C" int sb_lde_batch(sb_ctx *ctx, const uint64_t *cols, size_t n_cols, size_t col_len, const uint64_t root_big[4],
                            uint32_t log_s, uint32_t log_ext, uint64_t *out) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !cols || !out || !root_big) return SB_ERR_ARG;
    if (log_s + log_ext > 28) return fail(ctx, SB_ERR_ARG, "extended domain 2^%u exceeds two-adicity 28", log_s + log_ext);
    const size_t N = (size_t)1 << (log_s + log_ext);
    DevBuf in(ctx), o(ctx);
    TRY(in.alloc(n_cols * col_len * 32));
    TRY(o.alloc(n_cols * N * 32));
    const hfp::el root = hfp::from_limbs(root_big);
    const size_t G = 2;                                  // columns per pipeline stage
    if (n_cols < 2 * G || N < ((size_t)1 << 18)) {
        CU(cudaMemcpyAsync(in.p, cols, n_cols * col_len * 32, cudaMemcpyHostToDevice, ctx->stream));
        TRY(lde_dev(ctx, (const uint4 *)in.p, n_cols, col_len, col_len, root, log_s, log_ext, (uint4 *)o.p));
        CU(cudaMemcpyAsync(out, o.p, n_cols * N * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return SB_OK;
    }
    // Large batches: upload of group g+1, transform of group g and download of group g-1 overlap (three streams).  The
    // download (32 N bytes per column) is the long pole over PCIe; the transform hides behind it.  Needs pinned host buffers
    // to overlap (pageable ones are staged by the driver and serialise, still correct).
    if (!ctx->h2d_stream) CU(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    if (!ctx->d2h_stream) CU(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    const size_t n_groups = (n_cols + G - 1) / G;
    std::vector<cudaEvent_t> up(n_groups), done(n_groups);
    cudaEvent_t ready;
    CU(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    CU(cudaEventRecord(ready, ctx->stream));             // the allocations above are ordered on ctx->stream
    CU(cudaStreamWaitEvent(ctx->h2d_stream, ready, 0));
    CU(cudaStreamWaitEvent(ctx->d2h_stream, ready, 0));
    int rc = SB_OK;
    for (size_t g = 0; g < n_groups; g++) {
        cudaEventCreateWithFlags(&up[g], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&done[g], cudaEventDisableTiming);
        const size_t c0 = g * G, nc = std::min(G, n_cols - c0);
        cudaMemcpyAsync((uint8_t *)in.p + c0 * col_len * 32, (const uint8_t *)cols + c0 * col_len * 32, nc * col_len * 32,
                        cudaMemcpyHostToDevice, ctx->h2d_stream);
        cudaEventRecord(up[g], ctx->h2d_stream);
    }
    for (size_t g = 0; g < n_groups && rc == SB_OK; g++) {
        const size_t c0 = g * G, nc = std::min(G, n_cols - c0);
        cudaStreamWaitEvent(ctx->stream, up[g], 0);
        rc = lde_dev(ctx, (const uint4 *)in.p + 2 * c0 * col_len, nc, col_len, col_len, root, log_s, log_ext, (uint4 *)o.p + 2 * c0 * N);
        cudaEventRecord(done[g], ctx->stream);
        cudaStreamWaitEvent(ctx->d2h_stream, done[g], 0);
        cudaMemcpyAsync((uint8_t *)out + c0 * N * 32, (const uint8_t *)o.p + c0 * N * 32, nc * N * 32, cudaMemcpyDeviceToHost, ctx->d2h_stream);
    }
    cudaError_t e1 = cudaStreamSynchronize(ctx->d2h_stream), e2 = cudaStreamSynchronize(ctx->h2d_stream), e3 = cudaStreamSynchronize(ctx->stream);
    for (size_t g = 0; g < n_groups; g++) {
        cudaEventDestroy(up[g]);
        cudaEventDestroy(done[g]);
    }
    cudaEventDestroy(ready);
    if (rc != SB_OK) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return fail(ctx, SB_ERR_CUDA, "sb_lde_batch pipeline: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    return SB_OK;
    });
}

extern "C" int sb_powers_dev(sb_ctx *ctx, const uint64_t root[4], size_t n, uint64_t *d_out) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !root || !d_out) return SB_ERR_ARG;
    return powers_into(ctx, hfp::from_limbs(root), n, (uint4 *)d_out);
    });
}
extern "C" int sb_powers(sb_ctx *ctx, const uint64_t root[4], size_t n, uint64_t *out) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !root || !out) return SB_ERR_ARG;
    DevBuf b(ctx);
    TRY(b.alloc(n * 32));
    TRY(powers_into(ctx, hfp::from_limbs(root), n, (uint4 *)b.p));
    CU(cudaMemcpyAsync(out, b.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
    });
}

extern "C" int sb_batch_inverse_dev(sb_ctx *ctx, uint64_t *d_vals, size_t n) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !d_vals) return SB_ERR_ARG;
    if (n == 0) return SB_OK;
    DevBuf scratch(ctx);
    TRY(scratch.alloc(batch_inverse_scratch_elems(n) * 32));
    KLAUNCH(SB_KIND_OTHER, batch_inverse_launch(ctx->stream, (uint4 *)d_vals, (uint4 *)scratch.p, n));
    CU(cudaGetLastError());
    return SB_OK;
    });
}
extern "C" int sb_batch_inverse(sb_ctx *ctx, uint64_t *vals, size_t n) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !vals) return SB_ERR_ARG;
    DevBuf b(ctx);
    TRY(b.alloc(n * 32));
    CU(cudaMemcpyAsync(b.p, vals, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    TRY(sb_batch_inverse_dev(ctx, (uint64_t *)b.p, n));
    CU(cudaMemcpyAsync(vals, b.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
    });
}

// ------------------------------------------------------------------------------------------------
// Merkle
// ------------------------------------------------------------------------------------------------
void free_tree(sb_tree *t) {
    if (!t) return;
    if (t->sh) {
        for (int d = 0; d < t->sh->g; d++) {
            if (!t->sh->ctxs[d]) continue;
            DevGuard g(t->sh->ctxs[d]);
            blk_release(t->sh->ctxs[d], t->sh->low[d]);
            blk_release(t->sh->ctxs[d], t->sh->sub[d]);
            blk_release(t->sh->ctxs[d], t->sh->stage[d]);
            blk_release(t->sh->ctxs[d], t->sh->recv[d]);
        }
        delete t->sh;
    }
    DevGuard g(t->device);
    if (t->d_nodes) {
        if (t->owner) blk_release(t->owner, t->d_nodes);
        else cudaFreeAsync(t->d_nodes, t->stream);
    }
    if (t->d_leaves) {
        if (t->owner) blk_release(t->owner, t->d_leaves);      // falls back to cudaFreeAsync for blocks the cache does not know
        else cudaFreeAsync(t->d_leaves, t->stream);
    }
    delete t;
}

static void launch_leaves(sb_ctx *ctx, sb_tree *t, uint32_t lv) {
    if (t->n_cols) {
        MerkleColsParams P;
        for (int k = 0; k < 8; k++) P.cols[k] = t->cols[k];
        P.nodes = t->d_nodes;
        P.n = t->n;
        P.nc = (uint32_t)t->n_cols;
        P.coset_log_s = 0;
        KLAUNCH(SB_KIND_MERKLE_LEAVES, merkle_launch_leaves_cols(ctx->stream, lv, P));
    } else {
        MerkleBytesParams P;
        P.leaves = t->d_leaves;
        P.nodes = t->d_nodes;
        P.n = t->n;
        P.leaf_bytes = (uint32_t)t->leaf_bytes;
        P.first = 0;
        P.count = t->n;
        KLAUNCH(SB_KIND_MERKLE_LEAVES, merkle_launch_leaves_bytes(ctx->stream, lv, P));
    }
}

// levels above `level` (already present in t->d_nodes); fetch_root: leaves the root in t->root (one 32-byte D2H + sync)
int merkle_finish(sb_ctx *ctx, sb_tree *t, uint32_t level, bool fetch_root) {
    while (level < t->depth) {
        const uint32_t lv = t->depth - level < 3 ? t->depth - level : 3;
        KLAUNCH(SB_KIND_MERKLE_NODES, merkle_launch_nodes(ctx->stream, lv, t->d_nodes, t->n, level));
        level += lv;
    }
    CU(cudaGetLastError());
    if (fetch_root) {
        CU(cudaMemcpyAsync(t->root, (const uint8_t *)t->d_nodes + (2 * t->n - 2) * 32, 32, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return SB_OK;
}

// hashes the leaves and every level; leaves the root in t->root.
// fold != NULL: the leaves are produced by the FRI fold of *fold (fused kernel), which also writes the column t->cols[0].
static int merkle_build(sb_ctx *ctx, sb_tree *t, const FriFoldParams *fold = nullptr, bool fetch_root = true) {
    uint32_t level = t->depth < 3 ? t->depth : 3;
    if (fold) {
        KLAUNCH(SB_KIND_FRI_FOLD, merkle_launch_leaves_fold(ctx->stream, level, *fold, t->d_nodes));
    } else {
        launch_leaves(ctx, t, level);
    }
    return merkle_finish(ctx, t, level, fetch_root);
}

int tree_new(sb_ctx *ctx, size_t n, size_t leaf_bytes, sb_tree **out) {
    if (!is_pow2(n)) return fail(ctx, SB_ERR_ARG, "leaf count %zu is not a power of two", n);   // merkle_proof_in_place.rs:113
    if (n > ((size_t)1 << 30)) return fail(ctx, SB_ERR_ARG, "tree too large");
    sb_tree *t = new sb_tree();
    t->n = n;
    t->depth = ilog2(n);
    t->leaf_bytes = leaf_bytes;
    t->stream = ctx->stream;
    t->device = ctx->device;
    t->owner = ctx;
    int rc = blk_alloc(ctx, (2 * n - 1) * 32, (void **)&t->d_nodes);
    if (rc != SB_OK) {
        delete t;
        return rc;
    }
    *out = t;
    return SB_OK;
}

extern "C" int sb_merkle_commit(sb_ctx *ctx, const void *leaves, size_t leaf_bytes, size_t n, uint8_t root[32], sb_tree **tree) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !tree) return SB_ERR_ARG;
    if (!leaves && n * leaf_bytes) return fail(ctx, SB_ERR_ARG, "leaves is NULL");
    if (leaf_bytes >= ((size_t)1 << 31)) return fail(ctx, SB_ERR_ARG, "leaf too long");
    sb_tree *t = nullptr;
    TRY(tree_new(ctx, n, leaf_bytes, &t));
    cudaError_t e = cudaSuccess;
    {
        int rc0 = blk_alloc(ctx, n * leaf_bytes, (void **)&t->d_leaves);
        if (rc0 != SB_OK) {
            free_tree(t);
            return rc0;
        }
    }
    int rc = SB_OK;
    const size_t total = n * leaf_bytes;
    const size_t CHUNK_BYTES = (size_t)64 << 20;
    if (total >= 4 * CHUNK_BYTES && t->depth >= 13) {
        // Large trees: the upload of chunk k + 1 overlaps the leaf hashing (+ 3 levels) of chunk k (two streams); only the last
        // chunk's hashing and the upper levels remain after the transfer.  Chunks are multiples of 1024 leaves (one CTA's share).
        size_t chunk = CHUNK_BYTES / leaf_bytes;
        chunk = chunk < 1024 ? 1024 : (chunk & ~(size_t)1023);
        if (!ctx->h2d_stream) CU(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
        cudaEvent_t ready, up;
        CU(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        CU(cudaEventRecord(ready, ctx->stream));             // the allocations above are ordered on ctx->stream
        CU(cudaStreamWaitEvent(ctx->h2d_stream, ready, 0));
        std::vector<cudaEvent_t> ups;
        for (size_t first = 0; first < n && e == cudaSuccess; first += chunk) {
            const size_t cnt = std::min(chunk, n - first);
            e = cudaMemcpyAsync(t->d_leaves + first * leaf_bytes, (const uint8_t *)leaves + first * leaf_bytes, cnt * leaf_bytes, cudaMemcpyHostToDevice, ctx->h2d_stream);
            if (e != cudaSuccess) break;
            cudaEventCreateWithFlags(&up, cudaEventDisableTiming);
            cudaEventRecord(up, ctx->h2d_stream);
            ups.push_back(up);
            cudaStreamWaitEvent(ctx->stream, up, 0);
            MerkleBytesParams P;
            P.leaves = t->d_leaves;
            P.nodes = t->d_nodes;
            P.n = t->n;
            P.leaf_bytes = (uint32_t)leaf_bytes;
            P.first = first;
            P.count = cnt;
            KLAUNCH(SB_KIND_MERKLE_LEAVES, merkle_launch_leaves_bytes(ctx->stream, 3, P));
        }
        if (e == cudaSuccess) rc = merkle_finish(ctx, t, 3, true);
        else rc = fail(ctx, SB_ERR_CUDA, "H2D leaves: %s", cudaGetErrorString(e));
        cudaStreamSynchronize(ctx->h2d_stream);
        for (auto ev : ups) cudaEventDestroy(ev);
        cudaEventDestroy(ready);
    } else {
        e = total ? cudaMemcpyAsync(t->d_leaves, leaves, total, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
        rc = e == cudaSuccess ? merkle_build(ctx, t) : fail(ctx, SB_ERR_CUDA, "H2D leaves: %s", cudaGetErrorString(e));
    }
    if (rc != SB_OK) {
        free_tree(t);
        return rc;
    }
    if (root) memcpy(root, t->root, 32);
    *tree = t;
    return SB_OK;
    });
}

// byte leaves already on the device; the tree takes ownership of d_leaves (allocated with cudaMallocAsync)
int commit_bytes_owned(sb_ctx *ctx, uint8_t *d_leaves, size_t leaf_bytes, size_t n, sb_tree **tree) {
    sb_tree *t = nullptr;
    int rc = tree_new(ctx, n, leaf_bytes, &t);
    if (rc != SB_OK) {
        cudaFreeAsync(d_leaves, ctx->stream);
        return rc;
    }
    t->d_leaves = d_leaves;
    rc = merkle_build(ctx, t);
    if (rc != SB_OK) {
        free_tree(t);
        return rc;
    }
    *tree = t;
    return SB_OK;
}

int commit_cols(sb_ctx *ctx, const uint4 *const *d_cols, size_t n_cols, size_t n, sb_tree **tree) {
    NvtxRange nvtx("commit_cols");
    if (n_cols < 1 || n_cols > 8) return fail(ctx, SB_ERR_ARG, "1..8 columns per leaf supported, got %zu", n_cols);
    sb_tree *t = nullptr;
    TRY(tree_new(ctx, n, 32 * n_cols, &t));
    t->n_cols = (int)n_cols;
    for (size_t k = 0; k < n_cols; k++) t->cols[k] = d_cols[k];
    int rc = merkle_build(ctx, t);
    if (rc != SB_OK) {
        free_tree(t);
        return rc;
    }
    *tree = t;
    return SB_OK;
}

// FRI: column = fold(values) and the tree over the column in one pass over the data (fri.rs:141-172)
// fetch_root = false: everything is only queued; the root stays in the node array until the caller reads it
static int commit_fold(sb_ctx *ctx, const FriFoldParams &F, sb_tree **tree, bool fetch_root = true) {
    const size_t q = F.n >> 2;
    sb_tree *t = nullptr;
    TRY(tree_new(ctx, q, 32, &t));
    t->n_cols = 1;
    t->cols[0] = F.col;
    int rc = merkle_build(ctx, t, &F, fetch_root);
    if (rc != SB_OK) {
        free_tree(t);
        return rc;
    }
    *tree = t;
    return SB_OK;
}

extern "C" int sb_merkle_commit_cols_dev(sb_ctx *ctx, const uint64_t *const *d_cols, size_t n_cols, size_t n, uint8_t root[32],
                                         sb_tree **tree) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !d_cols || !tree) return SB_ERR_ARG;
    TRY(commit_cols(ctx, (const uint4 *const *)d_cols, n_cols, n, tree));
    if (root) memcpy(root, (*tree)->root, 32);
    return SB_OK;
    });
}

// gen_proofs for several (tree, index list) pairs with ONE device round trip: every gather kernel is queued, the results of all
// requests come back in one copy into pinned staging memory, one synchronisation.  (A FRI layer needs the openings of two trees and
// the prover those of m_tree and l_tree; one call each used to cost a synchronisation -- and the copy into the caller's pageable
// memory a second, implicit one.)  The staging buffer is the context's pinned scratch: no other user may have a transfer in flight.
int merkle_open_many(sb_ctx *ctx, const OpenReq *reqs, int n_req) {
    struct Lay { size_t idx_off, nodes_off, leaves_off, nodes_bytes, leaves_bytes; uint32_t dlow; };
    std::vector<Lay> lay(n_req);
    size_t total = 0;
    for (int r = 0; r < n_req; r++) {
        const OpenReq &R = reqs[r];
        if (!R.t || (!R.idx && R.n_idx)) return SB_ERR_ARG;
        if (R.n_idx > ((size_t)1 << 24)) return fail(ctx, SB_ERR_ARG, "too many openings");
        for (size_t i = 0; i < R.n_idx; i++)
            if (R.idx[i] >= R.t->n) return fail(ctx, SB_ERR_ARG, "leaf index %zu out of range (width %zu)", R.idx[i], R.t->n);
        Lay &L = lay[r];
        L.dlow = R.t->sh ? R.t->sh->lv + R.t->sh->log_s : R.t->depth;          // levels that live on the device(s)
        L.nodes_bytes = R.nodes_out ? R.n_idx * L.dlow * 32 : 0;
        L.leaves_bytes = R.leaves_out ? R.n_idx * R.t->leaf_bytes : 0;
        L.idx_off = total;
        total += (R.n_idx * 8 + 31) & ~(size_t)31;
    }
    const size_t out_off = total;
    for (int r = 0; r < n_req; r++) {
        lay[r].nodes_off = total;
        total += (lay[r].nodes_bytes + 31) & ~(size_t)31;
        lay[r].leaves_off = total;
        total += (lay[r].leaves_bytes + 31) & ~(size_t)31;
    }
    if (total == 0 || total == out_off) return SB_OK;
    uint8_t *host = (uint8_t *)pinned_scratch(ctx, total);
    if (!host) return fail(ctx, SB_ERR_OOM, "pinned staging for the openings");
    DevBuf dev(ctx);
    TRY(dev.alloc(total));
    uint8_t *d = (uint8_t *)dev.p;
    for (int r = 0; r < n_req; r++) {
        unsigned long long *h = (unsigned long long *)(host + lay[r].idx_off);
        for (size_t i = 0; i < reqs[r].n_idx; i++) h[i] = reqs[r].idx[i];
    }
    CU(cudaMemcpyAsync(d, host, out_off, cudaMemcpyHostToDevice, ctx->stream));
    for (int r = 0; r < n_req; r++) {
        const OpenReq &R = reqs[r];
        const Lay &L = lay[r];
        const sb_tree *t = R.t;
        if (!R.n_idx) continue;
        const unsigned long long *d_idx = (const unsigned long long *)(d + L.idx_off);
        uint4 *d_nodes = L.nodes_bytes ? (uint4 *)(d + L.nodes_off) : nullptr, *d_leaves = L.leaves_bytes ? (uint4 *)(d + L.leaves_off) : nullptr;
        if (t->sh) {
            // sharded tree: one gather kernel on this (the primary) device reads the shards through peer pointers; the top
            // log2 g levels are appended from the host copy
            const TreeShards &sh = *t->sh;
            ExtOpenParams P;
            memset(&P, 0, sizeof P);
            for (int dd = 0; dd < sh.g; dd++) {
                P.low[dd] = sh.low[dd];
                P.sub[dd] = sh.sub[dd];
                for (int c = 0; c < 8; c++) P.cols[dd][c] = sh.cols[dd][c];
            }
            P.nc = (uint32_t)t->n_cols; P.log_s = sh.log_s; P.cpd = sh.cpd; P.lv = sh.lv; P.g = (uint32_t)sh.g;
            if (d_nodes || d_leaves) KLAUNCH(SB_KIND_OPEN, merkle_launch_open_ext(ctx->stream, P, d_idx, (uint32_t)R.n_idx, d_nodes, d_leaves));
            continue;
        }
        if (d_nodes) KLAUNCH(SB_KIND_OPEN, merkle_launch_open(ctx->stream, t->d_nodes, t->n, t->depth, d_idx, (uint32_t)R.n_idx, d_nodes));
        if (d_leaves) {
            if (t->n_cols) {
                MerkleColsParams P;
                for (int k = 0; k < 8; k++) P.cols[k] = t->cols[k];
                P.nodes = nullptr;
                P.n = t->n;
                P.nc = (uint32_t)t->n_cols;
                P.coset_log_s = t->coset_log_s;
                KLAUNCH(SB_KIND_OPEN, merkle_launch_open_leaves_cols(ctx->stream, P, d_idx, (uint32_t)R.n_idx, d_leaves));
            } else {
                KLAUNCH(SB_KIND_OPEN, merkle_launch_gather_bytes(ctx->stream, t->d_leaves, t->leaf_bytes, d_idx, (uint32_t)R.n_idx, (uint8_t *)d_leaves));
            }
        }
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host + out_off, d + out_off, total - out_off, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < n_req; r++) {
        const OpenReq &R = reqs[r];
        const Lay &L = lay[r];
        if (L.leaves_bytes) memcpy(R.leaves_out, host + L.leaves_off, L.leaves_bytes);
        if (!R.nodes_out) continue;
        if (!R.t->sh) {
            if (L.nodes_bytes) memcpy(R.nodes_out, host + L.nodes_off, L.nodes_bytes);
            continue;
        }
        const TreeShards &sh = *R.t->sh;
        for (size_t q = 0; q < R.n_idx; q++) {
            uint8_t *o = R.nodes_out + q * R.t->depth * 32;
            if (L.dlow) memcpy(o, host + L.nodes_off + q * L.dlow * 32, L.dlow * 32);
            for (uint32_t l = L.dlow; l < R.t->depth; l++) {
                const size_t m = (R.idx[q] >> l) ^ 1;
                memcpy(o + l * 32, sh.top.data() + (merkle_level_off((size_t)sh.g, l - L.dlow) + m) * 32, 32);
            }
        }
    }
    return SB_OK;
}

extern "C" int sb_merkle_open(sb_ctx *ctx, const sb_tree *t, const size_t *idx, size_t n_idx, uint8_t *leaves_out, uint8_t *nodes_out) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !t || (!idx && n_idx)) return SB_ERR_ARG;
    if (n_idx == 0) return SB_OK;
    const OpenReq req{t, idx, n_idx, t->leaf_bytes ? leaves_out : nullptr, t->depth ? nodes_out : nullptr};
    return merkle_open_many(ctx, &req, 1);
    });
}

extern "C" size_t sb_tree_width(const sb_tree *t) { return t ? t->n : 0; }
extern "C" size_t sb_tree_leaf_bytes(const sb_tree *t) { return t ? t->leaf_bytes : 0; }
extern "C" int sb_tree_root(const sb_tree *t, uint8_t root[32]) {
    if (!t || !root) return SB_ERR_ARG;
    memcpy(root, t->root, 32);
    return SB_OK;
}
extern "C" void sb_tree_free(sb_ctx *ctx, sb_tree *t) {
    DevGuard g(ctx);
    if (t && ctx && t->device == ctx->device && !t->owner) t->stream = ctx->stream;   // the context's stream may have been replaced since the tree was built (sb_set_stream)
    free_tree(t);      // stream-ordered: work already queued on the stream finishes first
}

// ------------------------------------------------------------------------------------------------
// Fiat-Shamir helpers (host)
// ------------------------------------------------------------------------------------------------
extern "C" void sb_blake2s(const uint8_t *msg, size_t len, uint8_t out[32]) { b2s::hash_bytes(out, msg, len); }

// fri/src/utils.rs:82-109.  `extended`: lift the reference's `modulus < 2^24` assert (utils.rs:88) for domains the
// reference cannot handle (BASELINE config 5, N = 2^26); the rule itself is unchanged and the u32 arithmetic stays
// exact as long as modulus * (excl - 1) < 2^32.
int pseudorandom_indices(const uint8_t *seed, size_t seed_len, uint32_t modulus, size_t count, uint32_t excl, uint32_t *out,
                         bool extended) {
    if (!seed || !out) return SB_ERR_ARG;
    if (modulus == 0) return SB_ERR_ARG;
    if (!extended && modulus >= (1u << 24)) return SB_ERR_ARG;   // utils.rs:88 assert
    if (excl > 1 && (uint64_t)modulus * (excl - 1) >= ((uint64_t)1 << 32)) return SB_ERR_ARG;
    std::vector<uint8_t> data(seed, seed + seed_len);
    while (data.size() < 4 * count) {
        uint8_t d[32];
        if (data.size() < 32) return SB_ERR_ARG;     // reference: slice underflow panic (utils.rs:92)
        b2s::hash_bytes(d, data.data() + data.size() - 32, 32);
        data.insert(data.end(), d, d + 32);
    }
    if (excl == 1 || (excl > 1 && modulus * (excl - 1) / excl == 0)) return SB_ERR_ARG;   // reference: division by zero
    for (size_t i = 0; i < count; i++) {
        uint32_t v = ((uint32_t)data[4 * i] << 24) | ((uint32_t)data[4 * i + 1] << 16) | ((uint32_t)data[4 * i + 2] << 8) | data[4 * i + 3];
        if (excl == 0) {
            out[i] = v % modulus;
        } else {
            uint32_t real = modulus * (excl - 1) / excl;
            uint32_t t = v % real;
            out[i] = t + 1 + t / (excl - 1);
        }
    }
    return SB_OK;
}
extern "C" int sb_pseudorandom_indices(const uint8_t *seed, size_t seed_len, uint32_t modulus, size_t count, uint32_t excl,
                                       uint32_t *out) {
    return pseudorandom_indices(seed, seed_len, modulus, count, excl, out, false);
}
extern "C" int sb_pseudorandom_indices_ctx(const sb_ctx *ctx, const uint8_t *seed, size_t seed_len, uint32_t modulus, size_t count,
                                           uint32_t excl, uint32_t *out) {
    return pseudorandom_indices(seed, seed_len, modulus, count, excl, out, ctx && ctx->extended_domain);
}
extern "C" int sb_set_extended_domain(sb_ctx *ctx, int enable) {
    if (!ctx) return SB_ERR_ARG;
    for (sb_ctx *c : ctx->dev) c->extended_domain = enable != 0;
    return SB_OK;
}

// ------------------------------------------------------------------------------------------------
// FRI
// ------------------------------------------------------------------------------------------------
int fri_prove_dev(sb_ctx *ctx, const uint4 *d_vals, size_t n, const hfp::el &root, size_t max_deg_plus_1, uint32_t excl,
                         const sb_tree *values_tree, sb_fri_proof **out) {
    NvtxRange nvtx("fri_prove_dev");
    if (!is_pow2(n)) return fail(ctx, SB_ERR_ARG, "FRI needs a power-of-two number of values, got %zu", n);
    const uint32_t log_n0 = ilog2(n);
    const uint4 *tw;
    uint32_t tw_log_n, log_stride;
    TRY(get_table(ctx, root, log_n0, &tw, &tw_log_n, &log_stride));

    sb_fri_proof *proof = new sb_fri_proof();
    std::vector<sb_tree *> owned_trees;
    std::vector<void *> owned_cols;
    auto cleanup = [&]() {
        for (auto t : owned_trees) free_tree(t);
        for (auto p : owned_cols) blk_release(ctx, p);
    };
    // All layers are queued back to back: from the second layer on the fold kernel derives special_x from the previous column
    // tree's root where it lies in device memory (FriFoldParams::special_root), so the host does not wait for a root between
    // layers; the roots come back in one pinned buffer, then the host samples every layer's positions and all openings of all
    // layers are one round trip (merkle_open_many).  Before: two host round trips per layer (root, openings).
    int rc = SB_OK;
    const uint4 *cur = d_vals;
    const sb_tree *cur_tree = values_tree;
    size_t cur_n = n, bound = max_deg_plus_1;
    uint32_t cur_stride = log_stride;
    struct Mid {
        const sb_tree *poly_tree;      // tree over the layer's values
        sb_tree *col_tree;             // tree over the folded column (root fetched late)
        size_t q;
    };
    std::vector<Mid> mids;
    FriLayer last;
    DevBuf last_bytes(ctx);
    size_t max_layers = 2;
    for (size_t b = max_deg_plus_1; b > FRI_MIN_DEG_DIRECT; b /= 4) max_layers++;
    uint8_t *h_roots = (uint8_t *)pinned_scratch(ctx, 32 * max_layers);
    if (!h_roots) {
        delete proof;
        return fail(ctx, SB_ERR_OOM, "pinned scratch for the FRI roots");
    }
    while (true) {
        if (bound <= FRI_MIN_DEG_DIRECT) {
            // fri.rs:88-112: the remaining values go into the proof verbatim (to_bytes_le each)
            last.is_last = true;
            last.last.resize(cur_n * 32);
            if ((rc = last_bytes.alloc(cur_n * 32)) != SB_OK) break;
            KLAUNCH(SB_KIND_OTHER, fp_launch_to_bytes(ctx->stream, cur, (uint4 *)last_bytes.p, cur_n));
            break;
        }
        if (cur_n < 4 || (cur_n / 4 >= (1u << 24) && !(ctx->extended_domain && cur_n / 4 <= (1u << 28)))) { rc = fail(ctx, SB_ERR_ARG, "FRI layer of %zu values unsupported", cur_n); break; }
        // fri.rs:120-131: tree over the values (reused when the caller / previous layer built it)
        if (!cur_tree) {
            sb_tree *t = nullptr;
            const uint4 *cols[1] = {cur};
            if ((rc = commit_cols(ctx, cols, 1, cur_n, &t)) != SB_OK) break;
            owned_trees.push_back(t);
            cur_tree = t;
        }
        // fri.rs:141-164
        const size_t q = cur_n / 4;
        void *d_col = nullptr;
        if ((rc = blk_alloc(ctx, q * 32, &d_col)) != SB_OK) break;
        owned_cols.push_back(d_col);
        FriFoldParams P;
        memset(&P, 0, sizeof P);
        P.vals = cur;
        P.col = (uint4 *)d_col;
        P.tw = tw;
        P.n = cur_n;
        P.tw_log_n = tw_log_n;
        P.tw_log_stride = cur_stride;
        if (mids.empty()) {
            const hfp::el special_x = hfp::from_bytes_le32(cur_tree->root);      // fri.rs:135: the first tree's root is on the host
            memcpy(P.special_x, special_x.l, 32);
        } else {
            P.special_root = cur_tree->d_nodes + 2 * (2 * cur_n - 2);            // previous layer's column tree (this function built it)
        }
        // fri.rs:141-172: fold and commit the column in one kernel
        sb_tree *t2 = nullptr;
        if ((rc = commit_fold(ctx, P, &t2, false)) != SB_OK) break;
        owned_trees.push_back(t2);
        cudaError_t e = cudaMemcpyAsync(h_roots + 32 * mids.size(), (const uint8_t *)t2->d_nodes + (2 * q - 2) * 32, 32, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) { rc = fail(ctx, SB_ERR_CUDA, "FRI root: %s", cudaGetErrorString(e)); break; }
        mids.push_back(Mid{cur_tree, t2, q});
        // fri.rs:215-223
        cur = (const uint4 *)d_col;
        cur_tree = t2;
        cur_n = q;
        bound /= 4;
        cur_stride += 2;
    }
    if (rc == SB_OK) {
        cudaError_t e = last.is_last ? cudaMemcpyAsync(last.last.data(), last_bytes.p, last.last.size(), cudaMemcpyDeviceToHost, ctx->stream) : cudaSuccess;
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, SB_ERR_CUDA, "FRI layers: %s", cudaGetErrorString(e));
    }
    if (rc == SB_OK) {
        // fri.rs:181-204 for every layer, then all openings at once
        proof->layers.resize(mids.size());
        std::vector<std::vector<size_t>> idx(2 * mids.size());
        std::vector<OpenReq> reqs;
        for (size_t k = 0; k < mids.size() && rc == SB_OK; k++) {
            const Mid &M = mids[k];
            memcpy(M.col_tree->root, h_roots + 32 * k, 32);
            FriLayer &L = proof->layers[k];
            memcpy(L.values_root, M.poly_tree->root, 32);
            memcpy(L.root2, M.col_tree->root, 32);
            uint32_t ys[FRI_QUERIES];
            if (pseudorandom_indices(M.col_tree->root, 32, (uint32_t)M.q, FRI_QUERIES, excl, ys, ctx->extended_domain) != SB_OK) {
                rc = fail(ctx, SB_ERR_ARG, "sampler: column length %zu out of range", M.q);
                break;
            }
            std::vector<size_t> &yi = idx[2 * k], &pp = idx[2 * k + 1];
            yi.resize(FRI_QUERIES);
            pp.resize(4 * FRI_QUERIES);
            for (size_t i = 0; i < FRI_QUERIES; i++) {
                yi[i] = ys[i];
                for (size_t j = 0; j < 4; j++) pp[4 * i + j] = ys[i] + M.q * j;     // fri.rs:193-204
            }
            L.n_column = FRI_QUERIES;
            L.depth_column = M.col_tree->depth;
            L.column_leaves.resize(FRI_QUERIES * 32);
            L.column_nodes.resize(FRI_QUERIES * M.col_tree->depth * 32);
            L.n_poly = 4 * FRI_QUERIES;
            L.depth_poly = M.poly_tree->depth;
            L.poly_leaves.resize(L.n_poly * 32);
            L.poly_nodes.resize(L.n_poly * M.poly_tree->depth * 32);
            reqs.push_back(OpenReq{M.col_tree, yi.data(), FRI_QUERIES, L.column_leaves.data(), L.column_nodes.data()});
            reqs.push_back(OpenReq{M.poly_tree, pp.data(), L.n_poly, L.poly_leaves.data(), L.poly_nodes.data()});
        }
        if (rc == SB_OK && !reqs.empty()) rc = merkle_open_many(ctx, reqs.data(), (int)reqs.size());
        if (rc == SB_OK && last.is_last) proof->layers.push_back(std::move(last));
    }
    cleanup();
    if (rc != SB_OK) {
        delete proof;
        return rc;
    }
    *out = proof;
    return SB_OK;
}

// one fold (fri.rs:135-164) on device-resident values: the sharded prover runs layer 0 against a values tree whose
// subtrees live on several GPUs and only needs the column from this GPU
extern "C" int sb_fri_fold_dev(sb_ctx *ctx, const uint64_t *d_vals, size_t n, const uint64_t root[4], const uint8_t values_root[32],
                               uint64_t *d_col) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !d_vals || !root || !values_root || !d_col) return SB_ERR_ARG;
    if (!is_pow2(n) || n < 4) return fail(ctx, SB_ERR_ARG, "FRI fold needs a power-of-two number of values >= 4, got %zu", n);
    const uint4 *tw;
    uint32_t tw_log_n, log_stride;
    TRY(get_table(ctx, hfp::from_limbs(root), ilog2(n), &tw, &tw_log_n, &log_stride));
    hfp::el special_x = hfp::from_bytes_le32(values_root);      // fri.rs:135
    FriFoldParams P;
    memset(&P, 0, sizeof P);
    P.vals = (const uint4 *)d_vals;
    P.col = (uint4 *)d_col;
    P.tw = tw;
    P.n = n;
    P.tw_log_n = tw_log_n;
    P.tw_log_stride = log_stride;
    memcpy(P.special_x, special_x.l, 32);
    KLAUNCH(SB_KIND_FRI_FOLD, fri_launch_fold(ctx->stream, P));
    CU(cudaGetLastError());
    return SB_OK;
    });
}

extern "C" int sb_fri_prove_dev(sb_ctx *ctx, const uint64_t *d_vals, size_t n, const uint64_t root[4], size_t max_deg_plus_1,
                                uint32_t excl, const sb_tree *values_tree, sb_fri_proof **out) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !d_vals || !root || !out) return SB_ERR_ARG;
    if (values_tree && (values_tree->n != n || values_tree->leaf_bytes != 32)) return fail(ctx, SB_ERR_ARG, "values_tree does not match the values");
    return fri_prove_dev(ctx, (const uint4 *)d_vals, n, hfp::from_limbs(root), max_deg_plus_1, excl, values_tree, out);
    });
}

extern "C" int sb_fri_prove(sb_ctx *ctx, const uint64_t *vals, size_t n, const uint64_t root[4], size_t max_deg_plus_1, uint32_t excl,
                            sb_fri_proof **out) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !vals || !root || !out) return SB_ERR_ARG;
    void *d = nullptr;
    CU(cudaMallocAsync(&d, n * 32 ? n * 32 : 16, ctx->stream));
    cudaError_t e = cudaMemcpyAsync(d, vals, n * 32, cudaMemcpyHostToDevice, ctx->stream);
    int rc = e == cudaSuccess ? fri_prove_dev(ctx, (const uint4 *)d, n, hfp::from_limbs(root), max_deg_plus_1, excl, nullptr, out)
                              : fail(ctx, SB_ERR_CUDA, "H2D values: %s", cudaGetErrorString(e));
    cudaFreeAsync(d, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return rc;
    });
}

extern "C" size_t sb_fri_n_layers(const sb_fri_proof *p) { return p ? p->layers.size() : 0; }
extern "C" int sb_fri_layer_is_last(const sb_fri_proof *p, size_t i) { return (p && i < p->layers.size()) ? p->layers[i].is_last : -1; }
extern "C" int sb_fri_middle(const sb_fri_proof *p, size_t i, const uint8_t **root2, size_t *n_column, size_t *depth_column,
                             const uint8_t **column_leaves, const uint8_t **column_nodes, size_t *n_poly, size_t *depth_poly,
                             const uint8_t **poly_leaves, const uint8_t **poly_nodes) {
    if (!p || i >= p->layers.size() || p->layers[i].is_last) return SB_ERR_ARG;
    const FriLayer &L = p->layers[i];
    if (root2) *root2 = L.root2;
    if (n_column) *n_column = L.n_column;
    if (depth_column) *depth_column = L.depth_column;
    if (column_leaves) *column_leaves = L.column_leaves.data();
    if (column_nodes) *column_nodes = L.column_nodes.data();
    if (n_poly) *n_poly = L.n_poly;
    if (depth_poly) *depth_poly = L.depth_poly;
    if (poly_leaves) *poly_leaves = L.poly_leaves.data();
    if (poly_nodes) *poly_nodes = L.poly_nodes.data();
    return SB_OK;
}
extern "C" int sb_fri_last(const sb_fri_proof *p, size_t i, const uint8_t **values, size_t *n_last) {
    if (!p || i >= p->layers.size() || !p->layers[i].is_last) return SB_ERR_ARG;
    if (values) *values = p->layers[i].last.data();
    if (n_last) *n_last = p->layers[i].last.size() / 32;
    return SB_OK;
}
extern "C" int sb_fri_layer_root(const sb_fri_proof *p, size_t i, uint8_t root[32]) {
    if (!p || i >= p->layers.size() || !root || p->layers[i].is_last) return SB_ERR_ARG;
    memcpy(root, p->layers[i].values_root, 32);
    return SB_OK;
}

// serde_json prints a Vec<u8> / BlakeDigest(Vec<u8>) as an array of decimal integers; the proof is megabytes of these.
// One 8-byte table entry per byte value: up to three digits and the comma, and the length in the last byte.
namespace {
struct ByteTab {
    uint64_t e[256];
    ByteTab() {
        for (int v = 0; v < 256; v++) {
            char t[8] = {0};
            const int len = snprintf(t, sizeof t, "%d,", v);
            t[7] = (char)len;
            memcpy(&e[v], t, 8);
        }
    }
};
const ByteTab &byte_tab() {
    static const ByteTab T;        // C++11 magic static: initialised once, thread-safe
    return T;
}
}  // namespace
// writes "[b0,b1,...]" at w (at most 4 n + 2 characters, the caller provides 4 more of slack); returns the end
char *json_bytes_raw(char *w, const uint8_t *b, size_t n) {
    const uint64_t *tab = byte_tab().e;
    *w++ = '[';
    for (size_t i = 0; i < n; i++) {
        const uint64_t t = tab[b[i]];
        memcpy(w, &t, 4);          // (little endian: the low four bytes are the characters)
        w += t >> 56;
    }
    if (n) w--;                    // the last element's comma
    *w++ = ']';
    return w;
}
void json_bytes(std::string &s, const uint8_t *b, size_t n) {
    const size_t at = s.size();
    s.resize(at + 4 * n + 6);
    char *w = json_bytes_raw(&s[at], b, n);
    s.resize((size_t)(w - &s[0]));
}
// branches q0 .. q1 - 1 of a Proof{leaf,nodes} array (merkle_tree.rs:14-18), comma separated, without the enclosing brackets
char *json_branches_raw(char *w, const uint8_t *leaves, size_t leaf_bytes, const uint8_t *nodes, size_t depth, size_t q0, size_t q1) {
    for (size_t q = q0; q < q1; q++) {
        if (q != q0) *w++ = ',';
        memcpy(w, "{\"leaf\":", 8);
        w = json_bytes_raw(w + 8, leaves + q * leaf_bytes, leaf_bytes);
        memcpy(w, ",\"nodes\":[", 10);
        w += 10;
        for (size_t l = 0; l < depth; l++) {
            if (l) *w++ = ',';
            w = json_bytes_raw(w, nodes + (q * depth + l) * 32, 32);
        }
        *w++ = ']';
        *w++ = '}';
    }
    return w;
}
size_t json_branches_bound(size_t leaf_bytes, size_t depth, size_t count) {
    return count * (8 + 4 * leaf_bytes + 2 + 10 + depth * (4 * 32 + 3) + 3) + 8;
}
// Proof{leaf,nodes} (merkle_tree.rs:14-18)
void json_branches(std::string &s, const uint8_t *leaves, size_t leaf_bytes, const uint8_t *nodes, size_t depth, size_t count) {
    s.push_back('[');
    for (size_t q = 0; q < count; q++) {
        if (q) s.push_back(',');
        s += "{\"leaf\":";
        json_bytes(s, leaves + q * leaf_bytes, leaf_bytes);
        s += ",\"nodes\":[";
        for (size_t l = 0; l < depth; l++) {
            if (l) s.push_back(',');
            json_bytes(s, nodes + (q * depth + l) * 32, 32);
        }
        s += "]}";
    }
    s.push_back(']');
}
void fri_layer_json_into(std::string &s, const FriLayer &L) {
    if (L.is_last) {
        s += "{\"Last\":{\"last\":[";
        for (size_t k = 0; k < L.last.size() / 32; k++) {
            if (k) s.push_back(',');
            json_bytes(s, L.last.data() + 32 * k, 32);
        }
        s += "]}}";
    } else {
        s += "{\"Middle\":{\"root2\":";
        json_bytes(s, L.root2, 32);
        s += ",\"column_branches\":";
        json_branches(s, L.column_leaves.data(), 32, L.column_nodes.data(), L.depth_column, L.n_column);
        s += ",\"poly_branches\":";
        json_branches(s, L.poly_leaves.data(), 32, L.poly_nodes.data(), L.depth_poly, L.n_poly);
        s += "}}";
    }
}
void fri_proof_json_into(std::string &s, const sb_fri_proof *p) {
    s.push_back('[');
    for (size_t i = 0; i < p->layers.size(); i++) {
        if (i) s.push_back(',');
        fri_layer_json_into(s, p->layers[i]);
    }
    s.push_back(']');
}
extern "C" char *sb_fri_proof_json(const sb_fri_proof *p) {
    if (!p) return nullptr;
    std::string s;
    fri_proof_json_into(s, p);
    char *r = (char *)malloc(s.size() + 1);
    if (!r) return nullptr;
    memcpy(r, s.c_str(), s.size() + 1);
    return r;
}
extern "C" void sb_free_string(char *s) { free(s); }
extern "C" void sb_fri_proof_free(sb_fri_proof *p) { delete p; }

// element-wise field op on raw 256-bit limbs (no range checks): unit-test hook for the device field library
extern "C" int sb_fp_vec_op(sb_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !a || !b || !out) return SB_ERR_ARG;
    DevBuf da(ctx), db(ctx), dout(ctx);
    TRY(da.alloc(n * 32));
    TRY(db.alloc(n * 32));
    TRY(dout.alloc(n * 32));
    CU(cudaMemcpyAsync(da.p, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(db.p, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    KLAUNCH(SB_KIND_OTHER, fp_launch_vec_op(ctx->stream, op, (const uint4 *)da.p, (const uint4 *)db.p, (uint4 *)dout.p, n));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, dout.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
    });
}

// Measured issue-rate ceilings for bench.py's integer roofline: a register-only kernel of independent chains, timed with
// CUDA events on the context's stream (best of 5).  which = 0: Montgomery products per second, 1: IMAD.WIDE.U32 per second.
extern "C" int sb_pipe_peak(sb_ctx *ctx, int which, double *ops_per_s) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !ops_per_s || which < 0 || which > 1) return SB_ERR_ARG;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    const unsigned blocks = (unsigned)prop.multiProcessorCount * 2;        // 16 warps per SM, like ntt_pass_kernel
    DevBuf out(ctx);
    TRY(out.alloc((size_t)blocks * 256 * 32));
    hfp::el g = hfp::from_u64(7);
    fp seed;
    memcpy(seed.l, g.l, 32);
    const uint32_t iters = which == 0 ? 512 : 4096;
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(a, ctx->stream);
        const double ops = pipe_probe_launch(ctx->stream, which, (uint4 *)out.p, blocks, iters, seed);
        cudaEventRecord(b, ctx->stream);
        ctx->launches++;
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms > 0 && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    CU(cudaGetLastError());
    *ops_per_s = best;
    return SB_OK;
    });
}
