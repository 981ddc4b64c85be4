// prover.cu -- device-resident mk_r1cs_proof (r1cs-stark/src/prove.rs:14-378): the caller of the hot path,
// so that only the traces go in and only roots / openings / the FRI proof come out.
// Stage order and every formula follow the reference; citations are to r1cs-stark/src/prove.rs unless noted.
#include <atomic>
#include <memory>
#include <thread>

#include "internal.h"
#include "pointwise.cuh"

#include <algorithm>
#include <thread>

static const uint32_t EXTENSION_FACTOR = 8, LOG_EXTENSION_FACTOR = 3;    // utils.rs:134-135
static const size_t SPOT_CHECK_SECURITY_FACTOR = 80;                     // utils.rs:136

struct sb_stark_proof {
    uint8_t m_root[32], l_root[32], a_root[32];
    size_t depth = 0;                                       // log2(precision)
    std::vector<uint8_t> main_leaves, main_nodes;           // 320 openings of the m_tree (256-byte leaves)
    std::vector<uint8_t> lc_leaves, lc_nodes;               // 80 openings of the l_tree (32-byte leaves)
    sb_fri_proof *fri = nullptr;
    double stage_ms[5] = {0, 0, 0, 0, 0};                   // lde, merkle, fri, pointwise, total (CUDA events / host clock)
    ~sb_stark_proof() { delete fri; }
};

static unsigned nblk(size_t n, unsigned t = 128) { return (unsigned)((n + t - 1) / t); }

// utils.rs:14-23: named log2_ceil, really floor(log2 v) + 1
static uint32_t log2_ceil_quirk(size_t v) {
    uint32_t l = 1;
    while (v > 1) {
        v /= 2;
        l++;
    }
    return l;
}

// inclusive prefix product of n elements, in place allowed (out may equal in)
static int prefix_product(sb_ctx *ctx, const uint4 *in, uint4 *out, size_t n) {
    size_t len = 64;
    while ((n + len - 1) / len > 1024 * 64) len *= 2;
    const size_t chunks = (n + len - 1) / len;
    DevBuf part(ctx);
    TRY(part.alloc(chunks * 32));
    prof_begin(ctx, SB_KIND_OTHER);
    scan_chunk_reduce_kernel<<<nblk(chunks), 128, 0, ctx->stream>>>(in, (uint4 *)part.p, n, len);
    scan_partials_kernel<<<1, 1024, 0, ctx->stream>>>((uint4 *)part.p, chunks);
    scan_chunk_apply_kernel<<<nblk(chunks), 128, 0, ctx->stream>>>(in, (const uint4 *)part.p, out, n, len);
    prof_end(ctx);
    ctx->launches += 3;
    CU(cudaGetLastError());
    return SB_OK;
}

// host: coefficients of the interpolant through (xs[i], ys[i]) (poly_utils.rs:409-439) and of
// prod (X - xs[i]) (poly_utils.rs:362-373).  O(n^2) scalar work, n = number of public wires in use (1062 for the
// reference's `bits` circuit): the n synthetic divisions run on host threads, the denominators share one inversion.
static void host_zpoly(std::vector<hfp::el> &root, const std::vector<hfp::el> &xs) {
    root.assign(xs.size() + 1, hfp::ZERO);
    root[0] = hfp::ONE;
    size_t deg = 0;
    for (const auto &x : xs) {                                      // multiply by (X - x), in place from the top down
        root[deg + 1] = root[deg];
        for (size_t j = deg; j > 0; j--) root[j] = hfp::add(root[j - 1], hfp::neg(hfp::mul(root[j], x)));
        root[0] = hfp::neg(hfp::mul(root[0], x));
        deg++;
    }
}
static void host_lagrange(std::vector<hfp::el> &out, const std::vector<hfp::el> &xs, const std::vector<hfp::el> &ys,
                          std::vector<hfp::el> *root_out = nullptr) {
    const size_t n = xs.size();
    out.assign(n, hfp::ZERO);
    std::vector<hfp::el> root_local;
    std::vector<hfp::el> &root = root_out ? *root_out : root_local;
    host_zpoly(root, xs);                                          // degree n, root[n] = 1 (the vanishing polynomial, also wanted by the caller)
    if (!n) return;
    unsigned hw = std::thread::hardware_concurrency();
    const size_t T = std::max<size_t>(1, std::min<size_t>(std::min<unsigned>(hw ? hw : 1, 16), n / 64));
    // pass 1: denominators denom_i = (root / (X - x_i))(x_i) = root'(x_i)
    std::vector<hfp::el> denom(n);
    auto run = [&](auto body) {
        if (T == 1) { body((size_t)0, n); return; }
        std::vector<std::thread> th;
        for (size_t t = 0; t < T; t++) th.emplace_back([=]() { body(n * t / T, n * (t + 1) / T); });
        for (auto &x : th) x.join();
    };
    run([&](size_t lo, size_t hi) {
        std::vector<hfp::el> num(n);
        for (size_t i = lo; i < hi; i++) {
            num[n - 1] = root[n];
            for (size_t j = n - 1; j-- > 0;) num[j] = hfp::add(root[j + 1], hfp::mul(num[j + 1], xs[i]));
            hfp::el d = hfp::ZERO;
            for (size_t j = n; j-- > 0;) d = hfp::add(hfp::mul(d, xs[i]), num[j]);
            denom[i] = d;
        }
    });
    // one inversion for all denominators (distinct points: none is zero)
    std::vector<hfp::el> pre(n), scale(n);
    hfp::el acc = hfp::ONE;
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        acc = hfp::mul(acc, denom[i]);
    }
    hfp::el inv = hfp::inv(acc);
    for (size_t i = n; i-- > 0;) {
        scale[i] = hfp::mul(ys[i], hfp::mul(inv, pre[i]));          // y_i / denom_i
        inv = hfp::mul(inv, denom[i]);
    }
    // pass 2: out = sum_i scale_i * root / (X - x_i); every thread sums its share, the shares are added at the end
    std::vector<std::vector<hfp::el>> part(T, std::vector<hfp::el>(n, hfp::ZERO));
    std::vector<size_t> owner_lo(T);
    for (size_t t = 0; t < T; t++) owner_lo[t] = n * t / T;
    run([&](size_t lo, size_t hi) {
        size_t t = 0;
        while (t + 1 < T && owner_lo[t + 1] <= lo) t++;
        std::vector<hfp::el> num(n);
        std::vector<hfp::el> &o = part[t];
        for (size_t i = lo; i < hi; i++) {
            num[n - 1] = root[n];
            for (size_t j = n - 1; j-- > 0;) num[j] = hfp::add(root[j + 1], hfp::mul(num[j + 1], xs[i]));
            for (size_t j = 0; j < n; j++) o[j] = hfp::add(o[j], hfp::mul(num[j], scale[i]));
        }
    });
    for (size_t t = 0; t < T; t++)
        for (size_t j = 0; j < n; j++) out[j] = hfp::add(out[j], part[t][j]);
}

static void put_const(uint32_t (&dst)[8], const hfp::el &v) { memcpy(dst, v.l, 32); }

// lagrange_interp + zpoly (poly_utils.rs:409-439, :362-373) on the device: d_x = the n interpolation points (device),
// ys on the host; returns the n coefficients of the interpolant and the n + 1 of the vanishing polynomial on the host.
// Used from 32 points on (below that the scalar host code is faster than the launches); n is bounded by the shared memory
// the coefficient buffers of interp_zpoly_kernel need.
static const size_t INTERP_DEVICE_MIN = 32, INTERP_DEVICE_MAX = 3000;
static int device_lagrange(sb_ctx *ctx, const uint4 *d_x, const std::vector<hfp::el> &ys, std::vector<hfp::el> &interp, std::vector<hfp::el> &zroot) {
    const uint32_t n = (uint32_t)ys.size();
    DevBuf z(ctx), s(ctx), y(ctx), tab(ctx), P(ctx), out(ctx);
    TRY(z.alloc((size_t)(n + 1) * 32));
    TRY(s.alloc((size_t)n * 32));
    TRY(y.alloc((size_t)n * 32));
    TRY(tab.alloc((size_t)n * n * 32));
    TRY(P.alloc((size_t)n * 32));
    TRY(out.alloc((size_t)n * 32));
    CU(cudaMemcpyAsync(y.p, ys.data(), (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream));
    const size_t smem = (size_t)4 * (n + 1) * sizeof(uint4);
    static bool attr_set = false;            // benign race: the attribute is idempotent
    if (!attr_set) {
        CU(cudaFuncSetAttribute(interp_zpoly_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)4 * (INTERP_DEVICE_MAX + 1) * sizeof(uint4))));
        attr_set = true;
    }
    prof_begin(ctx, SB_KIND_OTHER);
    interp_zpoly_kernel<<<1, 1024, smem, ctx->stream>>>(d_x, n, (uint4 *)z.p);
    interp_denoms_kernel<<<nblk(n), 128, 0, ctx->stream>>>(d_x, n, (uint4 *)s.p);
    prof_end(ctx);
    ctx->launches += 2;
    TRY(sb_batch_inverse_dev(ctx, (uint64_t *)s.p, n));
    prof_begin(ctx, SB_KIND_OTHER);
    pw_mul_kernel<<<nblk(n), 128, 0, ctx->stream>>>((const uint4 *)s.p, (const uint4 *)y.p, (uint4 *)s.p, n);
    interp_powers_kernel<<<nblk(n), 128, 0, ctx->stream>>>(d_x, (const uint4 *)s.p, n, (uint4 *)tab.p);
    interp_powersum_kernel<<<n, 128, 0, ctx->stream>>>((const uint4 *)tab.p, n, (uint4 *)P.p);
    interp_conv_kernel<<<nblk(n), 128, 0, ctx->stream>>>((const uint4 *)z.p, (const uint4 *)P.p, n, (uint4 *)out.p);
    prof_end(ctx);
    ctx->launches += 4;
    CU(cudaGetLastError());
    interp.resize(n);
    zroot.resize(n + 1);
    CU(cudaMemcpyAsync(interp.data(), out.p, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(zroot.data(), z.p, (size_t)(n + 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

// mk_r1cs_proof on the devices of `ctx` (one, or the several of sb_init_multi).  Every N-point column is coset-major and
// sharded by cosets (sb_ext): pointwise kernels, leaf hashing and the first FRI fold run where the data is, the trees are
// built as per-device subtrees with the top finished on the host, and only S-point coefficient vectors and 32-byte digests
// cross NVLink.  The S-point accumulator chain (prefix products) runs on the device that holds the witness column.
extern "C" int sb_prove_r1cs(sb_ctx *ctx, const sb_trace *t, sb_stark_proof **out) {
    return prove_r1cs_impl(ctx, t, nullptr, out);
}
// flags != NULL: t->flag0 .. flag2 are not read, the flag columns are generated on the device from the constraints' last rows
int prove_r1cs_impl(sb_ctx *ctx, const sb_trace *t, const FlagSpec *flags, sb_stark_proof **out) {
    return guarded(ctx, "sb_prove_r1cs", [&]() -> int {
    if (!ctx || !t || !out) return SB_ERR_ARG;
    const size_t os = t->original_steps;
    if (os == 0 || os % 3 != 0) return fail(ctx, SB_ERR_ARG, "original_steps %zu must be a positive multiple of 3", os);   // :33
    if (!t->witness_trace || !t->computational_trace || !t->coefficients || !t->permuted_indices || (!flags && (!t->flag0 || !t->flag1 || !t->flag2)))
        return fail(ctx, SB_ERR_ARG, "missing trace array");
    if (flags) {
        if (flags->a * 3 != os || (flags->n_last && !flags->last_rows)) return fail(ctx, SB_ERR_ARG, "flag description does not match the trace");
        for (size_t i = 0; i < flags->n_last; i++)
            if (flags->last_rows[i] >= flags->a) return fail(ctx, SB_ERR_ARG, "constraint row out of range");
    }
    // :37-53 sizes
    const uint32_t log_steps = log2_ceil_quirk(os - 1);
    const size_t S = (size_t)1 << log_steps;
    if (S < 8) return fail(ctx, SB_ERR_ARG, "traces shorter than 8 steps hit the reference's unpatched log_steps (prove.rs:39-41)");
    const uint32_t log_prec = log_steps + LOG_EXTENSION_FACTOR;
    if (log_prec > 28) return fail(ctx, SB_ERR_ARG, "precision 2^%u exceeds the field's two-adicity (prove.rs:51-53)", log_prec);
    const size_t N = S * EXTENSION_FACTOR, sk = EXTENSION_FACTOR, o3 = os / 3;
    if (N >= ((size_t)1 << 24)) return fail(ctx, SB_ERR_ARG, "precision 2^%u: the sampler asserts modulus < 2^24 (fri/src/utils.rs:88)", log_prec);
    if (os > S) return fail(ctx, SB_ERR_ARG, "internal: steps < original_steps");
    for (size_t i = 0; i < os; i++)
        if (t->permuted_indices[i] >= S) return fail(ctx, SB_ERR_ARG, "permuted index out of range");
    const size_t np = t->n_pfi;
    if (np + 1 > S) return fail(ctx, SB_ERR_ARG, "more public wires than steps");
    for (size_t i = 0; i < np; i++)
        if (t->pfi_w[i] >= S || t->pfi_k[i] >= t->n_public) return fail(ctx, SB_ERR_ARG, "public_first_indices out of range");

    // columns: nine low-degree extensions (prove.rs:100-124, :160-167, :183-184), six quotient / combination columns, I2
    enum { K_ = 0, F0_, F1_, F2_, S_, P_, IDX_, PIDX_, A_, D1_, D2_, D3_, B2_, B3_, L_, I2_, N_COLS };
    sb_ext *E = nullptr;
    dbg_check("sb_prove_r1cs entry");
    TRY(ext_create(ctx, N_COLS, A_ + 1, log_steps, &E));
    dbg_check("after ext_create");
    const int g = E->g;
    const uint32_t cpd = E->cpd;
    const int dS = E->owner[S_];               // the device with the witness column runs the accumulator chain
    E->owner[A_] = dS;
    sb_stark_proof *proof = new sb_stark_proof();
    proof->depth = log_prec;
    std::vector<sb_tree *> trees;
    struct PerDev {
        void *perm = nullptr, *err = nullptr, *coef = nullptr, *amini = nullptr, *aleaves = nullptr;
    } pd[SB_MAX_DEV];
    struct Cleanup {
        sb_ctx *root;
        sb_ext *&E;
        std::vector<sb_tree *> &trees;
        PerDev *pd;
        sb_stark_proof *&proof;
        bool ok = false;
        ~Cleanup() {
            for (auto x : trees) free_tree(x);
            for (int d = 0; d < root->n_dev(); d++) {
                sb_ctx *c = root->dev[d];
                DevGuard dg(c);
                for (void *p : {pd[d].perm, pd[d].err, pd[d].coef, pd[d].amini, pd[d].aleaves})
                    if (p) cudaFreeAsync(p, c->stream);
            }
            ext_free(E);
            if (!ok) delete proof;
        }
    } cleanup{ctx, E, trees, pd, proof};
#define DCU(call)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? SB_ERR_OOM : SB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
    auto now = []() {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    };
    // stage times: host clock at points where every device has been waited for
    double tm[8];
    int n_tm = 0;
    auto mark = [&]() -> int {
        TRY(sync_all(ctx));
        tm[n_tm++] = now();
        return SB_OK;
    };
    auto sub_err = [&](sb_ctx *c, int rc) {           // an error recorded on a sub-context is reported on the primary
        if (c != ctx) fail(ctx, rc, "%s", c->err);
        return rc;
    };
    const hfp::el g2 = E->g2;                  // :71-94: g2 = 7^((p-1)/N); its power table doubles as `xs`
    const uint4 *xs[SB_MAX_DEV];
    for (int d = 0; d < g; d++) {
        sb_ctx *c = ctx->dev[d];
        DevGuard dg(c);
        uint32_t tw_log_n, tw_stride;
        int rc = get_table(c, g2, log_prec, &xs[d], &tw_log_n, &tw_stride, true);
        if (rc != SB_OK) return sub_err(c, rc);
        if (tw_stride != 0) return fail(ctx, SB_ERR_ARG, "internal: power table of g2 must have stride 1");
        DCU(cudaMallocAsync(&pd[d].err, sizeof(int), c->stream));
        DCU(cudaMemsetAsync(pd[d].err, 0, sizeof(int), c->stream));
    }
    TRY(mark());

    NvtxStages nvtx;
    nvtx.next("sb_prove_r1cs: inputs + 8 LDEs");
    // ---- inputs: every column goes to the device that runs its inverse transform (:55-68, :105-113, :160-163), on that
    // device's copy stream so that the uploads overlap the transforms of the columns that have already arrived ----------
    static_assert(sizeof(size_t) == sizeof(unsigned long long), "permuted_indices are uploaded as 64-bit words");
    cudaEvent_t up[6] = {0}, perm_up[SB_MAX_DEV] = {0};
    struct EvGuard {
        cudaEvent_t *a, *b;
        ~EvGuard() {
            for (int i = 0; i < 6; i++)
                if (a[i]) cudaEventDestroy(a[i]);
            for (int i = 0; i < SB_MAX_DEV; i++)
                if (b[i]) cudaEventDestroy(b[i]);
        }
    } ev_guard{up, perm_up};
    for (int d = 0; d < g; d++) {           // the copy streams start behind ext_create's allocations / memsets
        sb_ctx *c = ctx->dev[d];
        DevGuard dg(c);
        if (!c->h2d_stream) DCU(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
        cudaEvent_t ready;
        DCU(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        DCU(cudaEventRecord(ready, c->stream));
        DCU(cudaStreamWaitEvent(c->h2d_stream, ready, 0));
        cudaEventDestroy(ready);
    }
    if (flags) {
        // run.rs:283-308 on the devices that extend the flag columns: f0 = 1, f1 = 1 except after a constraint's last row, f2 = 1 only there.
        // Issued BEFORE the column uploads: the copy engine serves its queue in order, and the small upload of the row list on the
        // main stream would otherwise wait behind all of them (measured: +3.5 ms on the 2^23 proof).
        fp one;
        memcpy(one.l, hfp::ONE.l, 32);
        unsigned long long *h_rows = (unsigned long long *)pinned_scratch(ctx, flags->n_last * 8 + 8);
        if (!h_rows) return fail(ctx, SB_ERR_OOM, "pinned scratch");
        memcpy(h_rows, flags->last_rows, flags->n_last * 8);
        for (int col : {F0_, F1_, F2_}) {
            sb_ctx *c = ctx->dev[E->owner[col]];
            DevGuard dg(c);
            if (col != F2_) {
                pw_fill_kernel<<<nblk(os), 128, 0, c->stream>>>(E->input(col), os, one);
                c->launches++;
            }
            if (col != F0_ && flags->n_last) {
                DevBuf rows(c);
                TRY(rows.alloc(flags->n_last * 8));
                DCU(cudaMemcpyAsync(rows.p, h_rows, flags->n_last * 8, cudaMemcpyHostToDevice, c->stream));
                pw_flags_scatter_kernel<<<nblk(flags->n_last), 128, 0, c->stream>>>((const unsigned long long *)rows.p, flags->n_last, flags->a,
                                                                                   col == F1_ ? E->input(F1_) : nullptr, col == F2_ ? E->input(F2_) : nullptr, one);
                c->launches++;
            }
        }
    }
    for (int d : {E->owner[PIDX_], dS}) {   // the copy permutation: its padding (:55-56) is written on the device
        if (pd[d].perm) continue;
        sb_ctx *c = ctx->dev[d];
        DevGuard dg(c);
        DCU(cudaMallocAsync(&pd[d].perm, S * 8, c->stream));
        DCU(cudaEventCreateWithFlags(&perm_up[d], cudaEventDisableTiming));
        DCU(cudaEventRecord(perm_up[d], c->stream));
        DCU(cudaStreamWaitEvent(c->h2d_stream, perm_up[d], 0));
        DCU(cudaMemcpyAsync(pd[d].perm, t->permuted_indices, os * 8, cudaMemcpyHostToDevice, c->h2d_stream));
        DCU(cudaEventRecord(perm_up[d], c->h2d_stream));
        DCU(cudaStreamWaitEvent(c->stream, perm_up[d], 0));
        if (S > os) pw_perm_pad_kernel<<<nblk(S - os), 128, 0, c->stream>>>((unsigned long long *)pd[d].perm, os, S);
        c->launches++;
        DCU(cudaEventRecord(perm_up[d], c->stream));          // from here on: "permutation complete on device d"
    }
    const uint64_t *srcs[6] = {t->coefficients, flags ? nullptr : t->flag0, flags ? nullptr : t->flag1, flags ? nullptr : t->flag2, t->witness_trace,
                               t->computational_trace};
    for (int c = 0; c < 6; c++) {
        if (!srcs[c]) continue;
        sb_ctx *o = ctx->dev[E->owner[c]];
        DevGuard dg(o);
        DCU(cudaMemcpyAsync(E->input(c), srcs[c], os * 32, cudaMemcpyHostToDevice, o->h2d_stream));      // the tail stays zero (ext_create)
        DCU(cudaEventCreateWithFlags(&up[c], cudaEventDisableTiming));
        DCU(cudaEventRecord(up[c], o->h2d_stream));
    }
    {
        sb_ctx *c = ctx->dev[E->owner[IDX_]];
        DevGuard dg(c);
        pw_u64_to_fp_kernel<<<nblk(S), 128, 0, c->stream>>>(nullptr, E->input(IDX_), S);
        c->launches++;
    }
    {
        sb_ctx *c = ctx->dev[E->owner[PIDX_]];
        DevGuard dg(c);
        pw_u64_to_fp_kernel<<<nblk(S), 128, 0, c->stream>>>((const unsigned long long *)pd[E->owner[PIDX_]].perm, E->input(PIDX_), S);
        c->launches++;
    }
    // :100-124, :160-167 the eight LDEs: the two index columns first (nothing to wait for), then pairs of columns as their
    // uploads land (every owner waits for its own columns only)
    auto wait_cols = [&](int c0, int c1) -> int {
        for (int c = c0; c < c1; c++) {
            if (!up[c]) continue;                     // generated on the device
            sb_ctx *o = ctx->dev[E->owner[c]];
            DevGuard dg(o);
            DCU(cudaStreamWaitEvent(o->stream, up[c], 0));
        }
        return SB_OK;
    };
    if (S >= ((size_t)1 << 16)) {
        TRY(ext_extend(E, IDX_, 2));
        for (int c = 0; c < 6; c += 2) {
            TRY(wait_cols(c, c + 2));
            TRY(ext_extend(E, c, 2));
        }
    } else {
        TRY(wait_cols(0, 6));
        TRY(ext_extend(E, 0, 8));
    }
    // ---- constants of the pointwise stage ---------------------------------------------------------------
    PwConsts Cst;
    memset(&Cst, 0, sizeof Cst);
    {
        // g2^S: primitive 8th root of unity (:287-290); xs[N - sk] = g2^-8 (utils.rs:459) -- host scalars, no round trip
        const hfp::el gs = hfp::pow_u64(g2, S), x_last = hfp::inv(hfp::pow_u64(g2, sk));
        hfp::el pw = hfp::ONE;
        for (int i = 0; i < 8; i++) {
            put_const(Cst.pw8[i], pw);
            hfp::el z = hfp::add(pw, hfp::neg(hfp::ONE));                   // z[j] = g2^(jS) - 1 (utils.rs:173-178)
            put_const(Cst.inv_z8[i], hfp::is_zero(z) ? hfp::ZERO : hfp::inv(z));   // multi_inv: 0 -> 0
            pw = hfp::mul(pw, gs);
        }
        put_const(Cst.one, hfp::ONE);
        put_const(Cst.x_last, x_last);
    }
    auto view = [&](int d) {
        PwView v;
        v.log_s = log_steps;
        v.r0 = (uint32_t)d * cpd;
        v.n = (unsigned long long)cpd << log_steps;
        return v;
    };
    const unsigned pw_blocks = nblk((size_t)cpd << log_steps);

    // :133-151 + :203-214 (d1, d2)
    for (int d = 0; d < g; d++) {
        sb_ctx *c = ctx->dev[d];
        DevGuard dg(c);
        PwQ12Params P;
        P.k = E->col(d, K_); P.f0 = E->col(d, F0_); P.f1 = E->col(d, F1_); P.f2 = E->col(d, F2_); P.s = E->col(d, S_); P.p = E->col(d, P_);
        P.d1 = E->col(d, D1_); P.d2 = E->col(d, D2_);
        P.v = view(d); P.o3 = o3; P.err = (int *)pd[d].err;
        pw_q12_kernel<<<pw_blocks, 128, 0, c->stream>>>(P, Cst);
        c->launches++;
    }

    nvtx.next("sb_prove_r1cs: a_tree + accumulator + LDE(A)");
    // :171-184 a_root, the challenges r and the accumulator: an S-point, latency-bound chain (tree, two prefix-product scans,
    // batch inverse) on the device that holds the witness column.  (Tried: running it on a side stream next to the eight LDEs.
    // The stream-ordered pool then serves two streams and its cross-stream reuse rules made the prover slower and erratic --
    // 60-180 ms per 2^23 proof against a steady 29 ms -- so the chain stays on the device's main stream.)
    sb_tree *a_tree = nullptr;
    {
        sb_ctx *c = ctx->dev[dS];
        DevGuard dg(c);
        // :171 a_root (utils.rs:250-270): S leaves of 40 bytes
        DCU(cudaMallocAsync(&pd[dS].aleaves, S * 40, c->stream));
        pw_a_leaves_kernel<<<nblk(S), 128, 0, c->stream>>>((const unsigned long long *)pd[dS].perm, E->input(S_), (uint32_t *)pd[dS].aleaves, S);
        c->launches++;
        uint8_t *leaves = (uint8_t *)pd[dS].aleaves;
        pd[dS].aleaves = nullptr;                       // the tree takes ownership
        int rc = commit_bytes_owned(c, leaves, 40, S, &a_tree);
        if (rc != SB_OK) return sub_err(c, rc);
        trees.push_back(a_tree);
        memcpy(proof->a_root, a_tree->root, 32);
        // :172 r = get_random_ff_values(a_root, precision, 3, 0) (utils.rs:272-290)
        uint32_t idx[24];
        if (sb_pseudorandom_indices(proof->a_root, 32, (uint32_t)N, 24, 0, idx) != SB_OK) return fail(ctx, SB_ERR_ARG, "sampler rejected precision %zu", N);
        for (int i = 0; i < 3; i++) {
            uint8_t b[32];
            for (int j = 0; j < 8; j++) {     // utils.rs:29-38: each u32 big-endian, the 32 bytes then read little-endian
                uint32_t v = idx[8 * i + j];
                b[4 * j] = (uint8_t)(v >> 24); b[4 * j + 1] = (uint8_t)(v >> 16); b[4 * j + 2] = (uint8_t)(v >> 8); b[4 * j + 3] = (uint8_t)v;
            }
            put_const(Cst.r[i], hfp::from_bytes_le32(b));
        }
        // :175-184 accumulator (utils.rs:293-339)
        DCU(cudaMallocAsync(&pd[dS].amini, 2 * S * 32, c->stream));
        uint4 *nmr = (uint4 *)pd[dS].amini, *dnm = nmr + 2 * S;
        pw_acc_terms_kernel<<<nblk(S), 128, 0, c->stream>>>((const unsigned long long *)pd[dS].perm, E->input(S_), nmr, dnm, S, Cst);
        c->launches++;
        rc = prefix_product(c, nmr, nmr, S);
        if (rc == SB_OK) rc = prefix_product(c, dnm, dnm, S);
        if (rc == SB_OK) rc = sb_batch_inverse_dev(c, (uint64_t *)dnm, S);
        if (rc != SB_OK) return sub_err(c, rc);
        pw_mul_kernel<<<nblk(S), 128, 0, c->stream>>>(nmr, dnm, E->input(A_), S);
        c->launches++;
    }
    TRY(ext_extend(E, A_, 1));
    // :192-214 d3
    for (int d = 0; d < g; d++) {
        sb_ctx *c = ctx->dev[d];
        DevGuard dg(c);
        PwQ3Params P;
        P.a = E->col(d, A_); P.s = E->col(d, S_); P.idx = E->col(d, IDX_); P.pidx = E->col(d, PIDX_);
        P.d3 = E->col(d, D3_); P.v = view(d); P.err = (int *)pd[d].err;
        pw_q3_kernel<<<pw_blocks, 128, 0, c->stream>>>(P, Cst);
        c->launches++;
    }
    // :216-232 boundary quotients.  i2 and zb2 are polynomials of degree <= n_pub evaluated on the whole domain: the
    // reference does that with eval_poly_at / a product per point (O(N n_pub)); coset transforms of the same coefficients
    // (or Horner per point for a handful of public wires) give the same field elements.
    {
        std::vector<hfp::el> xv(np), yv(np), interp, zroot;
        for (size_t i = 0; i < np; i++) yv[i] = hfp::from_limbs(t->public_wires + 4 * t->pfi_k[i]);   // utils.rs:421-435
        const bool on_device = np >= INTERP_DEVICE_MIN && np <= INTERP_DEVICE_MAX;
        if (np && on_device) {
            std::vector<unsigned long long> pos(np);
            for (size_t i = 0; i < np; i++) pos[i] = sk * t->pfi_w[i];
            DevBuf dpos(ctx), dx(ctx);
            TRY(dpos.alloc(np * 8));
            TRY(dx.alloc(np * 32));
            DCU(cudaMemcpyAsync(dpos.p, pos.data(), np * 8, cudaMemcpyHostToDevice, ctx->stream));
            ctx->launches += merkle_launch_gather_bytes(ctx->stream, (const uint8_t *)xs[0], 32, (const unsigned long long *)dpos.p, (uint32_t)np, (uint8_t *)dx.p);
            TRY(device_lagrange(ctx, (const uint4 *)dx.p, yv, interp, zroot));
        } else {
            for (size_t i = 0; i < np; i++) xv[i] = hfp::pow_u64(g2, sk * t->pfi_w[i]);     // xs[sk w]: a handful of host scalars
        }
        if (!on_device) host_lagrange(interp, xv, yv, &zroot);
        const bool horner = np + 1 <= 24;      // ~np products per point against the ~log2(S)/2 + 2 of a transform
        // the 2 np + 1 coefficients go to every device from pinned memory (a pageable source would make each upload wait for
        // its device's stream and run the devices one after the other)
        hfp::el *hcoef = (hfp::el *)pinned_scratch(ctx, (2 * np + 1) * 32);
        if (!hcoef) return fail(ctx, SB_ERR_OOM, "pinned scratch");
        if (np) memcpy(hcoef, interp.data(), np * 32);
        memcpy(hcoef + np, zroot.data(), (np + 1) * 32);
        for (int d = 0; d < g; d++) {
            sb_ctx *c = ctx->dev[d];
            DevGuard dg(c);
            DCU(cudaMallocAsync(&pd[d].coef, (2 * np + 1) * 32, c->stream));
            uint4 *ci = (uint4 *)pd[d].coef, *cz = ci + 2 * np;
            DCU(cudaMemcpyAsync(ci, hcoef, (2 * np + 1) * 32, cudaMemcpyHostToDevice, c->stream));
            uint4 *zb2 = E->col(d, B2_), *zb3 = E->col(d, B3_);          // adjacent columns: one batch inverse over both
            if (horner) {
                if (np) {
                    pw_poly_eval_kernel<<<pw_blocks, 128, 0, c->stream>>>(xs[d], ci, (uint32_t)np, E->col(d, I2_), view(d));
                    c->launches++;
                }
                pw_poly_eval_kernel<<<pw_blocks, 128, 0, c->stream>>>(xs[d], cz, (uint32_t)np + 1, zb2, view(d));
                c->launches++;
            } else {
                CosetSpec cs;
                cs.log_ext = LOG_EXTENSION_FACTOR;
                cs.store = NTT_STORE_PLAIN;
                cs.r0 = (uint32_t)d * cpd;
                cs.cnt = cpd;
                int rc = ntt_dev_tw(c, ci, np, np, E->col(d, I2_), S, 1, log_steps, 0, xs[d], log_prec, LOG_EXTENSION_FACTOR, &cs);
                if (rc == SB_OK) rc = ntt_dev_tw(c, cz, np + 1, np + 1, zb2, S, 1, log_steps, 0, xs[d], log_prec, LOG_EXTENSION_FACTOR, &cs);
                if (rc != SB_OK) return sub_err(c, rc);
            }
            pw_zb3_kernel<<<pw_blocks, 128, 0, c->stream>>>(xs[d], zb3, view(d), Cst);
            c->launches++;
            int rc = sb_batch_inverse_dev(c, (uint64_t *)zb2, 2 * ((size_t)cpd << log_steps));
            if (rc != SB_OK) return sub_err(c, rc);
            PwB23Params P;
            P.s = E->col(d, S_); P.a = E->col(d, A_); P.i2 = np ? E->col(d, I2_) : nullptr;
            P.inv_zb2 = zb2; P.inv_zb3 = zb3; P.v = view(d); P.err = (int *)pd[d].err;
            pw_b23_kernel<<<pw_blocks, 128, 0, c->stream>>>(P, Cst);
            c->launches++;
        }
    }
    // the reference's asserts (utils.rs:379-418, :489, :514) -> error code
    {
        int err[SB_MAX_DEV] = {0};
        for (int d = 0; d < g; d++) {
            sb_ctx *c = ctx->dev[d];
            DevGuard dg(c);
            DCU(cudaMemcpyAsync(&err[d], pd[d].err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            DCU(cudaGetLastError());
        }
        TRY(mark());
        int bad = 0;
        for (int d = 0; d < g; d++)
            if (err[d] && (!bad || err[d] < bad)) bad = err[d];
        if (bad)
            return fail(ctx, SB_ERR_ARG, bad == 1 ? "invalid D1/D2/D3: the witness does not satisfy the constraints (utils.rs:379-418)"
                                                  : (bad == 2 ? "invalid B2: public wires do not match the trace (utils.rs:489)" : "invalid B3 (utils.rs:514)"));
    }
    nvtx.next("sb_prove_r1cs: m_tree");
    // :235-264 m_tree over p a s d1 d2 d3 b2 b3
    sb_tree *m_tree = nullptr;
    {
        const size_t ids[8] = {P_, A_, S_, D1_, D2_, D3_, B2_, B3_};
        TRY(ext_commit(E, ids, 8, &m_tree));
        trees.push_back(m_tree);
        memcpy(proof->m_root, m_tree->root, 32);
    }
    TRY(mark());
    nvtx.next("sb_prove_r1cs: l + l_tree + openings");
    // :274-283 k[i] = int_BE(blake(m_root || i)) mod p
    put_const(Cst.k[0], hfp::ONE);
    for (int i = 1; i < 11; i++) {
        uint8_t msg[33], h[32], le[32];
        memcpy(msg, proof->m_root, 32);
        msg[32] = (uint8_t)i;
        b2s::hash_bytes(h, msg, 33);
        for (int b = 0; b < 32; b++) le[b] = h[31 - b];
        put_const(Cst.k[i], hfp::from_bytes_le32(le));
    }
    for (int r = 0; r < 8; r++) {         // X = (g2^S)^(j mod 8) in the p, b2, b3 terms (:287-322)
        hfp::el X, kk[11];
        memcpy(X.l, Cst.pw8[r], 32);
        for (int i = 0; i < 11; i++) memcpy(kk[i].l, Cst.k[i], 32);
        put_const(Cst.kx[0][r], hfp::add(kk[3], hfp::mul(kk[4], X)));
        put_const(Cst.kx[1][r], hfp::add(kk[5], hfp::mul(kk[6], X)));
        put_const(Cst.kx[2][r], hfp::add(kk[7], hfp::mul(kk[8], X)));
    }
    // :287-322 l
    for (int d = 0; d < g; d++) {
        sb_ctx *c = ctx->dev[d];
        DevGuard dg(c);
        PwLParams P;
        P.d1 = E->col(d, D1_); P.d2 = E->col(d, D2_); P.d3 = E->col(d, D3_); P.p = E->col(d, P_); P.b2 = E->col(d, B2_); P.b3 = E->col(d, B3_);
        P.a = E->col(d, A_); P.s = E->col(d, S_); P.l = E->col(d, L_); P.v = view(d);
        pw_l_kernel<<<pw_blocks, 128, 0, c->stream>>>(P, Cst);
        c->launches++;
    }
    // :324-332 l_tree
    sb_tree *l_tree = nullptr;
    {
        const size_t ids[1] = {L_};
        TRY(ext_commit(E, ids, 1, &l_tree));
        trees.push_back(l_tree);
        memcpy(proof->l_root, l_tree->root, 32);
    }
    // :337-362 spot-check positions and openings
    {
        uint32_t pos32[SPOT_CHECK_SECURITY_FACTOR];
        if (sb_pseudorandom_indices(proof->l_root, 32, (uint32_t)N, SPOT_CHECK_SECURITY_FACTOR, (uint32_t)sk, pos32) != SB_OK)
            return fail(ctx, SB_ERR_ARG, "sampler rejected precision %zu", N);
        std::vector<size_t> positions(SPOT_CHECK_SECURITY_FACTOR), aug(4 * SPOT_CHECK_SECURITY_FACTOR);
        for (size_t i = 0; i < SPOT_CHECK_SECURITY_FACTOR; i++) {
            const size_t j = positions[i] = pos32[i];
            aug[4 * i + 0] = j;
            aug[4 * i + 1] = (j + N - sk) % N;
            aug[4 * i + 2] = (j + o3 * sk) % N;
            aug[4 * i + 3] = (j + o3 * 2 * sk) % N;
        }
        proof->lc_leaves.resize(positions.size() * 32);
        proof->lc_nodes.resize(positions.size() * log_prec * 32);
        proof->main_leaves.resize(aug.size() * 256);
        proof->main_nodes.resize(aug.size() * log_prec * 32);
        const OpenReq reqs[2] = {{l_tree, positions.data(), positions.size(), proof->lc_leaves.data(), proof->lc_nodes.data()},
                                 {m_tree, aug.data(), aug.size(), proof->main_leaves.data(), proof->main_nodes.data()}};
        TRY(merkle_open_many(ctx, reqs, 2));
    }
    TRY(mark());
    nvtx.next("sb_prove_r1cs: FRI");
    // :367 FRI on l with the committed l_tree
    TRY(ext_fri_prove(E, L_, l_tree, N / 4, (uint32_t)sk, &proof->fri));
    TRY(mark());
    // marks: 0 start, 1 pointwise stage checked (inputs, nine LDEs, accumulator chain, quotients), 2 m_tree, 3 l + l_tree +
    // openings, 4 FRI (host clock; every device has been waited for at each mark)
    proof->stage_ms[0] = tm[1] - tm[0];
    proof->stage_ms[1] = tm[2] - tm[1];
    proof->stage_ms[2] = tm[4] - tm[3];
    proof->stage_ms[4] = tm[4] - tm[0];
    proof->stage_ms[3] = tm[3] - tm[2];
    cleanup.ok = true;
    *out = proof;
    return SB_OK;
#undef DCU
    });
}

extern "C" int sb_stark_proof_roots(const sb_stark_proof *p, uint8_t m_root[32], uint8_t l_root[32], uint8_t a_root[32]) {
    if (!p) return SB_ERR_ARG;
    if (m_root) memcpy(m_root, p->m_root, 32);
    if (l_root) memcpy(l_root, p->l_root, 32);
    if (a_root) memcpy(a_root, p->a_root, 32);
    return SB_OK;
}
extern "C" int sb_stark_proof_stage_ms(const sb_stark_proof *p, double ms[5]) {
    if (!p || !ms) return SB_ERR_ARG;
    memcpy(ms, p->stage_ms, sizeof p->stage_ms);
    return SB_OK;
}
// serde_json::to_string(&StarkProof) (utils.rs:122-130 field order; run.rs:549 compact)
// The proof is megabytes of decimal byte arrays.  The text is cut into pieces -- literals, and runs of at most JSON_RUN Merkle
// branches of the main / linear-combination / FRI sections -- that are formatted independently on a few host threads into
// their own buffers and then copied to their offsets in the result (round 1: one thread, 1.2 of poseidon3_test's 5 ms and
// 3.5 of the 2^23 circuit's 54 ms; first threaded version: one thread per section, bounded by the largest section).
namespace {
struct JsonPiece {
    std::string lit;                              // literal text, or
    const uint8_t *leaves = nullptr, *nodes = nullptr;   // branches q0 .. q1 - 1 of an array of Proof{leaf,nodes}
    size_t leaf_bytes = 0, depth = 0, q0 = 0, q1 = 0;
    const uint8_t *digests = nullptr;             // or n_digests 32-byte arrays, comma separated (FRI last layer)
    size_t n_digests = 0;
    std::unique_ptr<char[]> buf;
    size_t n = 0;
    void make() {
        if (leaves) {
            buf.reset(new char[json_branches_bound(leaf_bytes, depth, q1 - q0)]);
            n = (size_t)(json_branches_raw(buf.get(), leaves, leaf_bytes, nodes, depth, q0, q1) - buf.get());
        } else if (digests) {
            buf.reset(new char[n_digests * (4 * 32 + 3) + 8]);
            char *w = buf.get();
            for (size_t k = 0; k < n_digests; k++) {
                if (k) *w++ = ',';
                w = json_bytes_raw(w, digests + 32 * k, 32);
            }
            n = (size_t)(w - buf.get());
        } else {
            n = lit.size();
        }
    }
    const char *data() const { return buf ? buf.get() : lit.data(); }
};
const size_t JSON_RUN = 48;       // branches per piece (~25-60 KB of text)

void json_pieces_branches(std::vector<JsonPiece> &out, const uint8_t *leaves, size_t leaf_bytes, const uint8_t *nodes, size_t depth, size_t count) {
    JsonPiece open;
    open.lit = "[";
    out.push_back(std::move(open));
    for (size_t q0 = 0; q0 < count; q0 += JSON_RUN) {
        if (q0) {
            JsonPiece sep;
            sep.lit = ",";
            out.push_back(std::move(sep));
        }
        JsonPiece pc;
        pc.leaves = leaves;
        pc.nodes = nodes;
        pc.leaf_bytes = leaf_bytes;
        pc.depth = depth;
        pc.q0 = q0;
        pc.q1 = std::min(count, q0 + JSON_RUN);
        out.push_back(std::move(pc));
    }
    JsonPiece close;
    close.lit = "]";
    out.push_back(std::move(close));
}
void json_lit(std::vector<JsonPiece> &out, const std::string &text) {
    JsonPiece pc;
    pc.lit = text;
    out.push_back(std::move(pc));
}
}  // namespace

extern "C" char *sb_stark_proof_json(const sb_stark_proof *p, size_t *len) {
    if (!p) return nullptr;
    try {
        std::vector<JsonPiece> pcs;
        {
            std::string head = "{\"m_root\":";
            json_bytes(head, p->m_root, 32);
            head += ",\"l_root\":";
            json_bytes(head, p->l_root, 32);
            head += ",\"a_root\":";
            json_bytes(head, p->a_root, 32);
            head += ",\"main_branches\":";
            json_lit(pcs, head);
        }
        json_pieces_branches(pcs, p->main_leaves.data(), 256, p->main_nodes.data(), p->depth, p->main_leaves.size() / 256);
        json_lit(pcs, ",\"linear_comb_branches\":");
        json_pieces_branches(pcs, p->lc_leaves.data(), 32, p->lc_nodes.data(), p->depth, p->lc_leaves.size() / 32);
        json_lit(pcs, ",\"fri_proof\":[");
        const size_t n_layers = p->fri ? p->fri->layers.size() : 0;
        for (size_t i = 0; i < n_layers; i++) {           // fri.rs:16-26, same text as fri_layer_json_into
            const FriLayer &L = p->fri->layers[i];
            if (i) json_lit(pcs, ",");
            if (L.is_last) {
                json_lit(pcs, "{\"Last\":{\"last\":[");
                JsonPiece pc;
                pc.digests = L.last.data();
                pc.n_digests = L.last.size() / 32;
                if (pc.n_digests) pcs.push_back(std::move(pc));
                json_lit(pcs, "]}}");
            } else {
                std::string h = "{\"Middle\":{\"root2\":";
                json_bytes(h, L.root2, 32);
                h += ",\"column_branches\":";
                json_lit(pcs, h);
                json_pieces_branches(pcs, L.column_leaves.data(), 32, L.column_nodes.data(), L.depth_column, L.n_column);
                json_lit(pcs, ",\"poly_branches\":");
                json_pieces_branches(pcs, L.poly_leaves.data(), 32, L.poly_nodes.data(), L.depth_poly, L.n_poly);
                json_lit(pcs, "}}");
            }
        }
        json_lit(pcs, "]}");
        // format the pieces: a few threads take them in turn (the calling thread included)
        const size_t raw = p->main_leaves.size() + p->main_nodes.size() + p->lc_nodes.size();
        unsigned T = std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency()));
        if (raw < ((size_t)1 << 16)) T = 1;
        std::atomic<size_t> next(0);
        std::atomic<bool> failed(false);
        auto work = [&]() {
            try {
                for (size_t i = next.fetch_add(1); i < pcs.size(); i = next.fetch_add(1)) pcs[i].make();
            } catch (...) {
                failed.store(true);
            }
        };
        {
            std::vector<std::thread> th;
            for (unsigned t = 1; t < T; t++) {
                try {
                    th.emplace_back(work);
                } catch (...) {
                    break;          // no more threads to be had: the ones running (and this one) take all pieces
                }
            }
            work();
            for (auto &t : th) t.join();
        }
        if (failed.load()) return nullptr;
        size_t total = 0;
        std::vector<size_t> off(pcs.size());
        for (size_t i = 0; i < pcs.size(); i++) {
            off[i] = total;
            total += pcs[i].n;
        }
        char *r = (char *)malloc(total + 1);
        if (!r) return nullptr;
        next.store(0);
        auto gather = [&]() {
            for (size_t i = next.fetch_add(1); i < pcs.size(); i = next.fetch_add(1)) memcpy(r + off[i], pcs[i].data(), pcs[i].n);
        };
        {
            std::vector<std::thread> th;
            if (total >= ((size_t)1 << 21))
                for (unsigned t = 1; t < std::min(T, 4u); t++) {
                    try {
                        th.emplace_back(gather);
                    } catch (...) {
                        break;
                    }
                }
            gather();
            for (auto &t : th) t.join();
        }
        r[total] = 0;
        if (len) *len = total;
        return r;
    } catch (...) {
        return nullptr;
    }
}
extern "C" void sb_stark_proof_free(sb_stark_proof *p) { delete p; }

// =====================================================================================================================
// Verifier (r1cs-stark/src/verify.rs:13-258, fri/src/fri.rs:226-404, commitment/src/merkle_tree.rs:25-58): SURVEY.md 8f
// next-3.  The reference verifier interpolates K, F0, F1, F2 and the two index columns and evaluates them at the 80
// spot-check points with eval_poly_at (O(80 S) products) plus three N-point transforms; here the six columns are
// extended on the device by the same LDE the prover uses and read at the 80 positions (identical field elements).
// Merkle branches, the FRI layer checks and the constraint equations are O(proof size) scalar work on the host.
// Returns SB_OK when the proof is accepted and SB_ERR_VERIFY (reference: assert / unwrap panic) otherwise.
// =====================================================================================================================
namespace {

bool branch_ok(const uint8_t root[32], size_t index, const uint8_t *leaf, size_t leaf_bytes, const uint8_t *nodes, size_t depth) {
    uint8_t h[32], buf[64];
    b2s::hash_bytes(h, leaf, leaf_bytes);                         // merkle_tree.rs:26
    for (size_t l = 0; l < depth; l++) {                          // merkle_tree.rs:27-41
        const uint8_t *sib = nodes + 32 * l;
        if (index & 1) { memcpy(buf, sib, 32); memcpy(buf + 32, h, 32); } else { memcpy(buf, h, 32); memcpy(buf + 32, sib, 32); }
        b2s::hash_bytes(h, buf, 64);
        index >>= 1;
    }
    return memcmp(h, root, 32) == 0;
}

hfp::el sub(const hfp::el &a, const hfp::el &b) { return hfp::add(a, hfp::neg(b)); }

// value at x of the polynomial of degree < n through (xs[i], ys[i]) (poly_utils.rs:409-439 + eval_poly_at); the n
// denominators share one field inversion (Montgomery's trick)
hfp::el lagrange_eval(const std::vector<hfp::el> &xs, const std::vector<hfp::el> &ys, const hfp::el &x) {
    const size_t n = xs.size();
    if (n == 0) return hfp::ZERO;
    std::vector<hfp::el> num(n), den(n), pre(n);
    for (size_t i = 0; i < n; i++) {
        num[i] = hfp::ONE;
        den[i] = hfp::ONE;
        for (size_t j = 0; j < n; j++) {
            if (i == j) continue;
            num[i] = hfp::mul(num[i], sub(x, xs[j]));
            den[i] = hfp::mul(den[i], sub(xs[i], xs[j]));
        }
    }
    hfp::el run = hfp::ONE;
    for (size_t i = 0; i < n; i++) {
        pre[i] = run;
        run = hfp::mul(run, den[i]);
    }
    hfp::el inv = hfp::inv(run);          // the points are distinct, so no denominator is zero
    hfp::el acc = hfp::ZERO;
    for (size_t i = n; i-- > 0;) {
        const hfp::el di = hfp::mul(inv, pre[i]);
        inv = hfp::mul(inv, den[i]);
        acc = hfp::add(acc, hfp::mul(ys[i], hfp::mul(num[i], di)));
    }
    return acc;
}

// fri.rs:244-404
int fri_verify_host(sb_ctx *ctx, const sb_fri_proof *pr, const uint8_t values_root[32], hfp::el w, size_t n, size_t bound, uint32_t excl) {
    if (!pr || pr->layers.empty()) return fail(ctx, SB_ERR_VERIFY, "FRI proof is empty");
    uint8_t root[32];
    memcpy(root, values_root, 32);
    for (size_t li = 0; li + 1 < pr->layers.size(); li++) {
        const FriLayer &L = pr->layers[li];
        if (L.is_last) return fail(ctx, SB_ERR_VERIFY, "FRI proofs must consist of Middle layers except the last element (fri.rs:279)");
        if (n < 4) return fail(ctx, SB_ERR_VERIFY, "FRI layer %zu: domain exhausted", li);
        const size_t q = n / 4;
        const hfp::el special_x = hfp::from_bytes_le32(root);                      // fri.rs:285
        uint32_t ys[FRI_QUERIES];
        if (pseudorandom_indices(L.root2, 32, (uint32_t)q, FRI_QUERIES, excl, ys, ctx->extended_domain) != SB_OK)
            return fail(ctx, SB_ERR_VERIFY, "FRI layer %zu: sampler rejects column length %zu", li, q);
        uint32_t dq = 0, dn = 0;
        while (((size_t)1 << dq) < q) dq++;
        while (((size_t)1 << dn) < n) dn++;
        if (L.n_column != FRI_QUERIES || L.n_poly != 4 * FRI_QUERIES || L.depth_column != dq || L.depth_poly != dn ||
            L.column_leaves.size() != FRI_QUERIES * 32 || L.poly_leaves.size() != 4 * FRI_QUERIES * 32 ||
            L.column_nodes.size() != FRI_QUERIES * dq * 32 || L.poly_nodes.size() != 4 * FRI_QUERIES * dn * 32)
            return fail(ctx, SB_ERR_VERIFY, "FRI layer %zu: malformed branches", li);
        const hfp::el iota = hfp::pow_u64(w, q);                                    // quartic roots of unity, fri.rs:263-268
        for (size_t i = 0; i < FRI_QUERIES; i++) {
            if (!branch_ok(L.root2, ys[i], &L.column_leaves[32 * i], 32, &L.column_nodes[i * dq * 32], dq))
                return fail(ctx, SB_ERR_VERIFY, "FRI layer %zu: column branch %zu does not match root2 (fri.rs:310)", li, i);
            std::vector<hfp::el> xc(4), row(4);
            hfp::el x = hfp::pow_u64(w, ys[i]);
            for (size_t j = 0; j < 4; j++) {
                const size_t pos = j * q + ys[i];                                   // fri.rs:301-307
                if (!branch_ok(root, pos, &L.poly_leaves[(4 * i + j) * 32], 32, &L.poly_nodes[(4 * i + j) * dn * 32], dn))
                    return fail(ctx, SB_ERR_VERIFY, "FRI layer %zu: poly branch %zu does not match the layer root (fri.rs:311)", li, 4 * i + j);
                xc[j] = x;
                row[j] = hfp::from_bytes_le32(&L.poly_leaves[(4 * i + j) * 32]);
                x = hfp::mul(x, iota);
            }
            // fri.rs:340-346: the four row values and the column value lie on one polynomial of degree < 4
            if (!hfp::eq(lagrange_eval(xc, row, special_x), hfp::from_bytes_le32(&L.column_leaves[32 * i])))
                return fail(ctx, SB_ERR_VERIFY, "FRI layer %zu: query %zu is not on the degree-<4 interpolant (fri.rs:345)", li, i);
        }
        memcpy(root, L.root2, 32);
        w = hfp::sqr(hfp::sqr(w));
        n = q;
        bound /= 4;
    }
    const FriLayer &last = pr->layers.back();
    if (!last.is_last) return fail(ctx, SB_ERR_VERIFY, "the last element of FRI proofs must be Last (fri.rs:362)");
    if (bound < FRI_MIN_DEG_DIRECT / 2) return fail(ctx, SB_ERR_VERIFY, "the degree of direct checking is too low (fri.rs:355)");
    const size_t m = last.last.size() / 32;
    if (m != n || m <= bound) return fail(ctx, SB_ERR_VERIFY, "FRI last layer holds %zu values, expected %zu > %zu (fri.rs:366)", m, n, bound);
    // fri.rs:374-383: the Merkle root of the last values matches
    {
        std::vector<uint8_t> lv(m * 32);
        for (size_t i = 0; i < m; i++) b2s::hash_bytes(&lv[32 * i], &last.last[32 * i], 32);
        for (size_t w2 = m; w2 > 1; w2 /= 2)
            for (size_t i = 0; i < w2 / 2; i++) {
                uint8_t buf[64];
                memcpy(buf, &lv[64 * i], 64);
                b2s::hash_bytes(&lv[32 * i], buf, 64);
            }
        if (memcmp(lv.data(), root, 32) != 0) return fail(ctx, SB_ERR_VERIFY, "FRI last layer does not hash to the previous root2 (fri.rs:383)");
    }
    // fri.rs:385-401: degree check on the points that are not multiples of `excl`
    std::vector<size_t> pts;
    for (size_t pos = 0; pos < m; pos++)
        if (excl == 0 || pos % excl != 0) pts.push_back(pos);
    if (pts.size() < bound) return fail(ctx, SB_ERR_VERIFY, "FRI last layer: not enough points");
    std::vector<hfp::el> xv(bound), yv(bound);
    std::vector<hfp::el> pw(m);
    pw[0] = hfp::ONE;
    for (size_t i = 1; i < m; i++) pw[i] = hfp::mul(pw[i - 1], w);
    for (size_t i = 0; i < bound; i++) {
        xv[i] = pw[pts[i]];
        yv[i] = hfp::from_bytes_le32(&last.last[32 * pts[i]]);
    }
    for (size_t i = bound; i < pts.size(); i++)
        if (!hfp::eq(lagrange_eval(xv, yv, pw[pts[i]]), hfp::from_bytes_le32(&last.last[32 * pts[i]])))
            return fail(ctx, SB_ERR_VERIFY, "FRI last layer is not of degree < %zu (fri.rs:399)", bound);
    return SB_OK;
}

}  // namespace

extern "C" int sb_verify_r1cs(sb_ctx *ctx, const sb_trace *t, const sb_stark_proof *proof) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !t || !proof) return SB_ERR_ARG;
    const size_t os = t->original_steps;
    if (os == 0 || os % 3 != 0) return fail(ctx, SB_ERR_ARG, "original_steps %zu must be a positive multiple of 3", os);   // verify.rs:27
    if (!t->coefficients || !t->flag0 || !t->flag1 || !t->flag2 || !t->permuted_indices) return fail(ctx, SB_ERR_ARG, "missing public array");
    const uint32_t log_steps = log2_ceil_quirk(os - 1);                      // verify.rs:29-33
    const size_t S = (size_t)1 << log_steps;
    if (S < 8) return fail(ctx, SB_ERR_ARG, "traces shorter than 8 steps are not supported");
    const uint32_t log_prec = log_steps + LOG_EXTENSION_FACTOR;
    if (log_prec > 28) return fail(ctx, SB_ERR_ARG, "precision 2^%u exceeds the field's two-adicity (verify.rs:38)", log_prec);
    const size_t N = S * EXTENSION_FACTOR, sk = EXTENSION_FACTOR, o3 = os / 3;
    if (N >= ((size_t)1 << 24)) return fail(ctx, SB_ERR_ARG, "precision 2^%u: the sampler asserts modulus < 2^24", log_prec);
    if (proof->depth != log_prec) return fail(ctx, SB_ERR_VERIFY, "proof was made for precision 2^%zu, the circuit needs 2^%u", proof->depth, log_prec);
    const size_t V = SPOT_CHECK_SECURITY_FACTOR;
    if (proof->main_leaves.size() != 4 * V * 256 || proof->main_nodes.size() != 4 * V * log_prec * 32 || proof->lc_leaves.size() != V * 32 ||
        proof->lc_nodes.size() != V * log_prec * 32)
        return fail(ctx, SB_ERR_VERIFY, "malformed branches");
    int rc = SB_OK;
#define VTRY(expr) do { rc = (expr); if (rc != SB_OK) return rc; } while (0)
#define VCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, SB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

    // verify.rs:55-66 g2 and xs
    hfp::el g2;
    {
        uint64_t e[4] = {hfp::PMOD[0] - 1, hfp::PMOD[1], hfp::PMOD[2], hfp::PMOD[3]};
        for (uint32_t i = 0; i < log_prec; i++) {
            for (int k = 0; k < 3; k++) e[k] = (e[k] >> 1) | (e[k + 1] << 63);
            e[3] >>= 1;
        }
        g2 = hfp::pow_limbs(hfp::from_u64(7), e, 4);
    }
    const uint4 *xs;
    uint32_t tw_log_n, tw_stride;
    VTRY(get_table(ctx, g2, log_prec, &xs, &tw_log_n, &tw_stride, true));

    // verify.rs:83-87 low-degree proof of l
    VTRY(fri_verify_host(ctx, proof->fri, proof->l_root, g2, N, N / 4, (uint32_t)sk));

    // verify.rs:89-118 positions and branches
    uint32_t pos32[SPOT_CHECK_SECURITY_FACTOR];
    if (pseudorandom_indices(proof->l_root, 32, (uint32_t)N, V, (uint32_t)sk, pos32, false) != SB_OK) return fail(ctx, SB_ERR_ARG, "sampler rejected precision %zu", N);
    std::vector<size_t> aug(4 * V);
    for (size_t i = 0; i < V; i++) {
        const size_t j = pos32[i];
        aug[4 * i] = j;
        aug[4 * i + 1] = (j + N - sk) % N;
        aug[4 * i + 2] = (j + o3 * sk) % N;
        aug[4 * i + 3] = (j + o3 * 2 * sk) % N;
    }
    for (size_t i = 0; i < 4 * V; i++)
        if (!branch_ok(proof->m_root, aug[i], &proof->main_leaves[256 * i], 256, &proof->main_nodes[i * log_prec * 32], log_prec))
            return fail(ctx, SB_ERR_VERIFY, "main branch %zu does not match m_root (verify.rs:115)", i);
    for (size_t i = 0; i < V; i++)
        if (!branch_ok(proof->l_root, pos32[i], &proof->lc_leaves[32 * i], 32, &proof->lc_nodes[i * log_prec * 32], log_prec))
            return fail(ctx, SB_ERR_VERIFY, "linear combination branch %zu does not match l_root (verify.rs:117)", i);

    // verify.rs:72-80, 127-136: K F0 F1 F2 idx pidx extended to the N-point domain on the device, read at the positions
    DevBuf in6(ctx), ev6(ctx), perm_d(ctx), dpos(ctx), dgath(ctx);
    VTRY(in6.alloc(6 * S * 32));
    VTRY(ev6.alloc(6 * N * 32));
    VTRY(perm_d.alloc(S * 8));
    auto in_col = [&](int c) { return (uint4 *)in6.p + 2 * (size_t)c * S; };
    VCU(cudaMemsetAsync(in6.p, 0, 6 * S * 32, ctx->stream));
    const uint64_t *srcs[4] = {t->coefficients, t->flag0, t->flag1, t->flag2};
    for (int c = 0; c < 4; c++) VCU(cudaMemcpyAsync(in_col(c), srcs[c], os * 32, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<unsigned long long> perm(S);
    for (size_t i = 0; i < os; i++) {
        if (t->permuted_indices[i] >= S) return fail(ctx, SB_ERR_ARG, "permuted index out of range");
        perm[i] = t->permuted_indices[i];
    }
    for (size_t i = os; i < S; i++) perm[i] = i;                                   // verify.rs:41-42
    VCU(cudaMemcpyAsync(perm_d.p, perm.data(), S * 8, cudaMemcpyHostToDevice, ctx->stream));
    pw_u64_to_fp_kernel<<<nblk(S), 128, 0, ctx->stream>>>(nullptr, in_col(4), S);
    pw_u64_to_fp_kernel<<<nblk(S), 128, 0, ctx->stream>>>((const unsigned long long *)perm_d.p, in_col(5), S);
    ctx->launches += 2;
    VTRY(lde_dev(ctx, in_col(0), 6, S, S, g2, log_steps, LOG_EXTENSION_FACTOR, (uint4 *)ev6.p));
    // gather: 6 columns + xs at the V positions, xs at the public wires' first uses, xs[N - sk]
    const size_t np = t->n_pfi;
    std::vector<unsigned long long> gidx;
    for (int c = 0; c < 6; c++)
        for (size_t i = 0; i < V; i++) gidx.push_back((unsigned long long)c * N + pos32[i]);
    const size_t n_ev = gidx.size();
    std::vector<unsigned long long> xidx;
    for (size_t i = 0; i < V; i++) xidx.push_back(pos32[i]);
    for (size_t i = 0; i < np; i++) {
        if (t->pfi_w[i] >= S || t->pfi_k[i] >= t->n_public) return fail(ctx, SB_ERR_ARG, "public_first_indices out of range");
        xidx.push_back(sk * t->pfi_w[i]);
    }
    xidx.push_back(N - sk);
    std::vector<hfp::el> ev(n_ev), xv(xidx.size());
    VTRY(dpos.alloc((n_ev + xidx.size()) * 8));
    VTRY(dgath.alloc((n_ev + xidx.size()) * 32));
    VCU(cudaMemcpyAsync(dpos.p, gidx.data(), n_ev * 8, cudaMemcpyHostToDevice, ctx->stream));
    VCU(cudaMemcpyAsync((uint8_t *)dpos.p + n_ev * 8, xidx.data(), xidx.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->launches += merkle_launch_gather_bytes(ctx->stream, (const uint8_t *)ev6.p, 32, (const unsigned long long *)dpos.p, (uint32_t)n_ev, (uint8_t *)dgath.p);
    ctx->launches += merkle_launch_gather_bytes(ctx->stream, (const uint8_t *)xs, 32, (const unsigned long long *)dpos.p + n_ev, (uint32_t)xidx.size(),
                                                (uint8_t *)dgath.p + n_ev * 32);
    VCU(cudaMemcpyAsync(ev.data(), dgath.p, n_ev * 32, cudaMemcpyDeviceToHost, ctx->stream));
    VCU(cudaMemcpyAsync(xv.data(), (uint8_t *)dgath.p + n_ev * 32, xidx.size() * 32, cudaMemcpyDeviceToHost, ctx->stream));
    VCU(cudaStreamSynchronize(ctx->stream));
    VCU(cudaGetLastError());
    auto col_at = [&](int c, size_t i) { return ev[(size_t)c * V + i]; };

    // verify.rs:152-173 boundary data and challenges
    std::vector<hfp::el> pub_x(np), pub_y(np);
    for (size_t i = 0; i < np; i++) {
        pub_x[i] = xv[V + i];
        pub_y[i] = hfp::from_limbs(t->public_wires + 4 * t->pfi_k[i]);
    }
    const hfp::el x_last = xv[V + np];
    std::vector<hfp::el> interp2;                                              // calc_i2_polynomial once (verify.rs:152), Horner per position
    if (np >= INTERP_DEVICE_MIN && np <= INTERP_DEVICE_MAX) {                  // many public wires (`bits`: 1062): O(np^2) on the device, like the prover
        std::vector<hfp::el> zroot;
        VTRY(device_lagrange(ctx, (const uint4 *)((const uint8_t *)dgath.p + (n_ev + V) * 32), pub_y, interp2, zroot));
    } else {
        host_lagrange(interp2, pub_x, pub_y);
    }
    hfp::el r[3], k[11];
    {
        uint32_t idx[24];
        if (pseudorandom_indices(proof->a_root, 32, (uint32_t)N, 24, 0, idx, false) != SB_OK) return fail(ctx, SB_ERR_ARG, "sampler rejected precision %zu", N);
        for (int i = 0; i < 3; i++) {
            uint8_t b[32];
            for (int j = 0; j < 8; j++) {
                uint32_t v = idx[8 * i + j];
                b[4 * j] = (uint8_t)(v >> 24); b[4 * j + 1] = (uint8_t)(v >> 16); b[4 * j + 2] = (uint8_t)(v >> 8); b[4 * j + 3] = (uint8_t)v;
            }
            r[i] = hfp::from_bytes_le32(b);
        }
        k[0] = hfp::ONE;
        for (int i = 1; i < 11; i++) {
            uint8_t msg[33], h[32], le[32];
            memcpy(msg, proof->m_root, 32);
            msg[32] = (uint8_t)i;
            b2s::hash_bytes(h, msg, 33);
            for (int b = 0; b < 32; b++) le[b] = h[31 - b];
            k[i] = hfp::from_bytes_le32(le);
        }
    }

    // verify.rs:175-252 the spot checks
    for (size_t i = 0; i < V; i++) {
        const hfp::el x = xv[i];
        auto leaf = [&](size_t br, size_t f) { return hfp::from_bytes_le32(&proof->main_leaves[256 * (4 * i + br) + 32 * f]); };
        const hfp::el p_x = leaf(0, 0), p_prev = leaf(1, 0), p_w = leaf(2, 0), p_2w = leaf(3, 0);
        const hfp::el a_x = leaf(0, 1), a_prev = leaf(1, 1), s_x = leaf(0, 2);
        const hfp::el d1 = leaf(0, 3), d2 = leaf(0, 4), d3 = leaf(0, 5), b2 = leaf(0, 6), b3 = leaf(0, 7);
        const hfp::el x_s = hfp::pow_u64(x, S);                                          // verify.rs:234
        const hfp::el z = sub(x_s, hfp::ONE);                                            // z_evaluations[pos], utils.rs:173-178
        const hfp::el k_x = col_at(0, i), f0 = col_at(1, i), f1 = col_at(2, i), f2 = col_at(3, i);
        if (!hfp::eq(hfp::mul(f0, sub(sub(p_x, hfp::mul(f1, p_prev)), hfp::mul(k_x, s_x))), hfp::mul(z, d1)))
            return fail(ctx, SB_ERR_VERIFY, "spot check %zu: Q1(x) != Z(x) D1(x) (verify.rs:205)", i);
        if (!hfp::eq(hfp::mul(f2, sub(p_2w, hfp::mul(p_x, p_w))), hfp::mul(z, d2)))
            return fail(ctx, SB_ERR_VERIFY, "spot check %zu: Q2(x) != Z(x) D2(x) (verify.rs:211)", i);
        const hfp::el t2 = hfp::mul(r[2], s_x);
        const hfp::el vn = hfp::add(hfp::add(r[0], hfp::mul(r[1], col_at(4, i))), t2);
        const hfp::el vd = hfp::add(hfp::add(r[0], hfp::mul(r[1], col_at(5, i))), t2);
        if (!hfp::eq(sub(hfp::mul(a_x, vd), hfp::mul(a_prev, vn)), hfp::mul(z, d3)))
            return fail(ctx, SB_ERR_VERIFY, "spot check %zu: Q3(x) != Z(x) D3(x) (verify.rs:220)", i);
        hfp::el zb2 = hfp::ONE;
        for (size_t j = 0; j < np; j++) zb2 = hfp::mul(zb2, sub(x, pub_x[j]));           // verify.rs:223-226
        hfp::el i2 = hfp::ZERO;
        for (size_t j = np; j-- > 0;) i2 = hfp::add(hfp::mul(i2, x), interp2[j]);          // eval_poly_at (verify.rs:227)
        if (!hfp::eq(sub(s_x, i2), hfp::mul(zb2, b2))) return fail(ctx, SB_ERR_VERIFY, "spot check %zu: S(x) - I2(x) != Zb2(x) B2(x) (verify.rs:228)", i);
        if (!hfp::eq(sub(a_x, hfp::ONE), hfp::mul(sub(x, x_last), b3)))                 // I3 = 1 (utils.rs:458-463)
            return fail(ctx, SB_ERR_VERIFY, "spot check %zu: A(x) - I3(x) != Zb3(x) B3(x) (verify.rs:232)", i);
        const hfp::el l_x = hfp::from_bytes_le32(&proof->lc_leaves[32 * i]);
        hfp::el acc = hfp::mul(k[0], d1);
        acc = hfp::add(acc, hfp::mul(k[1], d2));
        acc = hfp::add(acc, hfp::mul(k[2], d3));
        acc = hfp::add(acc, hfp::mul(k[3], p_x));
        acc = hfp::add(acc, hfp::mul(hfp::mul(k[4], p_x), x_s));
        acc = hfp::add(acc, hfp::mul(k[5], b2));
        acc = hfp::add(acc, hfp::mul(hfp::mul(k[6], b2), x_s));
        acc = hfp::add(acc, hfp::mul(k[7], b3));
        acc = hfp::add(acc, hfp::mul(hfp::mul(k[8], b3), x_s));
        acc = hfp::add(acc, hfp::mul(k[9], a_x));
        acc = hfp::add(acc, hfp::mul(k[10], s_x));
        if (!hfp::eq(l_x, acc)) return fail(ctx, SB_ERR_VERIFY, "spot check %zu: linear combination mismatch (verify.rs:237)", i);
    }
    return SB_OK;
#undef VTRY
#undef VCU
    });
}

// ---- serde_json reader for StarkProof (run.rs:578 serde_json::from_reader; the layout sb_stark_proof_json writes) ------
namespace {
struct JsonIn {
    const char *p, *end;
    bool ok = true;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    bool eat(char c) {
        ws();
        if (p < end && *p == c) { p++; return true; }
        return false;
    }
    void need(char c) { if (!eat(c)) ok = false; }
    bool key(const char *name) {                       // "name":
        ws();
        const size_t n = strlen(name);
        if ((size_t)(end - p) < n + 3 || *p != '"' || memcmp(p + 1, name, n) != 0 || p[n + 1] != '"') { ok = false; return false; }
        p += n + 2;
        need(':');
        return ok;
    }
    bool bytes(std::vector<uint8_t> &out) {            // [1,2,3] appended
        need('[');
        if (eat(']')) return ok;
        while (ok) {
            ws();
            unsigned v = 0;
            int digits = 0;
            while (p < end && *p >= '0' && *p <= '9' && digits < 4) { v = v * 10 + (unsigned)(*p - '0'); p++; digits++; }
            if (!digits || v > 255) { ok = false; break; }
            out.push_back((uint8_t)v);
            if (eat(',')) continue;
            need(']');
            break;
        }
        return ok;
    }
    // [{"leaf":[..],"nodes":[[..],..]},..]
    bool branches(std::vector<uint8_t> &leaves, std::vector<uint8_t> &nodes, size_t &count, size_t &leaf_bytes, size_t &depth) {
        count = 0; leaf_bytes = 0; depth = 0;
        need('[');
        if (eat(']')) return ok;
        while (ok) {
            need('{');
            key("leaf");
            const size_t l0 = leaves.size();
            bytes(leaves);
            const size_t lb = leaves.size() - l0;
            need(',');
            key("nodes");
            need('[');
            size_t d = 0;
            if (!eat(']')) {
                while (ok) {
                    const size_t n0 = nodes.size();
                    bytes(nodes);
                    if (nodes.size() - n0 != 32) ok = false;
                    d++;
                    if (eat(',')) continue;
                    need(']');
                    break;
                }
            }
            need('}');
            if (count == 0) { leaf_bytes = lb; depth = d; } else if (lb != leaf_bytes || d != depth) ok = false;
            count++;
            if (eat(',')) continue;
            need(']');
            break;
        }
        return ok;
    }
};
}  // namespace

// [{"Middle":{"root2":..,"column_branches":..,"poly_branches":..}},..,{"Last":{"last":[[..],..]}}]  (fri.rs:16-26)
static void parse_fri_layers(JsonIn &j, sb_fri_proof *fri) {
    j.need('[');
    if (j.eat(']')) return;
    while (j.ok) {
        FriLayer L;
        j.need('{');
        j.ws();
        if (j.p + 6 < j.end && memcmp(j.p, "\"Last\"", 6) == 0) {
            j.key("Last");
            j.need('{');
            j.key("last");
            j.need('[');
            if (!j.eat(']')) {
                while (j.ok) {
                    const size_t n0 = L.last.size();
                    j.bytes(L.last);
                    if (L.last.size() - n0 != 32) j.ok = false;
                    if (j.eat(',')) continue;
                    j.need(']');
                    break;
                }
            }
            j.need('}');
            L.is_last = true;
        } else {
            j.key("Middle");
            j.need('{');
            std::vector<uint8_t> v;
            j.key("root2");
            j.bytes(v);
            if (v.size() != 32) j.ok = false; else memcpy(L.root2, v.data(), 32);
            j.need(',');
            size_t lb2 = 0;
            j.key("column_branches");
            j.branches(L.column_leaves, L.column_nodes, L.n_column, lb2, L.depth_column);
            if (L.n_column && lb2 != 32) j.ok = false;
            j.need(',');
            j.key("poly_branches");
            j.branches(L.poly_leaves, L.poly_nodes, L.n_poly, lb2, L.depth_poly);
            if (L.n_poly && lb2 != 32) j.ok = false;
            j.need('}');
        }
        j.need('}');
        fri->layers.push_back(std::move(L));
        if (j.eat(',')) continue;
        j.need(']');
        break;
    }
}

extern "C" int sb_stark_proof_from_json(const char *text, size_t len, sb_stark_proof **out) {
    return guarded((sb_ctx *)nullptr, __func__, [&]() -> int {
    if (!text || !out) return SB_ERR_ARG;
    JsonIn j{text, text + len};
    sb_stark_proof *p = new sb_stark_proof();
    p->fri = new sb_fri_proof();
    auto root = [&](const char *name, uint8_t dst[32]) {
        std::vector<uint8_t> v;
        j.key(name);
        j.bytes(v);
        if (v.size() != 32) j.ok = false; else memcpy(dst, v.data(), 32);
    };
    j.need('{');
    root("m_root", p->m_root); j.need(',');
    root("l_root", p->l_root); j.need(',');
    root("a_root", p->a_root); j.need(',');
    size_t cnt = 0, lb = 0, depth = 0, depth2 = 0;
    j.key("main_branches");
    j.branches(p->main_leaves, p->main_nodes, cnt, lb, depth);
    if (cnt && lb != 256) j.ok = false;
    j.need(',');
    j.key("linear_comb_branches");
    j.branches(p->lc_leaves, p->lc_nodes, cnt, lb, depth2);
    if ((cnt && lb != 32) || depth2 != depth) j.ok = false;
    p->depth = depth;
    j.need(',');
    j.key("fri_proof");
    parse_fri_layers(j, p->fri);
    j.need('}');
    j.ws();
    if (!j.ok || j.p != j.end) {
        delete p;
        return SB_ERR_ARG;
    }
    *out = p;
    return SB_OK;
    });
}

// verify_low_degree_proof (fri.rs:226-404) on the serde text of Vec<FriProof>: host only, needs no device.  ctx may be NULL
// (then the sampler keeps the reference's 2^24 limit and no error text is recorded).
extern "C" int sb_fri_verify_json(sb_ctx *ctx, const char *text, size_t len, const uint8_t merkle_root[32], const uint64_t root_of_unity[4],
                                  size_t n, size_t max_deg_plus_1, uint32_t exclude_multiples_of) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!text || !merkle_root || !root_of_unity) return SB_ERR_ARG;
    JsonIn j{text, text + len};
    sb_fri_proof fri;
    parse_fri_layers(j, &fri);
    j.ws();
    if (!j.ok || j.p != j.end) return SB_ERR_ARG;
    sb_ctx scratch;               // only err / extended_domain are touched by the host verifier
    return fri_verify_host(ctx ? ctx : &scratch, &fri, merkle_root, hfp::from_limbs(root_of_unity), n, max_deg_plus_1, exclude_multiples_of);
    });
}
