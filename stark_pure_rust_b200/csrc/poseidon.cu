// poseidon.cu -- host side of the alternative digest (poseidon.cuh): parameter generation, the host hash used for branch checks,
// launchers and the C ABI (`sb_poseidon_hash`, `sb_merkle_commit_poseidon`, `sb_poseidon_hash_host`).
//
// Replaces commitment/src/poseidon.rs:30-63 and the Poseidon instantiation of the Merkle trees
// (commitment/src/pallarel_merkle_tree.rs:219-253).  The parameters are those of neptune 5.1.0
// `PoseidonConstants::<Fr, U2>::new_with_strength(Strength::Standard)`: see poseidon.cuh and DESIGN.md 4.8.
#include <mutex>

#include "internal.h"
#include "kernels.h"
#include "poseidon.cuh"

namespace hbls {          // BLS12-381 scalar field on the host: 4 x u64 limbs, Montgomery form with R = 2^256
typedef unsigned __int128 u128;
struct el { uint64_t l[4]; };

static const el MOD = {{0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull}};
static const uint64_t NINV = 0xfffffffeffffffffull;        // -r^-1 mod 2^64 (checked in params())

static bool below_mod(const el &a) {
    for (int i = 3; i >= 0; i--)
        if (a.l[i] != MOD.l[i]) return a.l[i] < MOD.l[i];
    return false;
}
static void sub_mod(el &a) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        const u128 d = (u128)a.l[i] - MOD.l[i] - borrow;
        a.l[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
}
static el add(const el &a, const el &b) {
    el r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.l[i] + b.l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    if (!below_mod(r)) sub_mod(r);
    return r;
}
static el dbl_plain(const el &a) {      // 2a mod r for any a < r (not a Montgomery operation)
    return add(a, a);
}
static el mul(const el &a, const el &b) {
    uint64_t t[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a.l[j] * b.l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        const uint64_t hi0 = (uint64_t)c, hi1 = (uint64_t)(c >> 64);
        const uint64_t m = t[0] * NINV;
        c = ((u128)m * MOD.l[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * MOD.l[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += hi0;
        t[3] = (uint64_t)c;
        t[4] = hi1 + (uint64_t)(c >> 64);
    }
    el r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || !below_mod(r)) sub_mod(r);
    return r;
}
static el sbox(const el &x) {
    const el x2 = mul(x, x);
    return mul(mul(x2, x2), x);
}
} // namespace hbls

namespace {

struct PoseidonParams {
    hbls::el rc[POS_N_RC], mds[POS_T * POS_T], r2, one;     // Montgomery form; r2 = 2^512 mod r, one = 2^256 mod r
};

// Grain LFSR of the Poseidon paper, 80-bit state kept in the low bits of a 128-bit word (b0 = bit 79)
struct Grain {
    unsigned __int128 s = 0;
    int bit(int i) const { return (int)((s >> (79 - i)) & 1); }
    int step() {
        const int b = bit(62) ^ bit(51) ^ bit(38) ^ bit(23) ^ bit(13) ^ bit(0);
        s = ((s << 1) | (unsigned)b) & ((((unsigned __int128)1) << 80) - 1);
        return b;
    }
    int shrunk() {                        // self-shrinking mode: the second bit of a pair counts when the first is set
        for (;;) {
            const int a = step(), b = step();
            if (a) return b;
        }
    }
    void field(unsigned v, int width) { s = (s << width) | v; }
};

const PoseidonParams &params() {
    static PoseidonParams P;
    static std::once_flag once;
    std::call_once(once, [] {
        using namespace hbls;
        if ((uint64_t)(MOD.l[0] * NINV) != ~(uint64_t)0) abort();
        el x = {{1, 0, 0, 0}};
        for (int i = 0; i < 256; i++) x = dbl_plain(x);
        P.one = x;
        for (int i = 0; i < 256; i++) x = dbl_plain(x);
        P.r2 = x;
        Grain g;
        g.field(1, 2);                    // prime field
        g.field(1, 4);                    // S-box tag as neptune passes it
        g.field(255, 12);                 // bits of the modulus
        g.field(POS_T, 12);
        g.field(POS_RF, 10);
        g.field(POS_RP, 10);
        g.field(0x3fffffffu, 30);
        for (int i = 0; i < 160; i++) g.step();
        for (int n = 0; n < POS_N_RC;) {
            el c = {{0, 0, 0, 0}};
            for (int i = 254; i >= 0; i--) c.l[i >> 6] |= (uint64_t)g.shrunk() << (i & 63);
            if (!below_mod(c)) continue;  // rejection sampling
            P.rc[n++] = mul(c, P.r2);
        }
        // M[i][j] = 1 / (i + t + j): Fermat inverse of the small integers t .. 3t - 2
        el e = MOD;
        e.l[0] -= 2;
        for (int i = 0; i < POS_T; i++)
            for (int j = 0; j < POS_T; j++) {
                const el d = mul(el{{(uint64_t)(i + POS_T + j), 0, 0, 0}}, P.r2);
                el acc = P.one;
                for (int b = 254; b >= 0; b--) {
                    acc = mul(acc, acc);
                    if ((e.l[b >> 6] >> (b & 63)) & 1) acc = mul(acc, d);
                }
                P.mds[i * POS_T + j] = acc;
            }
    });
    return P;
}

// PoseidonDigest::hash on the host; false where the reference panics
bool hash_host(const uint8_t *msg, size_t len, uint8_t out[32]) {
    using namespace hbls;
    if (len == 0 || len > 64) return false;                 // poseidon.rs:33 (and the underflow of (len - 1) for an empty message)
    const PoseidonParams &P = params();
    uint8_t padded[64] = {0};
    memcpy(padded, msg, len);
    el st[POS_T];
    st[0] = mul(el{{3, 0, 0, 0}}, P.r2);
    for (int k = 0; k < 2; k++) {
        el c;
        memcpy(c.l, padded + 32 * k, 32);
        if (!below_mod(c)) return false;                    // Fr::from_bytes_le(..).unwrap(), poseidon.rs:38-48
        st[1 + k] = mul(c, P.r2);
    }
    for (int r = 0; r < POS_RF + POS_RP; r++) {
        for (int i = 0; i < POS_T; i++) st[i] = add(st[i], P.rc[r * POS_T + i]);
        const bool full = r < POS_RF / 2 || r >= POS_RF / 2 + POS_RP;
        st[0] = sbox(st[0]);
        if (full) {
            st[1] = sbox(st[1]);
            st[2] = sbox(st[2]);
        }
        el nx[POS_T];
        for (int j = 0; j < POS_T; j++) nx[j] = add(add(mul(st[0], P.mds[j]), mul(st[1], P.mds[POS_T + j])), mul(st[2], P.mds[2 * POS_T + j]));
        for (int j = 0; j < POS_T; j++) st[j] = nx[j];
    }
    const el d = mul(st[1], el{{1, 0, 0, 0}});
    memcpy(out, d.l, 32);
    return true;
}

// the device copy of the parameters, one per context (made on first use)
int device_consts(sb_ctx *ctx, const uint32_t **out) {
    if (!ctx->poseidon_consts) {
        const PoseidonParams &P = params();
        std::vector<uint32_t> h(POS_CONST_WORDS);
        memcpy(h.data(), P.rc, sizeof P.rc);
        memcpy(h.data() + POS_N_RC * 8, P.mds, sizeof P.mds);
        memcpy(h.data() + (POS_N_RC + POS_T * POS_T) * 8, P.r2.l, 32);
        uint32_t *d = nullptr;
        CU(cudaMalloc(&d, POS_CONST_WORDS * 4));
        cudaError_t e = cudaMemcpy(d, h.data(), POS_CONST_WORDS * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(d);
            return fail(ctx, SB_ERR_CUDA, "upload of the Poseidon parameters: %s", cudaGetErrorString(e));
        }
        ctx->poseidon_consts = d;
    }
    *out = ctx->poseidon_consts;
    return SB_OK;
}

unsigned blocks(size_t threads) { return (unsigned)((threads + POS_THREADS - 1) / POS_THREADS); }

// digests of n messages already on the device -> out; *d_err (device int, zeroed by the caller) flags a non-canonical chunk
int launch_leaves(sb_ctx *ctx, const uint8_t *d_msgs, size_t msg_bytes, size_t n, uint4 *d_out, int *d_err) {
    PoseidonLeavesParams P;
    P.msgs = d_msgs;
    P.out = d_out;
    P.n = n;
    P.msg_bytes = (uint32_t)msg_bytes;
    P.err = d_err;
    TRY(device_consts(ctx, &P.consts));
    if (n) {
        prof_begin(ctx, SB_KIND_MERKLE_LEAVES);
        poseidon_leaves_kernel<<<blocks(n), POS_THREADS, 0, ctx->stream>>>(P);
        ctx->launches++;
        prof_end(ctx);
    }
    return SB_OK;
}

int check_shape(sb_ctx *ctx, const void *msgs, size_t msg_bytes, size_t n) {
    if (msg_bytes == 0 || msg_bytes > 64) return fail(ctx, SB_ERR_ARG, "Poseidon messages are 1..64 bytes, got %zu (poseidon.rs:33)", msg_bytes);
    if (!msgs && n) return fail(ctx, SB_ERR_ARG, "messages is NULL");
    return SB_OK;
}

} // namespace

extern "C" int sb_poseidon_hash_host(const uint8_t *msg, size_t len, uint8_t out[32]) {
    if (!out || (!msg && len)) return SB_ERR_ARG;
    return hash_host(msg, len, out) ? SB_OK : SB_ERR_ARG;
}

extern "C" int sb_poseidon_hash(sb_ctx *ctx, const void *msgs, size_t msg_bytes, size_t n, uint8_t *out) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || (!out && n)) return SB_ERR_ARG;
    TRY(check_shape(ctx, msgs, msg_bytes, n));
    if (n == 0) return SB_OK;
    DevBuf in(ctx), dig(ctx), err(ctx);
    TRY(in.alloc(n * msg_bytes));
    TRY(dig.alloc(n * 32));
    TRY(err.alloc(sizeof(int)));
    CU(cudaMemsetAsync(err.p, 0, sizeof(int), ctx->stream));
    CU(cudaMemcpyAsync(in.p, msgs, n * msg_bytes, cudaMemcpyHostToDevice, ctx->stream));
    TRY(launch_leaves(ctx, (const uint8_t *)in.p, msg_bytes, n, (uint4 *)dig.p, (int *)err.p));
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out, dig.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (bad) return fail(ctx, SB_ERR_ARG, "a 32-byte chunk is not a canonical BLS12-381 scalar (poseidon.rs:48)");
    return SB_OK;
    });
}

extern "C" int sb_merkle_commit_poseidon(sb_ctx *ctx, const void *leaves, size_t leaf_bytes, size_t n, uint8_t root[32], sb_tree **tree) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !tree) return SB_ERR_ARG;
    TRY(check_shape(ctx, leaves, leaf_bytes, n));
    sb_tree *t = nullptr;
    TRY(tree_new(ctx, n, leaf_bytes, &t));
    struct Guard {                       // the tree is released on every early return
        sb_ctx *ctx;
        sb_tree *t;
        ~Guard() { if (t) sb_tree_free(ctx, t); }
    } guard{ctx, t};
    TRY(blk_alloc(ctx, n * leaf_bytes, (void **)&t->d_leaves));
    DevBuf err(ctx);
    TRY(err.alloc(sizeof(int)));
    CU(cudaMemsetAsync(err.p, 0, sizeof(int), ctx->stream));
    CU(cudaMemcpyAsync(t->d_leaves, leaves, n * leaf_bytes, cudaMemcpyHostToDevice, ctx->stream));
    TRY(launch_leaves(ctx, t->d_leaves, leaf_bytes, n, t->d_nodes, (int *)err.p));
    const uint32_t *consts;
    TRY(device_consts(ctx, &consts));
    for (uint32_t level = 0; level < t->depth; level++) {          // pallarel_merkle_tree.rs: parent = H(left || right)
        prof_begin(ctx, SB_KIND_MERKLE_NODES);
        poseidon_nodes_kernel<<<blocks(n >> (level + 1)), POS_THREADS, 0, ctx->stream>>>(t->d_nodes, n, level, consts);
        ctx->launches++;
        prof_end(ctx);
    }
    CU(cudaGetLastError());
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(t->root, (const uint8_t *)t->d_nodes + (2 * n - 2) * 32, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (bad) return fail(ctx, SB_ERR_ARG, "a leaf holds a 32-byte chunk that is not a canonical BLS12-381 scalar (poseidon.rs:48)");
    if (root) memcpy(root, t->root, 32);
    *tree = t;
    guard.t = nullptr;
    return SB_OK;
    });
}
