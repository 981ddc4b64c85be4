// frontend.cu -- host front end of the prover: iden3 .r1cs / .wtns parsing and the R1CS -> trace arrangement
// the reference does before calling mk_r1cs_proof.  O(nnz) scalar work, stays on the CPU by design
// (SURVEY.md §8f next-2); the vectors it produces go straight into sb_prove_r1cs.
//
//   read_r1cs      circom2bellman_core/src/reader.rs:4-89
//   read_witness   r1cs-stark/src/reader.rs:7-42
//   build_trace    r1cs-stark/src/run.rs:109-308 (calc_coefficients_and_witness, calc_flags),
//                  :390-419 (permuted_indices, public_first_indices), :344-361 (prime / witness[0] checks)
//   sb_prove_files r1cs-stark/src/run.rs:528-554 (prove_with_file_path)
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "internal.h"

#include <algorithm>
#include <atomic>
#include <thread>

namespace {

struct Reader {
    const uint8_t *p;
    size_t left;
    bool ok = true;
    bool need(size_t n) {
        if (left < n) ok = false;
        return ok;
    }
    uint32_t u32() {
        if (!need(4)) return 0;
        uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        p += 4;
        left -= 4;
        return v;
    }
    uint64_t u64() {
        uint64_t lo = u32(), hi = u32();
        return lo | (hi << 32);
    }
    void bytes(uint8_t *out, size_t n) {
        if (!need(n)) {
            memset(out, 0, n);
            return;
        }
        memcpy(out, p, n);
        p += n;
        left -= n;
    }
};

// One A / B / C factor of a constraint, left in place in the file image: n terms of (u32 wire, 32-byte LE coefficient)
struct FactorRef {
    const uint8_t *terms;
    uint32_t n;
};
struct R1cs {
    uint32_t field_size = 0, n_wires = 0, n_pub_out = 0, n_pub_in = 0, n_priv = 0, n_constraints = 0;
    uint64_t n_labels = 0;
    uint8_t prime[32];
    std::vector<FactorRef> factors;             // 3 per constraint: A, B, C
    std::vector<size_t> row_off;                // row_off[c] = rows of the constraints before c (run.rs:128-135: max of the three lengths)
};

const uint8_t BN254_FR_LE[32] = {1, 0, 0, 240, 147, 245, 225, 67, 145, 112, 185, 121, 72, 232, 51, 40,
                                 93, 88, 129, 129, 182, 69, 80, 184, 41, 160, 49, 225, 114, 78, 100, 48};   // run.rs:344-350

// A file as a read-only memory image: mapped (the parsers below leave the constraint terms in place, so a 17 MB .r1cs is never
// copied -- fread of it took 2.9 ms of the 2^23 proof's front end), read into a buffer where mapping is not possible.
// (The file must not be truncated while the proof runs.)
struct FileImage {
    const uint8_t *p = nullptr;
    size_t n = 0;
    bool mapped = false;
    std::vector<uint8_t> buf;
    bool open(const char *path) {
        const int fd = ::open(path, O_RDONLY | O_CLOEXEC);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
            ::close(fd);
            return slurp_into(path);
        }
        n = (size_t)st.st_size;
        if (n == 0) {
            ::close(fd);
            p = (const uint8_t *)"";
            return true;
        }
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        ::close(fd);
        if (m == MAP_FAILED) return slurp_into(path);
        p = (const uint8_t *)m;
        mapped = true;
        return true;
    }
    bool slurp_into(const char *path);
    ~FileImage() {
        if (mapped) munmap((void *)p, n);
    }
    FileImage() = default;
    FileImage(const FileImage &) = delete;
    FileImage &operator=(const FileImage &) = delete;
};

bool slurp(const char *path, std::vector<uint8_t> &out) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(sz > 0 ? (size_t)sz : 0);
    bool ok = sz <= 0 || fread(out.data(), 1, (size_t)sz, f) == (size_t)sz;
    fclose(f);
    return ok;
}

bool FileImage::slurp_into(const char *path) {
    if (!slurp(path, buf)) return false;
    p = buf.data();
    n = buf.size();
    return true;
}

// body(lo, hi) over [0, n) on up to 16 host threads (the O(nnz) scalar work below: ~2 Montgomery products per term)
template <class F>
void parallel_for(size_t n, size_t min_grain, F body) {
    unsigned hw = std::thread::hardware_concurrency();
    size_t T = std::min<size_t>(std::min<unsigned>(hw ? hw : 1, 16), n / (min_grain ? min_grain : 1));
    if (T <= 1) {
        body((size_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    size_t started = 0;
    for (; started + 1 < T; started++) {
        const size_t i = started;
        try {
            th.emplace_back([=]() { body(n * i / T, n * (i + 1) / T); });
        } catch (...) {
            break;                  // no more threads to be had: this thread takes the rest
        }
    }
    body(n * started / T, n);
    for (auto &t : th) t.join();
}

// reader.rs:4-89: magic, version 1, 3 sections, header section first, then the constraint section; labels ignored.
// Only the structure is decoded here (one u32 per factor); coefficients are converted where they are used.
const char *read_r1cs(const uint8_t *data, size_t size, R1cs &r) {
    Reader p{data, size};
    if (p.u32() != 0x73633172u) return "not an r1cs file (magic)";
    if (p.u32() != 1) return "r1cs version must be 1";
    if (p.u32() != 3) return "r1cs must have 3 sections";
    if (p.u32() != 1) return "first r1cs section must be the header";
    p.u64();
    r.field_size = p.u32();
    p.bytes(r.prime, 32);
    r.n_wires = p.u32();
    r.n_pub_out = p.u32();
    r.n_pub_in = p.u32();
    r.n_priv = p.u32();
    r.n_labels = p.u64();
    r.n_constraints = p.u32();
    if (p.u32() != 2) return "second r1cs section must be the constraints";
    p.u64();
    if (!p.ok) return "truncated r1cs header";
    // the third section maps every wire to a u64 label (reader.rs:66-74), so a well-formed file holds 8 bytes per wire:
    // bounds n_wires before anything is sized by it (the verifier has no witness to compare it with)
    if ((uint64_t)r.n_wires * 8 > size) return "r1cs header declares more wires than the file can describe";
    if ((size_t)r.n_constraints * 12 > p.left) return "truncated r1cs constraints";
    r.factors.resize((size_t)3 * r.n_constraints);
    r.row_off.assign((size_t)r.n_constraints + 1, 0);
    for (size_t c = 0; c < r.n_constraints; c++) {
        size_t rows = 0;
        for (int k = 0; k < 3; k++) {
            uint32_t n = p.u32();
            if (!p.ok || (size_t)n * 36 > p.left) return "truncated r1cs constraints";
            r.factors[3 * c + k] = FactorRef{p.p, n};
            p.p += (size_t)n * 36;
            p.left -= (size_t)n * 36;
            rows = std::max<size_t>(rows, n);
        }
        r.row_off[c + 1] = r.row_off[c] + rows;
    }
    return p.ok ? nullptr : "truncated r1cs constraints";
}

// reader.rs:7-42: "wtns", 5 skipped words, field size, modulus, n_wires, 3 skipped words, the values
const char *read_r1cs(const std::vector<uint8_t> &buf, R1cs &r) { return read_r1cs(buf.data(), buf.size(), r); }

const char *read_witness(const uint8_t *data, size_t size, std::vector<hfp::el> &w) {
    Reader p{data, size};
    if (p.u32() != 1936618615u) return "not a wtns file (magic)";
    for (int i = 0; i < 5; i++) p.u32();
    uint32_t field_size = p.u32();
    if (field_size != 32) return "witness field size must be 32 bytes";
    uint8_t tmp[32];
    p.bytes(tmp, 32);
    uint32_t n = p.u32();
    p.u32();
    p.u32();
    p.u32();
    if (!p.ok || (size_t)n * 32 > p.left) return "truncated wtns file";
    w.resize(n);
    const uint8_t *vals = p.p;
    parallel_for(n, 4096, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) w[i] = hfp::from_bytes_le32(vals + 32 * i);     // run.rs:353-357
    });
    return nullptr;
}

const char *read_witness(const std::vector<uint8_t> &buf, std::vector<hfp::el> &w) { return read_witness(buf.data(), buf.size(), w); }

// the six original_steps-long vectors mk_r1cs_proof takes, carved out of the context's pinned staging arena so that
// their upload runs at PCIe speed (the copy permutation sits behind them in the same arena); public data are small and stay
// in ordinary vectors
struct Trace {
    size_t os = 0;
    hfp::el *wit = nullptr, *comp = nullptr, *coef = nullptr, *f0 = nullptr, *f1 = nullptr, *f2 = nullptr;
    std::vector<hfp::el> pub, heap;
    size_t *perm = nullptr;                     // original_steps entries, in the arena (or perm_heap)
    std::vector<unsigned long long> last_rows;  // last row of every constraint (the flag vectors in compressed form)
    size_t a = 0;
    std::vector<uint32_t> wire_at;              // wire of every row (input of the copy permutation)
    std::vector<size_t> perm_heap, pfi_k, pfi_w;
};

// run.rs:109-281, :283-308, :390-419
// with_witness = false (verifier, run.rs:454-526): only the public part is built -- coefficients, flags, permutation,
// public wires and their first uses; `witness` then only needs the public wires
// flags_on_device: the three flag vectors are not materialised (sb_prove_files: the prover generates them from last_rows)
static double fe_now() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static bool fe_trace() {
    static const bool on = getenv("SB_TRACE_FRONT") != nullptr;      // phase times of the front end on stderr
    return on;
}
#define FE_LAP(what) do { if (fe_trace()) { const double t_ = fe_now(); fprintf(stderr, "[front end] %-24s %7.3f ms\n", what, t_ - fe_t0); fe_t0 = t_; } } while (0)

const char *finish_trace(const R1cs &r, const std::vector<hfp::el> &witness, Trace &t);

// defer_perm: stop after the row vectors; the caller runs finish_trace (copy permutation, public wires) later -- sb_prove_files
// does that inside the prover, while the device already transforms the columns that do not depend on it
const char *build_trace(sb_ctx *ctx, const R1cs &r, const std::vector<hfp::el> &witness, Trace &t, bool with_witness = true, bool flags_on_device = false,
                        bool defer_perm = false) {
    double fe_t0 = fe_now();
    const size_t n_wires = r.n_wires, nc = r.n_constraints;
    if ((with_witness && witness.size() < n_wires) || n_wires == 0) return "witness shorter than the circuit's wire count";
    const size_t a = r.row_off[nc], os = 3 * a;
    if (a == 0) return "circuit has no constraint rows";
    hfp::el *arena;
    if (ctx) {
        // always sized for six vectors: the verifier (which materialises the flag vectors) then reuses the prover's arena
        // instead of re-allocating ~200 MB of pinned memory (~100 ms) on the first verification after a proof
        const size_t n_vec = flags_on_device ? 3 : 6;
        arena = (hfp::el *)pinned_arena(ctx, 6 * os * sizeof(hfp::el) + os * sizeof(size_t));
        if (arena) t.perm = (size_t *)(arena + n_vec * os);
    } else {                      // host-only use (sb_trace_from_files): ordinary memory owned by the Trace
        t.heap.resize(6 * os);
        arena = t.heap.data();
        t.perm_heap.assign(os, 0);
        t.perm = t.perm_heap.data();
    }
    if (!arena) return "cannot allocate the pinned staging arena";
    t.os = os;
    t.a = a;
    if (flags_on_device && ctx) {
        t.coef = arena; t.wit = arena + os; t.comp = arena + 2 * os;
    } else {
        flags_on_device = false;
        t.coef = arena; t.f0 = arena + os; t.f1 = arena + 2 * os; t.f2 = arena + 3 * os; t.wit = arena + 4 * os; t.comp = arena + 5 * os;
    }
    std::vector<uint32_t> &wire_at = t.wire_at;
    wire_at.resize(os);
    std::atomic<int> bad(0);
    FE_LAP("arena + wire_at");
    // rows of constraint c sit at row_off[c] .. in each third k (A, B, C); threads take ranges of constraints
    parallel_for(nc, 64, [&](size_t c0, size_t c1) {
        for (size_t c = c0; c < c1; c++) {
            const size_t off = r.row_off[c], n = r.row_off[c + 1] - off;
            for (int k = 0; k < 3; k++) {
                const FactorRef &f = r.factors[3 * c + k];
                hfp::el run = hfp::ZERO;
                for (size_t i = 0; i < n; i++) {
                    const size_t pos = (size_t)k * a + off + i;
                    uint32_t w;
                    if (i < f.n) {
                        const uint8_t *term = f.terms + 36 * i;
                        memcpy(&w, term, 4);
                        if (w >= n_wires) {
                            bad.store(1);
                            w = 0;
                        }
                        const hfp::el cf = hfp::from_bytes_le32(term + 4);     // T::from_bytes_le(value), run.rs:156
                        if (with_witness) run = hfp::add(run, hfp::mul(cf, witness[w]));
                        t.coef[pos] = cf;
                    } else {                                           // padding row: LAST wire, coefficient 0 (run.rs:165-176)
                        w = (uint32_t)(n_wires - 1);
                        t.coef[pos] = hfp::ZERO;
                    }
                    wire_at[pos] = w;
                    if (with_witness) {
                        t.wit[pos] = witness[w];
                        t.comp[pos] = run;
                    }
                }
            }
        }
    });
    if (bad.load()) return "wire id out of range";
    FE_LAP("rows (coef, wit, comp)");
    // calc_flags, run.rs:283-308
    for (size_t c = 0; c < nc; c++)
        if (r.row_off[c + 1] != r.row_off[c]) t.last_rows.push_back(r.row_off[c + 1] - 1);     // (every well-formed constraint has a term)
    if (!flags_on_device) {
        parallel_for(os, 1 << 16, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; i++) {
                t.f0[i] = hfp::ONE;
                t.f1[i] = hfp::ONE;
                t.f2[i] = hfp::ZERO;
            }
        });
        for (unsigned long long l : t.last_rows) {
            const size_t k = (l + 1) % a;
            t.f1[k] = t.f1[k + a] = t.f1[k + 2 * a] = hfp::ZERO;
            t.f2[l] = hfp::ONE;
        }
    }
    // copy permutation, run.rs:390-401: the uses of a wire in (constraint, factor, row) order form a cycle,
    // perm[first use] = last use, perm[use j] = use j-1
    FE_LAP("flags");
    return defer_perm ? nullptr : finish_trace(r, witness, t);
}

// copy permutation, run.rs:390-401: the uses of a wire in (constraint, factor, row) order form a cycle,
// perm[first use] = last use, perm[use j] = use j-1
const char *finish_trace(const R1cs &r, const std::vector<hfp::el> &witness, Trace &t) {
    double fe_t0 = fe_now();
    const size_t n_wires = r.n_wires, nc = r.n_constraints, a = t.a, os = t.os;
    const std::vector<uint32_t> &wire_at = t.wire_at;
    const size_t NONE = (size_t)-1;
    std::vector<size_t> first(n_wires, NONE), prev(n_wires, NONE);
    memset(t.perm, 0, os * sizeof(size_t));
    for (size_t c = 0; c < nc; c++) {
        const size_t off = r.row_off[c], n = r.row_off[c + 1] - off;
        for (int k = 0; k < 3; k++) {
            for (size_t i = 0; i < n; i++) {
                const size_t pos = (size_t)k * a + off + i;
                const uint32_t w = wire_at[pos];
                if (first[w] == NONE) first[w] = pos; else t.perm[pos] = prev[w];
                prev[w] = pos;
            }
        }
    }
    for (size_t w = 0; w < n_wires; w++)
        if (first[w] != NONE) t.perm[first[w]] = prev[w];
    FE_LAP("permutation");
    // public wires and their first uses, run.rs:359-361, :413-419
    const size_t n_pub = 1 + (size_t)r.n_pub_in + r.n_pub_out;
    if (n_pub > witness.size() || n_pub > n_wires) return "more public wires than wires";
    t.pub.assign(witness.begin(), witness.begin() + n_pub);
    for (size_t w = 0; w < n_pub; w++) {
        if (first[w] != NONE) {
            t.pfi_k.push_back(w);
            t.pfi_w.push_back(first[w]);
        }
    }
    return nullptr;
}

}  // namespace

// prove_with_file_path (run.rs:528-554): files in, proof.json out (no trailing newline, run.rs:551).
// proof_path may be NULL (timing only).  stage_ms (may be NULL): [0] LDE [1] m_tree [2] FRI [3] rest [4] GPU total,
// [5] host front end (parse + trace arrangement), [6] JSON serialisation + write.
extern "C" int sb_prove_files(sb_ctx *ctx, const char *r1cs_path, const char *wtns_path, const char *proof_path, double stage_ms[7]) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !r1cs_path || !wtns_path) return SB_ERR_ARG;
    dbg_check("sb_prove_files entry");
    auto now = []() {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    };
    const double t0 = now();
    double fe_t0 = t0;
    FileImage rb, wb;
    if (!rb.open(r1cs_path)) return fail(ctx, SB_ERR_ARG, "cannot read %s", r1cs_path);
    if (!wb.open(wtns_path)) return fail(ctx, SB_ERR_ARG, "cannot read %s", wtns_path);
    FE_LAP("read files");
    R1cs r;
    std::vector<hfp::el> witness;
    const char *e = read_r1cs(rb.p, rb.n, r);
    if (e) return fail(ctx, SB_ERR_ARG, "%s: %s", r1cs_path, e);
    if (memcmp(r.prime, BN254_FR_LE, 32) != 0) return fail(ctx, SB_ERR_ARG, "%s: field is not BN254 Fr (run.rs:344-350)", r1cs_path);
    e = read_witness(wb.p, wb.n, witness);
    if (e) return fail(ctx, SB_ERR_ARG, "%s: %s", wtns_path, e);
    if (witness.empty() || !hfp::eq(witness[0], hfp::ONE)) return fail(ctx, SB_ERR_ARG, "witness[0] must be 1 (run.rs:358)");
    dbg_check("sb_prove_files before build_trace");
    FE_LAP("parse");
    Trace t;
    e = build_trace(ctx, r, witness, t, true, true);
    if (e) return fail(ctx, SB_ERR_ARG, "%s", e);
    dbg_check("sb_prove_files after build_trace");
    sb_trace st;
    memset(&st, 0, sizeof st);
    st.original_steps = t.os;
    st.witness_trace = (const uint64_t *)t.wit;
    st.computational_trace = (const uint64_t *)t.comp;
    st.coefficients = (const uint64_t *)t.coef;
    st.flag0 = (const uint64_t *)t.f0;
    st.flag1 = (const uint64_t *)t.f1;
    st.flag2 = (const uint64_t *)t.f2;
    st.permuted_indices = t.perm;
    st.n_public = t.pub.size();
    st.public_wires = (const uint64_t *)t.pub.data();
    st.n_pfi = t.pfi_k.size();
    st.pfi_k = t.pfi_k.data();
    st.pfi_w = t.pfi_w.data();
    // (Tried: building the copy permutation -- 1.8 ms of serial host work at 955 086 steps -- inside the prover, after the
    // transforms that do not need it had been queued.  The device timeline stayed gap-free, yet the proof took 1.9 ms longer
    // than with the permutation built up front, so it is built here.)
    const double t1 = now();
    sb_stark_proof *proof = nullptr;
    dbg_check("sb_prove_files after the front end");
    const FlagSpec flags{t.last_rows.data(), t.last_rows.size(), t.a};
    TRY(prove_r1cs_impl(ctx, &st, t.f0 ? nullptr : &flags, &proof));
    const double t2 = now();
    int rc = SB_OK;
    if (proof_path) {
        size_t len = 0;
        char *s = sb_stark_proof_json(proof, &len);
        FILE *f = s ? fopen(proof_path, "wb") : nullptr;
        if (!f || fwrite(s, 1, len, f) != len) rc = fail(ctx, SB_ERR_ARG, "cannot write %s", proof_path);
        if (f) fclose(f);
        free(s);
    }
    const double t3 = now();
    if (stage_ms) {
        sb_stark_proof_stage_ms(proof, stage_ms);
        stage_ms[4] = t2 - t1;          // wall clock of sb_prove_r1cs (includes host scalar work between kernels)
        stage_ms[5] = t1 - t0;
        stage_ms[6] = t3 - t2;
    }
    sb_stark_proof_free(proof);
    return rc;
    });
}

// verify_with_file_path (run.rs:556-590): the public wires are the head of the witness file (run.rs:582-585)
extern "C" int sb_verify_files(sb_ctx *ctx, const char *r1cs_path, const char *wtns_path, const char *proof_path, double verify_ms[2]) {
    return guarded(ctx, __func__, [&]() -> int {
    if (!ctx || !r1cs_path || !wtns_path || !proof_path) return SB_ERR_ARG;
    auto now = []() {
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    };
    const double t0 = now();
    std::vector<uint8_t> rb, wb, pb;
    if (!slurp(r1cs_path, rb)) return fail(ctx, SB_ERR_ARG, "cannot read %s", r1cs_path);
    if (!slurp(wtns_path, wb)) return fail(ctx, SB_ERR_ARG, "cannot read %s", wtns_path);
    if (!slurp(proof_path, pb)) return fail(ctx, SB_ERR_ARG, "cannot read %s", proof_path);
    R1cs r;
    std::vector<hfp::el> witness;
    const char *e = read_r1cs(rb, r);
    if (e) return fail(ctx, SB_ERR_ARG, "%s: %s", r1cs_path, e);
    if (memcmp(r.prime, BN254_FR_LE, 32) != 0) return fail(ctx, SB_ERR_ARG, "%s: field is not BN254 Fr (run.rs:344-350)", r1cs_path);
    e = read_witness(wb, witness);
    if (e) return fail(ctx, SB_ERR_ARG, "%s: %s", wtns_path, e);
    if (witness.empty() || !hfp::eq(witness[0], hfp::ONE)) return fail(ctx, SB_ERR_ARG, "public_wires[0] must be 1 (run.rs:479)");
    Trace t;
    e = build_trace(ctx, r, witness, t, false);
    if (e) return fail(ctx, SB_ERR_ARG, "%s", e);
    sb_stark_proof *proof = nullptr;
    if (sb_stark_proof_from_json((const char *)pb.data(), pb.size(), &proof) != SB_OK)
        return fail(ctx, SB_ERR_ARG, "%s: not a serialised StarkProof", proof_path);
    sb_trace st;
    memset(&st, 0, sizeof st);
    st.original_steps = t.os;
    st.coefficients = (const uint64_t *)t.coef;
    st.flag0 = (const uint64_t *)t.f0;
    st.flag1 = (const uint64_t *)t.f1;
    st.flag2 = (const uint64_t *)t.f2;
    st.permuted_indices = t.perm;
    st.n_public = t.pub.size();
    st.public_wires = (const uint64_t *)t.pub.data();
    st.n_pfi = t.pfi_k.size();
    st.pfi_k = t.pfi_k.data();
    st.pfi_w = t.pfi_w.data();
    const double t1 = now();
    int rc = sb_verify_r1cs(ctx, &st, proof);
    const double t2 = now();
    sb_stark_proof_free(proof);
    if (verify_ms) {
        verify_ms[0] = t1 - t0;
        verify_ms[1] = t2 - t1;
    }
    return rc;
    });
}

// The arguments run.rs:390-419 hands to mk_r1cs_proof, built on the host alone (no device, no context): lets the front
// end be checked without a GPU and lets a caller keep the trace around for several sb_prove_r1cs / sb_verify_r1cs calls.
struct sb_host_trace {
    Trace t;
    sb_trace view;
};
extern "C" int sb_trace_from_files(const char *r1cs_path, const char *wtns_path, sb_host_trace **out, const sb_trace **view) {
    return guarded((sb_ctx *)nullptr, __func__, [&]() -> int {
    if (!r1cs_path || !wtns_path || !out || !view) return SB_ERR_ARG;
    std::vector<uint8_t> rb, wb;
    if (!slurp(r1cs_path, rb) || !slurp(wtns_path, wb)) return SB_ERR_ARG;
    R1cs r;
    std::vector<hfp::el> witness;
    if (read_r1cs(rb, r) || memcmp(r.prime, BN254_FR_LE, 32) != 0) return SB_ERR_ARG;
    if (read_witness(wb, witness) || witness.empty() || !hfp::eq(witness[0], hfp::ONE)) return SB_ERR_ARG;
    sb_host_trace *h = new sb_host_trace();
    if (build_trace(nullptr, r, witness, h->t)) {
        delete h;
        return SB_ERR_ARG;
    }
    const Trace &t = h->t;
    sb_trace &st = h->view;
    memset(&st, 0, sizeof st);
    st.original_steps = t.os;
    st.witness_trace = (const uint64_t *)t.wit;
    st.computational_trace = (const uint64_t *)t.comp;
    st.coefficients = (const uint64_t *)t.coef;
    st.flag0 = (const uint64_t *)t.f0;
    st.flag1 = (const uint64_t *)t.f1;
    st.flag2 = (const uint64_t *)t.f2;
    st.permuted_indices = t.perm;
    st.n_public = t.pub.size();
    st.public_wires = (const uint64_t *)t.pub.data();
    st.n_pfi = t.pfi_k.size();
    st.pfi_k = t.pfi_k.data();
    st.pfi_w = t.pfi_w.data();
    *out = h;
    *view = &h->view;
    return SB_OK;
    });
}
extern "C" void sb_host_trace_free(sb_host_trace *h) { delete h; }
