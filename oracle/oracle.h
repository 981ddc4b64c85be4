/*
 * oracle.h — CPU restatement of the stark-pure-rust hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This directory is the bit-exact checker for the CUDA product in
 * stark_pure_rust_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load or execute anything built from it.
 * The product library (libstark_b200.so) never links or calls this code.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   pinned by the reference's own KATs : Blake2s (commitment/src/utils.rs:13-24),
 *       index sampler (fri/src/utils.rs:112-120), Merkle shape/proofs
 *       (commitment/src/pallarel_merkle_tree.rs:133-216,
 *        commitment/src/merkle_proof_in_place.rs:209-261),
 *       field byte codecs (ff_utils/src/fp.rs:28-68), parsers (r1cs-stark/src/reader.rs:45-89).
 *       the Poseidon digest and its tree (commitment/src/poseidon.rs:66-106, pallarel_merkle_tree.rs:219-253).
 *   PARITY UNPINNED by the reference    : F_p products, NTT outputs, FRI columns/roots and
 *       proof.json have no golden vectors in the reference and the Rust code cannot be built
 *       here (no cargo/rustc).  They are anchored instead by (i) uniqueness of exact field
 *       arithmetic + canonical encodings, (ii) an independent Python big-int restatement
 *       (oracle/py_model.py) that must agree byte-for-byte, and (iii) the restated verifier
 *       (verify.rs) accepting every proof.
 *
 * Third-party arithmetic restated here (not under /root/reference):
 *   ff / ff_derive 0.10.0  (Cargo.lock:448-449,496-497) — Montgomery F_p, R = 2^256, 4 x u64 limbs
 *   blake2 0.9.1           (Cargo.lock:103-104)         — Blake2s-256, RFC 7693
 *   neptune 5.1.0, blstrs 0.4.1 (commitment/Cargo.toml:10,13) — Poseidon arity 2 over the BLS12-381 scalar field (poseidon.c)
 *   num-bigint 0.4.0, serde_json 1.0.66                 — byte<->integer and compact JSON
 */
#ifndef STARK_ORACLE_H
#define STARK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- field: BN254 Fr (ff_utils/src/fp.rs:8-12) ---------------------------------------- */
/* Element = 4 x u64 little-endian limbs in Montgomery form (x * 2^256 mod p), i.e. exactly
 * the in-memory layout of the reference's `Fp([u64; 4])`. */
typedef struct { uint64_t l[4]; } fp_t;

extern const fp_t FP_ZERO, FP_ONE /* = R mod p */, FP_R2;
extern const uint64_t FP_P[4];

void fp_add(fp_t *r, const fp_t *a, const fp_t *b);
void fp_sub(fp_t *r, const fp_t *a, const fp_t *b);
void fp_neg(fp_t *r, const fp_t *a);
void fp_mul(fp_t *r, const fp_t *a, const fp_t *b);
void fp_sqr(fp_t *r, const fp_t *a);
void fp_pow_u64(fp_t *r, const fp_t *a, uint64_t e);
void fp_pow_limbs(fp_t *r, const fp_t *a, const uint64_t *e, size_t n_limbs); /* pow_vartime */
void fp_inv(fp_t *r, const fp_t *a);               /* a != 0 ; Fermat */
int  fp_eq(const fp_t *a, const fp_t *b);
int  fp_is_zero(const fp_t *a);
void fp_from_u64(fp_t *r, uint64_t v);
/* fp.rs:70-77  from_bytes_le / from_bytes_be: integer value of ANY length, reduced mod p */
void fp_from_bytes_le(fp_t *r, const uint8_t *b, size_t len);
void fp_from_bytes_be(fp_t *r, const uint8_t *b, size_t len);
/* fp.rs:35-44  to_bytes_le / to_bytes_be: canonical value, exactly 32 bytes */
void fp_to_bytes_le(uint8_t out[32], const fp_t *a);
void fp_to_bytes_be(uint8_t out[32], const fp_t *a);
void orc_fp_to_bytes_le_vec(uint8_t *out, const fp_t *a, size_t n); /* n x fp_to_bytes_le */
void fp_multiplicative_generator(fp_t *r);        /* 7, fp.rs:10 */
/* g = 7^((p-1)/2^log_n): prove.rs:71-82 */
void fp_root_of_unity(fp_t *r, uint32_t log_n);

/* ---- Blake2s-256 (commitment/src/utils.rs:5-10, fri/src/utils.rs:5-10) ------------------- */
void orc_blake2s(uint8_t out[32], const uint8_t *msg, size_t len);

/* ---- NTT (fri/src/fft.rs) ---------------------------------------------------------------- */
/* fft.rs:5-14; writes at most cap elements, returns the order of the root */
size_t orc_expand_root_of_unity(fp_t *out, size_t cap, const fp_t *root);
void orc_serial_fft(fp_t *values, const fp_t *root, uint32_t log_n);               /* :150-193 */
void orc_parallel_fft(fp_t *values, const fp_t *root, uint32_t log_n, uint32_t log_cpus); /* :195-251 */
/* fft.rs:327-357 / :359-379.  values has room for 2^log_n elements, the first len_in are the
 * caller's vector, the rest is zero-padded here.  cpus mirrors Worker::cpus (multicore.rs:44). */
void orc_best_fft(fp_t *values, size_t len_in, const fp_t *root, uint32_t log_n, unsigned cpus);
void orc_inv_best_fft(fp_t *values, size_t len_in, const fp_t *root, uint32_t log_n, unsigned cpus);

/* ---- poly utils (fri/src/poly_utils.rs) -------------------------------------------------- */
void orc_multi_inv(fp_t *out, const fp_t *values, size_t n);                        /* :38-70 */
void orc_eval_poly_at(fp_t *r, const fp_t *poly, size_t n, const fp_t *x);          /* :93-102 */
void orc_zpoly(fp_t *out /* n+1 */, const fp_t *xs, size_t n);                      /* :362-373 */
void orc_lagrange_interp(fp_t *out /* n */, const fp_t *xs, const fp_t *ys, size_t n); /* :409-439 */
void orc_eval_quartic(fp_t *r, const fp_t p[4], const fp_t *x);                     /* :442-446 */
void orc_multi_interp_4(fp_t *out /* rows*4 */, const fp_t *xsets, const fp_t *ysets, size_t rows); /* :449-511 */

/* ---- sampler (fri/src/utils.rs:82-109) --------------------------------------------------- */
void orc_get_pseudorandom_indices(uint32_t *out, const uint8_t *seed, size_t seed_len,
                                  uint32_t modulus, size_t count, uint32_t exclude_multiples_of);

/* ---- Merkle (commitment/src/merkle_proof_in_place.rs:54-206, merkle_tree.rs:15-58) ------- */
/* leaves: n contiguous records of leaf_bytes each.  Writes the root and, for every requested
 * index (caller order, duplicates allowed), log2(n) sibling digests leaf level first.
 * Rebuilds the whole tree on every call exactly like gen_proofs does. */
void orc_merkle_gen_proofs(const uint8_t *leaves, size_t leaf_bytes, size_t n,
                           const size_t *indices, size_t n_idx,
                           uint8_t root[32], uint8_t *nodes_out /* n_idx*log2(n)*32 */);
/* merkle_tree.rs:25-43: returns 1 when the branch hashes to root */
int orc_merkle_validate(const uint8_t root[32], size_t index, const uint8_t *leaf, size_t leaf_bytes,
                        const uint8_t *nodes, size_t depth);

/* ---- the alternative digest (commitment/src/poseidon.rs:30-63; neptune 5.1.0 arity 2 over the BLS12-381 scalar field,
 * restated in poseidon.c; pinned by the KATs of poseidon.rs:66-106 and pallarel_merkle_tree.rs:235-246) ------------------ */
/* 0, or -1 where the reference panics: len == 0, len > 64, or a 32-byte chunk that is not a canonical scalar */
int orc_poseidon_hash(uint8_t out[32], const uint8_t *msg, size_t len);
int orc_poseidon_merkle_gen_proofs(const uint8_t *leaves, size_t leaf_bytes, size_t n, const size_t *indices, size_t n_idx,
                                   uint8_t root[32], uint8_t *nodes_out);
int orc_poseidon_merkle_validate(const uint8_t root[32], size_t index, const uint8_t *leaf, size_t leaf_bytes,
                                 const uint8_t *nodes, size_t depth);

/* ---- growable byte buffer used for JSON output ------------------------------------------- */
typedef struct { char *p; size_t len, cap; } orc_buf;
void orc_buf_free(orc_buf *b);

/* ---- FRI (fri/src/fri.rs) ---------------------------------------------------------------- */
typedef struct {
    uint8_t *leaf;      /* leaf_bytes */
    size_t leaf_bytes;
    uint8_t *nodes;     /* depth*32 */
    size_t depth;
} orc_branch;

typedef struct {
    int is_last;
    /* Middle (fri.rs:21-25) */
    uint8_t root2[32];
    orc_branch *column_branches; size_t n_column;   /* 40 */
    orc_branch *poly_branches;   size_t n_poly;     /* 160 */
    /* Last (fri.rs:18-20) */
    uint8_t *last; size_t n_last;                    /* n_last*32 */
} orc_fri_layer;

typedef struct { orc_fri_layer *layers; size_t n_layers; } orc_fri_proof;

/* fri.rs:46-224.  Optional taps (may be NULL): layer_roots receives the m_root of every Middle
 * layer (32 bytes each, up to max_layers), columns receives a malloc'd copy of every folded
 * column (caller frees). */
void orc_prove_low_degree(orc_fri_proof *out, const fp_t *values, size_t n, const fp_t *root,
                          size_t max_deg_plus_1, uint32_t exclude_multiples_of);
/* fri.rs:226-404: returns 1 when accepted, 0 otherwise */
int orc_verify_low_degree_proof(const uint8_t merkle_root[32], const fp_t *root,
                                const orc_fri_proof *proof, size_t max_deg_plus_1,
                                uint32_t exclude_multiples_of);
void orc_fri_proof_free(orc_fri_proof *p);
void orc_fri_proof_json(orc_buf *b, const orc_fri_proof *p);

/* ---- STARK prover / verifier (r1cs-stark/src/{run,prove,verify,utils,reader}.rs) ---------- */
typedef struct {
    /* circom2bellman_core/src/r1csfile.rs:29-59 */
    uint32_t field_size;
    uint8_t prime[32];
    uint32_t n_wires, n_pub_out, n_pub_in, n_priv_in;
    uint64_t n_labels;
    uint32_t n_constraints;
    /* flattened constraints: for constraint c and factor k (A,B,C): entries
     * [off[3c+k], off[3c+k+1]) of wire_id / value */
    size_t *off;
    uint32_t *wire_id;
    uint8_t *value;   /* 32 bytes each, LE */
} orc_r1cs;

int  orc_read_r1cs(orc_r1cs *out, const uint8_t *bytes, size_t len);  /* circom2bellman_core/src/reader.rs:4-89 */
void orc_r1cs_free(orc_r1cs *r);
/* r1cs-stark/src/reader.rs:7-42; returns malloc'd n_wires field elements */
fp_t *orc_read_witness(const uint8_t *bytes, size_t len, size_t *n_wires);

typedef struct {
    /* run.rs:109-308, 390-419 */
    size_t original_steps;           /* 3 * a_trace_len */
    fp_t *witness_trace, *computational_trace, *coefficients, *flag0, *flag1, *flag2;
    size_t *permuted_indices;        /* original_steps */
    size_t n_public;                 /* 1 + n_pub_in + n_pub_out */
    fp_t *public_wires;
    size_t n_pfi;                    /* public_first_indices */
    size_t *pfi_k, *pfi_w;
    size_t n_constraints, n_wires;
} orc_trace;

void orc_build_trace(orc_trace *t, const orc_r1cs *r, const fp_t *witness, int with_witness);
void orc_trace_free(orc_trace *t);

typedef struct {
    uint8_t m_root[32], l_root[32], a_root[32];
    orc_branch *main_branches; size_t n_main;       /* 320 */
    orc_branch *lc_branches;   size_t n_lc;         /* 80  */
    orc_fri_proof fri;
} orc_stark_proof;

/* Optional taps for intermediate-parity tests: when non-NULL the prover stores malloc'd copies. */
typedef struct {
    size_t steps, precision;
    fp_t *lde[9];      /* k f0 f1 f2 s p idx pidx a : N each */
    fp_t *tree_cols[8];/* p a s d1 d2 d3 b2 b3 : N each */
    fp_t *l_evals;     /* N */
    fp_t r[3], k[11];
    uint32_t positions[80];
} orc_prove_taps;
void orc_prove_taps_free(orc_prove_taps *t);

/* prove.rs:14-378.  cpus: the Worker::cpus value the NTTs dispatch on (results do not depend on it). */
void orc_mk_r1cs_proof(orc_stark_proof *out, const orc_trace *t, unsigned cpus, orc_prove_taps *taps);
/* verify.rs:13-258 */
int orc_verify_r1cs_proof(const orc_stark_proof *proof, const orc_trace *t, unsigned cpus);
void orc_stark_proof_free(orc_stark_proof *p);
/* serde_json::to_string(&StarkProof) — run.rs:549, utils.rs:122-130 */
void orc_stark_proof_json(orc_buf *b, const orc_stark_proof *p);

/* convenience used by the CLI and ctypes: files in, JSON out.  Returns 0 on success, 1 when the
 * restated verifier rejects, <0 on IO/format errors. */
int orc_prove_files(const char *r1cs_path, const char *wtns_path, const char *proof_path,
                    unsigned cpus, int verify, double *t_prove_s);

/* trace arrays of a circuit for the test harness (caller frees with orc_trace_free) */
int orc_trace_from_files(const char *r1cs_path, const char *wtns_path, orc_trace *out);

/* per-stage wall-clock of the last orc_mk_r1cs_proof in this thread (seconds):
 * [0]=ntt/lde [1]=merkle [2]=fri [3]=pointwise+rest */
extern __thread double orc_stage_s[4];

#ifdef __cplusplus
}
#endif
#endif
