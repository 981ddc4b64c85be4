/*
 * stark_b200.h -- C ABI of the B200 (sm_100a) backend for the stark-pure-rust hot path:
 * NTT / low-degree extension, Blake2s Merkle commitment, FRI low-degree proof.
 *
 * This is the boundary a Rust FFI shim crate would bind so that the reference's own entry points
 * keep their signatures (packages/r1cs-stark/src/prove.rs:2-10 imports exactly these):
 *
 *   fri::fft::best_fft / inv_best_fft          packages/fri/src/fft.rs:327-331, :359-363
 *   fri::fft::expand_root_of_unity             packages/fri/src/fft.rs:5-14
 *   fri::poly_utils::multi_inv                 packages/fri/src/poly_utils.rs:38-70
 *   commitment::MerkleProofInPlace             packages/commitment/src/merkle_proof_in_place.rs:16-50
 *     (trait MerkleTree: width/get_root/update/gen_proofs, merkle_tree.rs:60-73)
 *   fri::fri::prove_low_degree                 packages/fri/src/fri.rs:46-62
 *   the inv_best_fft -> best_fft pairs         packages/r1cs-stark/src/prove.rs:100-124,160-167,183-184
 *
 * Conventions
 *   - Field element = the reference's in-memory `Fp([u64; 4])` (ff_utils/src/fp.rs:8-12): BN254 Fr,
 *     Montgomery form with R = 2^256, four little-endian u64 limbs, canonical (< p).  A Rust
 *     `Vec<Fp>` is passed as `uint64_t*` by pointer cast, no conversion.
 *   - Digest = 32 bytes, Blake2s-256 unkeyed (commitment/src/blake.rs:28-32).
 *   - Every function returns SB_OK (0) or a negative error; sb_last_error() gives the text.  The
 *     reference panics on these conditions (assert!/unwrap); a Rust shim turns non-zero into panic!.
 *   - There is NO CPU fallback: without a CUDA device of compute capability 10.x sb_init fails with
 *     SB_ERR_NO_DEVICE and nothing else can be called.
 *   - Pointers named `d_*` are device pointers obtained from sb_dev_alloc; all others are host
 *     pointers (pageable or pinned).  Calls are synchronous with respect to the host unless the name
 *     ends in `_async`; work is issued on the context's stream.
 *   - One context per GPU per thread; a context is not re-entrant.
 */
#ifndef STARK_B200_H
#define STARK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_NO_DEVICE (-1)   /* no sm_100 device / CUDA runtime failure at init            */
#define SB_ERR_CUDA (-2)        /* a CUDA call or kernel failed                                */
#define SB_ERR_ARG (-3)         /* size / alignment / range violation (reference: assert!)     */
#define SB_ERR_ROOT (-4)        /* root_of_unity is not a primitive 2^log_n-th root of unity   */
#define SB_ERR_OOM (-5)         /* device allocation failed                                    */
#define SB_ERR_VERIFY (-6)      /* a proof was rejected (reference verifier: assert / unwrap panic) */

typedef struct sb_ctx sb_ctx;
typedef struct sb_tree sb_tree;
typedef struct sb_fri_proof sb_fri_proof;

/* ---- context ------------------------------------------------------------------------------ */
int sb_init(int device, sb_ctx **out);
/* One context over n_devices GPUs of the node (1, 2, 4 or 8; one process, peer access over NVLink / NVSwitch; fails with
 * SB_ERR_NO_DEVICE when a pair has no peer access).  devices[0] is the primary device: every single-device entry point
 * runs there.  sb_prove_r1cs / sb_prove_files and the sb_ext_* chain spread ONE job over all devices: the
 * r1cs-stark binary proves on N GPUs with `r1cs-stark <r1cs> <wtns> <proof.json> --gpus N`.  An ordinal may be listed more
 * than once (logical devices on one GPU: the sharded code path without the speed-up; how it is tested on a 1-GPU box). */
int sb_init_multi(const int *devices, int n_devices, sb_ctx **out);
int sb_device_count(const sb_ctx *ctx);
void sb_destroy(sb_ctx *ctx);
const char *sb_last_error(const sb_ctx *ctx);
/* Use an external CUDA stream (cudaStream_t passed as void*, e.g. torch's current stream). */
int sb_set_stream(sb_ctx *ctx, void *cuda_stream);
int sb_sync(sb_ctx *ctx);
/* CUDA-event stopwatch on the context's stream (bench harness). */
int sb_timer_start(sb_ctx *ctx);
int sb_timer_stop(sb_ctx *ctx, float *ms);
/* Domains beyond the reference's limits (BASELINE.json configs[4], N = 2^26): the reference sampler asserts
 * modulus < 2^24 (fri/src/utils.rs:88), so prove_low_degree cannot take more than 2^26 values with columns of 2^24.
 * enable = 1 applies the same rule to larger moduli (exact in u32 while modulus * 7 < 2^32); default 0 = the
 * reference's behaviour (SB_ERR_ARG where it would panic).  No parity target exists in this range. */
int sb_set_extended_domain(sb_ctx *ctx, int enable);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t sb_launch_count(const sb_ctx *ctx);

/* Per-kernel-family timing: when enabled every launch is bracketed by CUDA events on the context's stream
 * (bench.py's roofline figures come from this, measured inside the timed region).  sb_profile resets the
 * counters; sb_profile_read synchronises and returns launches and summed milliseconds of one family. */
#define SB_KIND_NTT_PASS 0      /* ntt_pass_kernel<B>                                   */
#define SB_KIND_MERKLE_LEAVES 1 /* merkle_leaves_{cols,bytes}_kernel (leaf hash + 3 levels) */
#define SB_KIND_MERKLE_NODES 2  /* merkle_nodes_kernel (3 levels per launch)             */
#define SB_KIND_FRI_FOLD 3      /* fri_fold_kernel                                       */
#define SB_KIND_OPEN 4          /* opening gathers                                       */
#define SB_KIND_OTHER 5         /* table building, batch inverse, codecs                 */
#define SB_KIND_COUNT 6
int sb_profile(sb_ctx *ctx, int enable);
int sb_profile_read(sb_ctx *ctx, int kind, uint64_t *launches, double *total_ms);

/* Measured issue-rate ceiling of the integer pipe the field arithmetic runs on (bench.py's integer roofline): a
 * register-only probe kernel timed with CUDA events.  which = 0: Montgomery products / s, 1: IMAD.WIDE.U32 / s. */
int sb_pipe_peak(sb_ctx *ctx, int which, double *ops_per_s);

/* ---- device memory (pipelines that keep vectors resident in HBM) --------------------------- */
int sb_dev_alloc(sb_ctx *ctx, size_t bytes, void **d_ptr);
int sb_dev_free(sb_ctx *ctx, void *d_ptr);
int sb_h2d(sb_ctx *ctx, void *d_dst, const void *src, size_t bytes);
int sb_d2h(sb_ctx *ctx, void *dst, const void *d_src, size_t bytes);
int sb_host_alloc_pinned(sb_ctx *ctx, size_t bytes, void **ptr);
int sb_host_free_pinned(sb_ctx *ctx, void *ptr);

/* ---- NTT (fri/src/fft.rs) -------------------------------------------------------------------- */
/* best_fft (inverse = 0) / inv_best_fft (inverse = 1): `vals` has room for 2^log_n elements, the
 * first len_in are the caller's vector, the rest is zero padding supplied here (fft.rs:335-338).
 * In place, natural order in and out: out[k] = sum_j v[j] root^(jk) (inverse: root^-1, then / n).
 * len_in > 2^log_n -> SB_ERR_ARG (reference: assert at fft.rs:162). */
int sb_ntt(sb_ctx *ctx, uint64_t *vals, size_t len_in, const uint64_t root[4], uint32_t log_n, int inverse);
/* Same on device-resident data, batched: n_polys vectors, src element stride src_stride, dst stride
 * dst_stride (elements), d_dst must not alias d_src. */
int sb_ntt_dev(sb_ctx *ctx, const uint64_t *d_src, size_t len_in, size_t src_stride, uint64_t *d_dst,
               size_t dst_stride, size_t n_polys, const uint64_t root[4], uint32_t log_n, int inverse);
/* ONE best_fft / inv_best_fft (fft.rs:327-379) spread over the g = 2, 4 or 8 devices of an sb_init_multi context
 * (SURVEY.md 8e(5)): device d holds elements [d n/g, (d+1) n/g) of the vector in d_slabs[d] (natural order, memory of that
 * device, e.g. from sb_dev_alloc_on) on entry and the same range of the result on return; 2^20 <= n <= 2^28.  The exchanges of
 * the four-step formulation happen inside the pass kernels' loads and stores over peer memory (no transposes, no staging
 * copies).  sb_ntt on host vectors takes this path by itself on a multi-device context from 2^20 points on. */
int sb_ntt_multi_dev(sb_ctx *ctx, uint64_t *const *d_slabs, const uint64_t root[4], uint32_t log_n, int inverse);
/* device memory on device dev_index of the context (0 .. sb_device_count - 1); free with sb_dev_free */
int sb_dev_alloc_on(sb_ctx *ctx, int dev_index, size_t bytes, void **d_ptr);
/* Twiddle step of a 2^log_n-point transform split four-step style over several GPUs (stark_pure_rust_b200/sharded.py,
 * SURVEY.md 8e(5)): d_vals[r * cols + c] *= root^((row0 + r) * c) (inverse: root^-1), canonical result. */
int sb_twiddle_mul_dev(sb_ctx *ctx, uint64_t *d_vals, size_t rows, size_t cols, size_t row0, const uint64_t root[4],
                       uint32_t log_n, int inverse);
/* Low-degree extension of n_cols columns: best_fft(inv_best_fft(col, root_big^(2^log_ext), log_s),
 * root_big, log_s + log_ext) for every column (prove.rs:100-124).  cols: n_cols x col_len elements
 * (col_len <= 2^log_s, zero padded like inv_best_fft does); out: n_cols x 2^(log_s+log_ext). */
int sb_lde_batch(sb_ctx *ctx, const uint64_t *cols, size_t n_cols, size_t col_len, const uint64_t root_big[4],
                 uint32_t log_s, uint32_t log_ext, uint64_t *out);
int sb_lde_batch_dev(sb_ctx *ctx, const uint64_t *d_cols, size_t n_cols, size_t col_len, size_t col_stride,
                     const uint64_t root_big[4], uint32_t log_s, uint32_t log_ext, uint64_t *d_out);
/* expand_root_of_unity (fft.rs:5-14): out[i] = root^i for i < n (the reference stops when the power
 * wraps to 1, i.e. n = order of root). */
int sb_powers(sb_ctx *ctx, const uint64_t root[4], size_t n, uint64_t *out);
int sb_powers_dev(sb_ctx *ctx, const uint64_t root[4], size_t n, uint64_t *d_out);
/* multi_inv (poly_utils.rs:38-70): element-wise inverse in place, 0 -> 0. */
int sb_batch_inverse(sb_ctx *ctx, uint64_t *vals, size_t n);
int sb_batch_inverse_dev(sb_ctx *ctx, uint64_t *d_vals, size_t n);

/* ---- Merkle (commitment/src/merkle_proof_in_place.rs, merkle_tree.rs) ------------------------- */
/* update(leaves) + gen_proofs(&[]) + get_root(): n leaves (power of two) of leaf_bytes each, flattened.
 * All levels stay resident in HBM inside *tree. */
int sb_merkle_commit(sb_ctx *ctx, const void *leaves, size_t leaf_bytes, size_t n, uint8_t root[32], sb_tree **tree);
/* Leaves taken from device-resident field columns: leaf i = to_bytes_le(col_0[i]) || ... (prove.rs:235-258,
 * :324-327; fri.rs:120-123).  The columns must outlive the tree (openings re-read them). */
int sb_merkle_commit_cols_dev(sb_ctx *ctx, const uint64_t *const *d_cols, size_t n_cols, size_t n, uint8_t root[32],
                              sb_tree **tree);
/* gen_proofs(idx): caller order, duplicates allowed (merkle_proof_in_place.rs:199-205).
 * leaves_out: n_idx x leaf_bytes; nodes_out: n_idx x log2(n) x 32, sibling at the leaf level first, root
 * excluded (merkle_tree.rs:25-43).  Either output may be NULL. */
int sb_merkle_open(sb_ctx *ctx, const sb_tree *tree, const size_t *idx, size_t n_idx, uint8_t *leaves_out,
                   uint8_t *nodes_out);
size_t sb_tree_width(const sb_tree *tree);      /* MerkleTree::width */
size_t sb_tree_leaf_bytes(const sb_tree *tree);
int sb_tree_root(const sb_tree *tree, uint8_t root[32]);
void sb_tree_free(sb_ctx *ctx, sb_tree *tree);

/* ---- FRI (fri/src/fri.rs:46-224) --------------------------------------------------------------- */
/* prove_low_degree(values, root_of_unity, max_deg_plus_1, exclude_multiples_of). n = number of values =
 * order of root (power of two). */
int sb_fri_prove(sb_ctx *ctx, const uint64_t *vals, size_t n, const uint64_t root[4], size_t max_deg_plus_1,
                 uint32_t exclude_multiples_of, sb_fri_proof **out);
/* device-resident values; `values_tree` may be the already committed tree of d_vals (the prover's l_tree,
 * prove.rs:324-332) or NULL. */
int sb_fri_prove_dev(sb_ctx *ctx, const uint64_t *d_vals, size_t n, const uint64_t root[4], size_t max_deg_plus_1,
                     uint32_t exclude_multiples_of, const sb_tree *values_tree, sb_fri_proof **out);
/* One fold step (fri.rs:135-164): d_col[i] (n/4 elements) = the degree-<4 interpolant through
 * (x * iota^j, d_vals[i + j n/4]), j < 4, evaluated at special_x = int_LE(values_root) mod p.  Used by the sharded
 * prover, whose first values tree is spread over several GPUs (stark_pure_rust_b200/sharded.py). */
int sb_fri_fold_dev(sb_ctx *ctx, const uint64_t *d_vals, size_t n, const uint64_t root[4], const uint8_t values_root[32],
                    uint64_t *d_col);
/* Proof accessors.  Layers 0..n_layers-2 are FriProof::Middle, the last one is FriProof::Last (fri.rs:16-26). */
size_t sb_fri_n_layers(const sb_fri_proof *p);
int sb_fri_layer_is_last(const sb_fri_proof *p, size_t layer);
/* Middle: root2 (32 B); column_branches: 40 openings of the column tree (leaf 32 B, depth_col siblings);
 * poly_branches: 160 openings of the values tree (leaf 32 B, depth_poly siblings). */
int sb_fri_middle(const sb_fri_proof *p, size_t layer, const uint8_t **root2, size_t *n_column, size_t *depth_column,
                  const uint8_t **column_leaves, const uint8_t **column_nodes, size_t *n_poly, size_t *depth_poly,
                  const uint8_t **poly_leaves, const uint8_t **poly_nodes);
/* Last: n_last elements, 32 bytes each (to_bytes_le). */
int sb_fri_last(const sb_fri_proof *p, size_t layer, const uint8_t **values, size_t *n_last);
/* Roots of the per-layer value trees (layer 0 = commitment of the input values): test / bench taps. */
int sb_fri_layer_root(const sb_fri_proof *p, size_t layer, uint8_t root[32]);
/* serde_json::to_string(&Vec<FriProof>) (compact; fri.rs:16-26 + merkle_tree.rs:14-18): returns a malloc'd
 * NUL-terminated string the caller frees with sb_free_string. */
char *sb_fri_proof_json(const sb_fri_proof *p);
void sb_free_string(char *s);
void sb_fri_proof_free(sb_fri_proof *p);

/* ---- the LDE -> commit -> FRI chain with the extended columns kept on the device(s) (prove.rs:100-124, :235-264, :324-332,
 * :367), the building blocks of sb_prove_r1cs.  Columns are stored coset-major: the value at position 8 k + r of the
 * extended domain lives in coset array r at index k, and a context of g devices gives device d the cosets
 * [8 d / g, 8 (d + 1) / g) of every column.  All shifts of the prover are whole steps (multiples of 8), so pointwise work,
 * leaf hashing and the first FRI fold stay on the device that holds the data; S-point coefficient vectors (once per
 * column) and 32-byte digests are all that crosses NVLink.  Results (roots, openings, proofs, sb_ext_read) are identical to
 * the natural-order single-GPU entry points above.  log_ext must be 3 (EXTENSION_FACTOR = 8, utils.rs:134). */
typedef struct sb_ext sb_ext;
int sb_ext_create(sb_ctx *ctx, size_t n_cols, uint32_t log_s, uint32_t log_ext, sb_ext **out);
void sb_ext_free(sb_ctx *ctx, sb_ext *e);
int sb_ext_devices(const sb_ext *e);       /* devices the columns are spread over (1 for domains below 2^13) */
/* host columns [first, first + count) (count x col_len elements, col_len <= 2^log_s, zero padded like inv_best_fft) ->
 * the device that runs the column's inverse transform (column c: device c mod g) */
int sb_ext_load(sb_ctx *ctx, sb_ext *e, size_t first, size_t count, const uint64_t *cols, size_t col_len);
/* best_fft(inv_best_fft(col, root^8, log_s), root, log_s + 3) for the loaded columns [first, first + count) */
int sb_ext_extend(sb_ctx *ctx, sb_ext *e, size_t first, size_t count);
/* MerkleProofInPlace over the rows of the given columns (leaf i = to_bytes_le(col_ids[0][i]) || ...); per-device subtrees,
 * top finished on the host.  sb_merkle_open / sb_tree_* / sb_tree_free work on the returned tree. */
int sb_ext_commit(sb_ctx *ctx, const sb_ext *e, const size_t *col_ids, size_t n_ids, uint8_t root[32], sb_tree **tree);
/* prove_low_degree(column, root, max_deg_plus_1, exclude_multiples_of); values_tree = sb_ext_commit of that column or NULL */
int sb_ext_fri_prove(sb_ctx *ctx, const sb_ext *e, size_t col, const sb_tree *values_tree, size_t max_deg_plus_1,
                     uint32_t exclude_multiples_of, sb_fri_proof **out);
/* one extended column in natural order (2^(log_s+3) elements) to the host */
int sb_ext_read(sb_ctx *ctx, const sb_ext *e, size_t col, uint64_t *out);

/* ---- the caller of the hot path: mk_r1cs_proof (r1cs-stark/src/prove.rs:14-378), device resident -------- */
/* Same arguments as the reference function (prove.rs:14-26); n_constraints / n_wires only feed an assert there.
 * All field vectors are Montgomery limbs; traces / coefficients / flags have original_steps elements. */
typedef struct {
    size_t original_steps;              /* coefficients.len(), a multiple of 3 (prove.rs:30-35) */
    const uint64_t *witness_trace, *computational_trace, *coefficients, *flag0, *flag1, *flag2;
    const size_t *permuted_indices;     /* original_steps entries */
    size_t n_public;
    const uint64_t *public_wires;       /* n_public elements */
    size_t n_pfi;                       /* public_first_indices: (k, w) pairs -> (public_wires[k], xs[8 w]) */
    const size_t *pfi_k, *pfi_w;
} sb_trace;
typedef struct sb_stark_proof sb_stark_proof;
/* Returns SB_ERR_ARG where the reference panics (witness not satisfying the circuit: utils.rs:379-418, :489, :514;
 * precision beyond the sampler's 2^24 or the field's two-adicity). */
int sb_prove_r1cs(sb_ctx *ctx, const sb_trace *trace, sb_stark_proof **out);
int sb_stark_proof_roots(const sb_stark_proof *p, uint8_t m_root[32], uint8_t l_root[32], uint8_t a_root[32]);
/* Stage times of that proof (host clock, every device waited for at the stage boundaries): [0] inputs, the nine LDEs, the
 * accumulator chain and the pointwise stage (they overlap), [1] m_tree commit, [2] FRI, [3] l, l_tree and the openings,
 * [4] total (ms). */
int sb_stark_proof_stage_ms(const sb_stark_proof *p, double ms[5]);
/* serde_json::to_string(&StarkProof) (utils.rs:122-130, run.rs:549): malloc'd, free with sb_free_string. */
char *sb_stark_proof_json(const sb_stark_proof *p, size_t *len);
void sb_stark_proof_free(sb_stark_proof *p);

/* verify_r1cs_proof (r1cs-stark/src/verify.rs:13-258; fri/src/fri.rs:226-404; merkle_tree.rs:25-58).  Uses the public
 * members of `trace` (original_steps, coefficients, flag0..2, permuted_indices, public wires and their first uses);
 * witness_trace / computational_trace may be NULL.  SB_OK = accepted, SB_ERR_VERIFY = rejected (sb_last_error names the
 * failed check), other codes = malformed input.  The six interpolated columns the reference evaluates with eval_poly_at are
 * extended on the device instead (same field elements). */
int sb_verify_r1cs(sb_ctx *ctx, const sb_trace *trace, const sb_stark_proof *proof);
/* verify_low_degree_proof (fri/src/fri.rs:226-404) on the serde_json text of Vec<FriProof> (what sb_fri_proof_json writes):
 * merkle_root = root of the tree over the n values, root_of_unity of order n.  Host only; ctx may be NULL. */
int sb_fri_verify_json(sb_ctx *ctx, const char *text, size_t len, const uint8_t merkle_root[32], const uint64_t root_of_unity[4],
                       size_t n, size_t max_deg_plus_1, uint32_t exclude_multiples_of);
/* serde_json::from_reader::<StarkProof> (run.rs:578): parses the text sb_stark_proof_json writes. */
int sb_stark_proof_from_json(const char *text, size_t len, sb_stark_proof **out);
/* verify_with_file_path (run.rs:556-590): r1cs + witness (for the public wires) + proof.json.  verify_ms (may be NULL):
 * [0] host front end + JSON parse, [1] sb_verify_r1cs wall clock. */
int sb_verify_files(sb_ctx *ctx, const char *r1cs_path, const char *wtns_path, const char *proof_path, double verify_ms[2]);

/* prove_with_file_path (r1cs-stark/src/run.rs:528-554): parse <r1cs> (circom2bellman_core/src/reader.rs:4-89) and
 * <wtns> (r1cs-stark/src/reader.rs:7-42), arrange the traces (run.rs:109-308, :390-419), prove on the device and
 * write the proof as compact JSON.  proof_path may be NULL.  stage_ms (may be NULL): [0]..[3] as sb_stark_proof_stage_ms,
 * [4] sb_prove_r1cs wall clock [5] host front end [6] JSON + file write. */
int sb_prove_files(sb_ctx *ctx, const char *r1cs_path, const char *wtns_path, const char *proof_path, double stage_ms[7]);

/* The host front end alone (run.rs:109-308, :390-419; circom2bellman_core/src/reader.rs:4-89; r1cs-stark/src/reader.rs:7-42):
 * parses the two files and builds the arguments of mk_r1cs_proof in host memory.  Needs no device and no context.  *view
 * points into *out and stays valid until sb_host_trace_free. */
typedef struct sb_host_trace sb_host_trace;
int sb_trace_from_files(const char *r1cs_path, const char *wtns_path, sb_host_trace **out, const sb_trace **view);
void sb_host_trace_free(sb_host_trace *h);

/* ---- unit-test hook for the device field library (no reference counterpart: ff_derive's arithmetic is
 * generated code) ---- element-wise op on n raw 256-bit values, no range checks.
 * op: 0 Montgomery product (lazy, < 2p), 1 add, 2 sub, 3 a+2p-b, 4 canonicalise, 5 halve, 6 from Montgomery,
 *     7 to Montgomery, 8 inverse, 9 reduce [0,4p)->[0,2p), 10 canonical Montgomery product. */
int sb_fp_vec_op(sb_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);

/* ---- Fiat-Shamir helpers that sit between the kernels (host side, exact) ------------------------- */
/* get_pseudorandom_indices (fri/src/utils.rs:82-109) */
int sb_pseudorandom_indices(const uint8_t *seed, size_t seed_len, uint32_t modulus, size_t count,
                            uint32_t exclude_multiples_of, uint32_t *out);
/* same, honouring the context's sb_set_extended_domain setting (moduli >= 2^24) */
int sb_pseudorandom_indices_ctx(const sb_ctx *ctx, const uint8_t *seed, size_t seed_len, uint32_t modulus, size_t count,
                                uint32_t exclude_multiples_of, uint32_t *out);
/* blake (fri/src/utils.rs:5-10) */
void sb_blake2s(const uint8_t *msg, size_t len, uint8_t out[32]);

/* ---- the alternative digest -------------------------------------------------------------------
 * `PoseidonDigest` (commitment/src/poseidon.rs:10-63): neptune 5.1.0 Poseidon, arity 2, Strength::Standard,
 * HashMode::Correct over the BLS12-381 scalar field.  A message is 1..64 bytes, zero-padded to 32-byte chunks, every chunk a
 * canonical little-endian scalar; anything else makes the reference panic (:33, :48) and returns SB_ERR_ARG here. */
/* PoseidonDigest::hash of n messages of msg_bytes each, on the device; out = n x 32 bytes */
int sb_poseidon_hash(sb_ctx *ctx, const void *msgs, size_t msg_bytes, size_t n, uint8_t *out);
/* the same digest of one message on the host (Proof::validate with H = PoseidonDigest, merkle_tree.rs:25-43) */
int sb_poseidon_hash_host(const uint8_t *msg, size_t len, uint8_t out[32]);
/* ParallelMerkleTree<Vec<u8>, PoseidonDigest>::update + get_root (commitment/src/pallarel_merkle_tree.rs:219-253): the tree of
 * sb_merkle_commit with the other digest; sb_merkle_open / sb_tree_* / sb_tree_free apply to it unchanged */
int sb_merkle_commit_poseidon(sb_ctx *ctx, const void *leaves, size_t leaf_bytes, size_t n, uint8_t root[32], sb_tree **tree);

#ifdef __cplusplus
}
#endif
#endif
