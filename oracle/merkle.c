/*
 * merkle.c — the production Merkle tree of the reference, restating
 * /root/reference/packages/commitment/src/merkle_proof_in_place.rs:54-206 and the branch check
 * of commitment/src/merkle_tree.rs:25-43.
 * ORACLE / TEST INFRASTRUCTURE — see oracle.h.
 *
 * Shape: leaf digest = blake(leaf bytes); parent = blake(left32 || right32); root = top node.
 * The reference reduces the digest array in place in 2^floor(log2 cores) chunks and then across
 * the chunk roots; the node values and the sibling order do not depend on that chunking, so the
 * restatement reduces the whole array with the same in-place stride-doubling loop (:78-98).
 * Like the reference, every call hashes all leaves again (gen_proofs rebuilds the tree).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

void orc_merkle_gen_proofs(const uint8_t *leaves, size_t leaf_bytes, size_t n,
                           const size_t *indices, size_t n_idx,
                           uint8_t root[32], uint8_t *nodes_out) {
    if (n == 0 || (n & (n - 1))) abort(); /* :113 assert is_a_power_of_2 */
    uint8_t *cur = (uint8_t *)malloc(n * 32);
    for (size_t i = 0; i < n; i++) orc_blake2s(cur + 32 * i, leaves + i * leaf_bytes, leaf_bytes); /* :128-131 */
    size_t depth = 0;
    while (((size_t)1 << depth) < n) depth++;
    size_t log_interval = 0;
    while (((size_t)1 << log_interval) < n) { /* :78 */
        for (size_t i = 0; i < n_idx; i++) { /* :79-82 */
            size_t twin = ((indices[i] >> log_interval) ^ 1) << log_interval;
            memcpy(nodes_out + (i * depth + log_interval) * 32, cur + twin * 32, 32);
        }
        size_t interval = (size_t)1 << log_interval;
        for (size_t base = 0; base < n; base += 2 * interval) { /* :86-95 */
            uint8_t msg[64];
            memcpy(msg, cur + base * 32, 32);
            memcpy(msg + 32, cur + (base + interval) * 32, 32);
            orc_blake2s(cur + base * 32, msg, 64);
        }
        log_interval++;
    }
    memcpy(root, cur, 32); /* :189 */
    free(cur);
}

int orc_merkle_validate(const uint8_t root[32], size_t index, const uint8_t *leaf, size_t leaf_bytes,
                        const uint8_t *nodes, size_t depth) {
    uint8_t cur[32], msg[64];
    orc_blake2s(cur, leaf, leaf_bytes);
    size_t tmp = index;
    for (size_t d = 0; d < depth; d++) {
        if (tmp % 2 == 0) {
            memcpy(msg, cur, 32);
            memcpy(msg + 32, nodes + d * 32, 32);
        } else {
            memcpy(msg, nodes + d * 32, 32);
            memcpy(msg + 32, cur, 32);
        }
        orc_blake2s(cur, msg, 64);
        tmp /= 2;
    }
    return memcmp(cur, root, 32) == 0;
}
