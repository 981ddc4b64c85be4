#!/usr/bin/env python3
"""Byte-level goldens at the headline size (BASELINE.json configs[1], N = 2^24) from the CPU oracle: tests/golden/vectors_large.json.

    python tests/golden/make_golden_large.py          (about 6 minutes on 8 host cores, ~12 GB of RAM)

Inputs are seeded (tests/conftest.py::random_elems), only sha256 digests of the outputs are stored:
  ntt     best_fft / inv_best_fft (fri/src/fft.rs:327-379) of 2^22 and 2^24 points
  lde     best_fft(inv_best_fft(col, g1, 21), g2, 24) for 8 columns of 2^21 values (prove.rs:100-124)
  merkle8 MerkleProofInPlace over the 256-byte rows of the 8 extended columns (prove.rs:235-264) + 4 openings
  merkle1 tree over the 32-byte leaves of the last extended column (prove.rs:324-332)
  fri     serde_json text of prove_low_degree(last column, g2, N/4, 8) (fri.rs:46-224)
tests/test_gpu_parity.py::test_full_size_2_24_golden asserts the same digests on the GPU path."""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_bind as ob
from conftest import random_elems

LOG_S, LOG_N, NC = 21, 24, 8
OPEN = [0, (1 << LOG_N) - 1, (1 << LOG_N) // 3, 12345678]


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    ob.lib()
    t0 = time.time()
    out = {"_how": "python tests/golden/make_golden_large.py (oracle/liboracle.so)", "log_s": LOG_S, "log_n": LOG_N, "n_cols": NC}
    ntt = []
    for k in (22, 24):
        v = random_elems(1 << k, 0xB200 + k)
        w = ob.root_of_unity(k)
        ntt.append({"log_n": k, "seed": 0xB200 + k, "fwd_sha256": sha(ob.best_fft(v, w, k).tobytes()),
                    "inv_sha256": sha(ob.best_fft(v, w, k, inverse=True).tobytes())})
        print("ntt 2^%d done %.0f s" % (k, time.time() - t0), flush=True)
    out["ntt"] = ntt
    n, s = 1 << LOG_N, 1 << LOG_S
    g1, g2 = ob.root_of_unity(LOG_S), ob.root_of_unity(LOG_N)
    cols = random_elems(NC * s, 0x1DE).reshape(NC, s, 4)
    ext = [ob.best_fft(ob.best_fft(cols[c], g1, LOG_S, inverse=True), g2, LOG_N) for c in range(NC)]
    out["lde"] = {"seed": 0x1DE, "col_sha256": [sha(e.tobytes()) for e in ext]}
    print("lde done %.0f s" % (time.time() - t0), flush=True)
    rows = np.empty((n, NC, 32), dtype=np.uint8)
    for c in range(NC):
        rows[:, c, :] = ob.fp_to_bytes_le(ext[c]).reshape(n, 32)
    root, nodes = ob.merkle_gen_proofs(rows.tobytes(), 32 * NC, n, OPEN)
    out["merkle8"] = {"root": root.hex(), "open": OPEN, "nodes_sha256": sha(nodes.tobytes()),
                      "leaves_sha256": sha(b"".join(rows[i].tobytes() for i in OPEN))}
    del rows
    print("merkle8 done %.0f s" % (time.time() - t0), flush=True)
    root1, _ = ob.merkle_gen_proofs(ob.fp_to_bytes_le(ext[NC - 1]).tobytes(), 32, n, [])
    out["merkle1"] = {"root": root1.hex()}
    text, ok = ob.prove_low_degree_json(ext[NC - 1], g2, n // 4, 8, verify=True)
    assert ok
    out["fri"] = {"json_sha256": sha(text.encode()), "json_len": len(text)}
    print("fri done %.0f s" % (time.time() - t0), flush=True)
    json.dump(out, open(os.path.join(HERE, "vectors_large.json"), "w"), indent=1)
    print("wrote vectors_large.json")


if __name__ == "__main__":
    main()
