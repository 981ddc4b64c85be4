#!/usr/bin/env python3
"""Seeded synthetic circuit generator: writes an iden3 v1 `.r1cs` and a matching `.wtns` (formats:
SURVEY.md A.7; circom2bellman_core/src/reader.rs:4-89, r1cs-stark/src/reader.rs:7-42) whose witness
satisfies every constraint by construction.  Stand-in for tests/sha256_2_test.r1cs, which is missing from
the reference mount (.MISSING_LARGE_BLOBS): same wire count order of magnitude, row-length distribution
chosen so that the trace lands on the requested number of steps.

    python tools/gen_r1cs.py out_prefix --constraints 30000 --avg-terms 5.8 --pub 2 --seed 1

Constraint c:  (sum a_i w_i) * (sum b_i w_i) = w_new,  w_new a fresh wire set to the product; a few
constraints are linear (B = 1) or have multi-term C to exercise every padding case of run.rs:109-281.
"""
import argparse
import random
import struct

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def generate(n_constraints, avg_terms, n_pub, seed):
    rnd = random.Random(seed)
    wit = [1] + [rnd.randrange(P) for _ in range(n_pub)] + [rnd.randrange(1, 1 << 64) for _ in range(3)]
    cons = []

    def lin(max_terms):
        k = max(1, min(len(wit), int(rnd.expovariate(1.0 / max(avg_terms - 1, 0.5))) + 1))
        k = min(k, max_terms)
        wires = rnd.sample(range(len(wit)), k) if k <= len(wit) else list(range(len(wit)))
        terms = [(w, rnd.choice([1, 2, 3, P - 1, rnd.randrange(P)])) for w in sorted(wires)]
        return terms, sum(c * wit[w] for w, c in terms) % P

    for c in range(n_constraints):
        a, av = lin(40)
        kind = rnd.random()
        if kind < 0.1:
            b, bv = [(0, 1)], 1                      # linear constraint
        else:
            b, bv = lin(40)
        new = len(wit)
        wit.append(av * bv % P)
        if kind > 0.95 and len(wit) > 4:              # multi-term C: w_new + k*w_j - k*w_j
            j = rnd.randrange(1, new)
            cterm = sorted([(new, 1), (j, 0)])
            cterm = [(w, cf) for w, cf in cterm]
        else:
            cterm = [(new, 1)]
        cons.append((a, b, cterm))
    return wit, cons


def write_files(prefix, wit, cons, n_pub):
    n_wires = len(wit)
    parts = []                      # joined once: repeated bytes += is quadratic (20 MB at 30k constraints)
    for a, b, c in cons:
        for f in (a, b, c):
            parts.append(struct.pack("<I", len(f)))
            for w, cf in f:
                parts.append(struct.pack("<I", w))
                parts.append(int(cf % P).to_bytes(32, "little"))
    body = b"".join(parts)
    header = struct.pack("<I", 32) + P.to_bytes(32, "little") + struct.pack("<IIIIQI", n_wires, 0, n_pub, n_wires - 1 - n_pub, n_wires, len(cons))
    labels = b"".join(struct.pack("<Q", i) for i in range(n_wires))
    with open(prefix + ".r1cs", "wb") as f:
        f.write(b"r1cs" + struct.pack("<II", 1, 3))
        f.write(struct.pack("<IQ", 1, len(header)) + header)
        f.write(struct.pack("<IQ", 2, len(body)) + body)
        f.write(struct.pack("<IQ", 3, len(labels)) + labels)
    with open(prefix + ".wtns", "wb") as f:
        f.write(b"wtns" + struct.pack("<II", 2, 2))                 # magic, version, n_sections
        f.write(struct.pack("<IQ", 1, 40))                           # section 1 header: id + u64 size  (3 words)
        f.write(struct.pack("<I", 32) + P.to_bytes(32, "little") + struct.pack("<I", n_wires))
        f.write(struct.pack("<IQ", 2, 32 * n_wires))                 # section 2 header (3 words)
        f.write(b"".join(int(v).to_bytes(32, "little") for v in wit))
    rows = sum(max(len(a), len(b), len(c)) for a, b, c in cons)
    return {"n_wires": n_wires, "n_constraints": len(cons), "a_trace_len": rows, "original_steps": 3 * rows}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("prefix")
    ap.add_argument("--constraints", type=int, default=300)
    ap.add_argument("--avg-terms", type=float, default=3.0)
    ap.add_argument("--pub", type=int, default=2)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    wit, cons = generate(a.constraints, a.avg_terms, a.pub, a.seed)
    print(write_files(a.prefix, wit, cons, a.pub))
