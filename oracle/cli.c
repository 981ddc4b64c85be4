/*
 * cli.c — `r1cs_stark_oracle <r1cs> <wtns> <proof.json> [cpus]`: the CPU restatement of the
 * reference binary (r1cs-stark/src/main.rs:4-11 -> run.rs:590-625): prove, write proof.json,
 * verify.  ORACLE / TEST INFRASTRUCTURE — see oracle.h.
 */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s <r1cs> <wtns> <proof.json> [cpus]\n", argv[0]);
        return 2;
    }
    unsigned cpus = argc > 4 ? (unsigned)atoi(argv[4]) : (unsigned)sysconf(_SC_NPROCESSORS_ONLN);
    double t = 0;
    int rc = orc_prove_files(argv[1], argv[2], argv[3], cpus, 1, &t);
    fprintf(stderr, "oracle: rc=%d prove=%.3fs (ntt %.3f, merkle %.3f, fri %.3f, rest %.3f) cpus=%u\n", rc, t,
            orc_stage_s[0], orc_stage_s[1], orc_stage_s[2], orc_stage_s[3], cpus);
    return rc;
}
