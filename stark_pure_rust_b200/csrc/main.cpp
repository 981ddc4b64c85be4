// r1cs-stark <r1cs> <wtns> <proof.json> -- the reference binary's command line (r1cs-stark/src/main.rs:4-11) on the
// B200 backend: prove, write proof.json (compact serde_json layout, run.rs:549-551), then verify what was written like the
// reference's run_with_file_path does (run.rs:592-626).  `--gpus N` (or SB_GPUS=N) spreads the ONE proof over N GPUs of the
// node (sb_init_multi: 1, 2, 4 or 8; devices 0 .. N-1); a bare number as the fourth argument selects a single device.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/stark_b200.h"

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s <r1cs> <wtns> <proof.json> [device | --gpus N]\n", argv[0]);
        return 2;
    }
    sb_ctx *ctx = nullptr;
    int n_gpus = getenv("SB_GPUS") ? atoi(getenv("SB_GPUS")) : 0, device = 0;
    for (int i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--gpus") && i + 1 < argc) n_gpus = atoi(argv[++i]);
        else device = atoi(argv[i]);
    }
    int rc;
    if (n_gpus > 1) {
        int devs[8];
        for (int i = 0; i < 8; i++) devs[i] = i;
        rc = n_gpus <= 8 ? sb_init_multi(devs, n_gpus, &ctx) : SB_ERR_ARG;
    } else {
        rc = sb_init(device, &ctx);
    }
    if (rc != SB_OK) {
        fprintf(stderr, "r1cs-stark: no B200 (sm_100) device available; there is no CPU fallback (error %d)\n", rc);
        return 1;
    }
    double ms[7] = {0};
    rc = sb_prove_files(ctx, argv[1], argv[2], argv[3], ms);
    if (rc != SB_OK) {
        fprintf(stderr, "r1cs-stark: %s (error %d)\n", sb_last_error(ctx), rc);
        sb_destroy(ctx);
        return 1;
    }
    printf("Produced STARK proof: front end %.3f ms, GPU prove %.3f ms (LDE + pointwise %.3f, m_tree %.3f, FRI %.3f, l_tree + openings %.3f), JSON %.3f ms\n",
           ms[5], ms[4], ms[0], ms[1], ms[2], ms[3], ms[6]);
    // run_with_file_path (run.rs:592-626) verifies what it has just written
    double vms[2] = {0, 0};
    rc = sb_verify_files(ctx, argv[1], argv[2], argv[3], vms);
    if (rc != SB_OK) {
        fprintf(stderr, "r1cs-stark: proof rejected: %s (error %d)\n", sb_last_error(ctx), rc);
        sb_destroy(ctx);
        return 1;
    }
    printf("Done proof verification: front end + JSON parse %.3f ms, verify %.3f ms\n", vms[0], vms[1]);
    sb_destroy(ctx);
    return 0;
}
