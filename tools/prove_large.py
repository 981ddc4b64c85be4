#!/usr/bin/env python3
"""Prove a seeded synthetic circuit of sha256_2_test's scale (BASELINE.json configs[3]; the real .r1cs is missing
from the reference mount) on the B200 pipeline, print the stage times, optionally time the CPU oracle too.

    python tools/prove_large.py [--constraints 30000] [--avg-terms 8] [--cpu] [--reps 3]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_r1cs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--constraints", type=int, default=30000)
    ap.add_argument("--avg-terms", type=float, default=8.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--dir", default="/tmp")
    ap.add_argument("--devices", default=None, help="comma-separated device ordinals for ONE proof over several (logical) devices, e.g. 0,1,2,3 or 0,0")
    ap.add_argument("--profile", action="store_true", help="print the per-kernel-family times of the last repetition (CUDA events around every launch)")
    ap.add_argument("--timeline", action="store_true", help="after the repetitions: one more proof under torch.profiler (CUPTI), printing every device activity longer than 0.15 ms and every gap longer than 0.3 ms")
    a = ap.parse_args()
    prefix = os.path.join(a.dir, "syn_%d_%g_%d" % (a.constraints, a.avg_terms, a.seed))
    t0 = time.perf_counter()
    wit, cons = gen_r1cs.generate(a.constraints, a.avg_terms, 2, a.seed)
    info = gen_r1cs.write_files(prefix, wit, cons, 2)
    print("generated", info, "in %.1f s" % (time.perf_counter() - t0), flush=True)
    import stark_pure_rust_b200 as sb
    ctx = sb.Context(devices=[int(x) for x in a.devices.split(",")]) if a.devices else sb.default_context()
    out = prefix + ".proof.json"
    res = {"circuit": info}
    for r in range(a.reps):
        if a.profile and r == a.reps - 1:
            ctx.profile(True)
        t0 = time.perf_counter()
        ms = sb.prove.prove_with_file_path(prefix + ".r1cs", prefix + ".wtns", out, ctx=ctx)
        wall = (time.perf_counter() - t0) * 1e3
        if a.profile and r == a.reps - 1:
            kinds = ["ntt_pass", "merkle_leaves", "merkle_nodes", "fri_fold", "open", "other"]
            print("kernel families (launches, ms summed over devices):", {k: ctx.profile_read(i) for i, k in enumerate(kinds)})
            ctx.profile(False)
        import hashlib
        print("sha256", hashlib.sha256(open(out, "rb").read()).hexdigest())
        print("gpu rep %d: wall %.1f ms | LDE+pointwise %.2f m_tree %.2f FRI %.2f l_tree+openings %.2f | prove %.2f | front end %.2f | json+write %.2f | proof %d bytes" % (
            r, wall, ms[0], ms[1], ms[2], ms[3], ms[4], ms[5], ms[6], os.path.getsize(out)), flush=True)
        res["gpu"] = {"wall_ms": wall, "stage_ms": ms}
    if a.timeline:
        import torch
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            sb.prove.prove_with_file_path(prefix + ".r1cs", prefix + ".wtns", out, ctx=ctx)
            torch.cuda.synchronize()
        ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        t0, last_end = ev[0].time_range.start, ev[0].time_range.start
        for e in ev:
            st, en = e.time_range.start, e.time_range.end
            if st - last_end > 300:
                print("  %9.3f ms  ---- idle %.3f ms ----" % ((last_end - t0) / 1e3, (st - last_end) / 1e3))
            if en - st > 150:
                print("  %9.3f ms  %8.3f ms  %s" % ((st - t0) / 1e3, (en - st) / 1e3, e.name[:90]))
            last_end = max(last_end, en)
    if a.cpu:
        import oracle_bind as ob
        want = prefix + ".oracle.json"
        t0 = time.perf_counter()
        rc, _ = ob.prove_files(prefix + ".r1cs", prefix + ".wtns", want, verify=False)
        cpu_s = time.perf_counter() - t0
        same = ob.sha256_file(want) == ob.sha256_file(out)
        print("cpu oracle: rc %d, %.2f s on %d threads; proof.json identical: %s" % (rc, cpu_s, os.cpu_count(), same), flush=True)
        res["cpu"] = {"s": cpu_s, "identical": same}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
