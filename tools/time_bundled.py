import sys, time
sys.path.insert(0, "/root/repo")
import stark_pure_rust_b200 as sb
ctx = sb.default_context()
d = "/root/repo/tests/golden/circuits/"
for name in ("bits", "pedersen_test"):
    for _ in range(3):
        t = time.perf_counter(); ms = sb.prove.prove_with_file_path(d + name + ".r1cs", d + name + ".wtns", "/tmp/p.json", ctx=ctx); w = time.perf_counter() - t
        t = time.perf_counter(); sb.prove.verify_with_file_path(d + name + ".r1cs", d + name + ".wtns", "/tmp/p.json", ctx=ctx); v = time.perf_counter() - t
    print(name, "prove %.1f ms" % (w * 1e3), [round(x, 2) for x in ms], "verify %.1f ms" % (v * 1e3))
