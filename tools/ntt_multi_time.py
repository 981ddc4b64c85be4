"""runs bench.run_multi_records' ntt_multi record alone on logical or real devices: python tools/ntt_multi_time.py G [log_n] [--logical]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import stark_pure_rust_b200 as sb


class A:
    pass


g = int(sys.argv[1])
a = A()
a.sharded_log_n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 26
a.sharded_steps = 5
logical = "--logical" in sys.argv
ctxs = {1: sb.Context(devices=[0]), g: sb.Context(devices=[0] * g if logical else list(range(g)))}
print(json.dumps(bench.ntt_multi_record(a, ctxs, g)))
