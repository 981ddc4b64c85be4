import sys, os, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, stark_pure_rust_b200 as sb
from stark_pure_rust_b200 import field
from stark_pure_rust_b200._lib import _ptr
from conftest import random_elems
ctx = sb.Context(0); lib = ctx.lib
k = 24; n = 1 << k; cols = 10
src = ctx.to_device(random_elems(cols << k, 1).reshape(-1, 4)); dst = ctx.alloc((cols << k) * 32)
w = field.mont_scalar(field.root_of_unity(k))
def best(fn):
    fn(); t = []
    for _ in range(3):
        ctx.timer_start(); fn(); t.append(ctx.timer_stop())
    return min(t)
print("fft 2^24 x10 ms", best(lambda: ctx.check(lib.sb_ntt_dev(ctx.h, C.c_void_p(src), n, n, C.c_void_p(dst), n, cols, _ptr(w), k, 0))))
print("lde 2^21->2^24 x10 ms", best(lambda: ctx.check(lib.sb_lde_batch_dev(ctx.h, C.c_void_p(src), cols, n >> 3, n >> 3, _ptr(w), k - 3, 3, C.c_void_p(dst)))))
for kk in (16, 20, 22):
    nn = 1 << kk; ww = field.mont_scalar(field.root_of_unity(kk))
    print("fft 2^%d x10 ms" % kk, best(lambda: ctx.check(lib.sb_ntt_dev(ctx.h, C.c_void_p(src), nn, nn, C.c_void_p(dst), nn, cols, _ptr(ww), kk, 0))))
