#!/usr/bin/env python
"""Hand-over timeline of one conv_gemm launch (CTA 0, first tiles).  Needs a DEBUG build of the library
(make -C text-to-image_b200/csrc clean all EXTRA=-DT2I_TIMELINE_BUILD; the stamps slow the kernel by 10-25 %, the
default build compiles them out):  T2I_TIMELINE=1 python tools/conv_timeline.py [case]
Prints, per tile, the time (ns, relative to the first stamp) of: weights TMA issued, patch rows arrived, MMA saw the
stage, MMA committed, epilogue saw the accumulator, epilogue done, MMA got a free accumulator."""
import ctypes as C
import os
import sys

import torch

os.environ.setdefault("T2I_TIMELINE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2i_b200 import _lib, kernels as K  # noqa: E402


CHAIN = ("floor", "c0", "c5")


def chain(fn, n=8):
    """n back-to-back launches replayed from a CUDA graph: kernel-level stamps of CTA 0 of every launch (row 63 of its
    region) and the hand-over stamps of its first tile -- where the fixed cost of a small launch goes."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (32 * 512))()
    _lib.call("t2i_debug_timeline", buf, 32 * 512)
    # regions: 0 = the eager call, 1..n = the captured launches
    reg = lambda r, t, e: buf[r * 512 + t * 8 + e]
    t0 = reg(1, 63, 0)
    print("launch  entry  setup  pdl_ok | tile0: mma_see commit epi_see staged stored | stores_done roles_done synced dealloc   (ns)")
    for r in range(1, n + 1):
        k = [reg(r, 63, e) for e in range(8)]
        t = [reg(r, 0, e) for e in range(8)]
        f = lambda v: "%7d" % (v - t0) if v else "     -1"
        print("%4d " % r, f(k[0]), f(k[1]), f(k[2]), "|", f(t[2]), f(t[3]), f(t[4]), f(t[0]), f(t[5]), "|", f(k[6]), f(k[3]), f(k[4]), f(k[5]))


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "img"
    dev = torch.device("cuda")
    B = 1024
    if case in CHAIN:
        shp = {"floor": (256, 8, 1, 64, 128), "c0": (256, 4, 1, 1024, 256), "c5": (256, 8, 3, 128, 128)}[case]
        n, hw, k, ci, co = shp
        x = torch.randn(1, n, hw, hw, ci, device=dev).to(torch.bfloat16)
        w = (torch.randn(1, k * k, co, ci, device=dev) * 0.05).to(torch.bfloat16)
        y = torch.zeros(1, n, hw, hw, co, device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(co, device=dev)
        fn = lambda: K.conv_gemm(K.CONV_S1, k, 0, K.View(x), w, K.View(y), bias=bias)
    elif case == "img":
        img = torch.rand(B, 64, 64, 3, device=dev) * 2 - 1
        w = (torch.randn(1, 1, 128, 64, device=dev) * 0.1).to(torch.bfloat16)
        y = torch.zeros(1, B, 32, 32, 128, device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(128, device=dev)
        rows = torch.zeros(1, B, 64, K.img_row_pitch(64), device=dev, dtype=torch.bfloat16)
        K.img_to_rows(img, rows)
        fn = lambda: K.conv_gemm(K.CONV_K4S2, 4, 0, K.ImgPatches(rows, 64), w, K.View(y), bias=bias, act=K.ACT_LRELU)
    elif case == "c8":      # g_net 3x3 conv at 32x32, 128 -> 128 (N = 128, K = 1152)
        x = torch.randn(1, 256, 32, 32, 128, device=dev).to(torch.bfloat16)
        w = (torch.randn(1, 9, 128, 128, device=dev) * 0.05).to(torch.bfloat16)
        y = torch.zeros(1, 256, 32, 32, 128, device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(128, device=dev)
        fn = lambda: K.conv_gemm(K.CONV_S1, 3, 0, K.View(x), w, K.View(y), bias=bias)
    elif case in ("c8s", "t2s"):     # the same with the BatchNorm statistics in the epilogue; t2s: deconv 16x16 256 -> 128
        dec = case == "t2s"
        hw, ci, taps = (16, 256, 16) if dec else (32, 128, 9)
        x = torch.randn(1, 256, hw, hw, ci, device=dev).to(torch.bfloat16)
        w = (torch.randn(1, taps, 128, ci, device=dev) * 0.05).to(torch.bfloat16)
        y = torch.zeros(1, 256, 32, 32, 128, device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(128, device=dev)
        s1, s2 = torch.zeros(128, device=dev), torch.zeros(128, device=dev)
        fn = lambda: K.conv_gemm(K.DECONV_K4S2 if dec else K.CONV_S1, 4 if dec else 3, 0, K.View(x), w, K.View(y), bias=bias,
                                 stat_sum=s1, stat_sq=s2)
    elif case == "dh3":     # d_net 4x4/s2 conv at 8x8, 512 -> 1024 (N = 256, K = 8192)
        x = torch.randn(1, 256, 8, 8, 512, device=dev).to(torch.bfloat16)
        w = (torch.randn(1, 16, 1024, 512, device=dev) * 0.02).to(torch.bfloat16)
        y = torch.zeros(1, 256, 4, 4, 1024, device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(1024, device=dev)
        fn = lambda: K.conv_gemm(K.CONV_K4S2, 4, 0, K.View(x), w, K.View(y), bias=bias, act=K.ACT_LRELU)
    else:   # the same output through the plain kernel: K = 64 patch matrix
        x = torch.randn(1, B * 1024, 64, device=dev).to(torch.bfloat16)
        w = (torch.randn(1, 1, 128, 64, device=dev) * 0.1).to(torch.bfloat16)
        y = torch.zeros(1, B * 1024, 128, device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(128, device=dev)
        fn = lambda: K.conv_gemm(K.CONV_S1, 1, 0, K.View(x), w, K.View(y), bias=bias, act=K.ACT_LRELU)
    if case in CHAIN:
        return chain(fn)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (3 * 512))()
    _lib.call("t2i_debug_timeline", buf, 3 * 512)       # one region per launch: the third one
    rows = [[buf[1024 + t * 8 + e] for e in range(8)] for t in range(64)]
    t0 = min(v for r in rows for v in r if v)
    print("tile   tmaB|s0fence  patch|s0bar2  mma_see mma_commit epi_see epi_done acc_free epi_top  (ns since first stamp)")
    for t, r in enumerate(rows[:40]):
        print("%4d " % t + " ".join("%8d" % (v - t0 if v else -1) for v in r))


if __name__ == "__main__":
    main()
