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


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "img"
    dev = torch.device("cuda")
    B = 1024
    if case == "img":
        img = torch.rand(B, 64, 64, 3, device=dev) * 2 - 1
        w = (torch.randn(1, 1, 128, 64, device=dev) * 0.1).to(torch.bfloat16)
        y = torch.zeros(1, B, 32, 32, 128, device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(128, device=dev)
        fn = lambda: K.conv_gemm(K.CONV_K4S2, 4, 0, K.ImgPatches(img), w, K.View(y), bias=bias, act=K.ACT_LRELU)
    elif case == "c8":      # g_net 3x3 conv at 32x32, 128 -> 128 (N = 128, K = 1152)
        x = torch.randn(1, 256, 32, 32, 128, device=dev).to(torch.bfloat16)
        w = (torch.randn(1, 9, 128, 128, device=dev) * 0.05).to(torch.bfloat16)
        y = torch.zeros(1, 256, 32, 32, 128, device=dev, dtype=torch.bfloat16)
        bias = torch.zeros(128, device=dev)
        fn = lambda: K.conv_gemm(K.CONV_S1, 3, 0, K.View(x), w, K.View(y), bias=bias)
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
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 512)()
    _lib.call("t2i_debug_timeline", buf, 512)
    rows = [[buf[t * 8 + e] for e in range(8)] for t in range(64)]
    t0 = min(v for r in rows for v in r if v)
    print("tile   tmaB|s0fence  patch|s0bar2  mma_see mma_commit epi_see epi_done acc_free epi_top  (ns since first stamp)")
    for t, r in enumerate(rows[:40]):
        print("%4d " % t + " ".join("%8d" % (v - t0 if v else -1) for v in r))


if __name__ == "__main__":
    main()
