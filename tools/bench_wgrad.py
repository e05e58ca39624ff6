"""Times single wgrad_gemm launches of d_net's large layers on the 4B batch (CUDA events, L2 flushed between repetitions).
Development aid:  python tools/bench_wgrad.py [--reps 9]      (T2I_WGRAD_RANGE_MAJOR=0: the split-fastest work order)
DRAM bytes:  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:wgrad_gemm python tools/bench_wgrad.py --reps 1 --warm 0"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2i_b200 import kernels as K  # noqa: E402

SHAPES = {   # name: (mode, k, N, H, W, Cin, Cout)
    "dh1_4B": (K.CONV_K4S2, 4, 1024, 32, 32, 128, 256),
    "dh2_4B": (K.CONV_K4S2, 4, 1024, 16, 16, 256, 512),
    "dh3_4B": (K.CONV_K4S2, 4, 1024, 8, 8, 512, 1024),
    "dh5_4B": (K.CONV_S1, 3, 1024, 4, 4, 1152, 1024),
    "dr3_4B": (K.CONV_S1, 3, 1024, 4, 4, 512, 1024),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=9)
    ap.add_argument("--warm", type=int, default=2)
    a = ap.parse_args()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, (mode, k, N, H, W, ci, co) in SHAPES.items():
        oh, ow = (H // 2, W // 2) if mode == K.CONV_K4S2 else (H, W)
        taps = 16 if mode == K.CONV_K4S2 else k * k
        x = torch.randn(1, N, H, W, ci, device=dev).bfloat16()
        dy = torch.randn(1, N, oh, ow, co, device=dev).bfloat16()
        dw = torch.zeros(taps, co, ci, device=dev)
        ts = []
        for i in range(a.reps + a.warm):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.wgrad_gemm(mode, k, K.View(x), K.View(dy), dw)
            e1.record()
            torch.cuda.synchronize()
            if i >= a.warm:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        flops = 2.0 * N * oh * ow * co * ci * taps
        mb = (x.numel() + dy.numel()) * 2 / 1e6 + dw.numel() * 4 / 1e6
        print("%-8s median %.4f ms  %.0f TFLOP/s  (tensors %.0f MB)" % (name, med, flops / med / 1e9, mb))


if __name__ == "__main__":
    main()
