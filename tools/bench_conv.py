"""Times single conv_gemm launches (CUDA events, L2 flushed between repetitions) for chosen layer shapes, with
and without the epilogue statistics.  Development aid:  python tools/bench_conv.py [shape-name ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2i_b200 import kernels as K  # noqa: E402

SHAPES = {
    # name: (mode, k, flip, N, H, W, Cin, Cout, kn)
    "dh1_dgrad": (K.DECONV_K4S2, 4, 0, 1024, 16, 16, 256, 128, True),
    "dh2_dgrad": (K.DECONV_K4S2, 4, 0, 1024, 8, 8, 512, 256, True),
    "c8_fwd": (K.CONV_S1, 3, 0, 256, 32, 32, 128, 128, False),
    "t2_dgrad": (K.CONV_K4S2, 4, 0, 256, 32, 32, 128, 256, True),
    "t3_dgrad": (K.CONV_S1, 1, 0, 262144, 1, 1, 64, 128, True),
    "c7_fwd": (K.CONV_S1, 3, 0, 256, 16, 16, 256, 256, False),
    # small generator layers at batch 256 (launch / latency bound) and a one-k-block launch (fixed cost)
    "c1_fwd": (K.CONV_S1, 3, 0, 256, 4, 4, 256, 256, False),
    "c2_fwd": (K.CONV_S1, 3, 0, 256, 4, 4, 256, 1024, False),
    "c5_fwd": (K.CONV_S1, 3, 0, 256, 8, 8, 128, 128, False),
    "c3_fwd": (K.CONV_S1, 3, 0, 256, 8, 8, 512, 512, False),
    "t0_fwd": (K.DECONV_K4S2, 4, 0, 256, 4, 4, 1024, 512, False),
    "floor": (K.CONV_S1, 1, 0, 256, 8, 8, 64, 128, False),
    # the remaining distinct batch-256 shapes of the step (generator + the B-sample discriminator passes)
    "c0_fwd": (K.CONV_S1, 1, 0, 256, 4, 4, 1024, 256, False),
    "c4_fwd": (K.CONV_S1, 1, 0, 256, 8, 8, 512, 128, False),
    "c6_fwd": (K.CONV_S1, 3, 0, 256, 8, 8, 128, 512, False),
    "t1_fwd": (K.DECONV_K4S2, 4, 0, 256, 8, 8, 512, 256, False),
    "t2_fwd": (K.DECONV_K4S2, 4, 0, 256, 16, 16, 256, 128, False),
    "dh1_fwd": (K.CONV_K4S2, 4, 0, 256, 32, 32, 128, 256, False),
    "dh2_fwd": (K.CONV_K4S2, 4, 0, 256, 16, 16, 256, 512, False),
    "dh3_fwd": (K.CONV_K4S2, 4, 0, 256, 8, 8, 512, 1024, False),
    "dr1_fwd": (K.CONV_S1, 1, 0, 256, 4, 4, 1024, 256, False),
    "dr2_fwd": (K.CONV_S1, 3, 0, 256, 4, 4, 256, 512, False),
    "dr3_fwd": (K.CONV_S1, 3, 0, 256, 4, 4, 512, 1024, False),
    "dh5_fwd": (K.CONV_S1, 3, 0, 256, 4, 4, 1152, 1024, False),
    "dh6_fwd": (K.CONV_S1, 1, 0, 256, 4, 4, 1024, 1024, False),
    "fc0_fwd": (K.CONV_S1, 1, 0, 256, 1, 1, 256, 16384, False),
    "fc0_dgrad": (K.CONV_S1, 1, 0, 256, 1, 1, 16384, 256, True),
}


def run(name, variant, reps=int(os.environ.get('REPS', '20')), warm=int(os.environ.get('WARM', '3'))):
    mode, k, flip, N, H, W, Ci, Co, kn = SHAPES[name]
    dev = "cuda"
    oh, ow = (H, W) if mode == K.CONV_S1 else (H // 2, W // 2) if mode == K.CONV_K4S2 else (2 * H, 2 * W)
    taps = k * k if mode == K.CONV_S1 else 16
    x = torch.randn(1, N, H, W, Ci, device=dev).bfloat16()
    w = (torch.randn(1, taps, Ci, Co, device=dev) * 0.05).bfloat16() if kn else (torch.randn(1, taps, Co, Ci, device=dev) * 0.05).bfloat16()
    y = torch.zeros(1, N, oh, ow, Co, device=dev, dtype=torch.bfloat16)
    mask = torch.randn(1, N, oh, ow, Co, device=dev).bfloat16()
    sx = torch.randn(1, N, oh, ow, Co, device=dev).bfloat16()
    s1, s2 = torch.zeros(Co, device=dev), torch.zeros(Co, device=dev)
    kw = dict(w_kn=kn)
    if "mask" in variant:
        kw.update(mask=K.View(mask), mask_kind=K.MASK_RELU)
    if "sum" in variant:
        kw.update(stat_sum=s1)
    if "sq" in variant:
        kw.update(stat_sq=s2)
    if "dot" in variant:
        kw.update(stat_dot=s2, stat_x=K.View(sx))
    if "lim" in variant:
        kw.update(stat_n=3 * N // 4)
    if "chain" in variant:      # 20 back-to-back launches replayed from a CUDA graph: per-launch time in a real step
        for k2 in ("mask", "mask_kind", "stat_sum", "stat_sq", "stat_dot", "stat_x", "stat_n"):
            kw.pop(k2, None)
        K.conv_gemm(mode, k, flip, K.View(x), w, K.View(y), **kw)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(20):
                K.conv_gemm(mode, k, flip, K.View(x), w, K.View(y), **kw)
        ts = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / 20)
        ts.sort()
        flops = 2.0 * N * oh * ow * Co * Ci * (k * k if mode == K.CONV_S1 else 16 if mode == K.CONV_K4S2 else 4)
        print("%-10s %-18s per launch in a 20-launch graph: median %.4f ms  min %.4f  %.0f TFLOP/s" % (
            name, variant, ts[len(ts) // 2], ts[0], flops / ts[len(ts) // 2] / 1e9))
        return
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(reps + warm):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        K.conv_gemm(mode, k, flip, K.View(x), w, K.View(y), **kw)
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    ts.sort()
    flops = 2.0 * N * oh * ow * Co * Ci * (k * k if mode == K.CONV_S1 else 16 if mode == K.CONV_K4S2 else 4)
    med = ts[len(ts) // 2]
    print("%-10s %-18s dbg=%-2s median %.4f ms  min %.4f  %.0f TFLOP/s" % (name, variant, os.environ.get("T2I_STAT_DBG", "0"),
                                                                        med, ts[0], flops / med / 1e9))


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if a in SHAPES] or list(SHAPES)
    variants = [a for a in sys.argv[1:] if a not in SHAPES] or ["plain", "mask", "mask+sum", "mask+sum+lim", "sum+sq", "mask+sum+dot"]
    for n in names:
        for v in variants:
            run(n, v)
