#!/usr/bin/env python
"""Micro-benchmark of the direct kernels of the 3-channel ends and the fp32 head at the sizes of BASELINE config 2
(batch 256; the D run uses the 4B batch): CUDA events around each launch, a 512 MB write between launches to flush L2,
algorithmic bytes / time against the measured HBM peak.  Also the target of `ncu -k regex:...` captures.

    python tools/bench_img.py [--reps 5] [--only NAME]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2i_b200 import kernels as K  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default=None)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    B = a.batch
    dev = torch.device("cuda")
    peak = 6550.4
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    bf = dict(device=dev, dtype=torch.bfloat16)
    f32 = dict(device=dev, dtype=torch.float32)
    img4 = torch.rand(4 * B, 64, 64, 3, **f32) * 2 - 1
    a0 = torch.zeros(1, 4 * B, 32, 32, 128, **bf)
    d_a0 = torch.randn(1, 4 * B, 32, 32, 128, **f32).to(torch.bfloat16)
    w0 = (torch.randn(1, 1, 128, 64, **f32) * 0.1).to(torch.bfloat16)
    wt3 = (torch.randn(1, 1, 64, 128, **f32) * 0.1).to(torch.bfloat16)
    bias = torch.zeros(128, **f32)
    gx = torch.zeros(B, 64, 64, 3, **f32)
    u4, out = torch.zeros(B, 64, 64, 3, **f32), torch.zeros(B, 64, 64, 3, **f32)
    w9, b9, b3 = torch.randn(81, **f32) * 0.2, torch.zeros(3, **f32), torch.zeros(3, **f32)
    dw0, dwt3 = torch.zeros(1, 128, 64, **f32), torch.zeros(1, 64, 128, **f32)
    stat_a, stat_b = torch.zeros(128, **f32), torch.zeros(128, **f32)
    cond, wms, bms, ms = torch.randn(B, 1024, **f32), torch.randn(256, 1024, **f32) * 0.03, torch.zeros(256, **f32), torch.zeros(B, 256, **f32)
    c9_dx, c9_dw, c9_db, c9_sum = torch.zeros(B, 64, 64, 3, **f32), torch.zeros(81, **f32), torch.zeros(3, **f32), torch.zeros(3, **f32)
    V = K.View
    MB = 1e6
    rows4 = torch.zeros(1, 4 * B, 64, K.img_row_pitch(64), **bf)
    K.img_to_rows(img4, rows4)
    gxrows = torch.zeros(1, B, 64, K.img_row_pitch(64), **bf)
    P4, PB, PG = K.ImgPatches(rows4, 64), K.ImgPatches(rows4[:, :B], 64), K.ImgPatches(gxrows, 64)
    RB = 64 * K.img_row_pitch(64) * 2      # bytes of one image as padded bf16 rows
    cases = [
        ("h0_fwd_4B", lambda: K.conv_gemm(K.CONV_K4S2, 4, 0, P4, w0, V(a0), bias=bias, act=K.ACT_LRELU),
         (4 * B * RB + 4 * B * 1024 * 256) / MB),
        ("img_to_rows_4B", lambda: K.img_to_rows(img4, rows4), (4 * B * (12288 * 4 + RB)) / MB),
        ("h0_fwd_B", lambda: K.conv_gemm(K.CONV_K4S2, 4, 0, PB, w0, V(a0, 0, B), bias=bias, act=K.ACT_LRELU),
         (B * RB + B * 1024 * 256) / MB),
        ("h0_tangent_B", lambda: K.conv_gemm(K.CONV_K4S2, 4, 0, PB, w0, V(a0, 0, B), mask=V(a0, 0, B), mask_kind=K.MASK_LRELU),
         (B * RB + 2 * B * 1024 * 256) / MB),
        ("h0_wgrad_4B", lambda: K.wgrad_img(P4, V(d_a0), dw0, 1), (4 * B * RB + 4 * B * 1024 * 256) / MB),
        ("h0_dgrad_B", lambda: K.deconv_img(V(d_a0, 0, B), w0, gx, w_kn=True), (B * 12288 * 4 + B * 1024 * 256) / MB),
        ("up4_fwd_B", lambda: K.deconv_img(V(a0, 0, B), wt3, u4, bias3=b3, w9=w9, b9=b9, img=out), (2 * B * 12288 * 4 + B * 1024 * 256) / MB),
        ("up4_dgrad_B", lambda: K.conv_gemm(K.CONV_K4S2, 4, 0, PG, wt3, V(d_a0, 0, B), w_kn=True, mask=V(a0, 0, B), mask_kind=K.MASK_RELU,
                                            stat_sum=stat_a, stat_dot=stat_b, stat_x=V(a0, B, B)), (B * RB + 3 * B * 1024 * 256) / MB),
        ("up4_wgrad_B", lambda: K.wgrad_img(PG, V(a0, 0, B), dwt3, 2), (B * RB + B * 1024 * 256) / MB),
        ("c9_bwd_dx_B", lambda: K.conv3x3_c3_tanh_bwd(u4, w9.view(3, 3, 3, 3), out, gx, c9_dx, None, None, c9_sum), (3 * B * 12288 * 4) / MB),
        ("c9_bwd_dw_B", lambda: K.conv3x3_c3_tanh_bwd(u4, w9.view(3, 3, 3, 3), out, gx, None, c9_dw, c9_db), (3 * B * 12288 * 4) / MB),
        ("floor_tiny_kernel", lambda: K.scale_rows(cond[:4], bms[:4], ms[:4, :1024 // 4 * 0 + 256][:, :256].contiguous() if False else cond[4:8]), 0.03),
        ("dense_f32", lambda: K.dense_f32(cond, wms, bms, ms, act=K.ACT_LRELU), (B * 1024 * 4 + 256 * 1024 * 4 + B * 256 * 4) / MB),
    ]
    res = {}
    for name, fn, mb in cases:
        if a.only and a.only != name:
            continue
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        res[name] = {"us": t * 1e3, "algorithmic_MB": mb, "GB_per_s": mb / t, "frac_of_hbm_peak": mb / t / peak}
        print("%-14s %8.1f us  %7.1f MB algorithmic  %7.0f GB/s  (%.0f %% of %.0f)" % (name, t * 1e3, mb, mb / t, 100 * mb / t / peak, peak))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
