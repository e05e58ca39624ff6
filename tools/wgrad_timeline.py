#!/usr/bin/env python
"""Producer-loop timeline of t2i_wgrad_img (debug build only: make -C text-to-image_b200/csrc EXTRA=-DT2I_TIMELINE_BUILD,
then T2I_B200_LIB=<that .so> python tools/wgrad_timeline.py).  Prints %globaltimer marks of one CTA's producer trips."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from t2i_b200 import _lib, kernels as K
B = 1024; dev = "cuda"
img = torch.rand(B, 64, 64, 3, device=dev) * 2 - 1
rows = torch.zeros(1, B, 64, K.img_row_pitch(64), device=dev, dtype=torch.bfloat16); K.img_to_rows(img, rows)
dy = torch.randn(1, B, 32, 32, 128, device=dev).to(torch.bfloat16)
dw = torch.zeros(1, 128, 64, device=dev)
for _ in range(3): K.wgrad_img(K.ImgPatches(rows, 64), K.View(dy), dw, 1)
torch.cuda.synchronize()
lib = _lib.load(); buf = (C.c_ulonglong * 512)()
lib.t2i_debug_wgrad_timeline(buf, 512)
r = [[buf[t * 8 + e] for e in range(8)] for t in range(64)]
t0 = min(v for x in r for v in x if v)
print("trip  bar  fetch_done rows_ok empty_ok built  mma_see")
for t, x in enumerate(r[8:30]): print(t + 8, " ".join("%7d" % (v - t0 if v else -1) for v in x[:6]))
