"""Independent naive NumPy loops for conv / transposed conv in TF layout -- TEST INFRASTRUCTURE.

Written straight from the TF-1.4 definitions (not from torch) so that it cross-checks the
layout/padding assumptions of oracle/wgancls_oracle.py (utils/ops.py:58-71 of the reference).
Small shapes only.
"""
import numpy as np


def same_pad(in_size, k, s):
    """TF SAME: out = ceil(in/s); pad_total = max((out-1)*s + k - in, 0); before = total // 2."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return out, total // 2, total - total // 2


def conv2d_nhwc(x, w_hwio, b, s, padding):
    """tf.nn.conv2d: y[n,ho,wo,co] = sum x[n, ho*s+kh-pt, wo*s+kw-pl, ci] * w[kh,kw,ci,co] + b."""
    n, h, wd, ci = x.shape
    kh, kw, _, co = w_hwio.shape
    if padding.upper() == "SAME":
        ho, pt, _ = same_pad(h, kh, s)
        wo, pl, _ = same_pad(wd, kw, s)
    else:
        ho, wo, pt, pl = (h - kh) // s + 1, (wd - kw) // s + 1, 0, 0
    y = np.zeros((n, ho, wo, co), dtype=np.float64)
    for oh in range(ho):
        for ow in range(wo):
            for a in range(kh):
                for c in range(kw):
                    ih, iw = oh * s + a - pt, ow * s + c - pl
                    if 0 <= ih < h and 0 <= iw < wd:
                        y[:, oh, ow, :] += x[:, ih, iw, :].astype(np.float64) @ w_hwio[a, c].astype(np.float64)
    return y + b


def conv2d_transpose_nhwc(x, w_hwoi, b, s):
    """tf.nn.conv2d_transpose SAME (= gradient of conv2d w.r.t. its input): every input pixel
    scatters x[n,hi,wi,ci] * w[kh,kw,co,ci] to output (hi*s + kh - pt, wi*s + kw - pl)."""
    n, h, wd, ci = x.shape
    kh, kw, co, _ = w_hwoi.shape
    ho, wo = h * s, wd * s
    _, pt, _ = same_pad(ho, kh, s)
    _, pl, _ = same_pad(wo, kw, s)
    y = np.zeros((n, ho, wo, co), dtype=np.float64)
    for ih in range(h):
        for iw in range(wd):
            for a in range(kh):
                for c in range(kw):
                    oh, ow = ih * s + a - pt, iw * s + c - pl
                    if 0 <= oh < ho and 0 <= ow < wo:
                        y[:, oh, ow, :] += x[:, ih, iw, :].astype(np.float64) @ w_hwoi[a, c].astype(np.float64).T
    return y + b
