#!/usr/bin/env python
"""Where does the bf16-mode forward error come from?  CPU emulation (no GPU needed).

Runs the oracle's generator / discriminator forward (models/wgancls/model.py:129-225) in fp32 with the
storage roundings the CUDA engine performs in ``precision="bf16"``: weights rounded to bf16 once, every tensor
the engine writes to HBM as bf16 rounded at that point (GEMM accumulation and BatchNorm statistics stay fp32,
as in the kernels).  Prints the relative L2 error of the image / logits against the unrounded fp32 run for
  * all roundings on,
  * each storage site alone,
  * all but the weights, all but the activations,
so that the error budget in DESIGN.md section 5 is reproducible.  Test infrastructure only (imports oracle/).

usage: python tools/bf16_error_budget.py [--batch 16] [--seed 0] [--fmt bf16|fp16]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wgancls_oracle as O  # noqa: E402


class Q:
    """rounding policy: site name -> round or not"""

    round_head = False

    def __init__(self, fmt, sites=None, weights=True):
        self.dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[fmt]
        self.sites, self.weights, self.seen = sites, weights, []

    def a(self, name, x):
        if name not in self.seen:
            self.seen.append(name)
        if self.sites is None or name in self.sites:
            return x.to(self.dt).float()
        return x

    def w(self, x):
        return x.to(self.dt).float() if self.weights else x


def rel(a, b):
    return float((a - b).norm() / b.norm())


def generator(p, z, embed, tn_eps, cfg, q):
    g = "g_net/"
    gf = cfg.gf_dim

    def conv(scope, x, k, name):
        w = q.w(p[scope + "/weights"]).permute(3, 2, 0, 1)
        return q.a(name, F.conv2d(x, w, p[scope + "/biases"], padding={1: 0, 3: 1}[k]))

    def deconv(scope, x, name):
        w = q.w(p[scope + "/weights"]).permute(3, 2, 0, 1)
        return q.a(name, F.conv_transpose2d(x, w, p[scope + "/biases"], stride=2, padding=1))

    def bn(scope, x, name, act=None, residual=None):
        dims = (0, 2, 3) if x.dim() == 4 else (0,)
        shape = (1, -1, 1, 1) if x.dim() == 4 else (1, -1)
        # statistics come from the fp32 accumulators (GEMM epilogue), the normalised tensor is the bf16 stored one
        y = (x - x.mean(dims).view(shape)) * torch.rsqrt(x.var(dims, unbiased=False).view(shape) + O.BN_EPS)
        y = y * p[scope + "/gamma"].view(shape) + p[scope + "/beta"].view(shape)
        if residual is not None:
            y = y + residual
        if act is not None:
            y = act(y)
        return q.a(name, y)

    # the conditioning head runs in fp32 since round 2 (t2i_dense_f32: fp32 cond, fp32 master weights, fp32 [mean | log_sigma]);
    # --round-head emulates round 1, where cond, the weights and the head's output were bf16
    head = q.round_head
    cond = q.a("g.cond", embed) if head else embed
    wq = q.w if head else (lambda t: t)
    ms_w = torch.cat([wq(p[g + "dense/kernel"]), wq(p[g + "dense_1/kernel"])], 1)
    ms = O.lrelu(cond @ ms_w + torch.cat([p[g + "dense/bias"], p[g + "dense_1/bias"]]))
    if head:
        ms = q.a("g.ms", ms)
    ce = cfg.compressed_embed_dim
    c = ms[:, :ce] + torch.exp(ms[:, ce:]) * tn_eps
    zc = q.a("g.zc", torch.cat([z, c], 1))
    f0 = q.a("g.f0", zc @ q.w(p[g + "dense_2/kernel"]) + p[g + "dense_2/bias"])
    h0 = bn(g + "BatchNorm", f0, "g.h0").reshape(-1, gf * 8, 4, 4)

    def res(x, c1, b1, c2, b2, c3, b3, tag):
        n = bn(g + b1, conv(g + c1, x, 1, tag + ".t_a"), tag + ".u_a", torch.relu)
        n = bn(g + b2, conv(g + c2, n, 3, tag + ".t_b"), tag + ".u_b", torch.relu)
        return bn(g + b3, conv(g + c3, n, 3, tag + ".t_c"), tag + ".out", torch.relu, residual=x)

    h1 = res(h0, "Conv", "BatchNorm_1", "Conv_1", "BatchNorm_2", "Conv_2", "BatchNorm_3", "g.res1")
    h2 = bn(g + "BatchNorm_4", conv(g + "Conv_3", deconv(g + "Conv2d_transpose", h1, "g.d1"), 3, "g.t4"), "g.h2")
    h3 = res(h2, "Conv_4", "BatchNorm_5", "Conv_5", "BatchNorm_6", "Conv_6", "BatchNorm_7", "g.res2")
    h4 = bn(g + "BatchNorm_8", conv(g + "Conv_7", deconv(g + "Conv2d_transpose_1", h3, "g.d2"), 3, "g.t8"), "g.h4",
            torch.relu)
    h5 = bn(g + "BatchNorm_9", conv(g + "Conv_8", deconv(g + "Conv2d_transpose_2", h4, "g.d3"), 3, "g.t9"), "g.h5",
            torch.relu)
    # 3-channel end: deconv output and the 3x3 conv run in fp32 in the engine (u4 is an fp32 tensor)
    w = q.w(p[g + "Conv2d_transpose_3/weights"]).permute(3, 2, 0, 1)
    lg = F.conv_transpose2d(h5, w, p[g + "Conv2d_transpose_3/biases"], stride=2, padding=1)
    lg = F.conv2d(lg, p[g + "Conv_9/weights"].permute(3, 2, 0, 1), p[g + "Conv_9/biases"], padding=1)
    return torch.tanh(lg).permute(0, 2, 3, 1)


def discriminator(p, x_nhwc, embed, cfg, q):
    d = "d_net/"

    def conv(scope, x, k, s, name, act=True, add=None, pad=None):
        w = q.w(p[scope + "/weights"]).permute(3, 2, 0, 1)
        y = F.conv2d(x, w, p[scope + "/biases"], stride=s, padding={1: 0, 3: 1, 4: 1}[k] if pad is None else pad)
        if add is not None:
            y = y + add
        return q.a(name, O.lrelu(y) if act else y)

    x = q.a("d.img", x_nhwc.permute(0, 3, 1, 2))      # the bf16 patch matrix of the image
    a0 = conv(d + "Conv", x, 4, 2, "d.a0")
    a1 = conv(d + "Conv_1", a0, 4, 2, "d.a1")
    a2 = conv(d + "Conv_2", a1, 4, 2, "d.a2")
    a3 = conv(d + "Conv_3", a2, 4, 2, "d.a3", act=False)
    r1 = conv(d + "Conv_4", a3, 1, 1, "d.r1")
    r2 = conv(d + "Conv_5", r1, 3, 1, "d.r2")
    h4 = conv(d + "Conv_6", r2, 3, 1, "d.cat", add=a3)
    e = q.a("d.e", O.lrelu(q.a("d.cond", embed) @ q.w(p[d + "dense/kernel"]) + p[d + "dense/bias"]))
    h4c = torch.cat([h4, e[:, :, None, None].expand(-1, -1, 4, 4)], 1)
    a5 = conv(d + "Conv_7", h4c, 3, 1, "d.a5")
    a6 = conv(d + "Conv_8", a5, 1, 1, "d.a6")
    return F.conv2d(a6, p[d + "Conv_9/weights"].permute(3, 2, 0, 1), p[d + "Conv_9/biases"], stride=4)   # fp32 dot product


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--fmt", default="bf16")
    ap.add_argument("--per-site", action="store_true")
    ap.add_argument("--round-head", action="store_true", help="round the conditioning head to the storage format too (round 1)")
    a = ap.parse_args()
    Q.round_head = a.round_head
    torch.set_grad_enabled(False)
    cfg = O.OracleCfg(batch_size=a.batch)
    p = O.init_params(cfg, a.seed)
    f = O.make_feed(cfg, 1234)
    none = Q(a.fmt, sites=(), weights=False)
    G0 = generator(p, f["z"], f["cond"], f["tn_eps"], cfg, none)
    Gref, _, _ = O.generator(p, f["z"], f["cond"], f["tn_eps"], cfg)
    print("emulation with no rounding vs oracle: %.2e" % rel(G0, Gref))
    D0 = discriminator(p, f["x"], f["cond"], cfg, none)
    print("same for D(x): %.2e" % rel(D0, O.discriminator(p, f["x"], f["cond"], cfg)))
    full = Q(a.fmt)
    G1 = generator(p, f["z"], f["cond"], f["tn_eps"], cfg, full)
    D1 = discriminator(p, f["x"], f["cond"], cfg, full)
    print("batch %d %s: all roundings  G %.3e   D(x) %.3e" % (a.batch, a.fmt, rel(G1, G0), rel(D1, D0)))
    for label, q in (("weights only", Q(a.fmt, sites=(), weights=True)), ("activations only", Q(a.fmt, None, False))):
        print("  %-18s G %.3e   D(x) %.3e" % (label, rel(generator(p, f["z"], f["cond"], f["tn_eps"], cfg, q), G0),
                                              rel(discriminator(p, f["x"], f["cond"], cfg, q), D0)))
    if a.per_site:
        for s in [s for s in full.seen if s.startswith("g.")]:
            q = Q(a.fmt, sites=(s,), weights=False)
            print("  only %-12s G %.3e" % (s, rel(generator(p, f["z"], f["cond"], f["tn_eps"], cfg, q), G0)))
        for s in [s for s in full.seen if s.startswith("d.")]:
            q = Q(a.fmt, sites=(s,), weights=False)
            print("  only %-12s D %.3e" % (s, rel(discriminator(p, f["x"], f["cond"], cfg, q), D0)))


if __name__ == "__main__":
    main()
