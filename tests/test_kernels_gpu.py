"""-m gpu: every CUDA kernel, called through the C ABI, against its CPU restatement
(tests/fake_kernels.py, fp64 torch on the same bf16-plane inputs).

Tolerances: np = 1 (bf16 storage) outputs may differ from the fp64 result by bf16 rounding of the
output (2^-8 relative) plus fp32 accumulation; np = 2 (split bf16) carries ~2^-16.  fp32 outputs
(weight gradients, statistics) are compared at 2e-3 / 2e-5 of the tensor's scale.
"""
import numpy as np
import pytest
import torch

import fake_kernels as fk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from t2i_b200 import kernels
    return kernels


def rand_planes(np_, shape, gen, scale=1.0):
    v = torch.randn(*shape, generator=gen) * scale
    t = torch.zeros(np_, *shape, dtype=torch.bfloat16)
    fk.put(t, v)
    return t


def tol(np_):
    return (2.0 ** -7, 1e-2) if np_ == 1 else (1e-4, 0.25)


def check_close(name, got, ref, rtol, atol_scale):
    got, ref = got.double().cpu(), ref.double().cpu()
    scale = float(ref.abs().max()) + 1e-30
    err = (got - ref).abs()
    bound = rtol * ref.abs() + atol_scale * rtol * scale
    bad = err > bound
    if bool(bad.any()):
        idx = torch.nonzero(bad)
        worst = int(torch.argmax(err - bound))
        wi = np.unravel_index(worst, tuple(err.shape))
        last = idx[:, -1]
        rows = idx[:, :-1]
        msg = ("%s: %d/%d bad; worst at %s got %.6g ref %.6g; scale %.4g; rel-L2 %.3g; first bad %s; "
               "bad last-dim mod 8 hist %s; bad count by first index %s" % (
                   name, int(bad.sum()), bad.numel(), wi, float(got[wi]), float(ref[wi]), scale,
                   float((got - ref).norm() / (ref.norm() + 1e-30)), idx[:6].tolist(),
                   torch.bincount(last % 8, minlength=8).tolist(),
                   torch.bincount(rows[:, 0], minlength=1)[:16].tolist() if rows.numel() else []))
        raise AssertionError(msg)


CONV_CASES = [
    # name, mode, k, flip, N, H, W, Cin, Cout
    ("fc_1tile", fk.CONV_S1, 1, 0, 128, 1, 1, 64, 128),
    ("fc_k256", fk.CONV_S1, 1, 0, 128, 1, 1, 256, 128),
    ("fc_ragged", fk.CONV_S1, 1, 0, 300, 1, 1, 256, 384),
    ("fc_wide", fk.CONV_S1, 1, 0, 256, 1, 1, 256, 2048),
    ("c1_4x4", fk.CONV_S1, 1, 0, 16, 4, 4, 128, 64),
    ("c3_4x4", fk.CONV_S1, 3, 0, 16, 4, 4, 128, 128),
    ("c3_8x8", fk.CONV_S1, 3, 0, 6, 8, 8, 64, 256),
    ("c3_32x32", fk.CONV_S1, 3, 0, 2, 32, 32, 128, 128),
    ("c3_flip", fk.CONV_S1, 3, 1, 4, 8, 8, 128, 64),
    ("c3_tinych", fk.CONV_S1, 3, 0, 4, 4, 4, 8, 16),
    ("k4s1_4x4", fk.CONV_S1, 4, 0, 16, 4, 4, 128, 128),          # TF SAME for k = 4, s = 1: pad 1 before, 2 after
    ("k4s1_16_flip", fk.CONV_S1, 4, 1, 3, 16, 16, 64, 128),
    ("k2s1_8x8", fk.CONV_S1, 2, 0, 6, 8, 8, 64, 16),              # TF SAME for k = 2: pad 0 before, 1 after (PGGAN to_rgb)
    ("k2s1_flip", fk.CONV_S1, 2, 1, 3, 16, 16, 16, 64),
    ("k4s2_32", fk.CONV_K4S2, 4, 0, 2, 32, 32, 128, 256),
    ("k4s2_8", fk.CONV_K4S2, 4, 0, 16, 8, 8, 64, 128),
    ("k4s2_tiny", fk.CONV_K4S2, 4, 0, 2, 8, 8, 16, 32),
    ("deconv_4", fk.DECONV_K4S2, 4, 0, 16, 4, 4, 128, 64),
    ("deconv_16", fk.DECONV_K4S2, 4, 0, 2, 16, 16, 256, 128),
    ("deconv_tiny", fk.DECONV_K4S2, 4, 0, 2, 4, 4, 32, 16),
]


def out_hw(mode, h, w):
    return (h, w) if mode == fk.CONV_S1 else (h // 2, w // 2) if mode == fk.CONV_K4S2 else (2 * h, 2 * w)


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_gemm_plain(K, case, np_):
    name, mode, k, flip, N, H, W, Cin, Cout = case
    gen = torch.Generator().manual_seed(hash(name) % 1000)
    taps = k * k if mode == fk.CONV_S1 else 16
    x = rand_planes(np_, (N, H, W, Cin), gen)
    w = rand_planes(np_, (taps, Cout, Cin), gen, scale=(taps * Cin) ** -0.5)
    oh, ow = out_hw(mode, H, W)
    y = torch.zeros(np_, N, oh, ow, Cout, dtype=torch.bfloat16)
    fk.conv_gemm(mode, k, flip, fk.View(x), w, fk.View(y))
    xg, wg = x.cuda(), w.cuda()
    yg = torch.full_like(y, 7.0).cuda()
    K.conv_gemm(mode, k, flip, K.View(xg), wg, K.View(yg))
    torch.cuda.synchronize()
    check_close(name, fk.val(yg.cpu()), fk.val(y), *tol(np_))
    # the same product with the weights stored [tap][contraction][output channel] (MN-major B operand),
    # which is how a layer's packed forward weights serve its input-gradient
    w_kn = w.transpose(2, 3).contiguous()
    yg2 = torch.full_like(y, 5.0).cuda()
    K.conv_gemm(mode, k, flip, K.View(xg), w_kn.cuda(), K.View(yg2), w_kn=True)
    torch.cuda.synchronize()
    check_close(name + "_kn", fk.val(yg2.cpu()), fk.val(y), *tol(np_))


@pytest.mark.parametrize("np_", [1, 2])
def test_conv_gemm_epilogue_and_windows(K, np_):
    """bias + residual add + LeakyReLU + derivative mask, channel windows with pitch != c,
    sample sub-ranges (segment views)."""
    gen = torch.Generator().manual_seed(5)
    N, H, W, Cin, Cout = 24, 4, 4, 128, 256
    xbuf = rand_planes(np_, (N, H, W, 192), gen)        # window channels [64, 192)
    w = rand_planes(np_, (9, Cout, Cin), gen, scale=(9 * Cin) ** -0.5)
    ybuf = rand_planes(np_, (N, H, W, 320), gen)         # write channels [64, 320)
    addbuf = rand_planes(np_, (N, H, W, Cout), gen)
    maskbuf = rand_planes(np_, (N, H, W, Cout), gen)
    bias = torch.randn(Cout, generator=gen)
    n0, n = 8, 16
    args = dict(bias=None, act=fk.ACT_LRELU, mask_kind=fk.MASK_LRELU)
    yref = ybuf.clone()
    fk.conv_gemm(fk.CONV_S1, 3, 0, fk.View(xbuf, n0, n, 64, 128), w, fk.View(yref, n0, n, 64, 256), bias=bias,
                 add=fk.View(addbuf, n0, n), mask=fk.View(maskbuf, n0, n), act=fk.ACT_LRELU, mask_kind=fk.MASK_LRELU)
    xg, wg, yg, ag, mg, bg = xbuf.cuda(), w.cuda(), ybuf.cuda(), addbuf.cuda(), maskbuf.cuda(), bias.cuda()
    K.conv_gemm(K.CONV_S1, 3, 0, K.View(xg, n0, n, 64, 128), wg, K.View(yg, n0, n, 64, 256), bias=bg,
                add=K.View(ag, n0, n), mask=K.View(mg, n0, n), act=K.ACT_LRELU, mask_kind=K.MASK_LRELU)
    torch.cuda.synchronize()
    check_close("epilogue", fk.val(yg.cpu()), fk.val(yref), *tol(np_))
    # untouched regions (other samples, channels < 64) must be bit-identical to the input buffer
    assert torch.equal(yg.cpu()[:, :n0], ybuf[:, :n0]) and torch.equal(yg.cpu()[..., :64], ybuf[..., :64])
    # in-place mask (mask source == destination), ReLU mask, deconv phase writes
    ybuf2 = rand_planes(np_, (4, 8, 8, 64), gen)
    x2 = rand_planes(np_, (4, 4, 4, 128), gen)
    w2 = rand_planes(np_, (16, 64, 128), gen, scale=(4 * 128) ** -0.5)
    yref2 = ybuf2.clone()
    fk.conv_gemm(fk.DECONV_K4S2, 4, 0, fk.View(x2), w2, fk.View(yref2), mask=fk.View(yref2), mask_kind=fk.MASK_RELU)
    yg2 = ybuf2.cuda()
    K.conv_gemm(K.DECONV_K4S2, 4, 0, K.View(x2.cuda()), w2.cuda(), K.View(yg2), mask=K.View(yg2), mask_kind=K.MASK_RELU)
    torch.cuda.synchronize()
    check_close("inplace_mask", fk.val(yg2.cpu()), fk.val(yref2), *tol(np_))



STAT_CASES = [
    # name, mode, k, N, H, W, Cin, Cout, stat_n, stat_c
    ("fc_rows", fk.CONV_S1, 1, 300, 1, 1, 64, 384, 0, 0),              # ragged last tile: OOB rows must not count
    ("fc_few", fk.CONV_S1, 1, 4, 1, 1, 64, 256, 0, 0),                 # fewer rows than one tile
    ("c3_4x4_lim", fk.CONV_S1, 3, 24, 4, 4, 128, 320, 17, 256),        # sample limit inside a tile, channel limit
    ("c3_pair", fk.CONV_S1, 3, 64, 4, 4, 128, 256, 48, 0),             # CTA-pair tiles, limit on a tile boundary
    ("c3_32x32", fk.CONV_S1, 3, 3, 32, 32, 64, 128, 2, 0),
    ("k4s2_16", fk.CONV_K4S2, 4, 6, 16, 16, 64, 128, 0, 0),
    ("deconv_8", fk.DECONV_K4S2, 4, 5, 8, 8, 128, 64, 3, 0),           # four output phases
    ("tinych", fk.CONV_S1, 3, 4, 4, 4, 8, 16, 0, 0),                   # Cout < one 64-channel sub-tile
]


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("case", STAT_CASES, ids=[c[0] for c in STAT_CASES])
def test_conv_gemm_epilogue_statistics(K, case, np_):
    """stat_sum / stat_sq / stat_dot of the epilogue against the restatement, with bias + activation + mask in
    front of them, ragged tiles, sample / channel limits and accumulation (+=) semantics."""
    name, mode, k, N, H, W, Cin, Cout, stat_n, stat_c = case
    gen = torch.Generator().manual_seed(hash(name) % 1000 + 7)
    taps = k * k if mode == fk.CONV_S1 else 16
    oh, ow = out_hw(mode, H, W)
    x = rand_planes(np_, (N, H, W, Cin), gen)
    w = rand_planes(np_, (taps, Cout, Cin), gen, scale=(taps * Cin / (16 if mode == fk.DECONV_K4S2 else 1)) ** -0.5)
    bias = torch.randn(Cout, generator=gen) * 0.5
    maskbuf = rand_planes(np_, (N, oh, ow, Cout), gen)
    sx = rand_planes(np_, (N, oh, ow, Cout), gen)
    lim = dict(stat_n=stat_n, stat_c=stat_c)
    ftol = 3e-5 if np_ == 2 else 2e-4
    # (a) BatchNorm-forward form: bias, then sum and sum of squares
    y = torch.zeros(np_, N, oh, ow, Cout, dtype=torch.bfloat16)
    s1, s2 = torch.zeros(Cout, dtype=torch.float64), torch.zeros(Cout, dtype=torch.float64)
    fk.conv_gemm(mode, k, 0, fk.View(x), w, fk.View(y), bias=bias, stat_sum=s1, stat_sq=s2, **lim)
    yg = torch.zeros_like(y).cuda()
    s1g, s2g = torch.ones(Cout, device="cuda"), torch.full((Cout,), 2.0, device="cuda")     # += semantics
    K.conv_gemm(mode, k, 0, K.View(x.cuda()), w.cuda(), K.View(yg), bias=bias.cuda(), stat_sum=s1g, stat_sq=s2g, **lim)
    torch.cuda.synchronize()
    check_close(name + " y", fk.val(yg.cpu()), fk.val(y), *tol(np_))
    check_close(name + " sum", s1g.cpu() - 1.0, s1, ftol, 10.0)
    check_close(name + " sq", s2g.cpu() - 2.0, s2, ftol, 10.0)
    if stat_c:
        assert torch.all(s1g.cpu()[stat_c:] == 1.0) and torch.all(s2g.cpu()[stat_c:] == 2.0)
    # (b) BatchNorm-backward form: ReLU derivative mask, then sum and dot with a second tensor
    d1, d2 = torch.zeros(Cout, dtype=torch.float64), torch.zeros(Cout, dtype=torch.float64)
    fk.conv_gemm(mode, k, 0, fk.View(x), w, fk.View(y), mask=fk.View(maskbuf), mask_kind=fk.MASK_RELU, stat_sum=d1,
                 stat_dot=d2, stat_x=fk.View(sx), **lim)
    d1g, d2g = torch.zeros(Cout, device="cuda"), torch.zeros(Cout, device="cuda")
    K.conv_gemm(mode, k, 0, K.View(x.cuda()), w.cuda(), K.View(yg), mask=K.View(maskbuf.cuda()), mask_kind=K.MASK_RELU,
                stat_sum=d1g, stat_dot=d2g, stat_x=K.View(sx.cuda()), **lim)
    torch.cuda.synchronize()
    check_close(name + " masked y", fk.val(yg.cpu()), fk.val(y), *tol(np_))
    check_close(name + " dsum", d1g.cpu(), d1, ftol, 10.0)
    check_close(name + " ddot", d2g.cpu(), d2, ftol, 10.0)


@pytest.mark.parametrize("np_", [1, 2])
def test_conv_gemm_statistics_with_residual_and_kn(K, np_):
    """three epilogue tensors at once (residual add + mask + dot factor) on the KN (input-gradient) weight layout."""
    gen = torch.Generator().manual_seed(21)
    N, H, W, Cin, Cout = 20, 4, 4, 256, 192
    x = rand_planes(np_, (N, H, W, Cin), gen)
    w = rand_planes(np_, (9, Cin, Cout), gen, scale=(9 * Cin) ** -0.5)     # [tap][contraction][output channel]
    add, mask, sx = [rand_planes(np_, (N, H, W, Cout), gen) for _ in range(3)]
    y = torch.zeros(np_, N, H, W, Cout, dtype=torch.bfloat16)
    d1, d2 = torch.zeros(Cout, dtype=torch.float64), torch.zeros(Cout, dtype=torch.float64)
    fk.conv_gemm(fk.CONV_S1, 3, 1, fk.View(x), w, fk.View(y), add=fk.View(add), mask=fk.View(mask),
                 mask_kind=fk.MASK_LRELU, w_kn=True, stat_sum=d1, stat_dot=d2, stat_x=fk.View(sx))
    yg = torch.zeros_like(y).cuda()
    d1g, d2g = torch.zeros(Cout, device="cuda"), torch.zeros(Cout, device="cuda")
    K.conv_gemm(K.CONV_S1, 3, 1, K.View(x.cuda()), w.cuda(), K.View(yg), add=K.View(add.cuda()), mask=K.View(mask.cuda()),
                mask_kind=K.MASK_LRELU, w_kn=True, stat_sum=d1g, stat_dot=d2g, stat_x=K.View(sx.cuda()))
    torch.cuda.synchronize()
    check_close("y", fk.val(yg.cpu()), fk.val(y), *tol(np_))
    check_close("dsum", d1g.cpu(), d1, 3e-5 if np_ == 2 else 2e-4, 10.0)
    check_close("ddot", d2g.cpu(), d2, 3e-5 if np_ == 2 else 2e-4, 10.0)


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("rows,c", [(64, 16384), (4096, 256), (2048, 24), (70000, 128)])
def test_fused_batch_norm_kernels(K, np_, rows, c):
    """bn_apply_train (statistics finished in the apply kernel, moving statistics stepped) and bn_bwd_fused
    (reductions supplied, dgamma / bias gradient produced) against the restatement."""
    gen = torch.Generator().manual_seed(13)
    x, xg = both(np_, (rows, c), gen, 2.0)
    dy, dyg = both(np_, (rows, c), gen)
    res, resg = both(np_, (rows, c), gen)
    gamma, beta = torch.rand(c, generator=gen) + 0.5, torch.randn(c, generator=gen)
    xv = fk.val(x)
    sums = torch.cat([xv.sum(0), (xv * xv).sum(0)]).float()
    mean, rstd, var = torch.zeros(c), torch.zeros(c), torch.zeros(c)
    mm, mv = torch.randn(c, generator=gen), torch.rand(c, generator=gen)
    mmg, mvg = mm.cuda(), mv.cuda()
    y = torch.zeros_like(x)
    fk.bn_apply_train(x, sums.double(), 1e-5, gamma, beta, y, mean, rstd, var, residual=res, relu=True, moving=(mm, mv))
    yg = torch.zeros_like(x).cuda()
    mg, rg, vg = [torch.zeros(c, device="cuda") for _ in range(3)]
    K.bn_apply_train(xg, sums.cuda(), 1e-5, gamma.cuda(), beta.cuda(), yg, mg, rg, vg, residual=resg, relu=True,
                     moving=(mmg, mvg))
    check_close("mean", mg.cpu(), mean, 1e-4, 1.0)
    check_close("var", vg.cpu(), var, 1e-4, 1.0)
    check_close("rstd", rg.cpu(), rstd, 1e-4, 1.0)
    check_close("y", fk.val(yg.cpu()), fk.val(y), *tol(np_))
    check_close("mm", mmg.cpu(), mm, 1e-5, 1.0)
    check_close("mv", mvg.cpu(), mv, 1e-4, 1.0)
    # without the moving pair nothing else may change
    K.bn_apply_train(xg, sums.cuda(), 1e-5, gamma.cuda(), beta.cuda(), yg, mg, rg, vg)
    check_close("mm untouched", mmg.cpu(), mm, 1e-5, 1.0)
    dyv = fk.val(dy)
    dbeta, dot = dyv.sum(0).float(), (dyv * xv).sum(0).float()
    dga, dxs = torch.full((c,), 0.25, dtype=torch.float64), torch.zeros(c, dtype=torch.float64)
    dx = torch.zeros_like(x)
    fk.bn_bwd_fused(dy, x, mean, rstd, gamma, dot.double(), dbeta.double(), dga, dx, dxs)
    dgag, dxsg = torch.full((c,), 0.25, device="cuda"), torch.zeros(c, device="cuda")
    dxg = torch.zeros_like(x).cuda()
    K.bn_bwd_fused(dyg, xg, mean.cuda(), rstd.cuda(), gamma.cuda(), dot.cuda(), dbeta.cuda(), dgag, dxg, dxsg)
    check_close("dgamma", dgag.cpu(), dga, 5e-4, 5.0)
    check_close("dx", fk.val(dxg.cpu()), fk.val(dx), *tol(np_))
    # sum of dx is zero up to rounding (a bias in front of a BatchNorm has no gradient): bound it by the rounding scale
    scale = float(fk.val(dx).abs().sum(0).max())
    assert float((dxsg.cpu().double() - dxs).abs().max()) < (2.0 ** -8 if np_ == 1 else 1e-4) * scale + 1e-6


WGRAD_CASES = [
    ("fc", fk.CONV_S1, 1, 256, 1, 1, 128, 128),
    ("fc_ragged", fk.CONV_S1, 1, 200, 1, 1, 256, 384),
    ("c1_4x4", fk.CONV_S1, 1, 16, 4, 4, 128, 64),
    ("c3_4x4", fk.CONV_S1, 3, 16, 4, 4, 128, 128),
    ("c3_16x16", fk.CONV_S1, 3, 3, 16, 16, 64, 128),
    ("c3_tiny", fk.CONV_S1, 3, 4, 4, 4, 8, 16),
    ("k4s1_8x8", fk.CONV_S1, 4, 6, 8, 8, 64, 128),
    ("k2s1_16", fk.CONV_S1, 2, 3, 16, 16, 64, 16),                # PGGAN to_rgb: 2x2 SAME conv to 16 (9 used) channels
    ("c3_to8", fk.CONV_S1, 3, 2, 16, 16, 32, 8),
    ("c3_wide_map", fk.CONV_S1, 3, 9, 256, 256, 32, 32),         # > 64 MB of operands, one channel tile: taps-fastest work order                 # 8 padded output channels (stage-II image conv)
    ("k4s2_32", fk.CONV_K4S2, 4, 2, 32, 32, 128, 128),
    ("k4s2_8", fk.CONV_K4S2, 4, 16, 8, 8, 64, 256),
    ("deconv_4", fk.DECONV_K4S2, 4, 16, 4, 4, 128, 64),
    ("deconv_16", fk.DECONV_K4S2, 4, 2, 16, 16, 256, 128),
    # >= 256 channels on both sides: the CTA-pair (cta_group::2) tile
    ("c3_pair", fk.CONV_S1, 3, 8, 4, 4, 256, 512),
    ("k4s2_pair", fk.CONV_K4S2, 4, 4, 16, 16, 256, 256),
    ("deconv_pair", fk.DECONV_K4S2, 4, 4, 4, 4, 512, 320),
]


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("case", WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
def test_wgrad_gemm(K, case, np_):
    name, mode, k, N, H, W, Cin, Cout = case
    gen = torch.Generator().manual_seed(hash(name) % 1000 + 1)
    taps = k * k if mode == fk.CONV_S1 else 16
    oh, ow = out_hw(mode, H, W)
    x = rand_planes(np_, (N, H, W, Cin), gen)
    dy = rand_planes(np_, (N, oh, ow, Cout), gen)
    dw = torch.zeros(taps, Cout, Cin)
    fk.wgrad_gemm(mode, k, fk.View(x), fk.View(dy), dw)
    dwg = torch.zeros(taps, Cout, Cin, device="cuda")
    K.wgrad_gemm(mode, k, K.View(x.cuda()), K.View(dy.cuda()), dwg)
    torch.cuda.synchronize()
    check_close(name, dwg.cpu(), dw, 2e-5 if np_ == 2 else 1e-4, 10.0)
    # accumulation semantics (+=) and an explicit split-K
    # (a contraction over several 1e5 pixels in ONE fp32 accumulator loses ~1e-3: the wide map keeps short splits)
    K.wgrad_gemm(mode, k, K.View(x.cuda()), K.View(dy.cuda()), dwg, split_k=3 if N * H * W < 100000 else 300)
    torch.cuda.synchronize()
    check_close(name + "_acc", dwg.cpu(), 2 * dw, 2e-5 if np_ == 2 else 1e-4, 10.0)


# ------------------------------------------------------------------------------------------
# HBM-bound kernels
def both(np_, shape, gen, scale=1.0):
    t = rand_planes(np_, shape, gen, scale)
    return t, t.cuda()


@pytest.mark.parametrize("np_", [1, 2])
def test_planes_im2col_col2im(K, np_):
    gen = torch.Generator().manual_seed(1)
    src = torch.randn(6, 40, generator=gen)
    rs = torch.rand(6, generator=gen)
    dst = torch.zeros(np_, 6, 40, dtype=torch.bfloat16)
    fk.to_planes(src, dst, rs)
    dg = torch.zeros_like(dst).cuda()
    K.to_planes(src.cuda(), dg, rs.cuda())
    assert torch.equal(dg.cpu(), dst)
    back = torch.zeros(6, 40, device="cuda")
    K.from_planes(dg, back)
    np.testing.assert_allclose(back.cpu().numpy(), fk.val(dst).float().numpy(), rtol=1e-6)
    img = torch.rand(3, 16, 16, 3, generator=gen) * 2 - 1
    sc = torch.rand(3, generator=gen) + 0.5
    col = torch.zeros(np_, 3 * 64, 64, dtype=torch.bfloat16)
    fk.im2col_k4s2_c3(img, col, sc)
    cg = torch.full_like(col, 3.0).cuda()
    K.im2col_k4s2_c3(img.cuda(), cg, sc.cuda())
    check_close("im2col", fk.val(cg.cpu()), fk.val(col), 2.0 ** -8 if np_ == 1 else 1e-5, 1e-3)
    c2, c2g = both(np_, (3 * 64, 64), gen)
    bias = torch.randn(3, generator=gen)
    out = torch.zeros(3, 16, 16, 3)
    fk.col2im_k4s2_c3(c2, out, bias)
    og = torch.zeros(3, 16, 16, 3, device="cuda")
    K.col2im_k4s2_c3(c2g, og, bias.cuda())
    check_close("col2im", og.cpu(), out, 1e-5, 1.0)
    # 256-wide images (StackGAN stage-II discriminator): the patch rows are staged in 64-pixel column strips
    for (n, h, w) in ((1, 20, 256), (2, 256, 256), (1, 6, 136)):
        img = torch.rand(n, h, w, 3, generator=gen) * 2 - 1
        col = torch.zeros(np_, n * (h // 2) * (w // 2), 64, dtype=torch.bfloat16)
        fk.im2col_k4s2_c3(img, col, None)
        cg = torch.full_like(col, 3.0).cuda()
        K.im2col_k4s2_c3(img.cuda(), cg, None)
        check_close("im2col %dx%d" % (h, w), fk.val(cg.cpu()), fk.val(col), 2.0 ** -8 if np_ == 1 else 1e-5, 1e-3)
        c2, c2g = both(np_, (n * (h // 2) * (w // 2), 64), gen)
        out = torch.zeros(n, h, w, 3)
        fk.col2im_k4s2_c3(c2, out, None)
        og = torch.full((n, h, w, 3), 9.0, device="cuda")
        K.col2im_k4s2_c3(c2g, og, None)
        check_close("col2im %dx%d" % (h, w), og.cpu(), out, 1e-5, 1.0)


@pytest.mark.parametrize("np_", [1, 2])
def test_stage2_image_end_kernels(K, np_):
    gen = torch.Generator().manual_seed(8)
    img = torch.rand(2, 16, 16, 3, generator=gen) * 2 - 1
    col = torch.zeros(np_, 2 * 256, 32, dtype=torch.bfloat16)
    fk.im2col_k3s1_c3(img, col)
    cg = torch.full_like(col, 3.0).cuda()
    K.im2col_k3s1_c3(img.cuda(), cg)
    check_close("im2col3", fk.val(cg.cpu()), fk.val(col), 2.0 ** -8 if np_ == 1 else 1e-5, 1e-3)
    lg, lgg = both(np_, (2, 16, 16, 8), gen)
    y = torch.zeros(2, 16, 16, 3)
    fk.tanh_c3_fwd(lg, y)
    yg = torch.zeros_like(y).cuda()
    K.tanh_c3_fwd(lgg, yg)
    check_close("tanh fwd", yg.cpu(), y, 1e-5, 1.0)
    dy = torch.randn(2, 16, 16, 3, generator=gen)
    dl = torch.zeros_like(lg)
    fk.tanh_c3_bwd(y, dy, dl)
    dlg = torch.full_like(lg, 1.0).cuda()
    K.tanh_c3_bwd(yg, dy.cuda(), dlg)
    check_close("tanh bwd", fk.val(dlg.cpu()), fk.val(dl), *tol(np_))
    # the 32 -> 3 image conv as a GEMM with 8 padded output channels, forward and input-gradient
    x, xg = both(np_, (2, 16, 16, 32), gen)
    w = rand_planes(np_, (9, 8, 32), gen, scale=(9 * 32) ** -0.5)
    out = torch.zeros(np_, 2, 16, 16, 8, dtype=torch.bfloat16)
    fk.conv_gemm(fk.CONV_S1, 3, 0, fk.View(x), w, fk.View(out))
    og = torch.zeros_like(out).cuda()
    K.conv_gemm(K.CONV_S1, 3, 0, K.View(xg), w.cuda(), K.View(og))
    check_close("conv to 8", fk.val(og.cpu()), fk.val(out), *tol(np_))
    dx = torch.zeros_like(x)
    fk.conv_gemm(fk.CONV_S1, 3, 1, fk.View(out), w, fk.View(dx), w_kn=True)
    dxg = torch.zeros_like(x).cuda()
    K.conv_gemm(K.CONV_S1, 3, 1, K.View(out.cuda()), w.cuda(), K.View(dxg), w_kn=True)     # same input planes as the restatement
    check_close("dgrad from 8", fk.val(dxg.cpu()), fk.val(dx), *tol(np_))


def test_conv3x3_c3_tanh(K):
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(3, 16, 16, 3, generator=gen)
    w = torch.randn(3, 3, 3, 3, generator=gen) * 0.3
    b = torch.randn(3, generator=gen) * 0.1
    y = torch.zeros_like(x)
    fk.conv3x3_c3_tanh_fwd(x, w, b, y)
    yg = torch.zeros_like(x).cuda()
    K.conv3x3_c3_tanh_fwd(x.cuda(), w.cuda(), b.cuda(), yg)
    check_close("c9 fwd", yg.cpu(), y, 1e-5, 1.0)
    dy = torch.randn(3, 16, 16, 3, generator=gen)
    dx, dw, db, dxs = torch.zeros_like(x), torch.zeros(81), torch.zeros(3), torch.zeros(3)
    fk.conv3x3_c3_tanh_bwd(x, w, y, dy, dx, dw, db, dxs)
    dxg, dwg, dbg, dxsg = [torch.zeros_like(t).cuda() for t in (dx, dw, db, dxs)]
    K.conv3x3_c3_tanh_bwd(x.cuda(), w.cuda(), yg, dy.cuda(), dxg, dwg, dbg, dxsg)
    check_close("c9 dx", dxg.cpu(), dx, 1e-4, 1.0)
    check_close("c9 dw", dwg.cpu(), dw, 1e-4, 1.0)
    check_close("c9 db", dbg.cpu(), db, 1e-4, 1.0)
    check_close("c9 dxsum", dxsg.cpu(), dxs, 1e-4, 1.0)


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("rows,c", [(64, 16384), (4096, 256), (2048, 24), (70000, 128)])
def test_batch_norm_kernels(K, np_, rows, c):
    gen = torch.Generator().manual_seed(3)
    x, xg = both(np_, (rows, c), gen, 2.0)
    dy, dyg = both(np_, (rows, c), gen)
    res, resg = both(np_, (rows, c), gen)
    gamma, beta = torch.rand(c, generator=gen) + 0.5, torch.randn(c, generator=gen)
    mean, rstd, var = torch.zeros(c), torch.zeros(c), torch.zeros(c)
    fk.bn_stats(x, None, mean, rstd, var, 1e-5)
    mg, rg, vg = [torch.zeros(c, device="cuda") for _ in range(3)]
    scratch = torch.zeros(2 * c, device="cuda")
    K.bn_stats(xg, scratch, mg, rg, vg, 1e-5)
    K.bn_stats(xg, scratch, mg, rg, vg, 1e-5)          # the scratch is left zeroed: a second call is identical
    assert float(scratch.abs().max()) == 0.0
    check_close("bn mean", mg.cpu(), mean, 1e-4, 1.0)
    check_close("bn var", vg.cpu(), var, 1e-4, 1.0)
    check_close("bn rstd", rg.cpu(), rstd, 1e-4, 1.0)
    y = torch.zeros_like(x)
    fk.bn_apply(x, mean, rstd, gamma, beta, y, res, True)
    yg = torch.zeros_like(x).cuda()
    K.bn_apply(xg, mean.cuda(), rstd.cuda(), gamma.cuda(), beta.cuda(), yg, resg, True)
    check_close("bn apply", fk.val(yg.cpu()), fk.val(y), *tol(np_))
    dga, dbe = torch.zeros(c), torch.zeros(c)
    fk.bn_bwd_reduce(dy, x, mean, rstd, dga, dbe)
    dgag, dbeg = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    K.bn_bwd_reduce(dyg, xg, mean.cuda(), rstd.cuda(), dgag, dbeg)
    check_close("bn dgamma", dgag.cpu(), dga, 2e-4, 5.0)
    check_close("bn dbeta", dbeg.cpu(), dbe, 2e-4, 5.0)
    dx = torch.zeros_like(x)
    fk.bn_bwd_apply(dy, x, mean, rstd, gamma, dga, dbe, dx)
    dxg = torch.zeros_like(x).cuda()
    K.bn_bwd_apply(dyg, xg, mean.cuda(), rstd.cuda(), gamma.cuda(), dga.cuda(), dbe.cuda(), dxg)
    check_close("bn dx", fk.val(dxg.cpu()), fk.val(dx), *tol(np_))
    mm, mv = torch.randn(c, generator=gen), torch.rand(c, generator=gen)
    mmg, mvg = mm.cuda(), mv.cuda()
    fk.bn_update_moving(mm, mv, mean, var, rows, 0.9)
    K.bn_update_moving(mmg, mvg, mean.cuda(), var.cuda(), rows, 0.9)
    check_close("bn mm", mmg.cpu(), mm, 1e-5, 1.0)
    check_close("bn mv", mvg.cpu(), mv, 1e-5, 1.0)
    a = torch.zeros_like(x)
    fk.act_bwd(dy, res, a, fk.MASK_LRELU)
    ag = torch.zeros_like(x).cuda()
    K.act_bwd(dyg, resg, ag, K.MASK_LRELU)
    check_close("act_bwd", fk.val(ag.cpu()), fk.val(a), *tol(np_))
    s = torch.zeros(c)
    fk.colsum(fk.View(x), s)
    sg = torch.zeros(c, device="cuda")
    K.colsum(K.View(xg), sg)
    check_close("colsum", sg.cpu(), s, 2e-4, 5.0)


@pytest.mark.parametrize("np_", [1, 2])
def test_discriminator_side_kernels(K, np_):
    gen = torch.Generator().manual_seed(4)
    S, ce, C = 12, 16, 64
    e, eg = both(np_, (S, ce), gen)
    cat, catg = both(np_, (S, 4, 4, C + ce), gen)
    fk.embed_tile(e, cat, C)
    K.embed_tile(eg, catg, C)
    assert torch.equal(catg.cpu(), cat)
    de = torch.zeros_like(e)
    fk.embed_reduce(cat, de, C)
    deg = torch.zeros_like(e).cuda()
    K.embed_reduce(catg, deg, C)
    check_close("embed_reduce", fk.val(deg.cpu()), fk.val(de), *tol(np_))
    a, ag = both(np_, (S, 4, 4, C), gen)
    w = torch.randn(16 * C, generator=gen) * 0.05
    b = torch.randn(1, generator=gen)
    seed = torch.randn(S, generator=gen)
    logit = torch.zeros(S)
    fk.dout_fwd(a, w, b, logit)
    lg = torch.zeros(S, device="cuda")
    K.dout_fwd(ag, w.cuda(), b.cuda(), lg)
    check_close("dout_fwd", lg.cpu(), logit, 1e-5, 1.0)
    da = torch.zeros_like(a)
    fk.dout_bwd_data(a, w, seed, da)
    dag = torch.zeros_like(a).cuda()
    K.dout_bwd_data(ag, w.cuda(), seed.cuda(), dag)
    check_close("dout_bwd_data", fk.val(dag.cpu()), fk.val(da), *tol(np_))
    dw, db = torch.zeros(16 * C), torch.zeros(1)
    fk.dout_bwd_weight(a, seed, dw, db, 9)
    dwg, dbg = torch.zeros(16 * C, device="cuda"), torch.zeros(1, device="cuda")
    K.dout_bwd_weight(ag, seed.cuda(), dwg, dbg, 9)
    check_close("dout dw", dwg.cpu(), dw, 1e-4, 1.0)
    check_close("dout db", dbg.cpu(), db, 1e-4, 1.0)


def test_gp_ca_scalar_adam_pack_kernels(K):
    gen = torch.Generator().manual_seed(5)
    B = 8
    g, x = torch.rand(B, 64, 64, 3, generator=gen), torch.rand(B, 64, 64, 3, generator=gen)
    eps = torch.rand(B, generator=gen)
    xh = torch.zeros_like(g)
    fk.gp_interp(g, x, eps, xh)
    xhg = torch.zeros_like(g).cuda()
    K.gp_interp(g.cuda(), x.cuda(), eps.cuda(), xhg)
    check_close("gp_interp", xhg.cpu(), xh, 1e-6, 1.0)
    grad = torch.randn(B, 64, 64, 3, generator=gen) * 0.012
    grad[0] *= 0.1      # slope < 1: coefficient must be exactly 0
    sl, co, pen = torch.zeros(B), torch.zeros(B), torch.zeros(1)
    fk.gp_penalty(grad, 150.0, 1.0 / 16, sl, co, pen)
    slg, cog, peng = [torch.zeros_like(t).cuda() for t in (sl, co, pen)]
    K.gp_penalty(grad.cuda(), 150.0, 1.0 / 16, slg, cog, peng)
    check_close("slope", slg.cpu(), sl, 1e-5, 1.0)
    check_close("coef", cog.cpu(), co, 1e-4, 1.0)
    check_close("pen", peng.cpu(), pen, 1e-4, 1.0)
    assert float(co[0]) == 0.0 and float(cog[0]) == 0.0 and float(co[1]) > 0
    for np_ in (1, 2):
        ce, zd = 16, 24
        ms = torch.randn(B, 2 * ce, generator=gen) * 0.5          # fp32 [mean | log_sigma] (dense_f32)
        msg_ = ms.cuda()
        z, tn = torch.randn(B, zd, generator=gen), torch.randn(B, ce, generator=gen)
        zc, kl = torch.zeros(np_, B, zd + ce, dtype=torch.bfloat16), torch.zeros(1)
        fk.ca_fwd(ms, z, tn, zc, kl)
        zcg, klg = torch.zeros_like(zc).cuda(), torch.zeros(1, device="cuda")
        K.ca_fwd(msg_, z.cuda(), tn.cuda(), zcg, klg)
        check_close("ca_fwd", fk.val(zcg.cpu()), fk.val(zc), *tol(np_))
        check_close("kl", klg.cpu(), kl, 1e-4, 1.0)
        dzc, dzcg = both(np_, (B, zd + ce), gen)
        dms = torch.zeros(np_, B, 2 * ce, dtype=torch.bfloat16)
        fk.ca_bwd(ms, dzc, tn, dms, zd, 0.01)
        dmsg = torch.zeros_like(dms).cuda()
        K.ca_bwd(msg_, dzcg, tn.cuda(), dmsg, zd, 0.01)
        check_close("ca_bwd", fk.val(dmsg.cpu()), fk.val(dms), *tol(np_))
        w = torch.randn(9, 40, 24, generator=gen)
        f, bwd = torch.zeros(np_, 9, 40, 24, dtype=torch.bfloat16), torch.zeros(np_, 9, 24, 40, dtype=torch.bfloat16)
        fk.pack_weight(w, f, bwd)
        fg, bg = torch.zeros_like(f).cuda(), torch.zeros_like(bwd).cuda()
        K.pack_weight(w.cuda(), fg, bg)
        assert torch.equal(fg.cpu(), f) and torch.equal(bg.cpu(), bwd)
    # scalars
    kt = torch.tensor([0.7])
    logit = torch.randn(4 * B, generator=gen)
    seed, sums, sc = torch.zeros(4 * B), torch.zeros(8), torch.zeros(16)
    sums[4], sums[5] = 0.3, 0.2
    fk.d_seeds(kt, seed, B, 1.0 / 16)
    fk.d_sums(logit, B, sums)
    ktg, seedg, sumsg, scg = kt.cuda(), torch.zeros(4 * B, device="cuda"), torch.zeros(8, device="cuda"), torch.zeros(16, device="cuda")
    sumsg[4], sumsg[5] = 0.3, 0.2
    K.d_seeds(ktg, seedg, B, 1.0 / 16)
    K.d_sums(logit.cuda(), B, sumsg)
    check_close("seeds", seedg.cpu(), seed, 1e-6, 1.0)
    check_close("sums", sumsg.cpu(), sums, 1e-5, 1.0)
    fk.d_scalars(sums, kt, sc, 16, 150.0, 1e-3)
    K.d_scalars(sumsg, ktg, scg, 16, 150.0, 1e-3)
    check_close("d_scalars", scg.cpu()[:12], sc[:12], 1e-5, 1.0)
    check_close("kt", ktg.cpu(), kt, 1e-6, 1.0)
    gs, gsg = torch.tensor([1.5, 40.0] + [0.0] * 6), torch.tensor([1.5, 40.0] + [0.0] * 6).cuda()
    fk.g_sums(logit, B, gs)
    K.g_sums(logit.cuda(), B, gsg)
    fk.g_scalars(gs, sc, 16, 8, 1.0)
    K.g_scalars(gsg, scg, 16, 8, 1.0)
    check_close("g_scalars", scg.cpu()[12:14], sc[12:14], 1e-5, 1.0)
    # Adam (TF form), odd length exercises the tail
    n = 1003
    th, gr = torch.randn(n, generator=gen), torch.randn(n, generator=gen)
    m, v = torch.rand(n, generator=gen), torch.rand(n, generator=gen)
    for b1, np_ in ((0.5, 1), (0.0, 2)):
        th1, m1, v1 = th.clone(), m.clone(), v.clone()
        thg, mg_, vg_ = th.cuda(), m.cuda(), v.cuda()
        pk = torch.zeros(np_, 1008, dtype=torch.bfloat16)[:, :n]      # plane stride keeps 16-byte alignment
        pkg = torch.zeros(np_, 1008, dtype=torch.bfloat16).cuda()[:, :n]
        lr_t = torch.tensor([3e-5])
        fk.adam_tf(th1, gr, m1, v1, lr_t, b1, 0.9, packed=pk)
        K.adam_tf(thg, gr.cuda(), mg_, vg_, lr_t.cuda(), b1, 0.9, packed=pkg)
        check_close("adam theta", thg.cpu(), th1, 1e-6, 1.0)
        check_close("adam m", mg_.cpu(), m1, 1e-6, 1.0)     # untouched when beta1 == 0
        check_close("adam v", vg_.cpu(), v1, 1e-6, 1.0)
        check_close("adam packed", fk.val(pkg.cpu()), fk.val(pk), 2.0 ** -7 if np_ == 1 else 1e-4, 1e-2)


@pytest.mark.parametrize("np_", [1, 2])
def test_conv_gemm_output_channel_window(K, np_):
    """w_n0: the output is a window of the weight matrix's output channels (stage-II splits the gradient of the
    generator's concat buffer into two windows with different epilogues), in both weight layouts."""
    gen = torch.Generator().manual_seed(23)
    N, H, W, Cin, Cout = 16, 4, 4, 128, 640
    x, xg = both(np_, (N, H, W, Cin), gen)
    w = rand_planes(np_, (9, Cout, Cin), gen, scale=(9 * Cin) ** -0.5)
    w_kn = w.transpose(2, 3).contiguous()
    ybuf = rand_planes(np_, (N, H, W, Cout), gen)
    mask, maskg = both(np_, (N, H, W, Cout), gen)
    for n0, c, epi in ((0, 512, True), (512, 128, False), (128, 192, False)):
        yref = ybuf.clone()
        kw_f = dict(mask=fk.View(mask, coff=n0, c=c), mask_kind=fk.MASK_RELU) if epi else {}
        fk.conv_gemm(fk.CONV_S1, 3, 1, fk.View(x), w, fk.View(yref, coff=n0, c=c), w_n0=n0, **kw_f)
        for kn in (False, True):
            yg = ybuf.cuda()
            kw_g = dict(mask=K.View(maskg, coff=n0, c=c), mask_kind=K.MASK_RELU) if epi else {}
            K.conv_gemm(K.CONV_S1, 3, 1, K.View(xg), (w_kn if kn else w).cuda(), K.View(yg, coff=n0, c=c), w_kn=kn,
                        w_n0=n0, **kw_g)
            torch.cuda.synchronize()
            check_close("window %d+%d kn=%d" % (n0, c, kn), fk.val(yg.cpu()), fk.val(yref), *tol(np_))


@pytest.mark.parametrize("np_", [1, 2])
def test_fused_batch_norm_pitch_and_affine_scale(K, np_):
    """y_pitch / dy_pitch (the tensor is the leading channels of a wider concat buffer) and affine_scale = 2
    (models/stackgan/stageII/model.py:117 adds a BatchNorm output to itself), LeakyReLU activation."""
    gen = torch.Generator().manual_seed(29)
    rows, c, pitch = 512, 64, 96
    x, xg = both(np_, (rows, c), gen, 2.0)
    gamma, beta = torch.rand(c, generator=gen) + 0.5, torch.randn(c, generator=gen)
    xv = fk.val(x)
    sums = torch.cat([xv.sum(0), (xv * xv).sum(0)]).float()
    mean, rstd, var = torch.zeros(c), torch.zeros(c), torch.zeros(c)
    ybuf = rand_planes(np_, (rows, pitch), gen)
    y = ybuf.clone()
    fk.bn_apply_train(x, sums.double(), 1e-5, gamma, beta, y, mean, rstd, var, relu=2, y_pitch=pitch, affine_scale=2.0)
    yg = ybuf.cuda()
    mg, rg, vg = [torch.zeros(c, device="cuda") for _ in range(3)]
    K.bn_apply_train(xg, sums.cuda(), 1e-5, gamma.cuda(), beta.cuda(), yg, mg, rg, vg, relu=2, y_pitch=pitch,
                     affine_scale=2.0)
    check_close("y", fk.val(yg.cpu()), fk.val(y), *tol(np_))
    assert torch.equal(yg.cpu()[..., c:], ybuf[..., c:])          # the trailing channels are not touched
    dybuf, dybufg = both(np_, (rows, pitch), gen)
    dyv = fk.val(dybuf)[:, :c]
    dbeta, dot = dyv.sum(0).float(), (dyv * xv).sum(0).float()
    dga, dbo = torch.full((c,), 0.5, dtype=torch.float64), torch.full((c,), -1.0, dtype=torch.float64)
    dx = torch.zeros_like(x)
    fk.bn_bwd_fused(dybuf, x, mean, rstd, gamma, dot.double(), dbeta.double(), dga, dx, dbeta_out=dbo, out_scale=2.0,
                    dy_pitch=pitch, affine_scale=2.0)
    dgag, dbog = torch.full((c,), 0.5, device="cuda"), torch.full((c,), -1.0, device="cuda")
    dxg = torch.zeros_like(x).cuda()
    K.bn_bwd_fused(dybufg, xg, mean.cuda(), rstd.cuda(), gamma.cuda(), dot.cuda(), dbeta.cuda(), dgag, dxg,
                   dbeta_out=dbog, out_scale=2.0, dy_pitch=pitch, affine_scale=2.0)
    check_close("dgamma", dgag.cpu(), dga, 5e-4, 5.0)
    check_close("dbeta", dbog.cpu(), dbo, 5e-4, 5.0)
    check_close("dx", fk.val(dxg.cpu()), fk.val(dx), *tol(np_))


LN_CASES = [(6, 1, 1, 8192), (5, 4, 4, 512), (3, 32, 32, 64), (2, 64, 64, 32), (4, 8, 8, 8), (7, 16, 16, 24)]


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("case", LN_CASES, ids=["%dx%dx%dx%d" % c for c in LN_CASES])
def test_layer_norm_kernels(K, case, np_):
    """per-sample layer normalisation (utils/ops.py:74-81) forward and backward, incl. the rank-2 case (rows = 1,
    c = 16 * nf features: channel groups do not divide the block) and a channel count whose groups do not divide 256"""
    n, h, w, c = case
    gen = torch.Generator().manual_seed(31)
    shape = (n, h, w, c) if h > 1 else (n, c)
    x, xg = both(np_, shape, gen, 1.5)
    dy, dyg = both(np_, shape, gen)
    gamma, beta = torch.rand(c, generator=gen) + 0.5, torch.randn(c, generator=gen)
    sums = torch.zeros(n, 2, dtype=torch.float64)
    fk.ln_stats(x, sums)
    sg = torch.zeros(n, 2, device="cuda")
    K.ln_stats(xg, sg)
    check_close("sums", sg.cpu(), sums, 2e-5, 1.0)
    y = torch.zeros_like(x)
    fk.ln_apply(x, sums, 1e-12, gamma, beta, y, relu=True)
    yg = torch.zeros_like(x).cuda()
    K.ln_apply(xg, sg, 1e-12, gamma.cuda(), beta.cuda(), yg, relu=True)
    check_close("y", fk.val(yg.cpu()), fk.val(y), *tol(np_))
    ds, dga, dbe = torch.zeros(n, 2, dtype=torch.float64), torch.full((c,), 0.5, dtype=torch.float64), torch.zeros(c, dtype=torch.float64)
    fk.ln_bwd_reduce(dy, x, sums, 1e-12, gamma, ds, dga, dbe)
    dsg, dgag, dbeg = torch.zeros(n, 2, device="cuda"), torch.full((c,), 0.5, device="cuda"), torch.zeros(c, device="cuda")
    K.ln_bwd_reduce(dyg, xg, sg, 1e-12, gamma.cuda(), dsg, dgag, dbeg)
    check_close("dsums", dsg.cpu(), ds, 1e-4, 5.0)
    check_close("dgamma", dgag.cpu(), dga, 1e-4, 5.0)
    check_close("dbeta", dbeg.cpu(), dbe, 1e-4, 5.0)
    dx, dxs = torch.zeros_like(x), torch.zeros(c, dtype=torch.float64)
    fk.ln_bwd_apply(dy, x, sums, 1e-12, gamma, ds, dx, dxs)
    dxg, dxsg = torch.zeros_like(x).cuda(), torch.zeros(c, device="cuda")
    K.ln_bwd_apply(dyg, xg, sg, 1e-12, gamma.cuda(), dsg, dxg, dxsg)
    check_close("dx", fk.val(dxg.cpu()), fk.val(dx), *tol(np_))
    scale = float(fk.val(dx).abs().reshape(-1, c).sum(0).max())
    assert float((dxsg.cpu().double() - dxs).abs().max()) < (2.0 ** -8 if np_ == 1 else 1e-4) * scale + 1e-6


@pytest.mark.parametrize("np_", [1, 2])
def test_resample_blend_and_image_padding_kernels(K, np_):
    gen = torch.Generator().manual_seed(37)
    x, xg = both(np_, (5, 6, 10, 24), gen)
    up = torch.zeros(np_, 5, 12, 20, 24, dtype=torch.bfloat16)
    fk.upscale2x(x, up, 0.25)
    upg = torch.zeros_like(up).cuda()
    K.upscale2x(xg, upg, 0.25)
    check_close("upscale", fk.val(upg.cpu()), fk.val(up), *tol(np_))
    m, mg = both(np_, (5, 12, 20, 24), gen)
    fk.upscale2x(x, up, 0.25, mask=m, mask_kind=fk.MASK_LRELU)
    K.upscale2x(xg, upg, 0.25, mask=mg, mask_kind=K.MASK_LRELU)
    check_close("upscale_masked", fk.val(upg.cpu()), fk.val(up), *tol(np_))
    wide, wideg = both(np_, (5, 6, 10, 40), gen)
    fk.copy_window(x[:, 1:3], 8, wide[:, 1:3], 16, 16)
    K.copy_window(xg[:, 1:3], 8, wideg[:, 1:3], 16, 16)
    assert torch.equal(wideg.cpu(), wide)
    # sample sub-range (views along the sample axis keep the plane stride of the whole buffer)
    po = torch.zeros(np_, 5, 3, 5, 24, dtype=torch.bfloat16)
    fk.pool2x(x[:, 1:4], po[:, 1:4], 0.25)
    pog = torch.zeros_like(po).cuda()
    K.pool2x(xg[:, 1:4], pog[:, 1:4], 0.25)
    check_close("pool", fk.val(pog.cpu()), fk.val(po), *tol(np_))
    z, zg = both(np_, (5, 6, 10, 24), gen)
    ab = torch.tensor([0.3, 0.7])
    out = torch.zeros_like(x)
    fk.axpby(x, z, out, ab)
    og = torch.zeros_like(x).cuda()
    K.axpby(xg, zg, og, ab.cuda())
    check_close("axpby", fk.val(og.cpu()), fk.val(out), *tol(np_))
    fk.axpby(x, None, out, ab[1:])
    K.axpby(xg, None, xg, ab.cuda()[1:])            # in place
    check_close("scale", fk.val(xg.cpu()), fk.val(out), *tol(np_))
    img = torch.rand(3, 8, 8, 3, generator=gen) * 2 - 1
    sc = torch.rand(3, generator=gen) + 0.5
    c8 = torch.zeros(np_, 3, 8, 8, 8, dtype=torch.bfloat16)
    fk.img_to_c8(img, c8, sc)
    c8g = torch.full_like(c8, 3.0).cuda()
    K.img_to_c8(img.cuda(), c8g, sc.cuda())
    check_close("img_to_c8", fk.val(c8g.cpu()), fk.val(c8), 2.0 ** -8 if np_ == 1 else 1e-5, 1e-3)
    back = torch.zeros(3, 8, 8, 3)
    fk.c8_to_img(c8, back)
    bg = torch.zeros(3, 8, 8, 3, device="cuda")
    K.c8_to_img(c8g, bg)
    check_close("c8_to_img", bg.cpu(), back, 1e-6, 1.0)


def test_wgrad_range_major_order_large_operands(K):
    """Operands beyond L2 (100 MB) with two channel tiles: wgrad_gemm runs its work pixel-range-major (all (tap, channel
    tile) items of a range before the next range).  Too large for the fp64 CPU restatement: the checker is torch's fp32
    conv weight gradient on the device (TF32 off) on the same bf16 values."""
    import torch.nn.functional as F
    N, H, W, ci, co = 512, 16, 16, 256, 512
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(1, N, H, W, ci, device="cuda", generator=gen).bfloat16()
    dy = (torch.randn(1, N, H // 2, W // 2, co, device="cuda", generator=gen) * 0.1).bfloat16()
    dw = torch.zeros(16, co, ci, device="cuda")
    K.wgrad_gemm(K.CONV_K4S2, 4, K.View(x), K.View(dy), dw)
    torch.cuda.synchronize()
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        xc = x[0].float().permute(0, 3, 1, 2).contiguous()
        gc = dy[0].float().permute(0, 3, 1, 2).contiguous()
        ref = torch.zeros(co, ci, 4, 4, device="cuda")
        for lo in range(0, N, 64):          # chunks keep cuDNN's workspace small; fp32 accumulation across them
            ref += torch.nn.grad.conv2d_weight(xc[lo:lo + 64], (co, ci, 4, 4), gc[lo:lo + 64], stride=2, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    ref = ref.permute(2, 3, 0, 1).reshape(16, co, ci)          # [tap = kh*4 + kw][co][ci]
    err = float((dw - ref).abs().max() / ref.abs().max())
    print("[wgrad range-major, 100 MB operands] max abs err / max |dw| = %.2e" % err)
    assert err < 1e-4, "range-major wgrad: max abs err / max |dw| = %.3e" % err      # measured 2.05e-5 (fp32 accumulation order)
