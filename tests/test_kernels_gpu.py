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


WGRAD_CASES = [
    ("fc", fk.CONV_S1, 1, 256, 1, 1, 128, 128),
    ("fc_ragged", fk.CONV_S1, 1, 200, 1, 1, 256, 384),
    ("c1_4x4", fk.CONV_S1, 1, 16, 4, 4, 128, 64),
    ("c3_4x4", fk.CONV_S1, 3, 16, 4, 4, 128, 128),
    ("c3_16x16", fk.CONV_S1, 3, 3, 16, 16, 64, 128),
    ("c3_tiny", fk.CONV_S1, 3, 4, 4, 4, 8, 16),
    ("k4s2_32", fk.CONV_K4S2, 4, 2, 32, 32, 128, 128),
    ("k4s2_8", fk.CONV_K4S2, 4, 16, 8, 8, 64, 256),
    ("deconv_4", fk.DECONV_K4S2, 4, 16, 4, 4, 128, 64),
    ("deconv_16", fk.DECONV_K4S2, 4, 2, 16, 16, 256, 128),
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
    K.wgrad_gemm(mode, k, K.View(x.cuda()), K.View(dy.cuda()), dwg, split_k=3)
    torch.cuda.synchronize()
    check_close(name + "_acc", dwg.cpu(), 2 * dw, 2e-5 if np_ == 2 else 1e-4, 10.0)
