"""-m gpu: the direct kernels of the 3-channel ends (csrc/img_gemm.cu, the x_img form of t2i_conv_gemm, t2i_wgrad_img)
and the fp32 conditioning head, through the C ABI, against their CPU restatements (tests/fake_kernels.py, fp64 on the
same inputs).  Tolerances as in test_kernels_gpu.py: bf16 outputs 2^-7 relative (np = 1) / 1e-4 (np = 2); fp32 outputs
of bf16-operand products 2e-3 (np = 1: the operands the tensor core sees are the bf16 roundings of the fp32 image, the
restatement uses the unrounded image) / 2e-5 (np = 2) of the tensor's scale."""
import pytest
import torch

import fake_kernels as fk
from test_kernels_gpu import check_close, rand_planes, tol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from t2i_b200 import kernels
    return kernels


def f32tol(np_):
    return (4e-3, 1.0) if np_ == 1 else (2e-5, 1.0)


def rows_of(K, img, np_, scale=None):
    """the image's padded bf16 rows, by the CPU restatement and by the CUDA kernel (bit-identical, asserted)"""
    n, h, w, _ = img.shape
    r = torch.zeros(np_, n, h, fk.img_row_pitch(w), dtype=torch.bfloat16)
    fk.img_to_rows(img, r, scale)
    rg = torch.full_like(r, 3.0).cuda()
    K.img_to_rows(img.cuda(), rg, None if scale is None else scale.cuda())
    torch.cuda.synchronize()
    assert torch.equal(rg.cpu(), r)
    return r, rg


IMG_CASES = [
    # name, N, H, W, Cout
    ("d_h0", 3, 64, 64, 128),
    ("d_h0_tiny", 5, 64, 64, 8),
    ("wide_256", 1, 256, 256, 64),
    ("small_32", 2, 32, 32, 128),
    ("co_256", 2, 64, 64, 256),
]


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("case", IMG_CASES, ids=[c[0] for c in IMG_CASES])
def test_conv_from_image_forward_and_tangent(K, case, np_):
    """y = lrelu(conv4x4s2(img) + b) with the patches formed on chip; then the tangent form (no bias, derivative mask
    taken from the output buffer in place), as the gradient-penalty schedule uses it (engine.d_forward)."""
    name, N, H, W, Co = case
    gen = torch.Generator().manual_seed(len(name))
    img = torch.rand(N, H, W, 3, generator=gen) * 2 - 1
    w = rand_planes(np_, (1, Co, 64), gen, scale=48 ** -0.5)
    w[:, :, :, 48:] = 0
    bias = torch.randn(Co, generator=gen) * 0.1
    y = torch.zeros(np_, N, H // 2, W // 2, Co, dtype=torch.bfloat16)
    rows, rows_g = rows_of(K, img, np_)
    fk.conv_gemm(fk.CONV_K4S2, 4, 0, fk.ImgPatches(rows, W), w, fk.View(y), bias=bias, act=fk.ACT_LRELU)
    wg = w.cuda()
    yg = torch.full_like(y, 3.0).cuda()
    K.conv_gemm(K.CONV_K4S2, 4, 0, K.ImgPatches(rows_g, W), wg, K.View(yg), bias=bias.cuda(), act=K.ACT_LRELU)
    torch.cuda.synchronize()
    check_close(name, fk.val(yg.cpu()), fk.val(y), *tol(np_))
    # tangent pass over a sample sub-range, in place over the forward activations
    n0, n = (1, N - 1) if N > 1 else (0, 1)
    timg = torch.randn(N, H, W, 3, generator=gen)
    coef = torch.rand(N, generator=gen) + 0.5
    trows, trows_g = rows_of(K, timg, np_, coef)           # per-sample scaled, as the penalty's tangent seeds are
    y2 = yg.cpu().clone()
    fk.conv_gemm(fk.CONV_K4S2, 4, 0, fk.ImgPatches(trows[:, n0:n0 + n], W), w, fk.View(y2, n0, n), mask=fk.View(y2, n0, n),
                 mask_kind=fk.MASK_LRELU)
    K.conv_gemm(K.CONV_K4S2, 4, 0, K.ImgPatches(trows_g[:, n0:n0 + n], W), wg, K.View(yg, n0, n), mask=K.View(yg, n0, n),
                mask_kind=K.MASK_LRELU)
    torch.cuda.synchronize()
    check_close(name + "_tangent", fk.val(yg.cpu()), fk.val(y2), *tol(np_))


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("cin", [128, 8])
def test_conv_from_image_kn_with_mask_and_bn_reductions(K, np_, cin):
    """the input gradient of g_net's last transposed conv (engine.g_backward): patches of d_u4 times the layer's packed
    forward weights read [contraction][channel], ReLU mask, and the two BatchNorm-backward reductions in the epilogue"""
    gen = torch.Generator().manual_seed(17 + cin)
    N, H, W = 4, 64, 64
    dimg = torch.randn(N, H, W, 3, generator=gen)
    w = rand_planes(np_, (1, 64, cin), gen, scale=48 ** -0.5)
    w[:, :, 48:, :] = 0
    mask = rand_planes(np_, (N, 32, 32, cin), gen)
    xpre = rand_planes(np_, (N, 32, 32, cin), gen)
    y = torch.zeros(np_, N, 32, 32, cin, dtype=torch.bfloat16)
    s_sum, s_dot = torch.zeros(cin, dtype=torch.float64), torch.zeros(cin, dtype=torch.float64)
    rows, rows_g = rows_of(K, dimg, np_)
    fk.conv_gemm(fk.CONV_K4S2, 4, 0, fk.ImgPatches(rows, W), w, fk.View(y), w_kn=True, mask=fk.View(mask), mask_kind=fk.MASK_RELU,
                 stat_sum=s_sum, stat_dot=s_dot, stat_x=fk.View(xpre))
    yg = torch.zeros_like(y).cuda()
    g_sum, g_dot = torch.zeros(cin, device="cuda"), torch.zeros(cin, device="cuda")
    K.conv_gemm(K.CONV_K4S2, 4, 0, K.ImgPatches(rows_g, W), w.cuda(), K.View(yg), w_kn=True, mask=K.View(mask.cuda()),
                mask_kind=K.MASK_RELU, stat_sum=g_sum, stat_dot=g_dot, stat_x=K.View(xpre.cuda()))
    torch.cuda.synchronize()
    check_close("dgrad", fk.val(yg.cpu()), fk.val(y), *tol(np_))
    check_close("stat_sum", g_sum, s_sum, *f32tol(np_))
    check_close("stat_dot", g_dot, s_dot, *f32tol(np_))


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("case", [("b6_c128", 6, 128), ("b300_c128", 300, 128), ("b5_c8", 5, 8), ("b3_c64", 3, 64)],
                         ids=lambda c: c[0])
def test_deconv_img_input_gradient_and_fused_generator_end(K, case, np_):
    name, N, C = case
    gen = torch.Generator().manual_seed(N + C)
    a = rand_planes(np_, (N, 32, 32, C), gen)
    # (a) d_net's first conv, input gradient: weights [co = C][64] used [contraction][column]
    w_kn = rand_planes(np_, (1, C, 64), gen, scale=C ** -0.5)
    w_kn[:, :, :, 48:] = 0
    out = torch.zeros(N, 64, 64, 3, dtype=torch.float64)
    fk.deconv_img(fk.View(a), w_kn, out, w_kn=True)
    og = torch.full((N, 64, 64, 3), 9.0, device="cuda")
    K.deconv_img(K.View(a.cuda()), w_kn.cuda(), og, w_kn=True)
    torch.cuda.synchronize()
    check_close(name + "_gx", og, out, *f32tol(np_))
    # (b) g_net's end: transposed conv [64][C] + bias, then 3x3 conv + tanh
    w_nk = rand_planes(np_, (1, 64, C), gen, scale=(4 * C) ** -0.5)
    w_nk[:, :, 48:, :] = 0
    b3 = torch.randn(3, generator=gen) * 0.1
    w9 = torch.randn(81, generator=gen) * 0.3
    b9 = torch.randn(3, generator=gen) * 0.1
    u4, img = torch.zeros(N, 64, 64, 3, dtype=torch.float64), torch.zeros(N, 64, 64, 3, dtype=torch.float64)
    fk.deconv_img(fk.View(a), w_nk, u4, bias3=b3, w9=w9, b9=b9, img=img)
    ug, ig = torch.full((N, 64, 64, 3), 9.0, device="cuda"), torch.full((N, 64, 64, 3), 9.0, device="cuda")
    K.deconv_img(K.View(a.cuda()), w_nk.cuda(), ug, bias3=b3.cuda(), w9=w9.cuda(), b9=b9.cuda(), img=ig)
    torch.cuda.synchronize()
    check_close(name + "_u4", ug, u4, *f32tol(np_))
    check_close(name + "_img", ig, img, *f32tol(np_))


@pytest.mark.parametrize("np_", [1, 2])
def test_deconv_img_other_extents_and_sample_window(K, np_):
    """128-wide maps (one patch row per tile), 16x16 maps (eight rows per tile), a sample sub-range of a larger buffer"""
    gen = torch.Generator().manual_seed(3)
    for (N, P, Q, C) in ((2, 128, 128, 64), (3, 16, 16, 128), (2, 8, 16, 32)):
        a = rand_planes(np_, (N + 2, P, Q, C), gen)
        w = rand_planes(np_, (1, C, 64), gen, scale=C ** -0.5)
        w[:, :, :, 48:] = 0
        out = torch.zeros(N, 2 * P, 2 * Q, 3, dtype=torch.float64)
        fk.deconv_img(fk.View(a, 1, N), w, out, w_kn=True)
        og = torch.zeros(N, 2 * P, 2 * Q, 3, device="cuda")
        K.deconv_img(K.View(a.cuda(), 1, N), w.cuda(), og, w_kn=True)
        torch.cuda.synchronize()
        check_close("deconv_%dx%d" % (P, Q), og, out, *f32tol(np_))


@pytest.mark.parametrize("np_", [1, 2])
@pytest.mark.parametrize("case", [("h0_4b", 1, 12, 64, 64, 128), ("h0_tiny", 1, 4, 64, 64, 8), ("h0_256", 1, 2, 256, 256, 64),
                                  ("t3", 2, 6, 64, 64, 128), ("t3_tiny", 2, 4, 64, 64, 8), ("t3_c256", 2, 3, 64, 64, 256)],
                         ids=lambda c: c[0])
def test_wgrad_img(K, case, np_):
    name, side, N, H, W, C = case
    gen = torch.Generator().manual_seed(N * 7 + C)
    img = torch.randn(N, H, W, 3, generator=gen)
    other = rand_planes(np_, (N, H // 2, W // 2, C), gen)
    shape = (1, C, 64) if side == 1 else (1, 64, C)
    dw = torch.zeros(shape, dtype=torch.float64)
    rows, rows_g = rows_of(K, img, np_)
    fk.wgrad_img(fk.ImgPatches(rows, W), fk.View(other), dw, side)
    dg = torch.zeros(shape, device="cuda")
    K.wgrad_img(K.ImgPatches(rows_g, W), K.View(other.cuda()), dg, side)
    torch.cuda.synchronize()
    check_close(name, dg, dw, *f32tol(np_))
    # accumulates (+=) into what is there
    K.wgrad_img(K.ImgPatches(rows_g, W), K.View(other.cuda()), dg, side)
    torch.cuda.synchronize()
    check_close(name + "_acc", dg, 2 * dw, *f32tol(np_))


@pytest.mark.parametrize("shape", [(256, 1024, 256), (16, 1024, 256), (4, 32, 16), (37, 40, 24)])
def test_dense_f32_and_scale_rows(K, shape):
    rows, cin, cout = shape
    gen = torch.Generator().manual_seed(rows)
    x, w, b = torch.randn(rows, cin, generator=gen), torch.randn(cout, cin, generator=gen) * cin ** -0.5, torch.randn(cout, generator=gen)
    y = torch.zeros(rows, cout, dtype=torch.float64)
    fk.dense_f32(x, w, b, y, act=fk.ACT_LRELU)
    yg = torch.zeros(rows, cout, device="cuda")
    K.dense_f32(x.cuda(), w.cuda(), b.cuda(), yg, act=K.ACT_LRELU)
    torch.cuda.synchronize()
    check_close("dense_f32", yg, y, 1e-5, 1.0)
    src, sc = torch.randn(rows, 8, 4, 3, generator=gen), torch.randn(rows, generator=gen)
    dst = torch.zeros(rows, 8, 4, 3, device="cuda")
    K.scale_rows(src.cuda(), sc.cuda(), dst)
    assert torch.equal(dst.cpu(), src * sc.view(-1, 1, 1, 1))
