"""Tensor-level wrappers over the C ABI: every function takes torch CUDA tensors (used purely as
device-memory containers), builds the C descriptors and enqueues the kernel on torch's current
stream.  ``tests/fake_kernels.py`` restates each function's contract on the CPU so that the host
orchestration can be tested without a GPU; the product never imports it.

Storage convention ("bf16 planes"): a tensor of logical shape S is a torch.bfloat16 tensor of shape
[np, *S]; its value is the sum over the first axis (np = 1: plain bf16; np = 2: hi + lo).
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (ACT_LRELU, ACT_NONE, ACT_RELU, CONV_K4S2, CONV_S1, DECONV_K4S2, MASK_LRELU, MASK_NONE,
                   MASK_RELU)

__all__ = ["View", "CONV_S1", "CONV_K4S2", "DECONV_K4S2", "ACT_NONE", "ACT_LRELU", "ACT_RELU", "MASK_NONE",
           "MASK_LRELU", "MASK_RELU"]


class View:
    """A window of a planes tensor [np, N, H, W, pitch] (a [np, rows, pitch] tensor has H = W = 1):
    samples [n0, n0 + n) and channels [coff, coff + c)."""

    def __init__(self, t, n0=0, n=None, coff=0, c=None):
        assert t.dtype == torch.bfloat16 and t.is_contiguous()
        if t.dim() == 3:
            self.N, self.H, self.W, self.pitch = t.shape[1], 1, 1, t.shape[2]
        else:
            assert t.dim() == 5, t.shape
            self.N, self.H, self.W, self.pitch = t.shape[1:]
        self.t = t
        self.np = t.shape[0]
        self.n0 = n0
        self.n = self.N - n0 if n is None else n
        self.coff = coff
        self.c = self.pitch - coff if c is None else c
        assert 0 <= self.n0 and self.n0 + self.n <= self.N and self.coff + self.c <= self.pitch

    def values(self):
        """fp32 value of the window (test helper; works on any device)."""
        t = self.t.reshape(self.np, self.N, self.H, self.W, self.pitch)
        return t[:, self.n0:self.n0 + self.n, :, :, self.coff:self.coff + self.c].float().sum(0)

    def _act(self):
        a = _lib.Act()
        esz = 2
        a.ptr = self.t.data_ptr() + self.n0 * self.H * self.W * self.pitch * esz
        a.plane_stride = self.t.stride(0)
        a.n, a.h, a.w, a.c = self.n, self.H, self.W, self.c
        a.pitch, a.coff = self.pitch, self.coff
        return a


def _null_act():
    return _lib.Act()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32(t):
    assert t.dtype == torch.float32 and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def _ps(t):
    return t.stride(0)


# Optional per-launch timing of the two GEMM kernels (bench.py): when PROFILE is a list, every
# launch appends [kernel, tag, algorithmic flops, algorithmic HBM bytes, start event, end event].
PROFILE = None


def _taps_per_output(mode, k):
    return k * k if mode == CONV_S1 else 16 if mode == CONV_K4S2 else 4


# T2I_NVTX=1: every GEMM-type launch sits in an NVTX range named after its layer shape, so that
# `ncu --nvtx --print-nvtx-rename kernel` lists launches by layer (profiles/README.md)
NVTX = os.environ.get("T2I_NVTX") == "1"


class _Range:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.tag)

    def __exit__(self, *a):
        if NVTX:
            torch.cuda.nvtx.range_pop()


def _conv_tag(mode, k, x, y, img):
    if img:
        return "conv_gemm img k4s2 %dx%dx%d co%d" % (x.n, x.H, x.W, y.c)
    return "conv_gemm m%d k%d %dx%dx%d ci%d co%d" % (mode, k, x.n, x.H, x.W, x.c, y.c)


def _prof_begin():
    if PROFILE is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _prof_end(ev, kernel, tag, flops, nbytes):
    if ev is not None:
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        PROFILE.append([kernel, tag, flops, nbytes, ev, end])


def img_row_pitch(w):
    """entries per padded bf16 image row: 3 * w values behind 3 zeros, rounded up to a multiple of 4"""
    return (3 * w + 6 + 3) // 4 * 4


def img_to_rows(img, rows, sample_scale=None):
    """fp32 NHWC [n, h, w, 3] (times sample_scale[n]) -> padded bf16 rows, planes [np, n, h, img_row_pitch(w)]"""
    n, h, w, _ = img.shape
    assert rows.dtype == torch.bfloat16 and rows.shape[1:] == (n, h, img_row_pitch(w)) and rows[0].is_contiguous()
    _lib.call("t2i_img_to_rows", _f32(img), n, h, w, _p(sample_scale), _p(rows), _ps(rows), rows.shape[0], _stream())


class ImgPatches:
    """The 4x4 / stride-2 SAME patch matrix of a 3-channel image [n, h, w, 3] as the x operand of conv_gemm / wgrad_img
    (rows = pixels of the [n, h/2, w/2] grid, 48 columns (kh*4 + kw)*3 + c): it exists only in shared memory.  The image
    is given as its padded bf16 rows (img_to_rows), planes [np, n, h, pitch]; slices along n are fine."""

    def __init__(self, rows, w):
        assert rows.dtype == torch.bfloat16 and rows.dim() == 4 and rows.shape[3] == img_row_pitch(w) and rows[0].is_contiguous()
        self.rows = rows
        self.np = rows.shape[0]
        self.n, self.H, self.W, self.c = rows.shape[1], rows.shape[2], w, 48

    def _act(self):
        a = _lib.Act()
        a.n, a.h, a.w, a.c, a.pitch = self.n, self.H, self.W, 48, 48
        a.plane_stride = self.rows.stride(0)
        return a


def conv_gemm(mode, k, flip, x, w, y, bias=None, add=None, mask=None, act=ACT_NONE, mask_kind=MASK_NONE,
              algo_scale=1.0, w_kn=False, stat_sum=None, stat_sq=None, stat_dot=None, stat_x=None, stat_n=0,
              stat_c=0, w_n0=0):
    """y = act(conv(x, w) + bias + add) * mask'(mask); w: planes [np, taps, rows, cols].
    w_n0: y's channels are output channels [w_n0, w_n0 + y.c) of the weight matrix (output-channel window).
    w_kn=False: rows = output channels, cols = contraction (forward use of a layer's packed weights);
    w_kn=True : rows = contraction, cols = output channels (the same weights used for the input-gradient).
    algo_scale: fraction of the contraction that is algorithmic (0.75 for the 48-of-64 patch columns).
    stat_*: fp32 per-channel accumulators fed by the epilogue (+=): sum of the outputs, sum of squares, or
    dot with stat_x (a View on the output grid), over samples [0, stat_n) / channels [0, stat_c) (0 = all)."""
    ev = _prof_begin()
    d = _lib.ConvGemmDesc()
    img = isinstance(x, ImgPatches)
    if img:
        assert mode == CONV_K4S2 and k == 4
        d.x_img = x.rows.data_ptr()
    d.mode, d.k, d.flip, d.np = mode, k, flip, x.np
    d.x = x._act()
    assert w.dtype == torch.bfloat16 and w.dim() == 4 and w[0].is_contiguous() and w.shape[0] == x.np
    d.w = w.data_ptr()
    d.w_plane_stride = w.stride(0)
    d.w_rows, d.w_cols, d.w_layout = w.shape[2], w.shape[3], int(w_kn)
    d.y = y._act()
    d.bias = None if bias is None else bias.data_ptr()
    d.add = add._act() if add is not None else _null_act()
    d.mask = mask._act() if mask is not None else _null_act()
    d.act, d.mask_kind = act, mask_kind
    d.stat_sum = None if stat_sum is None else _f32(stat_sum).value
    d.stat_sq = None if stat_sq is None else _f32(stat_sq).value
    d.stat_dot = None if stat_dot is None else _f32(stat_dot).value
    d.stat_x = stat_x._act() if stat_x is not None else _null_act()
    d.stat_n, d.stat_c = stat_n, stat_c
    d.w_n0 = w_n0
    with _Range(_conv_tag(mode, k, x, y, img)):
        _lib.call("t2i_conv_gemm", C.byref(d), _stream())
    if ev is not None:
        opix = y.n * y.H * y.W
        if img:       # 48 = 16 taps x 3 channels per output pixel; the image is read once as fp32
            flops = 2.0 * opix * y.c * 48 * (3 if x.np == 2 else 1)
            nbytes = 2.0 * x.np * (x.n * x.H * x.W * 3 + opix * y.c * (1 + (add is not None) + (mask is not None)) + y.c * 48)
            _prof_end(ev, "conv_gemm", "img k4s2 %dx%dx%d co%d" % (x.n, x.H, x.W, y.c), flops, nbytes)
            return
        flops = 2.0 * opix * y.c * x.c * _taps_per_output(mode, k) * algo_scale * (3 if x.np == 2 else 1)
        nbytes = 2.0 * x.np * (x.n * x.H * x.W * x.c + opix * y.c * (1 + (add is not None) + (mask is not None))
                               + w.shape[1] * y.c * x.c)
        _prof_end(ev, "conv_gemm", "m%d k%d %dx%dx%d ci%d co%d" % (mode, k, x.n, x.H, x.W, x.c, y.c), flops, nbytes)


def wgrad_gemm(mode, k, x, dy, dw, split_k=0, algo_scale=1.0):
    """dw[tap, co, ci] += sum_pixels dy[pixel, co] * x[pixel + tap, ci]; dw fp32 [taps, cout, cin]."""
    ev = _prof_begin()
    d = _lib.WgradDesc()
    d.mode, d.k, d.np = mode, k, x.np
    d.x, d.dy = x._act(), dy._act()
    assert dw.dtype == torch.float32 and dw.dim() == 3 and dw.is_contiguous()
    d.dw = dw.data_ptr()
    d.cout, d.cin = dw.shape[1], dw.shape[2]
    d.split_k = split_k
    with _Range("wgrad_gemm m%d k%d %dx%dx%d ci%d co%d" % (mode, k, x.n, x.H, x.W, x.c, dy.c)):
        _lib.call("t2i_wgrad_gemm", C.byref(d), _stream())
    if ev is not None:
        vpix = x.n * x.H * x.W if mode != CONV_K4S2 else dy.n * dy.H * dy.W
        taps = k * k if mode == CONV_S1 else 16
        flops = 2.0 * vpix * dy.c * x.c * taps * algo_scale * (3 if x.np == 2 else 1)
        nbytes = 2.0 * x.np * (x.n * x.H * x.W * x.c + dy.n * dy.H * dy.W * dy.c) + 4.0 * taps * dy.c * x.c
        _prof_end(ev, "wgrad_gemm", "m%d k%d %dx%dx%d ci%d co%d" % (mode, k, x.n, x.H, x.W, x.c, dy.c), flops, nbytes)


def wgrad_img(img, other, dw, img_side):
    """Weight gradient between the on-chip patch matrix of img (an ImgPatches) and `other` (a View on the
    [n, h/2, w/2] grid): img_side 1: dw[co, 64] += other^T patch (the image is the conv input); img_side 2:
    dw[64, ci] += patch^T other (the image is the gradient at a transposed conv's output).  dw: fp32 [1, rows, cols]."""
    ev = _prof_begin()
    n, h, w = img.n, img.H, img.W
    assert dw.dtype == torch.float32 and dw.dim() == 3 and dw.shape[0] == 1 and dw.is_contiguous() and img.np == other.np
    a = other._act()
    with _Range("wgrad_img side%d %dx%dx%d c%d" % (img_side, n, h, w, other.c)):
        _lib.call("t2i_wgrad_img", _p(img.rows), _ps(img.rows), n, h, w, C.byref(a), img_side, other.np, _f32(dw), dw.shape[1],
                  dw.shape[2], _stream())
    if ev is not None:
        pix = n * (h // 2) * (w // 2)
        flops = 2.0 * pix * other.c * 48 * (3 if other.np == 2 else 1)
        nbytes = 2.0 * other.np * (n * h * w * 3 + pix * other.c) + 4.0 * other.c * 48
        _prof_end(ev, "wgrad_gemm", "img side%d %dx%dx%d c%d" % (img_side, n, h, w, other.c), flops, nbytes)


def deconv_img(a, w, out, bias3=None, w_kn=False, w9=None, b9=None, img=None):
    """out (fp32 NHWC [n, 2h, 2w, 3]) = bias3 + transposed 4x4/s2 conv of a (View [n, h, w, c]) with w (planes
    [np, 1, 64, K] or, w_kn, [np, 1, K, 64]); with w9 / b9 / img also img = tanh(conv3x3(out) + b9)."""
    ev = _prof_begin()
    assert w.dtype == torch.bfloat16 and w.dim() == 4 and w.shape[1] == 1 and w[0].is_contiguous() and w.shape[0] == a.np
    act = a._act()
    with _Range("deconv_img %dx%dx%d c%d%s" % (a.n, a.H, a.W, a.c, " +c9" if img is not None else "")):
        _lib.call("t2i_deconv_img", C.byref(act), _p(w), w.stride(0), w.shape[2], w.shape[3], int(w_kn), a.np, _p(bias3),
                  _f32(out), _p(w9), _p(b9), _p(img), _stream())
    if ev is not None:
        pix = a.n * a.H * a.W
        flops = 2.0 * pix * a.c * 48 * (3 if a.np == 2 else 1)
        nbytes = 2.0 * a.np * pix * a.c + 4.0 * pix * 4 * 3 * (2 if img is not None else 1)
        _prof_end(ev, "deconv_img", "%dx%dx%d c%d%s" % (a.n, a.H, a.W, a.c, " +c9" if img is not None else ""), flops, nbytes)


def dense_f32(x, w, bias, y, act=ACT_NONE):
    """y = act(x @ w^T + bias), all fp32; w [cout, cin]."""
    rows, cin = x.shape
    cout = w.shape[0]
    assert w.shape[1] == cin and y.shape == (rows, cout) and w.is_contiguous()
    _lib.call("t2i_dense_f32", _f32(x), rows, cin, _f32(w), _p(bias), cout, act, _f32(y), _stream())


def scale_rows(src, row_scale, dst):
    rows = src.shape[0]
    _lib.call("t2i_scale_rows", _f32(src), _f32(row_scale), _f32(dst), rows, src.numel() // rows, _stream())


def to_planes(src, dst, row_scale=None):
    rows = src.shape[0]
    cols = src.numel() // rows
    _lib.call("t2i_to_planes", _f32(src), _p(dst), _ps(dst), dst.shape[0], rows, cols, _p(row_scale), _stream())


def from_planes(src, dst):
    _lib.call("t2i_from_planes", _p(src), _ps(src), src.shape[0], _f32(dst), dst.numel(), _stream())


def im2col_k4s2_c3(img, col, sample_scale=None):
    n, h, w, _ = img.shape
    _lib.call("t2i_im2col_k4s2_c3", _f32(img), n, h, w, _p(sample_scale), _p(col), _ps(col), col.shape[0], _stream())


def col2im_k4s2_c3(col, img, bias3=None):
    n, h, w, _ = img.shape
    _lib.call("t2i_col2im_k4s2_c3", _p(col), _ps(col), col.shape[0], n, h, w, _p(bias3), _f32(img), _stream())


def im2col_k3s1_c3(img, col):
    """fp32 NHWC 3-channel image -> planes [np, n*h*w, 32] of 3x3 / stride-1 SAME patches (27 columns used)"""
    n, h, w, _ = img.shape
    _lib.call("t2i_im2col_k3s1_c3", _f32(img), n, h, w, _p(col), _ps(col), col.shape[0], _stream())


def tanh_c3_fwd(logits8, img):
    """planes [np, ..., 8] (3 channels used) -> tanh -> fp32 [..., 3]"""
    _lib.call("t2i_tanh_c3_fwd", _p(logits8), _ps(logits8), logits8.shape[0], _f32(img), img.numel() // 3, _stream())


def tanh_c3_bwd(img, dimg, dlogits8):
    _lib.call("t2i_tanh_c3_bwd", _f32(img), _f32(dimg), _p(dlogits8), _ps(dlogits8), dlogits8.shape[0], img.numel() // 3,
              _stream())


def conv3x3_c3_tanh_fwd(x, w, b, y):
    n, h, wd, _ = x.shape
    _lib.call("t2i_conv3x3_c3_tanh_fwd", _f32(x), _f32(w), _f32(b), _f32(y), n, h, wd, _stream())


def conv3x3_c3_tanh_bwd(x, w, y, dy, dx, dw, db, dx_sum=None):
    """dx (+ dx_sum) and / or dw, db of tanh(conv3x3(x) + b): pass dx=None or dw=db=None to get one of the two kernels
    alone (the input gradient is on the backward pass's critical path, the weight gradient is not)."""
    n, h, wd, _ = x.shape
    for t in (x, w, y, dy, dx, dw, db, dx_sum):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    _lib.call("t2i_conv3x3_c3_tanh_bwd", _f32(x), _f32(w), _f32(y), _f32(dy), _p(dx), _p(dw), _p(db),
              _p(dx_sum), n, h, wd, _stream())


def colsum(src, out):
    """out[c] += sum over all pixels of the view (bias gradients)."""
    a = src._act()
    rows = src.n * src.H * src.W
    _lib.call("t2i_colsum", C.c_void_p(a.ptr), a.plane_stride, src.np, rows, src.c, src.pitch, src.coff, _f32(out),
              _stream())


def _rows_c(t):
    c = t.shape[-1]
    return t[0].numel() // c, c


def bn_stats(x, sums, mean, rstd, var, eps):
    """sums: fp32 scratch [2c], zero on entry, left zero on return."""
    rows, c = _rows_c(x)
    _lib.call("t2i_bn_stats", _p(x), _ps(x), x.shape[0], rows, c, _f32(sums), _f32(mean), _f32(rstd), _f32(var), eps,
              _stream())


def bn_apply(x, mean, rstd, gamma, beta, y, residual=None, relu=False):
    rows, c = _rows_c(x)
    _lib.call("t2i_bn_apply", _p(x), _ps(x), _f32(mean), _f32(rstd), _f32(gamma), _f32(beta), _p(residual),
              0 if residual is None else _ps(residual), _p(y), _ps(y), x.shape[0], rows, c, int(relu), _stream())


def bn_apply_train(x, sums, eps, gamma, beta, y, mean, rstd, var, residual=None, relu=False, moving=None,
                   decay=0.9, stat_rows=0, y_pitch=0, affine_scale=1.0):
    """sums: fp32 [2c] = [sum x | sum x^2] from conv_gemm(stat_sum=, stat_sq=); moving: (mm, mv) or None;
    relu: False / True (ReLU) / 2 (LeakyReLU 0.2); stat_rows: values per channel behind the sums when they were
    all-reduced over ranks (0 = this tensor's rows); y_pitch: y is the leading c channels of a [.., y_pitch] buffer."""
    rows, c = _rows_c(x)
    mm, mv = moving if moving is not None else (None, None)
    _lib.call("t2i_bn_apply_train", _p(x), _ps(x), _f32(sums), eps, _f32(gamma), _f32(beta), _p(residual),
              0 if residual is None else _ps(residual), _p(y), _ps(y), x.shape[0], rows, c, int(relu), _f32(mean),
              _f32(rstd), _f32(var), _p(mm), _p(mv), decay, stat_rows, y_pitch, affine_scale, _stream())


def bn_bwd_fused(dy, x, mean, rstd, gamma, dot, dbeta, dgamma, dx, dx_sum=None, dbeta_out=None, out_scale=1.0,
                 dot_normalised=False, stat_rows=0, dy_pitch=0, affine_scale=1.0):
    """dbeta / dot: the reductions conv_gemm(stat_sum=dbeta, stat_dot=dot, stat_x=x) produced with dy
    (dot_normalised: dot = sum dy * xhat, as bn_bwd_reduce writes it).  dgamma += out_scale * (...),
    dbeta_out += out_scale * dbeta (when the sums live in a scratch that was all-reduced)."""
    rows, c = _rows_c(x)
    _lib.call("t2i_bn_bwd_fused", _p(dy), _ps(dy), _p(x), _ps(x), _f32(mean), _f32(rstd), _f32(gamma), _f32(dot),
              _f32(dbeta), _f32(dgamma), _p(dbeta_out), out_scale, int(dot_normalised), _p(dx), _ps(dx), _p(dx_sum),
              x.shape[0], rows, c, stat_rows, dy_pitch, affine_scale, _stream())


def ce_seeds(logit, n, label, weight, inv_global_batch, seed, loss_sum):
    _lib.call("t2i_ce_seeds", _f32(logit), n, label, weight, inv_global_batch, _f32(seed), _p(loss_sum), _stream())


def s1_scalars(sums, scalars, global_batch, ce, alpha, kl_coeff, which):
    _lib.call("t2i_s1_scalars", _f32(sums), _f32(scalars), global_batch, ce, alpha, kl_coeff, which, _stream())


def bn_bwd_reduce(dy, x, mean, rstd, dgamma, dbeta):
    rows, c = _rows_c(x)
    _lib.call("t2i_bn_bwd_reduce", _p(dy), _ps(dy), _p(x), _ps(x), _f32(mean), _f32(rstd), x.shape[0], rows, c,
              _f32(dgamma), _f32(dbeta), _stream())


def bn_bwd_apply(dy, x, mean, rstd, gamma, dgamma, dbeta, dx):
    rows, c = _rows_c(x)
    _lib.call("t2i_bn_bwd_apply", _p(dy), _ps(dy), _p(x), _ps(x), _f32(mean), _f32(rstd), _f32(gamma), _f32(dgamma),
              _f32(dbeta), _p(dx), _ps(dx), x.shape[0], rows, c, _stream())


def bn_update_moving(mm, mv, mean, var, rows, decay):
    _lib.call("t2i_bn_update_moving", _f32(mm), _f32(mv), _f32(mean), _f32(var), rows, mm.numel(), decay, _stream())


def act_bwd(dy, y, dst, mask_kind):
    _lib.call("t2i_act_bwd", _p(dy), _ps(dy), _p(y), _ps(y), _p(dst), _ps(dst), dy.shape[0], dy[0].numel(), mask_kind,
              _stream())


def embed_tile(e, cat, coff):
    """e planes [np, s, c] -> channels [coff, coff + c) of cat planes [np, s, h, w, pitch]."""
    s, c = e.shape[1], e.shape[2]
    hw = cat.shape[2] * cat.shape[3]
    _lib.call("t2i_embed_tile", _p(e), _ps(e), _p(cat), _ps(cat), e.shape[0], s, c, cat.shape[4], coff, hw, _stream())


def embed_reduce(dcat, de, coff):
    s, c = de.shape[1], de.shape[2]
    hw = dcat.shape[2] * dcat.shape[3]
    _lib.call("t2i_embed_reduce", _p(dcat), _ps(dcat), _p(de), _ps(de), de.shape[0], s, c, dcat.shape[4], coff, hw,
              _stream())


def dout_fwd(a, w, b, logit):
    s = a.shape[1]
    _lib.call("t2i_dout_fwd", _p(a), _ps(a), a.shape[0], _f32(w), _f32(b), _f32(logit), s, a[0, 0].numel(), _stream())


def dout_bwd_data(a, w, seed, da):
    s = a.shape[1]
    _lib.call("t2i_dout_bwd_data", _p(a), _ps(a), a.shape[0], _f32(w), _f32(seed), _p(da), _ps(da), s,
              a[0, 0].numel(), _stream())


def dout_bwd_weight(a, seed, dw, db, s_bias):
    s = a.shape[1]
    _lib.call("t2i_dout_bwd_weight", _p(a), _ps(a), a.shape[0], _f32(seed), _f32(dw), _p(db), s, s_bias,
              a[0, 0].numel(), _stream())


def gp_interp(g, x, eps, xhat):
    n = g.shape[0]
    _lib.call("t2i_gp_interp", _f32(g), _f32(x), _f32(eps), _f32(xhat), n, g[0].numel(), _stream())


def gp_penalty(grad, weight, inv_global_batch, slope, coef, pen_sum):
    n = grad.shape[0]
    _lib.call("t2i_gp_penalty", _f32(grad), n, grad[0].numel(), weight, inv_global_batch, _f32(slope), _f32(coef),
              _f32(pen_sum), _stream())


def ca_fwd(ms, z, tn_eps, zc, kl_sum):
    """ms: fp32 [b, 2*ce] = [mean | log_sigma] (dense_f32)"""
    b, z_dim = z.shape
    ce = tn_eps.shape[1]
    _lib.call("t2i_ca_fwd", _f32(ms), _f32(z), _f32(tn_eps), _p(zc), _ps(zc), zc.shape[0], b, z_dim, ce,
              _p(kl_sum), _stream())


def ca_bwd(ms, dzc, tn_eps, dms, z_dim, kl_scale):
    b, ce = tn_eps.shape
    _lib.call("t2i_ca_bwd", _f32(ms), _p(dzc), _ps(dzc), _f32(tn_eps), _p(dms), _ps(dms), dms.shape[0], b, z_dim,
              ce, kl_scale, _stream())


def d_seeds(kt, seed, b, inv_global_batch):
    _lib.call("t2i_d_seeds", _f32(kt), _f32(seed), b, inv_global_batch, _stream())


def d_sums(logit, b, sums):
    _lib.call("t2i_d_sums", _f32(logit), b, _f32(sums), _stream())


def d_scalars(sums, kt, scalars, global_batch, gp_weight, kt_lr):
    _lib.call("t2i_d_scalars", _f32(sums), _f32(kt), _f32(scalars), global_batch, gp_weight, kt_lr, _stream())


def g_sums(logit_fake, b, sums):
    _lib.call("t2i_g_sums", _f32(logit_fake), b, _f32(sums), _stream())


def g_scalars(sums, scalars, global_batch, ce, kl_coeff):
    _lib.call("t2i_g_scalars", _f32(sums), _f32(scalars), global_batch, ce, kl_coeff, _stream())


def pack_weight(w, fwd=None, bwd=None):
    """fp32 master [taps, cout, cin] -> planes fwd [np, taps, cout, cin] and bwd [np, taps, cin, cout]."""
    taps, cout, cin = w.shape
    ref = fwd if fwd is not None else bwd
    _lib.call("t2i_pack_weight", _f32(w), taps, cout, cin, _p(fwd), 0 if fwd is None else _ps(fwd), _p(bwd),
              0 if bwd is None else _ps(bwd), ref.shape[0], _stream())


def adam_tf(theta, grad, m, v, lr_t, beta1, beta2, eps=1e-8, grad_scale=1.0, packed=None):
    """lr_t: fp32 DEVICE tensor [1]; packed: optional bf16 planes [np, n] receiving the updated weights."""
    _lib.call("t2i_adam_tf", _f32(theta), _f32(grad), _f32(m), _f32(v), theta.numel(), _f32(lr_t), beta1, beta2, eps,
              grad_scale, _p(packed), 0 if packed is None else _ps(packed), 1 if packed is None else packed.shape[0],
              _stream())


# ---------------------------------------------------------------------------------------------------------------------
# conditional PGGAN (models/pggan/pggan.py): planes tensors [np, n, h, w, c] (or [np, n, c]); slices along the sample
# axis are fine (the plane stride travels separately)
def _nrc(t):
    """[np, n, ..., c] -> (n, rows per sample, c)"""
    n, c = t.shape[1], t.shape[-1]
    return n, t[0, 0].numel() // c, c


def ln_stats(x, sums):
    """sums fp32 [n, 2] += [sum x | sum x^2] per sample (layer_norm, utils/ops.py:74-81); the caller zeroes sums."""
    n, rows, c = _nrc(x)
    _lib.call("t2i_ln_stats", _p(x), _ps(x), x.shape[0], n, rows * c, _f32(sums), _stream())


def ln_apply(x, sums, eps, gamma, beta, y, relu=False):
    n, rows, c = _nrc(x)
    _lib.call("t2i_ln_apply", _p(x), _ps(x), _f32(sums), eps, _f32(gamma), _f32(beta), _p(y), _ps(y), x.shape[0], n, rows,
              c, int(relu), _stream())


def ln_bwd_reduce(dy, x, sums, eps, gamma, dsums, dgamma, dbeta):
    n, rows, c = _nrc(x)
    _lib.call("t2i_ln_bwd_reduce", _p(dy), _ps(dy), _p(x), _ps(x), _f32(sums), eps, _f32(gamma), _f32(dsums),
              _f32(dgamma), _f32(dbeta), x.shape[0], n, rows, c, _stream())


def ln_bwd_apply(dy, x, sums, eps, gamma, dsums, dx, dx_sum=None):
    n, rows, c = _nrc(x)
    _lib.call("t2i_ln_bwd_apply", _p(dy), _ps(dy), _p(x), _ps(x), _f32(sums), eps, _f32(gamma), _f32(dsums), _p(dx),
              _ps(dx), _p(dx_sum), x.shape[0], n, rows, c, _stream())


def upscale2x(x, y, scale=1.0, mask=None, mask_kind=MASK_NONE):
    """y[n, i, j] = scale * x[n, i // 2, j // 2] (resize_nearest_neighbor x2, utils/ops.py:109-111), optionally times
    act'(mask) (mask: post-activation values shaped like y)"""
    _, n, h, w, c = x.shape
    _lib.call("t2i_upscale2x", _p(x), _ps(x), _p(y), _ps(y), x.shape[0], n, h, w, c, scale, _p(mask),
              0 if mask is None else _ps(mask), mask_kind if mask is not None else MASK_NONE, _stream())


def copy_window(src, s_coff, dst, d_coff, c):
    """dst[..., d_coff : d_coff + c] = src[..., s_coff : s_coff + c] (planes [np, n, ..., pitch])"""
    rows = src[0].numel() // src.shape[-1]
    assert rows == dst[0].numel() // dst.shape[-1]
    _lib.call("t2i_copy_window", _p(src), _ps(src), src.shape[-1], s_coff, _p(dst), _ps(dst), dst.shape[-1], d_coff,
              src.shape[0], rows, c, _stream())


def pool2x(x, y, scale=0.25):
    """y = scale * sum over 2x2 blocks (scale 1/4: tf.nn.pool AVG 2, utils/ops.py:100-101)"""
    _, n, h, w, c = x.shape
    _lib.call("t2i_pool2x", _p(x), _ps(x), _p(y), _ps(y), x.shape[0], n, h, w, c, scale, _stream())


def axpby(x, z, out, ab):
    """out = ab[0] * x + ab[1] * z (z None: out = ab[0] * x); ab: fp32 DEVICE tensor"""
    _lib.call("t2i_axpby", _p(x), _ps(x), _p(z), 0 if z is None else _ps(z), _p(out), _ps(out), x.shape[0], x[0].numel(),
              _f32(ab), _stream())


def img_to_c8(img, dst, sample_scale=None):
    """fp32 NHWC [n, h, w, 3] -> planes [np, n, h, w, 8] (channels 3..7 zero), optionally scaled per sample"""
    n = img.shape[0]
    _lib.call("t2i_img_to_c8", _f32(img), n, img[0].numel() // 3, _p(sample_scale), _p(dst), _ps(dst), dst.shape[0],
              _stream())


def c8_to_img(src, img):
    _lib.call("t2i_c8_to_img", _p(src), _ps(src), src.shape[0], _f32(img), img.numel() // 3, _stream())
