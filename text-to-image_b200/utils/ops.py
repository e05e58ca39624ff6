"""The op wrappers of the reference's ``utils/ops.py`` (conv2d :58-63, conv2d_transpose :66-71, batch_norm :7-29,
fc :84-87, to_nchw/to_nhwc :129-134, lrelu_act :90-91 of the wgancls / StackGAN paths; layer_norm :74-81, pool :100-101,
resize_nearest_neighbor / upscale :104-111 of the PGGAN path) with the same names, arguments and defaults, executed
eagerly by the CUDA library.

Like the TF originals each wrapper owns its variables: they are created on first use under the
current ``variable_scope`` with TF's default names (``Conv``, ``Conv_1``, ``Conv2d_transpose``,
``dense``, ``BatchNorm`` + ``/weights``, ``/biases``, ``/kernel``, ``/bias``, ``/gamma`` ...) in TF
layouts, and shared when the scope is re-entered with ``reuse=True`` (models/wgancls/model.py:134,168).
Inputs/outputs are fp32 CUDA tensors in the requested data format.  The fused training step
(``t2i_b200.engine``) does not go through this module; it exists so that code written against the
reference's op surface (the other models' graphs) keeps working on the same kernels.

Supported geometries (everything the paths use): k1/k2/k3/k4 stride 1 SAME (TF's asymmetric padding for even k; k1
also VALID), k4 stride 2 SAME, k4 VALID on a 4x4 input (stride 4 or 1: one output pixel), transposed k4 stride 2 SAME,
2x2 average pool, 2x nearest-neighbour upscale.  Other shapes raise ValueError, as TF
raises on invalid arguments; channel counts that are not multiples of 8 are zero-padded internally.
"""
import contextlib
import math
from collections import OrderedDict

import torch

from .. import kernels as K

NHWC = 'NHWC'
NCHW = 'NCHW'

PRECISION_PLANES = 2        # eager ops default to the parity representation (split bf16)


class _State:
    def __init__(self):
        self.variables = OrderedDict()      # full name -> fp32 CUDA tensor (TF layout)
        self.scope = []                     # [(name, reuse, counters)]
        self.seed = 0
        self.update_ops = []                # (moving_mean, moving_var, mean, var, rows, decay) pending updates


_S = _State()


def reset_variables(seed=0):
    """Drop every variable (a fresh tf.Graph)."""
    _S.variables.clear()
    _S.scope.clear()
    _S.update_ops.clear()
    _S.seed = seed


def global_variables(prefix=""):
    return OrderedDict((k, v) for k, v in _S.variables.items() if k.startswith(prefix))


@contextlib.contextmanager
def variable_scope(name, reuse=False):
    """tf.variable_scope(name, reuse=reuse): default layer names restart on every entry."""
    _S.scope.append((name, reuse, {}))
    try:
        yield
    finally:
        _S.scope.pop()


def _layer_scope(default, name):
    """full scope name of a layer: explicit ``name`` or TF's uniquified default (Conv, Conv_1, ...)."""
    prefix = "/".join(s[0] for s in _S.scope)
    if name is None:
        counters = _S.scope[-1][2] if _S.scope else _S.__dict__.setdefault("_root_counters", {})
        i = counters.get(default, 0)
        counters[default] = i + 1
        name = default if i == 0 else "%s_%d" % (default, i)
    return (prefix + "/" if prefix else "") + name


def _reuse():
    return any(s[1] for s in _S.scope)


def _he_trunc_normal(shape, device):
    """variance_scaling_initializer(factor=2.0, mode='FAN_IN', uniform=False) (utils/ops.py:60)."""
    fan_in = shape[-2] * int(math.prod(shape[:-2])) if len(shape) > 1 else shape[0]
    std = math.sqrt(1.3 * 2.0 / fan_in)
    gen = torch.Generator().manual_seed(_S.seed + len(_S.variables))
    t = torch.empty(*shape)
    torch.nn.init.trunc_normal_(t, 0.0, 1.0, -2.0, 2.0, generator=gen)
    return (t * std).to(device)


def _get_variable(full, shape, init, device):
    if full in _S.variables:
        v = _S.variables[full]
        if tuple(v.shape) != tuple(shape):
            raise ValueError("variable %s exists with shape %s, requested %s" % (full, tuple(v.shape), tuple(shape)))
        return v
    if _reuse():
        raise ValueError("Variable %s does not exist, or was not created with reuse disabled" % full)
    if callable(init):
        v = init(shape).to(device=device, dtype=torch.float32)
    elif init == "zeros":
        v = torch.zeros(*shape, device=device)
    elif init == "ones":
        v = torch.ones(*shape, device=device)
    else:
        v = _he_trunc_normal(shape, device)
    _S.variables[full] = v.contiguous()
    return _S.variables[full]


# ---- activations: tagged callables so that the kernels' fused epilogue can be used ------------
class _Act:
    def __init__(self, kind, alpha=0.0):
        self.kind, self.alpha = kind, alpha

    def __call__(self, x):
        if self.kind == "relu":
            return torch.relu(x)
        if self.kind == "lrelu":
            return torch.maximum(x, self.alpha * x)
        return torch.tanh(x)


relu = _Act("relu")
tanh = _Act("tanh")


def lrelu_act(alpha=0.2):
    """utils/ops.py:90-91"""
    return _Act("lrelu", alpha)


def _fused(act):
    if act is None:
        return K.ACT_NONE, None
    if isinstance(act, _Act) and act.kind == "relu":
        return K.ACT_RELU, None
    if isinstance(act, _Act) and act.kind == "lrelu" and abs(act.alpha - 0.2) < 1e-12:
        return K.ACT_LRELU, None
    return K.ACT_NONE, act          # applied on the fp32 result


# ---- layout / precision plumbing ---------------------------------------------------------------
def to_nchw(x):
    """utils/ops.py:129-130"""
    return x.permute(0, 3, 1, 2).contiguous()


def to_nhwc(x):
    """utils/ops.py:133-134"""
    return x.permute(0, 2, 3, 1).contiguous()


def _pad8(c):
    return (c + 7) // 8 * 8


def _planes_from(x_nhwc):
    """fp32 NHWC -> bf16 planes with the channel count padded to a multiple of 8"""
    n, h, w, c = x_nhwc.shape
    cp = _pad8(c)
    if cp != c:
        x_nhwc = torch.nn.functional.pad(x_nhwc, (0, cp - c))
    x_nhwc = x_nhwc.contiguous().float()
    t = torch.empty(PRECISION_PLANES, n, h, w, cp, device=x_nhwc.device, dtype=torch.bfloat16)
    K.to_planes(x_nhwc.view(n * h * w, cp), t)
    return t


def _planes_to(t, c):
    out = torch.empty(t.shape[1:], device=t.device, dtype=torch.float32)
    K.from_planes(t, out)
    return out[..., :c]


def _pack(w_kernel_layout):
    """fp32 [taps, cout, cin] -> planes, both channel counts padded to multiples of 8"""
    taps, co, ci = w_kernel_layout.shape
    cop, cip = _pad8(co), _pad8(ci)
    w = torch.zeros(taps, cop, cip, device=w_kernel_layout.device)
    w[:, :co, :ci] = w_kernel_layout
    t = torch.empty(PRECISION_PLANES, taps, cop, cip, device=w.device, dtype=torch.bfloat16)
    K.to_planes(w.view(taps * cop, cip), t)
    return t


def _bias8(b, n):
    out = torch.zeros(_pad8(n), device=b.device)
    out[:n] = b
    return out


def _in_nhwc(x, df):
    if df not in (NHWC, NCHW):
        raise ValueError('Invalid data format %s' % df)
    return x if df == NHWC else x.permute(0, 2, 3, 1)


def _out_df(y_nhwc, df):
    return y_nhwc.contiguous() if df == NHWC else y_nhwc.permute(0, 3, 1, 2).contiguous()


# ---- the wrappers ----------------------------------------------------------------------------------
def conv2d(x, f, ks=(4, 4), s=(2, 2), padding='SAME', act=None, init=None, name=None, df=NHWC):
    """tf.contrib.layers.conv2d with bias (utils/ops.py:58-63); weights HWIO [kh, kw, Cin, f]."""
    xn = _in_nhwc(x, df)
    n, h, w, cin = xn.shape
    scope = _layer_scope("Conv", name)
    wv = _get_variable(scope + "/weights", (ks[0], ks[1], cin, f), init, x.device)
    bv = _get_variable(scope + "/biases", (f,), "zeros", x.device)
    pad = padding.upper()
    fused, post = _fused(act)
    if ks[0] != ks[1] or s[0] != s[1]:
        raise ValueError("only square kernels and strides are supported")
    k, st = ks[0], s[0]
    if k == 4 and st in (1, 4) and pad == 'VALID' and h == 4 and w == 4:
        # model.py:160: one dot product per sample and output channel (GEMM with H = W = 1 over the 16*Cin patch)
        xp = _planes_from(xn.reshape(n, 1, 1, 16 * cin))
        wk = _pack(wv.reshape(16 * cin, f).t().reshape(1, f, 16 * cin))
        y = torch.empty(PRECISION_PLANES, n, 1, 1, _pad8(f), device=x.device, dtype=torch.bfloat16)
        K.conv_gemm(K.CONV_S1, 1, 0, K.View(xp), wk, K.View(y), bias=_bias8(bv, f), act=fused)
    else:
        if k in (1, 2, 3, 4) and st == 1 and (pad == 'SAME' or k == 1):
            mode, oh, ow = K.CONV_S1, h, w
        elif (k, st) == (4, 2) and pad == 'SAME':
            mode, oh, ow = K.CONV_K4S2, h // 2, w // 2
        else:
            raise ValueError("conv2d: unsupported geometry ks=%s s=%s padding=%s on %dx%d" % (ks, s, padding, h, w))
        xp = _planes_from(xn)
        wk = _pack(wv.permute(0, 1, 3, 2).reshape(k * k, f, cin))
        y = torch.empty(PRECISION_PLANES, n, oh, ow, _pad8(f), device=x.device, dtype=torch.bfloat16)
        K.conv_gemm(mode, k, 0, K.View(xp), wk, K.View(y), bias=_bias8(bv, f), act=fused)
    out = _planes_to(y, f)
    if post is not None:
        out = post(out)
    return _out_df(out, df)


def conv2d_transpose(x, f, ks=(4, 4), s=(2, 2), padding='SAME', act=None, init=None, name=None, df=NHWC):
    """tf.contrib.layers.conv2d_transpose with bias (utils/ops.py:66-71); weights [kh, kw, f, Cin]."""
    xn = _in_nhwc(x, df)
    n, h, w, cin = xn.shape
    if tuple(ks) != (4, 4) or tuple(s) != (2, 2) or padding.upper() != 'SAME':
        raise ValueError("conv2d_transpose: only ks=(4,4), s=(2,2), SAME is supported")
    scope = _layer_scope("Conv2d_transpose", name)
    wv = _get_variable(scope + "/weights", (4, 4, f, cin), init, x.device)
    bv = _get_variable(scope + "/biases", (f,), "zeros", x.device)
    fused, post = _fused(act)
    xp = _planes_from(xn)
    wk = _pack(wv.reshape(16, f, cin))
    y = torch.empty(PRECISION_PLANES, n, 2 * h, 2 * w, _pad8(f), device=x.device, dtype=torch.bfloat16)
    K.conv_gemm(K.DECONV_K4S2, 4, 0, K.View(xp), wk, K.View(y), bias=_bias8(bv, f), act=fused)
    out = _planes_to(y, f)
    if post is not None:
        out = post(out)
    return _out_df(out, df)


def fc(x, units, act=None, init=None, bias=True, name=None):
    """tf.layers.dense (utils/ops.py:84-87); kernel [in, units]."""
    n, cin = x.shape
    scope = _layer_scope("dense", name)
    wv = _get_variable(scope + "/kernel", (cin, units), init, x.device)
    bv = _get_variable(scope + "/bias", (units,), "zeros", x.device) if bias else None
    fused, post = _fused(act)
    xp = _planes_from(x.reshape(n, 1, 1, cin))
    wk = _pack(wv.t().reshape(1, units, cin))
    y = torch.empty(PRECISION_PLANES, n, 1, 1, _pad8(units), device=x.device, dtype=torch.bfloat16)
    K.conv_gemm(K.CONV_S1, 1, 0, K.View(xp), wk, K.View(y), bias=None if bv is None else _bias8(bv, units), act=fused)
    out = _planes_to(y, units).reshape(n, units)
    return post(out) if post is not None else out


def batch_norm(x, train, init=None, act=None, name=None, eps=1e-5, decay=0.9, df=NHWC):
    """tf.contrib.layers.batch_norm(scale=True, fused=True) (utils/ops.py:7-29).  Rank-4 inputs are
    normalised per channel over (N, H, W), rank-2 inputs per feature over N.  In training mode the
    moving-average updates are queued like TF's UPDATE_OPS and applied by ``run_update_ops()``."""
    rank2 = x.dim() == 2
    xn = x.reshape(x.shape[0], 1, 1, x.shape[1]) if rank2 else _in_nhwc(x, df)
    c = xn.shape[-1]
    scope = _layer_scope("BatchNorm", name)
    init = init or {}
    beta = _get_variable(scope + "/beta", (c,), init.get("beta", "zeros"), x.device)
    gamma = _get_variable(scope + "/gamma", (c,), init.get("gamma", "ones"), x.device)
    mm = _get_variable(scope + "/moving_mean", (c,), "zeros", x.device)
    mv = _get_variable(scope + "/moving_variance", (c,), "ones", x.device)
    cp = _pad8(c)
    xp = _planes_from(xn)
    rows = xp[0].numel() // cp
    pad = lambda v, fill: torch.cat([v, torch.full((cp - c,), fill, device=v.device)]) if cp != c else v
    if train:
        mean, rstd, var = (torch.empty(cp, device=x.device) for _ in range(3))
        K.bn_stats(xp, torch.zeros(2 * cp, device=x.device), mean, rstd, var, eps)
        _S.update_ops.append((mm, mv, mean[:c].clone(), var[:c].clone(), rows, decay))
    else:
        mean, rstd = pad(mm, 0.0), torch.rsqrt(pad(mv, 1.0) + eps)
    fused, post = _fused(act)
    y = torch.empty_like(xp)
    K.bn_apply(xp, mean, rstd, pad(gamma, 1.0), pad(beta, 0.0), y, None, fused == K.ACT_RELU)
    out = _planes_to(y, c)
    if fused == K.ACT_LRELU:
        out = torch.maximum(out, 0.2 * out)
    if post is not None:
        out = post(out)
    return out.reshape(x.shape[0], c) if rank2 else _out_df(out, df)


def run_update_ops():
    """tf.get_collection(UPDATE_OPS): apply the queued moving-statistics updates (model.py:98,102)."""
    for mm, mv, mean, var, rows, decay in _S.update_ops:
        K.bn_update_moving(mm, mv, mean.contiguous(), var.contiguous(), rows, decay)
    _S.update_ops.clear()


def layer_norm(x, act=None, scope=None, df=NHWC):
    """tf.contrib.layers.layer_norm(begin_norm_axis=1, begin_params_axis=-1 / 1) (utils/ops.py:74-81): statistics per
    sample over all non-batch axes (epsilon 1e-12), gamma / beta per channel (rank 2: per feature)."""
    if df not in (NHWC, NCHW):
        raise ValueError('Invalid data format %s' % df)
    rank2 = x.dim() == 2
    xn = x.reshape(x.shape[0], 1, 1, x.shape[1]) if rank2 else _in_nhwc(x, df)
    n, c = xn.shape[0], xn.shape[-1]
    if c % 8 != 0:
        raise ValueError("layer_norm: the channel count must be a multiple of 8 (zero padding would change the statistics)")
    full = _layer_scope("LayerNorm", scope)
    beta = _get_variable(full + "/beta", (c,), "zeros", x.device)
    gamma = _get_variable(full + "/gamma", (c,), "ones", x.device)
    xp = _planes_from(xn)
    sums = torch.zeros(n, 2, device=x.device)
    K.ln_stats(xp, sums)
    fused, post = _fused(act)
    y = torch.empty_like(xp)
    K.ln_apply(xp, sums, 1e-12, gamma, beta, y, fused == K.ACT_RELU)
    out = _planes_to(y, c)
    if fused == K.ACT_LRELU:
        out = torch.maximum(out, 0.2 * out)
    if post is not None:
        out = post(out)
    return out.reshape(x.shape[0], c) if rank2 else _out_df(out, df)


def pool(x, s=2, p_type='AVG', df=NHWC):
    """tf.nn.pool(window [s, s], strides [s, s], SAME) (utils/ops.py:100-101); AVG with s = 2 on even extents."""
    xn = _in_nhwc(x, df)
    n, h, w, c = xn.shape
    if s != 2 or p_type != 'AVG' or (h & 1) or (w & 1):
        raise ValueError("pool: only the 2x2 average pool on even extents is supported")
    xp = _planes_from(xn)
    y = torch.empty(PRECISION_PLANES, n, h // 2, w // 2, xp.shape[-1], device=x.device, dtype=torch.bfloat16)
    K.pool2x(xp, y, 0.25)
    return _out_df(_planes_to(y, c), df)


def resize_nearest_neighbor(x, new_size):
    """tf.image.resize_nearest_neighbor (utils/ops.py:104-106), NHWC; the 2x case."""
    n, h, w, c = x.shape
    if tuple(new_size) != (2 * h, 2 * w):
        raise ValueError("resize_nearest_neighbor: only the 2x upscale is supported")
    xp = _planes_from(x)
    y = torch.empty(PRECISION_PLANES, n, 2 * h, 2 * w, xp.shape[-1], device=x.device, dtype=torch.bfloat16)
    K.upscale2x(xp, y, 1.0)
    return _planes_to(y, c).contiguous()


def upscale(x, s=2):
    """utils/ops.py:109-111"""
    _, h, w, _ = get_conv_shape(x)
    return resize_nearest_neighbor(x, (h * s, w * s))


def get_conv_shape(tensor):
    """utils/ops.py:119-126"""
    return [int(d) for d in tensor.shape]
