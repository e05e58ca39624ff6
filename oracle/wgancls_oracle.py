"""CPU oracle for the wgancls hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A PyTorch-CPU restatement (fp32 or fp64) of the reference's TensorFlow-1.4 graph for one
wgancls iteration.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may import this module; the product path (``text-to-image_b200``) never does.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and TensorFlow 1.4
cannot be installed here, so nothing of the reference itself pins these numbers.  The pins are
(a) an independent naive NumPy conv/deconv (``oracle/naive.py``), (b) fp64 finite differences and
closed forms in ``tests/test_oracle.py`` and (c) the committed goldens under ``tests/golden``.

Every function cites the reference lines it restates (paths relative to /root/reference).
TF-1.4 semantics encoded here (recalled):
  * SAME padding for k4/s2 and k3/s1 on even extents is symmetric 1/1.
  * contrib conv2d weights are HWIO, conv2d_transpose weights are [kh, kw, Cout, Cin]; both
    layers always carry a zero-initialised bias.  tf.layers.dense kernel is [in, out].
  * variance_scaling_initializer(factor=2, FAN_IN, uniform=False): truncated normal (+-2 sigma,
    resampled) with sigma = sqrt(1.3 * 2 / fan_in), fan_in = prod(shape[:-2]) * shape[-2].
  * fused batch norm: batch mean, biased variance, eps inside rsqrt; moving stats EMA with
    decay 0.9, the moving variance receives the unbiased (Bessel) estimate.
  * tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); eps outside the sqrt.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

GP_WEIGHT = 150.0  # models/wgancls/model.py:91 (cfg LAMBDA is read at :75 and ignored)
LRELU_ALPHA = 0.2  # models/wgancls/model.py:110,131
BN_EPS = 1e-5      # utils/ops.py:7
BN_DECAY = 0.9     # utils/ops.py:7
ADAM_EPS = 1e-8    # tf.train.AdamOptimizer default
KT_LR = 0.001      # models/wgancls/model.py:100
KT_INIT = 0.7      # models/wgancls/model.py:77


@dataclass
class OracleCfg:
    """The constants of models/wgancls/cfg/flowers.yml:9-37 (overridable for tiny test nets)."""
    batch_size: int = 16
    z_dim: int = 128
    embed_dim: int = 1024
    compressed_embed_dim: int = 128
    gf_dim: int = 128
    df_dim: int = 128
    output_size: int = 64
    image_c: int = 3
    beta1: float = 0.0
    beta2: float = 0.9
    kl_coeff: float = 1.0
    d_lr: float = 1e-4
    g_lr: float = 1e-4
    n_critic: int = 1


# ----------------------------------------------------------------------------- init

def truncated_normal_(t: torch.Tensor, std: float, gen: torch.Generator) -> torch.Tensor:
    """tf.truncated_normal: values beyond 2 sigma are re-drawn (models/wgancls/model.py:119)."""
    t.normal_(0.0, 1.0, generator=gen)
    bad = t.abs() > 2.0
    while bool(bad.any()):
        t[bad] = torch.empty(int(bad.sum()), dtype=t.dtype).normal_(0.0, 1.0, generator=gen)
        bad = t.abs() > 2.0
    return t.mul_(std)


def he_trunc_normal(shape, gen, dtype):
    """utils/ops.py:60,68,86 -- variance_scaling_initializer(2.0, 'FAN_IN', uniform=False)."""
    fan_in = shape[-2]
    for d in shape[:-2]:
        fan_in *= d
    std = math.sqrt(1.3 * 2.0 / fan_in)
    return truncated_normal_(torch.empty(*shape, dtype=torch.float32), std, gen).to(dtype)


def param_shapes(cfg: OracleCfg) -> "OrderedDict[str, tuple]":
    """Variable names and TF-layout shapes in creation order.

    g_net: models/wgancls/model.py:163-225 ; d_net: models/wgancls/model.py:129-161.
    Names are TF's default scopes.
    """
    gf, df, ce = cfg.gf_dim, cfg.df_dim, cfg.compressed_embed_dim
    s16 = cfg.output_size // 16
    c = cfg.image_c
    sh: "OrderedDict[str, tuple]" = OrderedDict()

    def dense(scope, i, o):
        sh[scope + "/kernel"] = (i, o)
        sh[scope + "/bias"] = (o,)

    def conv(scope, k, i, o):
        sh[scope + "/weights"] = (k, k, i, o)
        sh[scope + "/biases"] = (o,)

    def deconv(scope, k, i, o):
        sh[scope + "/weights"] = (k, k, o, i)
        sh[scope + "/biases"] = (o,)

    def bn(scope, ch):
        sh[scope + "/beta"] = (ch,)
        sh[scope + "/gamma"] = (ch,)
        sh[scope + "/moving_mean"] = (ch,)
        sh[scope + "/moving_variance"] = (ch,)

    g = "g_net/"
    dense(g + "dense", cfg.embed_dim, ce)            # mean          model.py:113
    dense(g + "dense_1", cfg.embed_dim, ce)          # log_sigma     model.py:114
    dense(g + "dense_2", cfg.z_dim + ce, gf * 8 * s16 * s16)  # model.py:175
    bn(g + "BatchNorm", gf * 8 * s16 * s16)          # model.py:176
    conv(g + "Conv", 1, gf * 8, gf * 2); bn(g + "BatchNorm_1", gf * 2)      # :184-185
    conv(g + "Conv_1", 3, gf * 2, gf * 2); bn(g + "BatchNorm_2", gf * 2)    # :186-187
    conv(g + "Conv_2", 3, gf * 2, gf * 8); bn(g + "BatchNorm_3", gf * 8)    # :188-189
    deconv(g + "Conv2d_transpose", 4, gf * 8, gf * 4)                       # :194
    conv(g + "Conv_3", 3, gf * 4, gf * 4); bn(g + "BatchNorm_4", gf * 4)    # :195-196
    conv(g + "Conv_4", 1, gf * 4, gf); bn(g + "BatchNorm_5", gf)            # :200-201
    conv(g + "Conv_5", 3, gf, gf); bn(g + "BatchNorm_6", gf)                # :202-203
    conv(g + "Conv_6", 3, gf, gf * 4); bn(g + "BatchNorm_7", gf * 4)        # :204-205
    deconv(g + "Conv2d_transpose_1", 4, gf * 4, gf * 2)                     # :210
    conv(g + "Conv_7", 3, gf * 2, gf * 2); bn(g + "BatchNorm_8", gf * 2)    # :211-212
    deconv(g + "Conv2d_transpose_2", 4, gf * 2, gf)                         # :214
    conv(g + "Conv_8", 3, gf, gf); bn(g + "BatchNorm_9", gf)                # :215-216
    deconv(g + "Conv2d_transpose_3", 4, gf, c)                              # :218
    conv(g + "Conv_9", 3, c, c)                                             # :219

    d = "d_net/"
    conv(d + "Conv", 4, c, df)               # :135
    conv(d + "Conv_1", 4, df, df * 2)        # :136
    conv(d + "Conv_2", 4, df * 2, df * 4)    # :137
    conv(d + "Conv_3", 4, df * 4, df * 8)    # :138
    conv(d + "Conv_4", 1, df * 8, df * 2)    # :142
    conv(d + "Conv_5", 3, df * 2, df * 4)    # :143
    conv(d + "Conv_6", 3, df * 4, df * 8)    # :144
    dense(d + "dense", cfg.embed_dim, ce)    # :150
    conv(d + "Conv_7", 3, df * 8 + ce, df * 8)  # :157
    conv(d + "Conv_8", 1, df * 8, df * 8)    # :158
    conv(d + "Conv_9", 4, df * 8, 1)         # :160
    return sh


def is_trainable(name: str) -> bool:
    return not (name.endswith("moving_mean") or name.endswith("moving_variance"))


def init_params(cfg: OracleCfg, seed: int = 0, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Reference initialisation: He-TN weights, zero biases, gamma=1, beta=0, moving 0/1."""
    gen = torch.Generator().manual_seed(seed)
    p: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("weights", "kernel"):
            p[name] = he_trunc_normal(shape, gen, dtype)
        elif leaf in ("gamma", "moving_variance"):
            p[name] = torch.ones(shape, dtype=dtype)
        else:
            p[name] = torch.zeros(shape, dtype=dtype)
    return p


def d_var_names(p):  # models/wgancls/model.py:59  tf.trainable_variables('d_net')
    return [n for n in p if n.startswith("d_net/") and is_trainable(n)]


def g_var_names(p):  # models/wgancls/model.py:60
    return [n for n in p if n.startswith("g_net/") and is_trainable(n)]


# ----------------------------------------------------------------------------- ops (utils/ops.py)

def lrelu(x):
    return torch.maximum(x, LRELU_ALPHA * x)  # tf.nn.leaky_relu(x, .2) == max(.2x, x)


def conv2d(p, scope, x, k, s, padding="SAME", act=None):
    """utils/ops.py:58-63 on NCHW input; weight HWIO -> OIHW."""
    w = p[scope + "/weights"].permute(3, 2, 0, 1)
    pad = 0
    if padding.upper() == "SAME":
        assert (k, s) in ((4, 2), (3, 1), (1, 1), (4, 1)), "only the SAME cases of the implemented paths"
        if (k, s) == (4, 1):     # pad_total = 3: TF puts 1 before and 2 after (StackGAN stage-II)
            x = F.pad(x, (1, 2, 1, 2))
        else:
            pad = {4: 1, 3: 1, 1: 0}[k]
    y = F.conv2d(x, w, p[scope + "/biases"], stride=s, padding=pad)
    return act(y) if act is not None else y


def conv2d_transpose(p, scope, x):
    """utils/ops.py:66-71, k4 s2 SAME; TF weight [kh,kw,Cout,Cin] -> torch [Cin,Cout,kh,kw]."""
    w = p[scope + "/weights"].permute(3, 2, 0, 1)
    return F.conv_transpose2d(x, w, p[scope + "/biases"], stride=2, padding=1)


def fc(p, scope, x, act=None):
    """utils/ops.py:84-87."""
    y = x @ p[scope + "/kernel"] + p[scope + "/bias"]
    return act(y) if act is not None else y


def batch_norm(p, scope, x, train, act=None, new_moving=None):
    """utils/ops.py:7-29 (NCHW rank-4 per channel, rank-2 per feature).

    When ``train`` and ``new_moving`` is a dict, the EMA-updated moving statistics are stored in
    it (they are applied only by the G step: models/wgancls/model.py:98,102).
    """
    dims = (0, 2, 3) if x.dim() == 4 else (0,)
    shape = (1, -1, 1, 1) if x.dim() == 4 else (1, -1)
    gamma, beta = p[scope + "/gamma"], p[scope + "/beta"]
    if train:
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        if new_moving is not None:
            n = x.numel() // x.shape[1]
            unbiased = var * (n / max(n - 1, 1))
            new_moving[scope + "/moving_mean"] = (
                BN_DECAY * p[scope + "/moving_mean"] + (1 - BN_DECAY) * mean).detach()
            new_moving[scope + "/moving_variance"] = (
                BN_DECAY * p[scope + "/moving_variance"] + (1 - BN_DECAY) * unbiased).detach()
    else:
        mean, var = p[scope + "/moving_mean"], p[scope + "/moving_variance"]
    y = (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + BN_EPS) * gamma.view(shape) + beta.view(shape)
    return act(y) if act is not None else y


# ----------------------------------------------------------------------------- the networks

def generator(p, z, embed, tn_eps, cfg: OracleCfg, is_training=True, cond_noise=True, new_moving=None,
              fc_reshape="nchw"):
    """models/wgancls/model.py:163-225 (+ :108-122).  ``tn_eps`` is the truncated-normal draw of
    :119 made explicit.  Returns (image NHWC, mean, log_sigma).
    fc_reshape="nhwc": the dense output is reshaped [-1, s16, s16, C] as the (otherwise layer-for-layer
    identical) StackGAN stage-I generator does (models/stackgan/stageI/model.py:129)."""
    g = "g_net/"
    gf = cfg.gf_dim
    s16 = cfg.output_size // 16
    bnk = dict(train=is_training, new_moving=new_moving)
    mean = fc(p, g + "dense", embed, lrelu)                      # :113
    log_sigma = fc(p, g + "dense_1", embed, lrelu)               # :114
    c = mean + torch.exp(log_sigma) * tn_eps if cond_noise else mean   # :117-122
    h0 = fc(p, g + "dense_2", torch.cat([z, c], 1))              # :174-175
    h0 = batch_norm(p, g + "BatchNorm", h0, **bnk)               # :176
    if fc_reshape == "nchw":
        h0 = h0.reshape(-1, gf * 8, s16, s16)                    # :179 (NCHW)
    else:
        h0 = h0.reshape(-1, s16, s16, gf * 8).permute(0, 3, 1, 2)

    def res(x, c1, b1, c2, b2, c3, b3):
        n = conv2d(p, g + c1, x, 1, 1, "valid")
        n = batch_norm(p, g + b1, n, act=torch.relu, **bnk)
        n = conv2d(p, g + c2, n, 3, 1)
        n = batch_norm(p, g + b2, n, act=torch.relu, **bnk)
        n = conv2d(p, g + c3, n, 3, 1)
        n = batch_norm(p, g + b3, n, **bnk)
        return torch.relu(x + n)

    h1 = res(h0, "Conv", "BatchNorm_1", "Conv_1", "BatchNorm_2", "Conv_2", "BatchNorm_3")   # :184-191
    h2 = conv2d_transpose(p, g + "Conv2d_transpose", h1)         # :194
    h2 = conv2d(p, g + "Conv_3", h2, 3, 1)                       # :195
    h2 = batch_norm(p, g + "BatchNorm_4", h2, **bnk)             # :196
    h3 = res(h2, "Conv_4", "BatchNorm_5", "Conv_5", "BatchNorm_6", "Conv_6", "BatchNorm_7")  # :200-207
    h4 = conv2d_transpose(p, g + "Conv2d_transpose_1", h3)       # :210
    h4 = conv2d(p, g + "Conv_7", h4, 3, 1)
    h4 = batch_norm(p, g + "BatchNorm_8", h4, act=torch.relu, **bnk)   # :212
    h5 = conv2d_transpose(p, g + "Conv2d_transpose_2", h4)       # :214
    h5 = conv2d(p, g + "Conv_8", h5, 3, 1)
    h5 = batch_norm(p, g + "BatchNorm_9", h5, act=torch.relu, **bnk)   # :216
    lg = conv2d_transpose(p, g + "Conv2d_transpose_3", h5)       # :218
    lg = conv2d(p, g + "Conv_9", lg, 3, 1)                       # :219
    out = torch.tanh(lg).permute(0, 2, 3, 1)                     # :221-224 (to_nhwc)
    return out, mean, log_sigma


def discriminator(p, x_nhwc, embed, cfg: OracleCfg):
    """models/wgancls/model.py:129-161.  Returns logits [B,1,1,1]."""
    d = "d_net/"
    x = x_nhwc.permute(0, 3, 1, 2)                               # :132 to_nchw
    h0 = conv2d(p, d + "Conv", x, 4, 2, act=lrelu)               # :135
    h1 = conv2d(p, d + "Conv_1", h0, 4, 2, act=lrelu)
    h2 = conv2d(p, d + "Conv_2", h1, 4, 2, act=lrelu)
    h3 = conv2d(p, d + "Conv_3", h2, 4, 2)                       # :138 (no activation)
    n = conv2d(p, d + "Conv_4", h3, 1, 1, "valid", act=lrelu)    # :142
    n = conv2d(p, d + "Conv_5", n, 3, 1, act=lrelu)
    n = conv2d(p, d + "Conv_6", n, 3, 1)
    h4 = lrelu(h3 + n)                                           # :145-146
    e = fc(p, d + "dense", embed, lrelu)                         # :150
    e = e[:, :, None, None].expand(-1, -1, 4, 4)                 # :153-154 (tile is hard-coded 4x4)
    h4c = torch.cat([h4, e], 1)                                  # :155
    h5 = conv2d(p, d + "Conv_7", h4c, 3, 1, "same", act=lrelu)   # :157
    h6 = conv2d(p, d + "Conv_8", h5, 1, 1, "valid", act=lrelu)   # :158
    return conv2d(p, d + "Conv_9", h6, 4, 4, "valid")            # :160


# ----------------------------------------------------------------------------- losses

def gradient_penalty(grad, dims):
    """models/wgancls/model.py:62-70: one-sided, no epsilon under the sqrt."""
    slopes = torch.sqrt(torch.sum(grad * grad, dim=dims))
    return torch.mean(torch.clamp(slopes - 1.0, min=0.0) ** 2), slopes


def kl_std_normal_loss(mean, log_sigma):
    """models/wgancls/model.py:124-127."""
    return torch.mean(-log_sigma + 0.5 * (-1 + torch.exp(2.0 * log_sigma) + mean * mean))


def d_forward_losses(p, kt, feed, cfg, create_graph):
    """models/wgancls/model.py:48-55,79-91: everything the D run evaluates.

    feed keys: x, x_mismatch [B,H,W,3]; cond [B,E]; z [B,Z]; epsilon [B,1,1,1]; tn_eps [B,ce].
    """
    G, mean, log_sigma = generator(p, feed["z"], feed["cond"], feed["tn_eps"], cfg)
    Dg = discriminator(p, G, feed["cond"], cfg)
    Dx = discriminator(p, feed["x"], feed["cond"], cfg)
    Dxmi = discriminator(p, feed["x_mismatch"], feed["cond"], cfg)
    x_hat = feed["epsilon"] * G + (1.0 - feed["epsilon"]) * feed["x"]      # :53
    cond_inp = feed["cond"] + 0.0                                          # :54
    if not x_hat.requires_grad:
        x_hat.requires_grad_(True)
    if not cond_inp.requires_grad:
        cond_inp.requires_grad_(True)
    Dxh = discriminator(p, x_hat, cond_inp, cfg)
    g_x, g_c = torch.autograd.grad(Dxh.sum(), [x_hat, cond_inp], create_graph=create_graph)
    gp, slopes = gradient_penalty(g_x, (1, 2, 3))
    gp2, slopes2 = gradient_penalty(g_c, (1,))
    out = OrderedDict()
    out["G"] = G
    out["embed_mean"], out["embed_log_sigma"] = mean, log_sigma
    out["Dg_logit"], out["Dx_logit"], out["Dxmi_logit"], out["Dx_hat_logit"] = Dg, Dx, Dxmi, Dxh
    out["x_hat"] = x_hat
    out["grad_x_hat"], out["grad_cond"] = g_x, g_c
    out["slopes"], out["slopes2"] = slopes, slopes2
    out["D_loss_real"] = Dx.mean()
    out["D_loss_fake"] = Dg.mean()
    out["D_loss_mismatch"] = Dxmi.mean()
    out["wdist"] = out["D_loss_real"] - out["D_loss_fake"]
    out["wdist2"] = out["D_loss_real"] - out["D_loss_mismatch"]
    out["reg_loss"] = (Dxmi * Dxmi).mean()                                  # :84 (unused by any update)
    out["balance_loss"] = (kt * out["wdist2"] - out["wdist"]) ** 2          # :85
    out["G_kl_loss"] = kl_std_normal_loss(mean, log_sigma)
    out["real_gp"], out["real_gp2"] = gp, gp2
    out["D_loss"] = -out["wdist"] - kt * out["wdist2"] + GP_WEIGHT * (gp + gp2)   # :91
    out["G_loss"] = -out["D_loss_fake"] + cfg.kl_coeff * out["G_kl_loss"]         # :92
    return out


def adam_tf(theta, grad, m, v, lr, beta1, beta2, t):
    """tf.train.AdamOptimizer (models/wgancls/model.py:94-96,103-105); t is the 1-based step."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m = beta1 * m + (1.0 - beta1) * grad
    v = beta2 * v + (1.0 - beta2) * grad * grad
    theta = theta - lr_t * m / (torch.sqrt(v) + ADAM_EPS)
    return theta, m, v


def new_state(p, dtype=None):
    """Optimizer state: Adam slots for every trainable variable, kt, step counters."""
    st = {"kt": torch.tensor(KT_INIT, dtype=dtype or next(iter(p.values())).dtype),
          "d_t": 0, "g_t": 0, "m": {}, "v": {}}
    for n, w in p.items():
        if is_trainable(n):
            st["m"][n] = torch.zeros_like(w)
            st["v"][n] = torch.zeros_like(w)
    return st


def d_step(p, st, feed, cfg: OracleCfg, lr_d=None):
    """sess.run([D_optim, kt_optim, D_loss]) -- models/wgancls/trainer.py:97.

    D gradients and the kt gradient are both taken at the OLD kt (the reference leaves the order
    unconstrained; this is the deterministic choice the product makes too).
    Returns (fetch dict incl. gradients, in-place updated p/st).
    """
    lr_d = cfg.d_lr if lr_d is None else lr_d
    names = d_var_names(p)
    kt = st["kt"].clone().requires_grad_(True)
    pl = {n: (w.detach().clone().requires_grad_(True) if n in names else w.detach()) for n, w in p.items()}
    out = d_forward_losses(pl, kt, feed, cfg, create_graph=True)
    grads = torch.autograd.grad(out["D_loss"], [pl[n] for n in names], retain_graph=True, allow_unused=True)
    (kt_grad,) = torch.autograd.grad(out["balance_loss"], [kt])
    st["d_t"] += 1
    res = {k: v.detach() for k, v in out.items()}
    res["grads"] = {}
    for n, g in zip(names, grads):
        g = torch.zeros_like(p[n]) if g is None else g.detach()
        res["grads"][n] = g
        p[n], st["m"][n], st["v"][n] = adam_tf(p[n], g, st["m"][n], st["v"][n], lr_d,
                                                cfg.beta1, cfg.beta2, st["d_t"])
    res["kt_grad"] = kt_grad.detach()
    st["kt"] = (st["kt"] - KT_LR * kt_grad).detach()              # model.py:100
    return res


def g_step(p, st, feed, cfg: OracleCfg, lr_g=None):
    """sess.run([G_optim, G_loss]) -- models/wgancls/trainer.py:101 (fresh tn_eps: feed['tn_eps_g'])."""
    lr_g = cfg.g_lr if lr_g is None else lr_g
    names = g_var_names(p)
    pl = {n: (w.detach().clone().requires_grad_(True) if n in names else w.detach()) for n, w in p.items()}
    new_moving = {}
    G, mean, log_sigma = generator(pl, feed["z"], feed["cond"], feed["tn_eps_g"], cfg, new_moving=new_moving)
    Dg = discriminator(pl, G, feed["cond"], cfg)
    kl = kl_std_normal_loss(mean, log_sigma)
    g_loss = -Dg.mean() + cfg.kl_coeff * kl                        # model.py:92
    grads = torch.autograd.grad(g_loss, [pl[n] for n in names], allow_unused=True)
    st["g_t"] += 1
    res = {"G": G.detach(), "G_loss": g_loss.detach(), "G_kl_loss": kl.detach(),
           "Dg_logit": Dg.detach(), "grads": {}}
    for n, g in zip(names, grads):
        g = torch.zeros_like(p[n]) if g is None else g.detach()
        res["grads"][n] = g
        p[n], st["m"][n], st["v"][n] = adam_tf(p[n], g, st["m"][n], st["v"][n], lr_g,
                                                cfg.beta1, cfg.beta2, st["g_t"])
    for n, w in new_moving.items():                                # UPDATE_OPS, model.py:98,102
        p[n] = w
    return res


def make_feed(cfg: OracleCfg, seed: int = 1234, dtype=torch.float32, batch=None):
    """Synthetic batch of SURVEY.md 8(d): x, x_mismatch ~ U(-1,1); cond, z ~ N(0,1);
    epsilon ~ U(0,1) (models/wgancls/trainer.py:79-80); two truncated-normal draws."""
    B = cfg.batch_size if batch is None else batch
    gen = torch.Generator().manual_seed(seed)
    hw = cfg.output_size
    f = {}
    f["x"] = (torch.rand(B, hw, hw, cfg.image_c, generator=gen) * 2 - 1).to(dtype)
    f["x_mismatch"] = (torch.rand(B, hw, hw, cfg.image_c, generator=gen) * 2 - 1).to(dtype)
    f["cond"] = torch.randn(B, cfg.embed_dim, generator=gen).to(dtype)
    f["z"] = torch.randn(B, cfg.z_dim, generator=gen).to(dtype)
    f["epsilon"] = torch.rand(B, 1, 1, 1, generator=gen).to(dtype)
    f["tn_eps"] = truncated_normal_(torch.empty(B, cfg.compressed_embed_dim), 1.0, gen).to(dtype)
    f["tn_eps_g"] = truncated_normal_(torch.empty(B, cfg.compressed_embed_dim), 1.0, gen).to(dtype)
    return f


def iteration(p, st, feed, cfg: OracleCfg):
    """One trainer iteration (models/wgancls/trainer.py:97-102, N_CRITIC=1)."""
    rd = d_step(p, st, feed, cfg)
    rg = g_step(p, st, feed, cfg)
    return rd, rg
