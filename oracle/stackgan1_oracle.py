"""CPU oracle for StackGAN stage-I (models/stackgan/stageI) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A PyTorch-CPU restatement of the reference's TF-1.4 graph for one ConditionalGan (stage-I) iteration:
models/stackgan/stageI/model.py (generator :114-171, discriminator :76-112) and the losses / optimizers of
models/stackgan/stageI/trainer.py:19-52.  PARITY UNPINNED like oracle/wgancls_oracle.py: the reference has no
tests or fixtures and TensorFlow 1.4 is not installable; the pins are the shared op restatements (checked
against oracle/naive.py and finite differences in tests/test_oracle.py), closed forms in
tests/test_stackgan1_cpu.py and torch.autograd.  Only tests/ may import this module.

Differences from wgancls that this file encodes:
  * weights ~ N(0, 0.02) (model.py:29), BatchNorm gamma ~ N(1, 0.02) (:30-32); dense layers of
    generate_conditionals use w_init too, the discriminator's embedding dense keeps the glorot default --
    initial values are irrelevant to parity, every test loads explicit parameters;
  * NHWC throughout (utils/ops.py defaults): the generator's dense output is reshaped [-1, 4, 4, C] (:129);
  * the discriminator has BatchNorm after every conv but the first and the logits conv (:83-109), so each of its
    three calls (synthetic, real/match, real/mismatch: :45-49) normalises with ITS OWN batch statistics;
  * sigmoid cross-entropy losses with one-sided label smoothing 0.9 (trainer.py:21-37), alpha-weighted (:40-44);
  * Adam(beta1 = D/G_BETA_DECAY, beta2 = 0.999) on d_vars / g_vars, both under ALL UPDATE_OPS (:50-55): every
    BatchNorm call of the graph steps its moving statistics in both runs (g_net: used by the sampler; d_net:
    never read anywhere in the reference, the discriminator is always built with is_training=True).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import wgancls_oracle as W
from .wgancls_oracle import batch_norm, conv2d, fc, is_trainable, lrelu, kl_std_normal_loss, truncated_normal_  # noqa: F401

ADAM_BETA2 = 0.999    # tf.train.AdamOptimizer default (trainer.py:52,54 pass beta1 only)
REAL_LABEL = 0.9      # trainer.py:26


@dataclass
class Stage1Cfg:
    """models/stackgan/stageI/cfg/flowers.yml:10-37 (overridable for tiny test nets)."""
    batch_size: int = 8
    z_dim: int = 100
    embed_dim: int = 1024
    compressed_embed_dim: int = 128
    gf_dim: int = 128
    df_dim: int = 64
    output_size: int = 64
    image_c: int = 3
    d_beta1: float = 0.5
    g_beta1: float = 0.5
    alpha_mismatch: float = 0.5
    kl_coeff: float = 2.0
    lr: float = 2e-4


def param_shapes(cfg: Stage1Cfg) -> "OrderedDict[str, tuple]":
    """g_net: identical variable list to wgancls (same layers, same default scope names); d_net: :76-112."""
    wc = W.OracleCfg(z_dim=cfg.z_dim, embed_dim=cfg.embed_dim, compressed_embed_dim=cfg.compressed_embed_dim,
                     gf_dim=cfg.gf_dim, df_dim=cfg.df_dim, output_size=cfg.output_size, image_c=cfg.image_c)
    sh = OrderedDict((k, v) for k, v in W.param_shapes(wc).items() if k.startswith("g_net/"))
    df, ce, c = cfg.df_dim, cfg.compressed_embed_dim, cfg.image_c
    s16 = cfg.output_size // 16
    d = "d_net/"

    def conv(scope, k, i, o):
        sh[scope + "/weights"] = (k, k, i, o)
        sh[scope + "/biases"] = (o,)

    def bn(scope, ch):
        for leaf in ("beta", "gamma", "moving_mean", "moving_variance"):
            sh[scope + "/" + leaf] = (ch,)

    conv(d + "Conv", 4, c, df)                                              # :81
    conv(d + "Conv_1", 4, df, df * 2); bn(d + "BatchNorm", df * 2)          # :82-83
    conv(d + "Conv_2", 4, df * 2, df * 4); bn(d + "BatchNorm_1", df * 4)    # :84-85
    conv(d + "Conv_3", 4, df * 4, df * 8); bn(d + "BatchNorm_2", df * 8)    # :86-87
    conv(d + "Conv_4", 1, df * 8, df * 2); bn(d + "BatchNorm_3", df * 2)    # :91-92
    conv(d + "Conv_5", 3, df * 2, df * 2); bn(d + "BatchNorm_4", df * 2)    # :93-94
    conv(d + "Conv_6", 3, df * 2, df * 8); bn(d + "BatchNorm_5", df * 8)    # :95-96
    sh[d + "dense/kernel"] = (cfg.embed_dim, ce)                            # :102
    sh[d + "dense/bias"] = (ce,)
    conv(d + "Conv_7", 1, df * 8 + ce, df * 8); bn(d + "BatchNorm_6", df * 8)   # :109-110
    conv(d + "Conv_8", s16, df * 8, 1)                                      # :112
    return sh


def init_params(cfg: Stage1Cfg, seed: int = 0, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("weights", "kernel"):
            p[name] = (torch.randn(*shape, generator=gen) * 0.02).to(dtype)          # model.py:29
        elif leaf == "gamma":
            p[name] = (1.0 + 0.02 * torch.randn(*shape, generator=gen)).to(dtype)    # model.py:30-32
        elif leaf == "moving_variance":
            p[name] = torch.ones(shape, dtype=dtype)
        else:
            p[name] = torch.zeros(shape, dtype=dtype)
    return p


def d_var_names(p):
    return [n for n in p if n.startswith("d_net/") and is_trainable(n)]


def g_var_names(p):
    return [n for n in p if n.startswith("g_net/") and is_trainable(n)]


def generator(p, z, embed, tn_eps, cfg: Stage1Cfg, is_training=True, cond_noise=True, new_moving=None):
    """models/stackgan/stageI/model.py:114-171: layer for layer the wgancls generator, NHWC reshape at :129."""
    return W.generator(p, z, embed, tn_eps, cfg, is_training=is_training, cond_noise=cond_noise, new_moving=new_moving,
                       fc_reshape="nhwc")


def discriminator(p, x_nhwc, embed, cfg: Stage1Cfg, new_moving=None):
    """models/stackgan/stageI/model.py:76-112 (is_training=True always).  Returns logits [B,1,1,1]."""
    d = "d_net/"
    bnk = dict(train=True, new_moving=new_moving)
    x = x_nhwc.permute(0, 3, 1, 2)
    h0 = conv2d(p, d + "Conv", x, 4, 2, act=lrelu)                                   # :81
    h1 = batch_norm(p, d + "BatchNorm", conv2d(p, d + "Conv_1", h0, 4, 2), act=lrelu, **bnk)      # :82-83
    h2 = batch_norm(p, d + "BatchNorm_1", conv2d(p, d + "Conv_2", h1, 4, 2), act=lrelu, **bnk)    # :84-85
    h3 = batch_norm(p, d + "BatchNorm_2", conv2d(p, d + "Conv_3", h2, 4, 2), **bnk)                # :86-87
    n = batch_norm(p, d + "BatchNorm_3", conv2d(p, d + "Conv_4", h3, 1, 1, "valid"), act=lrelu, **bnk)   # :91-92
    n = batch_norm(p, d + "BatchNorm_4", conv2d(p, d + "Conv_5", n, 3, 1), act=lrelu, **bnk)             # :93-94
    n = batch_norm(p, d + "BatchNorm_5", conv2d(p, d + "Conv_6", n, 3, 1), **bnk)                        # :95-96
    h4 = lrelu(h3 + n)                                                               # :97-98
    e = fc(p, d + "dense", embed, lrelu)                                             # :102
    e = e[:, :, None, None].expand(-1, -1, 4, 4)                                     # :105-106
    h4 = batch_norm(p, d + "BatchNorm_6", conv2d(p, d + "Conv_7", torch.cat([h4, e], 1), 1, 1, "valid"), act=lrelu,
                    **bnk)                                                           # :107-110
    s16 = cfg.output_size // 16
    w = p[d + "Conv_8/weights"].permute(3, 2, 0, 1)
    return F.conv2d(h4, w, p[d + "Conv_8/biases"], stride=s16).permute(0, 2, 3, 1)  # :112


def sigmoid_ce(logits, label):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x z + log(1 + exp(-|x|)) = log(1 + e^x) - x z, mean over
    the batch (written with logaddexp so that autograd gives sigmoid(x) - z at x = 0 too)."""
    x = logits.reshape(-1)
    return (torch.logaddexp(x, torch.zeros_like(x)) - x * label).mean()


def adam_tf(theta, grad, m, v, lr, beta1, t):
    lr_t = lr * math.sqrt(1.0 - ADAM_BETA2 ** t) / (1.0 - beta1 ** t)
    m = beta1 * m + (1.0 - beta1) * grad
    v = ADAM_BETA2 * v + (1.0 - ADAM_BETA2) * grad * grad
    return theta - lr_t * m / (torch.sqrt(v) + W.ADAM_EPS), m, v


def new_state(p):
    st = {"d_t": 0, "g_t": 0, "m": {}, "v": {}}
    for n, w in p.items():
        if is_trainable(n):
            st["m"][n] = torch.zeros_like(w)
            st["v"][n] = torch.zeros_like(w)
    return st


def _apply_moving(p, chain):
    """UPDATE_OPS: every BatchNorm call steps the moving statistics; calls on the same variables (the three
    discriminator calls) are applied one after the other, in graph-construction order."""
    for upd in chain:
        for n, w in upd.items():
            p[n] = w


def _forward_all(pl, p, feed, tn_eps, cfg, moving_chain):
    """model.py:44-49 with the moving-statistics updates of each BatchNorm call collected in call order."""
    mg = {}
    G, mean, log_sigma = generator(pl, feed["z"], feed["cond"], tn_eps, cfg, new_moving=mg)
    moving_chain.append(mg)
    logits = []
    for img in (G, feed["x"], feed["x_mismatch"]):
        # each later call sees the moving statistics left by the previous one
        cur = dict(pl)
        for upd in moving_chain:
            cur.update({k: v for k, v in upd.items() if k.startswith("d_net/")})
        md = {}
        logits.append(discriminator(cur, img, feed["cond"], cfg, new_moving=md))
        moving_chain.append(md)
    return G, mean, log_sigma, logits


def d_step(p, st, feed, cfg: Stage1Cfg, lr=None):
    """sess.run([D_optim, D_loss, ...]) -- trainer.py:139-140."""
    lr = cfg.lr if lr is None else lr
    names = d_var_names(p)
    pl = {n: (w.detach().clone().requires_grad_(True) if n in names else w.detach()) for n, w in p.items()}
    chain = []
    G, mean, log_sigma, (l_syn, l_real, l_mis) = _forward_all(pl, p, feed, feed["tn_eps"], cfg, chain)
    res = OrderedDict()
    res["G"] = G.detach()
    res["D_synthetic_logits"], res["D_real_match_logits"], res["D_real_mismatch_logits"] = (
        l_syn.detach(), l_real.detach(), l_mis.detach())
    res["D_synthetic_loss"] = sigmoid_ce(l_syn, 0.0)                       # trainer.py:21-23
    res["D_real_match_loss"] = sigmoid_ce(l_real, REAL_LABEL)              # :24-26
    res["D_real_mismatch_loss"] = sigmoid_ce(l_mis, 0.0)                   # :27-29
    a = cfg.alpha_mismatch
    d_loss = res["D_real_match_loss"] + a * res["D_real_mismatch_loss"] + (1.0 - a) * res["D_synthetic_loss"]   # :40-43
    grads = torch.autograd.grad(d_loss, [pl[n] for n in names], allow_unused=True)
    st["d_t"] += 1
    res = OrderedDict((k, v.detach()) for k, v in res.items())
    res["D_loss"] = d_loss.detach()
    res["grads"] = {}
    for n, g in zip(names, grads):
        g = torch.zeros_like(p[n]) if g is None else g.detach()
        res["grads"][n] = g
        p[n], st["m"][n], st["v"][n] = adam_tf(p[n], g, st["m"][n], st["v"][n], lr, cfg.d_beta1, st["d_t"])
    _apply_moving(p, chain)
    return res


def g_step(p, st, feed, cfg: Stage1Cfg, lr=None):
    """sess.run([G_optim, G_loss, ...]) -- trainer.py:144-145 (a fresh truncated-normal draw: feed['tn_eps_g'])."""
    lr = cfg.lr if lr is None else lr
    names = g_var_names(p)
    pl = {n: (w.detach().clone().requires_grad_(True) if n in names else w.detach()) for n, w in p.items()}
    chain = []
    G, mean, log_sigma, (l_syn, _, _) = _forward_all(pl, p, feed, feed["tn_eps_g"], cfg, chain)
    kl = kl_std_normal_loss(mean, log_sigma)                               # trainer.py:31,57-60
    gan = sigmoid_ce(l_syn, 1.0)                                           # :32-34
    g_loss = gan + cfg.kl_coeff * kl                                       # :44
    grads = torch.autograd.grad(g_loss, [pl[n] for n in names], allow_unused=True)
    st["g_t"] += 1
    res = {"G": G.detach(), "G_loss": g_loss.detach(), "G_gan_loss": gan.detach(), "G_kl_loss": kl.detach(),
           "D_synthetic_logits": l_syn.detach(), "grads": {}}
    for n, g in zip(names, grads):
        g = torch.zeros_like(p[n]) if g is None else g.detach()
        res["grads"][n] = g
        p[n], st["m"][n], st["v"][n] = adam_tf(p[n], g, st["m"][n], st["v"][n], lr, cfg.g_beta1, st["g_t"])
    _apply_moving(p, chain)
    return res


def make_feed(cfg: Stage1Cfg, seed: int = 1234, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    B, hw = cfg.batch_size, cfg.output_size
    f = {}
    f["x"] = (torch.rand(B, hw, hw, cfg.image_c, generator=gen) * 2 - 1).to(dtype)
    f["x_mismatch"] = (torch.rand(B, hw, hw, cfg.image_c, generator=gen) * 2 - 1).to(dtype)
    f["cond"] = torch.randn(B, cfg.embed_dim, generator=gen).to(dtype)
    f["z"] = torch.randn(B, cfg.z_dim, generator=gen).to(dtype)
    f["tn_eps"] = truncated_normal_(torch.empty(B, cfg.compressed_embed_dim), 1.0, gen).to(dtype)
    f["tn_eps_g"] = truncated_normal_(torch.empty(B, cfg.compressed_embed_dim), 1.0, gen).to(dtype)
    return f


def iteration(p, st, feed, cfg: Stage1Cfg):
    return d_step(p, st, feed, cfg), g_step(p, st, feed, cfg)
