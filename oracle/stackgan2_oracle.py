"""CPU oracle for StackGAN stage-II (models/stackgan/stageII) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PyTorch-CPU restatement of the reference's TF-1.4 graph for one stage-II iteration: models/stackgan/stageII/model.py
(generator :134-201, discriminator :78-132, build_model :40-60) with the losses / optimizers of
models/stackgan/stageII/trainer.py:20-59.  PARITY UNPINNED (no fixtures in the reference, TF 1.4 not installable);
pins as for oracle/stackgan1_oracle.py.  Only tests/ may import this module.

What the graph does (model.py:50-58): z, embedding -> the FROZEN stage-I generator in TRAINING mode (batch statistics,
its own conditioning-noise draw; its moving statistics keep stepping because every optimizer runs under all
UPDATE_OPS, trainer.py:54-59) -> 64x64 image -> stage-II generator (encode 64->16, concat the conditioning vector,
four residual blocks of 4x4 stride-1 convs, four transposed-conv + 3x3 upsampling stages to 256x256, tanh) ->
stage-II discriminator (six stride-2 convs 256->4, two 4x4 stride-1 convs, a residual branch that is ADDED TO ITSELF
(model.py:117 `tf.add(net, net)`: the skip input is dropped), embedding concat, 1x1 conv, logits).
SAME padding of the 4x4 stride-1 convs is asymmetric in TF (1 before, 2 after).  Label smoothing 0.95 (trainer.py:27).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import stackgan1_oracle as S1
from . import wgancls_oracle as W
from .stackgan1_oracle import adam_tf, sigmoid_ce
from .wgancls_oracle import batch_norm, conv2d, conv2d_transpose, fc, is_trainable, kl_std_normal_loss, lrelu, truncated_normal_

REAL_LABEL = 0.95     # trainer.py:27
G2, D2 = "stageII_g_net/", "stageII_d_net/"


@dataclass
class Stage2Cfg:
    """models/stackgan/stageII/cfg/flowers.yml (overridable for small test nets; gf_dim must be a multiple of 32)."""
    batch_size: int = 8
    z_dim: int = 100
    embed_dim: int = 1024
    compressed_embed_dim: int = 128
    gf_dim: int = 128
    df_dim: int = 64
    output_size: int = 256
    image_c: int = 3
    d_beta1: float = 0.5
    g_beta1: float = 0.5
    alpha_mismatch: float = 0.5
    kl_coeff: float = 2.0
    lr: float = 2e-4
    # the frozen stage-I generator (models/stackgan/stageI/cfg)
    s1_gf_dim: int = 128

    def stage1(self):
        return S1.Stage1Cfg(batch_size=self.batch_size, z_dim=self.z_dim, embed_dim=self.embed_dim,
                            compressed_embed_dim=self.compressed_embed_dim, gf_dim=self.s1_gf_dim, df_dim=8)


def param_shapes(cfg: Stage2Cfg):
    """stage-I g_net (frozen) + stageII_g_net + stageII_d_net, TF default names in creation order."""
    sh = OrderedDict((k, v) for k, v in S1.param_shapes(cfg.stage1()).items() if k.startswith("g_net/"))
    gf, df, ce, c = cfg.gf_dim, cfg.df_dim, cfg.compressed_embed_dim, cfg.image_c

    def conv(scope, k, i, o):
        sh[scope + "/weights"] = (k, k, i, o)
        sh[scope + "/biases"] = (o,)

    def deconv(scope, k, i, o):
        sh[scope + "/weights"] = (k, k, o, i)
        sh[scope + "/biases"] = (o,)

    def bn(scope, ch):
        for leaf in ("beta", "gamma", "moving_mean", "moving_variance"):
            sh[scope + "/" + leaf] = (ch,)

    def dense(scope, i, o):
        sh[scope + "/kernel"] = (i, o)
        sh[scope + "/bias"] = (o,)

    def bname(i):
        return "BatchNorm" + ("" if i == 0 else "_%d" % i)

    def cname(i):
        return "Conv" + ("" if i == 0 else "_%d" % i)

    g = G2
    conv(g + "Conv", 3, c, gf)                                               # model.py:135
    conv(g + "Conv_1", 4, gf, gf * 2); bn(g + "BatchNorm", gf * 2)           # :137-138
    conv(g + "Conv_2", 4, gf * 2, gf * 4); bn(g + "BatchNorm_1", gf * 4)     # :140-141
    dense(g + "dense", cfg.embed_dim, ce); dense(g + "dense_1", cfg.embed_dim, ce)   # :64-68
    conv(g + "Conv_3", 3, gf * 4 + ce, gf * 4); bn(g + "BatchNorm_2", gf * 4)        # :191-192
    for r in range(4):                                                       # :148-157, four blocks (:194-197)
        conv(g + cname(4 + 2 * r), 4, gf * 4, gf * 4); bn(g + bname(3 + 2 * r), gf * 4)
        conv(g + cname(5 + 2 * r), 4, gf * 4, gf * 4); bn(g + bname(4 + 2 * r), gf * 4)
    chans = [gf * 4, gf * 2, gf, gf // 2, gf // 4]
    for u in range(4):                                                       # :159-176
        deconv(g + "Conv2d_transpose" + ("" if u == 0 else "_%d" % u), 4, chans[u], chans[u + 1])
        conv(g + cname(12 + u), 3, chans[u + 1], chans[u + 1]); bn(g + bname(11 + u), chans[u + 1])
    conv(g + "Conv_16", 3, gf // 4, c)                                       # :178
    d = D2
    conv(d + "Conv", 4, c, df)                                               # :83
    dch = [df, df * 2, df * 4, df * 8, df * 16, df * 32]
    for i in range(5):                                                       # :85-98
        conv(d + cname(1 + i), 4, dch[i], dch[i + 1]); bn(d + bname(i), dch[i + 1])
    conv(d + "Conv_6", 4, df * 32, df * 16); bn(d + "BatchNorm_5", df * 16)  # :100-101
    conv(d + "Conv_7", 4, df * 16, df * 8); bn(d + "BatchNorm_6", df * 8)    # :103-104
    conv(d + "Conv_8", 1, df * 8, df * 2); bn(d + "BatchNorm_7", df * 2)     # :108-109
    conv(d + "Conv_9", 3, df * 2, df * 2); bn(d + "BatchNorm_8", df * 2)     # :111-112
    conv(d + "Conv_10", 3, df * 2, df * 8); bn(d + "BatchNorm_9", df * 8)    # :114-115
    dense(d + "dense", cfg.embed_dim, ce)                                    # :122
    conv(d + "Conv_11", 1, df * 8 + ce, df * 8); bn(d + "BatchNorm_10", df * 8)   # :129-130
    conv(d + "Conv_12", cfg.output_size // 64, df * 8, 1)                    # :132
    return sh


def init_params(cfg: Stage2Cfg, seed=0, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("weights", "kernel"):
            p[name] = (torch.randn(*shape, generator=gen) * 0.02).to(dtype)
        elif leaf == "gamma":
            p[name] = (1.0 + 0.02 * torch.randn(*shape, generator=gen)).to(dtype)
        elif leaf == "moving_variance":
            p[name] = torch.ones(shape, dtype=dtype)
        else:
            p[name] = torch.zeros(shape, dtype=dtype)
    return p


def d_var_names(p):      # tf.trainable_variables('stageII_d_net'), model.py:59
    return [n for n in p if n.startswith(D2) and is_trainable(n)]


def g_var_names(p):      # tf.trainable_variables('stageII_g_net'), model.py:60
    return [n for n in p if n.startswith(G2) and is_trainable(n)]


def generator(p, image_nhwc, embed, tn_eps, cfg: Stage2Cfg, is_training=True, cond_noise=True, new_moving=None):
    """models/stackgan/stageII/model.py:180-201 (+ :134-178).  image: the stage-I output, NHWC 64x64x3."""
    g, gf = G2, cfg.gf_dim
    bnk = dict(train=is_training, new_moving=new_moving)
    x = image_nhwc.permute(0, 3, 1, 2)
    h = conv2d(p, g + "Conv", x, 3, 1, act=torch.relu)                                       # :135
    h = batch_norm(p, g + "BatchNorm", conv2d(p, g + "Conv_1", h, 4, 2), act=torch.relu, **bnk)     # :137-138
    enc = batch_norm(p, g + "BatchNorm_1", conv2d(p, g + "Conv_2", h, 4, 2), act=torch.relu, **bnk)  # :140-141
    mean = fc(p, g + "dense", embed, lrelu)                                                  # :64-66
    log_sigma = fc(p, g + "dense_1", embed, lrelu)                                           # :67-68
    c = mean + torch.exp(log_sigma) * tn_eps if cond_noise else mean                         # :71-76
    s4 = enc.shape[-1]
    cat = torch.cat([enc, c[:, :, None, None].expand(-1, -1, s4, s4)], 1)                    # :187-189
    h = batch_norm(p, g + "BatchNorm_2", conv2d(p, g + "Conv_3", cat, 3, 1), act=torch.relu, **bnk)  # :191-192
    for r in range(4):                                                                       # :194-197
        def cn(i):
            return g + ("Conv_%d" % i)

        def bn_(i):
            return g + ("BatchNorm_%d" % i)
        n = batch_norm(p, bn_(3 + 2 * r), conv2d(p, cn(4 + 2 * r), h, 4, 1), act=torch.relu, **bnk)   # :151-152
        n = batch_norm(p, bn_(4 + 2 * r), conv2d(p, cn(5 + 2 * r), n, 4, 1), **bnk)                   # :154-155
        h = torch.relu(h + n)                                                                          # :157
    for u in range(4):                                                                       # :160-176
        h = conv2d_transpose(p, g + "Conv2d_transpose" + ("" if u == 0 else "_%d" % u), h)
        h = batch_norm(p, g + "BatchNorm_%d" % (11 + u), conv2d(p, g + "Conv_%d" % (12 + u), h, 3, 1), act=torch.relu, **bnk)
    out = conv2d(p, g + "Conv_16", h, 3, 1, act=torch.tanh)                                   # :178
    return out.permute(0, 2, 3, 1), mean, log_sigma


def discriminator(p, x_nhwc, embed, cfg: Stage2Cfg, new_moving=None):
    """models/stackgan/stageII/model.py:78-132 (is_training=True always).  Returns logits [B,1,1,1]."""
    d = D2
    bnk = dict(train=True, new_moving=new_moving)
    h = conv2d(p, d + "Conv", x_nhwc.permute(0, 3, 1, 2), 4, 2, act=lrelu)                   # :83
    for i in range(5):                                                                       # :85-98
        h = batch_norm(p, d + ("BatchNorm" if i == 0 else "BatchNorm_%d" % i), conv2d(p, d + "Conv_%d" % (1 + i), h, 4, 2),
                       act=lrelu, **bnk)
    h = batch_norm(p, d + "BatchNorm_5", conv2d(p, d + "Conv_6", h, 4, 1), act=lrelu, **bnk)  # :100-101
    h7 = batch_norm(p, d + "BatchNorm_6", conv2d(p, d + "Conv_7", h, 4, 1), **bnk)            # :103-104
    n = batch_norm(p, d + "BatchNorm_7", conv2d(p, d + "Conv_8", h7, 1, 1), act=lrelu, **bnk)  # :108-109
    n = batch_norm(p, d + "BatchNorm_8", conv2d(p, d + "Conv_9", n, 3, 1), act=lrelu, **bnk)   # :111-112
    n = batch_norm(p, d + "BatchNorm_9", conv2d(p, d + "Conv_10", n, 3, 1), **bnk)             # :114-115
    h8 = lrelu(n + n)                                                                         # :117-118 (sic: h7 is dropped)
    e = fc(p, d + "dense", embed, lrelu)                                                      # :122
    s = h8.shape[-1]
    h9 = batch_norm(p, d + "BatchNorm_10", conv2d(p, d + "Conv_11", torch.cat([h8, e[:, :, None, None].expand(-1, -1, s, s)], 1),
                                                  1, 1), act=lrelu, **bnk)                    # :125-130
    k = cfg.output_size // 64
    w = p[d + "Conv_12/weights"].permute(3, 2, 0, 1)
    return F.conv2d(h9, w, p[d + "Conv_12/biases"], stride=k).permute(0, 2, 3, 1)             # :132


def new_state(p):
    st = {"d_t": 0, "g_t": 0, "m": {}, "v": {}}
    for n, w in p.items():
        if is_trainable(n) and not n.startswith("g_net/"):
            st["m"][n] = torch.zeros_like(w)
            st["v"][n] = torch.zeros_like(w)
    return st


def _forward_all(pl, feed, tn1, tn2, cfg, chain):
    m1 = {}
    img64, _, _ = S1.generator(pl, feed["z"], feed["cond"], tn1, cfg.stage1(), new_moving=m1)    # model.py:50 (training mode)
    chain.append(m1)
    mg = {}
    G, mean, log_sigma = generator(pl, img64, feed["cond"], tn2, cfg, new_moving=mg)              # :52
    chain.append(mg)
    logits = []
    for img in (G, feed["x"], feed["x_mismatch"]):                                                # :53-56
        cur = dict(pl)
        for upd in chain:
            cur.update({k: v for k, v in upd.items() if k.startswith(D2)})
        md = {}
        logits.append(discriminator(cur, img, feed["cond"], cfg, new_moving=md))
        chain.append(md)
    return img64, G, mean, log_sigma, logits


def d_step(p, st, feed, cfg: Stage2Cfg, lr=None):
    lr = cfg.lr if lr is None else lr
    names = d_var_names(p)
    pl = {n: (w.detach().clone().requires_grad_(True) if n in names else w.detach()) for n, w in p.items()}
    chain = []
    img64, G, mean, log_sigma, (l_syn, l_real, l_mis) = _forward_all(pl, feed, feed["tn_s1"], feed["tn_eps"], cfg, chain)
    res = OrderedDict()
    res["stage1_G"], res["G"] = img64.detach(), G.detach()
    res["D_synthetic_logits"], res["D_real_match_logits"], res["D_real_mismatch_logits"] = l_syn.detach(), l_real.detach(), l_mis.detach()
    res["D_synthetic_loss"] = sigmoid_ce(l_syn, 0.0)
    res["D_real_match_loss"] = sigmoid_ce(l_real, REAL_LABEL)
    res["D_real_mismatch_loss"] = sigmoid_ce(l_mis, 0.0)
    a = cfg.alpha_mismatch
    d_loss = res["D_real_match_loss"] + a * res["D_real_mismatch_loss"] + (1.0 - a) * res["D_synthetic_loss"]
    grads = torch.autograd.grad(d_loss, [pl[n] for n in names], allow_unused=True)
    st["d_t"] += 1
    res = OrderedDict((k, v.detach()) for k, v in res.items())
    res["D_loss"] = d_loss.detach()
    res["grads"] = {}
    for n, g in zip(names, grads):
        g = torch.zeros_like(p[n]) if g is None else g.detach()
        res["grads"][n] = g
        p[n], st["m"][n], st["v"][n] = adam_tf(p[n], g, st["m"][n], st["v"][n], lr, cfg.d_beta1, st["d_t"])
    S1._apply_moving(p, chain)
    return res


def g_step(p, st, feed, cfg: Stage2Cfg, lr=None):
    lr = cfg.lr if lr is None else lr
    names = g_var_names(p)
    pl = {n: (w.detach().clone().requires_grad_(True) if n in names else w.detach()) for n, w in p.items()}
    chain = []
    img64, G, mean, log_sigma, (l_syn, _, _) = _forward_all(pl, feed, feed["tn_s1_g"], feed["tn_eps_g"], cfg, chain)
    kl = kl_std_normal_loss(mean, log_sigma)
    gan = sigmoid_ce(l_syn, 1.0)
    g_loss = gan + cfg.kl_coeff * kl
    grads = torch.autograd.grad(g_loss, [pl[n] for n in names], allow_unused=True)
    st["g_t"] += 1
    res = {"G": G.detach(), "G_loss": g_loss.detach(), "G_gan_loss": gan.detach(), "G_kl_loss": kl.detach(),
           "D_synthetic_logits": l_syn.detach(), "grads": {}}
    for n, g in zip(names, grads):
        g = torch.zeros_like(p[n]) if g is None else g.detach()
        res["grads"][n] = g
        p[n], st["m"][n], st["v"][n] = adam_tf(p[n], g, st["m"][n], st["v"][n], lr, cfg.g_beta1, st["g_t"])
    S1._apply_moving(p, chain)
    return res


def make_feed(cfg: Stage2Cfg, seed=1234, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    B, hw = cfg.batch_size, cfg.output_size
    f = {}
    f["x"] = (torch.rand(B, hw, hw, cfg.image_c, generator=gen) * 2 - 1).to(dtype)
    f["x_mismatch"] = (torch.rand(B, hw, hw, cfg.image_c, generator=gen) * 2 - 1).to(dtype)
    f["cond"] = torch.randn(B, cfg.embed_dim, generator=gen).to(dtype)
    f["z"] = torch.randn(B, cfg.z_dim, generator=gen).to(dtype)
    for k in ("tn_s1", "tn_eps", "tn_s1_g", "tn_eps_g"):      # stage-I / stage-II conditioning noise of the D run, of the G run
        f[k] = truncated_normal_(torch.empty(B, cfg.compressed_embed_dim), 1.0, gen).to(dtype)
    return f


def iteration(p, st, feed, cfg: Stage2Cfg):
    return d_step(p, st, feed, cfg), g_step(p, st, feed, cfg)
