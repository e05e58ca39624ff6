"""CPU oracle for the conditional progressive-growing WGAN (models/pggan) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PyTorch-CPU restatement of the reference's TF-1.4 graph for one iteration of one stage of models/pggan/pggan.py
(generator :279-316, discriminator :251-277, to_rgb :367-371, from_rgb :343-345, losses / optimizers :94-130) with the
op wrappers of utils/ops.py (layer_norm :74-81, pool :100-101, upscale :109-111, conv2d :58-63, fc :84-87).
PARITY UNPINNED (no fixtures in the reference, TF 1.4 not installable); pins as for oracle/wgancls_oracle.py
(tests/test_pggan_cpu.py: naive loops for layer-norm / pool / upscale / the 2x2 SAME conv, fp64 finite differences of
the gradient penalty, torch.autograd with create_graph=True).  Only tests/ may import this module.

TF semantics restated (recalled from TF 1.4; cannot be checked against TF here):
  * tf.contrib.layers.layer_norm(begin_norm_axis=1, begin_params_axis=-1): per SAMPLE mean / biased variance over all
    non-batch axes, variance_epsilon 1e-12, gamma (ones) / beta (zeros) of shape [C] (rank 2: [features]).
  * tf.nn.pool AVG 2x2 stride 2 SAME on even extents = plain 2x2 mean; resize_nearest_neighbor x2: out[i,j] = in[i//2,j//2].
  * conv2d 2x2 stride 1 SAME: pad_total = 1 -> 0 before, 1 after (bottom / right).
  * the graph is NHWC throughout (utils/ops.py default df=NHWC); the dense output is reshaped to [-1, 4, 4, C].
  * the graph rebinds self.epsilon to a tf.random_uniform tensor (:68) and the trainer feeds THAT tensor (:205), so the
    fed U(0,1) draw is the value used: an explicit input here (feed['epsilon']).
  * alpha_tra = iter / steps is assigned under a control dependency of D_optim only (:84,119); the forward ops read the
    variable unordered with respect to the assign.  Deterministic choice (also the product's): the assign happens
    first, both runs of iteration `iter` use alpha = iter / steps.
  * D_loss = -wdist - wdist2 + 200 (gp + gp2) (:108), G_loss = -D_fake + 5 KL (:109), Adam(2e-6, beta1 0, beta2 0.99)
    for both (:111-112; the learning_rate placeholder is fed and unused).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from .wgancls_oracle import adam_tf, fc, gradient_penalty, he_trunc_normal, kl_std_normal_loss, lrelu, truncated_normal_

GP_WEIGHT = 200.0     # pggan.py:108
KL_COEFF = 5.0        # pggan.py:109
LN_EPS = 1e-12        # tf.contrib.layers.layer_norm variance_epsilon


@dataclass
class PgganCfg:
    """Constants of models/pggan/pggan.py:28-38 and the channel schedule :339-343 (base / cap overridable for tests)."""
    batch_size: int = 16
    stage: int = 1
    trans: bool = False
    z_dim: int = 128
    embed_dim: int = 1024
    compr_embed_dim: int = 128
    nf_base: int = 1024       # get_nf(i) = min(base // 2**i * 4, cap), get_dnf(i) = min(base // 2**i * 2, cap)
    nf_cap: int = 512
    rgb_mid: int = 9          # to_rgb: conv 2x2 -> 9 channels -> conv 1x1 -> 3 (:369-370)
    d_embed: int = 128        # the discriminator's compressed embedding, literal 128 (:271)
    lr: float = 0.000002      # :111-112
    beta1: float = 0.0
    beta2: float = 0.99

    @property
    def out_size(self):       # :30
        return 4 * 2 ** (self.stage - 1)

    def nf(self, i):
        return min(self.nf_base // (2 ** i) * 4, self.nf_cap)

    def dnf(self, i):
        return min(self.nf_base // (2 ** i) * 2, self.nf_cap)


def param_shapes(cfg: PgganCfg):
    """Variables of the graph of (stage, trans), TF default names."""
    sh = OrderedDict()
    E, ce, Z = cfg.embed_dim, cfg.compr_embed_dim, cfg.z_dim

    def conv(scope, k, i, o):
        sh[scope + "/weights"] = (k, k, i, o)
        sh[scope + "/biases"] = (o,)

    def dense(scope, i, o):
        sh[scope + "/kernel"] = (i, o)
        sh[scope + "/bias"] = (o,)

    def ln(scope, ch):
        sh[scope + "/beta"] = (ch,)
        sh[scope + "/gamma"] = (ch,)

    g = "g_net/"
    s0 = g + "conv_stage_0/"
    dense(s0 + "dense", E, ce); dense(s0 + "dense_1", E, ce)                  # :343-347 via :284
    dense(s0 + "dense_2", Z + ce, 16 * cfg.nf(0)); ln(s0 + "LayerNorm", 16 * cfg.nf(0))     # :288-289
    conv(s0 + "Conv", 3, cfg.nf(0), cfg.nf(0)); ln(s0 + "LayerNorm_1", cfg.nf(0))           # :292-293
    conv(s0 + "Conv_1", 3, cfg.nf(0), cfg.nf(0)); ln(s0 + "LayerNorm_2", cfg.nf(0))         # :294-295
    for i in range(1, cfg.stage):                                             # :298-309
        s = g + "conv_stage_%d/" % i
        conv(s + "Conv", 3, cfg.nf(i - 1), cfg.nf(i)); ln(s + "LayerNorm", cfg.nf(i))
        conv(s + "Conv_1", 3, cfg.nf(i), cfg.nf(i)); ln(s + "LayerNorm_1", cfg.nf(i))
    rgb_stages = [cfg.stage - 1] + ([cfg.stage - 2] if cfg.trans else [])
    for k in sorted(rgb_stages):                                              # :367-371
        s = g + "rgb_stage_%d/" % k
        conv(s + "Conv", 2, cfg.nf(k), cfg.rgb_mid); conv(s + "Conv_1", 1, cfg.rgb_mid, 3)
    d = "d_net/"
    for k in sorted(rgb_stages):                                              # :343-345
        conv(d + "rgb_stage_%d/Conv" % k, 1, 3, cfg.dnf(k))
    for i in range(cfg.stage - 1, 0, -1):                                     # :261-267
        s = d + "conv_stage_%d/" % i
        conv(s + "Conv", 3, cfg.dnf(i), cfg.dnf(i)); conv(s + "Conv_1", 3, cfg.dnf(i), cfg.dnf(i - 1))
    s0 = d + "conv_stage_0/"
    dense(s0 + "dense", E, cfg.d_embed)                                       # :271
    conv(s0 + "Conv", 3, cfg.dnf(0) + cfg.d_embed, cfg.dnf(0))                # :273
    conv(s0 + "Conv_1", 4, cfg.dnf(0), cfg.dnf(0))                            # :274
    dense(s0 + "dense_1", cfg.dnf(0), 1)                                      # :275
    return sh


def init_params(cfg: PgganCfg, seed=0, dtype=torch.float32):
    """He truncated-normal kernels (utils/ops.py:60,86: sigma = sqrt(1.3 * 2 / fan_in), fan_in = kh*kw*Cin), zero
    biases / beta, gamma = 1."""
    gen = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("weights", "kernel"):
            p[name] = he_trunc_normal(shape, gen, dtype)
        elif leaf == "gamma":
            p[name] = torch.ones(shape, dtype=dtype)
        else:
            p[name] = torch.zeros(shape, dtype=dtype)
    return p


def d_var_names(p):
    return [n for n in p if n.startswith("d_net/")]


def g_var_names(p):
    return [n for n in p if n.startswith("g_net/")]


# ----------------------------------------------------------------------------------------------------- ops (NHWC)
def conv2d(p, scope, x, k, padding="SAME", act=None):
    """utils/ops.py:58-63 with df=NHWC, stride 1; weight HWIO."""
    w = p[scope + "/weights"].permute(3, 2, 0, 1)
    x = x.permute(0, 3, 1, 2)
    if padding.upper() == "SAME":
        before = (k - 1) // 2
        x = F.pad(x, (before, k - 1 - before, before, k - 1 - before))
    y = F.conv2d(x, w, p[scope + "/biases"]).permute(0, 2, 3, 1)
    return act(y) if act is not None else y


def layer_norm(p, scope, x, act=None):
    """utils/ops.py:74-81."""
    dims = tuple(range(1, x.dim()))
    mean = x.mean(dims, keepdim=True)
    var = ((x - mean) ** 2).mean(dims, keepdim=True)
    y = (x - mean) * torch.rsqrt(var + LN_EPS) * p[scope + "/gamma"] + p[scope + "/beta"]
    return act(y) if act is not None else y


def pool(x):
    """utils/ops.py:100-101 (AVG, 2)."""
    n, h, w, c = x.shape
    return x.reshape(n, h // 2, 2, w // 2, 2, c).mean((2, 4))


def upscale(x):
    """utils/ops.py:109-111."""
    return x.repeat_interleave(2, 1).repeat_interleave(2, 2)


def to_rgb(p, x, k):
    s = "g_net/rgb_stage_%d/" % k
    return conv2d(p, s + "Conv_1", conv2d(p, s + "Conv", x, 2, act=torch.relu), 1)      # :367-371


def from_rgb(p, x, k):
    return conv2d(p, "d_net/rgb_stage_%d/Conv" % k, x, 1, act=lrelu)                    # :343-345


def generator(p, z, embed, tn_eps, cfg: PgganCfg, alpha, cond_noise=True):
    """pggan.py:279-316.  Returns (image NHWC [B, S, S, 3] -- no tanh --, mean, log_sigma)."""
    s0 = "g_net/conv_stage_0/"
    mean = fc(p, s0 + "dense", embed, lrelu)                                              # :343-347
    log_sigma = fc(p, s0 + "dense_1", embed, lrelu)
    c = mean + torch.exp(log_sigma) * tn_eps if cond_noise else mean                      # :349-354
    x = fc(p, s0 + "dense_2", torch.cat([z, c], 1))                                       # :287-288
    x = layer_norm(p, s0 + "LayerNorm", x).reshape(-1, 4, 4, cfg.nf(0))                   # :289-290
    x = layer_norm(p, s0 + "LayerNorm_1", conv2d(p, s0 + "Conv", x, 3), torch.relu)       # :292-293
    x = layer_norm(p, s0 + "LayerNorm_2", conv2d(p, s0 + "Conv_1", x, 3), torch.relu)     # :294-295
    x_iden = None
    for i in range(1, cfg.stage):
        if i == cfg.stage - 1 and cfg.trans:                                              # :300-302
            x_iden = upscale(to_rgb(p, x, cfg.stage - 2))
        s = "g_net/conv_stage_%d/" % i
        x = upscale(x)                                                                    # :305
        x = layer_norm(p, s + "LayerNorm", conv2d(p, s + "Conv", x, 3), torch.relu)
        x = layer_norm(p, s + "LayerNorm_1", conv2d(p, s + "Conv_1", x, 3), torch.relu)
    x = to_rgb(p, x, cfg.stage - 1)                                                       # :311
    if cfg.trans:
        x = (1.0 - alpha) * x_iden + alpha * x                                            # :313-314
    return x, mean, log_sigma


def discriminator(p, inp, cond, cfg: PgganCfg, alpha):
    """pggan.py:251-277.  inp NHWC [B, S, S, 3] -> logits [B, 1, 1, 1]."""
    x_iden = None
    if cfg.trans:
        x_iden = from_rgb(p, pool(inp), cfg.stage - 2)                                    # :255-257
    x = from_rgb(p, inp, cfg.stage - 1)                                                   # :259
    for i in range(cfg.stage - 1, 0, -1):                                                 # :261-267
        s = "d_net/conv_stage_%d/" % i
        x = conv2d(p, s + "Conv", x, 3, act=lrelu)
        x = conv2d(p, s + "Conv_1", x, 3, act=lrelu)
        x = pool(x)
        if i == cfg.stage - 1 and cfg.trans:
            x = alpha * x + (1.0 - alpha) * x_iden
    s0 = "d_net/conv_stage_0/"
    e = fc(p, s0 + "dense", cond, lrelu)                                                  # :271
    x = torch.cat([x, e[:, None, None, :].expand(-1, 4, 4, -1)], 3)                       # :272, :318-322
    x = conv2d(p, s0 + "Conv", x, 3, act=lrelu)                                           # :273
    x = conv2d(p, s0 + "Conv_1", x, 4, padding="VALID", act=lrelu)                        # :274
    return fc(p, s0 + "dense_1", x)                                                       # :275  [B,1,1,1]


def d_forward_losses(p, feed, cfg: PgganCfg, alpha, create_graph):
    """pggan.py:62-77 and :94-108."""
    G, mean, log_sigma = generator(p, feed["z"], feed["cond"], feed["tn_eps"], cfg, alpha)
    x, xm, cond = feed["x"], feed["x_mismatch"], feed["cond"]
    Dg = discriminator(p, G, cond, cfg, alpha)
    Dx = discriminator(p, x, cond, cfg, alpha)
    Dxmi = discriminator(p, xm, cond, cfg, alpha)
    eps = feed["epsilon"].reshape(-1, 1, 1, 1)
    x_hat = (eps * G + (1.0 - eps) * x).detach().requires_grad_(True)     # :69 (the D update does not reach g_vars)
    cond_inp = (cond + 0.0).detach().requires_grad_(True)                 # :70
    Dxh = discriminator(p, x_hat, cond_inp, cfg, alpha)
    gx, gc = torch.autograd.grad(Dxh.sum(), [x_hat, cond_inp], create_graph=create_graph)
    out = OrderedDict()
    out["G"], out["x_hat"] = G, x_hat
    out["Dg_logit"], out["Dx_logit"], out["Dxmi_logit"], out["Dx_hat_logit"] = Dg, Dx, Dxmi, Dxh
    out["grad_x_hat"], out["grad_cond"] = gx, gc
    out["D_loss_real"], out["D_loss_fake"], out["D_loss_mismatch"] = Dx.mean(), Dg.mean(), Dxmi.mean()   # :95-97
    out["wdist"] = out["D_loss_real"] - out["D_loss_fake"]
    out["wdist2"] = out["D_loss_real"] - out["D_loss_mismatch"]
    out["reg_loss"] = (Dxmi ** 2).mean()
    out["real_gp"], out["slopes"] = gradient_penalty(gx, (1, 2, 3))       # :85-88
    out["real_gp2"], out["slopes2"] = gradient_penalty(gc, (1,))          # :90-93
    out["D_loss"] = -out["wdist"] - out["wdist2"] + GP_WEIGHT * (out["real_gp"] + out["real_gp2"])       # :108
    return out


def new_state(p):
    st = {"d_t": 0, "g_t": 0, "m": {}, "v": {}}
    for n, w in p.items():
        st["m"][n] = torch.zeros_like(w)
        st["v"][n] = torch.zeros_like(w)
    return st


def alpha_of(idx, steps):
    return float(idx) / float(steps)      # :78-79


def d_step(p, st, feed, cfg: PgganCfg, alpha):
    """sess.run([D_optim, D_loss]) -- pggan.py:218."""
    names = d_var_names(p)
    pl = {n: (w.detach().clone().requires_grad_(True) if n in names else w.detach()) for n, w in p.items()}
    out = d_forward_losses(pl, feed, cfg, alpha, create_graph=True)
    grads = torch.autograd.grad(out["D_loss"], [pl[n] for n in names], allow_unused=True)
    st["d_t"] += 1
    res = {k: v.detach() for k, v in out.items()}
    res["grads"] = {}
    for n, g in zip(names, grads):
        g = torch.zeros_like(p[n]) if g is None else g.detach()
        res["grads"][n] = g
        p[n], st["m"][n], st["v"][n] = adam_tf(p[n], g, st["m"][n], st["v"][n], cfg.lr, cfg.beta1, cfg.beta2, st["d_t"])
    return res


def g_step(p, st, feed, cfg: PgganCfg, alpha):
    """sess.run([G_optim, G_loss]) -- pggan.py:219 (fresh conditioning noise: feed['tn_eps_g'])."""
    names = g_var_names(p)
    pl = {n: (w.detach().clone().requires_grad_(True) if n in names else w.detach()) for n, w in p.items()}
    G, mean, log_sigma = generator(pl, feed["z"], feed["cond"], feed["tn_eps_g"], cfg, alpha)
    Dg = discriminator(pl, G, feed["cond"], cfg, alpha)
    kl = kl_std_normal_loss(mean, log_sigma)
    g_loss = -Dg.mean() + KL_COEFF * kl                                   # :109
    grads = torch.autograd.grad(g_loss, [pl[n] for n in names], allow_unused=True)
    st["g_t"] += 1
    res = {"G": G.detach(), "G_loss": g_loss.detach(), "G_kl_loss": kl.detach(), "Dg_logit": Dg.detach(), "grads": {}}
    for n, g in zip(names, grads):
        g = torch.zeros_like(p[n]) if g is None else g.detach()
        res["grads"][n] = g
        p[n], st["m"][n], st["v"][n] = adam_tf(p[n], g, st["m"][n], st["v"][n], cfg.lr, cfg.beta1, cfg.beta2, st["g_t"])
    return res


def make_feed(cfg: PgganCfg, seed=1234, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    B, S = cfg.batch_size, cfg.out_size
    f = {}
    f["x"] = (torch.rand(B, S, S, 3, generator=gen) * 2 - 1).to(dtype)
    f["x_mismatch"] = (torch.rand(B, S, S, 3, generator=gen) * 2 - 1).to(dtype)
    f["cond"] = torch.randn(B, cfg.embed_dim, generator=gen).to(dtype)
    f["z"] = torch.randn(B, cfg.z_dim, generator=gen).to(dtype)
    f["epsilon"] = torch.rand(B, generator=gen).to(dtype)
    f["tn_eps"] = truncated_normal_(torch.empty(B, cfg.compr_embed_dim), 1.0, gen).to(dtype)
    f["tn_eps_g"] = truncated_normal_(torch.empty(B, cfg.compr_embed_dim), 1.0, gen).to(dtype)
    return f


def iteration(p, st, feed, cfg: PgganCfg, alpha):
    return d_step(p, st, feed, cfg, alpha), g_step(p, st, feed, cfg, alpha)
