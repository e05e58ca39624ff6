"""StackGAN stage-II host logic on CPU: StageIIEngine (text-to-image_b200/engine_stage2.py)
driven by the CPU restatement of the kernels must reproduce the oracle's D run and G run -- the frozen stage-I
generator in training mode, the 4x4 stride-1 convs with TF's asymmetric SAME padding, the concat gradient split into
two output-channel windows, the residual branch added to itself (affine_scale = 2), three discriminator calls with
their own BatchNorm statistics, cross-entropy losses with label smoothing 0.95, Adam(0.5, 0.999), moving statistics.
np_ = 0: exact fp64 storage (any schedule / layout / formula error shows at 1e-9); np_ = 2: split-bf16 rounding."""
import numpy as np
import pytest
import torch

import fake_kernels as fk
from oracle import stackgan2_oracle as S2

TINY = dict(batch_size=2, z_dim=12, embed_dim=32, compressed_embed_dim=8, gf_dim=32, df_dim=8, s1_gf_dim=8)


def make_engine(cfg, np_=2, batch=None):
    from t2i_b200.engine_stage2 import StageIIEngine
    kw = dict(act_dtype=torch.float64, f32_dtype=torch.float64) if np_ == 0 else {}
    return StageIIEngine(fk, "cpu", batch or cfg.batch_size, max(np_, 1), cfg.z_dim, cfg.embed_dim,
                         cfg.compressed_embed_dim, cfg.gf_dim, cfg.df_dim, cfg.d_beta1, cfg.g_beta1, cfg.alpha_mismatch,
                         cfg.kl_coeff, s1_gf=cfg.s1_gf_dim, **kw)


def boosted_params(cfg, seed=0):
    """reference-style init is tiny (sigma 0.02); scale it up so that every term of the iteration is exercised"""
    p = S2.init_params(cfg, seed, torch.float64)
    g = torch.Generator().manual_seed(5)
    for n in p:
        if n.endswith("weights") or n.endswith("kernel"):
            fan_in = p[n].shape[-2] * int(np.prod(p[n].shape[:-2])) if p[n].dim() > 1 else 1
            p[n] = torch.randn(p[n].shape, generator=g, dtype=torch.float64) * (2.0 / fan_in) ** 0.5
        if n.endswith("biases") or n.endswith("bias") or n.endswith("beta"):
            p[n] = torch.randn(p[n].shape, generator=g, dtype=torch.float64) * 0.1
        if n.endswith("gamma"):
            p[n] = 1 + 0.2 * torch.randn(p[n].shape, generator=g, dtype=torch.float64)
    return p


def rel(a, b):
    a, b = torch.as_tensor(a).double().reshape(-1), torch.as_tensor(b).double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def _bias_before_bn(n):
    """a conv bias directly in front of a training-mode BatchNorm has an exactly-zero gradient"""
    if not n.endswith("biases"):
        return False
    if n.startswith(S2.G2):
        return "Conv2d_transpose" not in n and n not in (S2.G2 + "Conv/biases", S2.G2 + "Conv_16/biases")
    return n not in (S2.D2 + "Conv/biases", S2.D2 + "Conv_12/biases")


def test_param_layout_roundtrip():
    cfg = S2.Stage2Cfg(**TINY)
    p = boosted_params(cfg)
    eng = make_engine(cfg)
    eng.set_params_tf(p)
    q = eng.get_params_tf()
    assert set(q) == set(p)
    for n in p:
        np.testing.assert_allclose(q[n].numpy(), p[n].float().numpy(), rtol=0, atol=0, err_msg=n)


@pytest.mark.parametrize("np_", [0, 2])
def test_iteration_matches_oracle(np_):
    cfg = S2.Stage2Cfg(**TINY)
    p = boosted_params(cfg)
    feed = S2.make_feed(cfg, 21, torch.float64)
    eng = make_engine(cfg, np_)
    eng.set_params_tf(p)
    st = S2.new_state(p)
    rd = S2.d_step(p, st, feed, cfg)
    eng.load_feed(x=feed["x"], x_mismatch=feed["x_mismatch"], cond=feed["cond"], z=feed["z"], tn_eps=feed["tn_eps"],
                  tn_s1=feed["tn_s1"])
    eng.d_step(cfg.lr)
    tol = 1e-9 if np_ == 0 else 5e-4
    B = cfg.batch_size
    assert rel(eng.g["img64"], rd["stage1_G"]) < tol
    assert rel(eng.d["img"][:B], rd["G"]) < (tol if np_ == 0 else 5e-3)
    lg = eng.d["logit"]
    for k, n in enumerate(["D_synthetic_logits", "D_real_match_logits", "D_real_mismatch_logits"]):
        assert rel(lg[k * B:(k + 1) * B], rd[n]) < (tol if np_ == 0 else 1e-2), n
    sc = eng.scalars_dict()
    for k in ["D_loss", "D_synthetic_loss", "D_real_match_loss", "D_real_mismatch_loss"]:
        assert abs(sc[k] - float(rd[k])) < 10 * tol * max(1.0, abs(float(rd[k]))), (k, sc[k], float(rd[k]))
    grads = eng.get_grads_tf()
    worst = max((rel(grads[n], rd["grads"][n]), n) for n in rd["grads"] if float(rd["grads"][n].abs().max()) > 1e-14
                and not _bias_before_bn(n))
    assert worst[0] < (1e-8 if np_ == 0 else 0.1), worst
    if np_ == 0:
        newp = eng.get_params_tf()
        for n in rd["grads"]:
            g = rd["grads"][n]
            sure = g.abs() > 1e-3 * g.abs().max()
            if bool(sure.any()) and not _bias_before_bn(n):
                assert float((newp[n].double() - p[n])[sure].abs().max()) < 1e-9, n
        for n in p:       # UPDATE_OPS: stage-I g_net and stageII_g_net once, stageII_d_net once per call
            if "moving" in n:
                assert rel(newp[n], p[n]) < 1e-9, n
        eng.set_params_tf({k: v for k, v in p.items()})
    # ---- G run
    rg = S2.g_step(p, st, feed, cfg)
    eng.load_feed(tn_eps=feed["tn_eps_g"], tn_s1=feed["tn_s1_g"])
    eng.g_step(cfg.lr)
    sc = eng.scalars_dict()
    for k in ["G_loss", "G_gan_loss", "G_kl_loss"]:
        assert abs(sc[k] - float(rg[k])) < 10 * tol * max(1.0, abs(float(rg[k]))), (k, sc[k], float(rg[k]))
    grads = eng.get_grads_tf()
    worst = max((rel(grads[n], rg["grads"][n]), n) for n in rg["grads"] if float(rg["grads"][n].abs().max()) > 1e-14
                and not _bias_before_bn(n))
    assert worst[0] < (1e-7 if np_ == 0 else 0.3), worst   # batch 2: one flipped LeakyReLU unit moves everything by percents
    if np_ == 0:
        newp = eng.get_params_tf()
        for n in rg["grads"]:
            g = rg["grads"][n]
            sure = g.abs() > 1e-3 * g.abs().max()
            if bool(sure.any()) and not _bias_before_bn(n):
                assert float((newp[n].double() - p[n])[sure].abs().max()) < 1e-9, n
        for n in p:
            if "moving" in n and not n.startswith(S2.D2):
                assert rel(newp[n], p[n]) < 1e-9, n
