"""Pins for the CPU oracle (parity is otherwise unpinned: the reference ships no tests).

Cross-checks of oracle/wgancls_oracle.py against (a) the independent naive NumPy loops of
oracle/naive.py, (b) fp64 finite differences / closed forms, (c) the committed goldens.
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import naive
from oracle import wgancls_oracle as O

TINY = dict(batch_size=4, z_dim=8, embed_dim=32, compressed_embed_dim=8, gf_dim=8, df_dim=8)
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def tiny_cfg(**kw):
    d = dict(TINY)
    d.update(kw)
    return O.OracleCfg(**d)


@pytest.mark.parametrize("k,s,pad,h", [(4, 2, "SAME", 8), (3, 1, "SAME", 6), (1, 1, "VALID", 4), (4, 4, "VALID", 4)])
def test_conv_matches_naive(k, s, pad, h):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, h, h, 5, generator=g, dtype=torch.float64)
    w = torch.randn(k, k, 5, 7, generator=g, dtype=torch.float64)
    b = torch.randn(7, generator=g, dtype=torch.float64)
    p = {"c/weights": w, "c/biases": b}
    y = O.conv2d(p, "c", x.permute(0, 3, 1, 2), k, s, pad).permute(0, 2, 3, 1)
    ref = naive.conv2d_nhwc(x.numpy(), w.numpy(), b.numpy(), s, pad)
    np.testing.assert_allclose(y.numpy(), ref, rtol=1e-12, atol=1e-12)


def test_deconv_matches_naive_and_is_conv_transpose():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 4, 4, 6, generator=g, dtype=torch.float64)
    w = torch.randn(4, 4, 3, 6, generator=g, dtype=torch.float64)   # [kh,kw,Cout,Cin]
    b = torch.randn(3, generator=g, dtype=torch.float64)
    p = {"t/weights": w, "t/biases": b}
    y = O.conv2d_transpose(p, "t", x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    ref = naive.conv2d_transpose_nhwc(x.numpy(), w.numpy(), b.numpy(), 2)
    np.testing.assert_allclose(y.numpy(), ref, rtol=1e-12, atol=1e-12)
    # duality: <deconv(x), u> == <x, conv(u)> with the same kernel read as HWIO [kh,kw,Cout->in,Cin->out]
    u = torch.randn(2, 8, 8, 3, generator=g, dtype=torch.float64)
    pc = {"c/weights": w, "c/biases": torch.zeros(6, dtype=torch.float64)}
    cu = O.conv2d(pc, "c", u.permute(0, 3, 1, 2), 4, 2).permute(0, 2, 3, 1)
    lhs = ((y - b) * u).sum()
    rhs = (x * cu).sum()
    assert abs(lhs - rhs) < 1e-9 * abs(lhs)


def test_batch_norm_train_and_moving():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(6, 5, 3, 3, generator=g, dtype=torch.float64) * 2 + 1
    p = {"b/gamma": torch.rand(5, generator=g, dtype=torch.float64) + .5, "b/beta": torch.randn(5, generator=g, dtype=torch.float64),
         "b/moving_mean": torch.zeros(5, dtype=torch.float64), "b/moving_variance": torch.ones(5, dtype=torch.float64)}
    nm = {}
    y = O.batch_norm(p, "b", x, True, new_moving=nm)
    xm = x.permute(1, 0, 2, 3).reshape(5, -1)
    mean, var = xm.mean(1), ((xm - xm.mean(1, keepdim=True)) ** 2).mean(1)
    ref = ((xm - mean[:, None]) / torch.sqrt(var[:, None] + 1e-5)) * p["b/gamma"][:, None] + p["b/beta"][:, None]
    np.testing.assert_allclose(y.permute(1, 0, 2, 3).reshape(5, -1).numpy(), ref.numpy(), rtol=1e-12, atol=1e-12)
    n = xm.shape[1]
    np.testing.assert_allclose(nm["b/moving_mean"].numpy(), (0.1 * mean).numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(nm["b/moving_variance"].numpy(), (0.9 + 0.1 * var * n / (n - 1)).numpy(), rtol=1e-12, atol=1e-12)
    # rank-2: per feature over the batch (generator fc0, model.py:176)
    x2 = torch.randn(7, 5, generator=g, dtype=torch.float64)
    y2 = O.batch_norm(p, "b", x2, True)
    ref2 = (x2 - x2.mean(0)) / torch.sqrt(x2.var(0, unbiased=False) + 1e-5) * p["b/gamma"] + p["b/beta"]
    np.testing.assert_allclose(y2.numpy(), ref2.numpy(), rtol=1e-12, atol=1e-12)
    # inference uses the moving statistics
    y3 = O.batch_norm(p, "b", x, False)
    np.testing.assert_allclose(y3.numpy(), (x / math.sqrt(1 + 1e-5) * p["b/gamma"].view(1, -1, 1, 1)
                                            + p["b/beta"].view(1, -1, 1, 1)).numpy(), rtol=1e-12, atol=1e-12)


def test_adam_tf_closed_form():
    th, g = torch.tensor([1.0, -2.0], dtype=torch.float64), torch.tensor([0.5, -0.25], dtype=torch.float64)
    m = v = torch.zeros(2, dtype=torch.float64)
    t1, m1, v1 = O.adam_tf(th, g, m, v, 1e-4, 0.0, 0.9, 1)
    # beta1 = 0: m = g; v = .1 g^2; lr_t = lr*sqrt(.1)  => step = lr*sqrt(.1)*g/(sqrt(.1)|g| + 1e-8)
    ref = th - 1e-4 * math.sqrt(0.1) * g / (math.sqrt(0.1) * g.abs() + 1e-8)
    np.testing.assert_allclose(t1.numpy(), ref.numpy(), rtol=1e-14)
    t2, m2, v2 = O.adam_tf(t1, g, m1, v1, 1e-4, 0.0, 0.9, 2)
    v_ref = 0.9 * 0.1 * g * g + 0.1 * g * g
    np.testing.assert_allclose(v2.numpy(), v_ref.numpy(), rtol=1e-14)
    np.testing.assert_allclose(t2.numpy(), (t1 - 1e-4 * math.sqrt(1 - 0.81) * g / (v_ref.sqrt() + 1e-8)).numpy(), rtol=1e-14)


def _boost(p):
    """make both penalties active (slopes > 1) on the tiny net"""
    p["d_net/dense/kernel"] = p["d_net/dense/kernel"] * 6.0
    return p


def test_discriminator_is_per_sample_and_batchable():
    cfg = tiny_cfg()
    p = O.init_params(cfg, 0, torch.float64)
    f = O.make_feed(cfg, 5, torch.float64)
    big = O.discriminator(p, torch.cat([f["x"], f["x_mismatch"]]), torch.cat([f["cond"], f["cond"]]), cfg)
    a = O.discriminator(p, f["x"], f["cond"], cfg)
    b = O.discriminator(p, f["x_mismatch"], f["cond"], cfg)
    np.testing.assert_allclose(big.numpy(), torch.cat([a, b]).numpy(), rtol=1e-12, atol=1e-12)
    assert a.shape == (cfg.batch_size, 1, 1, 1)


def test_x_hat_limits_and_gp_finite_difference():
    cfg = tiny_cfg()
    p = _boost(O.init_params(cfg, 0, torch.float64))
    f = O.make_feed(cfg, 6, torch.float64)
    kt = torch.tensor(0.7, dtype=torch.float64)
    f0 = dict(f); f0["epsilon"] = torch.zeros_like(f["epsilon"])
    f1 = dict(f); f1["epsilon"] = torch.ones_like(f["epsilon"])
    o0 = O.d_forward_losses(p, kt, f0, cfg, False)
    o1 = O.d_forward_losses(p, kt, f1, cfg, False)
    np.testing.assert_allclose(o0["x_hat"].detach().numpy(), f["x"].numpy())
    np.testing.assert_allclose(o1["x_hat"].detach().numpy(), o1["G"].detach().numpy())
    out = O.d_forward_losses(p, kt, f, cfg, False)
    assert float(out["real_gp"]) > 0 and float(out["real_gp2"]) > 0
    # directional finite difference of D(x_hat) along a random image direction
    xh = out["x_hat"].detach()
    g = torch.Generator().manual_seed(7)
    d = torch.randn(xh.shape, generator=g, dtype=torch.float64)
    h = 1e-6
    fp = O.discriminator(p, xh + h * d, f["cond"], cfg).sum()
    fm = O.discriminator(p, xh - h * d, f["cond"], cfg).sum()
    fd = (fp - fm) / (2 * h)
    an = (out["grad_x_hat"] * d).sum()
    assert abs(fd - an) < 1e-6 * max(1.0, abs(an))


def test_second_order_weight_gradient_finite_difference():
    """d(150*(gp+gp2))/dW for one D weight vs central differences in fp64."""
    cfg = tiny_cfg()
    p = _boost(O.init_params(cfg, 0, torch.float64))
    st = O.new_state(p)
    f = O.make_feed(cfg, 8, torch.float64)
    res = O.d_step({k: v.clone() for k, v in p.items()}, st, f, cfg)
    name = "d_net/Conv_2/weights"
    idx = (1, 2, 3, 5)
    h = 1e-6

    def loss(delta):
        q = {k: v.clone() for k, v in p.items()}
        q[name][idx] += delta
        return float(O.d_forward_losses(q, torch.tensor(0.7, dtype=torch.float64), f, cfg, False)["D_loss"])

    fd = (loss(h) - loss(-h)) / (2 * h)
    an = float(res["grads"][name][idx])
    assert abs(fd - an) < 1e-5 * max(1.0, abs(an)), (fd, an)


def test_kl_and_kt_update():
    m = torch.tensor([[0.5, -1.0]], dtype=torch.float64)
    ls = torch.tensor([[0.1, -0.3]], dtype=torch.float64)
    ref = np.mean(-ls.numpy() + 0.5 * (-1 + np.exp(2 * ls.numpy()) + m.numpy() ** 2))
    assert abs(float(O.kl_std_normal_loss(m, ls)) - ref) < 1e-15
    cfg = tiny_cfg()
    p = O.init_params(cfg, 0, torch.float64)
    st = O.new_state(p)
    f = O.make_feed(cfg, 9, torch.float64)
    r = O.d_step(p, st, f, cfg)
    kt_grad = 2 * (0.7 * r["wdist2"] - r["wdist"]) * r["wdist2"]
    assert abs(float(r["kt_grad"]) - float(kt_grad)) < 1e-12
    assert abs(float(st["kt"]) - (0.7 - 1e-3 * float(kt_grad))) < 1e-15


def test_param_counts_match_reference_layer_table():
    sh = O.param_shapes(O.OracleCfg())
    g = sum(math.prod(s) for n, s in sh.items() if n.startswith("g_net/") and O.is_trainable(n))
    d = sum(math.prod(s) for n, s in sh.items() if n.startswith("d_net/"))
    assert g == 22_643_287 and d == 28_995_329


def test_bn_moving_stats_only_move_in_g_step():
    cfg = tiny_cfg()
    p = O.init_params(cfg, 0, torch.float64)
    st = O.new_state(p)
    f = O.make_feed(cfg, 10, torch.float64)
    O.d_step(p, st, f, cfg)
    assert float(p["g_net/BatchNorm_9/moving_mean"].abs().sum()) == 0.0
    O.g_step(p, st, f, cfg)
    assert float(p["g_net/BatchNorm_9/moving_mean"].abs().sum()) > 0.0


def test_golden_tiny_roundtrip():
    path = os.path.join(GOLD, "wgancls_tiny.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet")
    z = np.load(path)
    cfg = tiny_cfg()
    p = O.OrderedDict((k[2:], torch.from_numpy(z[k]).double()) for k in z.files if k.startswith("p/"))
    f = {k[2:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("f/")}
    st = O.new_state(p)
    rd, rg = O.iteration(p, st, f, cfg)
    np.testing.assert_allclose(rd["D_loss"].numpy(), z["o/D_loss"], rtol=1e-10)
    np.testing.assert_allclose(rg["G_loss"].numpy(), z["o/G_loss"], rtol=1e-10)
    np.testing.assert_allclose(rd["G"].numpy(), z["o/G"], rtol=1e-9, atol=1e-12)
    for k in z.files:
        if k.startswith("q/"):
            np.testing.assert_allclose(p[k[2:]].numpy(), z[k], rtol=1e-6, atol=1e-7)
