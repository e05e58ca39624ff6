"""Host logic on CPU: the engine's orchestration (text-to-image_b200/engine.py) driven by the CPU
restatement of the kernels (tests/fake_kernels.py) must reproduce the oracle's D run and G run:
losses, every parameter gradient (including the second-order gradient-penalty term), Adam and
kt updates, BN moving statistics.  np = 2 (split bf16, ~2^-17 per value) is compared tightly."""
import numpy as np
import pytest
import torch

import fake_kernels as fk
from oracle import wgancls_oracle as O

TINY = dict(batch_size=4, z_dim=8, embed_dim=32, compressed_embed_dim=8, gf_dim=8, df_dim=8)


def make_engine(cfg, np_=2, world=1, allreduce=None, batch=None):
    """np_ = 0 selects exact fp64 storage (no rounding anywhere): the orchestration alone is tested."""
    from t2i_b200.engine import Engine
    kw = dict(act_dtype=torch.float64, f32_dtype=torch.float64) if np_ == 0 else {}
    return Engine(fk, "cpu", batch or cfg.batch_size, max(np_, 1), cfg.z_dim, cfg.embed_dim,
                  cfg.compressed_embed_dim, cfg.gf_dim, cfg.df_dim, cfg.beta1, cfg.beta2, cfg.kl_coeff, world,
                  allreduce, **kw)


def boosted_params(cfg, seed=0):
    p = O.init_params(cfg, seed, torch.float64)
    p["d_net/dense/kernel"] = p["d_net/dense/kernel"] * 6.0        # makes the cond penalty active too
    g = torch.Generator().manual_seed(3)
    for n in p:                                                       # non-trivial biases / BN affine
        if n.endswith("biases") or n.endswith("bias") or n.endswith("beta"):
            p[n] = torch.randn(p[n].shape, generator=g, dtype=torch.float64) * 0.1
        if n.endswith("gamma"):
            p[n] = 1 + 0.2 * torch.randn(p[n].shape, generator=g, dtype=torch.float64)
    return p


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_param_layout_roundtrip():
    cfg = O.OracleCfg(**TINY)
    p = boosted_params(cfg)
    eng = make_engine(cfg)
    eng.set_params_tf(p)
    q = eng.get_params_tf()
    assert set(q) == set(p)
    for n in p:
        np.testing.assert_allclose(q[n].numpy(), p[n].float().numpy(), rtol=0, atol=0, err_msg=n)


@pytest.mark.parametrize("np_", [0, 2, 1])
def test_iteration_matches_oracle(np_):
    """np_ = 0: exact storage, agreement to ~1e-9 (any scheduling / layout / formula error shows).
    np_ = 2 / 1: the rounding paths; ReLU sign flips at |x| ~ rounding error make single elements of
    a gradient disagree, so the G-run bounds are loose there by design."""
    cfg = O.OracleCfg(**TINY)
    p = boosted_params(cfg)
    # feed seed 12: under np = 2 no d_net unit changes the sign of its pre-activation against the exact run
    # (seeds 11, 13, 14, 16-18 each flip 1-3 units that sit within ~2e-6 of zero, and ONE flipped LeakyReLU
    # derivative in this 8-channel net moves dD/dx_hat by 1e-2 -- the mask-flip effect of DESIGN.md section 5,
    # not a scheduling error: np = 0 is exact for every seed).  The tight np = 2 bounds below therefore measure
    # rounding alone; flips are covered by the loose np = 1 bounds.
    feed = O.make_feed(cfg, 12, torch.float64)
    eng = make_engine(cfg, np_)
    eng.set_params_tf(p)
    st = O.new_state(p)
    rd = O.d_step(p, st, feed, cfg)

    eng.load_feed(x=feed["x"], x_mismatch=feed["x_mismatch"], cond=feed["cond"],
                  z=feed["z"], epsilon=feed["epsilon"], tn_eps=feed["tn_eps"])
    eng.d_step(cfg.d_lr)
    tol = {0: 1e-9, 2: 2e-4, 1: 6e-2}[np_]
    B = cfg.batch_size
    assert rel(eng.d["img"][:B], rd["G"]) < tol
    gtol = 0.5 if np_ == 1 else tol    # bf16 storage on an 8-channel net: LeakyReLU mask flips dominate
    assert rel(eng.d["gx"], rd["grad_x_hat"]) < gtol
    assert rel(eng.d["g2"], rd["grad_cond"]) < gtol
    assert float(rd["real_gp"]) > 0 and float(rd["real_gp2"]) > 0
    sc = eng.scalars_dict()
    for k in ["D_loss", "D_loss_real", "D_loss_fake", "D_loss_mismatch", "wdist", "wdist2", "reg_loss",
              "balance_loss", "real_gp", "real_gp2"]:
        assert abs(sc[k] - float(rd[k])) < 4 * tol * max(1.0, abs(float(rd[k]))), (k, sc[k], float(rd[k]))
    assert abs(sc["kt"] - float(st["kt"])) < max(tol * 1e-2, 1e-12)
    grads = eng.get_grads_tf()
    worst = max((rel(grads[n], rd["grads"][n]), n) for n in rd["grads"] if float(rd["grads"][n].abs().max()) > 0)
    assert worst[0] < {0: 1e-8, 2: 1e-3, 1: 0.6}[np_], worst
    if np_ in (0, 2):
        newp = eng.get_params_tf()
        for n in rd["grads"]:
            # beta1 = 0 => the first Adam step is ~lr * sign(g): exact wherever g is clearly non-zero
            g = rd["grads"][n]
            sure = g.abs() > 1e-3 * g.abs().max()
            if bool(sure.any()):     # p was updated in place by the oracle
                assert float((newp[n].double() - p[n])[sure].abs().max()) < (1e-9 if np_ == 0 else 2e-6), n

    # ---- G run
    if np_ == 0:      # keep both sides on identical weights for the exact comparison of the G run
        eng.set_params_tf({k: v for k, v in p.items()})
    rg = O.g_step(p, st, feed, cfg)
    eng.load_feed(tn_eps=feed["tn_eps_g"])
    eng.g_step(cfg.g_lr)
    sc = eng.scalars_dict()
    assert abs(sc["G_loss"] - float(rg["G_loss"])) < tol * max(1.0, abs(float(rg["G_loss"])))
    assert abs(sc["G_kl_loss"] - float(rg["G_kl_loss"])) < tol * max(1.0, abs(float(rg["G_kl_loss"])))
    grads = eng.get_grads_tf()
    bad = []
    for n in rg["grads"]:
        ref = rg["grads"][n]
        scale = float(ref.abs().max())
        if n.endswith("biases") and "Conv2d_transpose_3" not in n and "Conv_9" not in n or n.endswith("dense_2/bias"):
            # a bias in front of a training-mode BatchNorm has an exactly-zero gradient (cancellation)
            continue
        if scale > 0:
            bad.append((rel(grads[n], ref), n))
    worst = max(bad)
    assert worst[0] < {0: 1e-7, 2: 5e-2, 1: 0.8}[np_], worst
    if np_ in (0, 2):
        newp = eng.get_params_tf()
        for n in ("g_net/BatchNorm_9/moving_mean", "g_net/BatchNorm/moving_variance", "g_net/BatchNorm_4/moving_mean"):
            assert rel(newp[n], p[n]) < (1e-9 if np_ == 0 else 1e-4), n


def test_losses_are_final_when_published():
    """d_step / g_step compute the loss scalars BEFORE the rest of the run (tangent pass + weight gradients; both backward
    passes) so that a D_loss / G_loss fetch can return early: the values at that point must be the run's final ones, and
    the order of the phases must be the documented one"""
    cfg = O.OracleCfg(**TINY)
    eng = make_engine(cfg, np_=0)
    eng.set_params_tf(boosted_params(cfg))
    feed = O.make_feed(cfg, 11, torch.float64)
    eng.load_feed(**{k: feed[k] for k in ("x", "x_mismatch", "cond", "z", "epsilon")}, tn_eps=feed["tn_eps"])
    order, snap = [], {}
    real_run = eng._run

    def run(name, body):
        real_run(name, body)
        order.append(name)
        if name in ("d_b", "g_b"):
            snap[name] = eng.scalars.clone()

    eng._run = run
    eng.d_step(cfg.d_lr)
    assert order == ["d_a1", "d_a", "d_b", "d_a2", "d_c"], order
    assert torch.equal(snap["d_b"], eng.scalars)            # nothing after d_b touches the scalars of the D run
    del order[:]
    eng.load_feed(tn_eps=feed["tn_eps_g"] if "tn_eps_g" in feed else feed["tn_eps"])
    eng.g_step(cfg.g_lr)
    assert order == ["g_a1", "g_a2", "g_b", "g_a3", "g_c"], order
    assert torch.equal(snap["g_b"], eng.scalars)
    sc = eng.scalars_dict()
    assert all(v == v for v in sc.values())
