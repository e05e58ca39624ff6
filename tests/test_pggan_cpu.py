"""Conditional PGGAN host logic on CPU: PgganEngine (text-to-image_b200/engine_pggan.py) driven
by the CPU restatement of the kernels must reproduce the oracle's D run and G run for every kind of stage graph
(first stage, a transition stage with the fade-in blend on both networks, a stabilised later stage): layer_norm
forward / backward, nearest upscale / average pool and their transposes, the 2x2 SAME to_rgb conv with padded channels,
the 4B-batched discriminator with the tangent pass for the two gradient penalties, Adam(2e-6, 0, 0.99).
np_ = 0: exact fp64 storage (any schedule / layout / formula error shows at 1e-9); np_ = 2: split-bf16 rounding.
Also pins the oracle's own ops against naive loops and the penalty against finite differences."""
import numpy as np
import pytest
import torch

import fake_kernels as fk
from oracle import pggan_oracle as P

TINY = dict(batch_size=3, z_dim=16, embed_dim=32, compr_embed_dim=8, nf_base=16, nf_cap=16, d_embed=8)


def make_engine(cfg, np_=2, **kw):
    from t2i_b200.engine_pggan import PgganEngine
    if np_ == 0:
        kw.update(act_dtype=torch.float64, f32_dtype=torch.float64)
    return PgganEngine(fk, "cpu", cfg.batch_size, max(np_, 1), cfg.stage, cfg.trans, cfg.z_dim, cfg.embed_dim,
                       cfg.compr_embed_dim, cfg.nf_base, cfg.nf_cap, cfg.d_embed, cfg.rgb_mid, **kw)


def boosted_params(cfg, seed=0):
    """zero biases / beta and unit gamma hide terms: perturb them"""
    p = P.init_params(cfg, seed, torch.float64)
    g = torch.Generator().manual_seed(5)
    for n in p:
        if n.endswith("biases") or n.endswith("bias") or n.endswith("beta"):
            p[n] = torch.randn(p[n].shape, generator=g, dtype=torch.float64) * 0.1
        if n.endswith("gamma"):
            p[n] = 1 + 0.2 * torch.randn(p[n].shape, generator=g, dtype=torch.float64)
        if n.startswith("d_net/") and (n.endswith("weights") or n.endswith("kernel")):
            p[n] = p[n] * 1.6       # slopes > 1: both one-sided penalties (and their second-order terms) are active
    return p


def rel(a, b):
    a, b = torch.as_tensor(a).double().reshape(-1), torch.as_tensor(b).double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


# --------------------------------------------------------------------------------------------- the oracle's own ops
def test_oracle_ops_against_naive_loops():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 6, 3, generator=g, dtype=torch.float64)
    up, po = P.upscale(x), P.pool(x)
    for n in range(2):
        for i in range(8):
            for j in range(12):
                assert torch.equal(up[n, i, j], x[n, i // 2, j // 2])
        for i in range(2):
            for j in range(3):
                assert torch.allclose(po[n, i, j], x[n, 2 * i:2 * i + 2, 2 * j:2 * j + 2].reshape(4, 3).mean(0))
    # 2x2 SAME conv: TF pads 0 before, 1 after
    p = {"c/weights": torch.randn(2, 2, 3, 2, generator=g, dtype=torch.float64), "c/biases": torch.randn(2, generator=g, dtype=torch.float64)}
    y = P.conv2d(p, "c", x, 2)
    xp = torch.zeros(2, 5, 7, 3, dtype=torch.float64)
    xp[:, :4, :6] = x
    for i in range(4):
        for j in range(6):
            ref = p["c/biases"].clone()
            for kh in range(2):
                for kw in range(2):
                    ref = ref + xp[0, i + kh, j + kw] @ p["c/weights"][kh, kw]
            assert torch.allclose(y[0, i, j], ref)
    # layer_norm: per-sample statistics over (H, W, C), per-channel affine
    q = {"l/gamma": torch.rand(3, generator=g, dtype=torch.float64) + 0.5, "l/beta": torch.randn(3, generator=g, dtype=torch.float64)}
    ln = P.layer_norm(q, "l", x)
    for n in range(2):
        v = x[n].reshape(-1)
        m, var = v.mean(), ((v - v.mean()) ** 2).mean()
        assert torch.allclose(ln[n], (x[n] - m) / torch.sqrt(var + P.LN_EPS) * q["l/gamma"] + q["l/beta"])


def test_oracle_gradient_penalty_by_finite_differences():
    cfg = P.PgganCfg(stage=2, trans=True, **TINY)
    p = boosted_params(cfg)
    f = P.make_feed(cfg, 2, torch.float64)
    out = P.d_forward_losses(p, f, cfg, 0.4, create_graph=False)
    x_hat, cond = out["x_hat"].detach(), f["cond"]
    g = torch.Generator().manual_seed(9)
    for _ in range(3):
        dx = torch.randn(x_hat.shape, generator=g, dtype=torch.float64)
        h = 1e-6
        num = (P.discriminator(p, x_hat + h * dx, cond, cfg, 0.4).sum() - P.discriminator(p, x_hat - h * dx, cond, cfg, 0.4).sum()) / (2 * h)
        assert abs(float(num) - float((out["grad_x_hat"] * dx).sum())) < 1e-6 * max(1.0, abs(float(num)))


# --------------------------------------------------------------------------------------------- the engine
GRAPHS = [(1, False), (2, True), (3, False), (3, True)]


@pytest.mark.parametrize("stage,trans", GRAPHS)
def test_param_layout_roundtrip(stage, trans):
    cfg = P.PgganCfg(stage=stage, trans=trans, **TINY)
    p = boosted_params(cfg)
    eng = make_engine(cfg)
    eng.set_params_tf(p)
    q = eng.get_params_tf()
    assert set(q) == set(p)
    for n in p:
        np.testing.assert_allclose(q[n].numpy(), p[n].float().numpy(), rtol=0, atol=0, err_msg=n)


@pytest.mark.parametrize("np_", [0, 2])
@pytest.mark.parametrize("stage,trans", GRAPHS)
def test_iteration_matches_oracle(stage, trans, np_):
    cfg = P.PgganCfg(stage=stage, trans=trans, **TINY)
    alpha = 0.3
    p = boosted_params(cfg)
    feed = P.make_feed(cfg, 21, torch.float64)
    eng = make_engine(cfg, np_)
    eng.set_params_tf(p)
    st = P.new_state(p)
    rd = P.d_step(p, st, feed, cfg, alpha)
    eng.load_feed(x=feed["x"], x_mismatch=feed["x_mismatch"], cond=feed["cond"], z=feed["z"], epsilon=feed["epsilon"],
                  tn_eps=feed["tn_eps"])
    eng.d_step(alpha)
    tol = 1e-9 if np_ == 0 else 2e-3
    B = cfg.batch_size
    assert rel(eng.d["img"][:B], rd["G"]) < tol
    assert rel(eng.d["img"][3 * B:], rd["x_hat"]) < tol
    lg = eng.d["logit"]
    for k, n in enumerate(["Dg_logit", "Dx_logit", "Dxmi_logit", "Dx_hat_logit"]):
        assert rel(lg[k * B:(k + 1) * B], rd[n]) < (tol if np_ == 0 else 2e-2), n
    assert rel(eng.d["slope"], rd["slopes"]) < (tol if np_ == 0 else 5e-2)
    assert rel(eng.d["slope2"], rd["slopes2"]) < (tol if np_ == 0 else 5e-2)
    assert float(rd["real_gp"]) > 1e-3 and float(rd["real_gp2"]) > 1e-3, "the penalties must be active in this test"
    sc = eng.scalars_dict()
    for k in ["D_loss", "D_loss_real", "D_loss_fake", "D_loss_mismatch", "wdist", "wdist2", "reg_loss", "real_gp", "real_gp2"]:
        bound = (10 * tol if np_ == 0 else 0.1) * max(1.0, abs(float(rd[k])))
        assert abs(sc[k] - float(rd[k])) < bound, (k, sc[k], float(rd[k]))
    grads = eng.get_grads_tf()
    errs = sorted(((rel(grads[n], rd["grads"][n]), n) for n in rd["grads"] if float(rd["grads"][n].abs().max()) > 1e-14),
                  reverse=True)
    assert errs[0][0] < (1e-8 if np_ == 0 else 0.3), errs[:4]
    if np_ == 0:
        newp = eng.get_params_tf()
        for n in rd["grads"]:
            g = rd["grads"][n]
            sure = g.abs() > 1e-3 * g.abs().max()
            if bool(sure.any()):
                assert float((newp[n].double() - p[n])[sure].abs().max()) < 1e-9, n
        eng.set_params_tf({k: v for k, v in p.items()})
    # ---- G run
    rg = P.g_step(p, st, feed, cfg, alpha)
    eng.load_feed(tn_eps=feed["tn_eps_g"])
    eng.g_step()
    sc = eng.scalars_dict()
    for k in ["G_loss", "G_kl_loss"]:
        bound = (10 * tol if np_ == 0 else 5e-2) * max(1.0, abs(float(rg[k])))
        assert abs(sc[k] - float(rg[k])) < bound, (k, sc[k], float(rg[k]))
    grads = eng.get_grads_tf()
    errs = sorted(((rel(grads[n], rg["grads"][n]), n) for n in rg["grads"] if float(rg["grads"][n].abs().max()) > 1e-14),
                  reverse=True)
    assert errs[0][0] < (1e-7 if np_ == 0 else 0.3), errs[:4]
    if np_ == 0:
        newp = eng.get_params_tf()
        for n in rg["grads"]:
            g = rg["grads"][n]
            sure = g.abs() > 1e-3 * g.abs().max()
            if bool(sure.any()):
                assert float((newp[n].double() - p[n])[sure].abs().max()) < 1e-9, n


# --------------------------------------------------------------------------------------------- the mirrors
def _model(tmp_path, stage, trans, steps=40, **kw):
    from t2i_b200.models.pggan.pggan import PGGAN
    from t2i_b200.models.wgancls.trainer import SyntheticTextDataset
    data = SyntheticTextDataset(embed_dim=32, num_examples=32, image_size=4 * 2 ** (stage - 1))
    prev = stage - 1 if trans else stage
    return PGGAN(3, steps, str(tmp_path / ("stage%d" % stage)), str(tmp_path / ("stage%d" % prev)), data, str(tmp_path / "s"),
                 str(tmp_path / "l"), stage, trans, precision="bf16x3", device="cpu", kernels=fk, use_graphs=False,
                 nf_base=16, nf_cap=16, z_dim=16, embed_dim=32, compr_embed_dim=8, sample_num=2, **kw)


def test_schedule_matches_reference():
    from t2i_b200.models.pggan.train_pggan import schedule
    s = schedule()
    assert len(s) == 15 and s[0] == (1, 1, False, 16, 37500) and s[1] == (2, 1, True, 16, 37500)
    assert s[2] == (2, 2, False, 16, 37500) and s[9] == (6, 5, True, 8, 75000) and s[14] == (8, 8, False, 8, 75000)
    assert [t for (_, _, t, _, _) in s] == [i % 2 == 1 for i in range(15)]


def test_mirror_stage_to_stage_restore_and_fade_in(tmp_path):
    import os
    m1 = _model(tmp_path, 1, False)
    assert m1.output_size == 4 and m1.get_nf(0) == 16 and m1.get_dnf(2) == 8
    assert m1.get_variables_up_to_stage(1) == ["d_net/rgb_stage_0/", "g_net/rgb_stage_0/", "d_net/conv_stage_0/",
                                               "g_net/conv_stage_0/"]
    m1.train(max_updates=2)
    assert sorted(os.listdir(m1.check_dir_write)) == ["checkpoint", "pggan-2.npz"]
    v1 = m1.get_variables()
    # the transition graph of stage 2 restores everything of stage 1 and initialises its new layers
    m2 = _model(tmp_path, 2, True)
    assert m2.restore is not None and set(m2.d_vars + m2.g_vars) == set(m2.variable_names())
    m2.train(max_updates=1)
    assert abs(m2.alpha_tra - 1.0 / 40) < 1e-12                     # alpha_assign: iter / steps (pggan.py:78-79)
    v2 = m2.get_variables()
    moved = [n for n in v1 if not torch.equal(torch.as_tensor(v1[n]), torch.as_tensor(v2[n]))]
    assert moved, "one update must move the restored variables"
    for n in v1:       # restored, then one Adam step of at most ~lr * 10 per weight
        assert float((torch.as_tensor(v2[n]).double() - torch.as_tensor(v1[n]).double()).abs().max()) < 1e-4, n
    assert any(n.startswith("g_net/conv_stage_1/") for n in v2) and any(n.startswith("d_net/rgb_stage_1/") for n in v2)
    # eager sub-graphs at this object's stage only
    img, mean, ls = m2.generator(np.random.normal(0, 1, (3, 16)), np.random.normal(0, 1, (3, 32)))
    assert img.shape == (3, 8, 8, 3) and mean.shape == (3, 8)
    assert m2.discriminator(img, np.random.normal(0, 1, (3, 32))).shape == (3, 1, 1, 1)
    with pytest.raises(NotImplementedError):
        m2.generator(np.zeros((3, 16)), np.zeros((3, 32)), stages=1, t=False)
    with pytest.raises(RuntimeError):
        _model(tmp_path / "elsewhere", 2, True).train(max_updates=1)      # no stage-1 checkpoint to fade in from (:149-151)


def test_sampler_runs_in_training_batch_chunks(tmp_path):
    """sess.run(sampler) for sample_num images uses the training engine chunk by chunk (the generator has no batch
    statistics): same images as the generator called on each chunk, and no engine for a second batch size appears"""
    m = _model(tmp_path, 2, False)            # batch 3
    m.initialize()
    rng = np.random.RandomState(3)
    z, c, noise = rng.normal(0, 1, (5, 16)), rng.normal(0, 1, (5, 32)), rng.normal(0, 1, (5, 8)).clip(-2, 2)
    m.sample_num = 5
    got = m.run(m.sampler, feed_dict={m.z_sample: z, m.cond_sample: c, m.cond_noise_sample: noise})
    assert got.shape == (5, 8, 8, 3)
    for lo in (0, 3):
        idx = np.arange(lo, lo + 3) % 5
        ref, _, _ = m.generator(z[idx], c[idx], noise=noise[idx])
        n = min(3, 5 - lo)
        assert np.array_equal(got[lo:lo + n], ref[:n].cpu().numpy())
    assert list(m._engines) == [3], list(m._engines)


def test_schedule_driver_runs_consecutive_passes(tmp_path):
    """train_pggan.py:20-69: pass 0 (stage 1) writes stage1/, pass 1 (stage 2, transition) reads stage1/ and writes stage2/"""
    import os
    from t2i_b200.models.pggan import train_pggan
    from t2i_b200.models.wgancls.trainer import SyntheticTextDataset
    from t2i_b200.utils.config import AttrDict
    cfg = AttrDict({"CHECKPOINT_DIR": str(tmp_path / "ck"), "SAMPLE_DIR": str(tmp_path / "s"), "LOGS_DIR": str(tmp_path / "l"),
                    "MODEL": {"SIZES": [4, 8, 16, 32, 64, 128, 256, 512], "EMBED_DIM": 32}})
    seen = []

    def data(size):
        seen.append(size)
        return SyntheticTextDataset(embed_dim=32, num_examples=64, image_size=size)

    ms = train_pggan.train(cfg, data, images=160, passes=[0, 1], max_updates=2, precision="bf16x3", device="cpu", kernels=fk,
                           use_graphs=False, nf_base=16, nf_cap=16, z_dim=16, embed_dim=32, compr_embed_dim=8, sample_num=2,
                           d_embed=8)
    assert seen == [4, 8] and [(m.stage, m.trans, m.batch_size, m.steps) for m in ms] == [(1, False, 16, 10), (2, True, 16, 10)]
    assert sorted(os.listdir(os.path.join(cfg.CHECKPOINT_DIR, "stage1"))) == ["checkpoint", "pggan-2.npz"]
    assert sorted(os.listdir(os.path.join(cfg.CHECKPOINT_DIR, "stage2"))) == ["checkpoint", "pggan-2.npz"]
    assert abs(ms[1].alpha_tra - 0.2) < 1e-12
