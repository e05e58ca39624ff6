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


# ---------------------------------------------------------------------------------------------------------------------
# the reference-facing mirrors (models/stackgan/stageII/{model,trainer}.py) on the CPU restatement of the kernels
def _cfgs(tmp_path, o):
    from t2i_b200.utils.config import AttrDict
    model = {"Z_DIM": o.z_dim, "EMBED_DIM": o.embed_dim, "COMPRESSED_EMBED_DIM": o.compressed_embed_dim}
    train = {"BATCH_SIZE": o.batch_size, "SAMPLE_NUM": 2, "D_LR": o.lr, "G_LR": o.lr, "EPOCH": 1,
             "D_BETA_DECAY": o.d_beta1, "G_BETA_DECAY": o.g_beta1, "CHECKPOINTS_TO_KEEP": 2,
             "COEFF": {"ALPHA_MISMATCH_LOSS": o.alpha_mismatch, "KL": o.kl_coeff}}
    c1 = AttrDict({"CHECKPOINT_DIR": str(tmp_path / "s1"), "TRAIN": dict(train),
                   "MODEL": dict(model, OUTPUT_SIZE=64, GF_DIM=o.s1_gf_dim, DF_DIM=8,
                                 IMAGE_SHAPE={"W": 64, "H": 64, "D": 3})})
    c2 = AttrDict({"CHECKPOINT_DIR": str(tmp_path / "s2"), "TRAIN": dict(train),
                   "MODEL": dict(model, OUTPUT_SIZE=256, GF_DIM=o.gf_dim, DF_DIM=o.df_dim,
                                 IMAGE_SHAPE={"W": 256, "H": 256, "D": 3})})
    return c1, c2


def _models(tmp_path, o):
    from t2i_b200.models.stackgan.stageI.model import ConditionalGan as StageI
    from t2i_b200.models.stackgan.stageII.model import ConditionalGan as StageII
    c1, c2 = _cfgs(tmp_path, o)
    s1 = StageI(c1, precision="bf16x3", device="cpu", kernels=fk, use_graphs=False)
    s2 = StageII(s1, c2)
    return s1, s2, c1, c2


def test_model_mirror_shares_stage1_generator_and_matches_oracle(tmp_path):
    cfg = S2.Stage2Cfg(**TINY)
    p = boosted_params(cfg)
    s1, s2, c1, c2 = _models(tmp_path, cfg)
    s2.set_variables(p)
    # one copy of g_net/*: what stage-II received is what stage-I now holds
    v1 = s1.get_variables()
    for n in (k for k in p if k.startswith("g_net/")):
        assert torch.equal(torch.as_tensor(v1[n]), p[n].float()), n
    assert all(n.startswith(S2.G2) for n in s2.g_vars) and all(n.startswith(S2.D2) for n in s2.d_vars)
    assert set(s2.g_vars) == set(S2.g_var_names(p)) and set(s2.d_vars) == set(S2.d_var_names(p))
    feed = S2.make_feed(cfg, 3, torch.float32)
    pf = {k: v.float() for k, v in p.items()}
    with torch.no_grad():
        img64, _, _ = S2.S1.generator(pf, feed["z"], feed["cond"], feed["tn_s1"], cfg.stage1())
        G, mean, ls = S2.generator(pf, img64, feed["cond"], feed["tn_eps"], cfg)
        Dx = S2.discriminator(pf, feed["x"], feed["cond"], cfg)
    got64, _, _ = s1.generator(feed["z"], feed["cond"], noise=feed["tn_s1"])
    img, mean_g, ls_g = s2.generator(got64, feed["cond"], noise=feed["tn_eps"])
    assert rel(got64, img64) < 1e-3 and rel(img, G) < 5e-3 and rel(mean_g, mean) < 1e-3 and rel(ls_g, ls) < 1e-3
    prob, logits = s2.discriminator(feed["x"], feed["cond"])
    assert logits.shape == (cfg.batch_size, 1, 1, 1) and rel(logits, Dx) < 1e-2
    assert torch.allclose(prob, torch.sigmoid(logits))
    with pytest.raises(ValueError):
        bad = _cfgs(tmp_path, cfg)[1]
        bad.MODEL.OUTPUT_SIZE = 128
        type(s2)(s1, bad)


def test_trainer_two_checkpoint_restore(tmp_path):
    from t2i_b200.models.stackgan.stageII.trainer import ConditionalGanTrainer
    from t2i_b200.models.wgancls.trainer import SyntheticTextDataset
    from t2i_b200.utils import saver
    cfg = S2.Stage2Cfg(**TINY)
    s1, s2, c1, c2 = _models(tmp_path, cfg)
    # a stage-I checkpoint with recognisable generator weights (written by the stage-I model itself)
    s1.initialize(seed=11)
    saver.save(s1, c1.CHECKPOINT_DIR, 7, prefix="stageI")
    g1 = {k: torch.as_tensor(v).clone() for k, v in s1.get_variables().items() if k.startswith("g_net/")}
    s1.initialize(seed=12)       # clobber; the trainer must restore seed 11's generator from the stage-I directory
    data = SyntheticTextDataset(embed_dim=cfg.embed_dim, num_examples=16, image_size=256)
    tr = ConditionalGanTrainer(None, s2, data, c2, c1)
    tr.train(max_updates=1)
    assert len(tr.log) == 1 and np.isfinite(tr.log[0]["d_loss"]) and np.isfinite(tr.log[0]["g_loss"])
    after = s2.get_variables()
    for n, w in g1.items():
        if "moving" not in n:      # frozen: the two runs did not touch the stage-I generator's trainables
            assert torch.equal(torch.as_tensor(after[n]), w), n
    # counter 2 -> a stage-II checkpoint that holds only the stage-II scopes (trainer.py:48-52,175-176)
    import os
    assert sorted(os.listdir(c2.CHECKPOINT_DIR)) == ["checkpoint", "stageII-2.npz"]
    z = np.load(os.path.join(c2.CHECKPOINT_DIR, "stageII-2.npz"))
    assert all(k.startswith((S2.G2, S2.D2, "__adam__/")) for k in z.files)
    s2.initialize(seed=99)
    tr2 = ConditionalGanTrainer(None, s2, data, c2, c1)
    tr2.define_losses()
    assert saver.load(tr2.stageii_saver, c2.CHECKPOINT_DIR, prefix="stageII") == (True, 2)
    again = s2.get_variables()
    for n in again:
        if n.startswith((S2.G2, S2.D2)):
            assert torch.equal(torch.as_tensor(again[n]), torch.as_tensor(after[n])), n
    samples = s2.run(s2.sampler, feed_dict={s2.z_sample: np.random.normal(0, 1, (2, cfg.z_dim)),
                                           s2.embed_sample: np.random.normal(0, 1, (2, cfg.embed_dim))})
    assert samples.shape == (2, 256, 256, 3) and float(np.abs(samples).max()) <= 1.0


def test_oracle_k4s1_same_padding_against_naive_loops():
    """TF SAME for a 4x4 stride-1 conv: pad_total = 3 -> 1 before, 2 after (models/stackgan/stageII/model.py:102,105,
    151,154 via utils/ops.py:58-63).  The oracle's conv2d against explicit loops, and the self-added residual branch
    (model.py:117) against its literal reading."""
    from oracle import wgancls_oracle as W
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 5, 6, generator=g, dtype=torch.float64)          # NCHW
    p = {"c/weights": torch.randn(4, 4, 3, 2, generator=g, dtype=torch.float64),
         "c/biases": torch.randn(2, generator=g, dtype=torch.float64)}
    y = W.conv2d(p, "c", x, 4, 1)
    assert y.shape == (2, 2, 5, 6)
    xp = torch.zeros(2, 3, 5 + 3, 6 + 3, dtype=torch.float64)
    xp[:, :, 1:6, 1:7] = x
    for n in range(2):
        for i in range(5):
            for j in range(6):
                ref = p["c/biases"].clone()
                for kh in range(4):
                    for kw in range(4):
                        ref = ref + xp[n, :, i + kh, j + kw] @ p["c/weights"][kh, kw]
                assert torch.allclose(y[n, :, i, j], ref, atol=1e-12)
    # lrelu(net + net): doubling the BatchNorm output equals doubling gamma and beta
    cfg = S2.Stage2Cfg(**TINY)
    q = boosted_params(cfg)
    f = S2.make_feed(cfg, 3, torch.float64)
    a = S2.discriminator(q, f["x"], f["cond"], cfg)
    q2 = dict(q)
    q2[S2.D2 + "BatchNorm_9/gamma"] = 2 * q[S2.D2 + "BatchNorm_9/gamma"]
    q2[S2.D2 + "BatchNorm_9/beta"] = 2 * q[S2.D2 + "BatchNorm_9/beta"]

    def single(pp, x_nhwc, embed):     # the same graph with h8 = lrelu(n) instead of lrelu(n + n)
        import torch.nn.functional as F
        from oracle.wgancls_oracle import batch_norm, conv2d, fc, lrelu
        d, bnk = S2.D2, dict(train=True, new_moving=None)
        h = conv2d(pp, d + "Conv", x_nhwc.permute(0, 3, 1, 2), 4, 2, act=lrelu)
        for i in range(5):
            h = batch_norm(pp, d + ("BatchNorm" if i == 0 else "BatchNorm_%d" % i), conv2d(pp, d + "Conv_%d" % (1 + i), h, 4, 2),
                           act=lrelu, **bnk)
        h = batch_norm(pp, d + "BatchNorm_5", conv2d(pp, d + "Conv_6", h, 4, 1), act=lrelu, **bnk)
        h7 = batch_norm(pp, d + "BatchNorm_6", conv2d(pp, d + "Conv_7", h, 4, 1), **bnk)
        n_ = batch_norm(pp, d + "BatchNorm_7", conv2d(pp, d + "Conv_8", h7, 1, 1), act=lrelu, **bnk)
        n_ = batch_norm(pp, d + "BatchNorm_8", conv2d(pp, d + "Conv_9", n_, 3, 1), act=lrelu, **bnk)
        n_ = batch_norm(pp, d + "BatchNorm_9", conv2d(pp, d + "Conv_10", n_, 3, 1), **bnk)
        h8 = lrelu(n_)
        e = fc(pp, d + "dense", embed, lrelu)
        s = h8.shape[-1]
        h9 = batch_norm(pp, d + "BatchNorm_10", conv2d(pp, d + "Conv_11", torch.cat([h8, e[:, :, None, None].expand(-1, -1, s, s)], 1),
                                                       1, 1), act=lrelu, **bnk)
        w = pp[d + "Conv_12/weights"].permute(3, 2, 0, 1)
        return F.conv2d(h9, w, pp[d + "Conv_12/biases"], stride=cfg.output_size // 64).permute(0, 2, 3, 1)

    b = single(q2, f["x"], f["cond"])
    assert rel(a, b) < 1e-12


def test_run_entry_point_trains_stage2_around_a_stage1_generator(tmp_path):
    """models/stackgan/stageII/run.py:26-86: two configs, directories, TRAIN dispatch (stage-I built with
    build_model=False, only its generator is used), one update on the CPU restatement of the kernels."""
    import os
    import yaml
    from t2i_b200.models.stackgan.stageII import run
    cfg = S2.Stage2Cfg(**TINY)
    c1, c2 = _cfgs(tmp_path, cfg)
    for c, tag in ((c1, "s1"), (c2, "s2")):
        c["DATASET_DIR"] = str(tmp_path / "data")
        c["LOGS_DIR"], c["SAMPLE_DIR"] = str(tmp_path / tag / "logs"), str(tmp_path / tag / "samples")
        c["TRAIN"]["FLAG"], c["EVAL"] = True, {"FLAG": False}
    plain = lambda d: {k: (plain(v) if isinstance(v, dict) else v) for k, v in d.items()}
    p1, p2 = tmp_path / "c1.yml", tmp_path / "c2.yml"
    p1.write_text(yaml.safe_dump(plain(c1)))
    p2.write_text(yaml.safe_dump(plain(c2)))
    from t2i_b200.models.wgancls.trainer import SyntheticTextDataset
    data = SyntheticTextDataset(embed_dim=cfg.embed_dim, num_examples=16, image_size=256)
    tr = run.main(str(p1), str(p2), dataset=data, max_updates=1, precision="bf16x3", device="cpu", kernels=fk, use_graphs=False)
    assert len(tr.log) == 1 and np.isfinite(tr.log[0]["d_loss"]) and os.path.isdir(c2["SAMPLE_DIR"])
    c2["TRAIN"]["FLAG"] = False
    p2.write_text(yaml.safe_dump(plain(c2)))
    with pytest.raises(NotImplementedError, match="visualis"):
        run.main(str(p1), str(p2), dataset=data, precision="bf16x3", device="cpu", kernels=fk, use_graphs=False)
