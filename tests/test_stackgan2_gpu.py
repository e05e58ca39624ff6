"""-m gpu: StackGAN stage-II on the CUDA path, through the reference-facing API
(models/stackgan/stageII ConditionalGan / ConditionalGanTrainer) and the C ABI, against the CPU oracle on the same
seeded inputs.

Tolerances (relative L2): precision "bf16x3" (parity mode): generator image and discriminator logits <= 1e-3 / 5e-3 at
the reference widths; parameter gradients are bounded loosely (LeakyReLU / ReLU derivative flips at |x| ~ rounding
error, see tests/test_parity_gpu.py; the exact check of the schedule is tests/test_stackgan2_cpu.py).
precision "bf16": sanity (finite, close)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import stackgan2_oracle as S2

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(__file__))
TINY = dict(batch_size=4, z_dim=12, embed_dim=32, compressed_embed_dim=8, gf_dim=32, df_dim=8, s1_gf_dim=8)


def rel(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().cpu().reshape(-1)
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def cfgs(o, root="/tmp/t2i_s2"):
    from t2i_b200.utils.config import AttrDict
    model = {"Z_DIM": o.z_dim, "EMBED_DIM": o.embed_dim, "COMPRESSED_EMBED_DIM": o.compressed_embed_dim}
    train = {"BATCH_SIZE": o.batch_size, "SAMPLE_NUM": 4, "D_LR": o.lr, "G_LR": o.lr, "EPOCH": 1,
             "D_BETA_DECAY": o.d_beta1, "G_BETA_DECAY": o.g_beta1, "CHECKPOINTS_TO_KEEP": 2,
             "COEFF": {"ALPHA_MISMATCH_LOSS": o.alpha_mismatch, "KL": o.kl_coeff}}
    c1 = AttrDict({"CHECKPOINT_DIR": root + "/s1", "TRAIN": dict(train),
                   "MODEL": dict(model, OUTPUT_SIZE=64, GF_DIM=o.s1_gf_dim, DF_DIM=8,
                                 IMAGE_SHAPE={"W": 64, "H": 64, "D": 3})})
    c2 = AttrDict({"CHECKPOINT_DIR": root + "/s2", "TRAIN": dict(train),
                   "MODEL": dict(model, OUTPUT_SIZE=256, GF_DIM=o.gf_dim, DF_DIM=o.df_dim,
                                 IMAGE_SHAPE={"W": 256, "H": 256, "D": 3})})
    return c1, c2


def build(o, precision, params, root="/tmp/t2i_s2"):
    from t2i_b200.models.stackgan.stageI.model import ConditionalGan as StageI
    from t2i_b200.models.stackgan.stageII.model import ConditionalGan as StageII
    c1, c2 = cfgs(o, root)
    s1 = StageI(c1, precision=precision)
    s2 = StageII(s1, c2)
    s2.set_variables(params)
    return s1, s2, c1, c2


def trainer_for(m, c2, c1, data=None):
    from t2i_b200.models.stackgan.stageII.trainer import ConditionalGanTrainer
    tr = ConditionalGanTrainer(None, m, data, c2, c1)
    tr.define_losses()
    return tr


def feeds(m, tr, f, which, which1):
    return {m.inputs: f["x"], m.wrong_inputs: f["x_mismatch"], m.embed_inputs: f["cond"], m.z: f["z"],
            m.cond_noise: f[which], m.cond_noise_stagei: f[which1], tr.learning_rate: 2e-4}


def test_tiny_iteration_against_oracle():
    from test_stackgan2_cpu import boosted_params, _bias_before_bn
    ocfg = S2.Stage2Cfg(**TINY)
    p = boosted_params(ocfg)
    p = S2.OrderedDict((k, v.float().double()) for k, v in p.items())
    f = {k: v.float().double() for k, v in S2.make_feed(ocfg, 21, torch.float64).items()}
    s1, m, c1, c2 = build(ocfg, "bf16x3", p)
    tr = trainer_for(m, c2, c1)
    st = S2.new_state(p)
    rd = S2.d_step(p, st, f, ocfg)
    ff = {k: v.float() for k, v in f.items()}
    B = ocfg.batch_size
    _, d_loss, syn, real, mis = m.run([tr.D_optim, tr.D_loss, tr.D_synthetic_loss, tr.D_real_match_loss,
                                       tr.D_real_mismatch_loss], feeds(m, tr, ff, "tn_eps", "tn_s1"))
    eng = m._train_engine()
    e64, e256 = rel(eng.g["img64"], rd["stage1_G"]), rel(eng.d["img"][:B], rd["G"])
    print("\n[stage-II tiny] stage-I image rel-L2 %.3e  stage-II image rel-L2 %.3e" % (e64, e256))
    assert e64 < 1e-3 and e256 < 5e-3
    for k, n in enumerate(["D_synthetic_logits", "D_real_match_logits", "D_real_mismatch_logits"]):
        assert rel(eng.d["logit"][B * k:B * k + B], rd[n]) < 1e-2, n
    for got, n in ((d_loss, "D_loss"), (syn, "D_synthetic_loss"), (real, "D_real_match_loss"), (mis, "D_real_mismatch_loss")):
        assert abs(got - float(rd[n])) < 5e-3 * max(1.0, abs(float(rd[n]))), (n, got, float(rd[n]))
    grads = eng.get_grads_tf()
    errs = sorted(((rel(grads[n], rd["grads"][n]), n) for n in rd["grads"] if float(rd["grads"][n].abs().max()) > 1e-12
                   and not _bias_before_bn(n)), reverse=True)
    print("[stage-II tiny] D-run gradient rel-L2: worst %s  median %.3e" % (errs[0], errs[len(errs) // 2][0]))
    assert errs[0][0] < 0.15, errs[:4]
    rg = S2.g_step(p, st, f, ocfg)
    _, g_loss, gan, kl = m.run([tr.G_optim, tr.G_loss, tr.G_gan_loss, tr.G_kl_loss], feeds(m, tr, ff, "tn_eps_g", "tn_s1_g"))
    for got, n in ((g_loss, "G_loss"), (gan, "G_gan_loss"), (kl, "G_kl_loss")):
        assert abs(got - float(rg[n])) < 1e-2 * max(1.0, abs(float(rg[n]))), (n, got, float(rg[n]))
    grads = eng.get_grads_tf()
    errs = sorted(((rel(grads[n], rg["grads"][n]), n) for n in rg["grads"] if float(rg["grads"][n].abs().max()) > 1e-12
                   and not _bias_before_bn(n)), reverse=True)
    print("[stage-II tiny] G-run gradient rel-L2: worst %s  median %.3e" % (errs[0], errs[len(errs) // 2][0]))
    assert errs[0][0] < 0.3, errs[:4]


@pytest.mark.parametrize("precision,ftol", [("bf16x3", 1e-3), ("bf16", 5e-2)])
def test_reference_width_forward_parity(precision, ftol):
    """models/stackgan/stageII/cfg/flowers.yml widths (GF 128, DF 64, Z 100), batch 4: stage-I image -> stage-II G,
    and D on real and generated 256x256 images, against the oracle."""
    ocfg = S2.Stage2Cfg(batch_size=4)
    p = S2.init_params(ocfg, 0, torch.float32)
    g = torch.Generator().manual_seed(1)
    for n in p:      # N(0, 0.02) weights give near-constant outputs: use He-scaled weights for a meaningful comparison
        if n.endswith("weights") or n.endswith("kernel"):
            fan_in = p[n].shape[-2] * int(np.prod(p[n].shape[:-2])) if p[n].dim() > 1 else 1
            p[n] = torch.randn(p[n].shape, generator=g) * (2.0 / fan_in) ** 0.5
    f = S2.make_feed(ocfg, 7, torch.float32)
    s1, m, _, _ = build(ocfg, precision, p)
    with torch.no_grad():
        img64, _, _ = S2.S1.generator(p, f["z"], f["cond"], f["tn_s1"], ocfg.stage1())
        G, mean, ls = S2.generator(p, img64, f["cond"], f["tn_eps"], ocfg)
        Dx = S2.discriminator(p, f["x"], f["cond"], ocfg)
        Dg = S2.discriminator(p, G, f["cond"], ocfg)
    img, mean_g, ls_g = m.generator(img64, f["cond"], noise=f["tn_eps"])
    e_g = rel(img, G)
    e_dx = rel(m.discriminator(f["x"], f["cond"])[1], Dx)
    e_dg = rel(m.discriminator(G, f["cond"])[1], Dg)
    print("\n[stage-II parity] %s: G rel-L2 %.3e  D(x) rel-L2 %.3e  D(G) rel-L2 %.3e" % (precision, e_g, e_dx, e_dg))
    assert rel(mean_g, mean) < ftol and rel(ls_g, ls) < ftol
    assert e_g < ftol and e_dx < 5 * ftol and e_dg < 5 * ftol


def test_trainer_loop_sampler_and_graphs(tmp_path):
    """three updates through the trainer mirror (CUDA graphs replay from the second update on), then the sampler."""
    from t2i_b200.models.wgancls.trainer import SyntheticTextDataset
    ocfg = S2.Stage2Cfg(**TINY)
    s1, m, c1, c2 = build(ocfg, "bf16", S2.init_params(ocfg, 0, torch.float32), root=str(tmp_path))
    from t2i_b200.models.stackgan.stageII.trainer import ConditionalGanTrainer
    data = SyntheticTextDataset(embed_dim=ocfg.embed_dim, num_examples=64, image_size=256)
    tr = ConditionalGanTrainer(None, m, data, c2, c1)
    tr.train(max_updates=3)
    assert len(tr.log) == 3 and all(np.isfinite(r["d_loss"]) and np.isfinite(r["g_loss"]) for r in tr.log)
    samples = m.run(m.sampler, feed_dict={m.z_sample: np.random.normal(0, 1, (4, ocfg.z_dim)),
                                         m.embed_sample: np.random.normal(0, 1, (4, ocfg.embed_dim))})
    assert samples.shape == (4, 256, 256, 3) and float(np.abs(samples).max()) <= 1.0
    assert sorted(os.listdir(c2.CHECKPOINT_DIR)) == ["checkpoint", "stageII-2.npz"]
