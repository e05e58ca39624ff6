"""-m gpu: conditional PGGAN on the CUDA path, through the reference-facing API
(models/pggan/pggan.py PGGAN) and the C ABI, against the CPU oracle on the same seeded inputs.

Tolerances (relative L2): precision "bf16x3" (parity mode): generator image and discriminator logits <= 1e-3 / 5e-3 at
the reference widths; parameter gradients are bounded loosely (LeakyReLU / ReLU derivative flips at |x| ~ rounding
error; the exact check of the schedule is tests/test_pggan_cpu.py).  precision "bf16": sanity."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import pggan_oracle as P

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(__file__))
TINY = dict(batch_size=4, z_dim=16, embed_dim=32, compr_embed_dim=8, nf_base=16, nf_cap=16)


def rel(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().cpu().reshape(-1)
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def build(o, precision, root, params=None, steps=100, use_graphs=True):
    from t2i_b200.models.pggan.pggan import PGGAN
    from t2i_b200.models.wgancls.trainer import SyntheticTextDataset
    data = SyntheticTextDataset(embed_dim=o.embed_dim, num_examples=64, image_size=o.out_size)
    prev = o.stage - 1 if o.trans else o.stage
    m = PGGAN(o.batch_size, steps, os.path.join(root, "stage%d" % o.stage), os.path.join(root, "stage%d" % prev), data,
              os.path.join(root, "s"), os.path.join(root, "l"), o.stage, o.trans, precision=precision, nf_base=o.nf_base,
              nf_cap=o.nf_cap, z_dim=o.z_dim, embed_dim=o.embed_dim, compr_embed_dim=o.compr_embed_dim, sample_num=4, d_embed=o.d_embed,
              use_graphs=use_graphs)
    if params is not None:
        m.set_variables(params)
    return m


@pytest.mark.parametrize("stage,trans", [(1, False), (2, True), (3, False), (3, True)])
def test_tiny_iteration_against_oracle(stage, trans, tmp_path):
    from test_pggan_cpu import boosted_params
    ocfg = P.PgganCfg(stage=stage, trans=trans, **TINY)       # d_embed stays 128 as in the reference
    alpha, steps = 0.3, 100
    p = boosted_params(ocfg)
    p = P.OrderedDict((k, v.float().double()) for k, v in p.items())
    f = {k: v.float().double() for k, v in P.make_feed(ocfg, 21, torch.float64).items()}
    m = build(ocfg, "bf16x3", str(tmp_path), p, steps)
    st = P.new_state(p)
    rd = P.d_step(p, st, f, ocfg, alpha)
    ff = {k: v.float() for k, v in f.items()}
    B = ocfg.batch_size
    fd = {m.x: ff["x"], m.x_mismatch: ff["x_mismatch"], m.cond: ff["cond"], m.z: ff["z"], m.epsilon: ff["epsilon"],
          m.cond_noise: ff["tn_eps"], m.iter: alpha * steps}
    _, d_loss, gp, gp2, wd = m.run([m.D_optim, m.D_loss, m.real_gp, m.real_gp2, m.wdist], fd)
    eng = m._train_engine()
    e_img = rel(eng.d["img"][:B], rd["G"])
    print("\n[pggan tiny %d/%s] G rel-L2 %.3e  D_loss %.5f (oracle %.5f)  gp %.4f/%.4f (oracle %.4f/%.4f)" % (
        stage, trans, e_img, d_loss, float(rd["D_loss"]), gp, gp2, float(rd["real_gp"]), float(rd["real_gp2"])))
    assert e_img < 1e-3
    for k, n in enumerate(["Dg_logit", "Dx_logit", "Dxmi_logit", "Dx_hat_logit"]):
        assert rel(eng.d["logit"][B * k:B * k + B], rd[n]) < 5e-3, n
    assert rel(eng.d["slope"], rd["slopes"]) < 2e-2 and rel(eng.d["slope2"], rd["slopes2"]) < 2e-2
    assert float(rd["real_gp"]) > 1e-3 and abs(gp - float(rd["real_gp"])) < 5e-2 * max(1.0, float(rd["real_gp"]))
    assert abs(wd - float(rd["wdist"])) < 5e-3 * max(1.0, abs(float(rd["wdist"])))
    grads = eng.get_grads_tf()
    errs = sorted(((rel(grads[n], rd["grads"][n]), n) for n in rd["grads"] if float(rd["grads"][n].abs().max()) > 1e-12),
                  reverse=True)
    print("[pggan tiny] D-run gradient rel-L2: worst %s  median %.3e" % (errs[0], errs[len(errs) // 2][0]))
    assert errs[0][0] < 0.25, errs[:4]
    rg = P.g_step(p, st, f, ocfg, alpha)
    fd[m.cond_noise] = ff["tn_eps_g"]
    _, g_loss, kl = m.run([m.G_optim, m.G_loss, m.G_kl_loss], fd)
    assert abs(g_loss - float(rg["G_loss"])) < 1e-2 * max(1.0, abs(float(rg["G_loss"]))), (g_loss, float(rg["G_loss"]))
    assert abs(kl - float(rg["G_kl_loss"])) < 1e-3 * max(1.0, abs(float(rg["G_kl_loss"])))
    grads = eng.get_grads_tf()
    errs = sorted(((rel(grads[n], rg["grads"][n]), n) for n in rg["grads"] if float(rg["grads"][n].abs().max()) > 1e-12),
                  reverse=True)
    print("[pggan tiny] G-run gradient rel-L2: worst %s  median %.3e" % (errs[0], errs[len(errs) // 2][0]))
    assert errs[0][0] < 0.25, errs[:4]


@pytest.mark.parametrize("stage,trans,batch,precision,ftol", [(5, True, 4, "bf16x3", 1e-3), (5, True, 4, "bf16", 5e-2),
                                                             (7, False, 2, "bf16x3", 1e-3),
                                                             (8, True, 2, "bf16x3", 1e-3)])
def test_reference_width_forward_parity(stage, trans, batch, precision, ftol, tmp_path):
    """the reference's channel schedule (get_nf / get_dnf, pggan.py:339-343) at 64x64 (fade-in), 256x256 and the last
    entry of the schedule, 512x512 during fade-in (16 / 32 channels at full resolution): G and D forward against the
    oracle."""
    ocfg = P.PgganCfg(batch_size=batch, stage=stage, trans=trans)
    p = P.init_params(ocfg, 0, torch.float32)
    f = P.make_feed(ocfg, 7, torch.float32)
    alpha = 0.35
    m = build(ocfg, precision, str(tmp_path), p)
    with torch.no_grad():
        G, mean, ls = P.generator(p, f["z"], f["cond"], f["tn_eps"], ocfg, alpha)
        Dx = P.discriminator(p, f["x"], f["cond"], ocfg, alpha)
        Dg = P.discriminator(p, G, f["cond"], ocfg, alpha)
    img, mean_g, ls_g = m.generator(f["z"], f["cond"], noise=f["tn_eps"], alpha=alpha)
    e_g = rel(img, G)
    e_dx = rel(m.discriminator(f["x"], f["cond"], alpha=alpha), Dx)
    e_dg = rel(m.discriminator(G, f["cond"], alpha=alpha), Dg)
    print("\n[pggan parity stage %d trans %s] %s: G rel-L2 %.3e  D(x) rel-L2 %.3e  D(G) rel-L2 %.3e" % (
        stage, trans, precision, e_g, e_dx, e_dg))
    assert rel(mean_g, mean) < ftol and rel(ls_g, ls) < ftol
    assert e_g < ftol and e_dx < 5 * ftol and e_dg < 5 * ftol


def test_train_two_passes_with_graphs(tmp_path):
    """stage 1, then the transition pass of stage 2 restoring stage 1 (CUDA graphs replay from the second update on,
    the fade-in coefficient reaches them through device memory), then the sampler."""
    o1 = P.PgganCfg(stage=1, trans=False, **TINY)
    m1 = build(o1, "bf16", str(tmp_path), steps=50)
    m1.train(max_updates=4)
    o2 = P.PgganCfg(stage=2, trans=True, **TINY)
    m2 = build(o2, "bf16", str(tmp_path), steps=50)
    m2.train(max_updates=4)
    assert abs(m2.alpha_tra - 4.0 / 50) < 1e-9
    eng = m2._train_engine()
    assert abs(float(eng.ab[0]) - 4.0 / 50) < 1e-6 and eng.replayed_launches > 0
    sc = eng.scalars_dict()
    assert all(np.isfinite(v) for v in sc.values()), sc
    samples = m2.run(m2.sampler, feed_dict={m2.z_sample: np.random.normal(0, 1, (4, 16)),
                                            m2.cond_sample: np.random.normal(0, 1, (4, 32))})
    assert samples.shape == (4, 8, 8, 3) and np.isfinite(samples).all()
