"""Committed golden fixtures of the widened rows (tests/golden/make_golden_next_rows.py): StackGAN stage-I / stage-II
and conditional PGGAN tiny nets.

not gpu: the oracles reproduce their goldens (an accidental change of an oracle, of torch's CPU kernels or of the seeded
helpers shows up here).  gpu: the CUDA path, through the reference-facing APIs, against the COMMITTED values (forward
quantities at parity-mode tolerances; gradient norms loosely: derivative flips of units near zero, see
tests/test_parity_gpu.py)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from oracle import pggan_oracle as P
from oracle import stackgan1_oracle as S1
from oracle import stackgan2_oracle as S2
import test_pggan_cpu as TP
import test_stackgan1_cpu as T1
import test_stackgan2_cpu as T2

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name))
    return {k: z[k] for k in z.files}


def fp32_exact(d):
    return type(d)((k, v.float().double()) for k, v in d.items())


def feed_of(gold, make_feed_fallback=None):
    f = {k[2:]: torch.as_tensor(v).double() for k, v in gold.items() if k.startswith("f/")}
    if make_feed_fallback is not None:
        for k, v in make_feed_fallback.items():
            f.setdefault(k, v)
    return f


def check(gold, rd, rg, rtol):
    for k, v in gold.items():
        tag, name = k.split("/", 1)
        if tag == "f":
            continue
        got = {"d": rd, "g": rg}[tag[0]]
        if tag in ("dn", "gn"):
            val = float(got["grads"][name].double().norm())
            assert abs(val - float(v)) <= rtol * max(abs(float(v)), 1e-9) + 1e-12, (k, val, float(v))
        elif name == "G_mean_abs":
            assert abs(float(got["G"].abs().mean()) - float(v)) < rtol * float(v)
        elif name == "G_corner":
            np.testing.assert_allclose(got["G"][:, :8, :8].numpy(), v, rtol=rtol, atol=rtol)
        else:
            np.testing.assert_allclose(np.asarray(got[name].numpy()), v, rtol=rtol, atol=rtol * 10, err_msg=k)


def test_stage1_oracle_reproduces_golden():
    gold = load("stackgan1_tiny.npz")
    cfg = S1.Stage1Cfg(**T1.TINY)
    p = fp32_exact(T1.boosted_params(cfg))
    feed = feed_of(gold)
    st = S1.new_state(p)
    rd = S1.d_step(p, st, feed, cfg)
    rg = S1.g_step(p, st, feed, cfg)
    check(gold, rd, rg, 1e-8)


def test_stage2_oracle_reproduces_golden():
    gold = load("stackgan2_tiny.npz")
    cfg = S2.Stage2Cfg(**T2.TINY)
    p = fp32_exact(T2.boosted_params(cfg))
    feed = feed_of(gold, fp32_exact(S2.make_feed(cfg, 21, torch.float64)))
    st = S2.new_state(p)
    rd = S2.d_step(p, st, feed, cfg)
    rg = S2.g_step(p, st, feed, cfg)
    check(gold, rd, rg, 1e-8)


@pytest.mark.parametrize("name,stage,trans", [("pggan_tiny_s2t.npz", 2, True), ("pggan_tiny_s3.npz", 3, False)])
def test_pggan_oracle_reproduces_golden(name, stage, trans):
    gold = load(name)
    cfg = P.PgganCfg(stage=stage, trans=trans, **TP.TINY)
    p = fp32_exact(TP.boosted_params(cfg))
    feed = feed_of(gold)
    st = P.new_state(p)
    rd = P.d_step(p, st, feed, cfg, 0.3)
    rg = P.g_step(p, st, feed, cfg, 0.3)
    check(gold, rd, rg, 1e-8)


# ----------------------------------------------------------------------------------------------------------- CUDA path
def _rel(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().cpu().reshape(-1)
    b = torch.as_tensor(np.asarray(b)).double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("name,stage,trans", [("pggan_tiny_s2t.npz", 2, True), ("pggan_tiny_s3.npz", 3, False)])
def test_pggan_cuda_iteration_against_golden(name, stage, trans, tmp_path):
    from test_pggan_gpu import build
    gold = load(name)
    ocfg = P.PgganCfg(stage=stage, trans=trans, **TP.TINY)
    p = fp32_exact(TP.boosted_params(ocfg))
    m = build(ocfg, "bf16x3", str(tmp_path), p, steps=100)
    f = {k: v.float() for k, v in feed_of(gold).items()}
    B = ocfg.batch_size
    fd = {m.x: f["x"], m.x_mismatch: f["x_mismatch"], m.cond: f["cond"], m.z: f["z"], m.epsilon: f["epsilon"],
          m.cond_noise: f["tn_eps"], m.iter: 30}
    _, d_loss, wd, gp, gp2 = m.run([m.D_optim, m.D_loss, m.wdist, m.real_gp, m.real_gp2], fd)
    eng = m._train_engine()
    assert _rel(eng.d["img"][:B], gold["d/G"]) < 1e-3 and _rel(eng.d["img"][3 * B:], gold["d/x_hat"]) < 1e-3
    for k, n in enumerate(["Dg_logit", "Dx_logit", "Dxmi_logit", "Dx_hat_logit"]):
        assert _rel(eng.d["logit"][B * k:B * k + B], gold["d/" + n]) < 5e-3, n
    assert _rel(eng.d["gx"], gold["d/grad_x_hat"]) < 2e-2 and _rel(eng.d["g2"], gold["d/grad_cond"]) < 2e-2
    assert _rel(eng.d["slope"], gold["d/slopes"]) < 2e-2 and _rel(eng.d["slope2"], gold["d/slopes2"]) < 2e-2
    for got, n in ((d_loss, "D_loss"), (wd, "wdist"), (gp, "real_gp"), (gp2, "real_gp2")):
        assert abs(got - float(gold["d/" + n])) < 5e-2 * max(1.0, abs(float(gold["d/" + n]))), (n, got)
    grads = eng.get_grads_tf()
    for n in (k[3:] for k in gold if k.startswith("dn/")):
        ref = float(gold["dn/" + n])
        if ref > 1e-9:
            assert abs(float(grads[n].double().norm()) - ref) < 0.1 * ref, n
    fd[m.cond_noise] = f["tn_eps_g"]
    _, g_loss, kl = m.run([m.G_optim, m.G_loss, m.G_kl_loss], fd)
    assert abs(g_loss - float(gold["g/G_loss"])) < 1e-2 * max(1.0, abs(float(gold["g/G_loss"])))
    assert abs(kl - float(gold["g/G_kl_loss"])) < 1e-3 * max(1.0, abs(float(gold["g/G_kl_loss"])))
    assert _rel(eng.d["img"][:B], gold["g/G"]) < 1e-3
    grads = eng.get_grads_tf()
    for n in (k[3:] for k in gold if k.startswith("gn/")):
        ref = float(gold["gn/" + n])
        if ref > 1e-9:
            assert abs(float(grads[n].double().norm()) - ref) < 0.1 * ref, n


@pytest.mark.gpu
def test_stage1_cuda_forward_against_golden():
    from test_stackgan1_gpu import build, feeds, trainer_for
    gold = load("stackgan1_tiny.npz")
    ocfg = S1.Stage1Cfg(**T1.TINY)
    p = fp32_exact(T1.boosted_params(ocfg))
    m = build(ocfg, "bf16x3", p)
    tr = trainer_for(m)
    f = {k: v.float() for k, v in feed_of(gold).items()}
    _, d_loss = m.run([tr.D_optim, tr.D_loss], feeds(m, tr, f, "tn_eps"))
    eng = m._train_engine()
    B = ocfg.batch_size
    assert _rel(eng.d["img"][:B], gold["d/G"]) < 1e-3
    for k, n in enumerate(["D_synthetic_logits", "D_real_match_logits", "D_real_mismatch_logits"]):
        assert _rel(eng.d["logit"][B * k:B * k + B], gold["d/" + n]) < 5e-3, n
    assert abs(d_loss - float(gold["d/D_loss"])) < 2e-3 * max(1.0, abs(float(gold["d/D_loss"])))


@pytest.mark.gpu
def test_stage2_cuda_forward_against_golden():
    from test_stackgan2_gpu import build, feeds, trainer_for
    gold = load("stackgan2_tiny.npz")
    ocfg = S2.Stage2Cfg(**T2.TINY)
    p = fp32_exact(T2.boosted_params(ocfg))
    s1, m, c1, c2 = build(ocfg, "bf16x3", p)
    tr = trainer_for(m, c2, c1)
    f = {k: v.float() for k, v in feed_of(gold, fp32_exact(S2.make_feed(ocfg, 21, torch.float64))).items()}
    _, d_loss = m.run([tr.D_optim, tr.D_loss], feeds(m, tr, f, "tn_eps", "tn_s1"))
    eng = m._train_engine()
    B = ocfg.batch_size
    assert _rel(eng.g["img64"], gold["d/stage1_G"]) < 1e-3
    assert _rel(eng.d["img"][:B, :8, :8], gold["d/G_corner"]) < 5e-3
    for k, n in enumerate(["D_synthetic_logits", "D_real_match_logits", "D_real_mismatch_logits"]):
        assert _rel(eng.d["logit"][B * k:B * k + B], gold["d/" + n]) < 1e-2, n
    assert abs(d_loss - float(gold["d/D_loss"])) < 5e-3 * max(1.0, abs(float(gold["d/D_loss"])))
