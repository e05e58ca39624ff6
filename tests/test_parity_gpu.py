"""-m gpu: the CUDA path, driven through the reference-facing API (WGanCls) and the C ABI, against
(a) the committed golden fixtures and (b) the CPU oracle run live on the same seeded inputs.

Stated tolerances (relative L2 unless noted):
  precision "bf16x3" (parity mode):  G / D forward <= 1e-3 (north-star bar; measured ~1e-5),
      losses 1e-3, parameter gradients 5e-2 worst case (ReLU/LeakyReLU sign flips at |x| ~ 1e-6
      move single elements), gradient norms 2e-2.
  precision "bf16" (throughput mode): forward <= 3e-2, reported alongside; gradients sanity only.
"""
import os

import numpy as np
import pytest
import torch

from oracle import wgancls_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TINY = dict(batch_size=4, z_dim=8, embed_dim=32, compressed_embed_dim=8, gf_dim=8, df_dim=8)


def cfg_for(o):
    from t2i_b200.utils.config import AttrDict
    return AttrDict({"MODEL": {"Z_DIM": o.z_dim, "OUTPUT_SIZE": 64, "EMBED_DIM": o.embed_dim,
                               "COMPRESSED_EMBED_DIM": o.compressed_embed_dim, "GF_DIM": o.gf_dim, "DF_DIM": o.df_dim,
                               "IMAGE_SHAPE": {"W": 64, "H": 64, "D": 3}},
                     "TRAIN": {"BATCH_SIZE": o.batch_size, "SAMPLE_NUM": 4, "D_LR": o.d_lr, "G_LR": o.g_lr,
                               "BETA1": o.beta1, "BETA2": o.beta2, "N_CRITIC": 1,
                               "COEFF": {"KL": o.kl_coeff, "LAMBDA": 100.0}}})


def rel(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().cpu().reshape(-1)
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def feed_dict(m, f, which):
    return {m.x: f["x"], m.x_mismatch: f["x_mismatch"], m.cond: f["cond"], m.z: f["z"], m.epsilon: f["epsilon"],
            m.cond_noise: f[which], m.learning_rate_d: 1e-4, m.learning_rate_g: 1e-4}


def build(ocfg, precision, params):
    from t2i_b200.models.wgancls.model import WGanCls
    m = WGanCls(cfg_for(ocfg), precision=precision)
    m.set_variables({k: v for k, v in params.items()})
    return m


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_tiny_iteration_against_golden(precision):
    z = np.load(os.path.join(GOLD, "wgancls_tiny.npz"))
    ocfg = O.OracleCfg(**TINY)
    p = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p/")}
    f = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("f/")}
    m = build(ocfg, precision, p)
    tight = precision == "bf16x3"
    d_loss = m.run([m.D_optim, m.kt_optim, m.D_loss], feed_dict(m, f, "tn_eps"))[2]
    eng = m._train_engine()
    ftol, stol = (1e-3, 1e-3) if tight else (6e-2, 0.25)
    assert rel(eng.d["img"][:4], z["o/G"]) < ftol
    lg = eng.d["logit"].cpu()
    # segment 3 was overwritten by the tangent pass only in its activations, not in the logits
    for i, k in enumerate(["Dg_logit", "Dx_logit", "Dxmi_logit", "Dx_hat_logit"]):
        assert rel(lg[4 * i:4 * i + 4], z["o/" + k]) < ftol * 5, k
    # dD/dx_hat on this 8-channel net: ONE LeakyReLU unit whose pre-activation lies inside the split-bf16
    # rounding band (|a| ~ 1e-6; the golden feed has two such units) flips its derivative and moves the
    # gradient by ~1e-2 -- the mask effect described in the module docstring, not an arithmetic error
    # (tests/test_engine_cpu.py checks the schedule itself to 1e-9).  Forward quantities keep the 1e-3 bar.
    gtol = 3e-2 if tight else 0.5
    assert rel(eng.d["gx"], z["o/grad_x_hat"]) < gtol
    assert rel(eng.d["g2"], z["o/grad_cond"]) < gtol
    sc = eng.scalars_dict()
    assert abs(d_loss - sc["D_loss"]) == 0
    for k in ["D_loss", "D_loss_real", "D_loss_fake", "D_loss_mismatch", "wdist", "wdist2", "reg_loss", "balance_loss",
              "real_gp", "real_gp2"]:
        assert abs(sc[k] - float(z["o/" + k])) < stol * max(1.0, abs(float(z["o/" + k]))), (k, sc[k], float(z["o/" + k]))
    if tight:
        assert abs(sc["kt"] - float(z["o/kt"])) < 1e-6
        grads = eng.get_grads_tf()
        for n in m.d_vars:
            ref = float(z["gn/" + n])
            if ref > 1e-12:
                got = float(grads[n].double().norm())
                assert abs(got - ref) < 2e-2 * ref, (n, got, ref)
    g_loss = m.run([m.G_optim, m.G_loss], feed_dict(m, f, "tn_eps_g"))[1]
    assert abs(g_loss - float(z["o/G_loss"])) < stol * abs(float(z["o/G_loss"]))
    assert rel(eng.d["img"][:4], z["o/G_run_G"]) < ftol
    if tight:
        grads = eng.get_grads_tf()
        for n in m.g_vars:
            ref = float(z["gn/" + n])
            if ref > 1e-9:
                got = float(grads[n].double().norm())
                assert abs(got - ref) < 5e-2 * ref, (n, got, ref)
        # parameters after the whole iteration: |step| ~ lr for every weight, compare to the oracle's
        newp = m.get_variables()
        for n in m.d_vars + m.g_vars:
            if float(z["gn/" + n]) < 1e-9:
                continue      # biases in front of a training-mode BN: exactly-zero gradient, Adam amplifies noise
            q = torch.from_numpy(z["q/" + n]).double()
            bad = ((newp[n].double() - q).abs() > 2e-5)
            # sign(g) decides the step: isolated elements with |g| ~ 0 may go the other way
            assert float(bad.double().mean()) < 0.02 or int(bad.sum()) <= 2, (n, int(bad.sum()), bad.numel())


def test_tiny_gradients_against_live_oracle():
    """element-wise parameter gradients (both runs, incl. the second-order GP term) vs the fp64 oracle"""
    z = np.load(os.path.join(GOLD, "wgancls_tiny.npz"))
    ocfg = O.OracleCfg(**TINY)
    p = O.OrderedDict((k[2:], torch.from_numpy(z[k]).double()) for k in z.files if k.startswith("p/"))
    f = {k[2:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("f/")}
    m = build(ocfg, "bf16x3", p)
    st = O.new_state(p)
    rd = O.d_step(p, st, f, ocfg)
    ff = {k: v.float() for k, v in f.items()}
    m.run([m.D_optim, m.kt_optim, m.D_loss], feed_dict(m, ff, "tn_eps"))
    grads = m._train_engine().get_grads_tf()
    worst = max((rel(grads[n], rd["grads"][n]), n) for n in m.d_vars if float(rd["grads"][n].abs().max()) > 0)
    # 1e-2, not the 2^-17 of the arithmetic: the golden feed puts two d_net units within ~1e-6 of zero, and a
    # LeakyReLU derivative that flips there moves the layers behind it by a few 1e-3 (see the note in
    # test_tiny_iteration_against_golden); a schedule / formula error would show as O(1).
    assert worst[0] < 1e-2, worst


@pytest.mark.parametrize("precision,ftol", [("bf16x3", 1e-3), ("bf16", 3e-2)])
def test_full_width_forward_parity_batch16(precision, ftol):
    """BASELINE config 1 (batch 16, GF = DF = 128, 1024-d embeddings): G and D forward vs the oracle."""
    ocfg = O.OracleCfg(batch_size=16)
    p = O.init_params(ocfg, 0, torch.float32)
    f = O.make_feed(ocfg, 1234, torch.float32)
    m = build(ocfg, precision, p)
    with torch.no_grad():
        G, mean, ls = O.generator(p, f["z"], f["cond"], f["tn_eps"], ocfg)
        Dx = O.discriminator(p, f["x"], f["cond"], ocfg)
        Dg = O.discriminator(p, G, f["cond"], ocfg)
    img, mean_g, ls_g = m.generator(f["z"], f["cond"], noise=f["tn_eps"])
    e_g = rel(img, G)
    e_dx = rel(m.discriminator(f["x"], f["cond"]), Dx)
    e_dg = rel(m.discriminator(G, f["cond"]), Dg)
    print("\n[parity] %s: G rel-L2 %.3e  D(x) rel-L2 %.3e  D(G) rel-L2 %.3e" % (precision, e_g, e_dx, e_dg))
    assert rel(mean_g, mean) < ftol and rel(ls_g, ls) < ftol
    assert e_g < ftol and e_dx < ftol and e_dg < ftol
    assert img.shape == (16, 64, 64, 3) and float(img.abs().max()) <= 1.0


def test_full_width_iteration_against_golden_b4():
    """128-wide net, batch 4, one whole iteration in parity mode vs the committed golden."""
    z = np.load(os.path.join(GOLD, "wgancls_full_b4.npz"))
    ocfg = O.OracleCfg(batch_size=4)
    p = O.init_params(ocfg, 0, torch.float32)
    p["d_net/dense/kernel"] = p["d_net/dense/kernel"] * 4.0
    f = O.make_feed(ocfg, 1234, torch.float32)
    m = build(ocfg, "bf16x3", p)
    m.run([m.D_optim, m.kt_optim, m.D_loss], feed_dict(m, f, "tn_eps"))
    eng = m._train_engine()
    assert rel(eng.d["img"][:4], z["o/G"]) < 1e-3
    lg = eng.d["logit"].cpu()
    for i, k in enumerate(["Dg_logit", "Dx_logit", "Dxmi_logit", "Dx_hat_logit"]):
        assert rel(lg[4 * i:4 * i + 4], z["o/" + k]) < 1e-3, k
    # Input gradients of a 1.2M-unit LeakyReLU net: a handful of units with |pre-activation| below the
    # 1e-5 arithmetic difference flip their mask and each moves the gradient by ~0.5 %; the forward
    # quantities above are tight, the mask-dependent ones below are bounded accordingly.
    assert rel(eng.d["gx"], z["o/grad_x_hat"]) < 3e-2
    assert rel(eng.d["g2"], z["o/grad_cond"]) < 3e-2
    sc = eng.scalars_dict()
    assert float(z["o/real_gp"]) > 0 and float(z["o/real_gp2"]) > 0
    for k in ["wdist", "wdist2", "balance_loss", "D_loss_real", "D_loss_fake", "D_loss_mismatch"]:
        assert abs(sc[k] - float(z["o/" + k])) < 1e-3 * max(1.0, abs(float(z["o/" + k]))), (k, sc[k], float(z["o/" + k]))
    for k in ["D_loss", "real_gp", "real_gp2"]:
        assert abs(sc[k] - float(z["o/" + k])) < 6e-2 * abs(float(z["o/" + k])), (k, sc[k], float(z["o/" + k]))
    grads = eng.get_grads_tf()
    for n in m.d_vars:
        ref = float(z["gn/" + n])
        if ref > 1e-9:
            got = float(grads[n].double().norm())
            assert abs(got - ref) < 6e-2 * ref, (n, got, ref)
    g_loss = m.run([m.G_optim, m.G_loss], feed_dict(m, f, "tn_eps_g"))[1]
    assert abs(g_loss - float(z["o/G_loss"])) < 2e-3 * abs(float(z["o/G_loss"]))
    grads = eng.get_grads_tf()
    for n in m.g_vars:
        ref = float(z["gn/" + n])
        if ref > 1e-6:
            got = float(grads[n].double().norm())
            # small bias vectors are sums with heavy cancellation: one flipped unit in d_net moves them coherently
            assert abs(got - ref) < (6e-2 if grads[n].numel() >= 1024 else 0.25) * ref, (n, got, ref)


def test_properties_at_bench_size():
    """BASELINE config 2 size (batch 256, bf16): size-independent properties instead of the oracle:
    D is per-sample (a 4B batch equals four B batches), x_hat(eps=0) = x and x_hat(eps=1) = G,
    determinism, finite losses, kt moves by -1e-3 * its gradient."""
    from t2i_b200.models.wgancls.model import WGanCls
    ocfg = O.OracleCfg(batch_size=256)
    m = WGanCls(cfg_for(ocfg), precision="bf16")
    m.initialize(0)
    f = O.make_feed(ocfg, 1234, torch.float32)
    eng = m._train_engine()
    big = m.discriminator(torch.cat([f["x"], f["x_mismatch"]])[:256], torch.cat([f["cond"], f["cond"]])[:256])
    a = m.discriminator(f["x"][:128], f["cond"][:128])
    assert torch.equal(big[:128], a)                                      # bit-exact: batch-independent tiles
    fd = feed_dict(m, f, "tn_eps")
    fd[m.epsilon] = torch.zeros(256, 1, 1, 1)
    xh = m.run([m.D_optim, m.kt_optim, m.D_loss, m.x_hat], fd)[3]
    assert np.array_equal(xh, eng.d["img"][256:512].cpu().numpy())         # x_hat == x
    kt0 = float(eng.kt.item())
    fd[m.epsilon] = torch.ones(256, 1, 1, 1)
    d_loss, xh = m.run([m.D_optim, m.kt_optim, m.D_loss, m.x_hat], fd)[2:]
    assert np.array_equal(xh, eng.d["img"][:256].cpu().numpy())            # x_hat == G
    sc = eng.scalars_dict()
    assert np.isfinite(d_loss) and all(np.isfinite(v) for v in sc.values())
    assert abs(float(eng.kt.item()) - (kt0 - 1e-3 * sc["kt_grad"])) < 1e-6
    g_loss = m.run([m.G_optim, m.G_loss], feed_dict(m, f, "tn_eps_g"))[1]
    assert np.isfinite(g_loss)


def test_graph_replay_and_side_stream_match_eager_schedule():
    """The same iteration from the same state three times: eager (first call), captured + replayed
    (second call), pure replay (third call), with side-stream weight gradients; and once more on a
    model without graphs and without the side stream.  Losses and updated weights must agree up to the
    fp32 atomic-accumulation order of the reduction / weight-gradient kernels."""
    from t2i_b200.models.wgancls.model import WGanCls
    ocfg = O.OracleCfg(batch_size=32)
    f = O.make_feed(ocfg, 77, torch.float32)
    names = ("d_net/Conv_2/weights", "g_net/Conv_3/weights", "g_net/BatchNorm_4/moving_mean", "d_net/dense/kernel")

    def one_iteration(m, v0):
        eng = m._train_engine()
        m.set_variables(v0)
        for k in ("d", "g"):
            eng.adam_m[k].zero_(); eng.adam_v[k].zero_()
        eng.d_t = eng.g_t = 0
        d = m.run([m.D_optim, m.kt_optim, m.D_loss, m.real_gp, m.wdist, m.kt], feed_dict(m, f, "tn_eps"))
        g = m.run([m.G_optim, m.G_loss], feed_dict(m, f, "tn_eps_g"))
        v = m.get_variables()
        return torch.tensor([d[2], d[3], d[4], d[5], g[1]], dtype=torch.float64), {n: v[n] for n in names}

    m = WGanCls(cfg_for(ocfg), precision="bf16x3", use_graphs=True)
    m.initialize(3)
    v0 = m.get_variables()
    runs = [one_iteration(m, v0) for _ in range(3)]
    eng = m._train_engine()
    assert all(eng._graphs[k]["graph"] is not None for k in ("d_a", "d_b", "d_c", "g_a1", "g_a2", "g_b", "g_c"))
    assert eng.replayed_launches > 0 and eng.side_stream is not None
    plain = WGanCls(cfg_for(ocfg), precision="bf16x3", use_graphs=False)
    plain._train_engine().side_stream = None
    runs.append(one_iteration(plain, v0))
    s0, w0 = runs[0]
    assert torch.isfinite(s0).all()
    for s, w in runs[1:]:
        assert float(((s - s0).abs() / (1.0 + s0.abs())).max()) < 1e-3, (s, s0)
        for n in names:   # one Adam step of ~lr: identical except where a near-zero gradient flips its sign
            assert float(((w[n] - w0[n]).abs() > 5e-5).double().mean()) < 0.03, n


def _d_masked(p, x_nhwc, cond, slopes):
    """models/wgancls/model.py:129-161 with every LeakyReLU replaced by a FIXED per-element slope (1 or 0.2): inside the
    linear region the CUDA path's saved activation signs select, d_net is exactly this map."""
    import torch.nn.functional as F
    d = "d_net/"

    def conv(scope, x, k, s, pad):
        return F.conv2d(x, p[scope + "/weights"].permute(3, 2, 0, 1), p[scope + "/biases"], stride=s, padding=pad)

    x = x_nhwc.permute(0, 3, 1, 2)
    h0 = conv(d + "Conv", x, 4, 2, 1) * slopes["a0"]
    h1 = conv(d + "Conv_1", h0, 4, 2, 1) * slopes["a1"]
    h2 = conv(d + "Conv_2", h1, 4, 2, 1) * slopes["a2"]
    h3 = conv(d + "Conv_3", h2, 4, 2, 1)
    n = conv(d + "Conv_4", h3, 1, 1, 0) * slopes["r1"]
    n = conv(d + "Conv_5", n, 3, 1, 1) * slopes["r2"]
    n = conv(d + "Conv_6", n, 3, 1, 1)
    h4 = (h3 + n) * slopes["h4"]
    e = (cond @ p[d + "dense/kernel"] + p[d + "dense/bias"]) * slopes["e"]
    h4c = torch.cat([h4, e[:, :, None, None].expand(-1, -1, 4, 4)], 1)
    h5 = conv(d + "Conv_7", h4c, 3, 1, 1) * slopes["a5"]
    h6 = conv(d + "Conv_8", h5, 1, 1, 0) * slopes["a6"]
    return conv(d + "Conv_9", h6, 4, 4, 0)


def _slopes_from_engine(eng, s0, n, df8):
    """slope tensors (NCHW, fp64) of samples [s0, s0 + n) from the engine's post-activation buffers"""
    def sl(t, c0=0, c1=None):
        v = t.float().sum(0)[s0:s0 + n]
        v = v[..., c0:c1] if v.dim() == 4 else v[:, c0:c1]
        s = torch.where(v > 0, 1.0, 0.2).double().cpu()
        return s.permute(0, 3, 1, 2) if s.dim() == 4 else s
    d = eng.d
    return {"a0": sl(d["a0"]), "a1": sl(d["a1"]), "a2": sl(d["a2"]), "r1": sl(d["r1"]), "r2": sl(d["r2"]),
            "h4": sl(d["cat"], 0, df8), "e": sl(d["e"]), "a5": sl(d["a5"]), "a6": sl(d["a6"])}


def test_gradient_penalty_path_against_mask_consistent_oracle():
    """The tolerance of the gradient-penalty quantities, made tight.  d_net is piecewise linear; units whose
    pre-activation is below the arithmetic difference between CUDA and CPU take the other LeakyReLU branch, which moves
    dD/dx_hat by percents although nothing is wrong.  Here the CPU side evaluates the SAME linear piece the CUDA path
    is on (its saved activation signs, all four segments) in fp64: slopes, both penalties, D_loss and every d_net
    gradient (first-order + the second-order penalty term) must then agree to 1e-3 (model.py:62-70,88-97)."""
    ocfg = O.OracleCfg(batch_size=4)
    p = O.init_params(ocfg, 0, torch.float32)
    p["d_net/dense/kernel"] = p["d_net/dense/kernel"] * 4.0       # makes the cond penalty active (as in the b4 golden)
    f = O.make_feed(ocfg, 1234, torch.float32)
    m = build(ocfg, "bf16x3", p)
    eng = m._train_engine()
    B, df8 = 4, 8 * ocfg.df_dim
    out = m.run([m.D_optim, m.kt_optim, m.D_loss, m.G, m.x_hat], feed_dict(m, f, "tn_eps"))
    G_img, x_hat = torch.from_numpy(out[3]).double(), torch.from_numpy(out[4]).double()
    sc = eng.scalars_dict()
    gx, g2 = eng.d["gx"].double().cpu(), eng.d["g2"].double().cpu()
    grads = eng.get_grads_tf()
    # signs: segments 0..2 are still in the buffers; the x_hat segment was overwritten by the tangent pass, so its
    # forward is repeated on a second model that still holds the weights of BEFORE the update (bit-identical logits:
    # d_net is per-sample) and read back
    slopes = [_slopes_from_engine(eng, s * B, B, df8) for s in range(3)]
    m2 = build(ocfg, "bf16x3", p)
    lg = m2.discriminator(x_hat.float(), f["cond"])
    assert torch.equal(lg.reshape(-1).cpu(), eng.d["logit"][3 * B:].cpu())
    slopes.append(_slopes_from_engine(m2._train_engine(), 0, B, df8))
    pd = {k: v.double() for k, v in p.items()}
    names = [n for n in pd if n.startswith("d_net/")]
    for n in names:
        pd[n] = pd[n].clone().requires_grad_(True)
    cond = f["cond"].double()
    kt = torch.tensor(O.KT_INIT, dtype=torch.float64)
    Dg = _d_masked(pd, G_img, cond, slopes[0])
    Dx = _d_masked(pd, f["x"].double(), cond, slopes[1])
    Dm = _d_masked(pd, f["x_mismatch"].double(), cond, slopes[2])
    xh = x_hat.clone().requires_grad_(True)
    ci = cond.clone().requires_grad_(True)
    Dh = _d_masked(pd, xh, ci, slopes[3])
    g_x, g_c = torch.autograd.grad(Dh.sum(), [xh, ci], create_graph=True)
    gp, s1 = O.gradient_penalty(g_x, (1, 2, 3))
    gp2, s2 = O.gradient_penalty(g_c, (1,))
    wdist, wdist2 = Dx.mean() - Dg.mean(), Dx.mean() - Dm.mean()
    d_loss = -wdist - kt * wdist2 + O.GP_WEIGHT * (gp + gp2)
    ref_grads = torch.autograd.grad(d_loss, [pd[n] for n in names])
    assert float(gp) > 0 and float(gp2) > 0          # both penalties are active on this feed
    e_gx, e_g2 = rel(gx, g_x.detach()), rel(g2, g_c.detach())
    print("\n[gp] mask-consistent oracle: dD/dx_hat %.2e  dD/dcond %.2e  real_gp %.6f vs %.6f  real_gp2 %.6f vs %.6f" % (
        e_gx, e_g2, sc["real_gp"], float(gp), sc["real_gp2"], float(gp2)))
    assert e_gx < 1e-3 and e_g2 < 1e-3
    for k, v in (("real_gp", gp), ("real_gp2", gp2), ("D_loss", d_loss), ("wdist", wdist), ("wdist2", wdist2)):
        assert abs(sc[k] - float(v)) < 1e-3 * max(1.0, abs(float(v))), (k, sc[k], float(v))
    worst = max((rel(grads[n], g), n) for n, g in zip(names, ref_grads) if float(g.abs().max()) > 0)
    print("[gp] worst d_net gradient (first order + second-order penalty term) rel-L2 %.2e (%s)" % worst)
    assert worst[0] < 1e-3, worst
