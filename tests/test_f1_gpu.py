"""-m gpu: the rows around the hot path on the real kernels and the benched configuration itself.

  * BASELINE config 2 exactly (batch 256, GF = DF = 128, bf16 throughput mode AND bf16x3 parity mode): G and D forward
    against the CPU oracle, with the tolerance each mode states.
  * the trainer mirror (models/wgancls/trainer.py:73-127) for three iterations on CUDA against three oracle iterations
    replayed on the recorded feeds: losses, kt, Adam step counters, moving statistics, weights, and the sampler
    (model.py:57: is_training=False -> moving statistics) after those three G steps.
  * the remaining fetches of build_model (x_hat, the four logits, embed_mean / embed_log_sigma).
  * two ranks over NCCL (skipped below two GPUs): the sharded iteration with sync_bn equals one GPU on the whole batch.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import wgancls_oracle as O
from test_parity_gpu import TINY, build, cfg_for, feed_dict, rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Stated forward tolerances (relative L2 against the fp32 oracle):
#   bf16x3 (parity mode): 1e-3, the north-star bar (measured ~5e-5 / 1e-4).
#   bf16 (throughput mode, the benched arithmetic): G 3e-2, D 1.5e-2.  This is the storage format's error, not a
#   kernel defect: tools/bf16_error_budget.py reproduces it on the CPU by rounding the same ~26 tensors (and the
#   weights) to bf16 inside the fp32 oracle -- every site contributes 2.5-4.5e-3, DESIGN.md section 5.
FWD_TOL = {"bf16x3": (1e-3, 1e-3), "bf16": (3e-2, 1.5e-2)}


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_bench_config_forward_vs_oracle_batch256(precision):
    """BASELINE config 2 (the configuration bench.py times): batch 256, full width."""
    ocfg = O.OracleCfg(batch_size=256)
    p = O.init_params(ocfg, 0, torch.float32)
    f = O.make_feed(ocfg, 1234, torch.float32)
    m = build(ocfg, precision, p)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        G, mean, ls = O.generator(p, f["z"], f["cond"], f["tn_eps"], ocfg)
        Dx = O.discriminator(p, f["x"], f["cond"], ocfg)
        Dg = O.discriminator(p, G, f["cond"], ocfg)
    img, mean_g, ls_g = m.generator(f["z"], f["cond"], noise=f["tn_eps"])
    e_g = rel(img, G)
    e_dx = rel(m.discriminator(f["x"], f["cond"]), Dx)
    e_dg = rel(m.discriminator(G, f["cond"]), Dg)
    e_ms = max(rel(mean_g, mean), rel(ls_g, ls))
    print("\n[parity b256] %s: G rel-L2 %.3e  D(x) %.3e  D(G) %.3e  mean/log_sigma %.3e" % (precision, e_g, e_dx, e_dg, e_ms))
    gtol, dtol = FWD_TOL[precision]
    assert e_g < gtol and e_dx < dtol and e_dg < dtol
    assert e_ms < 1e-3          # the conditioning head runs in fp32 in both modes
    assert img.shape == (256, 64, 64, 3) and float(img.abs().max()) <= 1.0


class _Recorder(object):
    """records every truncated-normal draw the model makes (model.py:119 draws inside the graph)"""

    def __init__(self, module):
        self.module, self.orig, self.draws = module, module._truncated_normal, []

    def __enter__(self):
        def rec(shape, device, generator=None):
            t = self.orig(shape, device, generator)
            self.draws.append(t.detach().cpu().clone())
            return t
        self.module._truncated_normal = rec
        return self

    def __exit__(self, *a):
        self.module._truncated_normal = self.orig


def test_trainer_three_iterations_and_sampler_vs_oracle(tmp_path):
    from t2i_b200.models.wgancls import model as model_mod
    from t2i_b200.models.wgancls.trainer import SyntheticTextDataset, WGanClsTrainer
    ocfg = O.OracleCfg(**TINY)
    p = O.init_params(ocfg, 3, torch.float64)
    cfg = cfg_for(ocfg)
    cfg.CHECKPOINT_DIR = str(tmp_path / "ckpt")
    cfg.TRAIN.MAX_STEPS, cfg.TRAIN.SUMMARY_PERIOD, cfg.TRAIN.SAMPLE_PERIOD, cfg.TRAIN.CHECKPOINTS_TO_KEEP = 4, 1, 3, 2
    m = model_mod.WGanCls(cfg, precision="bf16x3")
    m.seed_noise(11)
    runs, samples = [], []
    real_run, real_init = m.run, m.initialize

    def run(fetches, feed_dict=None):
        names = [f.name for f in (fetches if isinstance(fetches, (list, tuple)) else [fetches])]
        runs.append((names, {k.name: v for k, v in (feed_dict or {}).items()}))
        return real_run(fetches, feed_dict)

    def initialize(seed=0):            # the trainer initialises, then we load the oracle's parameters
        real_init(seed)
        m.set_variables({k: v.float() for k, v in p.items()})

    m.run, m.initialize = run, initialize
    np.random.seed(5)
    tr = WGanClsTrainer(None, m, SyntheticTextDataset(embed_dim=ocfg.embed_dim, num_examples=64), cfg,
                        on_samples=lambda idx, s, c: samples.append((idx, s)))
    with _Recorder(model_mod) as rec:
        tr.train()
    d_runs = [r for r in runs if "D_optim" in r[0]]
    g_runs = [r for r in runs if "G_optim" in r[0]]
    assert len(d_runs) == 3 and len(g_runs) == 3 and len(samples) == 1 and samples[0][0] == 3
    assert len(rec.draws) == 7          # D run + G run per iteration, then the sampler's draw
    eng = m._train_engine()
    assert (eng.d_t, eng.g_t, m.global_step) == (3, 3, 3)
    # replay on the oracle
    st = O.new_state(p)
    for it in range(3):
        fd = d_runs[it][1]
        f = {"x": fd["real_images"], "x_mismatch": fd["wrong_images"], "cond": fd["cond"], "z": fd["z"], "epsilon": fd["eps"]}
        f = {k: torch.as_tensor(np.asarray(v)).double() for k, v in f.items()}
        f["tn_eps"], f["tn_eps_g"] = rec.draws[2 * it].double(), rec.draws[2 * it + 1].double()
        rd, rg = O.iteration(p, st, f, ocfg)
        log = tr.log[it]
        # iteration 0 starts from identical weights: 2e-3.  Later iterations start from weights that differ in the few
        # elements whose near-zero gradient took the other sign in the Adam step before (|step| ~ lr whatever |g| is);
        # on this 8-channel net single weights move the penalty terms by a few 1e-3: 2e-2.  A wrong Adam step count,
        # kt order or moving-average update would be O(1) here and in the weight deltas checked below.
        tol_it = 2e-3 if it == 0 else 2e-2
        for k in ("D_loss", "D_loss_real", "D_loss_fake", "D_loss_mismatch", "wdist", "wdist2", "real_gp", "real_gp2",
                  "balance_loss", "reg_loss"):
            assert abs(log[k] - float(rd[k])) < tol_it * max(1.0, abs(float(rd[k]))), (it, k, log[k], float(rd[k]))
        assert abs(log["G_loss"] - float(rg["G_loss"])) < tol_it * max(1.0, abs(float(rg["G_loss"]))), (it, log["G_loss"])
        # the summary's kt is read in the D run that also steps it: either value is a legal TF ordering, ours is the new one
        assert abs(log["kt"] - float(st["kt"])) < 1e-5
    assert abs(float(eng.kt.item()) - float(st["kt"])) < 1e-5
    got = m.get_variables()
    p0 = O.init_params(ocfg, 3, torch.float64)
    worst_delta = (0.0, "")
    for n, w in p.items():
        if "moving_" not in n and float(st["v"][n].abs().max()) > 1e-18 and w.numel() >= 64:
            da, db = torch.as_tensor(got[n]).double() - p0[n], w.double() - p0[n]
            worst_delta = max(worst_delta, (rel(da, db), n))
    print("\n[f1] three Adam steps, worst relative L2 of (weights - initial) vs the oracle: %.3f (%s)" % worst_delta)
    assert worst_delta[0] < 0.25, worst_delta     # a step-size error (Adam t, lr_t) of 25 % or a skipped step shows here
    for n, w in p.items():
        a, b = torch.as_tensor(got[n]).double(), w.double()
        if n.endswith("moving_variance"):   # three EMA steps of the batch statistics (utils/ops.py:20-29), G runs only
            assert rel(a, b) < 1e-3, (n, rel(a, b))
        elif n.endswith("moving_mean"):
            # measured against the spread of the normalised tensor: behind another BatchNorm the true channel means are
            # ~1e-6 (beta = 0), pure cancellation noise that a relative comparison would amplify
            spread = torch.sqrt(p[n.replace("moving_mean", "moving_variance")].double())
            err = float((a - b).norm() / spread.norm())
            assert err < 1e-3, (n, err)
        elif float(st["v"][n].abs().max()) > 1e-18:
            # three Adam steps of ~lr each; sign(g) decides a step, isolated near-zero gradients may go the other way
            bad = (a - b).abs() > 6e-5
            assert float(bad.double().mean()) < 0.03 or int(bad.sum()) <= 3, (n, int(bad.sum()), bad.numel())
    # the sampler after three G steps: is_training=False -> moving statistics (model.py:57)
    z_s, c_s = g_runs[-1][1]["z_sample"], g_runs[-1][1]["cond_sample"]
    with torch.no_grad():
        ref, _, _ = O.generator({k: torch.as_tensor(got[k]).double() if k in got else v for k, v in p.items()},
                                torch.as_tensor(z_s).double(), torch.as_tensor(c_s).double(), rec.draws[6].double(), ocfg,
                                is_training=False)
    e = rel(samples[0][1], ref)
    print("\n[f1] sampler after 3 iterations vs oracle (same weights, moving statistics): rel-L2 %.3e" % e)
    assert e < 1e-3
    assert sorted(os.listdir(cfg.CHECKPOINT_DIR)) == ["checkpoint", "wgancls-2.npz"]


def test_sampler_full_width_moving_statistics():
    """full width, batch 16: three iterations move the moving statistics; the sampler (inference-mode BatchNorm through
    bn_apply with moving mean / variance) must match the oracle evaluated on the model's own variables."""
    ocfg = O.OracleCfg(batch_size=16)
    p = O.init_params(ocfg, 0, torch.float32)
    m = build(ocfg, "bf16x3", p)
    for it in range(3):
        f = O.make_feed(ocfg, 100 + it, torch.float32)
        m.run([m.D_optim, m.kt_optim, m.D_loss], feed_dict(m, f, "tn_eps"))
        m.run([m.G_optim, m.G_loss], feed_dict(m, f, "tn_eps_g"))
    v = m.get_variables()
    mm = torch.as_tensor(v["g_net/BatchNorm_4/moving_mean"])
    mv = torch.as_tensor(v["g_net/BatchNorm_4/moving_variance"])
    assert float(mm.abs().max()) > 0 and float((mv - 1).abs().max()) > 1e-3        # they moved
    z = torch.randn(4, ocfg.z_dim, generator=torch.Generator().manual_seed(1))
    c = torch.randn(4, ocfg.embed_dim, generator=torch.Generator().manual_seed(2))
    tn = torch.randn(4, ocfg.compressed_embed_dim, generator=torch.Generator().manual_seed(3)).clamp_(-2, 2)
    got = m.run(m.sampler, {m.z_sample: z, m.cond_sample: c, m.cond_noise_sample: tn})
    pv = {k: torch.as_tensor(v[k]).float() for k in p}
    with torch.no_grad():
        ref, _, _ = O.generator(pv, z, c, tn, ocfg, is_training=False)
        ref_train, _, _ = O.generator(pv, z, c, tn, ocfg, is_training=True)
    e = rel(got, ref)
    print("\n[f1] full-width sampler vs oracle: rel-L2 %.3e (batch-statistics image differs by %.2f)" % (e, rel(ref_train, ref)))
    assert e < 1e-3
    assert rel(ref_train, ref) > 0.05       # the check can tell moving from batch statistics


def test_remaining_graph_fetches():
    """x_hat, Dg / Dx / Dxmi / Dx_hat logits and the conditioning head are fetchable after a D run (model.py:48-55)."""
    ocfg = O.OracleCfg(**TINY)
    p = O.init_params(ocfg, 0, torch.float64)
    f = O.make_feed(ocfg, 7, torch.float64)
    m = build(ocfg, "bf16x3", p)
    ff = {k: v.float() for k, v in f.items()}
    out = m.run([m.D_optim, m.kt_optim, m.D_loss, m.G, m.x_hat, m.Dg_logit, m.Dx_logit, m.Dxmi_logit, m.Dx_hat_logit,
                 m.embed_mean, m.embed_log_sigma], feed_dict(m, ff, "tn_eps"))
    ref = O.d_forward_losses(p, torch.tensor(O.KT_INIT, dtype=torch.float64), f, ocfg, create_graph=False)
    ref = {k: v.detach() for k, v in ref.items()}
    for got, k in zip(out[3:], ["G", "x_hat", "Dg_logit", "Dx_logit", "Dxmi_logit", "Dx_hat_logit", "embed_mean",
                                "embed_log_sigma"]):
        assert got.shape == tuple(ref[k].shape), k
        assert rel(got, ref[k]) < 1e-3, (k, rel(got, ref[k]))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_iteration_equals_single_gpu():
    """tools/dp_check.py under torchrun, 2 ranks, NCCL: batch shards + sync_bn == one GPU on the whole batch."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "[dp_check] OK" in r.stdout
