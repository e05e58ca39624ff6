"""Generates the committed golden fixtures from the CPU oracle (run here, in the dev container):
    python tests/golden/make_golden.py
wgancls_tiny.npz   : 8-channel net, batch 4, fp64 oracle; parameters, feed, and every output.
wgancls_full_b4.npz: the real 128-wide config, batch 4, fp32 oracle; parameters are NOT stored
                     (regenerated from seed 0 by oracle.init_params), only the feed seed and outputs.
The reference itself cannot produce these (TF 1.4 is not installable, it ships no fixtures): the
oracle is the pin, see oracle/wgancls_oracle.py.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import wgancls_oracle as O  # noqa: E402

TINY = dict(batch_size=4, z_dim=8, embed_dim=32, compressed_embed_dim=8, gf_dim=8, df_dim=8)


def boosted(p, dtype):
    p["d_net/dense/kernel"] = p["d_net/dense/kernel"] * 6.0
    g = torch.Generator().manual_seed(3)
    for n in p:
        if n.endswith("biases") or n.endswith("bias") or n.endswith("beta"):
            p[n] = (torch.randn(p[n].shape, generator=g, dtype=torch.float64) * 0.1).to(dtype)
        if n.endswith("gamma"):
            p[n] = (1 + 0.2 * torch.randn(p[n].shape, generator=g, dtype=torch.float64)).to(dtype)
    return p


def run(cfg, p, feed, store_params):
    out = {}
    if store_params:
        for k, v in p.items():
            out["p/" + k] = v.numpy().astype(np.float32)      # exact: values were rounded to fp32 first
    for k, v in feed.items():
        out["f/" + k] = v.numpy().astype(np.float32)
    st = O.new_state(p)
    rd, rg = O.iteration(p, st, feed, cfg)
    for k in ("G", "Dg_logit", "Dx_logit", "Dxmi_logit", "Dx_hat_logit", "grad_x_hat", "grad_cond", "slopes",
              "slopes2", "D_loss", "D_loss_real", "D_loss_fake", "D_loss_mismatch", "wdist", "wdist2", "reg_loss",
              "balance_loss", "real_gp", "real_gp2", "kt_grad"):
        out["o/" + k] = rd[k].numpy()
    out["o/kt"] = st["kt"].numpy()
    out["o/G_loss"] = rg["G_loss"].numpy()
    out["o/G_kl_loss"] = rg["G_kl_loss"].numpy()
    out["o/G_run_G"] = rg["G"].numpy()
    out["o/G_run_Dg_logit"] = rg["Dg_logit"].numpy()
    for k, v in list(rd["grads"].items()) + list(rg["grads"].items()):
        out["gn/" + k] = np.asarray(float(v.double().norm()))
    if store_params:
        for k, v in p.items():
            out["q/" + k] = v.numpy().astype(np.float32)
    return out


def main():
    cfg = O.OracleCfg(**TINY)
    p = boosted(O.init_params(cfg, 0, torch.float64), torch.float64)
    p = O.OrderedDict((k, v.float().double()) for k, v in p.items())          # fp32-representable
    feed = {k: v.float().double() for k, v in O.make_feed(cfg, 11, torch.float64).items()}
    np.savez_compressed(os.path.join(HERE, "wgancls_tiny.npz"), **run(cfg, p, feed, True))
    cfg = O.OracleCfg(batch_size=4)
    p = O.init_params(cfg, 0, torch.float32)
    p["d_net/dense/kernel"] = p["d_net/dense/kernel"] * 4.0     # activates the cond penalty as well
    feed = O.make_feed(cfg, 1234, torch.float32)
    out = run(cfg, p, feed, False)
    out = {k: v for k, v in out.items() if not k.startswith("f/")}
    np.savez_compressed(os.path.join(HERE, "wgancls_full_b4.npz"), **out)


if __name__ == "__main__":
    main()
