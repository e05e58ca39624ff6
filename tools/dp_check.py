"""Multi-GPU equivalence check (run under torchrun, NCCL):  every rank runs its shard of one wgancls iteration in
parity mode (bf16x3) with sync_bn=True; rank 0 additionally runs the whole batch on one GPU; images, losses and
parameter gradients must agree to parity-mode rounding.  Also prints the per-replica-BN (default) deviation.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import model_cfg  # noqa: E402
from t2i_b200.models.wgancls.model import WGanCls  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def iteration(model, feed, tn_d, tn_g, sel):
    eng = model._train_engine()
    eng.load_feed(**{k: v[sel] for k, v in feed.items()}, tn_eps=tn_d[sel])
    eng.d_step(1e-4)
    img = eng.d["img"][:eng.B].clone()
    gd = {k: v.clone() for k, v in eng.get_grads_tf().items() if k.startswith("d_net/")}
    sc_d = dict(eng.scalars_dict())
    eng.load_feed(tn_eps=tn_g[sel])
    eng.g_step(1e-4)
    gg = {k: v.clone() for k, v in eng.get_grads_tf().items() if k.startswith("g_net/")}
    return img, gd, gg, sc_d, dict(eng.scalars_dict())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    gb = 32 * world
    b = gb // world
    gen = torch.Generator().manual_seed(5)
    feed = {"x": torch.rand(gb, 64, 64, 3, generator=gen) * 2 - 1, "x_mismatch": torch.rand(gb, 64, 64, 3, generator=gen) * 2 - 1,
            "cond": torch.randn(gb, 1024, generator=gen), "z": torch.randn(gb, 128, generator=gen),
            "epsilon": torch.rand(gb, 1, 1, 1, generator=gen)}
    tn_d = torch.randn(gb, 128, generator=gen).clamp_(-2, 2)
    tn_g = torch.randn(gb, 128, generator=gen).clamp_(-2, 2)
    sl = slice(rank * b, (rank + 1) * b)
    out = {}
    for name, sync in (("sync_bn", True), ("per_replica_bn", False)):
        m = WGanCls(model_cfg(b), precision="bf16x3", device=dev, distributed=True, use_graphs=False, sync_bn=sync)
        m.initialize(0)
        out[name] = iteration(m, feed, tn_d, tn_g, sl)
        del m
    if rank == 0:
        ref = WGanCls(model_cfg(gb), precision="bf16x3", device=dev, use_graphs=False)
        ref.initialize(0)
        r_img, r_gd, r_gg, r_sd, r_sg = iteration(ref, feed, tn_d, tn_g, slice(None))
        for name in out:
            img, gd, gg, sd, sg = out[name]
            worst_d = max((rel(gd[k], r_gd[k]), k) for k in r_gd if float(r_gd[k].abs().max()) > 0)
            skip = lambda k: (k.endswith("biases") and "Conv2d_transpose" not in k and "Conv_9" not in k) or k.endswith("dense_2/bias")
            worst_g = max((rel(gg[k], r_gg[k]), k) for k in r_gg if float(r_gg[k].abs().max()) > 1e-12 and not skip(k))
            print("[dp_check] %-15s world %d: G image rel-L2 %.2e | D_loss %.6f vs %.6f | G_loss %.6f vs %.6f | worst d-grad %.2e (%s) "
                  "| worst g-grad %.2e (%s)" % (name, world, rel(img, r_img[sl]), sd["D_loss"], r_sd["D_loss"], sg["G_loss"],
                                                r_sg["G_loss"], worst_d[0], worst_d[1], worst_g[0], worst_g[1]))
        img, gd, gg, sd, sg = out["sync_bn"]
        assert rel(img, r_img[sl]) < 1e-3 and abs(sd["D_loss"] - r_sd["D_loss"]) < 1e-3 * max(1, abs(r_sd["D_loss"]))
        assert abs(sg["G_loss"] - r_sg["G_loss"]) < 1e-3 * max(1, abs(r_sg["G_loss"]))
        print("[dp_check] OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
