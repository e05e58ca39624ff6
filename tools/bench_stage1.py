"""StackGAN stage-I throughput on one GPU: images/s of one D run + G run at the reference
widths (GF 128, DF 64, Z 100), captured CUDA graphs, inputs resident.  python tools/bench_stage1.py [--batch 256]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2i_b200.models.stackgan.stageI.model import ConditionalGan  # noqa: E402
from t2i_b200.utils.config import config_from_yaml  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    cfg = config_from_yaml(os.path.join(ROOT, "text-to-image_b200", "models", "stackgan", "stageI", "cfg", "flowers.yml"))
    cfg.TRAIN.BATCH_SIZE = args.batch
    m = ConditionalGan(cfg, precision="bf16")
    m.initialize(0)
    eng = m._train_engine()
    B = args.batch
    gen = torch.Generator().manual_seed(1)
    eng.load_feed(x=torch.rand(B, 64, 64, 3, generator=gen) * 2 - 1, x_mismatch=torch.rand(B, 64, 64, 3, generator=gen) * 2 - 1,
                  cond=torch.randn(B, 1024, generator=gen), z=torch.randn(B, 100, generator=gen),
                  tn_eps=torch.randn(B, 128, generator=gen).clamp_(-2, 2))
    for _ in range(5):
        eng.d_step(2e-4)
        eng.g_step(2e-4)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        eng.d_step(2e-4)
        eng.g_step(2e-4)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    sc = eng.scalars_dict()
    print(json.dumps({"metric": "images/sec (D run + G run) 64x64 StackGAN stage-I", "value": B / (ms * 1e-3), "ms_per_step": ms,
                      "batch": B, "dtype": "bf16", "launch": "CUDA graphs (D run, G run), collectives outside", "finite": all(v == v for v in sc.values()),
                      "scalars": sc}))


if __name__ == "__main__":
    main()
