"""StackGAN stage-II throughput on one GPU: 256x256 images/s of one D run + G run at the
reference widths (GF 128, DF 64, Z 100; frozen stage-I generator inside the graph), captured CUDA graphs, inputs
resident.  python tools/bench_stage2.py [--batch 64]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2i_b200.models.stackgan.stageI.model import ConditionalGan as StageI  # noqa: E402
from t2i_b200.models.stackgan.stageII.model import ConditionalGan as StageII  # noqa: E402
from t2i_b200.utils.config import config_from_yaml  # noqa: E402
from t2i_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-graphs", action="store_true", help="eager launches (ncu launch lists)")
    args = ap.parse_args()
    mdir = os.path.join(ROOT, "text-to-image_b200", "models", "stackgan")
    c1 = config_from_yaml(os.path.join(mdir, "stageI", "cfg", "flowers.yml"))
    c2 = config_from_yaml(os.path.join(mdir, "stageII", "cfg", "flowers.yml"))
    B = args.batch
    c1.TRAIN.BATCH_SIZE = c2.TRAIN.BATCH_SIZE = B
    c1.TRAIN.SAMPLE_NUM = c2.TRAIN.SAMPLE_NUM = B
    s1 = StageI(c1, precision="bf16", use_graphs=not args.no_graphs)
    s1.initialize(0)
    m = StageII(s1, c2, use_graphs=not args.no_graphs)
    m.initialize(1)
    eng = m._train_engine()
    gen = torch.Generator().manual_seed(1)
    eng.load_feed(x=torch.rand(B, 256, 256, 3, generator=gen) * 2 - 1, x_mismatch=torch.rand(B, 256, 256, 3, generator=gen) * 2 - 1,
                  cond=torch.randn(B, 1024, generator=gen), z=torch.randn(B, 100, generator=gen),
                  tn_eps=torch.randn(B, 128, generator=gen).clamp_(-2, 2), tn_s1=torch.randn(B, 128, generator=gen).clamp_(-2, 2))
    for _ in range(args.warmup):
        eng.d_step(2e-4)
        eng.g_step(2e-4)
    torch.cuda.synchronize()
    l0 = _lib.launch_count() + eng.replayed_launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        eng.d_step(2e-4)
        eng.g_step(2e-4)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    sc = eng.scalars_dict()
    print(json.dumps({"metric": "images/sec (D run + G run) 256x256 StackGAN stage-II", "value": B / (ms * 1e-3), "ms_per_step": ms,
                      "batch": B, "dtype": "bf16", "launch": "CUDA graphs (D run, G run), collectives outside",
                      "kernels_per_step": (_lib.launch_count() + eng.replayed_launches - l0) / args.steps,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                      "finite": all(v == v for v in sc.values()), "scalars": sc}))


if __name__ == "__main__":
    main()
