"""Conditional PGGAN throughput on one GPU: images/s of one D run + G run of a stage at the
reference's channel schedule, captured CUDA graphs, inputs resident.
python tools/bench_pggan.py [--stage 7] [--trans 0] [--batch 16]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2i_b200.models.pggan.pggan import PGGAN  # noqa: E402
from t2i_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", type=int, default=7)
    ap.add_argument("--trans", type=int, default=0)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-graphs", action="store_true", help="eager launches (ncu launch lists)")
    args = ap.parse_args()
    B, S = args.batch, 4 * 2 ** (args.stage - 1)
    m = PGGAN(B, 1000, "/tmp/pggan_w", "/tmp/pggan_r", None, "/tmp/pggan_s", "/tmp/pggan_l", args.stage, bool(args.trans),
              precision="bf16", sample_num=B, use_graphs=not args.no_graphs)
    m.initialize(0)
    eng = m._train_engine()
    gen = torch.Generator().manual_seed(1)
    eng.load_feed(x=torch.rand(B, S, S, 3, generator=gen) * 2 - 1, x_mismatch=torch.rand(B, S, S, 3, generator=gen) * 2 - 1,
                  cond=torch.randn(B, 1024, generator=gen), z=torch.randn(B, 128, generator=gen),
                  epsilon=torch.rand(B, generator=gen), tn_eps=torch.randn(B, 128, generator=gen).clamp_(-2, 2))
    for i in range(args.warmup):
        eng.d_step(0.1 * i)
        eng.g_step()
    torch.cuda.synchronize()
    l0 = _lib.launch_count() + eng.replayed_launches
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.steps):
        eng.d_step(0.3 + 0.01 * i)
        eng.g_step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    sc = eng.scalars_dict()
    print(json.dumps({"metric": "images/sec (D run + G run) %dx%d conditional PGGAN stage %d%s" % (S, S, args.stage, " (transition)" if args.trans else ""),
                      "value": B / (ms * 1e-3), "ms_per_step": ms, "batch": B, "dtype": "bf16",
                      "launch": "eager" if args.no_graphs else "CUDA graphs, collectives outside",
                      "kernels_per_step": (_lib.launch_count() + eng.replayed_launches - l0) / args.steps,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                      "finite": all(v == v for v in sc.values()), "scalars": sc}))


if __name__ == "__main__":
    main()
