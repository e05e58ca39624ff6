"""Where a data-parallel iteration spends its time (run under torchrun, NCCL): CUDA events around every graph replay and
every all-reduce of the wgancls step at the bench configuration, on the stream each one is enqueued on; rank 0 prints
start offset (from the iteration's first launch) and duration of each, averaged over the timed iterations.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_timeline.py
"""
import collections
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import model_cfg  # noqa: E402
from t2i_b200.models.wgancls.model import WGanCls  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = 256
    model = WGanCls(model_cfg(B), precision="bf16", device=dev, distributed=world > 1, use_graphs=True)
    model.initialize(0)
    eng = model._train_engine()
    gen = torch.Generator().manual_seed(1 + rank)
    eng.load_feed(x=torch.rand(B, 64, 64, 3, generator=gen) * 2 - 1, x_mismatch=torch.rand(B, 64, 64, 3, generator=gen) * 2 - 1,
                  cond=torch.randn(B, 1024, generator=gen), z=torch.randn(B, 128, generator=gen),
                  epsilon=torch.rand(B, 1, 1, 1, generator=gen), tn_eps=torch.randn(B, 128, generator=gen).clamp_(-2, 2))
    for _ in range(8):
        eng.d_step(1e-4)
        eng.g_step(1e-4)
    torch.cuda.synchronize()
    log = []
    real_run, real_ar = eng._run, eng.allreduce

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def run(name, body):
        a = ev()
        real_run(name, body)
        log.append((name, a, ev()))

    def allreduce(t):
        a = ev()
        real_ar(t)
        log.append(("allreduce %.1f MB" % (t.numel() * 4 / 1e6), a, ev()))

    eng._run = run
    if world > 1:
        eng.allreduce = allreduce
    iters, marks = 10, []
    for _ in range(iters):
        marks.append((len(log), ev()))
        eng.d_step(1e-4)
        eng.g_step(1e-4)
    end = ev()
    torch.cuda.synchronize()
    total = marks[0][1].elapsed_time(end) / iters
    agg = collections.OrderedDict()
    for i, (lo, m) in enumerate(marks):
        hi = marks[i + 1][0] if i + 1 < iters else len(log)
        seen = collections.Counter()
        for name, a, b in log[lo:hi]:
            seen[name] += 1
            key = name if seen[name] == 1 else "%s #%d" % (name, seen[name])
            st = agg.setdefault(key, [0.0, 0.0])
            st[0] += m.elapsed_time(a) / iters
            st[1] += a.elapsed_time(b) / iters
    if rank == 0:
        print("world %d: %.3f ms per iteration" % (world, total))
        print("%-22s %9s %9s %9s" % ("item", "start ms", "dur ms", "end ms"))
        for k, (s, d) in agg.items():
            print("%-22s %9.3f %9.3f %9.3f" % (k, s, d, s + d))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
