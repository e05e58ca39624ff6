#!/usr/bin/env python
"""bench.py -- images/sec of one wgancls iteration (D+GP run, then G run) at 64x64, per BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 256] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (N > 1)

A step is one trainer iteration (models/wgancls/trainer.py:97-102 of the reference) on one
synthetic batch of 256 images per GPU (BASELINE config 2; config 3 = 8 x 256, weak scaling), bf16
throughput mode.  `value`: inputs already resident in HBM.  `e2e`: the same iteration through the
reference-facing API (WGanCls.run with host feed dicts): pinned-host -> device copies of the feeds
and a device -> host read of D_loss / G_loss inside the timed region.  `roofline`: the dominant
kernel (conv_gemm_kernel, all of its launches in a step) timed with CUDA events on the launch stream.
`cpu_baseline` / `--impl reference`: the reference's TF-1.4 graph cannot run, so the
CPU arm is the oracle port (oracle/wgancls_oracle.py, PyTorch-CPU fp32, all host threads) on
BASELINE config 1 (batch 16).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "images/sec (G+D+GP step) 64x64 wgancls"

_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout when
    NCCL_DEBUG is set in the environment): keep a private handle of the real stdout for the result line and point file
    descriptor 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()
GFLOP_PER_IMAGE = 28.585          # 2 * (4 G_f + 15 D_f) MACs per image per iteration


def measured_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu launch list of this round's build
    (profiles/r02_traffic.json, written by tools/summarize_ncu.py layers; falls back to round 1's), or None"""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        with open(path) as f:
            t = json.load(f)
        t = t.get("by_kernel", t)
        if kernel in t:
            return t[kernel].get("dram_bytes_per_launch"), name
    return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"], "hbm": p["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def mark(self):
        """index of the next sample: rows before it were taken outside the timed region"""
        return len(self.rows)

    def finish(self, first=0, last=None):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        rows = self.rows[first:last]
        if len(rows) < 3:      # a very short timed region: fall back to the samples around it (still under load)
            rows = self.rows[max(0, first - 3):]
        self.rows = rows
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_cpu_throughput(batch, steps, warmup):
    """images/s of the oracle port (one D run + one G run per step) on the host cores."""
    import torch
    from oracle import wgancls_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.OracleCfg(batch_size=batch)
    p = O.init_params(cfg, 0, torch.float32)
    st = O.new_state(p)
    feed = O.make_feed(cfg, 1234, torch.float32)
    for _ in range(warmup):
        O.iteration(p, st, feed, cfg)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.iteration(p, st, feed, cfg)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def model_cfg(batch):
    from t2i_b200.utils.config import config_from_yaml
    cfg = config_from_yaml(os.path.join(ROOT, "text-to-image_b200", "models", "wgancls", "cfg", "flowers.yml"))
    cfg.TRAIN.BATCH_SIZE = batch
    return cfg


def workload_name(batch, world):
    return ("wgancls 64x64, batch=%d per GPU (BASELINE config %s), 1024-d random text embeds, GF=DF=128, one D+GP run then "
            "one G run per step (N_CRITIC=1), TF-form Adam, reference init" % (batch, "2" if world == 1 else "3-style weak scaling"))


def run_reference(args, rank):
    """--impl reference: the CPU arm.  Rank 0 alone works; other ranks exit 0."""
    if rank != 0:
        return
    batch = 16
    ips, spi = oracle_cpu_throughput(batch, args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": spi * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.batch, int(os.environ.get("WORLD_SIZE", "1"))),
                       "ran": "batch=16 per step (BASELINE config 1, the reference's CPU-runnable case), NOT the GPU arm's batch "
                              "of %d: images/s of a full iteration is what is compared" % args.batch,
                       "batch_per_step": batch,
                       "sample": "each step is one full iteration on a 16-image batch of that workload (BASELINE config 1: "
                                 "the reference's CPU-runnable case), executed by the CPU restatement of the reference's "
                                 "TF-1.4 graph (TF 1.4 is not installable, SURVEY.md 8c)"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": "%d full iterations (D run + G run) at batch 16, fp32, PyTorch-CPU, all host threads"
                                       % args.steps},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs 4 and 5 (the rows next to the hot path): the same line for PGGAN stage 7 (256x256) and
# StackGAN stage-II (256x256).  Device-resident value, e2e through the model's own run(fetches, feed_dict) with pinned
# host feeds, conv_gemm roofline from per-launch CUDA events; data-parallel over N GPUs exactly like wgancls.
WIDEN = {
    "pggan7": {"metric": "images/sec (D run + G run) 256x256 conditional PGGAN stage 7", "batch": 16,
               "gflop_per_image": None,
               "workload": "pggan stage 7 (256x256, stabilised), batch=%d per GPU (the reference's 2 x 8), reference channel schedule, "
                           "1024-d random text embeds, D run + G run per step (BASELINE config 4)"},
    "stackgan2": {"metric": "images/sec (D run + G run) 256x256 StackGAN stage-II", "batch": 64, "gflop_per_image": None,
                  "workload": "stackgan stage-II 256x256 on a frozen stage-I 64x64 generator, batch=%d per GPU, GF 128 / DF 64, "
                              "1024-d random text embeds, D run + G run per step (BASELINE config 5)"},
}


def run_widen(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from t2i_b200 import _lib, kernels
    spec = WIDEN[args.workload]
    B = args.batch if args.batch_given else spec["batch"]
    distributed = True if world > 1 else None
    gen = torch.Generator().manual_seed(1234 + rank)
    pin = lambda t: t.pin_memory()
    if args.workload == "pggan7":
        from t2i_b200.models.pggan.pggan import PGGAN
        m = PGGAN(B, 1000, "/tmp/pggan_w", "/tmp/pggan_r", None, "/tmp/pggan_s", "/tmp/pggan_l", 7, False, precision="bf16",
                  sample_num=B, use_graphs=not args.no_graphs, distributed=distributed)
        m.initialize(0)
        S, zd = 256, 128
        host = {"x": pin(torch.rand(B, S, S, 3, generator=gen) * 2 - 1), "xm": pin(torch.rand(B, S, S, 3, generator=gen) * 2 - 1),
                "cond": pin(torch.randn(B, 1024, generator=gen)), "z": pin(torch.randn(B, zd, generator=gen)),
                "eps": pin(torch.rand(B, 1, 1, 1, generator=gen))}
        fd = {m.x: host["x"], m.x_mismatch: host["xm"], m.cond: host["cond"], m.z: host["z"], m.epsilon: host["eps"],
              m.learning_rate: 2e-6, m.iter: 1}
        d_fetch, g_fetch = [m.D_optim, m.D_loss], [m.G_optim, m.G_loss]
        eng = m._train_engine()
        eng.load_feed(x=host["x"], x_mismatch=host["xm"], cond=host["cond"], z=host["z"], epsilon=host["eps"].reshape(-1),
                      tn_eps=torch.randn(B, 128, generator=gen).clamp_(-2, 2))
        step = lambda: (eng.d_step(0.5), eng.g_step())
        feed_keys = ("x", "xm", "cond", "z", "eps")
    else:
        from t2i_b200.models.stackgan.stageI.model import ConditionalGan as StageI
        from t2i_b200.models.stackgan.stageII.model import ConditionalGan as StageII
        from t2i_b200.utils.config import config_from_yaml
        mdir = os.path.join(ROOT, "text-to-image_b200", "models", "stackgan")
        c1 = config_from_yaml(os.path.join(mdir, "stageI", "cfg", "flowers.yml"))
        c2 = config_from_yaml(os.path.join(mdir, "stageII", "cfg", "flowers.yml"))
        c1.TRAIN.BATCH_SIZE = c2.TRAIN.BATCH_SIZE = B
        c1.TRAIN.SAMPLE_NUM = c2.TRAIN.SAMPLE_NUM = B
        s1 = StageI(c1, precision="bf16", use_graphs=not args.no_graphs, distributed=distributed)
        s1.initialize(0)
        m = StageII(s1, c2, use_graphs=not args.no_graphs, distributed=distributed)
        m.initialize(1)
        S, zd = 256, 100
        host = {"x": pin(torch.rand(B, S, S, 3, generator=gen) * 2 - 1), "xm": pin(torch.rand(B, S, S, 3, generator=gen) * 2 - 1),
                "cond": pin(torch.randn(B, 1024, generator=gen)), "z": pin(torch.randn(B, zd, generator=gen))}
        fd = {m.inputs: host["x"], m.wrong_inputs: host["xm"], m.embed_inputs: host["cond"], m.z: host["z"]}
        from t2i_b200.models.wgancls.model import Fetch       # the stage-II train ops belong to its trainer (trainer.py:51-57)
        d_fetch, g_fetch = [Fetch("D_optim", "op"), Fetch("D_loss", "scalar")], [Fetch("G_optim", "op"), Fetch("G_loss", "scalar")]
        eng = m._train_engine()
        eng.load_feed(x=host["x"], x_mismatch=host["xm"], cond=host["cond"], z=host["z"],
                      tn_eps=torch.randn(B, 128, generator=gen).clamp_(-2, 2), tn_s1=torch.randn(B, 128, generator=gen).clamp_(-2, 2))
        step = lambda: (eng.d_step(2e-4), eng.g_step(2e-4))
        feed_keys = ("x", "xm", "cond", "z")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if sampler is not None:
        t_end = time.time() + 1.0
        while sampler.mark() == 0 and time.time() < t_end:     # nvidia-smi needs ~0.3 s for its first row
            time.sleep(0.05)
    barrier()
    s_first = sampler.mark() if sampler is not None else 0
    l0 = _lib.launch_count() + eng.replayed_launches - eng.captured_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = _lib.launch_count() + eng.replayed_launches - eng.captured_launches - l0
    clocks = sampler.finish(s_first, sampler.mark()) if sampler is not None else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    value = B * world / (ms_per_step * 1e-3)

    def e2e_step():
        d_loss = m.run(d_fetch, fd)[1]
        g_loss = m.run(g_fetch, fd)[1]
        return d_loss, g_loss
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        losses = e2e_step()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = B * world * n_e2e / float(dt.item())
    # the D run stages every feed, the G run the conditioning and z only (it reads no real image)
    h2d = sum(host[k].numel() * 4 for k in feed_keys) + sum(host[k].numel() * 4 for k in ("cond", "z"))
    # conv_gemm roofline: eager, single stream, one event pair per launch
    pk = peaks()
    side, eng.side_stream = eng.side_stream, None
    comm, eng.comm_stream = getattr(eng, "comm_stream", None), None
    torch.cuda._sleep(int(2e7))          # the host enqueues the eager step while the device runs the spin kernel and two
    for _ in range(2):                   # replayed steps: its launches then run back to back at the sustained power state
        step()
    kernels.PROFILE = []
    step()
    barrier()
    prof, kernels.PROFILE = kernels.PROFILE, None
    eng.side_stream, eng.comm_stream = side, comm
    roofline = None
    if rank == 0:
        agg = {}
        for k, tag, fl, nb, a, b in prof:
            v = agg.setdefault(k, [0.0, 0.0, 0.0, 0])
            v[0] += fl; v[1] += nb; v[2] += a.elapsed_time(b); v[3] += 1
        top = max(agg, key=lambda k: agg[k][2])
        fl, nb, t, n = agg[top]
        roofline = {"kernel": top + "_kernel", "bound": "tensor", "achieved": fl / (t * 1e-3) / 1e12, "peak": pk["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": fl / (t * 1e-3) / 1e12 / pk["bf16_sustained"], "traffic": None,
                    "algorithmic_bytes_per_launch": nb / n, "launches_per_step": n, "ms_per_step_in_kernel": t,
                    "peak_source": pk["source"] + ", sustained",
                    "gemm_tflop_per_step": sum(v[0] for v in agg.values()) / 1e12,
                    "other_kernels": {k: {"ms_per_step": v[2], "tflops": v[0] / (v[2] * 1e-3) / 1e12} for k, v in agg.items() if k != top}}
        emit({"metric": spec["metric"], "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
              "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
              "config": {"workload": spec["workload"] % B, "global_batch": B * world,
                         "parallelism": "dp%d (batch shards, one NCCL allreduce per optimizer step)" % world,
                         "l2": "per-step working set exceeds the 126 MB L2; no explicit flush"},
              "roofline": roofline,
              "cpu_baseline": None, "cpu_baseline_note": "the CPU arm (bench.py --impl reference) exists for the north-star workload only",
              "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 2 * 16 * 4},
              "gpu_launches": int(launches), "clocks": clocks, "finite": bool(all(np.isfinite(v) for v in losses)),
              "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30})
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default: 256; pggan7 16; stackgan2 64)")
    ap.add_argument("--workload", default="wgancls", choices=["wgancls", "pggan7", "stackgan2"],
                    help="wgancls = the north-star workload (BASELINE config 2 / 3); the others are the rows next to it")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-mode", action="store_true", help="skip the bf16x3 sub-record")
    ap.add_argument("--only-resident", action="store_true", help="profiling aid: skip e2e / roofline / cpu legs")
    ap.add_argument("--no-graphs", action="store_true", help="profiling aid: eager launches (ncu launch lists)")
    ap.add_argument("--profile-out", default=None, help="write the per-launch GEMM table (JSON) here")
    args = ap.parse_args()
    args.batch_given = args.batch is not None
    if args.batch is None:
        args.batch = 256
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.workload != "wgancls":
        return run_widen(args, rank, world, local_rank)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from t2i_b200 import _lib, kernels
    from t2i_b200.models.wgancls.model import WGanCls

    B = args.batch
    model = WGanCls(model_cfg(B), precision=args.precision, device=dev, distributed=True if world > 1 else None,
                    use_graphs=not args.no_graphs)
    model.initialize(0)                       # reference init, identical on all ranks
    eng = model._train_engine()
    gen = torch.Generator().manual_seed(1234 + rank)
    host = {
        "x": (torch.rand(B, 64, 64, 3, generator=gen) * 2 - 1).pin_memory(),
        "x_mismatch": (torch.rand(B, 64, 64, 3, generator=gen) * 2 - 1).pin_memory(),
        "cond": torch.randn(B, 1024, generator=gen).pin_memory(),
        "z": torch.randn(B, 128, generator=gen).pin_memory(),
        "epsilon": torch.rand(B, 1, 1, 1, generator=gen).pin_memory(),
    }
    model.seed_noise(99 + rank)
    lr = 1e-4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------ device-resident timing (`value`)
    eng.load_feed(x=host["x"], x_mismatch=host["x_mismatch"], cond=host["cond"], z=host["z"], epsilon=host["epsilon"])
    tn = [torch.empty(B, 128, device=dev).normal_().clamp_(-2, 2) for _ in range(2)]

    def resident_step():
        eng.g["tn"].copy_(tn[0])
        eng.d_step(lr)
        eng.g["tn"].copy_(tn[1])
        eng.g_step(lr)

    # rank 0 samples its GPU's clocks (one nvidia-smi poller per rank would put N pollers on the host cores the ranks'
    # own launch threads need)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.start()                  # started before the warm-up: nvidia-smi needs ~0.3 s to deliver its first row
    for _ in range(args.warmup):
        resident_step()
    barrier()
    live = torch.tensor([1 if (sampler is None or sampler.mark() > 0) else 0], device=dev)
    for _ in range(30):                  # keep the GPUs loaded until the sampler is live (all ranks take the same trips)
        if world > 1:
            dist.all_reduce(live, op=dist.ReduceOp.MIN)
        if int(live.item()) == 1:
            break
        resident_step()
        live.fill_(1 if (sampler is None or sampler.mark() > 0) else 0)
    barrier()
    s_first = sampler.mark() if sampler is not None else 0
    l0 = _lib.launch_count() + eng.replayed_launches - eng.captured_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        resident_step()
    e1.record()
    barrier()
    launches = _lib.launch_count() + eng.replayed_launches - eng.captured_launches - l0
    clocks = sampler.finish(s_first, sampler.mark()) if sampler is not None else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    value = B * world / (ms_per_step * 1e-3)
    finite = all(v == v for v in eng.scalars_dict().values())

    if args.only_resident:
        if rank == 0:
            emit({"metric": METRIC, "value": value, "unit": "images/s", "ms_per_step": ms_per_step,
                  "gpu_launches": int(launches), "note": "--only-resident (profiling aid)"})
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------ end-to-end through the public API
    def e2e_step():
        fd = {model.x: host["x"], model.x_mismatch: host["x_mismatch"], model.cond: host["cond"], model.z: host["z"],
              model.epsilon: host["epsilon"], model.learning_rate_d: lr, model.learning_rate_g: lr}
        d_loss = model.run([model.D_optim, model.kt_optim, model.D_loss], fd)[2]
        g_loss = model.run([model.G_optim, model.G_loss], fd)[1]
        return d_loss, g_loss

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / float(dt.item())
    h2d = sum(host[k].numel() * 4 for k in ("x", "x_mismatch", "cond", "z", "epsilon")) + \
        sum(host[k].numel() * 4 for k in ("cond", "z"))       # the G run re-feeds cond and z
    d2h = 2 * 16 * 4                                           # the scalars vector, read once per run

    # ------------------------------------------------------------ per-launch GEMM timing (roofline)
    pk = peaks()
    roofline = None
    # every rank runs these steps (they contain the allreduce); they run eagerly, outside the CUDA graphs
    # and single-stream (side and communication streams off), so that each kernel's event pair times that kernel alone.
    # Ahead of each profiled step the device gets a spin kernel and three replayed (graph) steps: while they run, the
    # host enqueues the eager step's ~300 launches, which then execute back to back (an eager launch on an idle device
    # has its launch latency inside the event pair) and at the SUSTAINED power state of the replayed steps (after an
    # idle spin alone the first launches would run at boost clocks and flatter the fraction of the sustained peak).
    side, eng.side_stream = eng.side_stream, None
    comm, eng.comm_stream = eng.comm_stream, None

    def profiled_step():
        prof_list, tl_list = kernels.PROFILE, _lib.TIMELINE
        kernels.PROFILE, _lib.TIMELINE = None, None
        torch.cuda._sleep(int(2e7))
        for _ in range(3):
            resident_step()
        kernels.PROFILE, _lib.TIMELINE = prof_list, tl_list
        resident_step()

    kernels.PROFILE = []
    psteps = 2
    for _ in range(psteps):
        profiled_step()
    barrier()
    prof = kernels.PROFILE
    timeline = None
    if args.profile_out:     # a second pair of steps with an event pair around EVERY C-ABI call
        kernels.PROFILE = []
        _lib.TIMELINE = []
        for _ in range(psteps):
            profiled_step()
        barrier()
        timeline, _lib.TIMELINE = _lib.TIMELINE, None
    kernels.PROFILE = None
    eng.side_stream, eng.comm_stream = side, comm
    if rank == 0:
        rows = [(k, tag, fl, nb, a.elapsed_time(b)) for k, tag, fl, nb, a, b in prof]
        agg = {}
        for k, tag, fl, nb, t in rows:
            a = agg.setdefault(k, [0.0, 0.0, 0.0, 0])
            a[0] += fl; a[1] += nb; a[2] += t; a[3] += 1
        top = max(agg, key=lambda k: agg[k][2])
        fl, nb, t, n = agg[top]
        achieved = fl / (t * 1e-3) / 1e12
        roofline = {"kernel": top + "_kernel", "bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                    "traffic": measured_traffic(top + "_kernel")[0], "traffic_unit": "dram bytes per launch (ncu, mean over the "
                    "launches of one step; profiles/%s)" % measured_traffic(top + "_kernel")[1], "algorithmic_bytes_per_launch": nb / n,
                    "peak_source": pk["source"] + ", sustained (kernel timed inside a long step)",
                    "launches_per_step": n // psteps, "ms_per_step_in_kernel": t / psteps,
                    "share_of_step": (t / psteps) / ms_per_step,
                    "whole_step_tflops": GFLOP_PER_IMAGE * 1e9 * value / world / 1e12,
                    "whole_step_frac_of_sustained": GFLOP_PER_IMAGE * 1e9 * value / world / 1e12 / pk["bf16_sustained"],
                    "other_kernels": {k: {"ms_per_step": v[2] / psteps, "tflops": v[0] / (v[2] * 1e-3) / 1e12}
                                      for k, v in agg.items() if k != top}}
        if args.profile_out:
            per = {}
            for k, tag, fl, nb, t in rows:
                a = per.setdefault(k + " " + tag, [0.0, 0.0, 0.0, 0])
                a[0] += fl; a[1] += nb; a[2] += t; a[3] += 1
            table = [{"launch": k, "calls_per_step": v[3] / psteps, "ms_per_step": v[2] / psteps,
                      "tflops": v[0] / (v[2] * 1e-3) / 1e12, "algo_gbytes_per_s": v[1] / (v[2] * 1e-3) / 1e9}
                     for k, v in sorted(per.items(), key=lambda kv: -kv[1][2])]
            ep = {}
            for name, a, b in timeline:
                v = ep.setdefault(name, [0.0, 0])
                v[0] += a.elapsed_time(b); v[1] += 1
            entry = [{"entry_point": k, "calls_per_step": v[1] / psteps, "ms_per_step": v[0] / psteps}
                     for k, v in sorted(ep.items(), key=lambda kv: -kv[1][0])]
            with open(args.profile_out, "w") as f:
                json.dump({"ms_per_step": ms_per_step, "table": table, "entry_points": entry}, f, indent=1)

    # ------------------------------------------------------------ parity mode (the arithmetic that meets 1e-3)
    # The same iteration in precision="bf16x3" (split bf16: three tensor-core products per contraction, two planes of
    # every activation), device-resident, timed like `value`: accuracy and speed on one line.
    parity_mode = None
    if not args.no_parity_mode and args.precision == "bf16":
        pm = WGanCls(model_cfg(B), precision="bf16x3", device=dev, distributed=True if world > 1 else None,
                     use_graphs=not args.no_graphs)
        pm.initialize(0)
        pe = pm._train_engine()
        pe.load_feed(x=host["x"], x_mismatch=host["x_mismatch"], cond=host["cond"], z=host["z"], epsilon=host["epsilon"])

        def parity_step():
            pe.g["tn"].copy_(tn[0]); pe.d_step(lr)
            pe.g["tn"].copy_(tn[1]); pe.g_step(lr)

        for _ in range(3):
            parity_step()
        barrier()
        psn = max(3, min(args.steps, 10))
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(psn):
            parity_step()
        p1.record()
        barrier()
        pms = torch.tensor([p0.elapsed_time(p1)], device=dev)
        if world > 1:
            dist.all_reduce(pms, op=dist.ReduceOp.MAX)
        pms = float(pms.item()) / psn
        pfinite = all(v == v for v in pe.scalars_dict().values())
        parity_mode = {"precision": "bf16x3", "value": B * world / (pms * 1e-3), "unit": "images/s", "ms_per_step": pms,
                       "steps": psn, "finite": bool(pfinite),
                       "tensor_tflops_executed": 3 * GFLOP_PER_IMAGE * 1e9 * B / (pms * 1e-3) / 1e12,
                       "forward_rel_l2_bar": 1e-3,
                       "forward_rel_l2_evidence": "tests/test_f1_gpu.py::test_bench_config_forward_vs_oracle_batch256[bf16x3] "
                                                  "(this configuration against the CPU oracle; measured ~5e-5 G / ~1e-4 D)"}
        del pm, pe
        torch.cuda.empty_cache()

    # ------------------------------------------------------------ CPU baseline (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ips, spi = oracle_cpu_throughput(16, 3, 1)
        cpu = {"value": ips, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "3 full iterations (D run + G run) at batch 16 (BASELINE config 1), fp32, PyTorch-CPU oracle port, "
                         "all host threads; %.2f s per iteration" % spi}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16 (split x3, parity mode)",
                "data": "synthetic",
                "config": {"workload": workload_name(B, world),
                           "global_batch": B * world, "parallelism": "dp%d (batch shards, one NCCL allreduce per optimizer step)" % world,
                           "l2": "per-step working set (activations + gradients > 2 GB) exceeds the 126 MB L2; no explicit flush",
                           "bn": "per-replica batch statistics" if world > 1 else "single replica"},
                "roofline": roofline, "cpu_baseline": cpu,
                "parity": {"mode": args.precision, "forward_rel_l2_tolerance": {"G": 3e-2, "D": 1.5e-2} if args.precision == "bf16"
                           else {"G": 1e-3, "D": 1e-3},
                           "evidence": "tests/test_f1_gpu.py::test_bench_config_forward_vs_oracle_batch256 (batch 256, this "
                                       "configuration, against the CPU oracle); error budget: tools/bf16_error_budget.py, "
                                       "DESIGN.md section 5"},
                "parity_mode": parity_mode,
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches), "clocks": clocks, "finite": bool(finite)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
