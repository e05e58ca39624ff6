"""Times the captured CUDA graphs of one wgancls iteration one by one (events on the launch stream), next to the
algorithmic GEMM FLOPs each contains.  Development aid:  python tools/step_breakdown.py [--batch 256]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import model_cfg  # noqa: E402
from t2i_b200.models.wgancls.model import WGanCls  # noqa: E402

G_F, D_F = 969.478e6, 694.305e6      # forward MACs per image


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    B = args.batch
    dev = torch.device("cuda", 0)
    model = WGanCls(model_cfg(B), precision="bf16", device=dev, use_graphs=True)
    model.initialize(0)
    eng = model._train_engine()
    gen = torch.Generator().manual_seed(1)
    eng.load_feed(x=torch.rand(B, 64, 64, 3, generator=gen) * 2 - 1, x_mismatch=torch.rand(B, 64, 64, 3, generator=gen) * 2 - 1,
                  cond=torch.randn(B, 1024, generator=gen), z=torch.randn(B, 128, generator=gen),
                  epsilon=torch.rand(B, 1, 1, 1, generator=gen), tn_eps=torch.randn(B, 128, generator=gen).clamp_(-2, 2))
    for _ in range(4):
        eng.d_step(1e-4)
        eng.g_step(1e-4)
    torch.cuda.synchronize()
    flops = {"d_a1": 2 * G_F * B, "d_a": 2 * 8 * D_F * B, "d_a2": 2 * 5 * D_F * B, "d_b": 0.0, "d_c": 0.0, "g_a1": 2 * G_F * B,
             "g_a2": 2 * D_F * B, "g_a3": 2 * (2 * G_F + D_F) * B, "g_b": 0.0, "g_c": 0.0}
    total = 0.0
    # d_b / g_b: loss scalars, d_c / g_c: Adam; g_a2: d_net forward of the G run, g_a3: both backward passes
    # d_a: the D run up to its losses (4B forward, seeded backward, penalties), d_a2: tangent pass + weight gradients
    for name in ("d_a1", "d_a", "d_b", "d_a2", "d_c", "g_a1", "g_a2", "g_b", "g_a3", "g_c"):
        graph = eng._graphs[name]["graph"]
        ts = []
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        med = ts[len(ts) // 2]
        total += med
        print("%-5s %7.3f ms  launches %3d  %s" % (name, med, eng._graphs[name]["launches"],
                                                  "%.0f TFLOP/s" % (flops[name] / med / 1e9) if flops[name] else ""))
    print("sum of graphs %.3f ms (serial; the step overlaps d_b / d_c with g_a1)" % total)

    def timed(label, body):
        body()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()
        ts = []
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        print("%-44s %7.3f ms" % (label, ts[len(ts) // 2]))

    g, d = eng.g, eng.d
    timed("g_forward", lambda: eng.g_forward(g["z"], eng.feed["cond"], g["tn"], d["img"][:B], g["kl_scratch"]))
    timed("g_forward + moving statistics", lambda: eng.g_forward(g["z"], eng.feed["cond"], g["tn"], d["img"][:B], g["kl_scratch"],
                                                                  update_moving=True))
    K = eng.K
    real_conv, real_bn, real_wgrad, real_bnb = K.conv_gemm, K.bn_apply_train, K.wgrad_gemm, K.bn_bwd_fused
    gf = lambda: eng.g_forward(g["z"], eng.feed["cond"], g["tn"], d["img"][:B], g["kl_scratch"])
    K.conv_gemm = lambda *a, **k: None
    timed("g_forward without its GEMMs", gf)
    K.bn_apply_train = lambda *a, **k: None
    timed("g_forward without GEMMs and BatchNorm", gf)
    K.conv_gemm = real_conv
    timed("g_forward without BatchNorm", gf)
    K.bn_apply_train = real_bn
    gb = lambda: eng.g_backward(d["gx"])
    K.wgrad_gemm = lambda *a, **k: None
    timed("g_backward without weight gradients", gb)
    K.bn_bwd_fused = lambda *a, **k: None
    timed("g_backward without weight gradients, BatchNorm", gb)
    K.conv_gemm = lambda *a, **k: None
    timed("g_backward: the rest (c9 bwd, im2col, bn0, ca)", gb)
    K.conv_gemm, K.wgrad_gemm, K.bn_bwd_fused = real_conv, real_wgrad, real_bnb
    timed("_g_body_fwd (fresh capture)", eng._g_body_fwd)

    def replay_timed(label, graph):
        ts = []
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        print("%-44s %7.3f ms  (min %.3f max %.3f)" % (label, ts[len(ts) // 2], ts[0], ts[-1]))

    replay_timed("g_a1 (the step's graph) again", eng._graphs["g_a1"]["graph"])
    replay_timed("d_a1 (the step's graph) again", eng._graphs["d_a1"]["graph"])
    with torch.cuda.stream(eng.comm_stream):
        replay_timed("g_a1 replayed on the comm stream", eng._graphs["g_a1"]["graph"])
    timed("_d_body_gen (fresh capture)", eng._d_body_gen)
    timed("zero g + g_forward(sums tail) ", lambda: (eng.grad["g"].zero_(), eng.g_forward(g["z"], eng.feed["cond"], g["tn"], d["img"][:B], eng.sums["g"][1:2])))
    timed("g_forward + to_planes(cond)", lambda: (eng.g_forward(g["z"], eng.feed["cond"], g["tn"], d["img"][:B], g["kl_scratch"]),
                                                 eng.K.to_planes(eng.feed["cond"], d["cond"][:, :B])))
    timed("grad[g].zero_", lambda: eng.grad["g"].zero_())
    timed("grad[d].zero_", lambda: eng.grad["d"].zero_())
    timed("d_forward(B)", lambda: eng.d_forward(0, B))
    timed("d_forward(4B)", lambda: eng.d_forward(0, 4 * B))
    timed("d_backward(B)", lambda: eng.d_backward(0, B, d["gseed"], 0, B, False))
    timed("d_backward(4B) + bias statistics", lambda: eng.d_backward(0, 4 * B, d["seed"], 3 * B, B, True, bias_n=3 * B))
    timed("g_backward (side stream joins)", lambda: eng.g_backward(d["gx"]))
    timed("tangent d_forward(B) + merged wgrad(4B)", lambda: (eng.d_forward(3 * B, B, tangent=True, after=lambda buf: None),
                                                              [eng.d_wgrad_layer(l, 4 * B, 3 * B) for l in eng.D_WGRAD]))
    timed("merged wgrad(4B) alone", lambda: [eng.d_wgrad_layer(l, 4 * B, 3 * B) for l in eng.D_WGRAD])
    timed("adam d", lambda: eng._adam("d"))


if __name__ == "__main__":
    main()
