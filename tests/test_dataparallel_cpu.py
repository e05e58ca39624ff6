"""world_size-2 gloo test of the data-parallel host logic: each rank runs the engine (CPU restatement
of the kernels, exact fp64 storage) on its shard of a global batch; ONE allreduce of the flat
gradient buffer (with the scalar sums in its tail) per optimizer step must reproduce the
single-process D run on the whole batch exactly (d_net has no BatchNorm)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    """a TCP port that is free right now (a fixed pid-derived port can still be in TIME_WAIT from an earlier run)"""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker_buckets(rank, world, port, out):
    _worker(rank, world, port, out, g_buckets=2)


def _worker(rank, world, port, out, g_buckets=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels as fk
    from oracle import wgancls_oracle as O
    from t2i_b200.engine import Engine
    from test_engine_cpu import TINY, boosted_params
    torch.set_num_threads(1)
    cfg = O.OracleCfg(**TINY)
    gb = cfg.batch_size
    b = gb // world
    p = boosted_params(cfg)
    feed = O.make_feed(cfg, 11, torch.float64)
    calls = []

    def allreduce(t):
        calls.append(t.numel())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    eng = Engine(fk, "cpu", b, 1, cfg.z_dim, cfg.embed_dim, cfg.compressed_embed_dim, cfg.gf_dim, cfg.df_dim,
                 cfg.beta1, cfg.beta2, cfg.kl_coeff, world, allreduce, act_dtype=torch.float64, f32_dtype=torch.float64,
                 g_buckets=g_buckets)
    eng.set_params_tf(p)
    sl = slice(rank * b, (rank + 1) * b)
    eng.load_feed(**{k: feed[k][sl] for k in ("x", "x_mismatch", "cond", "z", "epsilon")}, tn_eps=feed["tn_eps"][sl])
    # the generator's BatchNorm is per replica under data parallelism; feed the D run the images of the
    # single-process generator so that the comparison isolates the sharded D math
    ref = Engine(fk, "cpu", gb, 1, cfg.z_dim, cfg.embed_dim, cfg.compressed_embed_dim, cfg.gf_dim, cfg.df_dim,
                 cfg.beta1, cfg.beta2, cfg.kl_coeff, act_dtype=torch.float64, f32_dtype=torch.float64)
    ref.set_params_tf(p)
    ref.load_feed(**{k: feed[k] for k in ("x", "x_mismatch", "cond", "z", "epsilon")}, tn_eps=feed["tn_eps"])
    ref.d_step(cfg.d_lr)
    orig = eng.g_forward

    def g_forward_fixed(z, cond, tn, img_out, kl, **kw):
        orig(z, cond, tn, img_out, kl, **kw)
        img_out.copy_(ref.d["img"][sl])
    eng.g_forward = g_forward_fixed
    eng.d_step(cfg.d_lr)
    eng.g_forward = orig
    g_sharded, g_full = eng.get_grads_tf(), ref.get_grads_tf()
    worst = max(float((g_sharded[n] - g_full[n]).abs().max() / (g_full[n].abs().max() + 1e-30))
                for n in g_full if n.startswith("d_net/"))
    sc, scr = eng.scalars_dict(), ref.scalars_dict()
    smax = max(abs(sc[k] - scr[k]) / max(1.0, abs(scr[k])) for k in sc)
    eng.g_step(cfg.g_lr)      # runs; per-replica BN, gradients identical on all ranks after the allreduce
    flat = eng.grad["g"].clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    # the G run's gradients against the single-process engine fed the same generator statistics are covered by the
    # sync_bn test; here: the bucketed all-reduce (g_buckets = 2) must produce the same flat gradient as one call
    if rank == 0:
        torch.save({"worst": worst, "smax": smax, "calls": calls, "same": same, "g_flat": flat, "g_n": eng.g_n,
                    "g_split": eng.g_split, "kt": float(eng.kt), "kt_ref": float(ref.kt),
                    "p_flat": eng.flat["g"].clone(), "m_flat": eng.adam_m["g"].clone(), "v_flat": eng.adam_v["g"].clone()}, out)
    dist.destroy_process_group()


def _worker_sync_bn(rank, world, port, out):
    """whole iteration with sync_bn=True on 2 ranks == the single-process iteration on the global batch"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels as fk
    from oracle import wgancls_oracle as O
    from t2i_b200.engine import Engine
    from test_engine_cpu import TINY, boosted_params
    torch.set_num_threads(1)
    cfg = O.OracleCfg(**TINY)
    gb = cfg.batch_size
    b = gb // world
    p = boosted_params(cfg)
    feed = O.make_feed(cfg, 12, torch.float64)
    calls = []

    def allreduce(t):
        calls.append(t.numel())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    exact = dict(act_dtype=torch.float64, f32_dtype=torch.float64)
    eng = Engine(fk, "cpu", b, 1, cfg.z_dim, cfg.embed_dim, cfg.compressed_embed_dim, cfg.gf_dim, cfg.df_dim,
                 cfg.beta1, cfg.beta2, cfg.kl_coeff, world, allreduce, sync_bn=True, **exact)
    ref = Engine(fk, "cpu", gb, 1, cfg.z_dim, cfg.embed_dim, cfg.compressed_embed_dim, cfg.gf_dim, cfg.df_dim,
                 cfg.beta1, cfg.beta2, cfg.kl_coeff, **exact)
    sl = slice(rank * b, (rank + 1) * b)
    keys = ("x", "x_mismatch", "cond", "z", "epsilon")
    res = {}
    for e, sel in ((eng, sl), (ref, slice(None))):
        e.set_params_tf(p)
        e.load_feed(**{k: feed[k][sel] for k in keys}, tn_eps=feed["tn_eps"][sel])
        e.d_step(cfg.d_lr)
    res["img"] = float((eng.d["img"][:b] - ref.d["img"][:gb][sl]).abs().max())
    gs, gf = eng.get_grads_tf(), ref.get_grads_tf()
    res["d_grads"] = max(float((gs[n] - gf[n]).abs().max() / (gf[n].abs().max() + 1e-30)) for n in gf if n.startswith("d_net/"))
    for e, sel in ((eng, sl), (ref, slice(None))):
        e.load_feed(tn_eps=feed["tn_eps_g"][sel])
        e.g_step(cfg.g_lr)
    gs, gf = eng.get_grads_tf(), ref.get_grads_tf()
    res["g_grads"] = max((float((gs[n] - gf[n]).abs().max() / (gf[n].abs().max() + 1e-30)), n)
                         for n in gf if n.startswith("g_net/") and float(gf[n].abs().max()) > 1e-12
                         and not (n.endswith("biases") and "Conv2d_transpose" not in n and "Conv_9" not in n)
                         and not n.endswith("dense_2/bias"))     # biases in front of a BatchNorm: exactly-zero gradient
    sc, scr = eng.scalars_dict(), ref.scalars_dict()
    res["scalars"] = max(abs(sc[k] - scr[k]) / max(1.0, abs(scr[k])) for k in sc)
    ps, pf = eng.get_params_tf(), ref.get_params_tf()
    res["moving"] = max(float((ps[n] - pf[n]).abs().max()) for n in pf if "moving" in n)
    res["calls"] = calls
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_two_rank_sync_bn_iteration_equals_single_process(tmp_path):
    """with synchronised BatchNorm sums the sharded iteration IS the reference's whole-batch
    iteration (exact storage, so any difference is a logic error)."""
    out = str(tmp_path / "s.pt")
    port = _free_port()
    mp.spawn(_worker_sync_bn, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["img"] < 1e-10, r
    assert r["d_grads"] < 1e-8, r
    assert r["g_grads"][0] < 1e-7, r
    assert r["scalars"] < 1e-9, r
    assert r["moving"] < 1e-10, r
    # D run: 10 BatchNorm layers in the generator forward + the loss sums (sent ahead, the losses are published before
    # the tangent pass) + 1 gradient all-reduce;
    # G run: 10 forward + 10 backward BatchNorm all-reduces + 1 gradient all-reduce
    assert len(r["calls"]) == 12 + 21, r["calls"]


def _worker_stage1(rank, world, port, out):
    """StackGAN stage-I under data parallelism: one gradient all-reduce per optimizer step, replicas stay identical"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels as fk
    from oracle import stackgan1_oracle as S
    from t2i_b200.engine_stage1 import StageIEngine
    from test_stackgan1_cpu import TINY, boosted_params
    torch.set_num_threads(1)
    cfg = S.Stage1Cfg(**TINY)
    b = cfg.batch_size // world
    p = boosted_params(cfg)
    feed = S.make_feed(cfg, 21, torch.float64)
    calls = []

    def allreduce(t):
        calls.append(t.numel())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    eng = StageIEngine(fk, "cpu", b, 1, cfg.z_dim, cfg.embed_dim, cfg.compressed_embed_dim, cfg.gf_dim, cfg.df_dim,
                       cfg.d_beta1, cfg.g_beta1, cfg.alpha_mismatch, cfg.kl_coeff, world, allreduce,
                       act_dtype=torch.float64, f32_dtype=torch.float64)
    eng.set_params_tf(p)
    sl = slice(rank * b, (rank + 1) * b)
    eng.load_feed(x=feed["x"][sl], x_mismatch=feed["x_mismatch"][sl], cond=feed["cond"][sl], z=feed["z"][sl],
                  tn_eps=feed["tn_eps"][sl])
    eng.d_step(cfg.lr)
    eng.load_feed(tn_eps=feed["tn_eps_g"][sl])
    eng.g_step(cfg.lr)
    flat = torch.cat([eng.flat["d"], eng.flat["g"]])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    sc = eng.scalars_dict()
    if rank == 0:
        torch.save({"calls": calls, "same": all(torch.equal(gathered[0], t) for t in gathered),
                    "finite": all(v == v for v in sc.values()), "d_n": eng.d_n + 8, "g_n": eng.g_n + 8}, out)
    dist.destroy_process_group()


def test_two_rank_stage1_iteration(tmp_path):
    out = str(tmp_path / "s1.pt")
    port = _free_port()
    mp.spawn(_worker_stage1, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["calls"] == [r["d_n"], r["g_n"]], r        # exactly one all-reduce per optimizer step
    assert r["same"] and r["finite"], r                  # replicas hold identical parameters afterwards


def test_two_rank_d_run_equals_single_process(tmp_path):
    out = str(tmp_path / "r.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["worst"] < 1e-9, r
    assert r["smax"] < 1e-9, r
    assert abs(r["kt"] - r["kt_ref"]) < 1e-12
    # one gradient all-reduce per optimizer step; each run sends its handful of loss sums ahead of it (the losses are
    # published before the rest of the backward work)
    assert len(r["calls"]) == 4 and r["calls"][0] <= 16 and r["calls"][2] <= 16 and r["same"], r


def test_two_rank_bucketed_g_allreduce_matches_single_call(tmp_path):
    """g_buckets = 2 (the CUDA default at world > 1): [t0 .. c9] goes out when the backward pass reaches the 4x4 maps,
    the rest afterwards; same reduced gradient, element for element, as the single all-reduce"""
    outs = []
    for worker in (_worker, _worker_buckets):
        out = str(tmp_path / ("r%d.pt" % len(outs)))
        mp.spawn(worker, args=(2, _free_port(), out), nprocs=2, join=True)
        outs.append(torch.load(out))
    one, two = outs
    # calls: [D loss sums, D gradient, G loss sums, G gradient ...] (the sums go ahead: losses are published early)
    assert len(one["calls"]) == 4 and len(two["calls"]) == 5, (one["calls"], two["calls"])
    assert one["calls"][2] == two["calls"][2] <= 16 and one["calls"][0] <= 16
    assert two["calls"][3] == two["g_split"] and two["calls"][3] + two["calls"][4] == one["calls"][3]
    assert 0.5 < two["g_split"] / two["g_n"] < 0.8          # the early bucket carries most of the bytes
    assert two["same"] and torch.equal(one["g_flat"], two["g_flat"])
    # ... and the per-bucket Adam steps (the first one right behind its all-reduce) leave the same parameters and slots
    for k in ("p_flat", "m_flat", "v_flat"):
        assert torch.equal(one[k], two[k]), k


def _worker_pggan(rank, world, port, out):
    """conditional PGGAN (transition stage) on 2 ranks == the single-process iteration on the global batch: layer_norm is
    per sample and d_net has no normalisation, so ONE gradient all-reduce per optimizer step is exact"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels as fk
    from oracle import pggan_oracle as P
    from t2i_b200.engine_pggan import PgganEngine
    from test_pggan_cpu import TINY, boosted_params
    torch.set_num_threads(1)
    cfg = P.PgganCfg(stage=2, trans=True, **dict(TINY, batch_size=4))
    gb = cfg.batch_size
    b = gb // world
    p = boosted_params(cfg)
    feed = P.make_feed(cfg, 12, torch.float64)
    calls = []

    def allreduce(t):
        calls.append(t.numel())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    exact = dict(act_dtype=torch.float64, f32_dtype=torch.float64)
    args = (cfg.stage, cfg.trans, cfg.z_dim, cfg.embed_dim, cfg.compr_embed_dim, cfg.nf_base, cfg.nf_cap, cfg.d_embed, cfg.rgb_mid)
    eng = PgganEngine(fk, "cpu", b, 1, *args, world, allreduce, **exact)
    ref = PgganEngine(fk, "cpu", gb, 1, *args, **exact)
    sl = slice(rank * b, (rank + 1) * b)
    keys = ("x", "x_mismatch", "cond", "z", "epsilon")
    res = {}
    for e, sel in ((eng, sl), (ref, slice(None))):
        e.set_params_tf(p)
        e.load_feed(**{k: feed[k][sel] for k in keys}, tn_eps=feed["tn_eps"][sel])
        e.d_step(0.4)
    res["img"] = float((eng.d["img"][:b] - ref.d["img"][:gb][sl]).abs().max())
    gs, gf = eng.get_grads_tf(), ref.get_grads_tf()
    res["d_grads"] = max(float((gs[n] - gf[n]).abs().max() / (gf[n].abs().max() + 1e-30)) for n in gf if n.startswith("d_net/"))
    sc, scr = eng.scalars_dict(), ref.scalars_dict()
    res["gp"] = scr["real_gp"]
    for e, sel in ((eng, sl), (ref, slice(None))):
        e.load_feed(tn_eps=feed["tn_eps_g"][sel])
        e.g_step()
    gs, gf = eng.get_grads_tf(), ref.get_grads_tf()
    res["g_grads"] = max(float((gs[n] - gf[n]).abs().max() / (gf[n].abs().max() + 1e-30)) for n in gf if n.startswith("g_net/"))
    sc, scr = eng.scalars_dict(), ref.scalars_dict()
    res["scalars"] = max(abs(sc[k] - scr[k]) / max(1.0, abs(scr[k])) for k in sc)
    ps, pf = eng.get_params_tf(), ref.get_params_tf()
    res["params"] = max(float((ps[n] - pf[n]).abs().max()) for n in pf)
    res["calls"], res["d_n"], res["g_n"] = calls, eng.d_n + 8, eng.g_n + 8
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_two_rank_pggan_iteration_equals_single_process(tmp_path):
    out = str(tmp_path / "p.pt")
    port = _free_port()
    mp.spawn(_worker_pggan, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["gp"] > 1e-3, r                                # the penalty (and its second-order term) is active
    assert r["img"] < 1e-10 and r["d_grads"] < 1e-8 and r["g_grads"] < 1e-8, r
    assert r["scalars"] < 1e-9 and r["params"] < 1e-10, r
    # one gradient all-reduce per optimizer step, the G run's loss sums ahead of its gradient
    assert len(r["calls"]) == 4 and r["calls"][0] <= 16 and r["calls"][2] <= 16, r
    assert r["calls"][0] + r["calls"][1] == r["d_n"] and r["calls"][2] + r["calls"][3] == r["g_n"], r


def _worker_stage2(rank, world, port, out):
    """StackGAN stage-II under data parallelism: one gradient all-reduce per optimizer step, replicas stay identical,
    the frozen stage-I generator is never touched"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels as fk
    from oracle import stackgan2_oracle as S2
    from t2i_b200.engine_stage2 import StageIIEngine
    from test_stackgan2_cpu import TINY, boosted_params
    torch.set_num_threads(1)
    cfg = S2.Stage2Cfg(**dict(TINY, batch_size=4))
    b = cfg.batch_size // world
    p = boosted_params(cfg)
    feed = S2.make_feed(cfg, 21, torch.float64)
    calls = []

    def allreduce(t):
        calls.append(t.numel())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    eng = StageIIEngine(fk, "cpu", b, 1, cfg.z_dim, cfg.embed_dim, cfg.compressed_embed_dim, cfg.gf_dim, cfg.df_dim,
                        cfg.d_beta1, cfg.g_beta1, cfg.alpha_mismatch, cfg.kl_coeff, world, allreduce, s1_gf=cfg.s1_gf_dim,
                        act_dtype=torch.float64, f32_dtype=torch.float64)
    eng.set_params_tf(p)
    sl = slice(rank * b, (rank + 1) * b)
    eng.load_feed(x=feed["x"][sl], x_mismatch=feed["x_mismatch"][sl], cond=feed["cond"][sl], z=feed["z"][sl],
                  tn_eps=feed["tn_eps"][sl], tn_s1=feed["tn_s1"][sl])
    eng.d_step(cfg.lr)
    eng.load_feed(tn_eps=feed["tn_eps_g"][sl], tn_s1=feed["tn_s1_g"][sl])
    eng.g_step(cfg.lr)
    flat = torch.cat([eng.flat["d"], eng.flat["g"]])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    sc = eng.scalars_dict()
    q = eng.get_params_tf()
    frozen = all(torch.equal(q[n].double(), p[n]) for n in p if n.startswith("g_net/") and "moving" not in n)
    if rank == 0:
        torch.save({"calls": calls, "same": all(torch.equal(gathered[0], t) for t in gathered), "frozen": frozen,
                    "finite": all(v == v for v in sc.values()), "d_n": eng.d_n + 8, "g_n": eng.g_n + 8}, out)
    dist.destroy_process_group()


def test_two_rank_stage2_iteration(tmp_path):
    out = str(tmp_path / "s2.pt")
    port = _free_port()
    mp.spawn(_worker_stage2, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["calls"] == [r["d_n"], r["g_n"]], r        # exactly one all-reduce per optimizer step
    assert r["same"] and r["finite"] and r["frozen"], r


def _worker_pggan_mirror(rank, world, port, out):
    """the reference-facing PGGAN object with distributed=True (gloo): every rank feeds its shard through run();
    the fetched losses are those of the GLOBAL batch and the replicas end up identical"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels as fk
    from oracle import pggan_oracle as P
    from t2i_b200.models.pggan.pggan import PGGAN
    from test_pggan_cpu import TINY, boosted_params
    torch.set_num_threads(1)
    cfg = P.PgganCfg(stage=2, trans=True, **dict(TINY, batch_size=4))
    gb, b = cfg.batch_size, cfg.batch_size // world
    p = boosted_params(cfg)
    feed = P.make_feed(cfg, 12, torch.float64)
    kw = dict(precision="bf16x3", device="cpu", kernels=fk, use_graphs=False, nf_base=cfg.nf_base, nf_cap=cfg.nf_cap,
              z_dim=cfg.z_dim, embed_dim=cfg.embed_dim, compr_embed_dim=cfg.compr_embed_dim, sample_num=2, d_embed=cfg.d_embed)
    m = PGGAN(b, 10, "/tmp/w", "/tmp/r", None, "/tmp/s", "/tmp/l", 2, True, distributed=True, **kw)
    m.set_variables(p)
    sl = slice(rank * b, (rank + 1) * b)

    def fd(model, sel, noise):
        return {model.x: feed["x"][sel].float(), model.x_mismatch: feed["x_mismatch"][sel].float(),
                model.cond: feed["cond"][sel].float(), model.z: feed["z"][sel].float(),
                model.epsilon: feed["epsilon"][sel].float(), model.cond_noise: feed[noise][sel].float(), model.iter: 4}
    _, d_loss, gp = m.run([m.D_optim, m.D_loss, m.real_gp], fd(m, sl, "tn_eps"))
    _, g_loss = m.run([m.G_optim, m.G_loss], fd(m, sl, "tn_eps_g"))
    eng = m._train_engine()
    flat = torch.cat([eng.flat["d"], eng.flat["g"]])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    res = {"same": all(torch.equal(gathered[0], t) for t in gathered), "d_loss": d_loss, "g_loss": g_loss, "gp": gp,
           "alpha": m.alpha_tra}
    if rank == 0:
        ref = PGGAN(gb, 10, "/tmp/w", "/tmp/r", None, "/tmp/s", "/tmp/l", 2, True, **kw)
        ref.set_variables(p)
        _, res["d_ref"], res["gp_ref"] = ref.run([ref.D_optim, ref.D_loss, ref.real_gp], fd(ref, slice(None), "tn_eps"))
        _, res["g_ref"] = ref.run([ref.G_optim, ref.G_loss], fd(ref, slice(None), "tn_eps_g"))
        torch.save(res, out)
    dist.destroy_process_group()


def test_two_rank_pggan_mirror_fetches_global_losses(tmp_path):
    out = str(tmp_path / "pm.pt")
    port = _free_port()
    mp.spawn(_worker_pggan_mirror, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["same"] and abs(r["alpha"] - 0.4) < 1e-12, r
    for a, b in (("d_loss", "d_ref"), ("gp", "gp_ref"), ("g_loss", "g_ref")):       # split-bf16 rounding only
        assert abs(r[a] - r[b]) < 2e-3 * max(1.0, abs(r[b])), (a, r[a], r[b])


def _worker_stage1_sync_bn(rank, world, port, out):
    """StackGAN stage-I with sync_bn=True on 2 ranks == the single-process iteration on the global batch: d_net has a
    BatchNorm after almost every conv (models/stackgan/stageI/model.py:81-112), so the equivalence needs the
    per-call statistics and the backward reductions of BOTH networks summed over the ranks"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels as fk
    from oracle import stackgan1_oracle as S
    from t2i_b200.engine_stage1 import StageIEngine
    from test_stackgan1_cpu import TINY, boosted_params, _bias_before_bn
    torch.set_num_threads(1)
    cfg = S.Stage1Cfg(**TINY)
    gb, b = cfg.batch_size, cfg.batch_size // world
    p = boosted_params(cfg)
    feed = S.make_feed(cfg, 21, torch.float64)
    calls = []

    def allreduce(t):
        calls.append(t.numel())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    exact = dict(act_dtype=torch.float64, f32_dtype=torch.float64)
    args = (cfg.z_dim, cfg.embed_dim, cfg.compressed_embed_dim, cfg.gf_dim, cfg.df_dim, cfg.d_beta1, cfg.g_beta1,
            cfg.alpha_mismatch, cfg.kl_coeff)
    eng = StageIEngine(fk, "cpu", b, 1, *args, world, allreduce, sync_bn=True, **exact)
    ref = StageIEngine(fk, "cpu", gb, 1, *args, **exact)
    sl = slice(rank * b, (rank + 1) * b)
    res = {}
    for e, sel in ((eng, sl), (ref, slice(None))):
        e.set_params_tf(p)
        e.load_feed(x=feed["x"][sel], x_mismatch=feed["x_mismatch"][sel], cond=feed["cond"][sel], z=feed["z"][sel],
                    tn_eps=feed["tn_eps"][sel])
        e.d_step(cfg.lr)
    res["img"] = float((eng.d["img"][:b] - ref.d["img"][:gb][sl]).abs().max())
    gs, gf = eng.get_grads_tf(), ref.get_grads_tf()
    ok = lambda n: float(gf[n].abs().max()) > 1e-12 and not _bias_before_bn(n)
    res["d_grads"] = max((float((gs[n] - gf[n]).abs().max() / (gf[n].abs().max() + 1e-30)), n) for n in gf
                         if n.startswith("d_net/") and ok(n))
    for e, sel in ((eng, sl), (ref, slice(None))):
        e.load_feed(tn_eps=feed["tn_eps_g"][sel])
        e.g_step(cfg.lr)
    gs, gf = eng.get_grads_tf(), ref.get_grads_tf()
    res["g_grads"] = max((float((gs[n] - gf[n]).abs().max() / (gf[n].abs().max() + 1e-30)), n) for n in gf
                         if n.startswith("g_net/") and ok(n))
    sc, scr = eng.scalars_dict(), ref.scalars_dict()
    res["scalars"] = max(abs(sc[k] - scr[k]) / max(1.0, abs(scr[k])) for k in sc)
    ps, pf = eng.get_params_tf(), ref.get_params_tf()
    res["moving"] = max(float((ps[n] - pf[n]).abs().max()) for n in pf if "moving" in n)
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_two_rank_stage1_sync_bn_iteration_equals_single_process(tmp_path):
    out = str(tmp_path / "s1s.pt")
    mp.spawn(_worker_stage1_sync_bn, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["img"] < 1e-10 and r["d_grads"][0] < 1e-8 and r["g_grads"][0] < 1e-7, r
    assert r["scalars"] < 1e-9 and r["moving"] < 1e-10, r


def _worker_stage2_sync_bn(rank, world, port, out):
    """StackGAN stage-II with sync_bn=True on 2 ranks == the single-process iteration on the global batch (the frozen
    stage-I generator, stageII_g_net and stageII_d_net all normalise with whole-batch statistics)"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels as fk
    from oracle import stackgan2_oracle as S2
    from t2i_b200.engine_stage2 import StageIIEngine
    from test_stackgan2_cpu import TINY, boosted_params, _bias_before_bn
    torch.set_num_threads(1)
    cfg = S2.Stage2Cfg(**dict(TINY, batch_size=4))
    gb, b = cfg.batch_size, cfg.batch_size // world
    p = boosted_params(cfg)
    feed = S2.make_feed(cfg, 21, torch.float64)

    def allreduce(t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    exact = dict(act_dtype=torch.float64, f32_dtype=torch.float64)
    args = (cfg.z_dim, cfg.embed_dim, cfg.compressed_embed_dim, cfg.gf_dim, cfg.df_dim, cfg.d_beta1, cfg.g_beta1,
            cfg.alpha_mismatch, cfg.kl_coeff)
    eng = StageIIEngine(fk, "cpu", b, 1, *args, world, allreduce, s1_gf=cfg.s1_gf_dim, sync_bn=True, **exact)
    ref = StageIIEngine(fk, "cpu", gb, 1, *args, s1_gf=cfg.s1_gf_dim, **exact)
    sl = slice(rank * b, (rank + 1) * b)
    res = {}
    for e, sel in ((eng, sl), (ref, slice(None))):
        e.set_params_tf(p)
        e.load_feed(x=feed["x"][sel], x_mismatch=feed["x_mismatch"][sel], cond=feed["cond"][sel], z=feed["z"][sel],
                    tn_eps=feed["tn_eps"][sel], tn_s1=feed["tn_s1"][sel])
        e.d_step(cfg.lr)
    res["img"] = float((eng.d["img"][:b] - ref.d["img"][:gb][sl]).abs().max())
    gs, gf = eng.get_grads_tf(), ref.get_grads_tf()
    ok = lambda n: float(gf[n].abs().max()) > 1e-12 and not _bias_before_bn(n)
    res["d_grads"] = max((float((gs[n] - gf[n]).abs().max() / (gf[n].abs().max() + 1e-30)), n) for n in gf
                         if n.startswith(S2.D2) and ok(n))
    for e, sel in ((eng, sl), (ref, slice(None))):
        e.load_feed(tn_eps=feed["tn_eps_g"][sel], tn_s1=feed["tn_s1_g"][sel])
        e.g_step(cfg.lr)
    gs, gf = eng.get_grads_tf(), ref.get_grads_tf()
    res["g_grads"] = max((float((gs[n] - gf[n]).abs().max() / (gf[n].abs().max() + 1e-30)), n) for n in gf
                         if n.startswith(S2.G2) and ok(n))
    sc, scr = eng.scalars_dict(), ref.scalars_dict()
    res["scalars"] = max(abs(sc[k] - scr[k]) / max(1.0, abs(scr[k])) for k in sc)
    ps, pf = eng.get_params_tf(), ref.get_params_tf()
    res["moving"] = max(float((ps[n] - pf[n]).abs().max()) for n in pf if "moving" in n)
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_two_rank_stage2_sync_bn_iteration_equals_single_process(tmp_path):
    out = str(tmp_path / "s2s.pt")
    mp.spawn(_worker_stage2_sync_bn, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["img"] < 1e-10 and r["d_grads"][0] < 1e-8 and r["g_grads"][0] < 1e-7, r
    assert r["scalars"] < 1e-9 and r["moving"] < 1e-10, r
