"""Host orchestration of one wgancls iteration on the CUDA kernels (no arithmetic happens here).

What the reference expresses as a TF-1.4 graph plus tf.gradients / AdamOptimizer.minimize
(models/wgancls/model.py:34-106) is scheduled explicitly:

  D run (trainer.py:97):  G forward -> D forward on the 4B batch [fake | real | mismatch | x_hat]
      -> seeded D backward (input gradients at every layer; x_hat's goes down to the image and to
      cond) -> slopes / one-sided penalties / tangent seeds -> tangent forward of the x_hat segment
      IN PLACE over its activations (d_net is piecewise linear: d/dtheta of the penalty is the
      weight gradient of the JVP) -> ONE merged weight-gradient pass over the 4B
      batch -> [allreduce] -> kt SGD, Adam.
  G run (trainer.py:101): G forward (fresh noise) -> D forward -> D input-gradient -> G backward
      (BatchNorm backward included) -> [allreduce] -> Adam, BN moving statistics.

Parameters live in kernel layout in two flat fp32 buffers (d / g) with flat gradient and Adam
buffers beside them; conversion to/from the reference's TF variable layout happens only in
``set_params_tf`` / ``get_params_tf`` (checkpoint boundary).  ``K`` is the kernel module
(text-to-image_b200/kernels.py); tests inject their CPU restatement to check this file on CPU.
"""
import contextlib
import math
from collections import OrderedDict

import torch

GP_WEIGHT = 150.0   # models/wgancls/model.py:91
KT_LR = 0.001       # models/wgancls/model.py:100
KT_INIT = 0.7       # models/wgancls/model.py:77
BN_EPS = 1e-5       # utils/ops.py:7
BN_DECAY = 0.9
ADAM_EPS = 1e-8
SUMS = 8            # per-net tail of the flat gradient buffer holding the scalar sums
IMG = 64            # the path is only valid for 64x64 (model.py:154 hard-codes the 4x4 tile)


def _align(n, a=64):
    return (n + a - 1) // a * a


class Layer:
    """One contraction layer: kernel-layout master weight/bias views, gradient views, packed copies."""

    def __init__(self, name, kind, tf_w, tf_b, mode, k, taps, cout, cin, need_bwd=True):
        self.name, self.kind, self.tf_w, self.tf_b = name, kind, tf_w, tf_b
        self.mode, self.k, self.taps, self.cout, self.cin = mode, k, taps, cout, cin
        self.need_bwd = need_bwd
        self.w = self.b = self.gw = self.gb = self.Wf = None
        # kind "conv_pad": the TF tensor has cin_tf x cout_tf channels, zero-padded to cin x cout in kernel layout
        self.cin_tf = self.cout_tf = None


class HostScalarRing:
    """Small host -> device scalar uploads that stay correct when the host runs ahead of the device.

    A single pinned staging slot is wrong: ``dst.copy_(slot, non_blocking=True)`` only ENQUEUES the copy, so the host
    can overwrite the slot with step t+1's value before the copy for step t has executed (a loop that never fetches a
    loss, e.g. bench.py's resident steps, does not synchronise).  Here every upload takes the next of ``depth`` pinned
    slots and records an event behind its copy; a slot is rewritten only after its previous copy has completed."""

    def __init__(self, width, dtype, cuda, depth=16):
        self.cuda, self.i = cuda, 0
        self.slots = [torch.zeros(width, dtype=dtype) for _ in range(depth if cuda else 1)]
        if cuda:
            self.slots = [t.pin_memory() for t in self.slots]
        self.events = [None] * len(self.slots)

    def upload(self, values, dst):
        slot, ev = self.slots[self.i], self.events[self.i]
        if ev is not None:
            ev.synchronize()                 # normally long done: `depth` uploads ago
        for j, v in enumerate(values):
            slot[j] = v
        dst.copy_(slot, non_blocking=True)
        if self.cuda:
            if ev is None:
                ev = self.events[self.i] = torch.cuda.Event()
            ev.record()
        self.i = (self.i + 1) % len(self.slots)


class Engine:
    def __init__(self, K, device, batch, np_=1, z_dim=128, embed_dim=1024, ce=128, gf=128, df=128,
                 beta1=0.0, beta2=0.9, kl_coeff=1.0, world=1, allreduce=None, act_dtype=torch.bfloat16,
                 f32_dtype=torch.float32, share_from=None, use_graphs=False, concurrent=True, sync_bn=False,
                 beta1_g=None, g_buckets=None):
        # act_dtype / f32_dtype exist for the CPU host-logic tests only (exact fp64 storage with the
        # kernels' CPU restatement); the CUDA kernels accept bf16 planes and fp32 exclusively.
        self.K, self.dev, self.B, self.np = K, torch.device(device), batch, np_
        self.act_dtype, self.f32_dtype = act_dtype, f32_dtype
        # the generator's dense input [z | c] needs a multiple of 8 columns (16-byte TMA strides): z is zero-padded
        self.Z_tf = z_dim
        z_dim = _align(z_dim, 8)
        self.Z, self.E, self.ce, self.gf, self.df = z_dim, embed_dim, ce, gf, df
        self.beta1, self.beta2, self.kl_coeff = beta1, beta2, kl_coeff
        self.beta1_net = {"d": beta1, "g": beta1 if beta1_g is None else beta1_g}
        self.world, self.allreduce = world, allreduce
        # sync_bn: all-reduce g_net's BatchNorm sums over the ranks (forward and backward), i.e. normalise over the
        # GLOBAL batch exactly as the single-device reference does (utils/ops.py:20-29); default: per-replica
        # statistics and a single all-reduce per optimizer step
        self.sync_bn = bool(sync_bn) and world > 1
        self.GB = batch * world
        # g_buckets = 2: the G-gradient all-reduce goes out in two pieces in backward order -- the layers from the first
        # transposed conv on (two thirds of the bytes, final when the backward pass reaches the 4x4 maps) travel on the
        # communication stream UNDER the rest of the backward pass.  Default: whenever there are streams to overlap on.
        if g_buckets is None:
            import os
            env = os.environ.get("T2I_G_BUCKETS")          # development aid: A/B the bucketing
            g_buckets = int(env) if env else (2 if (world > 1 and self.dev.type == "cuda") else 1)
        self.g_buckets = g_buckets
        for v in (z_dim, embed_dim, ce, gf, df):
            assert v % 8 == 0, "channel counts must be multiples of 8"
        self.d_t = 0
        self.g_t = 0
        self.use_graphs = use_graphs and self.dev.type == "cuda" and not self.sync_bn   # collectives stay outside graphs
        self._graphs = {}
        self.side_stream = torch.cuda.Stream(self.dev) if (self.dev.type == "cuda" and concurrent) else None
        self.comm_stream = torch.cuda.Stream(self.dev) if (self.dev.type == "cuda" and concurrent) else None
        self.copy_stream = torch.cuda.Stream(self.dev) if (self.dev.type == "cuda" and concurrent) else None
        self._scalars_host = self._scalars_event = None
        self._scalars_published = False
        self._img_free = None           # recorded after the D run's last read of the fed image segments
        self.replayed_launches = 0      # kernels executed through graph replays
        self.captured_launches = 0      # kernels recorded (not executed) during captures
        if share_from is None:
            before = set(self.__dict__)
            self._build_params()
            self._param_attrs = sorted(set(self.__dict__) - before)
        else:   # a second batch size (sampler / eval) on the same parameters
            assert (share_from.np, share_from.gf, share_from.df) == (np_, gf, df) and type(share_from) is type(self)
            self._param_attrs = share_from._param_attrs
            for k in self._param_attrs:
                setattr(self, k, getattr(share_from, k))
        self._build_buffers()

    # ------------------------------------------------------------------ parameters
    FC0_NCHW = True      # dense output features ordered c*16 + hw (wgancls reshapes to NCHW, model.py:179)
    NORM_MOVING = True   # the normalisation layers carry moving statistics (BatchNorm; PGGAN's LayerNorm does not)

    def _d_layers(self):
        """d_net contraction layers in creation order (models/wgancls/model.py:129-161) and its BatchNorm list."""
        K = self.K
        df, ce, E = self.df, self.ce, self.E
        S1, K4 = K.CONV_S1, K.CONV_K4S2
        d, L = "d_net/", Layer
        layers = [
            L("h0", "col_in", d + "Conv", d + "Conv", K4, 4, 1, df, 64),
            L("h1", "conv", d + "Conv_1", d + "Conv_1", K4, 4, 16, 2 * df, df),
            L("h2", "conv", d + "Conv_2", d + "Conv_2", K4, 4, 16, 4 * df, 2 * df),
            L("h3", "conv", d + "Conv_3", d + "Conv_3", K4, 4, 16, 8 * df, 4 * df),
            L("r1", "conv", d + "Conv_4", d + "Conv_4", S1, 1, 1, 2 * df, 8 * df),
            L("r2", "conv", d + "Conv_5", d + "Conv_5", S1, 3, 9, 4 * df, 2 * df),
            L("r3", "conv", d + "Conv_6", d + "Conv_6", S1, 3, 9, 8 * df, 4 * df),
            L("efc", "dense", d + "dense", d + "dense", S1, 1, 1, ce, E),
            L("h5", "conv", d + "Conv_7", d + "Conv_7", S1, 3, 9, 8 * df, 8 * df + ce),
            L("h6", "conv", d + "Conv_8", d + "Conv_8", S1, 1, 1, 8 * df, 8 * df),
            L("out", "dout", d + "Conv_9", d + "Conv_9", None, 4, 1, 1, 16 * 8 * df, need_bwd=False),
        ]
        return layers, [], []       # layers, BatchNorm channel counts, BatchNorm TF scopes

    def _g_layers(self):
        """g_net contraction layers in creation order (models/wgancls/model.py:163-225) and its BatchNorm list."""
        K = self.K
        gf, ce, E, Z = self.gf, self.ce, self.E, self.Z
        C1, C2, C4, C8 = gf, 2 * gf, 4 * gf, 8 * gf
        S1, DC = K.CONV_S1, K.DECONV_K4S2
        g, L = "g_net/", Layer
        layers = [
            L("ms", "ms", (g + "dense", g + "dense_1"), (g + "dense", g + "dense_1"), S1, 1, 1, 2 * ce, E,
              need_bwd=False),
            L("fc0", "fc0", g + "dense_2", g + "dense_2", S1, 1, 1, 16 * C8, Z + ce),
            L("c0", "conv", g + "Conv", g + "Conv", S1, 1, 1, C2, C8),
            L("c1", "conv", g + "Conv_1", g + "Conv_1", S1, 3, 9, C2, C2),
            L("c2", "conv", g + "Conv_2", g + "Conv_2", S1, 3, 9, C8, C2),
            L("t0", "deconv", g + "Conv2d_transpose", g + "Conv2d_transpose", DC, 4, 16, C4, C8),
            L("c3", "conv", g + "Conv_3", g + "Conv_3", S1, 3, 9, C4, C4),
            L("c4", "conv", g + "Conv_4", g + "Conv_4", S1, 1, 1, C1, C4),
            L("c5", "conv", g + "Conv_5", g + "Conv_5", S1, 3, 9, C1, C1),
            L("c6", "conv", g + "Conv_6", g + "Conv_6", S1, 3, 9, C4, C1),
            L("t1", "deconv", g + "Conv2d_transpose_1", g + "Conv2d_transpose_1", DC, 4, 16, C2, C4),
            L("c7", "conv", g + "Conv_7", g + "Conv_7", S1, 3, 9, C2, C2),
            L("t2", "deconv", g + "Conv2d_transpose_2", g + "Conv2d_transpose_2", DC, 4, 16, C1, C2),
            L("c8", "conv", g + "Conv_8", g + "Conv_8", S1, 3, 9, C1, C1),
            L("t3", "col_out", g + "Conv2d_transpose_3", g + "Conv2d_transpose_3", S1, 1, 1, 64, C1),
            L("c9", "c9", g + "Conv_9", g + "Conv_9", None, 3, 9, 3, 3, need_bwd=False),
        ]
        # BatchNorm layers of g_net in creation order (model.py:176-216)
        bn_ch = [16 * C8, C2, C2, C8, C4, C1, C1, C4, C2, C1]
        bn_tf = [g + "BatchNorm" + ("" if i == 0 else "_%d" % i) for i in range(10)]
        return layers, bn_ch, bn_tf

    def _build_params(self):
        d_layers, self.dbn_ch, self.dbn_tf = self._d_layers()
        g_layers, self.bn_ch, self.bn_tf = self._g_layers()
        self.dl = OrderedDict((l.name, l) for l in d_layers)
        self.gl = OrderedDict((l.name, l) for l in g_layers)

        def bias_len(l):
            return {"dout": 1, "dout_fc": 1, "col_out": 3, "c9": 3}.get(l.kind, l.cout)     # img_out: 8 (3 used, zero padded)

        # flat order of g_net: the layers whose gradients are final FIRST in the backward pass (t0 .. c9) lead, so that
        # each all-reduce bucket is one contiguous range: [t0 .. c9 | ms, fc0, c0, c1, c2 | BatchNorm | sums]
        early = ("ms", "fc0", "c0", "c1", "c2")
        bucketed = "t0" in self.gl and all(n in self.gl for n in early)

        def layout(layers, bn_ch):
            off, table = 0, OrderedDict()
            if layers is self.gl and bucketed:
                layers = OrderedDict([(n, l) for n, l in layers.items() if n not in early] +
                                     [(n, layers[n]) for n in early])
            for l in layers.values():
                wn = l.taps * l.cout * l.cin if l.kind not in ("dout", "dout_fc", "c9") else (81 if l.kind == "c9" else l.cin)
                table[l.name + ".w"] = (off, wn); off = _align(off + wn)
                table[l.name + ".b"] = (off, bias_len(l)); off = _align(off + bias_len(l))
            for i, c in enumerate(bn_ch):
                table["bn%d.gamma" % i] = (off, c); off = _align(off + c)
                table["bn%d.beta" % i] = (off, c); off = _align(off + c)
            return off, table

        self.d_n, self.d_table = layout(self.dl, self.dbn_ch)
        self.g_n, self.g_table = layout(self.gl, self.bn_ch)
        self.g_split = self.g_table["ms.w"][0] if bucketed else None      # first element of the second bucket
        f32 = dict(device=self.dev, dtype=self.f32_dtype)
        self.flat = {"d": torch.zeros(self.d_n, **f32), "g": torch.zeros(self.g_n, **f32)}
        self.grad = {"d": torch.zeros(self.d_n + SUMS, **f32), "g": torch.zeros(self.g_n + SUMS, **f32)}
        self.adam_m = {k: torch.zeros_like(v) for k, v in self.flat.items()}
        self.adam_v = {k: torch.zeros_like(v) for k, v in self.flat.items()}
        self.sums = {"d": self.grad["d"][self.d_n:], "g": self.grad["g"][self.g_n:]}
        # bf16-plane mirror of the whole flat master (written by Adam itself); the layers' tensor-core
        # weights are views into it.  One layout serves forward (NK) and input-gradient (KN) GEMMs.
        self.packed = {k: torch.zeros(self.np, v.numel(), device=self.dev, dtype=self.act_dtype)
                       for k, v in self.flat.items()}
        self.lr_t = {k: torch.zeros(1, **f32) for k in self.flat}
        self.lr_ring = HostScalarRing(1, self.f32_dtype, self.dev.type == "cuda")
        self.P, self.G = {}, {}
        for net, table in (("d", self.d_table), ("g", self.g_table)):
            for name, (off, n) in table.items():
                self.P[net + "." + name] = self.flat[net][off:off + n]
                self.G[net + "." + name] = self.grad[net][off:off + n]
        bf = dict(device=self.dev, dtype=self.act_dtype)
        for net, layers in (("d", self.dl), ("g", self.gl)):
            for l in layers.values():
                l.w, l.b = self.P["%s.%s.w" % (net, l.name)], self.P["%s.%s.b" % (net, l.name)]
                l.gw, l.gb = self.G["%s.%s.w" % (net, l.name)], self.G["%s.%s.b" % (net, l.name)]
                if l.kind in ("dout", "dout_fc", "c9"):
                    continue
                l.w = l.w.view(l.taps, l.cout, l.cin)
                l.gw = l.gw.view(l.taps, l.cout, l.cin)
                off, n = (self.d_table if net == "d" else self.g_table)[l.name + ".w"]
                l.Wf = self.packed[net][:, off:off + n].view(self.np, l.taps, l.cout, l.cin)
        ng = range(len(self.bn_ch))
        self.bn_gamma = [self.P["g.bn%d.gamma" % i] for i in ng]
        self.bn_beta = [self.P["g.bn%d.beta" % i] for i in ng]
        self.bn_dgamma = [self.G["g.bn%d.gamma" % i] for i in ng]
        self.bn_dbeta = [self.G["g.bn%d.beta" % i] for i in ng]
        self.bn_mm = [torch.zeros(c, **f32) for c in self.bn_ch]
        self.bn_mv = [torch.ones(c, **f32) for c in self.bn_ch]
        self.bn_mean = [torch.zeros(c, **f32) for c in self.bn_ch]
        self.bn_var = [torch.zeros(c, **f32) for c in self.bn_ch]
        self.bn_rstd = [torch.zeros(c, **f32) for c in self.bn_ch]
        # d_net BatchNorm (none in wgancls; StackGAN stage-I has seven)
        nd = range(len(self.dbn_ch))
        self.dbn_gamma = [self.P["d.bn%d.gamma" % i] for i in nd]
        self.dbn_beta = [self.P["d.bn%d.beta" % i] for i in nd]
        self.dbn_dgamma = [self.G["d.bn%d.gamma" % i] for i in nd]
        self.dbn_dbeta = [self.G["d.bn%d.beta" % i] for i in nd]
        self.dbn_mm = [torch.zeros(c, **f32) for c in self.dbn_ch]
        self.dbn_mv = [torch.ones(c, **f32) for c in self.dbn_ch]
        self.dbn_mean = [torch.zeros(c, **f32) for c in self.dbn_ch]
        self.dbn_var = [torch.zeros(c, **f32) for c in self.dbn_ch]
        self.dbn_rstd = [torch.zeros(c, **f32) for c in self.dbn_ch]
        for gmm in self.bn_gamma + self.dbn_gamma:
            gmm.fill_(1.0)
        self.kt = torch.full((1,), KT_INIT, **f32)
        self.scalars = torch.zeros(16, **f32)

    # -- layout conversion between the reference's TF variables and the kernel layout ---------
    def _perm_fc0(self, v_tf, inverse=False):
        """feature order: TF c*16 + hw (NCHW reshape, model.py:179)  <->  kernel hw*C8 + c (NHWC)."""
        if not self.FC0_NCHW:      # NHWC reshape: the dense features already are hw*C8 + c
            return v_tf
        C8 = 8 * self.gf
        lead = v_tf.shape[:-1]
        if not inverse:
            return v_tf.reshape(*lead, C8, 16).transpose(-1, -2).reshape(*lead, 16 * C8)
        return v_tf.reshape(*lead, 16, C8).transpose(-1, -2).reshape(*lead, 16 * C8)

    def _w_to_kernel(self, l, p):
        if l.kind == "conv":
            w = p[l.tf_w + "/weights"]
            return w.permute(0, 1, 3, 2).reshape(l.taps, l.cout, l.cin)
        if l.kind == "deconv":
            return p[l.tf_w + "/weights"].reshape(l.taps, l.cout, l.cin)
        if l.kind == "dense":
            return p[l.tf_w + "/kernel"].t().reshape(1, l.cout, l.cin)
        if l.kind == "ms":
            return torch.cat([p[l.tf_w[0] + "/kernel"].t(), p[l.tf_w[1] + "/kernel"].t()], 0).reshape(1, l.cout, l.cin)
        if l.kind == "fc0":      # TF rows [z (Z_tf) | c (ce)] -> kernel columns [z padded to Z | c]
            w = self._perm_fc0(p[l.tf_w + "/kernel"]).t()
            out = torch.zeros(l.cout, l.cin, dtype=w.dtype)
            out[:, :self.Z_tf] = w[:, :self.Z_tf]
            out[:, self.Z:] = w[:, self.Z_tf:]
            return out.reshape(1, l.cout, l.cin)
        if l.kind == "col_in":   # [4,4,3,co] -> [co, (kh*4+kw)*3+c], zero padded to 64 columns
            w = p[l.tf_w + "/weights"]
            out = torch.zeros(1, l.cout, 64, dtype=w.dtype)
            out[0, :, :48] = w.reshape(48, l.cout).t()
            return out
        if l.kind == "col_out":  # deconv [4,4,co=3,ci] -> [(kh*4+kw)*3+co, ci], zero padded to 64 rows
            w = p[l.tf_w + "/weights"]
            out = torch.zeros(1, 64, l.cin, dtype=w.dtype)
            out[0, :48] = w.reshape(48, l.cin)
            return out
        if l.kind == "col3_in":  # 3x3 conv on 3 channels: [3,3,3,co] -> [co, (kh*3+kw)*3+c], zero padded to 32 columns
            w = p[l.tf_w + "/weights"]
            out = torch.zeros(1, l.cout, 32, dtype=w.dtype)
            out[0, :, :27] = w.reshape(27, l.cout).t()
            return out
        if l.kind == "img_out":  # 3x3 conv to 3 channels: [3,3,ci,3] -> [tap][8][ci], output rows 3..7 zero
            w = p[l.tf_w + "/weights"]
            out = torch.zeros(9, 8, l.cin, dtype=w.dtype)
            out[:, :3] = w.permute(0, 1, 3, 2).reshape(9, 3, l.cin)
            return out
        if l.kind == "dout":     # [4,4,C,1] -> (kh,kw,c) flat == NHWC order of the 4x4xC activation
            return p[l.tf_w + "/weights"].reshape(-1)
        if l.kind == "dout_fc":  # dense [C, 1] on the last axis of a [B,1,1,C] tensor (models/pggan/pggan.py:275)
            return p[l.tf_w + "/kernel"].reshape(-1)
        if l.kind == "conv_pad":  # [k,k,ci_tf,co_tf] -> [tap][co][ci], zero padded in both channel counts
            w = p[l.tf_w + "/weights"]
            out = torch.zeros(l.taps, l.cout, l.cin, dtype=w.dtype)
            out[:, :l.cout_tf, :l.cin_tf] = w.permute(0, 1, 3, 2).reshape(l.taps, l.cout_tf, l.cin_tf)
            return out
        if l.kind == "flat4":    # 4x4 VALID conv on a 4x4 map = dense over the NHWC-flattened map: [4,4,ci,co] -> [co][(kh,kw,ci)]
            w = p[l.tf_w + "/weights"]
            return w.reshape(-1, l.cout).t().reshape(1, l.cout, l.cin)
        if l.kind == "c9":
            return p[l.tf_w + "/weights"].reshape(-1)
        raise ValueError(l.kind)

    def _w_to_tf(self, l, w, out):
        if l.kind == "conv":
            out[l.tf_w + "/weights"] = w.reshape(l.k, l.k, l.cout, l.cin).permute(0, 1, 3, 2).contiguous()
        elif l.kind == "deconv":
            out[l.tf_w + "/weights"] = w.reshape(4, 4, l.cout, l.cin).clone()
        elif l.kind == "dense":
            out[l.tf_w + "/kernel"] = w.reshape(l.cout, l.cin).t().contiguous()
        elif l.kind == "ms":
            w2 = w.reshape(l.cout, l.cin)
            out[l.tf_w[0] + "/kernel"] = w2[:self.ce].t().contiguous()
            out[l.tf_w[1] + "/kernel"] = w2[self.ce:].t().contiguous()
        elif l.kind == "fc0":
            w2 = w.reshape(l.cout, l.cin)
            w2 = torch.cat([w2[:, :self.Z_tf], w2[:, self.Z:]], 1)
            out[l.tf_w + "/kernel"] = self._perm_fc0(w2.t(), inverse=True).contiguous()
        elif l.kind == "col_in":
            out[l.tf_w + "/weights"] = w.reshape(l.cout, 64)[:, :48].t().reshape(4, 4, 3, l.cout).contiguous()
        elif l.kind == "col_out":
            out[l.tf_w + "/weights"] = w.reshape(64, l.cin)[:48].reshape(4, 4, 3, l.cin).clone()
        elif l.kind == "col3_in":
            out[l.tf_w + "/weights"] = w.reshape(l.cout, 32)[:, :27].t().reshape(3, 3, 3, l.cout).contiguous()
        elif l.kind == "img_out":
            out[l.tf_w + "/weights"] = w.reshape(3, 3, 8, l.cin)[:, :, :3].permute(0, 1, 3, 2).contiguous()
        elif l.kind == "dout":
            out[l.tf_w + "/weights"] = w.reshape(4, 4, -1, 1).clone()
        elif l.kind == "dout_fc":
            out[l.tf_w + "/kernel"] = w.reshape(-1, 1).clone()
        elif l.kind == "conv_pad":
            w = w.reshape(l.k, l.k, l.cout, l.cin)[:, :, :l.cout_tf, :l.cin_tf]
            out[l.tf_w + "/weights"] = w.permute(0, 1, 3, 2).contiguous()
        elif l.kind == "flat4":
            out[l.tf_w + "/weights"] = w.reshape(l.cout, l.cin).t().reshape(4, 4, l.cin // 16, l.cout).contiguous()
        elif l.kind == "c9":
            out[l.tf_w + "/weights"] = w.reshape(3, 3, 3, 3).clone()

    def _b_names(self, l):
        leaf = "/bias" if l.kind in ("dense", "ms", "fc0", "dout_fc") else "/biases"
        return [n + leaf for n in (l.tf_b if isinstance(l.tf_b, tuple) else (l.tf_b,))]

    def _views(self, flats):
        """name -> view for a pair of flat buffers laid out like the parameters ({"d": ..., "g": ...})"""
        out = {}
        for net, table in (("d", self.d_table), ("g", self.g_table)):
            for name, (off, n) in table.items():
                out[net + "." + name] = flats[net][off:off + n]
        return out

    def _import(self, p, views, with_moving):
        p = {k: torch.as_tensor(v).detach().to("cpu", self.f32_dtype) for k, v in p.items()}
        for net, layers in (("d", self.dl), ("g", self.gl)):
            for l in layers.values():
                views["%s.%s.w" % (net, l.name)].copy_(self._w_to_kernel(l, p).reshape(-1))
                b = torch.cat([p[n] for n in self._b_names(l)])
                if l.kind == "fc0":
                    b = self._perm_fc0(b)
                if l.kind == "img_out":
                    b = torch.cat([b, torch.zeros(5, dtype=b.dtype)])
                if l.kind == "conv_pad":
                    b = torch.cat([b, torch.zeros(l.cout - l.cout_tf, dtype=b.dtype)])
                views["%s.%s.b" % (net, l.name)].copy_(b)
        for i, scope in enumerate(self.bn_tf):
            perm = (lambda v: self._perm_fc0(v)) if i == 0 else (lambda v: v)
            views["g.bn%d.gamma" % i].copy_(perm(p[scope + "/gamma"]))
            views["g.bn%d.beta" % i].copy_(perm(p[scope + "/beta"]))
            if with_moving and self.NORM_MOVING:
                self.bn_mm[i].copy_(perm(p[scope + "/moving_mean"]))
                self.bn_mv[i].copy_(perm(p[scope + "/moving_variance"]))
        for i, scope in enumerate(self.dbn_tf):
            views["d.bn%d.gamma" % i].copy_(p[scope + "/gamma"])
            views["d.bn%d.beta" % i].copy_(p[scope + "/beta"])
            if with_moving:
                self.dbn_mm[i].copy_(p[scope + "/moving_mean"])
                self.dbn_mv[i].copy_(p[scope + "/moving_variance"])

    def set_params_tf(self, p):
        """Load parameters given in the reference's TF variable layout (oracle / checkpoint names)."""
        self.join_comm()
        self._import(p, self.P, True)
        self.repack("d")
        self.repack("g")

    def get_adam_tf(self):
        """Adam slots in TF layout ('m/<variable>', 'v/<variable>') and the two step counters."""
        self.join_comm()
        out = OrderedDict()
        for tag, flats in (("m", self.adam_m), ("v", self.adam_v)):
            for k, v in self._export(self._views(flats), False).items():
                out[tag + "/" + k] = v
        out["d_t"], out["g_t"] = torch.tensor(self.d_t), torch.tensor(self.g_t)
        return out

    def set_adam_tf(self, state):
        self.join_comm()
        for tag, flats in (("m", self.adam_m), ("v", self.adam_v)):
            sub = {k[2:]: v for k, v in state.items() if k.startswith(tag + "/")}
            if sub:
                self._import(sub, self._views(flats), False)
        self.d_t = int(state.get("d_t", self.d_t))
        self.g_t = int(state.get("g_t", self.g_t))

    def _export(self, flat_views, include_moving):
        out = OrderedDict()
        for net, layers in (("d", self.dl), ("g", self.gl)):
            for l in layers.values():
                self._w_to_tf(l, flat_views["%s.%s.w" % (net, l.name)].detach().cpu(), out)
                b = flat_views["%s.%s.b" % (net, l.name)].detach().cpu()
                if l.kind == "fc0":
                    b = self._perm_fc0(b, inverse=True)
                if l.kind == "img_out":
                    b = b[:3]
                if l.kind == "conv_pad":
                    b = b[:l.cout_tf]
                names = self._b_names(l)
                for j, n in enumerate(names):
                    out[n] = b.reshape(len(names), -1)[j].clone()
        for i, scope in enumerate(self.bn_tf):
            perm = (lambda v: self._perm_fc0(v, inverse=True)) if i == 0 else (lambda v: v)
            out[scope + "/gamma"] = perm(flat_views["g.bn%d.gamma" % i].detach().cpu()).clone()
            out[scope + "/beta"] = perm(flat_views["g.bn%d.beta" % i].detach().cpu()).clone()
            if include_moving and self.NORM_MOVING:
                out[scope + "/moving_mean"] = perm(self.bn_mm[i].detach().cpu()).clone()
                out[scope + "/moving_variance"] = perm(self.bn_mv[i].detach().cpu()).clone()
        for i, scope in enumerate(self.dbn_tf):
            out[scope + "/gamma"] = flat_views["d.bn%d.gamma" % i].detach().cpu().clone()
            out[scope + "/beta"] = flat_views["d.bn%d.beta" % i].detach().cpu().clone()
            if include_moving:
                out[scope + "/moving_mean"] = self.dbn_mm[i].detach().cpu().clone()
                out[scope + "/moving_variance"] = self.dbn_mv[i].detach().cpu().clone()
        return out

    def get_params_tf(self):
        self.join_comm()
        return self._export(self.P, True)

    def get_grads_tf(self):
        """Gradients of the last d_step / g_step in TF layout (the other net's entries are stale)."""
        self.join_comm()
        return self._export(self.G, False)

    def repack(self, net):
        """fp32 master -> its bf16-plane mirror (checkpoint load; the optimizer step does it itself)."""
        self.K.to_planes(self.flat[net].view(-1, 64), self.packed[net].view(self.np, -1, 64))

    # ------------------------------------------------------------------ buffers
    def _build_buffers(self):
        self._build_d_buffers()
        self._build_g_buffers()

    def _planes(self, *shape):
        return torch.zeros(self.np, *shape, device=self.dev, dtype=self.act_dtype)

    def _build_d_buffers(self):
        B, S = self.B, 4 * self.B
        df, ce, E = self.df, self.ce, self.E
        f32 = dict(device=self.dev, dtype=self.f32_dtype)
        planes = self._planes
        d = self.d = {}
        d["img"] = torch.zeros(S, IMG, IMG, 3, **f32)     # [fake | real | mismatch | x_hat]
        # the same images as padded bf16 rows (what the first conv's kernels stream); after the penalty seeds exist
        # segment 3 holds the TANGENT image coef * dD/dx_hat instead
        d["rows"] = planes(S, IMG, self.K.img_row_pitch(IMG))
        dshapes = {"a0": (S, 32, 32, df), "a1": (S, 16, 16, 2 * df), "a2": (S, 8, 8, 4 * df), "a3": (S, 4, 4, 8 * df),
                   "r1": (S, 4, 4, 2 * df), "r2": (S, 4, 4, 4 * df), "cat": (S, 4, 4, 8 * df + ce),
                   "a5": (S, 4, 4, 8 * df), "a6": (S, 4, 4, 8 * df)}
        for n, sh in dshapes.items():
            d[n] = planes(*sh)
            d["d_" + n] = planes(*sh)
        d["cond"] = planes(S, E)
        d["e"] = planes(S, ce)
        d["d_e"] = planes(S, ce)
        d["logit"] = torch.zeros(S, **f32)
        d["seed"] = torch.zeros(S, **f32)
        d["gseed"] = torch.full((B,), -1.0 / self.GB, **f32)    # G_loss = -mean D(G) + ...  (model.py:92)
        d["gx"] = torch.zeros(B, IMG, IMG, 3, **f32)       # dD/d image
        d["d_cond"] = planes(B, E)
        d["g2"] = torch.zeros(B, E, **f32)                  # dD/d cond
        for n in ("slope", "coef", "slope2", "coef2"):
            d[n] = torch.zeros(B, **f32)

    def _build_g_buffers(self):
        B = self.B
        gf, ce, E, Z = self.gf, self.ce, self.E, self.Z
        C1, C2, C4, C8 = gf, 2 * gf, 4 * gf, 8 * gf
        f32 = dict(device=self.dev, dtype=self.f32_dtype)
        planes = self._planes
        g = self.g = {}
        gshapes = {"f0": (B, 4, 4, C8), "h0": (B, 4, 4, C8), "t1": (B, 4, 4, C2), "u1": (B, 4, 4, C2),
                   "t2": (B, 4, 4, C2), "u2": (B, 4, 4, C2), "t3": (B, 4, 4, C8), "h1": (B, 4, 4, C8),
                   "d1": (B, 8, 8, C4), "t4": (B, 8, 8, C4), "h2": (B, 8, 8, C4), "t5": (B, 8, 8, C1),
                   "u5": (B, 8, 8, C1), "t6": (B, 8, 8, C1), "u6": (B, 8, 8, C1), "t7": (B, 8, 8, C4),
                   "h3": (B, 8, 8, C4), "d2": (B, 16, 16, C2), "t8": (B, 16, 16, C2), "h4": (B, 16, 16, C2),
                   "d3": (B, 32, 32, C1), "t9": (B, 32, 32, C1), "h5": (B, 32, 32, C1)}
        for n, sh in gshapes.items():
            g[n] = planes(*sh)
            g["d_" + n] = planes(*sh)
        for n, sh in {"cond": (B, E), "zc": (B, Z + ce)}.items():
            g[n] = planes(*sh)
            g["d_" + n] = planes(*sh)
        g["ms"] = torch.zeros(B, 2 * ce, **f32)            # [mean | log_sigma]: fp32 (feeds exp(), model.py:121)
        g["d_ms"] = planes(B, 2 * ce)
        g["u4"] = torch.zeros(B, IMG, IMG, 3, **f32)
        g["d_u4"] = torch.zeros(B, IMG, IMG, 3, **f32)
        g["d_u4_rows"] = planes(B, IMG, self.K.img_row_pitch(IMG))
        g["tn"] = torch.zeros(B, ce, **f32)
        g["z"] = torch.zeros(B, Z, **f32)
        g["kl_scratch"] = torch.zeros(1, **f32)
        # BatchNorm accumulators fed by the conv epilogues: per layer [sum x | sum x^2] (forward) and
        # sum dy * x (backward), one contiguous buffer cleared by a single memset per generator forward
        tot = sum(self.bn_ch)
        self.bn_scratch = torch.zeros(4 * tot, **f32)
        self.bn_fwd_sums, self.bn_bwd_sums, off = [], [], 0
        for c in self.bn_ch:
            self.bn_fwd_sums.append(self.bn_scratch[off:off + 2 * c]); off += 2 * c
        for c in self.bn_ch:      # backward: [sum dy | sum dy * x]; the first half is used only under sync_bn / by BN 0
            self.bn_bwd_sums.append(self.bn_scratch[off:off + 2 * c]); off += 2 * c
        self.bn_dsum = [t[:c] for t, c in zip(self.bn_bwd_sums, self.bn_ch)]
        self.bn_dot = [t[c:] for t, c in zip(self.bn_bwd_sums, self.bn_ch)]
        self.feed = {"cond": torch.zeros(B, E, **f32), "epsilon": torch.zeros(B, **f32)}

    @staticmethod
    def _rows(t):
        """[np, n, h, w, c] planes -> [np, n*h*w, c] alias (rows = pixels)."""
        return t.view(t.shape[0], -1, t.shape[-1])

    # ------------------------------------------------------------------ generator
    def _bn(self, i, x, y, residual=None, relu=False, train=True, update_moving=False):
        """BatchNorm i on x -> y.  Training mode: the conv that produced x left [sum x | sum x^2] in
        bn_fwd_sums[i] (epilogue statistics); one kernel finishes them, normalises and (G run) steps the
        moving statistics.  Inference (sampler, model.py:57): moving statistics."""
        K = self.K
        if train:
            stat_rows = 0
            if self.sync_bn:      # whole-batch statistics: sum the per-rank sums
                self.allreduce(self.bn_fwd_sums[i])
                stat_rows = (x[0].numel() // x.shape[-1]) * self.world
            K.bn_apply_train(x, self.bn_fwd_sums[i], BN_EPS, self.bn_gamma[i], self.bn_beta[i], y, self.bn_mean[i],
                             self.bn_rstd[i], self.bn_var[i], residual=residual, relu=relu,
                             moving=(self.bn_mm[i], self.bn_mv[i]) if update_moving else None, decay=BN_DECAY,
                             stat_rows=stat_rows)
        else:
            K.bn_apply(x, self.bn_mm[i], torch.rsqrt(self.bn_mv[i] + BN_EPS), self.bn_gamma[i], self.bn_beta[i], y,
                       residual, relu)

    def g_forward(self, z, cond, tn_eps, img_out, kl_sum, train=True, cond_noise=True, update_moving=False):
        """models/wgancls/model.py:163-225.  z, cond, tn_eps: fp32 device tensors; image -> img_out.
        update_moving: also run the UPDATE_OPS of utils/ops.py:20-29 (the G run, model.py:98,102)."""
        K, g, gl, V = self.K, self.g, self.gl, self.K.View
        S1, DC = K.CONV_S1, K.DECONV_K4S2
        if train:
            self.bn_scratch.zero_()

        def stats(i):     # batch statistics of BatchNorm i, accumulated by the epilogue of the conv in front of it
            if not train:
                return {}
            c = self.bn_ch[i]
            return dict(stat_sum=self.bn_fwd_sums[i][:c], stat_sq=self.bn_fwd_sums[i][c:])

        def bn(i, x, y, **kw):
            self._bn(i, x, y, train=train, update_moving=update_moving, **kw)

        self._g_cond = cond          # g_backward forms the head's weight gradient from it
        K.dense_f32(cond, gl["ms"].w.view(2 * self.ce, self.E), gl["ms"].b, g["ms"], act=K.ACT_LRELU)   # :113-114, fp32
        if not cond_noise:
            tn_eps = torch.zeros_like(tn_eps)
        K.ca_fwd(g["ms"], z, tn_eps, g["zc"], kl_sum)                                                   # :117-122,174
        f0r = g["f0"].view(self.np, self.B, -1)
        K.conv_gemm(S1, 1, 0, V(g["zc"]), gl["fc0"].Wf, V(f0r), bias=gl["fc0"].b, **stats(0))            # :175
        bn(0, f0r, g["h0"].view(self.np, self.B, -1))                                                    # :176

        def conv(l, x, y, bn_i=None):
            K.conv_gemm(gl[l].mode, gl[l].k, 0, V(g[x]), gl[l].Wf, V(g[y]), bias=gl[l].b,
                        **(stats(bn_i) if bn_i is not None else {}))

        def res(x, c_a, t_a, bn_a, u_a, c_b, t_b, bn_b, u_b, c_c, t_c, bn_c, out):
            conv(c_a, x, t_a, bn_a); bn(bn_a, g[t_a], g[u_a], relu=True)
            conv(c_b, u_a, t_b, bn_b); bn(bn_b, g[t_b], g[u_b], relu=True)
            conv(c_c, u_b, t_c, bn_c); bn(bn_c, g[t_c], g[out], residual=g[x], relu=True)

        res("h0", "c0", "t1", 1, "u1", "c1", "t2", 2, "u2", "c2", "t3", 3, "h1")                        # :184-191
        conv("t0", "h1", "d1"); conv("c3", "d1", "t4", 4); bn(4, g["t4"], g["h2"])                       # :194-196
        res("h2", "c4", "t5", 5, "u5", "c5", "t6", 6, "u6", "c6", "t7", 7, "h3")                        # :200-207
        conv("t1", "h3", "d2"); conv("c7", "d2", "t8", 8); bn(8, g["t8"], g["h4"], relu=True)            # :210-212
        conv("t2", "h4", "d3"); conv("c8", "d3", "t9", 9); bn(9, g["t9"], g["h5"], relu=True)            # :214-216
        # :218-221 in one kernel: transposed conv 128 -> 3 (overlap-add on chip) + bias -> u4, 3x3 conv 3 -> 3 + tanh
        K.deconv_img(V(g["h5"]), gl["t3"].Wf, g["u4"], bias3=gl["t3"].b, w9=gl["c9"].w, b9=gl["c9"].b, img=img_out)

    def g_backward(self, d_img, part=None):
        """Backward of g_forward given dLoss/d image (fp32 [B,64,64,3]); fills the g gradient buffer.
        part = 1: down to the first transposed conv (every gradient of the first all-reduce bucket is final, side stream
        joined); part = 2: the rest; None: all of it.
        Every input-gradient GEMM whose output is the gradient at a BatchNorm(+ReLU) output applies the ReLU
        derivative mask and accumulates that BatchNorm's two backward reductions (sum dy -> dbeta, sum dy * x)
        in its epilogue; bn_bwd_fused then needs one pass.  Bias gradients ride along the same way."""
        K, g, gl, V = self.K, self.g, self.gl, self.K.View
        S1, K4, DC = K.CONV_S1, K.CONV_K4S2, K.DECONV_K4S2
        B, np_ = self.B, self.np
        rows = self._rows
        RELU = K.MASK_RELU
        img = self.d["img"][:B]

        sync = self.sync_bn

        def bn_red(i, x_pre):
            # sum dy goes straight into the dbeta gradient; under sync_bn into the scratch that is all-reduced first
            return dict(stat_sum=self.bn_dsum[i] if sync else self.bn_dbeta[i], stat_dot=self.bn_dot[i], stat_x=V(x_pre))

        def bn_bwd(i, dy, x_pre, dx, conv_bias_grad, via_scratch=False, dot_normalised=False):
            """dy: gradient at the BatchNorm output (already masked); dx: gradient at its input = at the output
            of the conv in front, whose bias gradient is the per-channel sum of dx."""
            kw = {}
            if sync:
                self.allreduce(self.bn_bwd_sums[i])
                kw = dict(out_scale=1.0 / self.world, stat_rows=(x_pre[0].numel() // x_pre.shape[-1]) * self.world)
            if sync or via_scratch:
                K.bn_bwd_fused(dy, x_pre, self.bn_mean[i], self.bn_rstd[i], self.bn_gamma[i], self.bn_dot[i],
                               self.bn_dsum[i], self.bn_dgamma[i], dx, conv_bias_grad, dbeta_out=self.bn_dbeta[i],
                               dot_normalised=dot_normalised, **kw)
            else:
                K.bn_bwd_fused(dy, x_pre, self.bn_mean[i], self.bn_rstd[i], self.bn_gamma[i], self.bn_dot[i],
                               self.bn_dbeta[i], self.bn_dgamma[i], dx, conv_bias_grad)

        def conv_bwd(l, x, dy, dx, **epi):
            """weight gradient of layer l (input x, output gradient dy) and its input gradient -> dx (+ epilogue)."""
            L = gl[l]
            with self._side():
                K.wgrad_gemm(L.mode, L.k, V(g[x]), V(g[dy]), L.gw)
            mode = {S1: S1, DC: K4, K4: DC}[L.mode]
            K.conv_gemm(mode, L.k, 1 if L.k == 3 else 0, V(g[dy]), L.Wf, V(dx), w_kn=True, **epi)

        def relu_of(y_post):
            return dict(mask=V(g[y_post]), mask_kind=RELU)

        def res_bwd(x, c_a, t_a, bn_a, u_a, c_b, t_b, bn_b, u_b, c_c, t_c, bn_c, out, **last_epi):
            ds = g["d_" + out]         # gradient at the residual sum x + bn(t_c): ReLU mask already applied by its producer
            bn_bwd(bn_c, ds, g[t_c], g["d_" + t_c], gl[c_c].gb)
            conv_bwd(c_c, u_b, "d_" + t_c, g["d_" + u_b], **relu_of(u_b), **bn_red(bn_b, g[t_b]))
            bn_bwd(bn_b, g["d_" + u_b], g[t_b], g["d_" + t_b], gl[c_b].gb)
            conv_bwd(c_b, u_a, "d_" + t_b, g["d_" + u_a], **relu_of(u_a), **bn_red(bn_a, g[t_a]))
            bn_bwd(bn_a, g["d_" + u_a], g[t_a], g["d_" + t_a], gl[c_a].gb)
            conv_bwd(c_a, x, "d_" + t_a, g["d_" + x], add=V(ds), **last_epi)     # skip connection joins here

        if part == 2:
            return self._g_backward_tail(res_bwd, bn_bwd, V)
        K.conv3x3_c3_tanh_bwd(g["u4"], gl["c9"].w, img, d_img, g["d_u4"], None, None, gl["t3"].gb)     # the chain needs d_u4
        with self._side():
            K.conv3x3_c3_tanh_bwd(g["u4"], gl["c9"].w, img, d_img, None, gl["c9"].gw, gl["c9"].gb)
        # the 4x4/s2 patches of d_u4 are formed on chip by both kernels from its padded bf16 rows (no patch matrix in HBM)
        K.img_to_rows(g["d_u4"], g["d_u4_rows"])
        du4 = K.ImgPatches(g["d_u4_rows"], IMG)
        with self._side():
            K.wgrad_img(du4, V(g["h5"]), gl["t3"].gw, 2)
        K.conv_gemm(K4, 4, 0, du4, gl["t3"].Wf, V(g["d_h5"]), w_kn=True,
                    mask=V(g["h5"]), mask_kind=RELU, **bn_red(9, g["t9"]))
        bn_bwd(9, g["d_h5"], g["t9"], g["d_t9"], gl["c8"].gb)
        conv_bwd("c8", "d3", "d_t9", g["d_d3"], stat_sum=gl["t2"].gb)      # gradient at t2's output: its bias gradient
        conv_bwd("t2", "h4", "d_d3", g["d_h4"], **relu_of("h4"), **bn_red(8, g["t8"]))
        bn_bwd(8, g["d_h4"], g["t8"], g["d_t8"], gl["c7"].gb)
        conv_bwd("c7", "d2", "d_t8", g["d_d2"], stat_sum=gl["t1"].gb)
        conv_bwd("t1", "h3", "d_d2", g["d_h3"], **relu_of("h3"), **bn_red(7, g["t7"]))
        res_bwd("h2", "c4", "t5", 5, "u5", "c5", "t6", 6, "u6", "c6", "t7", 7, "h3", **bn_red(4, g["t4"]))
        bn_bwd(4, g["d_h2"], g["t4"], g["d_t4"], gl["c3"].gb)
        conv_bwd("c3", "d1", "d_t4", g["d_d1"], stat_sum=gl["t0"].gb)
        conv_bwd("t0", "h1", "d_d1", g["d_h1"], **relu_of("h1"), **bn_red(3, g["t3"]))
        if part == 1:
            self._join()
            return
        self._g_backward_tail(res_bwd, bn_bwd, V)

    def _g_backward_tail(self, res_bwd, bn_bwd, V):
        """the first residual block, BatchNorm 0, the dense layer and the conditioning head (second all-reduce bucket)"""
        K, g, gl = self.K, self.g, self.gl
        S1 = K.CONV_S1
        B, np_ = self.B, self.np
        res_bwd("h0", "c0", "t1", 1, "u1", "c1", "t2", 2, "u2", "c2", "t3", 3, "h1")
        # BatchNorm 0 normalises per FEATURE of the [B, 16*C8] dense output (model.py:176), not per channel of the
        # 4x4 map the conv above wrote, so its reductions stay a separate pass
        flat = lambda t: t.view(np_, B, -1)
        L = gl["fc0"]
        K.bn_bwd_reduce(flat(g["d_h0"]), flat(g["f0"]), self.bn_mean[0], self.bn_rstd[0], self.bn_dot[0], self.bn_dsum[0])
        bn_bwd(0, flat(g["d_h0"]), flat(g["f0"]), flat(g["d_f0"]), L.gb, via_scratch=True, dot_normalised=True)
        with self._side():
            K.wgrad_gemm(S1, 1, V(g["zc"]), V(flat(g["d_f0"])), L.gw)
        K.conv_gemm(S1, 1, 0, V(flat(g["d_f0"])), L.Wf, V(g["d_zc"]), w_kn=True)
        K.ca_bwd(g["ms"], g["d_zc"], g["tn"], g["d_ms"], self.Z, self.kl_coeff / (self.GB * self.ce))
        L = gl["ms"]
        K.colsum(V(g["d_ms"]), L.gb)
        K.to_planes(self._g_cond, g["cond"])
        K.wgrad_gemm(S1, 1, V(g["cond"]), V(g["d_ms"]), L.gw)
        self._join()

    # ------------------------------------------------------------------ discriminator
    def d_forward(self, s0, n, tangent=False, after=None):
        """models/wgancls/model.py:129-161 on samples [s0, s0+n) of the D buffers.  tangent=True
        propagates a tangent instead: no biases, LeakyReLU replaced by its saved derivative mask,
        written in place over the forward activations of those samples (no logit)."""
        K, d, dl = self.K, self.d, self.dl
        S1, K4 = K.CONV_S1, K.CONV_K4S2
        rows = self._rows
        df8 = 8 * self.df
        if self.comm_stream is not None and not torch.cuda.is_current_stream_capturing():
            self.join_comm()    # d_net weights may still be in flight on the communication stream

        def V(t, **kw):
            return K.View(t, s0, n, **kw)

        def cg(l, x, y, act=True, add=None):
            L = dl[l]
            kw = {}
            if tangent:
                if act:
                    kw = dict(mask=y, mask_kind=K.MASK_LRELU)
            else:
                kw = dict(bias=L.b, act=K.ACT_LRELU if act else K.ACT_NONE)
            K.conv_gemm(L.mode, L.k, 0, x, L.Wf, y, add=add, **kw)

        done = after if after is not None else (lambda buf: None)   # `buf` holds its final values for this pass
        # :135 from the images' padded bf16 rows (tangent pass: _d_body left the tangent image's rows in segment 3)
        if not tangent:
            K.img_to_rows(d["img"][s0:s0 + n], d["rows"][:, s0:s0 + n])
        done("img"); done("cond")
        cg("h0", K.ImgPatches(d["rows"][:, s0:s0 + n], IMG), V(d["a0"])); done("a0")
        cg("h1", V(d["a0"]), V(d["a1"])); done("a1")                                # :136
        cg("h2", V(d["a1"]), V(d["a2"])); done("a2")                                # :137
        cg("h3", V(d["a2"]), V(d["a3"]), act=False); done("a3")                     # :138
        cg("r1", V(d["a3"]), V(d["r1"])); done("r1")                                # :142
        cg("r2", V(d["r1"]), V(d["r2"])); done("r2")                                # :143
        cg("r3", V(d["r2"]), V(d["cat"], coff=0, c=df8), add=V(d["a3"]))            # :144-146
        cg("efc", V(d["cond"]), V(d["e"]))                                          # :150
        K.embed_tile(d["e"][:, s0:s0 + n], d["cat"][:, s0:s0 + n], df8); done("cat")   # :153-155
        cg("h5", V(d["cat"]), V(d["a5"])); done("a5")                               # :157
        cg("h6", V(d["a5"]), V(d["a6"])); done("a6")                                # :158
        if not tangent:
            K.dout_fwd(d["a6"][:, s0:s0 + n], dl["out"].w, dl["out"].b, d["logit"][s0:s0 + n])   # :160

    def d_backward(self, s0, n, seed, g0, gn, want_cond_grad, bias_n=0):
        """Seeded backward of d_forward through the inputs of every layer (no weight gradients).
        Samples [g0, g0+gn) additionally get dD/d image -> d['gx'] (and dD/d cond -> d['g2']).
        bias_n > 0: the first bias_n samples' output gradients are summed into the bias gradients of the
        layers by the epilogues that produce them (d_bias_grads covers the two that no GEMM writes)."""
        K, d, dl = self.K, self.d, self.dl
        S1, DC = K.CONV_S1, K.DECONV_K4S2
        rows = self._rows
        df8 = 8 * self.df
        LR = K.MASK_LRELU

        def V(t, **kw):
            return K.View(t, s0, n, **kw)

        def bias_of(l, c=0):    # this GEMM's output is the gradient at layer l's output: column sums = l's bias gradient
            return dict(stat_sum=dl[l].gb, stat_n=bias_n, stat_c=c) if bias_n > 0 else {}

        K.dout_bwd_data(d["a6"][:, s0:s0 + n], dl["out"].w, seed[s0:s0 + n], d["d_a6"][:, s0:s0 + n])
        KN = dict(w_kn=True)    # the packed forward weights, read as [contraction][output channel]
        K.conv_gemm(S1, 1, 0, V(d["d_a6"]), dl["h6"].Wf, V(d["d_a5"]), mask=V(d["a5"]), mask_kind=LR, **KN, **bias_of("h5"))
        K.conv_gemm(S1, 3, 1, V(d["d_a5"]), dl["h5"].Wf, V(d["d_cat"]), mask=V(d["cat"]), mask_kind=LR, **KN,
                    **bias_of("r3", df8))
        K.embed_reduce(d["d_cat"][:, s0:s0 + n], d["d_e"][:, s0:s0 + n], df8)
        K.conv_gemm(S1, 3, 1, V(d["d_cat"], coff=0, c=df8), dl["r3"].Wf, V(d["d_r2"]), mask=V(d["r2"]), mask_kind=LR, **KN,
                    **bias_of("r2"))
        K.conv_gemm(S1, 3, 1, V(d["d_r2"]), dl["r2"].Wf, V(d["d_r1"]), mask=V(d["r1"]), mask_kind=LR, **KN, **bias_of("r1"))
        K.conv_gemm(S1, 1, 0, V(d["d_r1"]), dl["r1"].Wf, V(d["d_a3"]), add=V(d["d_cat"], coff=0, c=df8), **KN,
                    **bias_of("h3"))
        K.conv_gemm(DC, 4, 0, V(d["d_a3"]), dl["h3"].Wf, V(d["d_a2"]), mask=V(d["a2"]), mask_kind=LR, **KN, **bias_of("h2"))
        K.conv_gemm(DC, 4, 0, V(d["d_a2"]), dl["h2"].Wf, V(d["d_a1"]), mask=V(d["a1"]), mask_kind=LR, **KN, **bias_of("h1"))
        K.conv_gemm(DC, 4, 0, V(d["d_a1"]), dl["h1"].Wf, V(d["d_a0"]), mask=V(d["a0"]), mask_kind=LR, **KN, **bias_of("h0"))
        if gn > 0:
            K.deconv_img(K.View(d["d_a0"], g0, gn), dl["h0"].Wf, d["gx"], w_kn=True)     # transposed conv, overlap-add on chip
            if want_cond_grad:
                K.conv_gemm(S1, 1, 0, K.View(d["d_e"], g0, gn), dl["efc"].Wf, K.View(d["d_cond"], 0, gn), **KN)
                K.from_planes(d["d_cond"], d["g2"])

    # weight-gradient jobs of d_net: name -> (buffer holding the layer input, buffer holding the output gradient)
    D_WGRAD = OrderedDict([("h0", ("img", "d_a0")), ("h1", ("a0", "d_a1")), ("h2", ("a1", "d_a2")), ("h3", ("a2", "d_a3")),
                           ("r1", ("a3", "d_r1")), ("r2", ("r1", "d_r2")), ("r3", ("r2", "d_cat")), ("efc", ("cond", "d_e")),
                           ("h5", ("cat", "d_a5")), ("h6", ("a5", "d_a6")), ("out", ("a6", None))])

    def d_wgrad_layer(self, l, n, n_bias):
        """Weight gradient of one d_net layer over samples [0, n): (layer input) x (seeded output gradient)."""
        K, d, dl = self.K, self.d, self.dl
        rows = self._rows
        x, dy = self.D_WGRAD[l]
        if l == "h0":
            K.wgrad_img(K.ImgPatches(d["rows"][:, :n], IMG), K.View(d["d_a0"], 0, n), dl["h0"].gw, 1)
        elif l == "out":
            K.dout_bwd_weight(d["a6"][:, :n], d["seed"][:n], dl["out"].gw, dl["out"].gb, n_bias)
        else:
            kw = dict(coff=0, c=8 * self.df) if l == "r3" else {}
            K.wgrad_gemm(dl[l].mode, dl[l].k, K.View(d[x], 0, n), K.View(d[dy], 0, n, **kw), dl[l].gw)

    def d_bias_grads(self, n_bias):
        """Bias gradients of the two d_net layers whose output gradient no GEMM epilogue produces (h6: from
        dout_bwd_data, efc: from embed_reduce): sums over samples [0, n_bias).  The others: d_backward(bias_n=)."""
        K, d, dl = self.K, self.d, self.dl
        K.colsum(K.View(d["d_a6"], 0, n_bias), dl["h6"].gb)
        K.colsum(K.View(d["d_e"], 0, n_bias), dl["efc"].gb)

    # ------------------------------------------------------------------ optimizer plumbing
    def _set_lr(self, net, lr, t):
        """lr_t of tf.train.AdamOptimizer, staged into device memory OUTSIDE any captured graph."""
        self.lr_ring.upload([lr * math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1_net[net] ** t)], self.lr_t[net])

    def _adam(self, net, lo=0, hi=None):
        """TF-form Adam over parameters [lo, hi) of the flat buffer (default: all of them) + the bf16 repack."""
        n = self.d_n if net == "d" else self.g_n
        hi = n if hi is None else hi
        assert lo % 8 == 0 and 0 <= lo <= hi <= n
        packed = self.packed[net]
        self.K.adam_tf(self.flat[net][lo:hi], self.grad[net][lo:hi], self.adam_m[net][lo:hi], self.adam_v[net][lo:hi],
                       self.lr_t[net], self.beta1_net[net], self.beta2, ADAM_EPS, 1.0,
                       None if packed is None else packed[:, lo:hi])

    def _reduce(self, net):
        if self.world > 1:
            self.allreduce(self.grad[net])

    # ------------------------------------------------------------------ the two runs of an iteration
    def load_feed(self, x=None, x_mismatch=None, cond=None, z=None, epsilon=None, tn_eps=None):
        """Stage inputs (fp32 host or device tensors) into the engine's device buffers."""
        B, d, g = self.B, self.d, self.g
        if x is not None or x_mismatch is not None:
            # The two image batches are 90 % of the feed bytes and are first needed AFTER the generator
            # forward of the D run: copy them on their own stream so that the transfer hides under it.
            cs = self.copy_stream if (self.copy_stream is not None and not torch.as_tensor(x if x is not None else x_mismatch).is_cuda) else None
            if cs is not None:
                # img[B:3B] is free once the last D run has read it (d_step records the event): the copy need not
                # queue behind the G run still executing, it travels under it
                if self._img_free is not None:
                    cs.wait_event(self._img_free)
                else:
                    cs.wait_stream(torch.cuda.current_stream())
            with (torch.cuda.stream(cs) if cs is not None else contextlib.nullcontext()):
                if x is not None:
                    d["img"][B:2 * B].copy_(x, non_blocking=True)
                if x_mismatch is not None:
                    d["img"][2 * B:3 * B].copy_(x_mismatch, non_blocking=True)
        if cond is not None:
            self.feed["cond"].copy_(cond, non_blocking=True)
        if z is not None:
            g["z"][:, :self.Z_tf].copy_(z, non_blocking=True)      # columns Z_tf..Z stay zero (padding)
        if epsilon is not None:
            self.feed["epsilon"].copy_(epsilon.reshape(-1), non_blocking=True)
        if tn_eps is not None:
            g["tn"].copy_(tn_eps, non_blocking=True)

    @contextlib.contextmanager
    def _side(self):
        """Enqueue the enclosed launches on the side stream, ordered after everything enqueued so far on
        the current stream: weight/bias gradients run beside the critical chain (input gradients, tangent
        pass) and fill its tail waves.  Captured graphs record the same fork/join dependencies."""
        if self.side_stream is None:
            yield
            return
        main = torch.cuda.current_stream()
        self.side_stream.wait_stream(main)
        with torch.cuda.stream(self.side_stream):
            yield

    def _join(self):
        if self.side_stream is not None:
            torch.cuda.current_stream().wait_stream(self.side_stream)

    @contextlib.contextmanager
    def _on_comm(self):
        if self.comm_stream is None:
            yield
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            yield

    def join_comm(self):
        """Make the current stream wait for the D update running on the communication stream."""
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def _run(self, name, body):
        """Launch a step body: eagerly, or (use_graphs) captured once into a CUDA graph and replayed.
        The body enqueues only kernels / memsets / the allreduce on the current stream; everything that
        varies between steps reaches it through device memory (feeds, kt, lr_t)."""
        if not self.use_graphs or self.K.__dict__.get("PROFILE") is not None:
            return body()
        st = self._graphs.setdefault(name, {"n": 0, "graph": None, "launches": 0})
        st["n"] += 1
        if st["graph"] is not None:
            self.replayed_launches += st["launches"]     # our kernels inside the replayed graph
            return st["graph"].replay()
        if st["n"] < 2:          # first call eager: one-time attribute setup, allocator warm-up
            return body()
        torch.cuda.synchronize()
        from . import _lib
        graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            body()
        st["launches"] = _lib.launch_count() - l0          # counted at capture, no work was executed
        self.captured_launches += st["launches"]
        st["graph"] = graph
        self.replayed_launches += st["launches"]
        graph.replay()

    def d_step(self, lr_d):
        """sess.run([D_optim, kt_optim, D_loss]) -- models/wgancls/trainer.py:97."""
        self.d_t += 1
        self.join_comm()                        # a previous D update must have landed (N_CRITIC > 1)
        self._set_lr("d", lr_d, self.d_t)
        self._run("d_a1", self._d_body_gen)
        if self.copy_stream is not None:        # the real / mismatching images arrive on the copy stream
            torch.cuda.current_stream().wait_stream(self.copy_stream)
        self._run("d_a", self._d_body_loss)
        if self.copy_stream is not None:        # the fed image segments may be overwritten from here on (load_feed)
            if self._img_free is None:
                self._img_free = torch.cuda.Event()
            self._img_free.record()
        # D_loss and the kt step need the four logit sums and the two penalties only: they are published here (own small
        # all-reduce of the sums), with the tangent pass and every weight gradient still to come -- a caller that
        # fetches D_loss stages the G run while the device works on.
        with self._on_comm():
            if self.world > 1:
                self.allreduce(self.sums["d"])
            self._run("d_b", self._d_tail_scalars)
            self._publish_scalars()
        self._run("d_a2", self._d_body_rest)
        # The gradient collective and Adam go to the communication stream: the G run's generator forward does not
        # depend on them and overlaps (it joins before its d_net forward).
        with self._on_comm():
            if self.world > 1:
                self.allreduce(self.grad["d"][:self.d_n])      # outside the graphs
            self._run("d_c", self._d_tail_adam)

    def _d_tail_scalars(self):
        self.K.d_scalars(self.sums["d"], self.kt, self.scalars, self.GB, GP_WEIGHT, KT_LR)  # :79-91,100

    def _d_tail_adam(self):
        self._adam("d")                                                                    # :94-97

    def _publish_scalars(self):
        """Copy the scalars vector to pinned host memory on the current stream and mark the point with an event:
        ``scalars_dict`` (the D_loss / G_loss fetch of sess.run) then waits for the losses only, not for the
        optimizer kernels queued behind them."""
        if self.dev.type != "cuda":
            return
        if self._scalars_host is None:
            self._scalars_host = torch.zeros(16, dtype=self.f32_dtype).pin_memory()
            self._scalars_event = torch.cuda.Event()
        self._scalars_host.copy_(self.scalars, non_blocking=True)
        self._scalars_event.record()
        self._scalars_published = True

    def _d_body_gen(self):
        g = self.g
        self.grad["d"].zero_()
        g["kl_scratch"].zero_()
        self.g_forward(g["z"], self.feed["cond"], g["tn"], self.d["img"][:self.B], g["kl_scratch"])   # model.py:48

    def _d_body(self):
        self._d_body_loss()
        self._d_body_rest()

    def _d_body_loss(self):
        """Everything D_loss depends on: the 4B forward, the seeded backward down to the images, the two penalties."""
        K, d, g, B = self.K, self.d, self.g, self.B
        S = 4 * B
        cond = self.feed["cond"]
        K.gp_interp(d["img"][:B], d["img"][B:2 * B], self.feed["epsilon"], d["img"][3 * B:])   # model.py:53
        for seg in range(4):
            K.to_planes(cond, d["cond"][:, seg * B:(seg + 1) * B])                        # cond_inp = cond (:54)
        self.d_forward(0, S)                                                             # model.py:49-55
        K.d_seeds(self.kt, d["seed"], B, 1.0 / self.GB)
        K.d_sums(d["logit"], B, self.sums["d"])
        self.d_backward(0, S, d["seed"], 3 * B, B, True, bias_n=3 * B)                   # tf.gradients, :63,68
        inv = 1.0 / self.GB
        K.gp_penalty(d["gx"], GP_WEIGHT, inv, d["slope"], d["coef"], self.sums["d"][4:5])       # :62-65
        K.gp_penalty(d["g2"], GP_WEIGHT, inv, d["slope2"], d["coef2"], self.sums["d"][5:6])     # :67-70

    def _d_body_rest(self):
        """The second-order term and every weight gradient of the D run."""
        K, d, g, B = self.K, self.d, self.g, self.B
        S = 4 * B
        # second-order term: tangent (coef * g) through d_net, in place over the x_hat segment
        K.img_to_rows(d["gx"], d["rows"][:, 3 * B:], d["coef"])
        K.to_planes(d["g2"], d["cond"][:, 3 * B:], d["coef2"])
        with self._side():
            self.d_bias_grads(3 * B)             # first-order only (the JVP does not depend on biases)
        consumer = {x: l for l, (x, dy) in self.D_WGRAD.items()}

        def wgrad_when_ready(buf):               # layer input final -> its merged weight gradient can start
            with self._side():
                self.d_wgrad_layer(consumer[buf], S, 3 * B)

        self.d_forward(3 * B, B, tangent=True, after=wgrad_when_ready)
        self._join()

    def g_step(self, lr_g):
        """sess.run([G_optim, G_loss]) -- models/wgancls/trainer.py:101."""
        self.g_t += 1
        self._set_lr("g", lr_g, self.g_t)
        self._run("g_a1", self._g_body_fwd)
        self.join_comm()                        # d_net's weights of this iteration are final from here on
        if not self.sync_bn:
            return self._g_step_early_loss()
        if self.g_buckets == 2 and self.world > 1 and self.g_split is not None and not self.sync_bn:
            self._run("g_a2", lambda: self._g_body(part=1))
            with self._on_comm():                  # bucket 1 travels under the rest of the backward pass
                self.allreduce(self.grad["g"][:self.g_split])
            self._run("g_a3", lambda: self.g_backward(self.d["gx"], part=2))
            self.allreduce(self.grad["g"][self.g_split:])
            self.join_comm()
        else:
            self._run("g_a2", self._g_body)
            self._reduce("g")
        self._run("g_b", self._g_tail_scalars)
        self._publish_scalars()
        self._run("g_c", self._g_tail_adam)

    def _g_step_early_loss(self):
        """The rest of the G run with the losses published as soon as they exist: G_loss needs d_net's logits of the
        fake batch and the KL sum only, both final after the d_net forward -- two thirds of the run are still to come
        (d_net backward, g_net backward, all-reduce, Adam).  A caller that fetches G_loss (sess.run([G_optim, G_loss]))
        gets it then and stages the next run's feeds while the device works on; the sums travel in their own small
        all-reduce, the gradient all-reduce leaves them out."""
        bucketed = self.g_buckets == 2 and self.world > 1 and self.g_split is not None
        self._run("g_a2", self._g_body_dfwd)
        with self._on_comm():
            if self.world > 1:
                self.allreduce(self.sums["g"])
            self._run("g_b", self._g_tail_scalars)
            self._publish_scalars()
        if bucketed:
            self._run("g_a3", lambda: self._g_body_bwd(part=1))
            with self._on_comm():                  # bucket 1 and its Adam step travel under the rest of the backward pass
                self.allreduce(self.grad["g"][:self.g_split])      # (which reads none of these layers' weights any more)
                self._run("g_c1", lambda: self._adam("g", 0, self.g_split))
            self._run("g_a4", lambda: self.g_backward(self.d["gx"], part=2))
            self.allreduce(self.grad["g"][self.g_split:self.g_n])
            self._run("g_c2", lambda: self._adam("g", self.g_split, self.g_n))
            self.join_comm()
            return
        self._run("g_a3", self._g_body_bwd)
        if self.world > 1:
            self.allreduce(self.grad["g"][:self.g_n])
        self.join_comm()
        self._run("g_c", self._g_tail_adam)

    def _g_body_dfwd(self):
        self.d_forward(0, self.B)
        self.K.g_sums(self.d["logit"], self.B, self.sums["g"])

    def _g_body_bwd(self, part=None):
        d, B = self.d, self.B
        self.d_backward(0, B, d["gseed"], 0, B, False)
        if part is None:          # (subclasses override g_backward without the bucket split)
            self.g_backward(d["gx"])
        else:
            self.g_backward(d["gx"], part)

    def _g_tail_scalars(self):
        self.K.g_scalars(self.sums["g"], self.scalars, self.GB, self.ce, self.kl_coeff)   # model.py:92

    def _g_tail_adam(self):
        self._adam("g")                                                                  # :103-106

    def _g_body_fwd(self):
        K, d, g, B = self.K, self.d, self.g, self.B
        cond = self.feed["cond"]
        self.grad["g"].zero_()
        # UPDATE_OPS (moving statistics, model.py:98,102) ride in the BatchNorm kernels of this forward
        self.g_forward(g["z"], cond, g["tn"], d["img"][:B], self.sums["g"][1:2], update_moving=True)
        K.to_planes(cond, d["cond"][:, :B])

    def _g_body(self, part=None):
        K, d, g, B = self.K, self.d, self.g, self.B
        self.d_forward(0, B)
        K.g_sums(d["logit"], B, self.sums["g"])
        self.d_backward(0, B, d["gseed"], 0, B, False)
        if part is None:          # (subclasses override g_backward without the bucket split)
            self.g_backward(d["gx"])
        else:
            self.g_backward(d["gx"], part)

    def sample(self, z, cond, tn_eps, out, cond_noise=True):
        """generator(z, cond, is_training=False) -- the sampler of model.py:57 (batch = engine batch)."""
        self.g["kl_scratch"].zero_()
        self.g_forward(z, cond, tn_eps, out, self.g["kl_scratch"], train=False, cond_noise=cond_noise)

    def scalars_dict(self):
        from ._lib import SCALARS
        if self._scalars_published:             # wait for the published copy only (see _publish_scalars)
            self._scalars_event.synchronize()
            vals = self._scalars_host.tolist()
            return {n: vals[i] for i, n in enumerate(SCALARS)}
        self.join_comm()
        vals = self.scalars.detach().cpu().tolist()
        return {n: vals[i] for i, n in enumerate(SCALARS)}
