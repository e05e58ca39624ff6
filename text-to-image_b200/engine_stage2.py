"""Host orchestration of one StackGAN stage-II iteration (models/stackgan/stageII) on the same CUDA kernels.

Graph (models/stackgan/stageII/model.py:40-60): z, embedding -> the FROZEN stage-I generator in training mode
(batch statistics, its own conditioning noise; engine_stage1.StageIEngine.g_forward) -> 64x64 image -> stage-II
generator (:134-201: encode 64 -> 16, concat the tiled conditioning vector, four residual blocks of 4x4 stride-1
convs, four transposed-conv + 3x3 stages up to 256x256, tanh) -> stage-II discriminator (:78-132: six stride-2
convs 256 -> 4, two 4x4 stride-1 convs, a residual branch added to ITSELF (:117), embedding concat, 1x1 conv,
logits).  Losses / optimizers as stage-I with label smoothing 0.95 (models/stackgan/stageII/trainer.py:20-59).

All 15 + 11 BatchNorms use the fused forms (statistics in the producing GEMM's epilogue, one normalise pass, backward
reductions in the input-gradient GEMM's epilogue, one input-gradient pass).  New pieces relative to stage-I: the
4x4 stride-1 SAME conv with TF's asymmetric padding (conv_gemm / wgrad_gemm mode CONV_S1, k = 4), the 3-channel ends
at 64x64 (3x3 patch matrix) and 256x256 (8 padded output channels + tanh), an output-channel window (w_n0) that
splits the gradient of the generator's concat buffer into its image part (ReLU mask, BatchNorm reductions) and its
conditioning part, and affine_scale = 2 for the doubled residual branch.
"""
import contextlib

import torch

from .engine import BN_DECAY, BN_EPS, Engine, Layer
from .engine_stage1 import SCALARS_S1, StageIEngine

REAL_LABEL = 0.95     # models/stackgan/stageII/trainer.py:27
G2, D2 = "stageII_g_net/", "stageII_d_net/"
RELU_ACT, LRELU_ACT = 1, 2     # bn_apply_train activation codes


class StageIIEngine(Engine):
    FC0_NCHW = False

    def __init__(self, K, device, batch, np_=1, z_dim=100, embed_dim=1024, ce=128, gf=128, df=64, d_beta1=0.5,
                 g_beta1=0.5, alpha=0.5, kl_coeff=2.0, world=1, allreduce=None, s1_gf=128, image=256, s1_engine=None,
                 **kw):
        assert gf % 32 == 0, "stage-II needs GF_DIM to be a multiple of 32 (GF_DIM / 4 channels at 256x256)"
        assert image == 256, "the reference's stage-II is defined for 256x256 output (model.py:79: s16 = size // 64)"
        self.alpha, self.image = alpha, image
        share = kw.get("share_from")
        # the frozen stage-I generator: parameters 'g_net/*', training-mode BatchNorm, moving statistics keep stepping.
        # s1_engine: a StageIEngine whose parameters are shared (the reference's stage-II graph contains the stage-I
        # model object it was constructed with, models/stackgan/stageII/model.py:8,50); its d_net is never run here.
        if share is None and s1_engine is not None:
            s1_gf, s1_df, s1_from = s1_engine.gf, s1_engine.df, s1_engine
        else:
            s1_df, s1_from = (8, None) if share is None else (share.s1.df, share.s1)
            s1_gf = s1_gf if share is None else share.s1.gf
        self.s1 = StageIEngine(K, device, batch, np_, z_dim, embed_dim, ce, s1_gf, s1_df, world=world, allreduce=allreduce,
                               share_from=s1_from,
                               **{k: v for k, v in kw.items() if k in ("act_dtype", "f32_dtype", "sync_bn")},
                               use_graphs=False, concurrent=False)
        super().__init__(K, device, batch, np_, z_dim, embed_dim, ce, gf, df, beta1=d_beta1, beta2=0.999,
                         kl_coeff=kl_coeff, world=world, allreduce=allreduce, beta1_g=g_beta1, **kw)

    # ------------------------------------------------------------------ parameters
    @staticmethod
    def _cn(i):
        return "Conv" + ("" if i == 0 else "_%d" % i)

    @staticmethod
    def _bnn(i):
        return "BatchNorm" + ("" if i == 0 else "_%d" % i)

    def _g_layers(self):
        """models/stackgan/stageII/model.py:134-201, variables of scope stageII_g_net in creation order."""
        K = self.K
        gf, ce, E = self.gf, self.ce, self.E
        S1, K4, DC = K.CONV_S1, K.CONV_K4S2, K.DECONV_K4S2
        g, L, cn = G2, Layer, self._cn
        C4 = 4 * gf
        layers = [
            L("e0", "col3_in", g + "Conv", g + "Conv", S1, 1, 1, gf, 32),                        # :135
            L("e1", "conv", g + "Conv_1", g + "Conv_1", K4, 4, 16, 2 * gf, gf),                  # :137
            L("e2", "conv", g + "Conv_2", g + "Conv_2", K4, 4, 16, C4, 2 * gf),                  # :140
            L("ms", "ms", (g + "dense", g + "dense_1"), (g + "dense", g + "dense_1"), S1, 1, 1, 2 * ce, E),   # :64-68
            L("p0", "conv", g + "Conv_3", g + "Conv_3", S1, 3, 9, C4, C4 + ce),                  # :191
        ]
        for r in range(4):                                                                       # :148-157 x 4
            layers.append(L("ra%d" % r, "conv", g + cn(4 + 2 * r), g + cn(4 + 2 * r), S1, 4, 16, C4, C4))
            layers.append(L("rb%d" % r, "conv", g + cn(5 + 2 * r), g + cn(5 + 2 * r), S1, 4, 16, C4, C4))
        ch = [C4, 2 * gf, gf, gf // 2, gf // 4]
        for u in range(4):                                                                       # :160-176
            tname = g + "Conv2d_transpose" + ("" if u == 0 else "_%d" % u)
            layers.append(L("t%d" % u, "deconv", tname, tname, DC, 4, 16, ch[u + 1], ch[u]))
            layers.append(L("u%d" % u, "conv", g + cn(12 + u), g + cn(12 + u), S1, 3, 9, ch[u + 1], ch[u + 1]))
        layers.append(L("out", "img_out", g + "Conv_16", g + "Conv_16", S1, 3, 9, 8, gf // 4))    # :178
        bn_ch = [2 * gf, C4, C4] + [C4] * 8 + [2 * gf, gf, gf // 2, gf // 4]
        return layers, bn_ch, [g + self._bnn(i) for i in range(15)]

    def _d_layers(self):
        """models/stackgan/stageII/model.py:78-132, scope stageII_d_net."""
        K = self.K
        df, ce, E = self.df, self.ce, self.E
        S1, K4 = K.CONV_S1, K.CONV_K4S2
        d, L, cn = D2, Layer, self._cn
        ch = [df, 2 * df, 4 * df, 8 * df, 16 * df, 32 * df]
        layers = [L("h0", "col_in", d + "Conv", d + "Conv", S1, 1, 1, df, 64)]                    # :83
        for i in range(5):                                                                       # :85-98
            layers.append(L("h%d" % (i + 1), "conv", d + cn(i + 1), d + cn(i + 1), K4, 4, 16, ch[i + 1], ch[i]))
        layers += [
            L("h6", "conv", d + "Conv_6", d + "Conv_6", S1, 4, 16, 16 * df, 32 * df),             # :100
            L("h7", "conv", d + "Conv_7", d + "Conv_7", S1, 4, 16, 8 * df, 16 * df),              # :103
            L("r1", "conv", d + "Conv_8", d + "Conv_8", S1, 1, 1, 2 * df, 8 * df),                # :108
            L("r2", "conv", d + "Conv_9", d + "Conv_9", S1, 3, 9, 2 * df, 2 * df),                # :111
            L("r3", "conv", d + "Conv_10", d + "Conv_10", S1, 3, 9, 8 * df, 2 * df),              # :114
            L("efc", "dense", d + "dense", d + "dense", S1, 1, 1, ce, E),                         # :122
            L("h9", "conv", d + "Conv_11", d + "Conv_11", S1, 1, 1, 8 * df, 8 * df + ce),         # :129
            L("out", "dout", d + "Conv_12", d + "Conv_12", None, 4, 1, 1, 16 * 8 * df, need_bwd=False),   # :132
        ]
        bn_ch = ch[1:] + [16 * df, 8 * df, 2 * df, 2 * df, 8 * df, 8 * df]
        return layers, bn_ch, [d + self._bnn(i) for i in range(11)]

    # stage-I generator variables ride along at the checkpoint boundary ('g_net/*', trainer.py:48-51)
    def set_params_tf(self, p):
        own = {k: v for k, v in p.items() if k.startswith(G2) or k.startswith(D2)}
        super().set_params_tf(own)
        s1 = {k: v for k, v in p.items() if k.startswith("g_net/")}
        if s1:
            cur = self.s1.get_params_tf()
            cur.update(s1)
            self.s1.set_params_tf(cur)

    def get_params_tf(self):
        out = super().get_params_tf()
        out.update({k: v for k, v in self.s1.get_params_tf().items() if k.startswith("g_net/")})
        return out

    # ------------------------------------------------------------------ buffers
    def _scratch(self, ch):
        f32 = dict(device=self.dev, dtype=self.f32_dtype)
        tot = sum(ch)
        buf = torch.zeros(4 * tot, **f32)
        fwd, bwd, off = [], [], 0
        for c in ch:
            fwd.append(buf[off:off + 2 * c]); off += 2 * c
        for c in ch:
            bwd.append(buf[off:off + 2 * c]); off += 2 * c
        return buf, fwd, bwd

    def _pair(self, store, name, *shape):
        store[name] = self._planes(*shape)
        store["d_" + name] = self._planes(*shape)

    def _build_g_buffers(self):
        B, gf, ce, E = self.B, self.gf, self.ce, self.E
        f32 = dict(device=self.dev, dtype=self.f32_dtype)
        C4 = 4 * gf
        g = self.g = {}
        g["img64"] = torch.zeros(B, 64, 64, 3, **f32)
        g["col"] = self._planes(B * 4096, 32)
        self._pair(g, "a0", B, 64, 64, gf)
        self._pair(g, "t1", B, 32, 32, 2 * gf); self._pair(g, "a1", B, 32, 32, 2 * gf)
        self._pair(g, "t2", B, 16, 16, C4)
        self._pair(g, "cat", B, 16, 16, C4 + ce)
        g["cond"] = self._planes(B, E)
        g["ms"] = torch.zeros(B, 2 * ce, **f32)            # [mean | log_sigma]: fp32 (feeds exp())
        g["d_ms"] = self._planes(B, 2 * ce)
        self._pair(g, "c", B, ce)
        self._pair(g, "t3", B, 16, 16, C4); self._pair(g, "r0", B, 16, 16, C4)
        for r in range(4):
            for n in ("ta", "ua", "tb"):
                self._pair(g, "%s%d" % (n, r), B, 16, 16, C4)
            self._pair(g, "r%d" % (r + 1), B, 16, 16, C4)
        ch = [C4, 2 * gf, gf, gf // 2, gf // 4]
        for u in range(4):
            s = 32 << u
            self._pair(g, "d%d" % u, B, s, s, ch[u + 1])      # transposed-conv output
            self._pair(g, "tu%d" % u, B, s, s, ch[u + 1])     # 3x3 conv output (pre-BatchNorm)
            self._pair(g, "u%d" % u, B, s, s, ch[u + 1])      # after BatchNorm + ReLU
        self._pair(g, "lg", B, 256, 256, 8)
        g["tn"] = torch.zeros(B, ce, **f32)       # stage-II conditioning noise
        g["tn1"] = torch.zeros(B, ce, **f32)      # stage-I conditioning noise
        g["z"] = torch.zeros(B, self.Z, **f32)
        g["z0"] = torch.zeros(B, 0, **f32)
        g["kl_scratch"] = torch.zeros(1, **f32)
        self.feed = {"cond": torch.zeros(B, E, **f32), "epsilon": torch.zeros(B, **f32)}
        self.gbn_scratch, self.gbn_fwd, self.gbn_bwd = self._scratch(self.bn_ch)

    def _build_d_buffers(self):
        B, df, ce, E = self.B, self.df, self.ce, self.E
        f32 = dict(device=self.dev, dtype=self.f32_dtype)
        d = self.d = {}
        S = self.image
        d["img"] = torch.zeros(3 * B, S, S, 3, **f32)       # [synthetic | real | mismatch]
        d["col0"] = self._planes(B * (S // 2) ** 2, 64)
        d["d_col0"] = self._planes(B * (S // 2) ** 2, 64)
        self._pair(d, "a0", B, S // 2, S // 2, df)
        ch = [df, 2 * df, 4 * df, 8 * df, 16 * df, 32 * df]
        for i in range(5):
            s = S >> (i + 2)
            self._pair(d, "t%d" % (i + 1), B, s, s, ch[i + 1]); self._pair(d, "a%d" % (i + 1), B, s, s, ch[i + 1])
        df8 = 8 * df
        for n, c in (("t6", 16 * df), ("a6", 16 * df), ("t7", df8), ("a7", df8), ("t8", 2 * df), ("a8", 2 * df), ("t9", 2 * df),
                     ("a9", 2 * df), ("t10", df8), ("cat", df8 + ce), ("t11", df8), ("a11", df8)):
            self._pair(d, n, B, 4, 4, c)
        d["cond"] = self._planes(B, E)
        d["e"] = self._planes(B, ce); d["d_e"] = self._planes(B, ce)
        d["logit"] = torch.zeros(3 * B, **f32)
        d["seed"] = torch.zeros(B, **f32)
        d["gx"] = torch.zeros(B, S, S, 3, **f32)
        self.dbn_scratch, self.dbn_fwd, self.dbn_bwd = self._scratch(self.dbn_ch)

    # ------------------------------------------------------------------ shared BatchNorm plumbing
    def _bnp(self, net):
        if net == "g":
            return (self.bn_ch, self.gbn_fwd, self.gbn_bwd, self.bn_gamma, self.bn_beta, self.bn_dgamma, self.bn_dbeta,
                    self.bn_mean, self.bn_rstd, self.bn_var, self.bn_mm, self.bn_mv)
        return (self.dbn_ch, self.dbn_fwd, self.dbn_bwd, self.dbn_gamma, self.dbn_beta, self.dbn_dgamma, self.dbn_dbeta,
                self.dbn_mean, self.dbn_rstd, self.dbn_var, self.dbn_mm, self.dbn_mv)

    def _conv_bn(self, net, layers, buf, l, x, t, i, train=True):
        """conv l: buf[x] -> buf[t] (+ the batch statistics of BatchNorm i in training mode)"""
        K, L = self.K, layers[l]
        ch, fwd = self._bnp(net)[:2]
        st = dict(stat_sum=fwd[i][:ch[i]], stat_sq=fwd[i][ch[i]:]) if train else {}
        K.conv_gemm(L.mode, L.k, 0, K.View(buf[x]), L.Wf, K.View(buf[t]), bias=L.b, **st)

    def _bn(self, net, buf, i, x, y, act, residual=None, y_pitch=0, affine_scale=1.0, train=True, update_moving=True):
        K = self.K
        ch, fwd, _, gamma, beta, _, _, mean, rstd, var, mm, mv = self._bnp(net)
        res = None if residual is None else buf[residual]
        if train:
            stat_rows = 0
            if self.sync_bn:      # whole-batch statistics: sum the per-rank sums (utils/ops.py:20-29)
                self.allreduce(fwd[i])
                stat_rows = (buf[x][0].numel() // buf[x].shape[-1]) * self.world
            K.bn_apply_train(buf[x], fwd[i], BN_EPS, gamma[i], beta[i], buf[y], mean[i], rstd[i], var[i], residual=res,
                             relu=act, moving=(mm[i], mv[i]) if update_moving else None, decay=BN_DECAY,
                             stat_rows=stat_rows, y_pitch=y_pitch, affine_scale=affine_scale)
        else:       # inference (sampler): moving statistics; never combined with y_pitch / affine_scale on this path
            assert y_pitch == 0 and affine_scale == 1.0 and act in (0, RELU_ACT)
            K.bn_apply(buf[x], mm[i], torch.rsqrt(mv[i] + BN_EPS), gamma[i], beta[i], buf[y], res, act == RELU_ACT)

    def _bn_red(self, net, buf, i, x_pre, **kw):
        ch, _, bwd = self._bnp(net)[:3]
        return dict(stat_sum=bwd[i][:ch[i]], stat_dot=bwd[i][ch[i]:], stat_x=self.K.View(buf[x_pre]), **kw)

    def _bn_bwd(self, net, buf, i, dy, x_pre, dx, bias_grad, dot_normalised=False, dy_pitch=0, affine_scale=1.0):
        ch, _, bwd, gamma, _, dgamma, dbeta, mean, rstd = self._bnp(net)[:9]
        c = ch[i]
        out_scale, kw = affine_scale, {}
        if self.sync_bn:      # reductions over the whole batch; 1 / world keeps the gradient all-reduce(sum) exact
            self.allreduce(bwd[i])
            out_scale = affine_scale / self.world
            kw = dict(stat_rows=(buf[x_pre][0].numel() // buf[x_pre].shape[-1]) * self.world)
        self.K.bn_bwd_fused(buf[dy], buf[x_pre], mean[i], rstd[i], gamma[i], bwd[i][c:], bwd[i][:c], dgamma[i], buf[dx],
                            bias_grad, dbeta_out=dbeta[i], out_scale=out_scale, dot_normalised=dot_normalised,
                            dy_pitch=dy_pitch, affine_scale=affine_scale, **kw)

    # ------------------------------------------------------------------ generator
    def g2_forward(self, img_out, kl_sum, train=True, cond_noise=True, update_moving=True, run_stage1=True):
        """stage-I generator (training-mode BatchNorm when train) -> stage-II generator; image -> img_out fp32 NHWC.
        run_stage1=False: the 64x64 input image is already in g['img64'] (eager ``generator(image, embed)`` calls)."""
        K, g, gl, V = self.K, self.g, self.gl, self.K.View
        S1 = K.CONV_S1
        rows = self._rows
        C4 = 4 * self.gf
        cbn = lambda l, x, t, i: self._conv_bn("g", gl, g, l, x, t, i, train)
        bn = lambda i, x, y, act, **kw: self._bn("g", g, i, x, y, act, train=train, update_moving=update_moving, **kw)
        # model.py:50-51: the stage-I generator runs inside the stage-II graph
        s1 = self.s1
        if run_stage1:
            s1.g["kl_scratch"].zero_()
            s1.g_forward(g["z"], self.feed["cond"], g["tn1"], g["img64"], s1.g["kl_scratch"], train=train,
                         cond_noise=cond_noise, update_moving=train and update_moving)
        if train:
            self.gbn_scratch.zero_()
        K.im2col_k3s1_c3(g["img64"], g["col"])
        K.conv_gemm(S1, 1, 0, V(g["col"]), gl["e0"].Wf, V(rows(g["a0"])), bias=gl["e0"].b, act=K.ACT_RELU,
                    algo_scale=27.0 / 32.0)                                                     # :135
        cbn("e1", "a0", "t1", 0); bn(0, "t1", "a1", RELU_ACT)                                     # :137-138
        cbn("e2", "a1", "t2", 1)                                                                 # :140
        if train:       # the encoded image is the leading part of the concat buffer (:187-189)
            bn(1, "t2", "cat", RELU_ACT, y_pitch=C4 + self.ce)
        else:
            self._bn("g", g, 1, "t2", "d_t2", RELU_ACT, train=False)       # contiguous scratch, then copied into the concat buffer
            g["cat"][..., :C4].copy_(g["d_t2"])
        K.dense_f32(self.feed["cond"], gl["ms"].w.view(2 * self.ce, self.E), gl["ms"].b, g["ms"], act=K.ACT_LRELU)   # :64-68, fp32
        tn = g["tn"] if cond_noise else torch.zeros_like(g["tn"])
        K.ca_fwd(g["ms"], g["z0"], tn, g["c"], kl_sum)                                           # :71-76
        K.embed_tile(g["c"], g["cat"], C4)
        cbn("p0", "cat", "t3", 2); bn(2, "t3", "r0", RELU_ACT)                                    # :191-192
        for r in range(4):                                                                       # :194-197
            rin, rout = "r%d" % r, "r%d" % (r + 1)
            cbn("ra%d" % r, rin, "ta%d" % r, 3 + 2 * r); bn(3 + 2 * r, "ta%d" % r, "ua%d" % r, RELU_ACT)
            cbn("rb%d" % r, "ua%d" % r, "tb%d" % r, 4 + 2 * r)
            bn(4 + 2 * r, "tb%d" % r, rout, RELU_ACT, residual=rin)
        x = "r4"
        for u in range(4):                                                                       # :160-176
            L = gl["t%d" % u]
            K.conv_gemm(L.mode, L.k, 0, V(g[x]), L.Wf, V(g["d%d" % u]), bias=L.b)
            cbn("u%d" % u, "d%d" % u, "tu%d" % u, 11 + u); bn(11 + u, "tu%d" % u, "u%d" % u, RELU_ACT)
            x = "u%d" % u
        L = gl["out"]
        K.conv_gemm(S1, 3, 0, V(g["u3"]), L.Wf, V(g["lg"]), bias=L.b)                             # :178
        K.tanh_c3_fwd(g["lg"], img_out)

    def g2_backward(self, img, d_img):
        """Backward of the stage-II generator from dLoss/d image (fp32 [B,256,256,3]); the stage-I generator is frozen."""
        K, g, gl, V = self.K, self.g, self.gl, self.K.View
        S1, K4, DC = K.CONV_S1, K.CONV_K4S2, K.DECONV_K4S2
        rows = self._rows
        C4, ce = 4 * self.gf, self.ce
        RELU = K.MASK_RELU
        red = lambda i, x_pre, **kw: self._bn_red("g", g, i, x_pre, **kw)
        bnb = lambda i, dy, x_pre, dx, l, **kw: self._bn_bwd("g", g, i, dy, x_pre, dx, gl[l].gb, **kw)
        relu_of = lambda y: dict(mask=V(g[y]), mask_kind=RELU)

        def wgrad(l, x, dy, **kw):
            L = gl[l]
            with self._side():
                K.wgrad_gemm(L.mode, L.k, V(g[x]), V(g[dy]), L.gw, **kw)

        def dgrad(l, dy, dx, **epi):
            L = gl[l]
            mode = {S1: S1, DC: K4, K4: DC}[L.mode]
            K.conv_gemm(mode, L.k, 1 if L.k in (3, 4) and L.mode == S1 else 0, V(g[dy]), L.Wf, V(g[dx]) if isinstance(dx, str) else dx,
                        w_kn=True, **epi)

        K.tanh_c3_bwd(img, d_img, g["d_lg"])
        with self._side():
            K.colsum(V(g["d_lg"]), gl["out"].gb)
        wgrad("out", "u3", "d_lg")
        dgrad("out", "d_lg", "d_u3", **relu_of("u3"), **red(14, "tu3"))
        x_prev = ["r4", "u0", "u1", "u2"]
        for u in (3, 2, 1, 0):
            bnb(11 + u, "d_u%d" % u, "tu%d" % u, "d_tu%d" % u, "u%d" % u)
            wgrad("u%d" % u, "d%d" % u, "d_tu%d" % u)
            dgrad("u%d" % u, "d_tu%d" % u, "d_d%d" % u, stat_sum=gl["t%d" % u].gb)       # transposed conv's bias gradient
            wgrad("t%d" % u, x_prev[u], "d_d%d" % u)
            prev_bn, prev_pre = (11 + u - 1, "tu%d" % (u - 1)) if u > 0 else (10, "tb3")
            dgrad("t%d" % u, "d_d%d" % u, "d_" + x_prev[u], **relu_of(x_prev[u]), **red(prev_bn, prev_pre))
        for r in (3, 2, 1, 0):
            rin, rout = "r%d" % r, "r%d" % (r + 1)
            ds = "d_" + rout          # gradient at the residual sum, ReLU mask applied by its producer
            bnb(4 + 2 * r, ds, "tb%d" % r, "d_tb%d" % r, "rb%d" % r)
            wgrad("rb%d" % r, "ua%d" % r, "d_tb%d" % r)
            dgrad("rb%d" % r, "d_tb%d" % r, "d_ua%d" % r, **relu_of("ua%d" % r), **red(3 + 2 * r, "ta%d" % r))
            bnb(3 + 2 * r, "d_ua%d" % r, "ta%d" % r, "d_ta%d" % r, "ra%d" % r)
            wgrad("ra%d" % r, rin, "d_ta%d" % r)
            prev_bn, prev_pre = (4 + 2 * (r - 1), "tb%d" % (r - 1)) if r > 0 else (2, "t3")
            dgrad("ra%d" % r, "d_ta%d" % r, "d_" + rin, add=V(g[ds]), **relu_of(rin), **red(prev_bn, prev_pre))
        bnb(2, "d_r0", "t3", "d_t3", "p0")
        wgrad("p0", "cat", "d_t3")
        # gradient of the concat buffer in two output-channel windows: the encoded image (ReLU mask, BatchNorm 1's
        # reductions) and the tiled conditioning vector (no activation)
        dgrad("p0", "d_t3", V(g["d_cat"], coff=0, c=C4), mask=V(g["cat"], coff=0, c=C4), mask_kind=RELU, **red(1, "t2"))
        dgrad("p0", "d_t3", V(g["d_cat"], coff=C4, c=ce), w_n0=C4)
        K.embed_reduce(g["d_cat"], g["d_c"], C4)
        K.ca_bwd(g["ms"], g["d_c"], g["tn"], g["d_ms"], 0, self.kl_coeff / (self.GB * ce))
        L = gl["ms"]
        with self._side():
            K.colsum(V(g["d_ms"]), L.gb)
            K.to_planes(self.feed["cond"], g["cond"])
            K.wgrad_gemm(S1, 1, V(g["cond"]), V(g["d_ms"]), L.gw)
        bnb(1, "d_cat", "t2", "d_t2", "e2", dy_pitch=C4 + ce)
        wgrad("e2", "a1", "d_t2")
        dgrad("e2", "d_t2", "d_a1", **relu_of("a1"), **red(0, "t1"))
        bnb(0, "d_a1", "t1", "d_t1", "e1")
        wgrad("e1", "a0", "d_t1")
        dgrad("e1", "d_t1", "d_a0", **relu_of("a0"), stat_sum=gl["e0"].gb)
        with self._side():
            K.wgrad_gemm(S1, 1, V(g["col"]), V(rows(g["d_a0"])), gl["e0"].gw, algo_scale=27.0 / 32.0)
        self._join()

    # ------------------------------------------------------------------ discriminator, one call of B samples
    def d2_forward(self, k, update_moving=True):
        """models/stackgan/stageII/model.py:78-132 on images d['img'][k*B:(k+1)*B]; logits -> d['logit'][k*B:...]."""
        K, d, dl, V, B = self.K, self.d, self.dl, self.K.View, self.B
        S1 = K.CONV_S1
        df8 = 8 * self.df
        cbn = lambda l, x, t, i: self._conv_bn("d", dl, d, l, x, t, i)
        bn = lambda i, x, y, act, **kw: self._bn("d", d, i, x, y, act, update_moving=update_moving, **kw)
        K.im2col_k4s2_c3(d["img"][k * B:(k + 1) * B], d["col0"])
        K.conv_gemm(S1, 1, 0, V(d["col0"]), dl["h0"].Wf, V(self._rows(d["a0"])), bias=dl["h0"].b, act=K.ACT_LRELU,
                    algo_scale=0.75)                                                            # :83
        for i in range(1, 6):                                                                   # :85-98
            cbn("h%d" % i, "a%d" % (i - 1), "t%d" % i, i - 1); bn(i - 1, "t%d" % i, "a%d" % i, LRELU_ACT)
        cbn("h6", "a5", "t6", 5); bn(5, "t6", "a6", LRELU_ACT)                                   # :100-101
        cbn("h7", "a6", "t7", 6); bn(6, "t7", "a7", 0)                                           # :103-104
        cbn("r1", "a7", "t8", 7); bn(7, "t8", "a8", LRELU_ACT)                                   # :108-109
        cbn("r2", "a8", "t9", 8); bn(8, "t9", "a9", LRELU_ACT)                                   # :111-112
        cbn("r3", "a9", "t10", 9)                                                               # :114
        # :115-118 lrelu(net + net): the BatchNorm output doubled, written as the leading part of the concat buffer
        bn(9, "t10", "cat", LRELU_ACT, y_pitch=df8 + self.ce, affine_scale=2.0)
        K.conv_gemm(S1, 1, 0, V(d["cond"]), dl["efc"].Wf, V(d["e"]), bias=dl["efc"].b, act=K.ACT_LRELU)   # :122
        K.embed_tile(d["e"], d["cat"], df8)                                                     # :125-127
        cbn("h9", "cat", "t11", 10); bn(10, "t11", "a11", LRELU_ACT)                             # :129-130
        K.dout_fwd(d["a11"], dl["out"].w, dl["out"].b, d["logit"][k * B:(k + 1) * B])            # :132

    def d2_backward(self, want_wgrad, want_gx):
        K, d, dl, V, B = self.K, self.d, self.dl, self.K.View, self.B
        S1, K4, DC = K.CONV_S1, K.CONV_K4S2, K.DECONV_K4S2
        df8 = 8 * self.df
        LRM = K.MASK_LRELU
        red = lambda i, x_pre, **kw: self._bn_red("d", d, i, x_pre, **kw)
        lrelu_of = lambda y: dict(mask=V(d[y]), mask_kind=LRM)

        def bnb(i, dy, x_pre, dx, l, **kw):
            self._bn_bwd("d", d, i, dy, x_pre, dx, dl[l].gb if want_wgrad else None, **kw)

        def wgrad(l, x, dy, **kw):
            if want_wgrad:
                L = dl[l]
                with self._side():
                    K.wgrad_gemm(L.mode, L.k, V(d[x]), V(d[dy]), L.gw, **kw)

        def dgrad(l, dy, dx, **epi):
            L = dl[l]
            mode = {S1: S1, K4: DC}[L.mode]
            K.conv_gemm(mode, L.k, 1 if L.k in (3, 4) and L.mode == S1 else 0, V(d[dy]), L.Wf, V(d[dx]), w_kn=True, **epi)

        K.dout_bwd_data(d["a11"], dl["out"].w, d["seed"], d["d_a11"])       # includes lrelu'(a11)
        if want_wgrad:
            with self._side():
                K.dout_bwd_weight(d["a11"], d["seed"], dl["out"].gw, dl["out"].gb, B)
        c10 = self.dbn_ch[10]
        K.bn_bwd_reduce(d["d_a11"], d["t11"], self.dbn_mean[10], self.dbn_rstd[10], self.dbn_bwd[10][c10:], self.dbn_bwd[10][:c10])
        bnb(10, "d_a11", "t11", "d_t11", "h9", dot_normalised=True)
        wgrad("h9", "cat", "d_t11")
        dgrad("h9", "d_t11", "d_cat", **lrelu_of("cat"), **red(9, "t10", stat_c=df8))
        K.embed_reduce(d["d_cat"], d["d_e"], df8)
        if want_wgrad:
            with self._side():
                K.colsum(V(d["d_e"]), dl["efc"].gb)
                K.wgrad_gemm(S1, 1, V(d["cond"]), V(d["d_e"]), dl["efc"].gw)
        bnb(9, "d_cat", "t10", "d_t10", "r3", dy_pitch=df8 + self.ce, affine_scale=2.0)
        wgrad("r3", "a9", "d_t10")
        dgrad("r3", "d_t10", "d_a9", **lrelu_of("a9"), **red(8, "t9"))
        bnb(8, "d_a9", "t9", "d_t9", "r2")
        wgrad("r2", "a8", "d_t9")
        dgrad("r2", "d_t9", "d_a8", **lrelu_of("a8"), **red(7, "t8"))
        bnb(7, "d_a8", "t8", "d_t8", "r1")
        wgrad("r1", "a7", "d_t8")
        dgrad("r1", "d_t8", "d_a7", **red(6, "t7"))          # a7 has no activation; the skip input is dropped (:117)
        bnb(6, "d_a7", "t7", "d_t7", "h7")
        wgrad("h7", "a6", "d_t7")
        dgrad("h7", "d_t7", "d_a6", **lrelu_of("a6"), **red(5, "t6"))
        bnb(5, "d_a6", "t6", "d_t6", "h6")
        wgrad("h6", "a5", "d_t6")
        dgrad("h6", "d_t6", "d_a5", **lrelu_of("a5"), **red(4, "t5"))
        for i in (5, 4, 3, 2):
            bnb(i - 1, "d_a%d" % i, "t%d" % i, "d_t%d" % i, "h%d" % i)
            wgrad("h%d" % i, "a%d" % (i - 1), "d_t%d" % i)
            dgrad("h%d" % i, "d_t%d" % i, "d_a%d" % (i - 1), **lrelu_of("a%d" % (i - 1)), **red(i - 2, "t%d" % (i - 1)))
        bnb(0, "d_a1", "t1", "d_t1", "h1")
        wgrad("h1", "a0", "d_t1")
        bias0 = dict(stat_sum=dl["h0"].gb) if want_wgrad else {}
        dgrad("h1", "d_t1", "d_a0", **lrelu_of("a0"), **bias0)
        rows = self._rows
        if want_wgrad:
            with self._side():
                K.wgrad_gemm(S1, 1, V(d["col0"]), V(rows(d["d_a0"])), dl["h0"].gw, algo_scale=0.75)
        if want_gx:
            K.conv_gemm(S1, 1, 0, V(rows(d["d_a0"])), dl["h0"].Wf, V(d["d_col0"]), algo_scale=0.75, w_kn=True)
            K.col2im_k4s2_c3(d["d_col0"], d["gx"], None)

    # ------------------------------------------------------------------ the two runs of an iteration
    def load_feed(self, x=None, x_mismatch=None, cond=None, z=None, tn_eps=None, tn_s1=None, **_):
        B, d, g = self.B, self.d, self.g
        if x is not None or x_mismatch is not None:
            # 2 x B x 256 x 256 x 3 floats (100 MB at batch 64) that the D run needs only after its generator forward:
            # they travel on the copy stream, behind the last D run's reads of these segments (not behind the G run)
            first = x if x is not None else x_mismatch
            cs = self.copy_stream if (self.copy_stream is not None and not torch.as_tensor(first).is_cuda) else None
            if cs is not None:
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
            g["z"][:, :self.Z_tf].copy_(z, non_blocking=True)
        if tn_eps is not None:
            g["tn"].copy_(tn_eps, non_blocking=True)
        if tn_s1 is not None:
            g["tn1"].copy_(tn_s1, non_blocking=True)

    def d_step(self, lr):
        """sess.run([D_optim, D_loss, ...]) -- models/stackgan/stageII/trainer.py:131-132."""
        self.d_t += 1
        self._set_lr("d", lr, self.d_t)
        self._run("s2_d_gen", self._s2_d_gen)
        if self.copy_stream is not None:        # the real / mismatching images arrive on the copy stream
            torch.cuda.current_stream().wait_stream(self.copy_stream)
        self._run("s2_d", self._s2_d_rest)
        if self.copy_stream is not None:        # the fed image segments may be overwritten from here on (load_feed)
            if self._img_free is None:
                self._img_free = torch.cuda.Event()
            self._img_free.record()
        self._reduce("d")
        self._run("s2_d_tail", self._s2_d_tail)

    def _s2_d_body(self):
        self._s2_d_gen()
        self._s2_d_rest()

    def _s2_d_gen(self):
        g = self.g
        self.grad["d"].zero_()
        g["kl_scratch"].zero_()
        self.g2_forward(self.d["img"][:self.B], g["kl_scratch"])                                 # model.py:50-52

    def _s2_d_rest(self):
        K, d, g, B = self.K, self.d, self.g, self.B
        K.to_planes(self.feed["cond"], d["cond"])
        inv = 1.0 / self.GB
        for k, (label, weight) in enumerate(((0.0, 1.0 - self.alpha), (REAL_LABEL, 1.0), (0.0, self.alpha))):   # :53-56
            self.dbn_scratch.zero_()
            self.d2_forward(k)
            K.ce_seeds(d["logit"][k * B:(k + 1) * B], B, label, weight, inv, d["seed"], self.sums["d"][k:k + 1])
            self.d2_backward(want_wgrad=True, want_gx=False)
        self._join()

    def _s2_d_tail(self):
        self.K.s1_scalars(self.sums["d"], self.scalars, self.GB, self.ce, self.alpha, self.kl_coeff, 0)
        self._adam("d")

    def g_step(self, lr):
        """sess.run([G_optim, G_loss, ...]) -- models/stackgan/stageII/trainer.py:136-137."""
        self.g_t += 1
        self._set_lr("g", lr, self.g_t)
        self._run("s2_g", self._s2_g_body)
        self._reduce("g")
        self._run("s2_g_tail", self._s2_g_tail)

    def _s2_g_body(self):
        K, d, g, B = self.K, self.d, self.g, self.B
        self.grad["g"].zero_()
        self.g2_forward(d["img"][:B], self.sums["g"][4:5])
        K.to_planes(self.feed["cond"], d["cond"])
        self.dbn_scratch.zero_()
        self.d2_forward(0)
        K.ce_seeds(d["logit"][:B], B, 1.0, 1.0, 1.0 / self.GB, d["seed"], self.sums["g"][3:4])
        self.d2_backward(want_wgrad=False, want_gx=True)
        self.g2_backward(d["img"][:B], d["gx"])

    def _s2_g_tail(self):
        self.K.s1_scalars(self.sums["g"], self.scalars, self.GB, self.ce, self.alpha, self.kl_coeff, 1)
        self._adam("g")

    def scalars_dict(self):
        vals = self.scalars.detach().cpu().tolist()
        return {n: vals[i] for i, n in enumerate(SCALARS_S1)}

    def discriminator_logits(self, images, cond):
        d, B = self.d, self.B
        d["img"][:B].copy_(images)
        self.K.to_planes(cond, d["cond"])
        self.dbn_scratch.zero_()
        self.d2_forward(0, update_moving=False)
        return d["logit"][:B].clone()
