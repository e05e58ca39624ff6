"""Host orchestration of one StackGAN stage-I iteration (models/stackgan/stageI) on the same CUDA kernels.

The generator is layer for layer the wgancls generator (models/stackgan/stageI/model.py:114-171 vs
models/wgancls/model.py:163-225; NHWC reshape of the dense output, N(0, 0.02) init), so ``Engine.g_forward`` /
``Engine.g_backward`` are inherited unchanged.  What differs is the discriminator and the losses:

  * d_net has a BatchNorm after every conv but the first and the logits conv (model.py:81-112).  BatchNorm couples
    the samples of a call, so the three discriminator calls of the D run (synthetic, real/match, real/mismatch:
    model.py:45-49) cannot be batched into one pass as in wgancls: each call runs forward and backward on its own B
    samples, with its own statistics, and the parameter gradients accumulate over the calls.
  * sigmoid cross-entropy with one-sided label smoothing 0.9 (trainer.py:21-37), D_loss = real + alpha * mismatch
    + (1 - alpha) * synthetic, G_loss = CE(synthetic, 1) + KL_coeff * KL (trainer.py:40-44); no gradient penalty.
  * Adam(beta1 = D/G_BETA_DECAY, beta2 = 0.999) under ALL UPDATE_OPS (trainer.py:50-55): g_net's moving
    statistics step in BOTH runs; d_net's step once per discriminator call that is executed here (the reference also
    evaluates the real / mismatch calls in the G run only to step d_net's moving statistics, which nothing ever
    reads -- the discriminator is always built with is_training=True; that dead work is not reproduced).

Every BatchNorm uses the fused forms: batch statistics from the producing GEMM's epilogue, one normalise pass, the
backward reductions from the input-gradient GEMM's epilogue, one input-gradient pass (see engine.py).
"""
from collections import OrderedDict

import torch

from .engine import ADAM_EPS, BN_DECAY, BN_EPS, IMG, SUMS, Engine, Layer   # noqa: F401

REAL_LABEL = 0.9      # models/stackgan/stageI/trainer.py:26
SCALARS_S1 = ["D_loss", "D_synthetic_loss", "D_real_match_loss", "D_real_mismatch_loss", "G_loss", "G_gan_loss",
              "G_kl_loss"]


class StageIEngine(Engine):
    FC0_NCHW = False      # model.py:129 reshapes the dense output to [-1, 4, 4, C] (NHWC)

    def __init__(self, K, device, batch, np_=1, z_dim=100, embed_dim=1024, ce=128, gf=128, df=64, d_beta1=0.5,
                 g_beta1=0.5, alpha=0.5, kl_coeff=2.0, world=1, allreduce=None, **kw):
        self.alpha = alpha
        super().__init__(K, device, batch, np_, z_dim, embed_dim, ce, gf, df, beta1=d_beta1, beta2=0.999,
                         kl_coeff=kl_coeff, world=world, allreduce=allreduce, beta1_g=g_beta1, **kw)

    # ------------------------------------------------------------------ parameters
    def _d_layers(self):
        """models/stackgan/stageI/model.py:76-112."""
        K = self.K
        df, ce, E = self.df, self.ce, self.E
        S1, K4 = K.CONV_S1, K.CONV_K4S2
        d, L = "d_net/", Layer
        layers = [
            L("h0", "col_in", d + "Conv", d + "Conv", S1, 1, 1, df, 64),                      # :81
            L("h1", "conv", d + "Conv_1", d + "Conv_1", K4, 4, 16, 2 * df, df),                # :82
            L("h2", "conv", d + "Conv_2", d + "Conv_2", K4, 4, 16, 4 * df, 2 * df),            # :84
            L("h3", "conv", d + "Conv_3", d + "Conv_3", K4, 4, 16, 8 * df, 4 * df),            # :86
            L("r1", "conv", d + "Conv_4", d + "Conv_4", S1, 1, 1, 2 * df, 8 * df),             # :91
            L("r2", "conv", d + "Conv_5", d + "Conv_5", S1, 3, 9, 2 * df, 2 * df),             # :93
            L("r3", "conv", d + "Conv_6", d + "Conv_6", S1, 3, 9, 8 * df, 2 * df),             # :95
            L("efc", "dense", d + "dense", d + "dense", S1, 1, 1, ce, E),                      # :102
            L("h5", "conv", d + "Conv_7", d + "Conv_7", S1, 1, 1, 8 * df, 8 * df + ce),        # :109
            L("out", "dout", d + "Conv_8", d + "Conv_8", None, 4, 1, 1, 16 * 8 * df, need_bwd=False),   # :112
        ]
        bn_ch = [2 * df, 4 * df, 8 * df, 2 * df, 2 * df, 8 * df, 8 * df]
        bn_tf = [d + "BatchNorm" + ("" if i == 0 else "_%d" % i) for i in range(7)]
        return layers, bn_ch, bn_tf

    # ------------------------------------------------------------------ buffers
    def _build_d_buffers(self):
        B = self.B
        df, ce, E = self.df, self.ce, self.E
        f32 = dict(device=self.dev, dtype=self.f32_dtype)
        planes = self._planes
        d = self.d = {}
        d["img"] = torch.zeros(3 * B, IMG, IMG, 3, **f32)      # [synthetic | real | mismatch]
        d["col0"] = planes(B * 1024, 64)
        d["d_col0"] = planes(B * 1024, 64)
        df8 = 8 * df
        shapes = {"a0": (B, 32, 32, df), "t1": (B, 16, 16, 2 * df), "a1": (B, 16, 16, 2 * df), "t2": (B, 8, 8, 4 * df),
                  "a2": (B, 8, 8, 4 * df), "t3": (B, 4, 4, df8), "a3": (B, 4, 4, df8), "tr1": (B, 4, 4, 2 * df),
                  "r1": (B, 4, 4, 2 * df), "tr2": (B, 4, 4, 2 * df), "r2": (B, 4, 4, 2 * df), "tr3": (B, 4, 4, df8),
                  "cat": (B, 4, 4, df8 + ce), "t5": (B, 4, 4, df8), "a5": (B, 4, 4, df8)}
        for n, sh in shapes.items():
            d[n] = planes(*sh)
            d["d_" + n] = planes(*sh)
        d["cond"] = planes(B, E)
        d["e"] = planes(B, ce)
        d["d_e"] = planes(B, ce)
        d["logit"] = torch.zeros(3 * B, **f32)
        d["seed"] = torch.zeros(B, **f32)
        d["gx"] = torch.zeros(B, IMG, IMG, 3, **f32)
        # BatchNorm accumulators of one discriminator call (cleared per call): forward [sum x | sum x^2],
        # backward [sum dy | sum dy * x]
        tot = sum(self.dbn_ch)
        self.dbn_scratch = torch.zeros(4 * tot, **f32)
        self.dbn_fwd, self.dbn_bwd, off = [], [], 0
        for c in self.dbn_ch:
            self.dbn_fwd.append(self.dbn_scratch[off:off + 2 * c]); off += 2 * c
        for c in self.dbn_ch:
            self.dbn_bwd.append(self.dbn_scratch[off:off + 2 * c]); off += 2 * c

    # ------------------------------------------------------------------ discriminator, one call of B samples
    def d_forward_s1(self, k, update_moving=True):
        """models/stackgan/stageI/model.py:76-112 on images d['img'][k*B:(k+1)*B]; logits -> d['logit'][k*B:...]."""
        K, d, dl, V, B = self.K, self.d, self.dl, self.K.View, self.B
        S1 = K.CONV_S1
        rows = self._rows
        df8 = 8 * self.df
        LR = 2      # LeakyReLU(0.2) in bn_apply_train

        def cbn(l, x, t, i):      # conv + the batch statistics of the BatchNorm behind it
            L, c = dl[l], self.dbn_ch[i]
            K.conv_gemm(L.mode, L.k, 0, V(d[x]), L.Wf, V(d[t]), bias=L.b, stat_sum=self.dbn_fwd[i][:c],
                        stat_sq=self.dbn_fwd[i][c:])

        def bn(i, x, y, act, residual=None, y_pitch=0):
            stat_rows = 0
            if self.sync_bn:      # whole-batch statistics of this call: sum the per-rank sums (utils/ops.py:20-29)
                self.allreduce(self.dbn_fwd[i])
                stat_rows = (d[x][0].numel() // d[x].shape[-1]) * self.world
            K.bn_apply_train(d[x], self.dbn_fwd[i], BN_EPS, self.dbn_gamma[i], self.dbn_beta[i], d[y], self.dbn_mean[i],
                             self.dbn_rstd[i], self.dbn_var[i], residual=None if residual is None else d[residual],
                             relu=act, moving=(self.dbn_mm[i], self.dbn_mv[i]) if update_moving else None,
                             decay=BN_DECAY, stat_rows=stat_rows, y_pitch=y_pitch)

        K.im2col_k4s2_c3(d["img"][k * B:(k + 1) * B], d["col0"])
        K.conv_gemm(S1, 1, 0, V(d["col0"]), dl["h0"].Wf, V(rows(d["a0"])), bias=dl["h0"].b, act=K.ACT_LRELU,
                    algo_scale=0.75)                                                          # :81
        cbn("h1", "a0", "t1", 0); bn(0, "t1", "a1", LR)                                        # :82-83
        cbn("h2", "a1", "t2", 1); bn(1, "t2", "a2", LR)                                        # :84-85
        cbn("h3", "a2", "t3", 2); bn(2, "t3", "a3", 0)                                         # :86-87
        cbn("r1", "a3", "tr1", 3); bn(3, "tr1", "r1", LR)                                      # :91-92
        cbn("r2", "r1", "tr2", 4); bn(4, "tr2", "r2", LR)                                      # :93-94
        cbn("r3", "r2", "tr3", 5)                                                              # :95
        # lrelu(h3 + bn(.)) written as the leading channels of the concat buffer (:96-98,107)
        bn(5, "tr3", "cat", LR, residual="a3", y_pitch=df8 + self.ce)
        K.conv_gemm(S1, 1, 0, V(d["cond"]), dl["efc"].Wf, V(d["e"]), bias=dl["efc"].b, act=K.ACT_LRELU)   # :102
        K.embed_tile(d["e"], d["cat"], df8)                                                    # :105-107
        cbn("h5", "cat", "t5", 6); bn(6, "t5", "a5", LR)                                       # :109-110
        K.dout_fwd(d["a5"], dl["out"].w, dl["out"].b, d["logit"][k * B:(k + 1) * B])           # :112

    def d_backward_s1(self, want_wgrad, want_gx):
        """Backward of the last d_forward_s1 from the per-sample seeds d['seed'] (dLoss/dlogit).  want_wgrad: weight,
        bias and BatchNorm gradients ACCUMULATE into the d gradient buffer (D run); want_gx: dLoss/d image -> d['gx']."""
        K, d, dl, V, B = self.K, self.d, self.dl, self.K.View, self.B
        S1, K4, DC = K.CONV_S1, K.CONV_K4S2, K.DECONV_K4S2
        rows = self._rows
        df8 = 8 * self.df
        LRM = K.MASK_LRELU

        def bn_red(i, x_pre, **kw):      # this call's sum dy / sum dy * x for BatchNorm i, from the GEMM epilogue
            c = self.dbn_ch[i]
            return dict(stat_sum=self.dbn_bwd[i][:c], stat_dot=self.dbn_bwd[i][c:], stat_x=V(d[x_pre]), **kw)

        def bn_bwd(i, dy, x_pre, dx, bias_of, dot_normalised=False, dy_pitch=0):
            c = self.dbn_ch[i]
            kw = {}
            if self.sync_bn:      # the two reductions run over the whole batch; every rank then holds the full dgamma /
                self.allreduce(self.dbn_bwd[i])      # dbeta, scaled by 1 / world for the gradient all-reduce(sum)
                kw = dict(out_scale=1.0 / self.world, stat_rows=(d[x_pre][0].numel() // d[x_pre].shape[-1]) * self.world)
            K.bn_bwd_fused(d[dy], d[x_pre], self.dbn_mean[i], self.dbn_rstd[i], self.dbn_gamma[i], self.dbn_bwd[i][c:],
                           self.dbn_bwd[i][:c], self.dbn_dgamma[i], d[dx], dl[bias_of].gb if want_wgrad else None,
                           dbeta_out=self.dbn_dbeta[i], dot_normalised=dot_normalised, dy_pitch=dy_pitch, **kw)

        def wgrad(l, x, dy, **kw):
            if want_wgrad:
                L = dl[l]
                with self._side():
                    K.wgrad_gemm(L.mode, L.k, V(d[x]), V(d[dy]), L.gw, **kw)

        def dgrad(l, dy, dx, **epi):
            L = dl[l]
            mode = {S1: S1, K4: DC}[L.mode]
            K.conv_gemm(mode, L.k, 1 if L.k == 3 else 0, V(d[dy]), L.Wf, V(d[dx]), w_kn=True, **epi)

        K.dout_bwd_data(d["a5"], dl["out"].w, d["seed"], d["d_a5"])           # includes lrelu'(a5)
        if want_wgrad:
            with self._side():
                K.dout_bwd_weight(d["a5"], d["seed"], dl["out"].gw, dl["out"].gb, B)
        c6 = self.dbn_ch[6]
        K.bn_bwd_reduce(d["d_a5"], d["t5"], self.dbn_mean[6], self.dbn_rstd[6], self.dbn_bwd[6][c6:], self.dbn_bwd[6][:c6])
        bn_bwd(6, "d_a5", "t5", "d_t5", "h5", dot_normalised=True)
        wgrad("h5", "cat", "d_t5")
        # gradient at the concat buffer: LeakyReLU derivative of both parts; its leading channels are the gradient
        # at (h3 + bn5(.)): BatchNorm 5's reductions ride along
        dgrad("h5", "d_t5", "d_cat", mask=V(d["cat"]), mask_kind=LRM, **bn_red(5, "tr3", stat_c=df8))
        K.embed_reduce(d["d_cat"], d["d_e"], df8)
        if want_wgrad:
            with self._side():
                K.colsum(V(d["d_e"]), dl["efc"].gb)
                K.wgrad_gemm(S1, 1, V(d["cond"]), V(d["d_e"]), dl["efc"].gw)
        bn_bwd(5, "d_cat", "tr3", "d_tr3", "r3", dy_pitch=df8 + self.ce)
        wgrad("r3", "r2", "d_tr3")
        dgrad("r3", "d_tr3", "d_r2", mask=V(d["r2"]), mask_kind=LRM, **bn_red(4, "tr2"))
        bn_bwd(4, "d_r2", "tr2", "d_tr2", "r2")
        wgrad("r2", "r1", "d_tr2")
        dgrad("r2", "d_tr2", "d_r1", mask=V(d["r1"]), mask_kind=LRM, **bn_red(3, "tr1"))
        bn_bwd(3, "d_r1", "tr1", "d_tr1", "r1")
        wgrad("r1", "a3", "d_tr1")
        dgrad("r1", "d_tr1", "d_a3", add=V(d["d_cat"], coff=0, c=df8), **bn_red(2, "t3"))     # skip connection joins
        bn_bwd(2, "d_a3", "t3", "d_t3", "h3")
        wgrad("h3", "a2", "d_t3")
        dgrad("h3", "d_t3", "d_a2", mask=V(d["a2"]), mask_kind=LRM, **bn_red(1, "t2"))
        bn_bwd(1, "d_a2", "t2", "d_t2", "h2")
        wgrad("h2", "a1", "d_t2")
        dgrad("h2", "d_t2", "d_a1", mask=V(d["a1"]), mask_kind=LRM, **bn_red(0, "t1"))
        bn_bwd(0, "d_a1", "t1", "d_t1", "h1")
        wgrad("h1", "a0", "d_t1")
        bias0 = dict(stat_sum=dl["h0"].gb) if want_wgrad else {}
        dgrad("h1", "d_t1", "d_a0", mask=V(d["a0"]), mask_kind=LRM, **bias0)
        if want_wgrad:
            with self._side():
                K.wgrad_gemm(S1, 1, V(d["col0"]), V(rows(d["d_a0"])), dl["h0"].gw, algo_scale=0.75)
        if want_gx:
            K.conv_gemm(S1, 1, 0, V(rows(d["d_a0"])), dl["h0"].Wf, V(d["d_col0"]), algo_scale=0.75, w_kn=True)
            K.col2im_k4s2_c3(d["d_col0"], d["gx"], None)

    # ------------------------------------------------------------------ the two runs of an iteration
    def d_step(self, lr):
        """sess.run([D_optim, D_loss, ...]) -- models/stackgan/stageI/trainer.py:139-140."""
        self.d_t += 1
        self.join_comm()
        self._set_lr("d", lr, self.d_t)
        if self.copy_stream is not None:        # the real / mismatching images arrive on the copy stream
            torch.cuda.current_stream().wait_stream(self.copy_stream)
        self._run("s1_d", self._s1_d_body)
        self._reduce("d")                       # outside the graphs
        self._run("s1_d_tail", self._s1_d_tail)

    def _s1_d_body(self):
        K, d, g, B = self.K, self.d, self.g, self.B
        self.grad["d"].zero_()
        g["kl_scratch"].zero_()
        self.g_forward(g["z"], self.feed["cond"], g["tn"], d["img"][:B], g["kl_scratch"], update_moving=True)   # model.py:44
        K.to_planes(self.feed["cond"], d["cond"])
        inv = 1.0 / self.GB
        # model.py:45-49 / trainer.py:21-29,40-43: (label, weight in D_loss) per call
        for k, (label, weight) in enumerate(((0.0, 1.0 - self.alpha), (REAL_LABEL, 1.0), (0.0, self.alpha))):
            self.dbn_scratch.zero_()
            self.d_forward_s1(k)
            K.ce_seeds(d["logit"][k * B:(k + 1) * B], B, label, weight, inv, d["seed"], self.sums["d"][k:k + 1])
            self.d_backward_s1(want_wgrad=True, want_gx=False)
        self._join()

    def _s1_d_tail(self):
        self.K.s1_scalars(self.sums["d"], self.scalars, self.GB, self.ce, self.alpha, self.kl_coeff, 0)
        self._adam("d")

    def g_step(self, lr):
        """sess.run([G_optim, G_loss, ...]) -- models/stackgan/stageI/trainer.py:144-145."""
        self.g_t += 1
        self._set_lr("g", lr, self.g_t)
        self._run("s1_g", self._s1_g_body)
        self._reduce("g")
        self._run("s1_g_tail", self._s1_g_tail)

    def _s1_g_body(self):
        K, d, g, B = self.K, self.d, self.g, self.B
        self.grad["g"].zero_()
        self.g_forward(g["z"], self.feed["cond"], g["tn"], d["img"][:B], self.sums["g"][4:5], update_moving=True)
        K.to_planes(self.feed["cond"], d["cond"])
        self.dbn_scratch.zero_()
        self.d_forward_s1(0)
        K.ce_seeds(d["logit"][:B], B, 1.0, 1.0, 1.0 / self.GB, d["seed"], self.sums["g"][3:4])    # trainer.py:32-34
        self.d_backward_s1(want_wgrad=False, want_gx=True)
        self.g_backward(d["gx"])

    def _s1_g_tail(self):
        self.K.s1_scalars(self.sums["g"], self.scalars, self.GB, self.ce, self.alpha, self.kl_coeff, 1)
        self._adam("g")

    def scalars_dict(self):
        self.join_comm()
        vals = self.scalars.detach().cpu().tolist()
        return {n: vals[i] for i, n in enumerate(SCALARS_S1)}

    def discriminator_logits(self, images, cond):
        """d_net on a batch of B images (training-mode BatchNorm, as the reference always builds it); no state kept."""
        d, B = self.d, self.B
        d["img"][:B].copy_(images)
        self.K.to_planes(cond, d["cond"])
        self.dbn_scratch.zero_()
        self.d_forward_s1(0, update_moving=False)
        return d["logit"][:B].clone()
