"""Host orchestration of one iteration of one stage of the conditional progressive-growing WGAN
(models/pggan/pggan.py of the reference) on the shared CUDA kernels.

Graph of (stage, trans) (pggan.py:49-84): generator :279-316 (dense -> layer_norm -> 4x4 map, two 3x3 convs with
layer_norm + ReLU per stage, nearest-neighbour upscale between stages, to_rgb = 2x2 SAME conv to 9 channels + ReLU +
1x1 conv to 3 channels, fade-in blend with the upscaled to_rgb of the previous stage while `trans`), discriminator
:251-277 (from_rgb = 1x1 conv + LeakyReLU, two 3x3 convs + LeakyReLU and a 2x2 average pool per stage, fade-in blend
with from_rgb of the pooled image, embedding concat, 3x3 conv, 4x4 VALID conv, dense logit), the two one-sided
gradient penalties with weight 200 and D_loss = -wdist - wdist2 + 200 (gp + gp2), G_loss = -D_fake + 5 KL (:94-109),
Adam(2e-6, 0, 0.99) (:111-112).

The discriminator has no normalisation and only piecewise-linear pieces (LeakyReLU, average pool, blend), so the
schedule of engine.py carries over unchanged: ONE forward over the 4B batch [fake | real | mismatch | x_hat], one
seeded input-gradient pass, the second-order term of the penalties as the weight gradient of a tangent pass through the
x_hat segment (saved activation signs as masks, no biases, in place), one merged weight-gradient GEMM per layer.
With kt fixed at 1 the loss kernels of wgancls (d_seeds / d_sums / d_scalars / g_sums / g_scalars) produce exactly
these losses.  The generator's layer_norm is per sample (utils/ops.py:74-81), so data-parallel replicas need no
statistics exchange.

Channel padding: the 3-channel image enters / leaves as 8-channel planes (t2i_img_to_c8 / t2i_c8_to_img), to_rgb's
9 intermediate channels are stored as 16; the padded weight rows / columns are zero and stay zero (their gradients
are never exported and Adam leaves a zero-gradient weight at zero).
"""
from collections import OrderedDict

import torch

from .engine import Engine, HostScalarRing, Layer

GP_WEIGHT = 200.0      # pggan.py:108
KL_COEFF = 5.0         # pggan.py:109
LN_EPS = 1e-12         # tf.contrib.layers.layer_norm
ADAM_LR, ADAM_BETA1, ADAM_BETA2 = 0.000002, 0.0, 0.99      # pggan.py:111-112


class PgganEngine(Engine):
    FC0_NCHW = False       # pggan.py:290 reshapes the dense output to [-1, 4, 4, C] (NHWC)
    NORM_MOVING = False    # layer_norm has gamma / beta only

    def __init__(self, K, device, batch, np_=1, stage=1, trans=False, z_dim=128, embed_dim=1024, ce=128, nf_base=1024,
                 nf_cap=512, d_embed=128, rgb_mid=9, world=1, allreduce=None, **kw):
        assert stage >= 1 and not (trans and stage < 2), "a transition needs a previous stage"
        assert not kw.get("sync_bn")
        self.stage, self.trans = stage, bool(trans)
        self.nf_base, self.nf_cap, self.dce, self.rgb_mid = nf_base, nf_cap, d_embed, rgb_mid
        self.rgb_pad = (rgb_mid + 7) // 8 * 8
        self.S = 4 * 2 ** (stage - 1)         # pggan.py:30
        share = kw.get("share_from")
        if share is not None:
            assert (share.stage, share.trans, share.nf_base, share.nf_cap) == (stage, self.trans, nf_base, nf_cap)
        super().__init__(K, device, batch, np_, z_dim, embed_dim, ce, 8, 8, beta1=ADAM_BETA1, beta2=ADAM_BETA2,
                         kl_coeff=KL_COEFF, world=world, allreduce=allreduce, **kw)

    def nf(self, i):       # pggan.py:342-343
        return min(self.nf_base // (2 ** i) * 4, self.nf_cap)

    def dnf(self, i):      # pggan.py:339-340
        return min(self.nf_base // (2 ** i) * 2, self.nf_cap)

    def _rgb_stages(self):
        return sorted([self.stage - 1] + ([self.stage - 2] if self.trans else []))

    # ------------------------------------------------------------------ parameters
    def _g_layers(self):
        """pggan.py:279-316 and :367-371, variables of scope g_net."""
        K = self.K
        ce, E, Z = self.ce, self.E, self.Z
        S1, L = K.CONV_S1, Layer
        s0 = "g_net/conv_stage_0/"
        n0 = self.nf(0)
        layers = [
            L("ms", "ms", (s0 + "dense", s0 + "dense_1"), (s0 + "dense", s0 + "dense_1"), S1, 1, 1, 2 * ce, E),   # :284
            L("fc0", "fc0", s0 + "dense_2", s0 + "dense_2", S1, 1, 1, 16 * n0, Z + ce),                          # :288
            L("c0a", "conv", s0 + "Conv", s0 + "Conv", S1, 3, 9, n0, n0),                                        # :292
            L("c0b", "conv", s0 + "Conv_1", s0 + "Conv_1", S1, 3, 9, n0, n0),                                    # :294
        ]
        bn_ch = [16 * n0, n0, n0]
        bn_tf = [s0 + "LayerNorm", s0 + "LayerNorm_1", s0 + "LayerNorm_2"]
        for i in range(1, self.stage):                                                                          # :298-309
            s = "g_net/conv_stage_%d/" % i
            layers.append(L("c%da" % i, "conv", s + "Conv", s + "Conv", S1, 3, 9, self.nf(i), self.nf(i - 1)))
            layers.append(L("c%db" % i, "conv", s + "Conv_1", s + "Conv_1", S1, 3, 9, self.nf(i), self.nf(i)))
            bn_ch += [self.nf(i), self.nf(i)]
            bn_tf += [s + "LayerNorm", s + "LayerNorm_1"]
        for k in self._rgb_stages():                                                                            # :367-371
            s = "g_net/rgb_stage_%d/" % k
            a = L("rgb%d_a" % k, "conv_pad", s + "Conv", s + "Conv", S1, 2, 4, self.rgb_pad, self.nf(k))
            a.cin_tf, a.cout_tf = self.nf(k), self.rgb_mid
            b = L("rgb%d_b" % k, "conv_pad", s + "Conv_1", s + "Conv_1", S1, 1, 1, 8, self.rgb_pad)
            b.cin_tf, b.cout_tf = self.rgb_mid, 3
            layers += [a, b]
        return layers, bn_ch, bn_tf

    def _d_layers(self):
        """pggan.py:251-277 and :343-345, variables of scope d_net."""
        K = self.K
        S1, L = K.CONV_S1, Layer
        layers = []
        for k in self._rgb_stages():
            n = "d_net/rgb_stage_%d/Conv" % k
            r = L("rgb%d" % k, "conv_pad", n, n, S1, 1, 1, self.dnf(k), 8)
            r.cin_tf, r.cout_tf = 3, self.dnf(k)
            layers.append(r)
        for i in range(self.stage - 1, 0, -1):                                                                  # :261-267
            s = "d_net/conv_stage_%d/" % i
            layers.append(L("a%d" % i, "conv", s + "Conv", s + "Conv", S1, 3, 9, self.dnf(i), self.dnf(i)))
            layers.append(L("b%d" % i, "conv", s + "Conv_1", s + "Conv_1", S1, 3, 9, self.dnf(i - 1), self.dnf(i)))
        s0 = "d_net/conv_stage_0/"
        n0 = self.dnf(0)
        layers += [
            L("efc", "dense", s0 + "dense", s0 + "dense", S1, 1, 1, self.dce, self.E),                           # :271
            L("h0", "conv", s0 + "Conv", s0 + "Conv", S1, 3, 9, n0, n0 + self.dce),                              # :273
            L("h1", "flat4", s0 + "Conv_1", s0 + "Conv_1", S1, 1, 1, n0, 16 * n0),                               # :274
            L("out", "dout_fc", s0 + "dense_1", s0 + "dense_1", None, 1, 1, 1, n0, need_bwd=False),              # :275
        ]
        return layers, [], []

    def _build_params(self):
        super()._build_params()
        self.kt.fill_(1.0)       # D_loss = -wdist - wdist2 + ... (pggan.py:108): the wgancls loss kernels with kt == 1
        top = self.stage - 1
        w = OrderedDict()
        w["rgb%d" % top] = ("x8", "d_y%d" % top)
        if self.trans:
            w["rgb%d" % (top - 1)] = ("x8p", "d_iden")
        for i in range(top, 0, -1):
            w["a%d" % i] = ("y%d" % i, "d_p%d" % i)
            w["b%d" % i] = ("p%d" % i, "d_q%d" % i)
        w.update([("efc", ("cond", "d_e")), ("h0", ("cat", "d_a5")), ("h1", ("a5", "d_a6")), ("out", ("a6", None))])
        self.D_WGRAD = w

    # ------------------------------------------------------------------ buffers
    def _pair(self, store, name, *shape):
        store[name] = self._planes(*shape)
        store["d_" + name] = self._planes(*shape)

    def _build_d_buffers(self):
        B, S4, S = self.B, 4 * self.B, self.S
        E, dce, n0 = self.E, self.dce, self.dnf(0)
        f32 = dict(device=self.dev, dtype=self.f32_dtype)
        top = self.stage - 1
        d = self.d = {}
        # fade-in coefficients [alpha, 1 - alpha] in device memory (pggan.py:78-79), staged by set_alpha()
        self.ab = torch.tensor([1.0, 0.0], **f32)
        self.ab_ring = HostScalarRing(2, self.f32_dtype, self.dev.type == "cuda")
        d["img"] = torch.zeros(S4, S, S, 3, **f32)          # [fake | real | mismatch | x_hat]
        d["x8"] = self._planes(S4, S, S, 8)
        d["d_x8"] = self._planes(B, S, S, 8)
        d["gx"] = torch.zeros(B, S, S, 3, **f32)            # dD/d image
        if self.trans:
            d["x8p"] = self._planes(S4, S // 2, S // 2, 8)
            self._pair(d, "iden", S4, S // 2, S // 2, self.dnf(top - 1))
            self._pair(d, "pl", S4, S // 2, S // 2, self.dnf(top - 1))
            d["d_x8p"] = self._planes(B, S // 2, S // 2, 8)
            d["d_x8u"] = self._planes(B, S, S, 8)
        self._pair(d, "y%d" % top, S4, S, S, self.dnf(top))
        for i in range(top, 0, -1):
            r = 4 * 2 ** i
            self._pair(d, "p%d" % i, S4, r, r, self.dnf(i))
            self._pair(d, "q%d" % i, S4, r, r, self.dnf(i - 1))
            self._pair(d, "y%d" % (i - 1), S4, r // 2, r // 2, self.dnf(i - 1))
        d["cond"] = self._planes(S4, E)
        self._pair(d, "e", S4, dce)
        self._pair(d, "cat", S4, 4, 4, n0 + dce)
        self._pair(d, "a5", S4, 4, 4, n0)
        self._pair(d, "a6", S4, n0)
        d["logit"] = torch.zeros(S4, **f32)
        d["seed"] = torch.zeros(S4, **f32)
        d["gseed"] = torch.full((B,), -1.0 / self.GB, **f32)     # G_loss = -mean D(G) + ... (pggan.py:109)
        d["d_cond"] = self._planes(B, E)
        d["g2"] = torch.zeros(B, E, **f32)                       # dD/d cond
        for n in ("slope", "coef", "slope2", "coef2"):
            d[n] = torch.zeros(B, **f32)

    def _build_g_buffers(self):
        B, S = self.B, self.S
        ce, E, Z, n0 = self.ce, self.E, self.Z, self.nf(0)
        f32 = dict(device=self.dev, dtype=self.f32_dtype)
        top = self.stage - 1
        g = self.g = {}
        g["cond"] = self._planes(B, E)
        g["ms"] = torch.zeros(B, 2 * ce, **f32)            # [mean | log_sigma]: fp32 (feeds exp())
        g["d_ms"] = self._planes(B, 2 * ce)
        self._pair(g, "zc", B, Z + ce)
        self._pair(g, "f0", B, 16 * n0)
        for n in ("h0", "t0a", "u0a", "t0b", "x0"):
            self._pair(g, n, B, 4, 4, n0)
        for i in range(1, top + 1):
            r = 4 * 2 ** i
            self._pair(g, "up%d" % i, B, r, r, self.nf(i - 1))
            for n in ("t%da", "u%da", "t%db", "x%d"):
                self._pair(g, n % i, B, r, r, self.nf(i))
            g["d_xp%d" % (i - 1)] = self._planes(B, r // 2, r // 2, self.nf(i - 1))
        self._pair(g, "mN", B, S, S, self.rgb_pad)
        self._pair(g, "rgbN8", B, S, S, 8)
        if self.trans:
            self._pair(g, "mO", B, S // 2, S // 2, self.rgb_pad)
            self._pair(g, "rgbO8", B, S // 2, S // 2, 8)
            self._pair(g, "iden", B, S, S, 8)
            self._pair(g, "out8", B, S, S, 8)
        else:
            g["out8"], g["d_out8"] = g["rgbN8"], g["d_rgbN8"]
        g["tn"] = torch.zeros(B, ce, **f32)
        g["z"] = torch.zeros(B, Z, **f32)
        g["kl_scratch"] = torch.zeros(1, **f32)
        n_ln = len(self.bn_ch)
        self.ln_scratch = torch.zeros(2 * n_ln * B * 2, **f32)      # per layer_norm: [B][sum x | sum x^2], [B][sum g | sum g xhat]
        self.ln_fwd = [self.ln_scratch[i * 2 * B:(i + 1) * 2 * B].view(B, 2) for i in range(n_ln)]
        self.ln_bwd = [self.ln_scratch[(n_ln + i) * 2 * B:(n_ln + i + 1) * 2 * B].view(B, 2) for i in range(n_ln)]
        self.feed = {"cond": torch.zeros(B, E, **f32), "epsilon": torch.zeros(B, **f32)}

    def set_alpha(self, alpha):
        """alpha_tra = iter / steps (pggan.py:78-79), staged into device memory OUTSIDE any captured graph."""
        self.ab_ring.upload([float(alpha), 1.0 - float(alpha)], self.ab)

    # ------------------------------------------------------------------ generator
    def _ln(self, i, x, y, relu):
        K = self.K
        K.ln_stats(x, self.ln_fwd[i])
        K.ln_apply(x, self.ln_fwd[i], LN_EPS, self.bn_gamma[i], self.bn_beta[i], y, relu)

    def _ln_bwd(self, i, dy, x_pre, dx, bias_grad):
        """dy: gradient at the layer_norm output (activation derivative applied); dx: at its input = at the output of
        the conv / dense layer in front, whose bias gradient is the per-channel sum of dx."""
        K = self.K
        K.ln_bwd_reduce(dy, x_pre, self.ln_fwd[i], LN_EPS, self.bn_gamma[i], self.ln_bwd[i], self.bn_dgamma[i],
                        self.bn_dbeta[i])
        K.ln_bwd_apply(dy, x_pre, self.ln_fwd[i], LN_EPS, self.bn_gamma[i], self.ln_bwd[i], dx, bias_grad)

    def g_forward(self, z, cond, tn_eps, img_out, kl_sum, train=True, cond_noise=True, update_moving=False):
        """pggan.py:279-316.  z, cond, tn_eps: fp32 device tensors; image (no tanh) -> img_out [B, S, S, 3].
        layer_norm has no training / inference distinction (train / update_moving are accepted for the base schedule)."""
        K, g, gl, V = self.K, self.g, self.gl, self.K.View
        S1 = K.CONV_S1
        np_, B, top = self.np, self.B, self.stage - 1
        self.ln_scratch.zero_()
        self._g_cond = cond
        K.dense_f32(cond, gl["ms"].w.view(2 * self.ce, self.E), gl["ms"].b, g["ms"], act=K.ACT_LRELU)    # :343-347, fp32
        if not cond_noise:
            tn_eps = torch.zeros_like(tn_eps)
        K.ca_fwd(g["ms"], z, tn_eps, g["zc"], kl_sum)                                                    # :349-354,287
        K.conv_gemm(S1, 1, 0, V(g["zc"]), gl["fc0"].Wf, V(g["f0"]), bias=gl["fc0"].b)                     # :288
        self._ln(0, g["f0"], g["h0"].view(np_, B, -1), False)                                            # :289-290

        def conv(l, x, y):
            K.conv_gemm(S1, gl[l].k, 0, V(g[x]), gl[l].Wf, V(g[y]), bias=gl[l].b)

        def to_rgb(k, x, m, out):                                                                        # :367-371
            La, Lb = gl["rgb%d_a" % k], gl["rgb%d_b" % k]
            K.conv_gemm(S1, 2, 0, V(g[x]), La.Wf, V(g[m]), bias=La.b, act=K.ACT_RELU)
            K.conv_gemm(S1, 1, 0, V(g[m]), Lb.Wf, V(g[out]), bias=Lb.b)

        conv("c0a", "h0", "t0a"); self._ln(1, g["t0a"], g["u0a"], True)                                   # :292-293
        conv("c0b", "u0a", "t0b"); self._ln(2, g["t0b"], g["x0"], True)                                   # :294-295
        for i in range(1, top + 1):
            if i == top and self.trans:                                                                  # :300-302
                to_rgb(top - 1, "x%d" % (i - 1), "mO", "rgbO8")
                K.upscale2x(g["rgbO8"], g["iden"])
            K.upscale2x(g["x%d" % (i - 1)], g["up%d" % i])                                                # :305
            conv("c%da" % i, "up%d" % i, "t%da" % i); self._ln(2 * i + 1, g["t%da" % i], g["u%da" % i], True)
            conv("c%db" % i, "u%da" % i, "t%db" % i); self._ln(2 * i + 2, g["t%db" % i], g["x%d" % i], True)
        to_rgb(top, "x%d" % top, "mN", "rgbN8")                                                           # :311
        if self.trans:
            K.axpby(g["rgbN8"], g["iden"], g["out8"], self.ab)                                            # :313-314
        K.c8_to_img(g["out8"], img_out)

    def g_backward(self, d_img):
        """Backward of g_forward given dLoss/d image (fp32 [B, S, S, 3]); fills the g gradient buffer."""
        K, g, gl, V = self.K, self.g, self.gl, self.K.View
        S1, RELU = K.CONV_S1, K.MASK_RELU
        np_, B, top = self.np, self.B, self.stage - 1
        KN = dict(w_kn=True)

        def wgrad(l, x, dy):
            with self._side():
                K.wgrad_gemm(S1, gl[l].k, V(g[x]), V(g[dy]), gl[l].gw)

        def to_rgb_bwd(k, x, m, d_out, d_x, add):
            """x: the (post-ReLU) input of to_rgb; gradient at x (ReLU derivative applied, + add) -> d_x."""
            La, Lb = gl["rgb%d_a" % k], gl["rgb%d_b" % k]
            with self._side():
                K.colsum(V(g[d_out]), Lb.gb)
            wgrad("rgb%d_b" % k, m, d_out)
            K.conv_gemm(S1, 1, 0, V(g[d_out]), Lb.Wf, V(g["d_" + m]), mask=V(g[m]), mask_kind=RELU, stat_sum=La.gb, **KN)
            wgrad("rgb%d_a" % k, x, "d_" + m)
            K.conv_gemm(S1, 2, 1, V(g["d_" + m]), La.Wf, V(g[d_x]), add=add, mask=V(g[x]), mask_kind=RELU, **KN)

        def block_bwd(i, x_in, mask_in):
            """the two conv + layer_norm + ReLU pairs of stage i from d_x{i}; gradient at the block input -> d_<x_in>"""
            a, b = "%da" % i, "%db" % i
            self._ln_bwd(2 * i + 2, g["d_x%d" % i], g["t" + b], g["d_t" + b], gl["c" + b].gb)
            wgrad("c" + b, "u" + a, "d_t" + b)
            K.conv_gemm(S1, 3, 1, V(g["d_t" + b]), gl["c" + b].Wf, V(g["d_u" + a]), mask=V(g["u" + a]), mask_kind=RELU, **KN)
            self._ln_bwd(2 * i + 1, g["d_u" + a], g["t" + a], g["d_t" + a], gl["c" + a].gb)
            wgrad("c" + a, x_in, "d_t" + a)
            K.conv_gemm(S1, 3, 1, V(g["d_t" + a]), gl["c" + a].Wf, V(g["d_" + x_in]), **KN)

        K.img_to_c8(d_img, g["d_out8"])
        if self.trans:                                                                                  # :313-314
            K.axpby(g["d_out8"], None, g["d_rgbN8"], self.ab[0:1])
            K.axpby(g["d_out8"], None, g["d_iden"], self.ab[1:2])
        to_rgb_bwd(top, "x%d" % top, "mN", "d_rgbN8", "d_x%d" % top, None)
        for i in range(top, 0, -1):
            block_bwd(i, "up%d" % i, None)
            prev = "x%d" % (i - 1)
            K.pool2x(g["d_up%d" % i], g["d_xp%d" % (i - 1)], 1.0)          # transpose of the nearest-neighbour upscale
            if i == top and self.trans:
                K.pool2x(g["d_iden"], g["d_rgbO8"], 1.0)
                to_rgb_bwd(top - 1, prev, "mO", "d_rgbO8", "d_" + prev, V(g["d_xp%d" % (i - 1)]))
            else:
                K.act_bwd(g["d_xp%d" % (i - 1)], g[prev], g["d_" + prev], RELU)
        block_bwd(0, "h0", None)
        flat = lambda t: t.view(np_, B, -1)
        L = gl["fc0"]
        self._ln_bwd(0, flat(g["d_h0"]), g["f0"], g["d_f0"], L.gb)
        with self._side():
            K.wgrad_gemm(S1, 1, V(g["zc"]), V(g["d_f0"]), L.gw)
        K.conv_gemm(S1, 1, 0, V(g["d_f0"]), L.Wf, V(g["d_zc"]), **KN)
        K.ca_bwd(g["ms"], g["d_zc"], g["tn"], g["d_ms"], self.Z, self.kl_coeff / (self.GB * self.ce))
        L = gl["ms"]
        K.colsum(V(g["d_ms"]), L.gb)
        K.to_planes(self._g_cond, g["cond"])
        K.wgrad_gemm(S1, 1, V(g["cond"]), V(g["d_ms"]), L.gw)
        self._join()

    # ------------------------------------------------------------------ discriminator
    def d_forward(self, s0, n, tangent=False, after=None):
        """pggan.py:251-277 on samples [s0, s0 + n) of the D buffers.  tangent=True propagates a tangent instead: no
        biases, LeakyReLU replaced by its saved derivative mask, in place over the forward activations (no logit)."""
        K, d, dl = self.K, self.d, self.dl
        S1, LR = K.CONV_S1, K.MASK_LRELU
        top, n0 = self.stage - 1, self.dnf(0)
        if self.comm_stream is not None and not torch.cuda.is_current_stream_capturing():
            self.join_comm()    # d_net weights may still be in flight on the communication stream

        def sl(name):
            return d[name][:, s0:s0 + n]

        def V(name, **kw):
            return K.View(d[name], s0, n, **kw)

        def cg(l, x, y, act=True):
            L = dl[l]
            if tangent:
                kw = dict(mask=y, mask_kind=LR) if act else {}
            else:
                kw = dict(bias=L.b, act=K.ACT_LRELU if act else K.ACT_NONE)
            K.conv_gemm(S1, L.k, 0, x, L.Wf, y, **kw)

        done = after if after is not None else (lambda buf: None)   # `buf` holds its final values for this pass
        if not tangent:
            K.img_to_c8(d["img"][s0:s0 + n], sl("x8"))
        done("x8"); done("cond")
        if self.trans:                                                                 # :255-257
            K.pool2x(sl("x8"), sl("x8p"), 0.25); done("x8p")
            cg("rgb%d" % (top - 1), V("x8p"), V("iden"))
        cg("rgb%d" % top, V("x8"), V("y%d" % top)); done("y%d" % top)                   # :259
        for i in range(top, 0, -1):                                                    # :261-267
            cg("a%d" % i, V("y%d" % i), V("p%d" % i)); done("p%d" % i)
            cg("b%d" % i, V("p%d" % i), V("q%d" % i))
            if i == top and self.trans:
                K.pool2x(sl("q%d" % i), sl("pl"), 0.25)
                K.axpby(sl("pl"), sl("iden"), sl("y%d" % (i - 1)), self.ab)             # :268
            else:
                K.pool2x(sl("q%d" % i), sl("y%d" % (i - 1)), 0.25)
            done("y%d" % (i - 1))
        cg("efc", V("cond"), V("e"))                                                   # :271
        K.copy_window(sl("y0"), 0, sl("cat"), 0, n0)                                   # :272, :318-322
        K.embed_tile(sl("e"), sl("cat"), n0); done("cat")
        cg("h0", V("cat"), V("a5")); done("a5")                                        # :273
        a5f = d["a5"].view(self.np, d["a5"].shape[1], -1)
        cg("h1", K.View(a5f, s0, n), V("a6")); done("a6")                              # :274 (4x4 VALID on a 4x4 map)
        if not tangent:
            K.dout_fwd(sl("a6"), dl["out"].w, dl["out"].b, d["logit"][s0:s0 + n])      # :275

    def d_backward(self, s0, n, seed, g0, gn, want_cond_grad, bias_n=0):
        """Seeded backward of d_forward through the inputs of every layer (no weight gradients).  Samples
        [g0, g0 + gn) additionally get dD/d image -> d['gx'] (and dD/d cond -> d['g2']).  bias_n > 0: the first
        bias_n samples' output gradients are summed into the bias gradients by the epilogues that produce them
        (d_bias_grads covers the ones no GEMM writes)."""
        K, d, dl = self.K, self.d, self.dl
        S1, LR = K.CONV_S1, K.MASK_LRELU
        top, n0 = self.stage - 1, self.dnf(0)
        KN = dict(w_kn=True)

        def sl(name):
            return d[name][:, s0:s0 + n]

        def V(name, **kw):
            return K.View(d[name], s0, n, **kw)

        def bias_of(l):
            return dict(stat_sum=dl[l].gb, stat_n=bias_n) if bias_n > 0 else {}

        flat = lambda name: K.View(d[name].view(self.np, d[name].shape[1], -1), s0, n)
        K.dout_bwd_data(sl("a6"), dl["out"].w, seed[s0:s0 + n], sl("d_a6"))             # includes lrelu'(a6)
        K.conv_gemm(S1, 1, 0, V("d_a6"), dl["h1"].Wf, flat("d_a5"), mask=flat("a5"), mask_kind=LR, **KN)
        K.conv_gemm(S1, 3, 1, V("d_a5"), dl["h0"].Wf, V("d_cat"), **KN)                 # the image part is no activation output
        K.embed_reduce(sl("d_cat"), sl("d_e"), n0)
        K.act_bwd(sl("d_e"), sl("e"), sl("d_e"), LR)
        K.copy_window(sl("d_cat"), 0, sl("d_y0"), 0, n0)
        if top == 0:                                   # y0 is from_rgb's LeakyReLU output
            K.act_bwd(sl("d_y0"), sl("y0"), sl("d_y0"), LR)
        for i in range(1, top + 1):
            src = "d_y%d" % (i - 1)
            if i == top and self.trans:                                                # :268
                K.axpby(sl(src), None, sl("d_pl"), self.ab[0:1])
                K.axpby(sl(src), None, sl("d_iden"), self.ab[1:2])
                K.act_bwd(sl("d_iden"), sl("iden"), sl("d_iden"), LR)
                src = "d_pl"
            # transpose of the average pool and LeakyReLU' of the conv in front of it, one pass
            K.upscale2x(sl(src), sl("d_q%d" % i), 0.25, mask=sl("q%d" % i), mask_kind=LR)
            K.conv_gemm(S1, 3, 1, V("d_q%d" % i), dl["b%d" % i].Wf, V("d_p%d" % i), mask=V("p%d" % i), mask_kind=LR, **KN,
                        **bias_of("a%d" % i))
            epi = dict(mask=V("y%d" % i), mask_kind=LR, **bias_of("rgb%d" % top)) if i == top else {}
            K.conv_gemm(S1, 3, 1, V("d_p%d" % i), dl["a%d" % i].Wf, V("d_y%d" % i), **KN, **epi)
        if gn > 0:
            assert gn == self.B
            VG = lambda name: K.View(d[name], g0, gn)
            V0 = lambda name: K.View(d[name], 0, gn)
            add = {}
            if self.trans:
                K.conv_gemm(S1, 1, 0, VG("d_iden"), dl["rgb%d" % (top - 1)].Wf, V0("d_x8p"), **KN)
                K.upscale2x(d["d_x8p"], d["d_x8u"], 0.25)
                add = dict(add=V0("d_x8u"))
            K.conv_gemm(S1, 1, 0, VG("d_y%d" % top), dl["rgb%d" % top].Wf, V0("d_x8"), **KN, **add)
            K.c8_to_img(d["d_x8"], d["gx"])
            if want_cond_grad:
                K.conv_gemm(S1, 1, 0, VG("d_e"), dl["efc"].Wf, V0("d_cond"), **KN)
                K.from_planes(d["d_cond"], d["g2"])

    def d_wgrad_layer(self, l, n, n_bias):
        """Weight gradient of one d_net layer over samples [0, n): (layer input) x (seeded output gradient)."""
        K, d, dl = self.K, self.d, self.dl
        x, dy = self.D_WGRAD[l]
        if l == "out":
            K.dout_bwd_weight(d["a6"][:, :n], d["seed"][:n], dl["out"].gw, dl["out"].gb, n_bias)
        elif l == "h1":
            K.wgrad_gemm(K.CONV_S1, 1, K.View(d["a5"].view(self.np, d["a5"].shape[1], -1), 0, n), K.View(d["d_a6"], 0, n),
                         dl["h1"].gw)
        else:
            K.wgrad_gemm(K.CONV_S1, dl[l].k, K.View(d[x], 0, n), K.View(d[dy], 0, n), dl[l].gw)

    def d_bias_grads(self, n_bias):
        """Bias gradients of the d_net layers whose output gradient no GEMM epilogue produces (pooled / reduced /
        dot-product gradients): sums over samples [0, n_bias).  The others: d_backward(bias_n=)."""
        K, d, dl = self.K, self.d, self.dl
        top = self.stage - 1
        cs = lambda buf, l: K.colsum(K.View(d[buf], 0, n_bias), dl[l].gb)
        cs("d_a6", "h1"); cs("d_a5", "h0"); cs("d_e", "efc")
        for i in range(1, top + 1):
            cs("d_q%d" % i, "b%d" % i)
        if self.trans:
            cs("d_iden", "rgb%d" % (top - 1))
        if top == 0:
            cs("d_y0", "rgb0")

    # ------------------------------------------------------------------ the D run
    def _d_tail_scalars(self):
        self.K.d_scalars(self.sums["d"], self.kt, self.scalars, self.GB, GP_WEIGHT, 0.0)     # kt stays 1 (:94-108)

    def _d_body_loss(self):
        K, d, B = self.K, self.d, self.B
        S4 = 4 * B
        cond = self.feed["cond"]
        K.gp_interp(d["img"][:B], d["img"][B:2 * B], self.feed["epsilon"], d["img"][3 * B:])       # :68-69
        for seg in range(4):
            K.to_planes(cond, d["cond"][:, seg * B:(seg + 1) * B])                           # cond_inp = cond (:70)
        self.d_forward(0, S4)                                                               # :64-66,71
        K.d_seeds(self.kt, d["seed"], B, 1.0 / self.GB)
        K.d_sums(d["logit"], B, self.sums["d"])
        self.d_backward(0, S4, d["seed"], 3 * B, B, True, bias_n=3 * B)                      # tf.gradients, :86,91
        inv = 1.0 / self.GB
        K.gp_penalty(d["gx"], GP_WEIGHT, inv, d["slope"], d["coef"], self.sums["d"][4:5])          # :85-88
        K.gp_penalty(d["g2"], GP_WEIGHT, inv, d["slope2"], d["coef2"], self.sums["d"][5:6])        # :90-93

    def _d_body_rest(self):
        K, d, B = self.K, self.d, self.B
        S4 = 4 * B
        # second-order term: tangent (coef * g) through d_net, in place over the x_hat segment
        K.img_to_c8(d["gx"], d["x8"][:, 3 * B:], d["coef"])
        K.to_planes(d["g2"], d["cond"][:, 3 * B:], d["coef2"])
        with self._side():
            self.d_bias_grads(3 * B)             # first-order only (the JVP does not depend on biases)
        consumer = {x: l for l, (x, dy) in self.D_WGRAD.items()}

        def wgrad_when_ready(buf):               # layer input final -> its merged weight gradient can start
            if buf in consumer:
                with self._side():
                    self.d_wgrad_layer(consumer[buf], S4, 3 * B)

        self.d_forward(3 * B, B, tangent=True, after=wgrad_when_ready)
        self._join()

    def d_step(self, alpha=None, lr=ADAM_LR):
        """sess.run([D_optim, D_loss]) -- pggan.py:218; alpha = iter / steps is assigned first (:84,119)."""
        if alpha is not None:
            self.set_alpha(alpha)
        super().d_step(lr)

    def g_step(self, lr=ADAM_LR):
        """sess.run([G_optim, G_loss]) -- pggan.py:219."""
        super().g_step(lr)
