"""CPU restatement of the contract of every function in text-to-image_b200/kernels.py --
TEST INFRASTRUCTURE ONLY (the product has no CPU path and never imports this file).

Two uses: (1) ``-m gpu`` kernel unit tests run the real CUDA kernel and this restatement on the
same planes tensors and compare; (2) ``-m "not gpu"`` host-logic tests monkeypatch the engine's
kernel module with this one, so that the orchestration (layer order, tap bookkeeping, second-order
gradient-penalty schedule, optimizer, data-parallel sums) is checked against the oracle on CPU.
All arithmetic is fp64 on the planes' exact values; results are rounded back to planes.
"""
import torch
import torch.nn.functional as F

CONV_S1, CONV_K4S2, DECONV_K4S2 = 0, 1, 2
ACT_NONE, ACT_LRELU, ACT_RELU = 0, 1, 2
MASK_NONE, MASK_LRELU, MASK_RELU = 0, 1, 2
S = {n: i for i, n in enumerate(["D_loss", "D_loss_real", "D_loss_fake", "D_loss_mismatch", "wdist", "wdist2",
                                 "reg_loss", "balance_loss", "real_gp", "real_gp2", "kt", "kt_grad", "G_loss",
                                 "G_kl_loss"])}


def val(t):
    """exact value of a planes tensor [np, ...] in fp64"""
    return t.double().sum(0)


def put(t, v):
    """round v into the planes tensor t ([np, ...]); float64 "planes" store the value exactly"""
    if t.dtype == torch.float64:
        t.zero_()
        t[0].copy_(v.reshape(t.shape[1:]))
        return
    v = v.to(torch.float32).reshape(t.shape[1:])
    hi = v.to(torch.bfloat16)
    t[0].copy_(hi)
    if t.shape[0] == 2:
        t[1].copy_((v - hi.float()).to(torch.bfloat16))


class View:
    def __init__(self, t, n0=0, n=None, coff=0, c=None):
        if t.dim() == 3:
            self.N, self.H, self.W, self.pitch = t.shape[1], 1, 1, t.shape[2]
        else:
            self.N, self.H, self.W, self.pitch = t.shape[1:]
        self.t, self.np, self.n0 = t, t.shape[0], n0
        self.n = self.N - n0 if n is None else n
        self.coff = coff
        self.c = self.pitch - coff if c is None else c

    def _win(self):
        t = self.t.view(self.np, self.N, self.H, self.W, self.pitch)
        return t[:, self.n0:self.n0 + self.n, :, :, self.coff:self.coff + self.c]

    def values(self):
        return self._win().double().sum(0)

    def put(self, v):
        w = self._win()
        if self.t.dtype == torch.float64:
            w.zero_()
            w[0].copy_(v)
            return
        v = v.to(torch.float32)
        hi = v.to(torch.bfloat16)
        w[0].copy_(hi)
        if self.np == 2:
            w[1].copy_((v - hi.float()).to(torch.bfloat16))


def img_row_pitch(w):
    return (3 * w + 6 + 3) // 4 * 4


def img_to_rows(img, rows, sample_scale=None):
    """padded rows: entry j of a row = x[j - 3] (zeros outside), as planes [np, n, h, pitch]"""
    n, h, w, _ = img.shape
    x = img.double()
    if sample_scale is not None:
        x = x * sample_scale.double().view(-1, 1, 1, 1)
    v = torch.zeros(n, h, img_row_pitch(w), dtype=torch.float64)
    v[:, :, 3:3 + 3 * w] = x.reshape(n, h, 3 * w)
    put(rows, v)


class ImgPatches:
    """kernels.ImgPatches: the 4x4 / s2 SAME patch matrix of a 3-channel image given as its padded rows (planes)"""

    def __init__(self, rows, w):
        self.rows, self.np = rows, rows.shape[0]
        self.n, self.H, self.W, self.c = rows.shape[1], rows.shape[2], w, 48

    def values(self):
        """[n, h/2, w/2, 48], column (kh*4 + kw)*3 + c"""
        n, h, w = self.n, self.H, self.W
        x = val(self.rows)[:, :, 3:3 + 3 * w].reshape(n, h, w, 3)
        xp = F.pad(x, (0, 0, 1, 1, 1, 1))
        out = torch.zeros(n, h // 2, w // 2, 48, dtype=torch.float64)
        for kh in range(4):
            for kw in range(4):
                out[..., (kh * 4 + kw) * 3:(kh * 4 + kw) * 3 + 3] = xp[:, kh:kh + h:2, kw:kw + w:2, :]
        return out


def _conv_core(mode, k, flip, x, w):
    """x [N,H,W,Ci] fp64, w [taps, Co, Ci] fp64 -> y [N,OH,OW,Co]"""
    xc = x.permute(0, 3, 1, 2)
    taps, co, ci = w.shape
    if mode == CONV_S1:
        wk = w.view(k, k, co, ci).permute(2, 3, 0, 1)
        lo, hi = (k - 1) // 2, k - 1 - (k - 1) // 2      # TF SAME: k = 4 pads 1 before, 2 after
        if flip:
            wk = wk.flip(2, 3)
            lo, hi = hi, lo
        y = F.conv2d(F.pad(xc, (lo, hi, lo, hi)), wk)
    elif mode == CONV_K4S2:
        wk = w.view(4, 4, co, ci).permute(2, 3, 0, 1)
        y = F.conv2d(xc, wk, stride=2, padding=1)
    else:
        wk = w.view(4, 4, co, ci).permute(3, 2, 0, 1)  # [Ci, Co, kh, kw]
        y = F.conv_transpose2d(xc, wk, stride=2, padding=1)
    return y.permute(0, 2, 3, 1)


def conv_gemm(mode, k, flip, x, w, y, bias=None, add=None, mask=None, act=ACT_NONE, mask_kind=MASK_NONE,
              algo_scale=1.0, w_kn=False, stat_sum=None, stat_sq=None, stat_dot=None, stat_x=None, stat_n=0,
              stat_c=0, w_n0=0):
    xv = x.values()
    wv = val(w)
    if w_kn:
        wv = wv.transpose(1, 2)
    wv = wv[:, w_n0:w_n0 + y.c, :xv.shape[-1]]
    if isinstance(x, ImgPatches):      # one contraction block of 48 per output pixel
        v = torch.einsum("nhwk,ok->nhwo", xv, wv[0])
    else:
        v = _conv_core(mode, k, flip, xv, wv)
    if bias is not None:
        v = v + bias.double()[:y.c]
    if add is not None:
        v = v + add.values()
    if act == ACT_LRELU:
        v = torch.maximum(v, 0.2 * v)
    elif act == ACT_RELU:
        v = torch.relu(v)
    if mask is not None:
        m = mask.values()
        neg = 0.2 if mask_kind == MASK_LRELU else 0.0
        v = v * torch.where(m > 0, torch.ones_like(m), torch.full_like(m, neg))
    y.put(v)
    n_lim = stat_n if 0 < stat_n < v.shape[0] else v.shape[0]
    c_lim = stat_c if 0 < stat_c < v.shape[-1] else v.shape[-1]
    vs = v[:n_lim, ..., :c_lim].reshape(-1, c_lim)
    if stat_sum is not None:
        stat_sum[:c_lim] += vs.sum(0)
    if stat_sq is not None:
        stat_sq[:c_lim] += (vs * vs).sum(0)
    if stat_dot is not None:
        stat_dot[:c_lim] += (vs * stat_x.values()[:n_lim, ..., :c_lim].reshape(-1, c_lim)).sum(0)


def wgrad_gemm(mode, k, x, dy, dw, split_k=0, algo_scale=1.0):
    xv = x.values().requires_grad_(False)
    dyv = dy.values()
    taps, co, ci = dw.shape
    w = torch.zeros(taps, dyv.shape[-1], xv.shape[-1], dtype=torch.float64, requires_grad=True)
    y = _conv_core(mode, k, 0, xv, w)
    (g,) = torch.autograd.grad((y * dyv).sum(), [w])
    dw[:, :g.shape[1], :g.shape[2]] += g


def wgrad_img(img, other, dw, img_side):
    pv = img.values().reshape(-1, 48)
    ov = other.values().reshape(-1, other.c)
    if img_side == 1:
        dw[0, :other.c, :48] += ov.t() @ pv
    else:
        dw[0, :48, :other.c] += pv.t() @ ov


def deconv_img(a, w, out, bias3=None, w_kn=False, w9=None, b9=None, img=None):
    av = a.values()
    wv = val(w)[0]
    if w_kn:
        wv = wv.t()
    col = torch.einsum("nhwk,ok->nhwo", av, wv[:, :av.shape[-1]])        # [n, h, w, 64], 48 used
    n, h, wd, _ = av.shape
    acc = torch.zeros(n, 2 * h + 2, 2 * wd + 2, 3, dtype=torch.float64)
    for kh in range(4):
        for kw in range(4):
            acc[:, kh:kh + 2 * h:2, kw:kw + 2 * wd:2, :] += col[..., (kh * 4 + kw) * 3:(kh * 4 + kw) * 3 + 3]
    o = acc[:, 1:2 * h + 1, 1:2 * wd + 1, :]
    if bias3 is not None:
        o = o + bias3.double()
    out.copy_(o)
    if w9 is not None:
        wk = w9.double().view(3, 3, 3, 3).permute(3, 2, 0, 1)
        v = F.conv2d(out.double().permute(0, 3, 1, 2), wk, b9.double(), padding=1)
        img.copy_(torch.tanh(v).permute(0, 2, 3, 1))


def dense_f32(x, w, bias, y, act=ACT_NONE):
    v = x.double() @ w.double().t()
    if bias is not None:
        v = v + bias.double()
    if act == ACT_LRELU:
        v = torch.maximum(v, 0.2 * v)
    elif act == ACT_RELU:
        v = torch.relu(v)
    y.copy_(v)


def scale_rows(src, row_scale, dst):
    dst.copy_(src.double() * row_scale.double().view(-1, *([1] * (src.dim() - 1))))


def to_planes(src, dst, row_scale=None):
    v = src.double().reshape(src.shape[0], -1)
    if row_scale is not None:
        v = v * row_scale.double().reshape(-1, 1)
    put(dst, v.reshape(dst.shape[1:]))


def from_planes(src, dst):
    dst.copy_(val(src).reshape(dst.shape))


def im2col_k4s2_c3(img, col, sample_scale=None):
    n, h, w, _ = img.shape
    x = img.double()
    if sample_scale is not None:
        x = x * sample_scale.double().view(-1, 1, 1, 1)
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    out = torch.zeros(n, h // 2, w // 2, 64, dtype=torch.float64)
    for kh in range(4):
        for kw in range(4):
            out[..., (kh * 4 + kw) * 3:(kh * 4 + kw) * 3 + 3] = xp[:, kh:kh + h:2, kw:kw + w:2, :]
    put(col, out.reshape(-1, 64))


def col2im_k4s2_c3(col, img, bias3=None):
    n, h, w, _ = img.shape
    c = val(col).reshape(n, h // 2, w // 2, 64)
    acc = torch.zeros(n, h + 2, w + 2, 3, dtype=torch.float64)
    for kh in range(4):
        for kw in range(4):
            acc[:, kh:kh + h:2, kw:kw + w:2, :] += c[..., (kh * 4 + kw) * 3:(kh * 4 + kw) * 3 + 3]
    out = acc[:, 1:h + 1, 1:w + 1, :]
    if bias3 is not None:
        out = out + bias3.double()
    img.copy_(out)


def im2col_k3s1_c3(img, col):
    n, h, w, _ = img.shape
    xp = F.pad(img.double(), (0, 0, 1, 1, 1, 1))
    out = torch.zeros(n, h, w, 32, dtype=torch.float64)
    for kh in range(3):
        for kw in range(3):
            out[..., (kh * 3 + kw) * 3:(kh * 3 + kw) * 3 + 3] = xp[:, kh:kh + h, kw:kw + w, :]
    put(col, out.reshape(-1, 32))


def tanh_c3_fwd(logits8, img):
    img.copy_(torch.tanh(val(logits8)[..., :3]).reshape(img.shape))


def tanh_c3_bwd(img, dimg, dlogits8):
    v = torch.zeros(dlogits8.shape[1:], dtype=torch.float64)
    v[..., :3] = (dimg.double() * (1 - img.double() ** 2)).reshape(v[..., :3].shape)
    put(dlogits8, v)


def conv3x3_c3_tanh_fwd(x, w, b, y):
    wk = w.double().view(3, 3, 3, 3).permute(3, 2, 0, 1)
    v = F.conv2d(x.double().permute(0, 3, 1, 2), wk, b.double(), padding=1)
    y.copy_(torch.tanh(v).permute(0, 2, 3, 1))


def conv3x3_c3_tanh_bwd(x, w, y, dy, dx, dw, db, dx_sum=None):
    xd = x.double().requires_grad_(True)
    wd = w.double().view(3, 3, 3, 3).clone().requires_grad_(True)
    bd = torch.zeros(3, dtype=torch.float64, requires_grad=True)
    v = F.conv2d(xd.permute(0, 3, 1, 2), wd.permute(3, 2, 0, 1), bd, padding=1).permute(0, 2, 3, 1)
    dl = dy.double() * (1 - y.double() ** 2)
    gx, gw, gb = torch.autograd.grad((v * dl).sum(), [xd, wd, bd])
    if dx is not None:
        dx.copy_(gx)
        if dx_sum is not None:
            dx_sum += gx.sum((0, 1, 2))
    if dw is not None:
        dw += gw.reshape(dw.shape)
        db += gb


def colsum(src, out):
    out[:src.c] += src.values().reshape(-1, src.c).sum(0)


def bn_stats(x, sums, mean, rstd, var, eps):
    v = val(x).reshape(-1, x.shape[-1])
    m = v.mean(0)
    s = v.var(0, unbiased=False)
    mean.copy_(m); var.copy_(s); rstd.copy_(torch.rsqrt(s + eps))


def bn_apply(x, mean, rstd, gamma, beta, y, residual=None, relu=False):
    v = (val(x) - mean.double()) * rstd.double() * gamma.double() + beta.double()
    if residual is not None:
        v = v + val(residual)
    if relu:
        v = torch.relu(v)
    put(y, v)


def bn_apply_train(x, sums, eps, gamma, beta, y, mean, rstd, var, residual=None, relu=False, moving=None,
                   decay=0.9, stat_rows=0, y_pitch=0, affine_scale=1.0):
    c = x.shape[-1]
    rows = stat_rows if stat_rows > 0 else x[0].numel() // c
    m = sums[:c].double() / rows
    s = torch.clamp(sums[c:2 * c].double() / rows - m * m, min=0)
    mean.copy_(m); var.copy_(s); rstd.copy_(torch.rsqrt(s + eps))
    if moving is not None:
        bn_update_moving(moving[0], moving[1], mean, var, rows, decay)
    v = ((val(x) - mean.double()) * rstd.double() * gamma.double() + beta.double()) * affine_scale
    if residual is not None:
        v = v + val(residual)
    if relu:
        v = torch.maximum(v, (0.2 if int(relu) == 2 else 0.0) * v)
    if y_pitch in (0, c):
        put(y, v)
    else:       # y: the leading c channels of a [np, rows, y_pitch] window starting at y's first element
        yw = y.as_strided((y.shape[0], v.numel() // c, c), (y.stride(0), y_pitch, 1))
        put(yw, v.reshape(-1, c))


def bn_bwd_fused(dy, x, mean, rstd, gamma, dot, dbeta, dgamma, dx, dx_sum=None, dbeta_out=None, out_scale=1.0,
                 dot_normalised=False, stat_rows=0, dy_pitch=0, affine_scale=1.0):
    c = x.shape[-1]
    rows = stat_rows if stat_rows > 0 else x[0].numel() // c
    if dy_pitch not in (0, c):
        dy = dy.as_strided((dy.shape[0], x[0].numel() // c, c), (dy.stride(0), dy_pitch, 1))
    db = dbeta.double().clone()
    dg = dot.double().clone() if dot_normalised else rstd.double() * (dot.double() - mean.double() * db)
    dgamma += dg * out_scale
    if dbeta_out is not None:
        dbeta_out += db * out_scale
    g = val(dy).reshape(val(x).shape)
    xh = (val(x) - mean.double()) * rstd.double()
    out = gamma.double() * affine_scale * rstd.double() * (g - db / rows - xh * dg / rows)
    put(dx, out)
    if dx_sum is not None:
        dx_sum += out.reshape(-1, c).sum(0)


def bn_bwd_reduce(dy, x, mean, rstd, dgamma, dbeta):
    g = val(dy).reshape(-1, x.shape[-1])
    xh = (val(x).reshape(-1, x.shape[-1]) - mean.double()) * rstd.double()
    dbeta += g.sum(0)
    dgamma += (g * xh).sum(0)


def bn_bwd_apply(dy, x, mean, rstd, gamma, dgamma, dbeta, dx):
    c = x.shape[-1]
    rows = x[0].numel() // c
    g = val(dy)
    xh = (val(x) - mean.double()) * rstd.double()
    put(dx, gamma.double() * rstd.double() * (g - dbeta.double() / rows - xh * dgamma.double() / rows))


def bn_update_moving(mm, mv, mean, var, rows, decay):
    mm.copy_(decay * mm + (1 - decay) * mean)
    mv.copy_(decay * mv + (1 - decay) * var * (rows / max(rows - 1, 1)))


def act_bwd(dy, y, dst, mask_kind):
    a = val(y)
    neg = 0.2 if mask_kind == MASK_LRELU else 0.0
    put(dst, val(dy) * torch.where(a > 0, torch.ones_like(a), torch.full_like(a, neg)))


def embed_tile(e, cat, coff):
    c = e.shape[2]
    cat[..., coff:coff + c] = e[:, :, None, None, :].expand(-1, -1, cat.shape[2], cat.shape[3], -1)


def embed_reduce(dcat, de, coff):
    c = de.shape[2]
    put(de, val(dcat)[..., coff:coff + c].sum((1, 2)))


def dout_fwd(a, w, b, logit):
    s = a.shape[1]
    logit[:s] = (val(a).reshape(s, -1) @ w.double().reshape(-1) + b.double()[0])


def dout_bwd_data(a, w, seed, da):
    s = a.shape[1]
    av = val(a).reshape(s, -1)
    m = torch.where(av > 0, torch.ones_like(av), torch.full_like(av, 0.2))
    put(da, (seed.double()[:s, None] * w.double().reshape(1, -1) * m).reshape(da.shape[1:]))


def dout_bwd_weight(a, seed, dw, db, s_bias):
    s = a.shape[1]
    dw += (seed.double()[:s, None] * val(a).reshape(s, -1)).sum(0).reshape(dw.shape)
    if db is not None and s_bias > 0:
        db += seed[:s_bias].double().sum()


def gp_interp(g, x, eps, xhat):
    e = eps.reshape(-1, 1, 1, 1)
    xhat.copy_(e * g + (1 - e) * x)


def gp_penalty(grad, weight, inv_global_batch, slope, coef, pen_sum):
    n = grad.shape[0]
    s = grad.double().reshape(n, -1).pow(2).sum(1).sqrt()
    ex = torch.clamp(s - 1, min=0)
    slope[:n] = s
    coef[:n] = torch.where(ex > 0, weight * 2 * ex / s * inv_global_batch, torch.zeros_like(s))
    pen_sum += (ex * ex).sum()


def ca_fwd(ms, z, tn_eps, zc, kl_sum):
    ce = tn_eps.shape[1]
    m = ms.double()                       # fp32 tensor (dense_f32)
    mean, ls = m[:, :ce], m[:, ce:]
    c = mean + torch.exp(ls) * tn_eps.double()
    put(zc, torch.cat([z.double(), c], 1))
    if kl_sum is not None:
        kl_sum += (-ls + 0.5 * (-1 + torch.exp(2 * ls) + mean * mean)).sum()


def ca_bwd(ms, dzc, tn_eps, dms, z_dim, kl_scale):
    ce = tn_eps.shape[1]
    m = ms.double()
    mean, ls = m[:, :ce], m[:, ce:]
    dc = val(dzc)[:, z_dim:]
    dmean = (dc + kl_scale * mean) * torch.where(mean > 0, 1.0, 0.2)
    dls = (dc * tn_eps.double() * torch.exp(ls) + kl_scale * (torch.exp(2 * ls) - 1)) * torch.where(ls > 0, 1.0, 0.2)
    put(dms, torch.cat([dmean, dls], 1))


def d_seeds(kt, seed, b, inv_global_batch):
    k = float(kt[0])
    seed[0:b] = inv_global_batch
    seed[b:2 * b] = -(1 + k) * inv_global_batch
    seed[2 * b:3 * b] = k * inv_global_batch
    seed[3 * b:4 * b] = 1.0


def d_sums(logit, b, sums):
    sums[0] += logit[0:b].sum()
    sums[1] += logit[b:2 * b].sum()
    sums[2] += logit[2 * b:3 * b].sum()
    sums[3] += (logit[2 * b:3 * b] ** 2).sum()


def d_scalars(sums, kt, scalars, global_batch, gp_weight, kt_lr):
    s = sums.double() / global_batch
    fake, real, mis, reg, gp, gp2 = [float(v) for v in s[:6]]
    k = float(kt[0])
    wdist, wdist2 = real - fake, real - mis
    bal = k * wdist2 - wdist
    out = {"D_loss_real": real, "D_loss_fake": fake, "D_loss_mismatch": mis, "wdist": wdist, "wdist2": wdist2,
           "reg_loss": reg, "balance_loss": bal * bal, "real_gp": gp, "real_gp2": gp2,
           "D_loss": -wdist - k * wdist2 + gp_weight * (gp + gp2), "kt_grad": 2 * bal * wdist2}
    out["kt"] = k - kt_lr * out["kt_grad"]
    kt[0] = out["kt"]
    for n, v in out.items():
        scalars[S[n]] = v


def ce_seeds(logit, n, label, weight, inv_global_batch, seed, loss_sum):
    x = logit[:n].double()
    seed[:n] = weight * (torch.sigmoid(x) - label) * inv_global_batch
    if loss_sum is not None:
        loss_sum += (torch.clamp(x, min=0) - x * label + torch.log1p(torch.exp(-x.abs()))).sum()


def s1_scalars(sums, scalars, global_batch, ce, alpha, kl_coeff, which):
    s = sums.double()
    if which == 0:
        syn, real, mis = float(s[0]) / global_batch, float(s[1]) / global_batch, float(s[2]) / global_batch
        scalars[0] = real + alpha * mis + (1 - alpha) * syn
        scalars[1], scalars[2], scalars[3] = syn, real, mis
    else:
        gan, kl = float(s[3]) / global_batch, float(s[4]) / (global_batch * ce)
        scalars[4] = gan + kl_coeff * kl
        scalars[5], scalars[6] = gan, kl


def g_sums(logit_fake, b, sums):
    sums[0] += logit_fake[:b].sum()


def g_scalars(sums, scalars, global_batch, ce, kl_coeff):
    fake = float(sums[0]) / global_batch
    kl = float(sums[1]) / (global_batch * ce)
    scalars[S["G_kl_loss"]] = kl
    scalars[S["G_loss"]] = -fake + kl_coeff * kl


def pack_weight(w, fwd=None, bwd=None):
    if fwd is not None:
        put(fwd, w.double())
    if bwd is not None:
        put(bwd, w.double().transpose(1, 2))


def adam_tf(theta, grad, m, v, lr_t, beta1, beta2, eps=1e-8, grad_scale=1.0, packed=None):
    g = grad * grad_scale
    mj = beta1 * m + (1 - beta1) * g
    if beta1 != 0:
        m.copy_(mj)
    v.copy_(beta2 * v + (1 - beta2) * g * g)
    theta -= lr_t[0] * mj / (v.sqrt() + eps)
    if packed is not None:
        put(packed, theta.double())


# ---------------------------------------------------------------------------------------------------------------------
# conditional PGGAN kernels (text-to-image_b200/csrc/pggan_ops.cu)
def ln_stats(x, sums):
    v = val(x).reshape(x.shape[1], -1)
    sums[:, 0] += v.sum(1).to(sums.dtype)
    sums[:, 1] += (v * v).sum(1).to(sums.dtype)


def _ln_mr(x, sums, eps):
    n = x.shape[1]
    m = x[0, 0].numel()
    mean = sums[:, 0].double() / m
    var = (sums[:, 1].double() / m - mean * mean).clamp_min(0.0)
    shape = [n] + [1] * (x.dim() - 2)
    return mean.reshape(shape), torch.rsqrt(var + eps).reshape(shape), m


def ln_apply(x, sums, eps, gamma, beta, y, relu=False):
    mean, rstd, _ = _ln_mr(x, sums, eps)
    v = (val(x) - mean) * rstd * gamma.double() + beta.double()
    put(y, v.clamp_min(0.0) if relu else v)


def ln_bwd_reduce(dy, x, sums, eps, gamma, dsums, dgamma, dbeta):
    mean, rstd, _ = _ln_mr(x, sums, eps)
    xh = (val(x) - mean) * rstd
    d = val(dy)
    g = d * gamma.double()
    n, c = x.shape[1], x.shape[-1]
    dsums[:, 0] += g.reshape(n, -1).sum(1).to(dsums.dtype)
    dsums[:, 1] += (g * xh).reshape(n, -1).sum(1).to(dsums.dtype)
    dgamma += (d * xh).reshape(-1, c).sum(0).to(dgamma.dtype)
    dbeta += d.reshape(-1, c).sum(0).to(dbeta.dtype)


def ln_bwd_apply(dy, x, sums, eps, gamma, dsums, dx, dx_sum=None):
    mean, rstd, m = _ln_mr(x, sums, eps)
    xh = (val(x) - mean) * rstd
    g = val(dy) * gamma.double()
    shape = [x.shape[1]] + [1] * (x.dim() - 2)
    out = rstd * (g - (dsums[:, 0].double() / m).reshape(shape) - xh * (dsums[:, 1].double() / m).reshape(shape))
    put(dx, out)
    if dx_sum is not None:
        dx_sum += out.reshape(-1, x.shape[-1]).sum(0).to(dx_sum.dtype)


def upscale2x(x, y, scale=1.0, mask=None, mask_kind=MASK_NONE):
    v = scale * val(x).repeat_interleave(2, 1).repeat_interleave(2, 2)
    if mask is not None and mask_kind != MASK_NONE:
        a = val(mask)
        v = v * torch.where(a > 0, torch.ones_like(a), torch.full_like(a, 0.2 if mask_kind == MASK_LRELU else 0.0))
    put(y, v)


def copy_window(src, s_coff, dst, d_coff, c):
    dst[..., d_coff:d_coff + c] = src[..., s_coff:s_coff + c]


def pool2x(x, y, scale=0.25):
    v = val(x)
    n, h, w, c = v.shape
    put(y, scale * v.reshape(n, h // 2, 2, w // 2, 2, c).sum((2, 4)))


def axpby(x, z, out, ab):
    v = float(ab[0]) * val(x)
    if z is not None:
        v = v + float(ab[1]) * val(z)
    put(out, v)


def img_to_c8(img, dst, sample_scale=None):
    v = img.double()
    if sample_scale is not None:
        v = v * sample_scale.double().reshape(-1, 1, 1, 1)
    put(dst, torch.cat([v, torch.zeros(*v.shape[:-1], 5, dtype=v.dtype)], -1))


def c8_to_img(src, img):
    img.copy_(val(src)[..., :3].reshape(img.shape).to(img.dtype))
