"""-m gpu: the reference-named op wrappers (t2i_b200.utils.ops, mirror of utils/ops.py) against the
oracle's restatement of the same TF ops, on the shapes the wgancls graph uses; default TF variable
names, reuse semantics and argument errors."""
import pytest
import torch

from oracle import wgancls_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_ops_match_oracle_and_name_variables_like_tf():
    from t2i_b200.utils import ops
    ops.reset_variables(seed=1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 3, 64, 64, generator=g).cuda()          # NCHW like the reference's d_net
    lrelu = ops.lrelu_act(0.2)
    with ops.variable_scope("d_net"):
        h0 = ops.conv2d(x, 16, ks=(4, 4), s=(2, 2), act=lrelu, df=ops.NCHW)
        h1 = ops.conv2d(h0, 32, ks=(4, 4), s=(2, 2), df=ops.NCHW, act=lrelu)
        r = ops.conv2d(h1, 8, ks=(1, 1), s=(1, 1), padding='valid', df=ops.NCHW, act=lrelu)
        r = ops.conv2d(r, 32, ks=(3, 3), s=(1, 1), df=ops.NCHW)
        e = ops.fc(torch.randn(4, 40, generator=g).cuda(), 8, act=lrelu)
    names = list(ops.global_variables("d_net/"))
    assert names == ["d_net/Conv/weights", "d_net/Conv/biases", "d_net/Conv_1/weights", "d_net/Conv_1/biases",
                     "d_net/Conv_2/weights", "d_net/Conv_2/biases", "d_net/Conv_3/weights", "d_net/Conv_3/biases",
                     "d_net/dense/kernel", "d_net/dense/bias"]
    p = {k: v.cpu().double() for k, v in ops.global_variables().items()}
    xo = x.cpu().double()
    o0 = O.conv2d(p, "d_net/Conv", xo, 4, 2, act=O.lrelu)
    o1 = O.conv2d(p, "d_net/Conv_1", o0, 4, 2, act=O.lrelu)
    orr = O.conv2d(p, "d_net/Conv_2", o1, 1, 1, "valid", act=O.lrelu)
    orr = O.conv2d(p, "d_net/Conv_3", orr, 3, 1)
    assert h0.shape == (4, 16, 32, 32) and rel(h0, o0) < 1e-4
    assert rel(h1, o1) < 1e-4 and rel(r, orr) < 1e-4
    assert e.shape == (4, 8)
    # reuse=True shares the variables and reproduces the result; a new name without reuse raises
    with ops.variable_scope("d_net", reuse=True):
        h0b = ops.conv2d(x, 16, ks=(4, 4), s=(2, 2), act=lrelu, df=ops.NCHW)
    assert torch.equal(h0, h0b) and len(ops.global_variables("d_net/")) == 10
    with ops.variable_scope("d_net", reuse=True):
        ops.conv2d(x, 16, df=ops.NCHW)
        with pytest.raises(ValueError):
            ops.conv2d(h0, 99, df=ops.NCHW)                     # Conv_1 exists with another shape
    with pytest.raises(ValueError):
        ops.conv2d(x, 8, ks=(5, 5), s=(1, 1), df=ops.NCHW)      # unsupported geometry
    with pytest.raises(ValueError):
        ops.conv2d(x, 8, df='NCWH')


def test_generator_side_ops():
    from t2i_b200.utils import ops
    ops.reset_variables(seed=2)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(6, 16, 8, 8, generator=g).cuda()
    with ops.variable_scope("g_net"):
        y = ops.conv2d_transpose(x, 24, ks=(4, 4), s=(2, 2), df=ops.NCHW)
        z = ops.batch_norm(y, train=True, act=ops.relu, df=ops.NCHW)
        f = ops.fc(torch.randn(6, 24, generator=g).cuda(), 64)
        fb = ops.batch_norm(f, train=True, act=None)
        img = ops.conv2d(ops.conv2d_transpose(z, 3, df=ops.NCHW), 3, ks=(3, 3), s=(1, 1), act=ops.tanh, df=ops.NCHW)
        out = ops.conv2d(torch.randn(6, 32, 4, 4, generator=g).cuda(), 1, ks=(4, 4), s=(4, 4), padding='valid', df=ops.NCHW)
    p = {k: v.cpu().double() for k, v in ops.global_variables().items()}
    yo = O.conv2d_transpose(p, "g_net/Conv2d_transpose", x.cpu().double())
    assert y.shape == (6, 24, 16, 16) and rel(y, yo) < 1e-4
    nm = {}
    zo = O.batch_norm(p, "g_net/BatchNorm", yo, True, act=torch.relu, new_moving=nm)
    assert rel(z, zo) < 1e-4
    assert fb.shape == (6, 64) and abs(float(fb.mean())) < 1e-4
    assert img.shape == (6, 3, 32, 32) and float(img.abs().max()) <= 1.0 and out.shape == (6, 1, 1, 1)
    io = torch.tanh(O.conv2d(p, "g_net/Conv", O.conv2d_transpose(p, "g_net/Conv2d_transpose_1", zo), 3, 1))
    assert rel(img, io) < 1e-3
    # UPDATE_OPS: moving statistics change only when the queued updates are run
    assert float(ops.global_variables()["g_net/BatchNorm/moving_mean"].abs().sum()) == 0.0
    ops.run_update_ops()
    assert rel(ops.global_variables()["g_net/BatchNorm/moving_mean"], nm["g_net/BatchNorm/moving_mean"]) < 1e-4
    assert rel(ops.global_variables()["g_net/BatchNorm/moving_variance"], nm["g_net/BatchNorm/moving_variance"]) < 1e-4
    # inference mode uses them
    zi = ops.batch_norm(y, train=False, act=ops.relu, name="BatchNorm", df=ops.NCHW) if False else None
    xn = ops.to_nhwc(x)
    assert xn.shape == (6, 8, 8, 16) and torch.equal(ops.to_nchw(xn), x)


def test_pggan_side_ops():
    """layer_norm / pool / upscale / the 2x2 and 4x4 stride-1 SAME convs (utils/ops.py:58-63,74-81,100-111) against the
    PGGAN oracle's restatement; TF default variable names (LayerNorm, LayerNorm_1)."""
    from oracle import pggan_oracle as P
    from t2i_b200.utils import ops
    ops.reset_variables(seed=3)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 8, 8, 16, generator=g).cuda()
    with ops.variable_scope("g_net"):
        with ops.variable_scope("conv_stage_1"):
            u = ops.upscale(x, 2)
            c = ops.conv2d(u, 24, ks=(3, 3), s=(1, 1))
            y = ops.layer_norm(c, act=ops.relu)
            f = ops.layer_norm(ops.fc(torch.randn(3, 40, generator=g).cuda(), 64))
        with ops.variable_scope("rgb_stage_1"):
            r = ops.conv2d(ops.conv2d(y, 9, ks=(2, 2), s=(1, 1), act=ops.relu), 3, ks=(1, 1), s=(1, 1))
        k4 = ops.conv2d(y, 8, ks=(4, 4), s=(1, 1))
        v = ops.conv2d(torch.randn(3, 4, 4, 16, generator=g).cuda(), 8, ks=(4, 4), s=(1, 1), padding='VALID')
    names = list(ops.global_variables("g_net/conv_stage_1/"))
    assert names == ["g_net/conv_stage_1/Conv/weights", "g_net/conv_stage_1/Conv/biases", "g_net/conv_stage_1/LayerNorm/beta",
                     "g_net/conv_stage_1/LayerNorm/gamma", "g_net/conv_stage_1/dense/kernel", "g_net/conv_stage_1/dense/bias",
                     "g_net/conv_stage_1/LayerNorm_1/beta", "g_net/conv_stage_1/LayerNorm_1/gamma"]
    gv = ops.global_variables()
    gv["g_net/conv_stage_1/LayerNorm/gamma"].uniform_(0.5, 1.5)
    gv["g_net/conv_stage_1/LayerNorm/beta"].normal_()
    y = None
    with ops.variable_scope("g_net", reuse=True):
        with ops.variable_scope("conv_stage_1", reuse=True):
            y = ops.layer_norm(c, act=ops.relu)
    p = {k: t.cpu().double() for k, t in gv.items()}
    xo = x.cpu().double()
    uo = P.upscale(xo)
    co = P.conv2d(p, "g_net/conv_stage_1/Conv", uo, 3)
    yo = P.layer_norm(p, "g_net/conv_stage_1/LayerNorm", co, torch.relu)
    assert u.shape == (3, 16, 16, 16) and rel(u, uo) < 1e-4 and rel(c, co) < 1e-4 and rel(y, yo) < 1e-4
    assert f.shape == (3, 64) and float(f.mean(1).abs().max()) < 1e-3
    with ops.variable_scope("g_net", reuse=True):
        with ops.variable_scope("rgb_stage_1", reuse=True):
            r = ops.conv2d(ops.conv2d(y, 9, ks=(2, 2), s=(1, 1), act=ops.relu), 3, ks=(1, 1), s=(1, 1))
    ro = P.conv2d(p, "g_net/rgb_stage_1/Conv_1", P.conv2d(p, "g_net/rgb_stage_1/Conv", yo, 2, act=torch.relu), 1)
    assert r.shape == (3, 16, 16, 3) and rel(r, ro) < 1e-3       # the 2x2 SAME conv pads bottom / right (TF)
    yb = ops.pool(y)
    assert yb.shape == (3, 8, 8, 24) and rel(yb, P.pool(yo)) < 1e-4
    assert k4.shape == (3, 16, 16, 8) and v.shape == (3, 1, 1, 8)
    with pytest.raises(ValueError):
        ops.pool(y, 3)
    with pytest.raises(ValueError):
        ops.layer_norm(torch.randn(2, 4, 4, 12).cuda())
