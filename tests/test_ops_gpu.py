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
