"""Parity of the kind::tf32 path (fp32 channels-last activations) against the UN-EMULATED fp32 torch-CPU oracle.

north_star's tolerance for this precision: per-layer activations and gradients within rel-L2 1e-3 of the reference's fp32
CPU arithmetic (networks_3d.py:48-50 nn.Conv3d on fp32 Tensors).  Inputs and weights here are plain fp32 draws — NOT
pre-rounded to the operand precision — so the measured error contains the operand rounding (2^-11 per tf32 operand),
the fp32 accumulation order and the tf32 rounding of stored activations."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-3        # north_star: 1e-3 at TF32


@pytest.fixture()
def tf32(hpvg_gpu):
    hpvg_gpu.set_precision("tf32")
    yield hpvg_gpu
    hpvg_gpu.set_precision("bf16")


def _conv_ref(x, w, b, act=None):
    y = F.conv3d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), padding=1)
    if act == "lrelu":
        y = F.leaky_relu(y, 0.2)
    if act == "tanh":
        y = torch.tanh(y)
    return y.numpy()


def _f32(rng, shape, scale=1.0):
    return (rng.standard_normal(shape) * scale).astype(np.float32)


@pytest.mark.parametrize("C,pitch,off", [(64, 64, 0), (128, 128, 0), (64, 128, 64), (20, 24, 0), (3, 4, 0), (1, 4, 0)])
def test_pack_unpack_f32(tf32, C, pitch, off):
    hp, ops = tf32, tf32.ops
    rng = np.random.default_rng(C + pitch + off)
    x = _f32(rng, (3, C, 2, 7, 13))
    cl = hp.Tensor((3, 2, 7, 13, pitch), hp.F32)
    cl.copy_from_host(np.full((3, 2, 7, 13, pitch), 1.0, np.float32))
    ops.pack_cl(hp.from_numpy(x), c_pitch=pitch, c_off=off, out=cl)
    raw = cl.numpy()
    got = np.moveaxis(raw[..., off:off + C], -1, 1)
    # values are rounded to tf32: 10 mantissa bits -> relative error <= 2^-11, low 13 bits cleared
    assert np.all(np.abs(got - x) <= np.abs(x) * 2.0 ** -11 + 1e-30)
    assert not (raw[..., off:off + C].view(np.uint32) & 0x1FFF).any()
    lo, hi = off + C, min(pitch, (off + C + 3) // 4 * 4)
    assert not raw[..., lo:hi].any()
    assert np.all(raw[..., :off] == 1.0) and np.all(raw[..., hi:] == 1.0)
    assert np.array_equal(ops.unpack_cl(cl, C=C, c_off=off).numpy(), got)


@pytest.mark.parametrize("shape", [(1, 4, 30, 41), (2, 5, 16, 8), (1, 1, 7, 5), (1, 3, 40, 17)])
def test_conv_64_64_tf32(tf32, shape):
    hp, ops = tf32, tf32.ops
    N, T, H, W = shape
    rng = np.random.default_rng(1)
    x, w, b = _f32(rng, (N, 64, T, H, W)), _f32(rng, (64, 64, 3, 3, 3), 0.05), _f32(rng, 64, 0.1)
    x_cl = ops.pack_cl(hp.from_numpy(x))
    assert x_cl.dtype == hp.F32
    aff = ops.affine_from_bias(hp.from_numpy(b))
    y_cl = ops.conv3d_cl_any(x_cl, hp.from_numpy(w), aff, ops.ACT_LRELU, 64, 64)
    assert y_cl.dtype == hp.F32
    err = rel_l2(ops.unpack_cl(y_cl).numpy(), _conv_ref(x, w, b, "lrelu"))
    assert err < TOL, "tf32 conv 64->64 %s rel-L2 %.3e" % (shape, err)


def test_conv_128_in_and_128_out_tf32(tf32):
    hp, ops = tf32, tf32.ops
    rng = np.random.default_rng(2)
    x, w, b = _f32(rng, (1, 128, 3, 20, 27)), _f32(rng, (64, 128, 3, 3, 3), 0.05), _f32(rng, 64, 0.1)
    y = ops.conv3d_cl_any(ops.pack_cl(hp.from_numpy(x)), hp.from_numpy(w), ops.affine_from_bias(hp.from_numpy(b)),
                          ops.ACT_NONE, 128, 64)
    assert rel_l2(ops.unpack_cl(y).numpy(), _conv_ref(x, w, b)) < TOL
    x2, w2, b2 = _f32(rng, (1, 64, 3, 20, 27)), _f32(rng, (128, 64, 3, 3, 3), 0.05), _f32(rng, 128, 0.1)
    y2 = ops.conv3d_cl_any(ops.pack_cl(hp.from_numpy(x2)), hp.from_numpy(w2), ops.affine_from_bias(hp.from_numpy(b2)),
                           ops.ACT_NONE, 64, 128)
    assert rel_l2(ops.unpack_cl(y2).numpy(), _conv_ref(x2, w2, b2)) < TOL


@pytest.mark.parametrize("shape,cin", [((1, 4, 30, 41), 3), ((2, 2, 16, 8), 3), ((1, 1, 9, 11), 1)])
def test_conv_head_tf32(tf32, shape, cin):
    hp, ops = tf32, tf32.ops
    N, T, H, W = shape
    rng = np.random.default_rng(3)
    x, w, b = _f32(rng, (N, cin, T, H, W)), _f32(rng, (64, cin, 3, 3, 3), 0.2), _f32(rng, 64, 0.1)
    x_cl = ops.pack_cl(hp.from_numpy(x), c_pitch=ops.narrow_pitch())
    assert x_cl.shape[-1] == 4
    y = ops.conv3d_cl_any(x_cl, hp.from_numpy(w), ops.affine_from_bias(hp.from_numpy(b)), ops.ACT_LRELU, cin, 64)
    err = rel_l2(ops.unpack_cl(y).numpy(), _conv_ref(x, w, b, "lrelu"))
    assert err < TOL, "tf32 head conv rel-L2 %.3e" % err


@pytest.mark.parametrize("shape,cout", [((1, 4, 30, 41), 3), ((3, 1, 17, 9), 3), ((1, 5, 16, 33), 1)])
def test_conv_tail_tf32(tf32, shape, cout):
    hp, ops = tf32, tf32.ops
    N, T, H, W = shape
    rng = np.random.default_rng(4)
    x, w, b = _f32(rng, (N, 64, T, H, W)), _f32(rng, (cout, 64, 3, 3, 3), 0.05), _f32(rng, cout, 0.1)
    res = _f32(rng, (N, cout, T, H, W), 0.3)
    y = ops.conv3d_cl_any(ops.pack_cl(hp.from_numpy(x)), hp.from_numpy(w), ops.affine_from_bias(hp.from_numpy(b)),
                          ops.ACT_TANH, 64, cout, residual=hp.from_numpy(res)).numpy()
    err = rel_l2(y, np.tanh(_conv_ref(x, w, b) + res))
    assert err < TOL, "tf32 tail conv rel-L2 %.3e" % err


def test_conv_2d_tf32(tf32):
    hp, ops = tf32, tf32.ops
    rng = np.random.default_rng(5)
    x, w, b = _f32(rng, (2, 64, 1, 30, 41)), _f32(rng, (64, 64, 3, 3), 0.05), _f32(rng, 64, 0.1)
    y = ops.conv3d_cl_any(ops.pack_cl(hp.from_numpy(x)), hp.from_numpy(w), ops.affine_from_bias(hp.from_numpy(b)),
                          ops.ACT_LRELU, 64, 64)
    ref = F.leaky_relu(F.conv2d(torch.from_numpy(x[:, :, 0]), torch.from_numpy(w), torch.from_numpy(b), padding=1), 0.2)
    assert rel_l2(ops.unpack_cl(y).numpy()[:, :, 0], ref.numpy()) < TOL


def test_conv_fused_bn_stats_tf32(tf32):
    """Training-mode ConvBlock3D forward: conv + bias with the batch statistics from the epilogue, then one BN+LReLU pass
    (networks_3d.py:45-54) against torch's batch_norm in fp32."""
    hp, ops = tf32, tf32.ops
    rng = np.random.default_rng(6)
    N, T, H, W = 1, 4, 30, 41
    x, w, b = _f32(rng, (N, 64, T, H, W)), _f32(rng, (64, 64, 3, 3, 3), 0.05), _f32(rng, 64, 0.1)
    gamma, beta = (1 + 0.1 * rng.standard_normal(64)).astype(np.float32), _f32(rng, 64, 0.1)
    stats = hp.Tensor((2, 64), hp.F64).zero_()
    y = ops.conv3d_cl_any(ops.pack_cl(hp.from_numpy(x)), hp.from_numpy(w), ops.affine_from_bias(hp.from_numpy(b)),
                          ops.ACT_NONE, 64, 64, stats=stats)
    mm, mv = hp.from_numpy(np.zeros(64, np.float32)), hp.from_numpy(np.ones(64, np.float32))
    saved = hp.Tensor((4, 64), hp.F32)
    a = ops.bn_train_fused_cl(y, stats, hp.from_numpy(gamma), hp.from_numpy(beta), mm, mv, ops.ACT_LRELU, saved=saved)
    yt = F.conv3d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), padding=1)
    rm, rv = torch.zeros(64), torch.ones(64)
    ref = F.leaky_relu(F.batch_norm(yt, rm, rv, torch.from_numpy(gamma), torch.from_numpy(beta), True, 0.1, 1e-5), 0.2)
    assert rel_l2(ops.unpack_cl(a).numpy(), ref.numpy()) < TOL
    assert rel_l2(mm.numpy(), rm.numpy()) < TOL
    # torch's running_var is the unbiased estimate, MindSpore's (and ours) the biased one
    n = N * T * H * W
    assert rel_l2(mv.numpy(), (0.9 + 0.1 * (rv.numpy() - 0.9) / 0.1 * (n - 1) / n)) < TOL
    # stand-alone statistics kernel == fused ones
    s2 = hp.Tensor((2, 64), hp.F64)
    hp.ops.check(hp.lib.hpvg_bn_stats_cl_f32(y.ptr, n, s2.ptr, s2.ptr + 512, None), "bn_stats_f32")
    # different summation trees (fp32 partials per block / per epilogue warp, fp64 atomics across them)
    assert np.allclose(s2.numpy(), stats.numpy(), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("shape", [(1, 4, 30, 41), (2, 3, 9, 70), (1, 1, 13, 64), (1, 5, 6, 129)])
def test_wgrad_tf32(tf32, shape):
    hp, ops = tf32, tf32.ops
    N, T, H, W = shape
    rng = np.random.default_rng(7)
    x, gy = _f32(rng, (N, 64, T, H, W)), _f32(rng, (N, 64, T, H, W))
    w = torch.zeros(64, 64, 3, 3, 3, requires_grad=True)
    F.conv3d(torch.from_numpy(x), w, None, padding=1).backward(torch.from_numpy(gy))
    dw = hp.Tensor((64, 64, 3, 3, 3), hp.F32)
    ops.conv_wgrad_cl(ops.pack_cl(hp.from_numpy(x)), ops.pack_cl(hp.from_numpy(gy)), dw)
    err = rel_l2(dw.numpy(), w.grad.numpy())
    assert err < TOL, "tf32 wgrad %s rel-L2 %.3e" % (shape, err)
    ops.conv_wgrad_cl(ops.pack_cl(hp.from_numpy(x)), ops.pack_cl(hp.from_numpy(gy)), dw, accumulate=True, scale=0.5)
    assert rel_l2(dw.numpy(), 1.5 * w.grad.numpy()) < TOL


def test_wgrad_narrow_and_wide_tf32(tf32):
    """Head conv (Cin = 3: narrow x), tail conv (Cout = 3: narrow gy), 128-channel operands (64-channel slices)."""
    hp, ops = tf32, tf32.ops
    N, T, H, W = 2, 3, 21, 70
    rng = np.random.default_rng(8)
    x3, gy = _f32(rng, (N, 3, T, H, W)), _f32(rng, (N, 64, T, H, W))
    w = torch.zeros(64, 3, 3, 3, 3, requires_grad=True)
    F.conv3d(torch.from_numpy(x3), w, None, padding=1).backward(torch.from_numpy(gy))
    dw = hp.Tensor((64, 3, 3, 3, 3), hp.F32)
    ops.conv_wgrad_cl(ops.pack_cl(hp.from_numpy(x3), c_pitch=4), ops.pack_cl(hp.from_numpy(gy)), dw, ci_n=3)
    assert rel_l2(dw.numpy(), w.grad.numpy()) < TOL
    x, gy3 = _f32(rng, (N, 64, T, H, W)), _f32(rng, (N, 3, T, H, W))
    w2 = torch.zeros(3, 64, 3, 3, 3, requires_grad=True)
    F.conv3d(torch.from_numpy(x), w2, None, padding=1).backward(torch.from_numpy(gy3))
    dw2 = hp.Tensor((3, 64, 3, 3, 3), hp.F32)
    ops.conv_wgrad_cl(ops.pack_cl(hp.from_numpy(x)), ops.pack_cl(hp.from_numpy(gy3), c_pitch=4), dw2, co_n=3)
    assert rel_l2(dw2.numpy(), w2.grad.numpy()) < TOL
    xw, gw = _f32(rng, (N, 128, T, H, W)), _f32(rng, (N, 128, T, H, W))
    w3 = torch.zeros(128, 128, 3, 3, 3, requires_grad=True)
    F.conv3d(torch.from_numpy(xw), w3, None, padding=1).backward(torch.from_numpy(gw))
    dw3 = hp.Tensor((128, 128, 3, 3, 3), hp.F32).zero_()
    xc, gc = ops.pack_cl(hp.from_numpy(xw)), ops.pack_cl(hp.from_numpy(gw))
    for ob in range(2):
        for ib in range(2):
            ops.conv_wgrad_cl(xc, gc, dw3, co_off=ob * 64, ci_off=ib * 64, x_coff=ib * 64, gy_coff=ob * 64)
    assert rel_l2(dw3.numpy(), w3.grad.numpy()) < TOL


def test_dgrad_tf32(tf32):
    """Data gradient = the forward kernel with the transposed / mirrored bank (64->64, 3->64 head, 64->3 tail)."""
    hp, ops = tf32, tf32.ops
    rng = np.random.default_rng(9)
    N, T, H, W = 1, 4, 20, 27
    unit = hp.from_numpy(np.concatenate([np.ones(64, np.float32), np.zeros(64, np.float32)]).reshape(2, 64))
    ss = (unit.view((64,), hp.F32, 0), unit.view((64,), hp.F32, 256))
    for cin, cout in ((64, 64), (3, 64), (64, 3)):
        w, gy = _f32(rng, (cout, cin, 3, 3, 3), 0.1), _f32(rng, (N, cout, T, H, W))
        xt = torch.zeros((N, cin, T, H, W), requires_grad=True)
        F.conv3d(xt, torch.from_numpy(w), padding=1).backward(torch.from_numpy(gy))
        gy_cl = ops.pack_cl(hp.from_numpy(gy))
        dx = ops.conv3d_cl_any(gy_cl, hp.from_numpy(w), None, ops.ACT_NONE, cout if cout > 4 else gy_cl.shape[-1], cin,
                               wimgs=ops.build_wimgs(hp.from_numpy(w), cout, cin, transpose_flip=True), scale_shift=ss)
        got = dx.numpy() if cin <= 4 else ops.unpack_cl(dx).numpy()
        err = rel_l2(got, xt.grad.numpy())
        assert err < TOL, "tf32 dgrad %d->%d rel-L2 %.3e" % (cin, cout, err)


def test_bn_lrelu_backward_tf32(tf32):
    hp, ops = tf32, tf32.ops
    rng = np.random.default_rng(10)
    V = 4 * 30 * 41
    y = _f32(rng, (1, 64, 4, 30, 41))
    ga = _f32(rng, (1, 64, 4, 30, 41))
    gamma, beta = (1 + 0.1 * rng.standard_normal(64)).astype(np.float32), _f32(rng, 64, 0.1)
    yt = torch.from_numpy(y).requires_grad_(True)
    g, bt = torch.from_numpy(gamma).requires_grad_(True), torch.from_numpy(beta).requires_grad_(True)
    out = F.leaky_relu(F.batch_norm(yt, None, None, g, bt, True, 0.1, 1e-5), 0.2)
    out.backward(torch.from_numpy(ga))
    # y is the conv's pre-BatchNorm output: the tf32 conv stores it UNROUNDED (its sign after the affine is the LeakyReLU
    # mask; rounding it flips ~1e-5 of the masks = 4e-3 rel-L2 on gy, measured), so it is uploaded as plain fp32 here
    y_cl = hp.from_numpy(np.ascontiguousarray(np.moveaxis(y, 1, -1)))
    ga_cl = ops.pack_cl(hp.from_numpy(ga))
    y_r, ga_r = y, ops.unpack_cl(ga_cl).numpy()
    mean, var = y_r.mean(axis=(0, 2, 3, 4)), y_r.var(axis=(0, 2, 3, 4))
    invstd = 1.0 / np.sqrt(var + 1e-5)
    saved = hp.from_numpy(np.stack([gamma * invstd, beta - mean * gamma * invstd, mean, invstd]).astype(np.float32))
    dg, db = hp.Tensor((64,), hp.F32), hp.Tensor((64,), hp.F32)
    gy = ops.bn_bwd_cl(ga_cl, y_cl, saved, ops.ACT_LRELU, dgamma=dg, dbeta=db)
    assert rel_l2(ops.unpack_cl(gy).numpy(), yt.grad.numpy()) < TOL
    assert rel_l2(dg.numpy(), g.grad.numpy()) < TOL and rel_l2(db.numpy(), bt.grad.numpy()) < TOL
    # LeakyReLU backward and the bias-gradient column sum
    a = _f32(rng, (1, 64, 4, 30, 41))
    gz = ops.lrelu_bwd_cl(ga_cl, ops.pack_cl(hp.from_numpy(a)))
    assert rel_l2(ops.unpack_cl(gz).numpy(), np.where(a > 0, ga, 0.2 * ga)) < TOL
    cs = hp.Tensor((64,), hp.F32)
    ops.colsum_cl(ga_cl, cs)
    assert rel_l2(cs.numpy(), ga_r.sum(axis=(0, 2, 3, 4))) < 1e-5
    assert V == y_r.size // 64


def test_upsample_noise_pack_tf32(tf32):
    hp, ops = tf32, tf32.ops
    rng = np.random.default_rng(11)
    x = _f32(rng, (2, 3, 4, 24, 33))
    noise = _f32(rng, (2, 3, 5, 30, 41))
    up, xin = ops.upsample_noise_pack(hp.from_numpy(x), (5, 30, 41), noise=hp.from_numpy(noise), amp=0.37)
    assert xin.dtype == hp.F32 and xin.shape == (2, 5, 30, 41, 4)
    ref_up = ops.resize3d(hp.from_numpy(x), (5, 30, 41)).numpy()
    assert np.array_equal(up.numpy(), ref_up)
    got = np.moveaxis(xin.numpy(), -1, 1)
    assert not got[:, 3].any()
    assert rel_l2(got[:, :3], ref_up + 0.37 * noise) < 5e-4
