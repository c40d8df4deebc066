"""GPU parity of the backward kernels (tcgen05 weight gradient, BatchNorm/LeakyReLU/tanh backward, spectral-norm chain
rule, WGAN-GP pieces) against torch-CPU autograd.  Tolerance: north_star rel-L2 <= 1e-2 for bf16 gradients."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hpvg_oracle as orc
from util import bf16_round, rel_l2

pytestmark = pytest.mark.gpu
TOL = 5e-3


@pytest.mark.parametrize("shape", [(1, 4, 24, 33), (2, 3, 9, 70), (1, 1, 13, 64), (1, 5, 6, 129)])
def test_wgrad_64_64(hpvg_gpu, shape):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    N, T, H, W = shape
    rng = np.random.default_rng(0)
    x = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    gy = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    w = torch.zeros(64, 64, 3, 3, 3, requires_grad=True)
    F.conv3d(torch.from_numpy(x), w, None, padding=1).backward(torch.from_numpy(gy))
    dw = hp.Tensor((64, 64, 3, 3, 3), hp.F32)
    ops.conv_wgrad_cl(ops.pack_cl(hp.from_numpy(x)), ops.pack_cl(hp.from_numpy(gy)), dw)
    err = rel_l2(dw.numpy(), w.grad.numpy())
    assert err < TOL, "wgrad %s rel-L2 %.3e" % (shape, err)
    # accumulate + scale
    ops.conv_wgrad_cl(ops.pack_cl(hp.from_numpy(x)), ops.pack_cl(hp.from_numpy(gy)), dw, accumulate=True, scale=0.5)
    assert rel_l2(dw.numpy(), 1.5 * w.grad.numpy()) < TOL


def test_wgrad_skinny_layers_via_zero_padding(hpvg_gpu):
    """Head conv (Cin=3) and tail conv (Cout=3): operands zero-padded to 64 channels, result cropped."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    N, T, H, W = 1, 4, 20, 27
    rng = np.random.default_rng(1)
    x3 = bf16_round(rng.standard_normal((N, 3, T, H, W)))
    gy = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    w = torch.zeros(64, 3, 3, 3, 3, requires_grad=True)
    F.conv3d(torch.from_numpy(x3), w, None, padding=1).backward(torch.from_numpy(gy))
    dw = hp.Tensor((64, 3, 3, 3, 3), hp.F32)
    ops.conv_wgrad_cl(ops.pack_cl(hp.from_numpy(x3), c_pitch=64, zero_to=64), ops.pack_cl(hp.from_numpy(gy)), dw, ci_n=3)
    assert rel_l2(dw.numpy(), w.grad.numpy()) < TOL
    x = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    gy3 = bf16_round(rng.standard_normal((N, 3, T, H, W)))
    w2 = torch.zeros(3, 64, 3, 3, 3, requires_grad=True)
    F.conv3d(torch.from_numpy(x), w2, None, padding=1).backward(torch.from_numpy(gy3))
    dw2 = hp.Tensor((3, 64, 3, 3, 3), hp.F32)
    ops.conv_wgrad_cl(ops.pack_cl(hp.from_numpy(x)), ops.pack_cl(hp.from_numpy(gy3), c_pitch=64, zero_to=64), dw2, co_n=3)
    assert rel_l2(dw2.numpy(), w2.grad.numpy()) < TOL


def test_wgrad_narrow_operands_equal_zero_padded_ones(hpvg_gpu):
    """Head / tail convs in the train step: the 8-channel tensors go to the weight-gradient kernel as they are (the
    tensor map carries their real channel count, TMA zero-fills channels 8..63 of the box) — bit-identical to the
    64-channel zero-padded copy that used to be materialised."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    N, T, H, W = 2, 5, 21, 70          # ragged last W segment, more than one H block
    rng = np.random.default_rng(11)
    x3 = hp.from_numpy(bf16_round(rng.standard_normal((N, 3, T, H, W))))
    gy3 = hp.from_numpy(bf16_round(rng.standard_normal((N, 3, T, H, W))))
    x = ops.pack_cl(hp.from_numpy(bf16_round(rng.standard_normal((N, 64, T, H, W)))))
    gy = ops.pack_cl(hp.from_numpy(bf16_round(rng.standard_normal((N, 64, T, H, W)))))
    a, b = hp.Tensor((64, 3, 3, 3, 3), hp.F32), hp.Tensor((64, 3, 3, 3, 3), hp.F32)
    ops.conv_wgrad_cl(ops.pack_cl(x3, c_pitch=64, zero_to=64), gy, a, ci_n=3)
    ops.conv_wgrad_cl(ops.pack_cl(x3, c_pitch=8), gy, b, ci_n=3)
    assert np.array_equal(a.numpy(), b.numpy()) and np.abs(a.numpy()).max() > 0
    c, d = hp.Tensor((3, 64, 3, 3, 3), hp.F32), hp.Tensor((3, 64, 3, 3, 3), hp.F32)
    ops.conv_wgrad_cl(x, ops.pack_cl(gy3, c_pitch=64, zero_to=64), c, co_n=3)
    ops.conv_wgrad_cl(x, ops.pack_cl(gy3, c_pitch=8), d, co_n=3)
    assert np.array_equal(c.numpy(), d.numpy()) and np.abs(c.numpy()).max() > 0
    with pytest.raises(hp.HpvgError):   # a block wider than the narrow operand
        ops.conv_wgrad_cl(ops.pack_cl(x3, c_pitch=8), gy, hp.Tensor((64, 64, 3, 3, 3), hp.F32), ci_n=64)


def test_wgrad_2d(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(2)
    x = bf16_round(rng.standard_normal((1, 64, 1, 30, 41)))
    gy = bf16_round(rng.standard_normal((1, 64, 1, 30, 41)))
    w = torch.zeros(64, 64, 3, 3, requires_grad=True)
    F.conv2d(torch.from_numpy(x[:, :, 0]), w, None, padding=1).backward(torch.from_numpy(gy[:, :, 0]))
    dw = hp.Tensor((64, 64, 3, 3), hp.F32)
    ops.conv_wgrad_cl(ops.pack_cl(hp.from_numpy(x)), ops.pack_cl(hp.from_numpy(gy)), dw)
    assert rel_l2(dw.numpy(), w.grad.numpy()) < TOL


def test_bn_lrelu_backward(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(3)
    y = bf16_round(rng.standard_normal((2, 64, 3, 11, 13)) * 1.5 + 0.2)
    ga = bf16_round(rng.standard_normal(y.shape))
    gamma = (1 + 0.1 * rng.standard_normal(64)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(64)).astype(np.float32)
    yt = torch.from_numpy(y).requires_grad_(True)
    gt, bt = torch.from_numpy(gamma).requires_grad_(True), torch.from_numpy(beta).requires_grad_(True)
    p = {"1.bn2d.gamma": gt, "1.bn2d.beta": bt, "1.bn2d.moving_mean": torch.zeros(64),
         "1.bn2d.moving_variance": torch.ones(64)}
    orc.lrelu(orc.batchnorm(yt, p, "1.", True)).backward(torch.from_numpy(ga))
    y_cl = ops.pack_cl(hp.from_numpy(y))
    _, saved = ops.bn_train_cl(y_cl, hp.from_numpy(gamma), hp.from_numpy(beta), hp.from_numpy(np.zeros(64, np.float32)),
                               hp.from_numpy(np.ones(64, np.float32)))
    dg, db = hp.Tensor((64,), hp.F32), hp.Tensor((64,), hp.F32)
    gy = ops.bn_bwd_cl(ops.pack_cl(hp.from_numpy(ga)), y_cl, saved, dgamma=dg, dbeta=db)
    assert rel_l2(ops.unpack_cl(gy).numpy(), yt.grad.numpy()) < 1e-2
    assert rel_l2(dg.numpy(), gt.grad.numpy()) < 2e-3 and rel_l2(db.numpy(), bt.grad.numpy()) < 2e-3


def test_small_backward_ops(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(4)
    a = bf16_round(rng.standard_normal((1, 64, 2, 9, 10)))
    ga = bf16_round(rng.standard_normal(a.shape))
    gz = ops.unpack_cl(ops.lrelu_bwd_cl(ops.pack_cl(hp.from_numpy(ga)), ops.pack_cl(hp.from_numpy(a)))).numpy()
    assert rel_l2(gz, bf16_round(ga * np.where(a > 0, 1.0, 0.2))) < 1e-6
    out = np.tanh(rng.standard_normal((1, 3, 4, 9, 11))).astype(np.float32)
    tgt = rng.standard_normal(out.shape).astype(np.float32)
    g = ops.mse_grad(hp.from_numpy(out), hp.from_numpy(tgt), 10.0 * 2 / out.size)
    gpre = ops.tanh_bwd(g, hp.from_numpy(out)).numpy()
    ot = torch.from_numpy(np.arctanh(np.clip(out, -0.999999, 0.999999))).requires_grad_(True)
    (10.0 * ((torch.tanh(ot) - torch.from_numpy(tgt)) ** 2).mean()).backward()
    assert rel_l2(gpre, ot.grad.numpy()) < 1e-4
    cs = hp.Tensor((3,), hp.F32)
    ops.channel_sum(hp.from_numpy(out), cs)
    assert np.allclose(cs.numpy(), out.sum(axis=(0, 2, 3, 4)), rtol=1e-5, atol=1e-4)
    col = hp.Tensor((64,), hp.F32)
    ops.colsum_cl(ops.pack_cl(hp.from_numpy(a)), col)
    assert np.allclose(col.numpy(), a.sum(axis=(0, 2, 3, 4)), rtol=1e-4, atol=1e-3)
    mu = rng.standard_normal((1, 128, 4, 6, 7)).astype(np.float32)
    lv = (0.3 * rng.standard_normal(mu.shape)).astype(np.float32)
    mt, lt = torch.from_numpy(mu).requires_grad_(True), torch.from_numpy(lv).requires_grad_(True)
    (1.0 * orc.kl_criterion(mt, lt)).backward()
    gmu, glv = ops.kl_grad(hp.from_numpy(mu), hp.from_numpy(lv), 1.0 / mu.size)
    assert rel_l2(gmu.numpy(), mt.grad.numpy()) < 1e-5 and rel_l2(glv.numpy(), lt.grad.numpy()) < 1e-5


def test_sn_grad_matches_autograd_with_constant_uv(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(5)
    w = (rng.standard_normal((64, 64, 3, 3, 3)) * 0.02).astype(np.float32)
    u = orc._l2normalize_np(rng.standard_normal((64, 1)).astype(np.float32))
    v = orc._l2normalize_np(rng.standard_normal((1728, 1)).astype(np.float32))
    G = rng.standard_normal(w.shape).astype(np.float32)
    wt = torch.from_numpy(w).requires_grad_(True)
    sigma, un, vn = orc.sn_power_iteration(wt, torch.from_numpy(u), torch.from_numpy(v))
    ((wt / sigma) * torch.from_numpy(G)).sum().backward()
    tu, tv = hp.from_numpy(u), hp.from_numpy(v)
    sg = ops.sn_power_iter(hp.from_numpy(w), tu, tv)
    gw = hp.Tensor(w.shape, hp.F32)
    ops.sn_grad(hp.from_numpy(G), hp.from_numpy(w), tu, tv, sg, gw)
    assert rel_l2(gw.numpy(), wt.grad.numpy()) < 1e-4


def test_gp_pieces(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(6)
    g = rng.standard_normal((1, 3, 4, 9, 11)).astype(np.float32)
    gt = torch.from_numpy(g).requires_grad_(True)
    gp_ref = ((torch.sqrt((gt ** 2).sum(dim=1)) - 1) ** 2).mean() * 0.1
    gp_ref.backward()
    Gout, gp = ops.gp_grad(hp.from_numpy(g), 0.1)
    assert abs(gp.numpy()[0] - float(gp_ref)) < 1e-5
    assert rel_l2(Gout.numpy(), gt.grad.numpy()) < 1e-5
    a, b = rng.standard_normal(100).astype(np.float32), rng.standard_normal(100).astype(np.float32)
    assert np.allclose(ops.lerp(hp.from_numpy(a), hp.from_numpy(b), 0.3).numpy(), 0.3 * a + 0.7 * b, atol=1e-6)
