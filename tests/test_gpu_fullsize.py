"""Parity at BASELINE.json's FULL sizes (finest scale 13 x 192 x 257, full 10-scale pyramid) through size-independent
properties — the CPU oracle would need minutes per case here, so these tests use identities the operators must
satisfy exactly or to bf16 round-off: linearity, adjointness (<Ax, y> == <x, A^T y>), Euler's identity for the weight
gradient, partition of unity, batch / shard invariance, and agreement between the fused and the stand-alone
statistics paths.  The small-size oracle comparisons live in the other test files."""
import numpy as np
import pytest

from util import bf16_round, rel_l2

pytestmark = pytest.mark.gpu
FULL = (13, 192, 257)          # finest scale of the default pyramid
PREV = (7, 153, 204)


def _dot(a, b):
    return float(np.vdot(a.astype(np.float64), b.astype(np.float64)))


def test_conv_64_64_linearity_and_adjoint_at_full_size(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(0)
    T, H, W = FULL
    unit = hp.from_numpy(np.concatenate([np.ones(64, np.float32), np.zeros(64, np.float32)]).reshape(2, 64))
    w = hp.from_numpy(bf16_round(rng.standard_normal((64, 64, 3, 3, 3)) * 0.05))
    # inputs that are exactly representable in bf16 and whose combination is too: small integers / 8
    x = (rng.integers(-8, 9, (1, T, H, W, 64)) / 8.0).astype(np.float32)
    y = (rng.integers(-8, 9, (1, T, H, W, 64)) / 8.0).astype(np.float32)

    def up(a):   # channels-last fp32 numpy -> bf16 cl device tensor (values are bf16-exact)
        t = hp.Tensor(a.shape, hp.BF16)
        t.copy_from_host((a.view(np.uint32) >> 16).astype(np.uint16))
        return t

    def conv(a_cl, flip=False):
        out = ops.conv_cl(ops.CONV_64_64, a_cl, ops.pack_weights(w, ops.CONV_64_64, flip), unit,
                          unit.view((64,), hp.F32, 256), ops.ACT_NONE, ops.OUT_F32_RAW)
        return out.numpy()

    cx, cy = conv(up(x)), conv(up(y))
    cxy = conv(up(2.0 * x - 0.5 * y))
    # fp32 accumulation of the same bf16 products in a different grouping: ~1e-6 relative
    assert rel_l2(cxy, 2.0 * cx - 0.5 * cy) < 1e-5
    # adjoint: <conv(x; W), g> == <x, dgrad(g; W)>  (dgrad = the same kernel with the transposed / mirrored bank)
    g = (rng.integers(-8, 9, (1, T, H, W, 64)) / 8.0).astype(np.float32)
    lhs, rhs = _dot(cx, g), _dot(x, conv(up(g), flip=True))
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)


def test_wgrad_euler_identity_at_full_size(hpvg_gpu):
    """conv is linear in W: sum_{W} W * dW(gy, x) == <gy, conv(x; W)> — ties the tcgen05 wgrad kernel to the fprop
    kernel on all 641 472 voxels without an oracle."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(1)
    T, H, W = FULL
    wn = bf16_round(rng.standard_normal((64, 64, 3, 3, 3)) * 0.05)
    w = hp.from_numpy(wn)
    unit = hp.from_numpy(np.concatenate([np.ones(64, np.float32), np.zeros(64, np.float32)]).reshape(2, 64))
    x = hp.Tensor((1, T, H, W, 64), hp.BF16)
    gy = hp.Tensor((1, T, H, W, 64), hp.BF16)
    xs = (rng.integers(-8, 9, x.shape) / 8.0).astype(np.float32)
    gs = (rng.integers(-4, 5, x.shape) / 64.0).astype(np.float32)
    x.copy_from_host((xs.view(np.uint32) >> 16).astype(np.uint16))
    gy.copy_from_host((gs.view(np.uint32) >> 16).astype(np.uint16))
    dw = hp.Tensor((64, 64, 3, 3, 3), hp.F32).zero_()
    ops.conv_wgrad_cl(x, gy, dw)
    y = ops.conv_cl(ops.CONV_64_64, x, ops.pack_weights(w, ops.CONV_64_64), unit, unit.view((64,), hp.F32, 256),
                    ops.ACT_NONE, ops.OUT_F32_RAW).numpy()
    lhs, rhs = _dot(wn, dw.numpy()), _dot(gs, y)
    assert abs(lhs - rhs) <= 2e-4 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)


def test_resize_full_size_partition_of_unity_and_adjoint(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(2)
    c = hp.from_numpy(np.full((1, 3) + PREV, 0.37, np.float32))
    up = ops.resize3d(c, FULL).numpy()
    assert np.abs(up - 0.37).max() < 1e-6                       # weights of every output voxel sum to one
    x = rng.standard_normal((2, 3) + PREV).astype(np.float32)
    y = rng.standard_normal((2, 3) + FULL).astype(np.float32)
    ax = ops.resize3d(hp.from_numpy(x), FULL).numpy()
    aty = ops.resize3d_bwd(hp.from_numpy(y), PREV).numpy()
    lhs, rhs = _dot(ax, y), _dot(x, aty)
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0)
    # the fused block-input kernel produces the same `up` bits as the stand-alone resize, and x_in = bf16(up) w/o noise
    up2, xin = ops.upsample_noise_pack(hp.from_numpy(x), FULL)
    assert np.array_equal(up2.numpy(), ax)
    got = hpvg_gpu.runtime.bf16_bits_to_f32(xin.numpy())[..., :3]
    assert np.array_equal(got, np.moveaxis(bf16_round(ax), 1, -1))


def test_fused_bn_statistics_equal_standalone_pass_at_full_size(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(3)
    T, H, W = FULL
    x = hp.Tensor((1, T, H, W, 64), hp.BF16)
    x.copy_from_host(((rng.integers(-8, 9, x.shape) / 8.0).astype(np.float32).view(np.uint32) >> 16).astype(np.uint16))
    w = hp.from_numpy(bf16_round(rng.standard_normal((64, 64, 3, 3, 3)) * 0.05))
    aff = ops.affine_from_bias(hp.from_numpy((0.1 * rng.standard_normal(64)).astype(np.float32)))
    stats = hp.Tensor((2, 64), hp.F64).zero_()
    y = ops.conv3d_cl_any(x, w, aff, ops.ACT_NONE, 64, 64, stats=stats)
    ref = hp.Tensor((2, 64), hp.F64)
    hp._lib.check(hp.lib.hpvg_bn_stats_cl(y.ptr, T * H * W, ref.ptr, ref.ptr + 512, None))
    a, b = stats.numpy(), ref.numpy()
    assert np.allclose(a, b, rtol=1e-6, atol=1e-3)              # same 41 M bf16 values, two reduction orders
    y2 = ops.conv3d_cl_any(x, w, aff, ops.ACT_NONE, 64, 64)     # the statistics variant stores the same output bits
    assert np.array_equal(y.numpy(), y2.numpy())


def test_full_pyramid_sample_is_batch_and_slot_invariant(hpvg_gpu):
    """eval_video.py semantics at the full 10-scale pyramid: a clip does not depend on what else is in the batch
    (eval-mode BatchNorm) nor on which output slot / sample index offset generated it."""
    hp = hpvg_gpu
    from hpvg import networks_3d as n3, sampling
    from hpvg.utils import images as uimg
    opt = uimg.default_opt()
    net = n3.GeneratorHPVAEGAN(opt, seed=0)
    for _ in range(opt.stop_scale):
        net.init_next_stage()
    amps = [1.0] + [0.1] * opt.stop_scale
    idx, clips = sampling.generate(net, amps, 3, batch=3, seed=5)
    assert clips.shape == (3, 3) + FULL and np.isfinite(clips).all() and np.abs(clips).max() <= 1.0
    net.out_slot = 1
    idx1, alone = sampling.generate(net, amps, 3, rank=2, world=3, batch=1, seed=5)     # only sample 2, batch of one
    assert idx1 == [2] and np.array_equal(alone[0], clips[2])
    assert not np.array_equal(clips[0], clips[1])
