"""Parity at BASELINE.json's FULL sizes (finest scale 13 x 192 x 257 / 16 x 192 x 257, full 10-scale pyramid).
The whole pyramid or a whole train step would cost the CPU oracle minutes per case, so those are covered by
size-independent identities the operators must satisfy exactly or to bf16 round-off: linearity, adjointness
(<Ax, y> == <x, A^T y>), Euler's identity for the weight gradient, partition of unity, batch / shard invariance, and
agreement between the fused and the stand-alone statistics paths.  Single layers and single networks at the full
size ARE within the oracle's reach (seconds of oneDNN): the second half of the file compares them value by value.
The small-size oracle comparisons live in the other test files."""
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


# ------------------------------------------------------------------------------------------- direct oracle, full size
# A single layer at the full size IS within reach of the CPU oracle (a few seconds of oneDNN on the box's host cores):
# the three operators that carry the work — conv fprop, its weight gradient, the pyramid resize — are also compared
# value by value at BASELINE.json's sizes, including the 16-frame finest scale of config 3.
@pytest.mark.parametrize("frames", [13, 16])
def test_conv_block_matches_oracle_at_full_size(hpvg_gpu, frames):
    """ConvBlock3D without BatchNorm (networks_3d.py:45-54: Conv3d 3x3x3 pad 1 + bias -> LeakyReLU(0.2)), 64 -> 64."""
    import torch
    import torch.nn.functional as F
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(frames)
    T, H, W = frames, FULL[1], FULL[2]
    x = bf16_round(rng.standard_normal((1, 64, T, H, W)))
    w = bf16_round(rng.standard_normal((64, 64, 3, 3, 3)) * 0.02)
    b = (rng.standard_normal(64) * 0.1).astype(np.float32)
    aff = hp.from_numpy(np.stack([np.ones(64, np.float32), b]))
    y = ops.unpack_cl(ops.conv3d_cl_any(ops.pack_cl(hp.from_numpy(x)), hp.from_numpy(w), aff, ops.ACT_LRELU, 64, 64)).numpy()
    with torch.no_grad():
        ref = F.leaky_relu(F.conv3d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), padding=1), 0.2).numpy()
    assert rel_l2(y, ref) < 4e-3          # bf16 output rounding only: operands are bf16-exact, accumulation fp32
    assert np.max(np.abs(y - ref)) < 2e-2


def test_wgrad_matches_oracle_at_full_size(hpvg_gpu):
    """MindSpore autodiff of Conv3d wrt its weight (restated with torch-CPU autograd) on one full-size layer."""
    import torch
    import torch.nn.functional as F
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(7)
    T, H, W = FULL
    x = bf16_round(rng.standard_normal((1, 64, T, H, W)))
    gy = bf16_round(rng.standard_normal((1, 64, T, H, W)))
    w = torch.zeros(64, 64, 3, 3, 3, requires_grad=True)
    F.conv3d(torch.from_numpy(x), w, None, padding=1).backward(torch.from_numpy(gy))
    dw = hp.Tensor((64, 64, 3, 3, 3), hp.F32)
    ops.conv_wgrad_cl(ops.pack_cl(hp.from_numpy(x)), ops.pack_cl(hp.from_numpy(gy)), dw)
    assert rel_l2(dw.numpy(), w.grad.numpy()) < 1e-4     # bf16-exact operands, fp32 accumulation over 641 k voxels


def test_resize_matches_oracle_at_full_size(hpvg_gpu):
    """UpsampleTrilinear3D(align_corners=True) s8 -> s9 (images.py:54-61) and its adjoint, against the numpy oracle."""
    from oracle import hpvg_oracle as orc
    hp = hpvg_gpu
    rng = np.random.default_rng(8)
    x = rng.standard_normal((2, 3) + PREV).astype(np.float32)
    y = hp.ops.resize3d(hp.from_numpy(x), FULL).numpy()
    ref = orc.resize_linear_np(x, FULL, True)
    assert np.max(np.abs(y - ref)) <= 1e-6 and float(np.mean(y == ref)) > 0.99
    gy = rng.standard_normal((2, 3) + FULL).astype(np.float32)
    gx = hp.ops.resize3d_bwd(hp.from_numpy(gy), PREV).numpy()
    assert rel_l2(gx, orc.resize_linear_bwd_np(gy, PREV, True)) < 1e-6



def test_discriminator_matches_oracle_at_full_size(hpvg_gpu):
    """WDiscriminator3D (networks_3d.py:170-193: SN head 3 -> 64, five SN 64 -> 64 blocks, plain tail 64 -> 1) on one
    full-size clip, spectral-norm power iteration included."""
    import torch
    from oracle import hpvg_oracle as orc
    hp = hpvg_gpu
    from hpvg import networks_3d as n3
    from hpvg.utils import images as uimg
    opt, oopt = uimg.default_opt(), orc.default_opt()
    params = orc.init_discriminator_params(oopt, seed=21)
    d = n3.WDiscriminator3D(opt)
    d.load_parameters(params)
    x = np.tanh(np.random.default_rng(3).standard_normal((1, 3) + FULL)).astype(np.float32)
    with torch.no_grad():
        ref = orc.discriminator(torch.from_numpy(x), orc.to_torch(params), oopt).numpy()
    out = d(hp.from_numpy(x)).numpy()
    assert out.shape == ref.shape == (1, 1) + FULL
    assert rel_l2(out, ref) < 1e-2          # 7 bf16-operand layers deep


def test_body_block_train_mode_matches_oracle_at_full_size(hpvg_gpu):
    """One refinement block (networks_3d.py:395-401: 3 -> 64 -> 64 x 4 -> 3, BatchNorm in training mode — batch statistics
    over all 641 k voxels fused into the conv epilogues) + residual + tanh at the finest scale."""
    import torch
    from oracle import hpvg_oracle as orc
    hp = hpvg_gpu
    from hpvg import networks_3d as n3
    from hpvg.utils import images as uimg
    opt, oopt = uimg.default_opt(), orc.default_opt()
    params = orc.init_generator_params(oopt, 1, seed=31)
    net = n3.GeneratorHPVAEGAN(opt)
    net.init_next_stage()
    net.load_parameters(params)
    net.set_train(True)
    up = np.tanh(np.random.default_rng(4).standard_normal((1, 3) + FULL)).astype(np.float32)
    pt = orc.to_torch(params)
    with torch.no_grad():
        ref = torch.tanh(orc.block_forward(torch.from_numpy(up), pt, "body.0.", oopt, True) + torch.from_numpy(up)).numpy()
    x_in = hp.ops.pack_cl(hp.from_numpy(up), c_pitch=8)
    net.bn_slab.reset()                     # what construct() does at the start of a pass
    out = net._run_block(net.body[0], x_in, hp.from_numpy(up), "t", None).numpy()
    assert rel_l2(out, ref) < 1e-2


def test_full_pyramid_sample_matches_oracle(hpvg_gpu):
    """eval_video.py:53-82 for one clip at the real size: decoder + all nine refinement stages of the default pyramid
    (1.33 TFLOP; 70 bf16-operand convs deep), eval-mode BatchNorm with randomised moving statistics, given z and given
    refinement noise — against the fp32 CPU oracle and against the same oracle with bf16-rounded conv operands.
    Per-layer errors of ~3e-3 accumulate through the random-init network: 3.2e-2 after 5 scales (the golden fixture),
    7.4e-2 after all 10."""
    import torch
    from oracle import hpvg_oracle as orc
    hp = hpvg_gpu
    from hpvg import networks_3d as n3
    from hpvg.utils import images as uimg
    opt, oopt = uimg.default_opt(), orc.default_opt()
    n_body = opt.stop_scale
    params = orc.randomize_bn_stats(orc.init_generator_params(oopt, n_body, seed=41), opt=oopt)
    net = n3.GeneratorHPVAEGAN(opt)
    for _ in range(n_body):
        net.init_next_stage()
    net.load_parameters(params)
    rng = np.random.default_rng(6)
    z = rng.standard_normal((1, opt.latent_dim) + orc.scale_shape(oopt, 0)).astype(np.float32)
    amps = [1.0] + [0.1] * n_body
    noises = {s: rng.standard_normal((1, 3) + orc.scale_shape(oopt, s)).astype(np.float32)
              for s in range(opt.vae_levels, n_body + 1)}
    tn = {k: torch.from_numpy(v) for k, v in noises.items()}
    with torch.no_grad():
        rx, rv = orc.generator_forward(None, amps, orc.to_torch(params), oopt, noise_init=torch.from_numpy(z),
                                       is_random=True, noises=tn)
        with orc.bf16_emulation():      # the same fp32 oracle with every conv operand rounded to bf16 first
            bx, bv = orc.generator_forward(None, amps, orc.to_torch(params), oopt, noise_init=torch.from_numpy(z),
                                           is_random=True, noises=tn)
    tz = hp.from_numpy(z)
    x, vae = net(tz, amps, noise_init=tz, isRandom=True, noises={k: hp.from_numpy(v) for k, v in noises.items()})
    assert tuple(x.shape) == tuple(rx.shape) == (1, 3) + FULL
    e_v, e_x = rel_l2(vae.numpy(), rv.numpy()), rel_l2(x.numpy(), rx.numpy())
    b_x, ob = rel_l2(x.numpy(), bx.numpy()), rel_l2(bx.numpy(), rx.numpy())
    print("full pyramid: vae_out rel-L2 %.3e, sample rel-L2 %.3e vs fp32 oracle, %.3e vs bf16-operand oracle "
          "(bf16-operand oracle vs fp32 oracle: %.3e)" % (e_v, e_x, b_x, ob))
    # 70 layers deep the sample is limited by bf16 operand rounding itself: the fp32 oracle with bf16-rounded conv
    # operands is 7.1e-2 away from the plain fp32 oracle (measured); the kernels must not be further away than that
    # rounding alone explains, and stay close to the bf16-operand oracle (two independent rounding patterns: 5e-2)
    assert e_v < 1e-2 and e_x < 1.3 * ob + 5e-3 and e_x < 0.12 and b_x < 8e-2
