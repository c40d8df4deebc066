"""GPU parity of the bandwidth-bound kernels (resize fwd/bwd, BN, spectral norm, losses, Adam) through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import hpvg_oracle as orc
from util import bf16_round, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_in,n_out", [(33, 41), (24, 30), (4, 5), (7, 13), (204, 257), (153, 192), (1, 4), (6, 1)])
@pytest.mark.parametrize("align", [True, False])
def test_resize_tables_bit_exact_device(hpvg_gpu, n_in, n_out, align):
    """north_star: resize index and weight computation must be bit-exact — as computed by the DEVICE code path."""
    i0, i1, l0, l1 = hpvg_gpu.ops.linear_taps_device(n_in, n_out, align)
    r0, r1, m0, m1 = orc.linear_taps(n_in, n_out, align)
    assert np.array_equal(i0, r0) and np.array_equal(i1, r1)
    assert np.array_equal(l0.view(np.uint32), m0.view(np.uint32))
    assert np.array_equal(l1.view(np.uint32), m1.view(np.uint32))


def test_resize_known_answer(hpvg_gpu):
    """src/tools/trilinear.py:222-233."""
    hp = hpvg_gpu
    x = np.arange(1, 5, dtype=np.float32).reshape(1, 1, 1, 2, 2)
    y = hp.ops.resize3d(hp.from_numpy(x), (2, 4, 4), align_corners=False).numpy()
    want = np.array([[1.0, 1.25, 1.75, 2.0], [1.5, 1.75, 2.25, 2.5], [2.5, 2.75, 3.25, 3.5], [3.0, 3.25, 3.75, 4.0]], np.float32)
    assert np.array_equal(y[0, 0, 0], want) and np.array_equal(y[0, 0, 1], want)


@pytest.mark.parametrize("case", [((4, 24, 33), (4, 30, 41)), ((5, 76, 102), (7, 96, 129)), ((7, 9, 11), (13, 12, 14)),
                                  ((1, 24, 33), (1, 29, 39))])
def test_resize_fwd_bwd(hpvg_gpu, case):
    hp = hpvg_gpu
    in_size, out_size = case
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3) + in_size).astype(np.float32)
    gy = rng.standard_normal((2, 3) + out_size).astype(np.float32)
    y = hp.ops.resize3d(hp.from_numpy(x), out_size).numpy()
    ref = orc.resize_linear_np(x, out_size, True)
    assert np.max(np.abs(y - ref)) <= 1e-6          # same rule, same evaluation order, fp32
    assert float(np.mean(y == ref)) > 0.99          # and in fact (nearly) always the same bits
    gx = hp.ops.resize3d_bwd(hp.from_numpy(gy), in_size).numpy()
    gref = orc.resize_linear_bwd_np(gy, in_size, True)
    assert rel_l2(gx, gref) < 1e-6


@pytest.mark.parametrize("case", [
    ((2, 6, 7), (13, 8, 9)),       # > 4 outputs per source frame: the register walk declines, tiled kernel takes over
    ((7, 20, 27), (16, 25, 33)),   # 7 -> 16 frames (BASELINE config 3): 3/2/3/2/3/2/1 outputs per source frame
    ((3, 5, 7), (3, 5, 7)),        # identity
    ((8, 12, 9), (5, 9, 5)),       # down-sampling: source frames that are nobody's lower tap
    ((4, 10, 1), (5, 13, 1)),      # a single column (no division by the row length)
    ((5, 1, 40), (7, 1, 51)),      # a single row
    ((9, 6, 6), (11, 8, 8)),       # more than 8 source frames: not the register walk
])
@pytest.mark.parametrize("align", [True, False])
def test_resize_edge_shapes(hpvg_gpu, case, align):
    """Every dispatch branch of the resize launchers (constant-bank T walk, window-staged adjoint, tiled, generic)."""
    hp = hpvg_gpu
    in_size, out_size = case
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 1) + in_size).astype(np.float32)
    gy = rng.standard_normal((3, 1) + out_size).astype(np.float32)
    y = hp.ops.resize3d(hp.from_numpy(x), out_size, align_corners=align).numpy()
    ref = orc.resize_linear_np(x, out_size, align)
    assert np.max(np.abs(y - ref)) <= 1e-6
    gx = hp.ops.resize3d_bwd(hp.from_numpy(gy), in_size, align_corners=align).numpy()
    gref = orc.resize_linear_bwd_np(gy, in_size, align)
    assert rel_l2(gx, gref) < 1e-6
    # adjointness <R x, gy> == <x, R^T gy>
    lhs, rhs = float(np.sum(y.astype(np.float64) * gy)), float(np.sum(x.astype(np.float64) * gx))
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))


def test_resize_rejects_bad_sizes(hpvg_gpu):
    hp = hpvg_gpu
    x = hp.from_numpy(np.zeros((1, 1, 2, 2, 2), np.float32))
    with pytest.raises(hp.HpvgError):
        hp.ops.resize3d(x, (0, 4, 4))


def test_upsample_noise_pack(hpvg_gpu):
    hp = hpvg_gpu
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 3, 4, 24, 33)).astype(np.float32)
    nz = rng.standard_normal((2, 3, 4, 30, 41)).astype(np.float32)
    up, xin = hp.ops.upsample_noise_pack(hp.from_numpy(x), (4, 30, 41), noise=hp.from_numpy(nz), amp=0.37)
    ref_up = orc.resize_linear_np(x, (4, 30, 41), True)
    assert np.max(np.abs(up.numpy() - ref_up)) <= 1e-6
    xin_f = hp.ops.unpack_cl(xin, C=8).numpy()
    assert np.max(np.abs(xin_f[:, :3] - bf16_round(ref_up + nz * np.float32(0.37)))) <= 2e-2
    assert np.all(xin_f[:, 3:] == 0)


def test_device_philox_noise_statistics_and_shard_invariance(hpvg_gpu):
    """Internally drawn noise is keyed by (seed, sample index, element): splitting a batch over ranks must give
    the same per-sample noise."""
    hp = hpvg_gpu
    x = hp.from_numpy(np.zeros((4, 3, 4, 24, 33), np.float32))
    _, xin = hp.ops.upsample_noise_pack(x, (4, 30, 41), amp=1.0, seed=1234, sample_base=10)
    a = hp.ops.unpack_cl(xin, C=3).numpy()
    assert abs(a.mean()) < 0.02 and abs(a.std() - 1.0) < 0.02
    x2 = hp.from_numpy(np.zeros((2, 3, 4, 24, 33), np.float32))
    _, xin2 = hp.ops.upsample_noise_pack(x2, (4, 30, 41), amp=1.0, seed=1234, sample_base=12)
    b = hp.ops.unpack_cl(xin2, C=3).numpy()
    assert np.array_equal(a[2:], b)


def test_bn_train_cl(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(2)
    y = bf16_round(rng.standard_normal((2, 64, 3, 17, 13)) * 1.7 + 0.3)
    gamma = (1 + 0.1 * rng.standard_normal(64)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(64)).astype(np.float32)
    mm, mv = np.zeros(64, np.float32), np.ones(64, np.float32)
    tmm, tmv = hp.from_numpy(mm), hp.from_numpy(mv)
    x_cl, saved = ops.bn_train_cl(ops.pack_cl(hp.from_numpy(y)), hp.from_numpy(gamma), hp.from_numpy(beta), tmm, tmv)
    p = {"1.bn2d.gamma": torch.from_numpy(gamma), "1.bn2d.beta": torch.from_numpy(beta),
         "1.bn2d.moving_mean": torch.from_numpy(mm.copy()), "1.bn2d.moving_variance": torch.from_numpy(mv.copy())}
    ref = orc.lrelu(orc.batchnorm(torch.from_numpy(y), p, "1.", True)).numpy()
    assert rel_l2(ops.unpack_cl(x_cl).numpy(), ref) < 5e-3
    assert np.allclose(tmm.numpy(), p["1.bn2d.moving_mean"].numpy(), atol=1e-5)
    assert np.allclose(tmv.numpy(), p["1.bn2d.moving_variance"].numpy(), rtol=1e-4)


@pytest.mark.parametrize("shape", [(2, 3, 17, 13), (1, 4, 33, 41), (3, 1, 9, 70)])
def test_conv_epilogue_bn_statistics_and_one_pass_bn(hpvg_gpu, shape):
    """Training-mode BatchNorm with the batch statistics fused into the conv epilogue (ragged tiles included):
    sums must equal the sums over the stored bf16 output, and conv -> BN -> LeakyReLU must match the oracle."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(21)
    N, T, H, W = shape
    x = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    w = bf16_round(rng.standard_normal((64, 64, 3, 3, 3)) * 0.05)
    bias = (0.1 * rng.standard_normal(64)).astype(np.float32)
    gamma = (1 + 0.1 * rng.standard_normal(64)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(64)).astype(np.float32)
    mm, mv = np.zeros(64, np.float32), np.ones(64, np.float32)
    tmm, tmv = hp.from_numpy(mm), hp.from_numpy(mv)
    stats = hp.Tensor((2, 64), hp.F64).zero_()
    aff = ops.affine_from_bias(hp.from_numpy(bias))
    y_cl = ops.conv3d_cl_any(ops.pack_cl(hp.from_numpy(x)), hp.from_numpy(w), aff, ops.ACT_NONE, 64, 64, stats=stats)
    y = ops.unpack_cl(y_cl).numpy().astype(np.float64)
    st = stats.numpy()
    assert np.allclose(st[0], y.sum(axis=(0, 2, 3, 4)), rtol=1e-5, atol=1e-3)
    assert np.allclose(st[1], (y * y).sum(axis=(0, 2, 3, 4)), rtol=1e-5, atol=1e-3)
    saved = hp.Tensor((4, 64), hp.F32)
    a_cl = ops.bn_train_fused_cl(y_cl, stats, hp.from_numpy(gamma), hp.from_numpy(beta), tmm, tmv, saved=saved)
    p = {"0.weight": torch.from_numpy(w), "0.bias": torch.from_numpy(bias),
         "1.bn2d.gamma": torch.from_numpy(gamma), "1.bn2d.beta": torch.from_numpy(beta),
         "1.bn2d.moving_mean": torch.from_numpy(mm.copy()), "1.bn2d.moving_variance": torch.from_numpy(mv.copy())}
    ref = orc.conv_block(torch.from_numpy(x), p, "", 1, True).numpy()
    assert rel_l2(ops.unpack_cl(a_cl).numpy(), ref) < 5e-3
    assert np.allclose(tmm.numpy(), p["1.bn2d.moving_mean"].numpy(), atol=1e-3)
    assert np.allclose(tmv.numpy(), p["1.bn2d.moving_variance"].numpy(), rtol=5e-3)
    sv = saved.numpy()
    mean = y.mean(axis=(0, 2, 3, 4))
    var = y.var(axis=(0, 2, 3, 4))
    assert np.allclose(sv[2], mean, atol=1e-5) and np.allclose(sv[3], 1 / np.sqrt(var + 1e-5), rtol=1e-5)
    assert np.allclose(sv[0], gamma * sv[3], rtol=1e-6) and np.allclose(sv[1], beta - sv[2] * sv[0], atol=1e-5)


def test_sn_power_iter_multi_matches_single_layer_oracle(hpvg_gpu):
    """All SN layers of a network in one launch: per-layer sigma, u, v, the snapshot copies and the fused
    (1/sigma, bias) epilogue vectors."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(31)
    layers, refs = [], []
    for cin in (3, 64, 64):
        w = (rng.standard_normal((64, cin, 3, 3, 3)) * 0.02).astype(np.float32)
        u = orc._l2normalize_np(rng.standard_normal((64, 1)).astype(np.float32))
        v = orc._l2normalize_np(rng.standard_normal((cin * 27, 1)).astype(np.float32))
        b = rng.standard_normal(64).astype(np.float32)
        layers.append({"w": hp.from_numpy(w), "u": hp.from_numpy(u), "v": hp.from_numpy(v), "bias": hp.from_numpy(b),
                       "sigma": hp.Tensor((2,), hp.F32), "aff": hp.Tensor((2, 64), hp.F32),
                       "u_copy": hp.Tensor((64, 1), hp.F32), "v_copy": hp.Tensor((cin * 27, 1), hp.F32)})
        refs.append([w, torch.from_numpy(u), torch.from_numpy(v), b])
    for it in range(2):
        ops.sn_power_iter_multi(layers)
        for l, ref in zip(layers, refs):
            w, u_, v_, b = ref
            sigma, un, vn = orc.sn_power_iteration(torch.from_numpy(w), u_, v_)
            ref[1], ref[2] = un, vn
            sg = l["sigma"].numpy()
            assert abs(sg[0] - float(sigma)) / float(sigma) < 1e-5 and abs(sg[0] * sg[1] - 1) < 1e-6
            assert np.allclose(l["u"].numpy(), un.numpy(), atol=1e-5) and np.allclose(l["v"].numpy(), vn.numpy(), atol=1e-5)
            assert np.array_equal(l["u"].numpy(), l["u_copy"].numpy()) and np.array_equal(l["v"].numpy(), l["v_copy"].numpy())
            aff = l["aff"].numpy()
            assert np.allclose(aff[0], sg[1]) and np.array_equal(aff[1], b)


@pytest.mark.parametrize("cin", [3, 64])
def test_sn_power_iter(hpvg_gpu, cin):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(3)
    w = (rng.standard_normal((64, cin, 3, 3, 3)) * 0.02).astype(np.float32)
    u = orc._l2normalize_np(rng.standard_normal((64, 1)).astype(np.float32))
    v = orc._l2normalize_np(rng.standard_normal((cin * 27, 1)).astype(np.float32))
    tu, tv = hp.from_numpy(u), hp.from_numpy(v)
    tw = hp.from_numpy(w)
    tu_, tv_ = torch.from_numpy(u), torch.from_numpy(v)
    for _ in range(3):   # Q5: u/v advance on every call
        sg = ops.sn_power_iter(tw, tu, tv).numpy()
        sigma, tu_, tv_ = orc.sn_power_iteration(torch.from_numpy(w), tu_, tv_)
        assert abs(sg[0] - float(sigma)) / float(sigma) < 1e-5
        assert abs(sg[1] * sg[0] - 1.0) < 1e-6
    assert np.allclose(tu.numpy(), tu_.numpy(), atol=1e-5) and np.allclose(tv.numpy(), tv_.numpy(), atol=1e-5)


def test_losses_and_reparam(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(4)
    a = rng.standard_normal((1, 3, 4, 30, 41)).astype(np.float32)
    b = rng.standard_normal((1, 3, 4, 30, 41)).astype(np.float32)
    ta, tb = hp.from_numpy(a), hp.from_numpy(b)
    assert abs(ops.mse(ta, tb).numpy()[0] - float(((a - b) ** 2).mean())) < 1e-5
    assert abs(ops.mean(ta).numpy()[0] - float(a.mean())) < 1e-6
    mu = rng.standard_normal((1, 128, 4, 24, 33)).astype(np.float32)
    lv = (0.3 * rng.standard_normal((1, 128, 4, 24, 33))).astype(np.float32)
    ref = float(orc.kl_criterion(torch.from_numpy(mu), torch.from_numpy(lv)))
    assert abs(ops.kl_criterion(hp.from_numpy(mu), hp.from_numpy(lv)).numpy()[0] - ref) < 1e-5 * max(1, abs(ref))
    eps = rng.standard_normal(mu.shape).astype(np.float32)
    z = ops.reparam(hp.from_numpy(mu), hp.from_numpy(lv), hp.from_numpy(eps)).numpy()
    assert np.allclose(z, eps * np.exp(0.5 * lv) + mu, atol=1e-6, rtol=1e-5)


def test_adam_clip_multi(hpvg_gpu):
    """ClippedAdam (optimizers.py:33-43): per-tensor ClipByNorm(5) then Adam(beta1=.5), per-group lr."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(5)
    shapes = [(64, 64, 3, 3, 3), (64,), (3, 64, 3, 3, 3), (64, 3, 3, 3, 3)]
    scales = [0.5, 0.001, 2.0, 0.01]     # first and third exceed the clip norm
    lrs = [5e-4, 5e-4, 1e-4, 2e-5]
    w = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    g = [(rng.standard_normal(s) * sc).astype(np.float32) for s, sc in zip(shapes, scales)]
    m = [np.zeros(s, np.float32) for s in shapes]
    v = [np.zeros(s, np.float32) for s in shapes]
    tw, tg, tm, tv = ([hp.from_numpy(a) for a in lst] for lst in (w, g, m, v))
    for step in (1, 2, 3):
        ops.adam_clip_multi(tw, tg, tm, tv, lrs, step, beta1=0.5, beta2=0.999, eps=1e-8, clip=5.0)
        for i in range(len(shapes)):
            w[i], m[i], v[i] = orc.adam_step(w[i], orc.clip_by_norm(g[i], 5.0), m[i], v[i], step, lrs[i])
    for i in range(len(shapes)):
        assert rel_l2(tw[i].numpy(), w[i]) < 1e-6
        assert rel_l2(tm[i].numpy(), m[i]) < 1e-5 and rel_l2(tv[i].numpy(), v[i]) < 1e-4   # fp32 norm reduction order
    # plain Adam (discriminator optimiser, train_video.py:65): clip <= 0 disables clipping
    w0 = rng.standard_normal(1000).astype(np.float32)
    g0 = (rng.standard_normal(1000) * 3).astype(np.float32)
    t = [hp.from_numpy(a) for a in (w0, g0, np.zeros(1000, np.float32), np.zeros(1000, np.float32))]
    ops.adam_clip_multi([t[0]], [t[1]], [t[2]], [t[3]], [5e-4], 1, clip=0.0)
    ref, _, _ = orc.adam_step(w0, g0, np.zeros(1000, np.float32), np.zeros(1000, np.float32), 1, 5e-4)
    assert rel_l2(t[0].numpy(), ref) < 1e-6


def test_operators_match_committed_golden_fixtures(hpvg_gpu):
    """The CUDA operators against tests/golden/ops_small.npz (frozen oracle outputs; the fixture travels to the GPU box,
    /root/reference does not)."""
    import os
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ops_small.npz"))
    # resize: tables bit-exact, values within 1e-6, adjoint within 1e-6 rel-L2
    i0, i1, l0, l1 = ops.linear_taps(33, 41, True)
    assert np.array_equal(i0, g["rs_i0"]) and np.array_equal(i1, g["rs_i1"])
    assert np.array_equal(l0.view(np.uint32), g["rs_l0"].view(np.uint32))
    assert np.array_equal(l1.view(np.uint32), g["rs_l1"].view(np.uint32))
    d0, d1, dl0, dl1 = ops.linear_taps_device(33, 41, True)
    assert np.array_equal(d0, g["rs_i0"]) and np.array_equal(dl1.view(np.uint32), g["rs_l1"].view(np.uint32))
    y = ops.resize3d(hp.from_numpy(g["rs_x"]), (4, 30, 41)).numpy()
    assert np.abs(y - g["rs_y"]).max() <= 1e-6 and float(np.mean(y == g["rs_y"])) > 0.99
    gx = ops.resize3d_bwd(hp.from_numpy(g["rs_gy"]), (4, 24, 33)).numpy()
    assert rel_l2(gx, g["rs_gx"]) < 1e-6
    # BatchNorm (train) + LeakyReLU, moving statistics
    tmm, tmv = hp.from_numpy(np.zeros(64, np.float32)), hp.from_numpy(np.ones(64, np.float32))
    x_cl, _ = ops.bn_train_cl(ops.pack_cl(hp.from_numpy(g["bn_y"])), hp.from_numpy(g["bn_gamma"]),
                              hp.from_numpy(g["bn_beta"]), tmm, tmv)
    assert rel_l2(ops.unpack_cl(x_cl).numpy(), g["bn_out"]) < 5e-3           # bf16 output
    assert np.allclose(tmm.numpy(), g["bn_mm"], atol=1e-5) and np.allclose(tmv.numpy(), g["bn_mv"], rtol=1e-4)
    # spectral norm
    tu, tv = hp.from_numpy(g["sn_u"]), hp.from_numpy(g["sn_v"])
    sg = ops.sn_power_iter(hp.from_numpy(g["sn_w"]), tu, tv).numpy()
    assert abs(sg[0] - float(g["sn_sigma"])) / float(g["sn_sigma"]) < 1e-5
    assert np.allclose(tu.numpy(), g["sn_u1"], atol=1e-5) and np.allclose(tv.numpy(), g["sn_v1"], atol=1e-5)
    # losses
    mu, lv = hp.from_numpy(g["kl_mu"]), hp.from_numpy(g["kl_lv"])
    assert abs(ops.kl_criterion(mu, lv).numpy()[0] - float(g["kl"])) < 1e-5 * max(1.0, abs(float(g["kl"])))
    assert abs(ops.mse(mu, lv).numpy()[0] - float(g["mse"])) < 1e-5 * float(g["mse"])
    # ClipByNorm + Adam, two steps (second one unclipped)
    w = hp.from_numpy(g["ad_w0"])
    m, v = hp.Tensor(w.shape, hp.F32).zero_(), hp.Tensor(w.shape, hp.F32).zero_()
    ops.adam_clip_multi([w], [hp.from_numpy(g["ad_g1"])], [m], [v], [5e-4], 1, clip=5.0)
    assert np.allclose(w.numpy(), g["ad_w1"], atol=1e-6)
    ops.adam_clip_multi([w], [hp.from_numpy(g["ad_g2"])], [m], [v], [5e-4], 2, clip=5.0)
    assert np.allclose(w.numpy(), g["ad_w2"], atol=1e-6)
