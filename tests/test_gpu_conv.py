"""Parity of the tcgen05 implicit-GEMM convolution (through the C ABI) against the torch-CPU oracle.

Tolerance: north_star states rel-L2 <= 1e-2 for bf16 compute.  Inputs/weights are pre-rounded to bf16 so the only
differences are accumulation order (fp32) and the bf16 rounding of the output — we assert 4e-3."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hpvg_oracle as orc
from util import bf16_round, rel_l2

pytestmark = pytest.mark.gpu
TOL = 4e-3


def _conv_ref(x, w, b, act=None):
    y = F.conv3d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), padding=1)
    if act == "lrelu":
        y = F.leaky_relu(y, 0.2)
    if act == "tanh":
        y = torch.tanh(y)
    return y.numpy()


@pytest.mark.parametrize("shape", [(1, 4, 24, 33), (2, 5, 16, 8), (1, 1, 7, 5), (1, 3, 40, 17), (1, 2, 33, 9)])
def test_conv_64_64(hpvg_gpu, shape):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    N, T, H, W = shape
    rng = np.random.default_rng(1)
    x = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    w = bf16_round(rng.standard_normal((64, 64, 3, 3, 3)) * 0.05)
    b = rng.standard_normal(64).astype(np.float32) * 0.1
    x_cl = ops.pack_cl(hp.from_numpy(x))
    aff = ops.affine_from_bias(hp.from_numpy(b))
    y_cl = ops.conv3d_cl_any(x_cl, hp.from_numpy(w), aff, ops.ACT_LRELU, 64, 64)
    y = ops.unpack_cl(y_cl).numpy()
    ref = _conv_ref(x, w, b, "lrelu")
    err = rel_l2(y, ref)
    assert err < TOL, "conv 64->64 %s rel-L2 %.3e" % (shape, err)


def test_conv_64_64_roundtrip_pack(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(2)
    x = bf16_round(rng.standard_normal((2, 64, 3, 9, 11)))
    back = ops.unpack_cl(ops.pack_cl(hp.from_numpy(x))).numpy()
    assert np.array_equal(back, x)


@pytest.mark.parametrize("C,pitch,off", [(64, 64, 0), (128, 128, 0), (64, 128, 64), (20, 24, 0), (72, 80, 0),
                                         (12, 64, 8), (3, 8, 0), (3, 64, 0)])
def test_pack_unpack_layouts(hpvg_gpu, C, pitch, off):
    """NCDHW fp32 <-> channels-last bf16 at the API edge: every kernel variant (shared-memory tile for wide tensors,
    skinny, direct), channel offsets inside a wider pitch, ragged channel counts, voxel counts off the tile size, and
    samples that straddle a tile.  Exact: the values are bf16-representable."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(C * 7 + pitch + off)
    x = bf16_round(rng.standard_normal((3, C, 2, 7, 13)))          # 182 voxels per sample: not a multiple of 64
    cl = hp.Tensor((3, 2, 7, 13, pitch), hp.BF16)
    cl.copy_from_host(np.full((3, 2, 7, 13, pitch), 0x3F80, np.uint16))   # 1.0 everywhere: untouched channels must stay
    ops.pack_cl(hp.from_numpy(x), c_pitch=pitch, c_off=off, out=cl)
    raw = hp.bf16_bits_to_f32(cl.numpy())
    assert np.array_equal(np.moveaxis(raw[..., off:off + C], -1, 1), x)
    lo, hi = off + C, min(pitch, (off + C + 7) // 8 * 8)                 # zero padding up to the 8-channel group
    assert not raw[..., lo:hi].any()
    assert np.all(raw[..., :off] == 1.0) and np.all(raw[..., hi:] == 1.0)
    assert np.array_equal(ops.unpack_cl(cl, C=C, c_off=off).numpy(), x)


@pytest.mark.parametrize("cout", [3, 1])
def test_conv_tail(hpvg_gpu, cout):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    N, T, H, W = 1, 4, 30, 41
    rng = np.random.default_rng(3)
    x = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    w = bf16_round(rng.standard_normal((cout, 64, 3, 3, 3)) * 0.05)
    b = rng.standard_normal(cout).astype(np.float32) * 0.1
    res = rng.standard_normal((N, cout, T, H, W)).astype(np.float32) * 0.3
    x_cl = ops.pack_cl(hp.from_numpy(x))
    aff = ops.affine_from_bias(hp.from_numpy(b))
    y = ops.conv3d_cl_any(x_cl, hp.from_numpy(w), aff, ops.ACT_TANH, 64, cout, residual=hp.from_numpy(res)).numpy()
    ref = np.tanh(_conv_ref(x, w, b) + res)
    err = rel_l2(y, ref)
    assert err < TOL, "tail conv 64->%d rel-L2 %.3e" % (cout, err)


@pytest.mark.parametrize("shape", [(3, 1, 17, 9), (1, 13, 35, 70), (2, 2, 16, 8), (1, 5, 1, 1)])
def test_conv_tail_folded_taps_ragged_shapes_and_legacy_variant(hpvg_gpu, shape):
    """The 64->3 tail kernel with the in-plane taps folded into N (conv3d_tail.cu): ragged tiles, T = 1, single-voxel
    planes, batch > 1; and it must agree with the generic skinny variant (HPVG_CONV_64_16) it replaces."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    N, T, H, W = shape
    rng = np.random.default_rng(13)
    x = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    w = bf16_round(rng.standard_normal((3, 64, 3, 3, 3)) * 0.05)
    b = rng.standard_normal(3).astype(np.float32) * 0.1
    x_cl = ops.pack_cl(hp.from_numpy(x))
    aff = ops.affine_from_bias(hp.from_numpy(b))
    tw = hp.from_numpy(w)
    y = ops.conv3d_cl_any(x_cl, tw, aff, ops.ACT_NONE, 64, 3).numpy()
    ref = _conv_ref(x, w, b)
    assert rel_l2(y, ref) < TOL
    s, sh = aff, aff.view((64,), hp.F32, 256)
    legacy = ops.conv_cl(ops.CONV_64_16, x_cl, ops.pack_weights(tw, ops.CONV_64_16), s, sh, ops.ACT_NONE,
                         ops.OUT_F32_NCDHW, cout_real=3).numpy()
    assert np.abs(y - legacy).max() < 1e-4 * max(1.0, np.abs(ref).max())
    # data-gradient filter bank (roles of Cin/Cout swapped, taps mirrored): dgrad of a 3 -> 64 head conv
    wh = bf16_round(rng.standard_normal((64, 3, 3, 3, 3)) * 0.05)
    gy = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    xt = torch.zeros((N, 3, T, H, W), requires_grad=True)
    F.conv3d(xt, torch.from_numpy(wh), padding=1).backward(torch.from_numpy(gy))
    unit = hp.from_numpy(np.concatenate([np.ones(64, np.float32), np.zeros(64, np.float32)]).reshape(2, 64))
    dx = ops.conv_cl(ops.CONV_64_3, ops.pack_cl(hp.from_numpy(gy)),
                     ops.pack_weights(hp.from_numpy(wh), ops.CONV_64_3, True, cout=3), unit, unit.view((64,), hp.F32, 256),
                     ops.ACT_NONE, ops.OUT_F32_NCDHW, cout_real=3).numpy()
    assert rel_l2(dx, xt.grad.numpy()) < TOL


def test_conv_epilogue_lrelu_backward_mask(hpvg_gpu):
    """HPVG_ACT_LRELU_MASK: the data-gradient conv multiplies its result by LeakyReLU'(stored activation), i.e. emits
    the next (lower) layer's pre-activation gradient directly — must equal dgrad followed by hpvg_lrelu_bwd_cl."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(15)
    N, T, H, W = 2, 3, 19, 21
    gy = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    a = bf16_round(rng.standard_normal((N, 64, T, H, W)))            # stored activation of the layer below
    w = bf16_round(rng.standard_normal((64, 64, 3, 3, 3)) * 0.05)
    unit = hp.from_numpy(np.concatenate([np.ones(64, np.float32), np.zeros(64, np.float32)]).reshape(2, 64))
    tw, gy_cl, a_cl = hp.from_numpy(w), ops.pack_cl(hp.from_numpy(gy)), ops.pack_cl(hp.from_numpy(a))
    plain = ops.conv3d_cl_any(gy_cl, tw, unit, ops.ACT_NONE, 64, 64, transpose_flip=True)
    want = ops.unpack_cl(ops.lrelu_bwd_cl(plain, a_cl)).numpy()
    got = ops.unpack_cl(ops.conv3d_cl_any(gy_cl, tw, unit, ops.ACT_LRELU_MASK, 64, 64, transpose_flip=True,
                                          mask=a_cl)).numpy()
    # one bf16 rounding fewer on the fused path (the unfused one rounds dgrad, then rounds again after the mask)
    assert rel_l2(got, want) < 3e-3
    xt = torch.from_numpy(a).clone().requires_grad_(True)    # autograd: d/dx of conv(lrelu(x)) at lrelu output a
    pre = torch.from_numpy(a).clone().requires_grad_(True)
    y = F.conv3d(F.leaky_relu(pre, 0.2), torch.from_numpy(w), padding=1)
    y.backward(torch.from_numpy(gy))
    assert rel_l2(got, pre.grad.numpy()) < TOL


def test_conv_tail_2d(hpvg_gpu):
    """Conv2d(N, nc_im) tail of the image path (networks_2d.py:212): 3x3 filter in the centre temporal tap."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(14)
    x = bf16_round(rng.standard_normal((2, 64, 1, 29, 39)))
    w = bf16_round(rng.standard_normal((3, 64, 3, 3)) * 0.05)
    b = rng.standard_normal(3).astype(np.float32) * 0.1
    aff = ops.affine_from_bias(hp.from_numpy(b))
    y = ops.conv3d_cl_any(ops.pack_cl(hp.from_numpy(x)), hp.from_numpy(w), aff, ops.ACT_TANH, 64, 3).numpy()
    ref = np.tanh(F.conv2d(torch.from_numpy(x[:, :, 0]), torch.from_numpy(w), torch.from_numpy(b), padding=1).numpy())
    assert rel_l2(y[:, :, 0], ref) < TOL


def test_conv_head_3_64(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    N, T, H, W = 2, 4, 24, 33
    rng = np.random.default_rng(4)
    x = bf16_round(rng.standard_normal((N, 3, T, H, W)))
    w = bf16_round(rng.standard_normal((64, 3, 3, 3, 3)) * 0.2)
    b = rng.standard_normal(64).astype(np.float32) * 0.1
    x_cl = ops.pack_cl(hp.from_numpy(x), c_pitch=8)
    aff = ops.affine_from_bias(hp.from_numpy(b))
    y = ops.unpack_cl(ops.conv3d_cl_any(x_cl, hp.from_numpy(w), aff, ops.ACT_LRELU, 3, 64)).numpy()
    ref = _conv_ref(x, w, b, "lrelu")
    err = rel_l2(y, ref)
    assert err < TOL, "head conv 3->64 rel-L2 %.3e" % err


def test_conv_128_64_and_64_128(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    N, T, H, W = 1, 4, 24, 33
    rng = np.random.default_rng(5)
    x = bf16_round(rng.standard_normal((N, 128, T, H, W)))
    w = bf16_round(rng.standard_normal((64, 128, 3, 3, 3)) * 0.04)
    b = rng.standard_normal(64).astype(np.float32) * 0.1
    x_cl = ops.pack_cl(hp.from_numpy(x))
    aff = ops.affine_from_bias(hp.from_numpy(b))
    y = ops.unpack_cl(ops.conv3d_cl_any(x_cl, hp.from_numpy(w), aff, ops.ACT_NONE, 128, 64)).numpy()
    err = rel_l2(y, _conv_ref(x, w, b))
    assert err < TOL, "conv 128->64 rel-L2 %.3e" % err
    x2 = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    w2 = bf16_round(rng.standard_normal((128, 64, 3, 3, 3)) * 0.05)
    b2 = rng.standard_normal(128).astype(np.float32) * 0.1
    aff2 = ops.affine_from_bias(hp.from_numpy(b2))
    y2 = ops.unpack_cl(ops.conv3d_cl_any(ops.pack_cl(hp.from_numpy(x2)), hp.from_numpy(w2), aff2, ops.ACT_NONE, 64, 128)).numpy()
    err2 = rel_l2(y2, _conv_ref(x2, w2, b2))
    assert err2 < TOL, "conv 64->128 rel-L2 %.3e" % err2


def test_conv_dgrad_matches_autograd(hpvg_gpu):
    """Data gradient = the same kernel with the transposed/mirrored filter bank."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    N, T, H, W = 1, 3, 20, 13
    rng = np.random.default_rng(6)
    gy = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    w = bf16_round(rng.standard_normal((64, 64, 3, 3, 3)) * 0.05)
    x = torch.zeros((N, 64, T, H, W), requires_grad=True)
    F.conv3d(x, torch.from_numpy(w), None, padding=1).backward(torch.from_numpy(gy))
    zero = hp.from_numpy(np.zeros(64, np.float32))
    aff = ops.affine_from_bias(zero)
    gx = ops.unpack_cl(ops.conv3d_cl_any(ops.pack_cl(hp.from_numpy(gy)), hp.from_numpy(w), aff, ops.ACT_NONE, 64, 64,
                                         transpose_flip=True)).numpy()
    err = rel_l2(gx, x.grad.numpy())
    assert err < TOL, "dgrad rel-L2 %.3e" % err


def test_conv2d_is_t1(hpvg_gpu):
    """Conv2d (networks_2d.py) = the T == 1 case with a (Cout, Cin, 3, 3) filter."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(7)
    x = bf16_round(rng.standard_normal((1, 64, 1, 39, 46)))
    w = bf16_round(rng.standard_normal((64, 64, 3, 3)) * 0.05)
    b = rng.standard_normal(64).astype(np.float32) * 0.1
    aff = ops.affine_from_bias(hp.from_numpy(b))
    y = ops.unpack_cl(ops.conv3d_cl_any(ops.pack_cl(hp.from_numpy(x)), hp.from_numpy(w), aff, ops.ACT_NONE, 64, 64)).numpy()
    ref = F.conv2d(torch.from_numpy(x[:, :, 0]), torch.from_numpy(w), torch.from_numpy(b), padding=1).numpy()[:, :, None]
    err = rel_l2(y, ref)
    assert err < TOL, "conv2d rel-L2 %.3e" % err


def test_conv_empty_input_is_noop(hpvg_gpu):
    hp = hpvg_gpu
    assert hp.lib.hpvg_conv_cl(0, 0, 4, 8, 8, None, 64, None, None, None, 0, 0, None, 64, 0, 64, None, None, None, 0, None) == 0


@pytest.mark.parametrize("nd", [3, 2])
def test_reflect_pad_branch_of_convblock_sn(hpvg_gpu, nd):
    """ConvBlock3DSN / ConvBlock2DSN(bn=False): nn.Pad(REFLECT) + pad_mode='valid' conv (reference networks_3d.py:64-73,
    networks_2d.py:63-72; reached through FeatureExtractor(return_linear=True)).  Pad bit-exact, conv within the bf16
    tolerance of the fp32 oracle."""
    hp = hpvg_gpu
    import torch
    from hpvg import networks_2d as n2, networks_3d as n3, ops
    rng = np.random.default_rng(21 + nd)
    N, T, H, W = (2, 4, 11, 13) if nd == 3 else (2, 1, 14, 9)
    x = rng.standard_normal((N, 64, T, H, W)).astype(np.float32)
    x_cl = ops.pack_cl(hp.from_numpy(x))
    # the pad itself, against numpy
    xp = ops.reflect_pad_cl(x_cl, pad_t=1 if nd == 3 else 0, pad_hw=1)
    got = ops.unpack_cl(xp).numpy()
    pads = ((0, 0), (0, 0), (1, 1) if nd == 3 else (0, 0), (1, 1), (1, 1))
    ref = np.pad(ops.unpack_cl(x_cl).numpy(), pads, mode="reflect")
    assert np.array_equal(got, ref)
    # the cell
    mk = n3.ConvBlock3DSN if nd == 3 else n2.ConvBlock2DSN
    cell = mk(64, 64, 3, 1, 1, bn=False, act="lrelu", rng=rng)
    names = sorted(cell.parameters_dict())
    assert names == (["1.weight"] if nd == 3 else ["1.bias", "1.weight"])
    w = cell.p["weight"].numpy()
    b = rng.standard_normal(64).astype(np.float32) * 0.1 if nd == 2 else None
    if b is not None:
        cell.load_parameters({"1.weight": w, "1.bias": b})
    y = ops.unpack_cl(cell.forward_cl(x_cl)).numpy()
    xt = torch.from_numpy(x if nd == 3 else x[:, :, 0])
    with torch.no_grad():
        yr = orc.reflect_conv(xt, torch.from_numpy(w), None if b is None else torch.from_numpy(b), act=True).numpy()
    if nd == 2:
        yr = yr[:, :, None]
    assert y.shape == yr.shape
    assert rel_l2(y, yr) < 1e-2, rel_l2(y, yr)
    # FeatureExtractor(return_linear=True): last block is that cell without activation
    fe = (n3.FeatureExtractor if nd == 3 else n2.FeatureExtractor)(64, 64, 3, 1, 1, num_blocks=2, return_linear=True, rng=rng)
    assert isinstance(fe.layers[-1], n3.ReflectConvLayer) and fe.layers[-1].act == ops.ACT_NONE
    assert fe.construct_cl(x_cl).shape == x_cl.shape


def _random_shapes(seed, n):
    """Shapes between the tiny cases above and the full size: several output planes per CTA pair, work ranges that start
    and end inside a strip, ragged last tiles, T from 1 up."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        out.append((int(rng.integers(1, 4)), int(rng.integers(1, 8)), int(rng.integers(8, 97)), int(rng.integers(8, 120))))
    return out


@pytest.mark.parametrize("shape", _random_shapes(2024, 10) + [(1, 2, 160, 200), (5, 3, 64, 64)])
def test_conv_variants_on_mid_sized_random_shapes(hpvg_gpu, shape):
    """64->64 (with and without the fused BatchNorm statistics), head 3->64 and tail 64->3 against the torch-CPU oracle
    on shapes where a CTA pair owns several output planes (the work partition cuts strips mid-way; the MMA issuer's
    wait-behind-the-MMAs order, the filter-bank fetch order and the two epilogue groups all see that case)."""
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    N, T, H, W = shape
    rng = np.random.default_rng(abs(hash(shape)) % (2 ** 31))
    x = bf16_round(rng.standard_normal((N, 64, T, H, W)))
    w = bf16_round(rng.standard_normal((64, 64, 3, 3, 3)) * 0.05)
    b = rng.standard_normal(64).astype(np.float32) * 0.1
    x_cl = ops.pack_cl(hp.from_numpy(x))
    aff = ops.affine_from_bias(hp.from_numpy(b))
    ref = _conv_ref(x, w, b, "lrelu")
    y = ops.unpack_cl(ops.conv3d_cl_any(x_cl, hp.from_numpy(w), aff, ops.ACT_LRELU, 64, 64)).numpy()
    assert rel_l2(y, ref) < TOL, "conv 64->64 %s rel-L2 %.3e" % (shape, rel_l2(y, ref))
    # the statistics variant: same output, sums of the values as stored
    stats = hp.Tensor((2, 64), hp.F64).zero_()
    y_cl = ops.conv3d_cl_any(x_cl, hp.from_numpy(w), aff, ops.ACT_NONE, 64, 64, stats=stats)
    ys = ops.unpack_cl(y_cl).numpy()
    s = stats.numpy()
    assert np.allclose(s[0], ys.sum(axis=(0, 2, 3, 4), dtype=np.float64), rtol=1e-6, atol=1e-3)
    assert np.allclose(s[1], (ys.astype(np.float64) ** 2).sum(axis=(0, 2, 3, 4)), rtol=1e-6, atol=1e-3)
    assert rel_l2(ys, _conv_ref(x, w, b)) < TOL
    # head 3 -> 64
    x3 = bf16_round(rng.standard_normal((N, 3, T, H, W)))
    w3 = bf16_round(rng.standard_normal((64, 3, 3, 3, 3)) * 0.1)
    x3_cl = ops.pack_cl(hp.from_numpy(x3))
    y3 = ops.unpack_cl(ops.conv3d_cl_any(x3_cl, hp.from_numpy(w3), aff, ops.ACT_LRELU, 3, 64)).numpy()
    assert rel_l2(y3, _conv_ref(x3, w3, b, "lrelu")) < TOL, "head %s" % (shape,)
    # tail 64 -> 3 (+ residual, tanh)
    wt = bf16_round(rng.standard_normal((3, 64, 3, 3, 3)) * 0.05)
    bt = rng.standard_normal(3).astype(np.float32) * 0.1
    res = rng.standard_normal((N, 3, T, H, W)).astype(np.float32) * 0.3
    yt = ops.conv3d_cl_any(x_cl, hp.from_numpy(wt), ops.affine_from_bias(hp.from_numpy(bt)), ops.ACT_TANH, 64, 3,
                           residual=hp.from_numpy(res)).numpy()
    reft = np.tanh(_conv_ref(x, wt, bt) + res)
    assert rel_l2(yt, reft) < TOL, "tail %s" % (shape,)
