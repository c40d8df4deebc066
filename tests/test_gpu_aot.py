"""The `Hpvg*` entry points are called here exactly the way MindSpore's ops.Custom(func_type="aot") runtime calls them
(INTEGRATION.md route A): `int f(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
void* stream, void* extra)` with caller-allocated device buffers, inputs first, outputs last, on the caller's stream;
0 = success, non-zero makes MindSpore raise.  MindSpore itself cannot be installed offline, so ctypes plays its part."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hpvg_oracle as orc
from util import bf16_round, rel_l2

pytestmark = pytest.mark.gpu


def _aot_call(hp, name, tensors, dtypes=None, stream=None):
    """tensors: device Tensors, inputs then outputs (what MindSpore passes in `params`)."""
    n = len(tensors)
    params = (ctypes.c_void_p * n)(*[ctypes.c_void_p(t.ptr) for t in tensors])
    ndims = (ctypes.c_int * n)(*[len(t.shape) for t in tensors])
    shape_arrays = [(ctypes.c_int64 * len(t.shape))(*t.shape) for t in tensors]
    shapes = (ctypes.POINTER(ctypes.c_int64) * n)(*[ctypes.cast(a, ctypes.POINTER(ctypes.c_int64)) for a in shape_arrays])
    names = (ctypes.c_char_p * n)(*[d.encode() for d in (dtypes or ["float32"] * n)])
    return getattr(hp.lib, name)(n, params, ndims, shapes, names, stream.handle if stream else None, None)


def test_aot_upsample_trilinear3d_forward_and_grad(hpvg_gpu):
    """Primitive UpsampleTrilinear3D (trilinear.py:171-254; io names x -> y) and its gradient as Custom ops."""
    hp = hpvg_gpu
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 4, 24, 33)).astype(np.float32)
    dy = rng.standard_normal((2, 3, 5, 30, 41)).astype(np.float32)
    st = hp.Stream()
    tx, ty = hp.from_numpy(x), hp.Tensor(dy.shape, hp.F32)
    assert _aot_call(hp, "HpvgUpsampleTrilinear3D", [tx, ty], stream=st) == 0
    st.sync()
    ref = orc.resize_linear_np(x, (5, 30, 41), True)
    assert np.max(np.abs(ty.numpy() - ref)) <= 1e-6
    tdy, tdx = hp.from_numpy(dy), hp.Tensor(x.shape, hp.F32)
    assert _aot_call(hp, "HpvgUpsampleTrilinear3DGrad", [tdy, tx, tdx], stream=st) == 0      # (dy, x) -> dx
    st.sync()
    assert rel_l2(tdx.numpy(), orc.resize_linear_bwd_np(dy, (4, 24, 33), True)) < 1e-6


def test_aot_conv3d_bias_lrelu(hpvg_gpu):
    """ConvBlock3D without BatchNorm (networks_3d.py:45-54) as ONE Custom op on fp32 NCDHW tensors."""
    hp = hpvg_gpu
    rng = np.random.default_rng(1)
    x = bf16_round(rng.standard_normal((1, 64, 4, 20, 27)))
    w = bf16_round(rng.standard_normal((64, 64, 3, 3, 3)) * 0.05)
    b = (rng.standard_normal(64) * 0.1).astype(np.float32)
    ty = hp.Tensor(x.shape, hp.F32)
    assert _aot_call(hp, "HpvgConv3dBiasLRelu", [hp.from_numpy(x), hp.from_numpy(w), hp.from_numpy(b), ty]) == 0
    hp.device_sync()
    with torch.no_grad():
        ref = F.leaky_relu(F.conv3d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), padding=1), 0.2).numpy()
    assert rel_l2(ty.numpy(), ref) < 4e-3


def test_aot_block_forward_and_bprop_wired_through_the_aot_convention(hpvg_gpu):
    """Route A end to end on one ConvBlock3DSN-shaped layer plus its neighbours: head conv (3 -> 64), 64 -> 64 conv +
    LeakyReLU, tail conv (64 -> 3) + tanh as forward Custom ops on fp32 NCDHW tensors, then the bprop of the middle layer
    ((x, w, y, dy) -> (dx, dw, db)) — every call through MindSpore's calling convention on a caller-owned stream, with
    no allocation or synchronisation per call (the second round of calls reuses the stream's workspace)."""
    hp = hpvg_gpu
    rng = np.random.default_rng(2)
    st = hp.Stream()
    x3 = bf16_round(np.tanh(rng.standard_normal((1, 3, 4, 20, 27))))
    w0 = bf16_round(rng.standard_normal((64, 3, 3, 3, 3)) * 0.2)
    w1 = bf16_round(rng.standard_normal((64, 64, 3, 3, 3)) * 0.05)
    w2 = bf16_round(rng.standard_normal((3, 64, 3, 3, 3)) * 0.05)
    b0, b1, b2 = (rng.standard_normal(n).astype(np.float32) * 0.1 for n in (64, 64, 3))
    dev = lambda a: hp.from_numpy(np.ascontiguousarray(a, np.float32))      # noqa: E731
    for _ in range(2):
        h0, h1 = hp.Tensor((1, 64, 4, 20, 27), hp.F32), hp.Tensor((1, 64, 4, 20, 27), hp.F32)
        out = hp.Tensor((1, 3, 4, 20, 27), hp.F32)
        assert _aot_call(hp, "HpvgConv3dBiasLRelu", [dev(x3), dev(w0), dev(b0), h0], stream=st) == 0
        assert _aot_call(hp, "HpvgConv3dBiasLRelu", [h0, dev(w1), dev(b1), h1], stream=st) == 0
        assert _aot_call(hp, "HpvgConv3dBiasTanh", [h1, dev(w2), dev(b2), out], stream=st) == 0
        st.sync()
    tx = torch.from_numpy(x3)
    r0 = F.leaky_relu(F.conv3d(tx, torch.from_numpy(w0), torch.from_numpy(b0), padding=1), 0.2)
    r0q = torch.from_numpy(bf16_round(r0.numpy())).requires_grad_(True)
    tw1, tb1 = torch.from_numpy(w1).requires_grad_(True), torch.from_numpy(b1).requires_grad_(True)
    r1 = F.leaky_relu(F.conv3d(r0q, tw1, tb1, padding=1), 0.2)
    rout = torch.tanh(F.conv3d(torch.from_numpy(bf16_round(r1.detach().numpy())), torch.from_numpy(w2),
                               torch.from_numpy(b2), padding=1))
    assert rel_l2(h0.numpy(), r0.numpy()) < 4e-3 and rel_l2(h1.numpy(), r1.detach().numpy()) < 6e-3
    assert rel_l2(out.numpy(), rout.numpy()) < 1e-2
    # bprop of the middle layer from an upstream gradient
    dy = bf16_round(rng.standard_normal((1, 64, 4, 20, 27)))
    r1.backward(torch.from_numpy(dy))
    dx, dw, db = hp.Tensor((1, 64, 4, 20, 27), hp.F32), hp.Tensor((64, 64, 3, 3, 3), hp.F32), hp.Tensor((64,), hp.F32)
    args = [dev(bf16_round(r0.numpy())), dev(w1), dev(r1.detach().numpy()), dev(dy), dx, dw, db]
    assert _aot_call(hp, "HpvgConv3dBiasLReluGrad", args, stream=st) == 0
    st.sync()
    assert rel_l2(dx.numpy(), r0q.grad.numpy()) < 1e-2
    assert rel_l2(dw.numpy(), tw1.grad.numpy()) < 1e-2
    assert rel_l2(db.numpy(), tb1.grad.numpy()) < 1e-2


def test_aot_batchnorm_spectral_norm_adam_and_losses(hpvg_gpu):
    """The rest of SURVEY §8b's per-op surface through the AOT convention: training-mode BatchNorm3d + LeakyReLU forward
    (moving statistics updated in place) and bprop, one spectral-norm power iteration, a ClippedAdam step with device-
    resident hyper-parameters, MSE and KL losses — each against the oracle."""
    hp = hpvg_gpu
    rng = np.random.default_rng(3)
    st = hp.Stream()
    dev = lambda a: hp.from_numpy(np.ascontiguousarray(a, np.float32))      # noqa: E731
    # ---- BatchNorm3d(train) + LeakyReLU
    x = bf16_round(rng.standard_normal((2, 64, 3, 10, 13)) * 1.5 + 0.3)
    gamma, beta = (1 + 0.1 * rng.standard_normal(64)).astype(np.float32), (0.1 * rng.standard_normal(64)).astype(np.float32)
    mm, mv = dev(np.zeros(64)), dev(np.ones(64))
    y, saved = hp.Tensor(x.shape, hp.F32), hp.Tensor((4, 64), hp.F32)
    assert _aot_call(hp, "HpvgBatchNorm3dLReluTrain", [dev(x), dev(gamma), dev(beta), mm, mv, y, saved], stream=st) == 0
    st.sync()
    tx = torch.from_numpy(x).requires_grad_(True)
    tg, tb = torch.from_numpy(gamma).requires_grad_(True), torch.from_numpy(beta).requires_grad_(True)
    p = {"1.bn2d.gamma": tg, "1.bn2d.beta": tb, "1.bn2d.moving_mean": torch.zeros(64), "1.bn2d.moving_variance": torch.ones(64)}
    ref = orc.lrelu(orc.batchnorm(tx, p, "1.", True))
    assert rel_l2(y.numpy(), ref.detach().numpy()) < 5e-3
    assert rel_l2(mm.numpy(), p["1.bn2d.moving_mean"].numpy()) < 1e-3
    assert rel_l2(mv.numpy(), p["1.bn2d.moving_variance"].numpy()) < 1e-3
    dy = bf16_round(rng.standard_normal(x.shape))
    ref.backward(torch.from_numpy(dy))
    dx, dg, db = hp.Tensor(x.shape, hp.F32), hp.Tensor((64,), hp.F32), hp.Tensor((64,), hp.F32)
    assert _aot_call(hp, "HpvgBatchNorm3dLReluTrainGrad", [dev(dy), dev(x), saved, dx, dg, db], stream=st) == 0
    st.sync()
    assert rel_l2(dx.numpy(), tx.grad.numpy()) < 1e-2
    assert rel_l2(dg.numpy(), tg.grad.numpy()) < 5e-3 and rel_l2(db.numpy(), tb.grad.numpy()) < 5e-3
    # ---- spectral norm: one power iteration
    w = (rng.standard_normal((64, 64, 3, 3, 3)) * 0.02).astype(np.float32)
    u = orc._l2normalize_np(rng.standard_normal((64, 1)).astype(np.float32))
    v = orc._l2normalize_np(rng.standard_normal((1728, 1)).astype(np.float32))
    s2, un, vn = hp.Tensor((2,), hp.F32), hp.Tensor((64, 1), hp.F32), hp.Tensor((1728, 1), hp.F32)
    assert _aot_call(hp, "HpvgSpectralNormIter", [dev(w), dev(u), dev(v), s2, un, vn], stream=st) == 0
    st.sync()
    sigma, ur, vr = orc.sn_power_iteration(torch.from_numpy(w), torch.from_numpy(u), torch.from_numpy(v))
    assert abs(s2.numpy()[0] - float(sigma)) < 1e-4 * float(sigma) and abs(s2.numpy()[0] * s2.numpy()[1] - 1) < 1e-5
    assert np.allclose(un.numpy(), ur.numpy(), atol=1e-5) and np.allclose(vn.numpy(), vr.numpy(), atol=1e-5)
    # ---- ClippedAdam, third step, clip 5
    pw, g = rng.standard_normal(5000).astype(np.float32), (rng.standard_normal(5000) * 3).astype(np.float32)
    m0, v0 = (rng.standard_normal(5000) * 0.1).astype(np.float32), (rng.random(5000) * 0.1).astype(np.float32)
    hyper = np.array([5e-4, 0.5, 0.999, 1e-8, 5.0, 3.0], np.float32)
    pn, mn, vn2 = (hp.Tensor((5000,), hp.F32) for _ in range(3))
    assert _aot_call(hp, "HpvgClipAdam", [dev(pw), dev(g), dev(m0), dev(v0), dev(hyper), pn, mn, vn2], stream=st) == 0
    st.sync()
    wr, mr, vr2 = orc.adam_step(pw, orc.clip_by_norm(g, 5.0), m0, v0, 3, 5e-4, 0.5, 0.999)
    assert rel_l2(pn.numpy(), wr) < 1e-6 and rel_l2(mn.numpy(), mr) < 1e-6 and rel_l2(vn2.numpy(), vr2) < 1e-6
    # ---- losses
    a, b = rng.standard_normal((1, 3, 4, 9, 11)).astype(np.float32), rng.standard_normal((1, 3, 4, 9, 11)).astype(np.float32)
    out = hp.Tensor((1,), hp.F32)
    assert _aot_call(hp, "HpvgMSELoss", [dev(a), dev(b), out], stream=st) == 0
    st.sync()
    assert abs(out.numpy()[0] - float(((a - b) ** 2).mean())) < 1e-5
    assert _aot_call(hp, "HpvgKLLoss", [dev(a), dev(0.3 * b), out], stream=st) == 0
    st.sync()
    assert abs(out.numpy()[0] - float(orc.kl_criterion(torch.from_numpy(a), torch.from_numpy(0.3 * b)))) < 1e-4


def test_per_stream_scratch_allows_concurrent_reductions(hpvg_gpu):
    """SURVEY §8b thread-safety: scratch is per stream, so two host threads driving two streams get the same results as
    one thread running the calls one after the other (BatchNorm backward + weight gradient: both use reduction scratch)."""
    import threading
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(4)
    shapes = [(1, 64, 4, 30, 41), (1, 64, 3, 24, 33)]
    data = [(bf16_round(rng.standard_normal(s)), bf16_round(rng.standard_normal(s))) for s in shapes]

    def work(i, stream, out):
        x, gy = data[i]
        # every tensor stays referenced until the stream is drained: the caching allocator hands a freed block to the next
        # requester at once, which is only safe within ONE stream (hpvg/runtime.py)
        xd, gd = hp.from_numpy(x, stream=stream), hp.from_numpy(gy, stream=stream)
        xc, gc = ops.pack_cl(xd, stream=stream), ops.pack_cl(gd, stream=stream)
        res = []
        for _ in range(20):
            dw = hp.Tensor((64, 64, 3, 3, 3), hp.F32)
            ops.conv_wgrad_cl(xc, gc, dw, stream=stream)
            cs = hp.Tensor((64,), hp.F32)
            ops.colsum_cl(gc, cs, stream=stream)
            res.append((dw, cs))
        stream.sync()
        out[i] = [(a.numpy(stream), b.numpy(stream)) for a, b in res]

    serial, conc = {}, {}
    s0, s1 = hp.Stream(), hp.Stream()
    work(0, s0, serial)
    work(1, s1, serial)
    th = [threading.Thread(target=work, args=(i, s, conc)) for i, s in ((0, s0), (1, s1))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for i in (0, 1):
        for (dw_a, cs_a), (dw_b, cs_b) in zip(serial[i], conc[i]):
            assert np.array_equal(dw_a, dw_b)                   # fixed-order reduction: bit-identical
            assert np.allclose(cs_a, cs_b, rtol=1e-6, atol=1e-6)   # fp64 atomics, rounded to fp32
        assert np.array_equal(serial[i][0][0], serial[i][-1][0])        # and each is reproducible run to run


def test_aot_entries_reject_what_mindspore_must_not_pass(hpvg_gpu):
    """Non-zero return (-> RuntimeError inside MindSpore) instead of a crash: wrong arity, rank, dtype, channel count."""
    hp = hpvg_gpu
    x5 = hp.Tensor((1, 3, 2, 4, 4), hp.F32)
    y5 = hp.Tensor((1, 3, 2, 8, 8), hp.F32)
    x4 = hp.Tensor((1, 3, 4, 4), hp.F32)
    assert _aot_call(hp, "HpvgUpsampleTrilinear3D", [x5]) != 0                                   # arity
    assert _aot_call(hp, "HpvgUpsampleTrilinear3D", [x4, y5]) != 0                               # rank
    assert _aot_call(hp, "HpvgUpsampleTrilinear3D", [x5, y5], dtypes=["float16", "float32"]) != 0   # dtype
    assert _aot_call(hp, "HpvgUpsampleTrilinear3D", [x5, hp.Tensor((2, 3, 2, 8, 8), hp.F32)]) != 0   # batch mismatch
    w = hp.Tensor((64, 64, 3, 3, 3), hp.F32)
    b = hp.Tensor((64,), hp.F32)
    x32 = hp.Tensor((1, 32, 2, 4, 4), hp.F32)
    assert _aot_call(hp, "HpvgConv3dBiasLRelu", [x32, w, b, hp.Tensor(x32.shape, hp.F32)]) != 0   # Cin != 64
