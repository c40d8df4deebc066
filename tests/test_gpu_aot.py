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
