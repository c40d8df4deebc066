"""Functional operators over libhpvg's C ABI: one Python function per reference library-op call site
(SURVEY.md §2.2).  Tensors are `hpvg.runtime.Tensor` (device memory); fp32 tensors use the reference's
NCDHW layout, "cl" tensors are the kernels' internal channels-last bf16 layout."""
import ctypes
import os

import numpy as np

from . import runtime as rt
from ._lib import HpvgError, check, lib
from .runtime import BF16, F32, F64, I32, Tensor, _s

CONV_64_64, CONV_64_16, CONV_8_64, CONV_64_3 = 0, 1, 2, 3
CONV_T32_64, CONV_T4_64, CONV_T32_3 = 4, 5, 6      # kind::tf32 variants over fp32 channels-last activations
ACT_NONE, ACT_LRELU, ACT_TANH, ACT_LRELU_MASK = 0, 1, 2, 3
OUT_BF16_CL, OUT_F32_NCDHW, OUT_F32_RAW, OUT_F32_CL = 0, 1, 2, 3
BN_EPS = 1e-5        # mindspore.nn.BatchNorm3d default
BN_MOMENTUM = 0.9    # mindspore.nn.BatchNorm3d default (moving = 0.9*moving + 0.1*batch)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.ptr)


# ------------------------------------------------------------------------------------------------ precision mode
# "bf16": activations between convs are channels-last bf16, tcgen05 kind::f16 (north_star's 1e-2 branch; the fast path).
# "tf32": activations are channels-last fp32 (values rounded to tf32), tcgen05 kind::tf32 — the fp32-accurate path
#         (north_star's 1e-3 branch; the reference computes in fp32, networks_3d.py:48-50).
# The mode decides the dtype of NEW channels-last tensors; every operator dispatches on the dtype of its operands.
_PRECISION = [os.environ.get("HPVG_PRECISION", "bf16").lower()]


def set_precision(mode):
    mode = str(mode).lower()
    if mode not in ("bf16", "tf32"):
        raise HpvgError("precision must be 'bf16' or 'tf32', got %r" % (mode,))
    _PRECISION[0] = mode


def precision():
    return _PRECISION[0]


def cl_dtype():
    """dtype of channels-last activation tensors in the current precision mode."""
    return F32 if _PRECISION[0] == "tf32" else BF16


def narrow_pitch(dtype=None):
    """Channels per voxel of a 'narrow' (16-byte) channels-last tensor: block inputs, 3-channel gradients."""
    return 4 if (dtype or cl_dtype()) == F32 else 8


def _isz(t):
    return 4 if t.dtype == F32 else 2


# ------------------------------------------------------------------------------------------------ in-situ kernel timing
_prof = None
_flops = None


def start_profile(mode):
    """Record a CUDA-event pair around every conv launch of kernel variant `mode` (bench.py's roofline leg).
    mode="all": every convolution AND weight-gradient launch, tagged (see stop_profile)."""
    global _prof
    _prof = {"mode": mode, "items": []}


def stop_profile():
    """-> [(voxels, milliseconds)] for the launches recorded since start_profile(mode);
    for mode="all": [(kind, mode, voxels, milliseconds)] with kind in {"conv", "wgrad"}."""
    global _prof
    items, tagged = (_prof["items"] if _prof else []), bool(_prof and _prof["mode"] == "all")
    _prof = None
    out = []
    for tag, e0, e1 in items:
        e1.sync()
        out.append(tag + (e0.elapsed_ms(e1),) if tagged else (tag, e0.elapsed_ms(e1)))
    return out


def start_flop_count():
    """Count the ALGORITHMIC FLOP of every convolution / data-gradient / weight-gradient call from here on:
    2 * taps * Cin * Cout per output voxel with the layer's real channel counts (3 for the clip-side convs, not the
    padded 8 / 64) and taps = 27 (3-D) or 9 (2-D)."""
    global _flops
    _flops = {"conv": 0, "wgrad": 0, "conv_calls": 0, "wgrad_calls": 0}


def stop_flop_count():
    global _flops
    out, _flops = _flops, None
    return out


# ------------------------------------------------------------------------------------------------ layout
def pack_cl(x, c_pitch=None, c_off=0, zero_to=None, out=None, stream=None, dtype=None):
    """fp32 (N,C,T,H,W) -> channels-last (N,T,H,W,c_pitch) in the current precision's dtype (or `out`'s)."""
    N, C, T, H, W = x.shape
    dtype = out.dtype if out is not None else (dtype or cl_dtype())
    g = narrow_pitch(dtype)
    if c_pitch is None:
        c_pitch = out.shape[-1] if out is not None else (C + g - 1) // g * g
    if zero_to is None:
        zero_to = min(c_pitch, (c_off + C + g - 1) // g * g)
    if out is None:
        out = Tensor((N, T, H, W, c_pitch), dtype)
    fn = lib.hpvg_pack_cl_f32 if dtype == F32 else lib.hpvg_pack_cl
    check(fn(_p(x), N, C, T, H, W, _p(out), c_pitch, c_off, zero_to, _s(stream)), "pack_cl")
    return out


def unpack_cl(x_cl, C=None, c_off=0, out=None, stream=None):
    """channels-last (N,T,H,W,c_pitch) bf16 / fp32 -> fp32 (N,C,T,H,W)."""
    N, T, H, W, pitch = x_cl.shape
    if C is None:
        C = pitch - c_off
    if out is None:
        out = Tensor((N, C, T, H, W), F32)
    fn = lib.hpvg_unpack_cl_f32 if x_cl.dtype == F32 else lib.hpvg_unpack_cl
    check(fn(_p(x_cl), N, C, T, H, W, pitch, c_off, _p(out), _s(stream)), "unpack_cl")
    return out


# ------------------------------------------------------------------------------------------------ conv
def conv_mode_for(cin, cout):
    if cin <= 8 and cout == 64:
        return CONV_8_64
    if cin == 64 and cout == 64:
        return CONV_64_64
    if cin == 64 and cout <= 3:
        return CONV_64_3
    if cin == 64 and cout <= 4:
        return CONV_64_16
    raise HpvgError("no single-pass conv kernel variant for Cin=%d Cout=%d" % (cin, cout))


def tail_mode(cout):
    """Kernel variant for a 64 -> cout (<= 4) convolution with fp32 NCDHW output."""
    return CONV_64_3 if cout <= 3 else CONV_64_16


def is_tf32_mode(mode):
    return mode >= CONV_T32_64


def pack_weights(w, mode, transpose_flip=False, cout_off=0, cout=None, cin_off=0, cin=None, out=None, stream=None):
    """w: fp32 (Cout, Cin, [kt,] 3, 3) device tensor -> packed filter-bank image for `mode`."""
    if len(w.shape) == 5:
        w_cout, w_cin, kt = w.shape[0], w.shape[1], w.shape[2]
    else:
        w_cout, w_cin, kt = w.shape[0], w.shape[1], 1
    eff_cout, eff_cin = (w_cin, w_cout) if transpose_flip else (w_cout, w_cin)
    cout = eff_cout - cout_off if cout is None else cout
    cin = eff_cin - cin_off if cin is None else cin
    if out is None:
        nb = lib.hpvg_conv_wimg_bytes(mode)
        out = Tensor((nb // 4,), F32) if is_tf32_mode(mode) else Tensor((nb // 2,), BF16)
    check(lib.hpvg_conv_pack_weights(_p(w), w_cout, w_cin, kt, mode, 1 if transpose_flip else 0, cout_off, cout,
                                     cin_off, cin, _p(out), _s(stream)), "conv_pack_weights")
    return out


def plan_wimgs(cin, cout, dtype=None):
    """The filter banks of a cin -> cout convolution, in the order conv3d_cl_any consumes them, for activations of `dtype`
    (default: the current precision mode): a list of pack_weights keyword dicts (mode, cout_off, cout, cin_off, cin).
    cin / cout are the EFFECTIVE channel counts (for the data-gradient conv they are the forward layer's cout / cin)."""
    tf = (dtype or cl_dtype()) == F32
    if cout <= 4:
        if cin != 64:
            raise HpvgError("tail convolutions read 64 channels")
        if tf:
            if cout > 3:
                raise HpvgError("tf32 tail convolutions write at most 3 channels")
            return [dict(mode=CONV_T32_3, cout=cout, cin_off=32 * ib, cin=32) for ib in range(2)]
        return [dict(mode=tail_mode(cout), cout=cout)]
    plan = []
    if cin <= 8:
        if tf and cin > 4:
            raise HpvgError("tf32 head convolutions read at most 4 channels")
        mode = CONV_T4_64 if tf else CONV_8_64
        return [dict(mode=mode, cout_off=ob * 64, cout=64, cin=cin) for ob in range(cout // 64)]
    blk = 32 if tf else 64
    mode = CONV_T32_64 if tf else CONV_64_64
    for ob in range(cout // 64):
        for ib in range(cin // blk):
            plan.append(dict(mode=mode, cout_off=ob * 64, cout=64, cin_off=ib * blk, cin=blk))
    return plan


def build_wimgs(w, cin, cout, transpose_flip=False, dtype=None, stream=None):
    """Packed filter banks of a cin -> cout convolution (see plan_wimgs), one launch per bank."""
    return [pack_weights(w, transpose_flip=transpose_flip, stream=stream, **kw) for kw in plan_wimgs(cin, cout, dtype)]


def wimg_tensor(mode):
    nb = lib.hpvg_conv_wimg_bytes(mode)
    return Tensor((nb // 4,), F32) if is_tf32_mode(mode) else Tensor((nb // 2,), BF16)


def pack_weights_multi(entries, stream=None):
    """Many filter banks (bf16 kernel variants) and (1, bias) epilogue-vector pairs in ONE launch.
    entries: dicts with `out` plus either (w, mode, transpose_flip, cout_off, cout, cin_off, cin) or (bias, cout) —
    the latter writes fp32 [2][64] = (1, bias)."""
    n = len(entries)
    if not n:
        return
    vp = ctypes.c_void_p
    ws, outs = (vp * n)(), (vp * n)()
    cols = {k: (ctypes.c_int * n)() for k in ("w_cout", "w_cin", "kt", "mode", "flip", "cout_off", "cout", "cin_off", "cin")}
    for i, e in enumerate(entries):
        outs[i] = e["out"].ptr
        if "bias" in e:
            ws[i] = e["bias"].ptr if e["bias"] is not None else None
            cols["mode"][i], cols["cout"][i], cols["kt"][i] = -1, int(e["cout"]), 3
            continue
        w = e["w"]
        kt = w.shape[2] if len(w.shape) == 5 else 1
        w_cout, w_cin = w.shape[0], w.shape[1]
        flip = 1 if e.get("transpose_flip") else 0
        eff_cout, eff_cin = (w_cin, w_cout) if flip else (w_cout, w_cin)
        co_off, ci_off = int(e.get("cout_off", 0)), int(e.get("cin_off", 0))
        ws[i] = w.ptr
        cols["w_cout"][i], cols["w_cin"][i], cols["kt"][i], cols["mode"][i], cols["flip"][i] = w_cout, w_cin, kt, e["mode"], flip
        cols["cout_off"][i], cols["cin_off"][i] = co_off, ci_off
        cols["cout"][i] = int(e["cout"]) if e.get("cout") is not None else eff_cout - co_off
        cols["cin"][i] = int(e["cin"]) if e.get("cin") is not None else eff_cin - ci_off
    check(lib.hpvg_conv_pack_weights_multi(n, ws, cols["w_cout"], cols["w_cin"], cols["kt"], cols["mode"], cols["flip"],
                                           cols["cout_off"], cols["cout"], cols["cin_off"], cols["cin"], outs,
                                           _s(stream)), "conv_pack_weights_multi")


def conv_cl(mode, x_cl, wimg, scale, shift, act=ACT_NONE, out_mode=None, out=None, out_pitch=64, out_coff=0,
            cout_real=64, addend=None, in_coff=0, stats=None, mask=None, mask_coff=0, stream=None):
    """Raw kernel call.  x_cl: channels-last (N,T,H,W,in_pitch), bf16 for the kind::f16 variants, fp32 for the kind::tf32
    ones; reads channels [in_coff, in_coff + 64|8 (32|4)).
    stats: optional fp64 (2,64) tensor accumulating the per-channel sum / sum of squares of the stored output.
    mask (act=ACT_LRELU_MASK, bf16 only): cl tensor of a stored LeakyReLU activation; output = v * LeakyReLU'(mask)."""
    N, T, H, W, in_pitch = x_cl.shape
    tf = is_tf32_mode(mode)
    if tf != (x_cl.dtype == F32):
        raise HpvgError("conv_cl: kernel variant %d does not read %s activations" % (mode, x_cl.dtype))
    if out_mode is None:
        out_mode = OUT_F32_CL if tf else OUT_BF16_CL
    if out is None:
        if out_mode == OUT_BF16_CL:
            out = Tensor((N, T, H, W, out_pitch), BF16)
        elif out_mode == OUT_F32_CL:
            out = Tensor((N, T, H, W, out_pitch), F32)
        elif out_mode == OUT_F32_RAW:
            out = Tensor((N, T, H, W, 64), F32)
        else:
            out = Tensor((N, cout_real, T, H, W), F32)
    in_ptr = ctypes.c_void_p(x_cl.ptr + _isz(x_cl) * in_coff)
    timed = _prof is not None and (_prof["mode"] == mode or _prof["mode"] is None or _prof["mode"] == "all")
    if timed:
        e0, e1 = rt.Event(), rt.Event()
        e0.record(stream)
    check(lib.hpvg_conv_cl(mode, N, T, H, W, in_ptr, in_pitch, _p(wimg), _p(scale), _p(shift), act, out_mode, _p(out),
                           out_pitch, out_coff, cout_real, _p(addend), _p(stats),
                           None if mask is None else ctypes.c_void_p(mask.ptr + 2 * mask_coff),
                           0 if mask is None else mask.shape[-1], _s(stream)), "conv_cl")
    if timed:
        e1.record(stream)
        if _prof["mode"] == "all":
            _prof["items"].append((("conv", mode, N * T * H * W), e0, e1))
        else:
            _prof["items"].append((N * T * H * W, e0, e1) if _prof["mode"] is not None
                                  else ((mode, N * T * H * W), e0, e1))
    return out


def affine_from_bias(bias, inv_sigma=None, C=None, out=None, stream=None):
    """(scale, shift) = (1 or 1/sigma, bias) epilogue vectors; returned as one (2, 64) tensor."""
    C = bias.shape[0] if C is None else C
    Cp = max(64, (C + 63) // 64 * 64)
    if out is None:
        out = Tensor((2, Cp), F32).zero_(stream)
    check(lib.hpvg_affine_from_bias(_p(bias), _p(inv_sigma), C, _p(out), ctypes.c_void_p(out.ptr + 4 * Cp),
                                    _s(stream)), "affine_from_bias")
    return out


def bn_fold_eval(gamma, beta, mean, var, bias, out=None, eps=BN_EPS, stream=None):
    """Eval-mode BatchNorm folded into the conv epilogue: scale = g/sqrt(v+eps), shift = (b-mean)*scale+beta."""
    C = gamma.shape[0]
    Cp = max(64, (C + 63) // 64 * 64)
    if out is None:
        out = Tensor((2, Cp), F32).zero_(stream)
    check(lib.hpvg_bn_fold_eval(_p(gamma), _p(beta), _p(mean), _p(var), eps, _p(bias), C, _p(out),
                                ctypes.c_void_p(out.ptr + 4 * Cp), _s(stream)), "bn_fold_eval")
    return out


def _sc(aff):
    return aff, aff.view((64,), F32, 256)


_SCRATCH = {}


def _zeros64():
    t = _SCRATCH.get("zeros64")
    if t is None:
        t = Tensor((64,), F32).zero_()
        rt.device_sync()
        _SCRATCH["zeros64"] = t
    return t


def _partial_buf(shape, stream):
    """fp32 raw partial sums of a split-Cin layer: one buffer per (stream, shape) — launches on a stream are ordered, so
    consecutive layers can share it; concurrent streams get their own."""
    key = ("partial", None if stream is None else stream.handle.value, tuple(shape))
    t = _SCRATCH.get(key)
    if t is None:
        t = Tensor(shape, F32)
        _SCRATCH[key] = t
    return t


def conv3d_cl_any(x_cl, w, aff, act, cin, cout, out_mode=None, residual=None, out=None, wimgs=None,
                  transpose_flip=False, stats=None, mask=None, stream=None, scale_shift=None):
    """Convolution for every channel combination on the hot path, composed from the kernel variants:
    Cin in {<=8, 64, 128}, Cout in {<=4, 64, 128}.  aff: (2,64*ceil(cout/64)) epilogue vectors (or scale_shift = one
    (scale, shift) pair of 64-vectors used for every output block).  The precision follows x_cl's dtype: bf16 ->
    kind::f16 kernels, fp32 -> kind::tf32 kernels (a 64-channel input block is consumed in two 32-channel launches).
    Returns channels-last (cout 64/128, same dtype as x_cl) or fp32 ncdhw (cout <= 4)."""
    N, T, H, W, pitch = x_cl.shape
    tf = x_cl.dtype == F32
    if _flops is not None and w is not None:
        taps = 27 if (len(w.shape) == 5 and w.shape[2] == 3) else 9
        _flops["conv"] += 2 * taps * w.shape[0] * w.shape[1] * N * T * H * W
        _flops["conv_calls"] += 1
    if wimgs is None:
        wimgs = build_wimgs(w, cin, cout, transpose_flip, dtype=x_cl.dtype, stream=stream)
    if mask is not None and tf:
        raise HpvgError("the fused LeakyReLU-mask epilogue exists for the bf16 kernels only")
    if cout <= 4:
        assert cin == 64
        s, b = scale_shift if scale_shift is not None else _sc(aff)
        if not tf:
            return conv_cl(tail_mode(cout), x_cl, wimgs[0], s, b, act, OUT_F32_NCDHW, out=out, cout_real=cout,
                           addend=residual, stream=stream)
        # two 32-channel launches: (conv_lo*scale + residual), then act(conv_hi*scale + shift + that)
        out = conv_cl(CONV_T32_3, x_cl, wimgs[0], s, _zeros64(), ACT_NONE, OUT_F32_NCDHW, out=out, cout_real=cout,
                      addend=residual, stream=stream)
        return conv_cl(CONV_T32_3, x_cl, wimgs[1], s, b, act, OUT_F32_NCDHW, out=out, cout_real=cout, addend=out,
                       in_coff=32, stream=stream)
    n_ob = cout // 64
    blk = 32 if tf else 64
    n_ib = 1 if cin <= 8 else cin // blk
    cl_out = OUT_F32_CL if tf else OUT_BF16_CL
    if out is None:
        out = Tensor((N, T, H, W, cout), x_cl.dtype)
    k = 0
    for ob in range(n_ob):
        if scale_shift is not None:
            s, b = scale_shift
        else:
            s = aff.view((64,), F32, ob * 256)
            b = aff.view((64,), F32, (n_ob + ob) * 256)
        partial = None
        for ib in range(n_ib):
            if cin <= 8:
                mode = CONV_T4_64 if tf else CONV_8_64
            else:
                mode = CONV_T32_64 if tf else CONV_64_64
            wi = wimgs[k]
            k += 1
            last = ib == n_ib - 1
            if not last:
                buf = _partial_buf((N, T, H, W, 64), stream) if tf else None
                partial = conv_cl(mode, x_cl, wi, s, b, ACT_NONE, OUT_F32_RAW, out=buf, addend=partial,
                                  in_coff=ib * blk, stream=stream)
            else:
                conv_cl(mode, x_cl, wi, s, b, act, cl_out, out=out, out_pitch=cout, out_coff=ob * 64,
                        addend=partial, in_coff=ib * blk, stats=stats, mask=mask, mask_coff=ob * 64, stream=stream)
    return out


# ------------------------------------------------------------------------------------------------ resize
def linear_taps(n_in, n_out, align_corners=True):
    """Host tables (i0, i1, l0, l1) of the resize rule — bit-exact contract with the reference."""
    i0 = np.empty(n_out, np.int32)
    i1 = np.empty(n_out, np.int32)
    l0 = np.empty(n_out, np.float32)
    l1 = np.empty(n_out, np.float32)
    c = ctypes
    check(lib.hpvg_linear_taps(n_in, n_out, int(align_corners), i0.ctypes.data_as(c.POINTER(c.c_int32)),
                               i1.ctypes.data_as(c.POINTER(c.c_int32)), l0.ctypes.data_as(c.POINTER(c.c_float)),
                               l1.ctypes.data_as(c.POINTER(c.c_float))), "linear_taps")
    return i0, i1, l0, l1


def linear_taps_device(n_in, n_out, align_corners=True):
    """The same tables as computed by the device code path."""
    i0, i1 = Tensor((n_out,), I32), Tensor((n_out,), I32)
    l0, l1 = Tensor((n_out,), F32), Tensor((n_out,), F32)
    check(lib.hpvg_linear_taps_dev(n_in, n_out, int(align_corners), _p(i0), _p(i1), _p(l0), _p(l1), None),
          "linear_taps_dev")
    return i0.numpy(), i1.numpy(), l0.numpy(), l1.numpy()


def resize3d(x, size, align_corners=True, out=None, stream=None):
    """UpsampleTrilinear3D(output_size=size, align_corners) forward (trilinear.py:171-254)."""
    N, C, Ti, Hi, Wi = x.shape
    To, Ho, Wo = (int(v) for v in size)
    if out is None:
        out = Tensor((N, C, To, Ho, Wo), F32)
    check(lib.hpvg_resize3d_fwd(_p(x), N, C, Ti, Hi, Wi, _p(out), To, Ho, Wo, int(align_corners), _s(stream)),
          "resize3d_fwd")
    return out


def resize3d_bwd(gy, in_size, align_corners=True, out=None, stream=None):
    N, C, To, Ho, Wo = gy.shape
    Ti, Hi, Wi = (int(v) for v in in_size)
    if out is None:
        out = Tensor((N, C, Ti, Hi, Wi), F32)
    check(lib.hpvg_resize3d_bwd(_p(gy), N, C, To, Ho, Wo, _p(out), Ti, Hi, Wi, int(align_corners), _s(stream)),
          "resize3d_bwd")
    return out


def frames_to_clip(frames, size, start=0, every=1, n_frames=None, hflip=False, bgr=False, out=None, stream=None):
    """Decoded uint8 frames (F, Hs, Ws, 3) on the device -> the fp32 clip (1, 3, T, H, W) in [-1, 1] that the
    reference's SingleVideoDataset yields (generate_frames.py:42-46 cv2.resize INTER_LINEAR, video.py:52-59 frame
    window + /255, video.py:75-86 flip + Normalize + CTHW)."""
    F_, Hs, Ws, C = frames.shape
    if C != 3 or frames.dtype != "uint8":
        raise HpvgError("frames_to_clip: frames must be uint8 (F, H, W, 3)")
    H, W = (int(v) for v in size)
    T = int(n_frames) if n_frames is not None else (F_ - int(start) + int(every) - 1) // int(every)
    if out is None:
        out = Tensor((1, 3, T, H, W), F32)
    check(lib.hpvg_frames_to_clip(_p(frames), F_, Hs, Ws, int(bool(bgr)), int(start), int(every), T, H, W,
                                  int(bool(hflip)), _p(out), _s(stream)), "frames_to_clip")
    return out


def randn(shape, seed, offset=0, d_offset=None, out=None, stream=None):
    """N(0,1) drawn on the device, keyed by (seed, offset [+ device counter], element)."""
    out = out or Tensor(shape, F32)
    check(lib.hpvg_randn(_p(out), out.size, int(seed) & 0xFFFFFFFFFFFFFFFF, int(offset), _p(d_offset), _s(stream)),
          "randn")
    return out


def box_muller_(z, stream=None):
    """In place: host-drawn uniforms in [0, 1) -> N(0, 1) (pairs of consecutive elements, Box-Muller)."""
    check(lib.hpvg_box_muller_inplace(_p(z), z.size, _s(stream)), "box_muller_inplace")
    return z


def counter_add(counter, inc=1, stream=None):
    check(lib.hpvg_counter_add(_p(counter), int(inc), _s(stream)), "counter_add")
    return counter


def upsample_noise_pack(x, size, noise=None, amp=0.0, seed=0, sample_base=0, up=None, xin=None, stream=None,
                        d_sample_offset=None):
    """Block input stage (networks_3d.py:440-446).  Returns (up fp32 ncdhw, x_in narrow channels-last: 8 bf16 channels
    or, in tf32 mode, 4 fp32 channels per voxel)."""
    N, C, Ti, Hi, Wi = x.shape
    To, Ho, Wo = (int(v) for v in size)
    if up is None:
        up = Tensor((N, C, To, Ho, Wo), F32)
    if xin is None:
        xin = Tensor((N, To, Ho, Wo, narrow_pitch()), cl_dtype())
    fn = lib.hpvg_upsample_noise_pack_f32 if xin.dtype == F32 else lib.hpvg_upsample_noise_pack
    check(fn(_p(x), N, C, Ti, Hi, Wi, To, Ho, Wo, _p(noise), float(amp), int(seed),
                                       int(sample_base), _p(d_sample_offset), _p(up), _p(xin), _s(stream)),
          "upsample_noise_pack")
    return up, xin


# ------------------------------------------------------------------------------------------------ batch norm (train)
def bn_train_cl(y_cl, gamma, beta, moving_mean, moving_var, act=ACT_LRELU, out=None, stats=None, stream=None):
    """Training-mode BatchNorm + LeakyReLU over a (.., 64) cl tensor (networks_3d.py:52).  Updates the moving
    statistics in place.  Returns (x_cl, saved) where saved = (scale, shift, mean, invstd) for the backward."""
    voxels = int(np.prod(y_cl.shape[:-1]))
    f32 = y_cl.dtype == F32
    if stats is None:
        stats = Tensor((2, 64), F64)
    sums, sumsq = stats, stats.view((64,), F64, 512)
    check((lib.hpvg_bn_stats_cl_f32 if f32 else lib.hpvg_bn_stats_cl)(_p(y_cl), voxels, _p(sums), _p(sumsq),
                                                                       _s(stream)), "bn_stats")
    saved = Tensor((4, 64), F32)
    sc, sh = saved, saved.view((64,), F32, 256)
    mean, invstd = saved.view((64,), F32, 512), saved.view((64,), F32, 768)
    check(lib.hpvg_bn_finalize(_p(sums), _p(sumsq), voxels, _p(gamma), _p(beta), BN_EPS, BN_MOMENTUM,
                               _p(moving_mean), _p(moving_var), _p(sc), _p(sh), _p(mean), _p(invstd), _s(stream)),
          "bn_finalize")
    if out is None:
        out = Tensor(y_cl.shape, y_cl.dtype)
    check((lib.hpvg_bn_apply_lrelu_cl_f32 if f32 else lib.hpvg_bn_apply_lrelu_cl)(
        _p(y_cl), voxels, _p(sc), _p(sh), act, _p(out), _s(stream)), "bn_apply")
    return out, saved


def bn_train_fused_cl(y_cl, stats, gamma, beta, moving_mean, moving_var, act=ACT_LRELU, out=None, saved=None,
                      stream=None, center=None):
    """One-pass training-mode BatchNorm + activation whose batch statistics were accumulated by the producing conv's
    epilogue (conv_cl(stats=...)).  Updates the moving statistics; `saved` (4,64) receives (scale, shift, mean, invstd).
    center: the per-channel offset the conv subtracted from y (kink-centred storage, bn_center_multi) — `saved` stays
    in that centred frame, the moving mean gets it added back."""
    voxels = int(np.prod(y_cl.shape[:-1]))
    if out is None:
        out = Tensor(y_cl.shape, y_cl.dtype)
    fn = lib.hpvg_bn_train_apply_cl_f32 if y_cl.dtype == F32 else lib.hpvg_bn_train_apply_cl
    check(fn(_p(y_cl), voxels, _p(stats), _p(gamma), _p(beta), BN_EPS, BN_MOMENTUM,
             _p(moving_mean), _p(moving_var), _p(saved), act, _p(out), _p(center), _s(stream)),
          "bn_train_apply")
    return out


def bn_center_multi(layers, stream=None):
    """layers: [(gamma, beta, moving_mean, moving_var, bias, center_out (64,), aff_out (2,64))] -> one launch computing
    every layer's kink estimate and its conv epilogue vectors (1, bias - center)."""
    n = len(layers)
    if n == 0:
        return
    VP = ctypes.c_void_p * n
    cols = [VP(*[t[i].ptr for t in layers]) for i in range(7)]
    check(lib.hpvg_bn_center_multi(n, *cols, BN_EPS, _s(stream)), "bn_center_multi")


def bn_moving_update_multi(items, stream=None):
    """items: [(saved (4,64), moving_mean, moving_var[, center or None])] in the order the forwards ran:
    moving = 0.9*moving + 0.1*batch (the batch mean = saved mean + center when the layer's y was stored centred)."""
    n = len(items)
    if n == 0:
        return
    VP = ctypes.c_void_p * n
    cen = [it[3] if len(it) > 3 else None for it in items]
    check(lib.hpvg_bn_moving_update_multi(n, VP(*[it[0].ptr for it in items]), VP(*[it[1].ptr for it in items]),
                                          VP(*[it[2].ptr for it in items]),
                                          VP(*[None if c is None else c.ptr for c in cen]), BN_EPS, BN_MOMENTUM,
                                          _s(stream)), "bn_moving_update_multi")


# ------------------------------------------------------------------------------------------------ spectral norm
def sn_power_iter_multi(layers, stream=None):
    """One launch for all spectrally normalised layers of a network.  layers: list of dicts with device tensors
    w, u, v, sigma (2,), bias, aff (2,64) [, u_copy, v_copy]: updates u, v in place, writes (sigma, 1/sigma), the conv
    epilogue vectors and (optionally) a snapshot of the updated u, v."""
    n = len(layers)
    VP = ctypes.c_void_p * n
    IA = ctypes.c_int * n
    couts = [l["w"].shape[0] for l in layers]
    ks = [l["w"].size // c for l, c in zip(layers, couts)]
    check(lib.hpvg_sn_power_iter_multi(n, VP(*[l["w"].ptr for l in layers]), IA(*couts), IA(*ks),
                                       VP(*[l["u"].ptr for l in layers]), VP(*[l["v"].ptr for l in layers]),
                                       VP(*[l["sigma"].ptr for l in layers]), VP(*[l["bias"].ptr for l in layers]),
                                       VP(*[l["aff"].ptr for l in layers]),
                                       VP(*[(l["u_copy"].ptr if l.get("u_copy") is not None else None) for l in layers]),
                                       VP(*[(l["v_copy"].ptr if l.get("v_copy") is not None else None) for l in layers]),
                                       _s(stream)), "sn_power_iter_multi")


def sn_power_iter(w, u, v, out=None, stream=None):
    """spectral_norm.py:142-151.  Updates u, v in place; returns a (2,) tensor (sigma, 1/sigma)."""
    cout = w.shape[0]
    k = w.size // cout
    if out is None:
        out = Tensor((2,), F32)
    check(lib.hpvg_sn_power_iter(_p(w), cout, k, _p(u), _p(v), _p(out), ctypes.c_void_p(out.ptr + 4), _s(stream)),
          "sn_power_iter")
    return out


# ------------------------------------------------------------------------------------------------ losses / misc
def mse(a, b, out=None, stream=None):
    out = out or Tensor((1,), F32)
    check(lib.hpvg_mse(_p(a), _p(b), a.size, _p(out), _s(stream)), "mse")
    return out


def mean(a, out=None, stream=None):
    out = out or Tensor((1,), F32)
    check(lib.hpvg_mean(_p(a), a.size, _p(out), _s(stream)), "mean")
    return out


def kl_criterion(mu, logvar, out=None, stream=None):
    """losses.py:5-7."""
    out = out or Tensor((1,), F32)
    check(lib.hpvg_kl(_p(mu), _p(logvar), mu.size, _p(out), _s(stream)), "kl")
    return out


def reparam(mu, logvar, eps, out=None, stream=None):
    """networks_3d.py:415-417."""
    out = out or Tensor(mu.shape, F32)
    check(lib.hpvg_reparam(_p(mu), _p(logvar), _p(eps), mu.size, _p(out), _s(stream)), "reparam")
    return out


def reparam_bwd(gz, eps, logvar, gmu, glv, stream=None):
    """Backward of networks_3d.py:415-417: gmu += gz ; glogvar += gz*eps*0.5*exp(0.5*logvar)."""
    check(lib.hpvg_reparam_bwd(_p(gz), _p(eps), _p(logvar), gz.size, _p(gmu), _p(glv), _s(stream)), "reparam_bwd")
    return gmu, glv


def adam_clip_multi(params, grads, ms, vs, lrs, step, beta1=0.5, beta2=0.999, eps=1e-8, clip=0.0, stream=None,
                    d_step=None):
    """ClippedAdam.construct (optimizers.py:41-43) / nn.Adam for D (train_video.py:65): per-tensor ClipByNorm then Adam."""
    n = len(params)
    VP = ctypes.c_void_p * n
    sizes = (ctypes.c_longlong * n)(*[p.size for p in params])
    lr_arr = (ctypes.c_float * n)(*[float(x) for x in lrs])
    check(lib.hpvg_adam_clip_multi(n, VP(*[p.ptr for p in params]), VP(*[g.ptr for g in grads]),
                                   VP(*[m.ptr for m in ms]), VP(*[v.ptr for v in vs]), sizes, lr_arr, beta1, beta2,
                                   eps, int(step), float(clip), _p(d_step), _s(stream)), "adam_clip_multi")


# ================================================================================================ backward operators
def conv_wgrad_cl(x_cl, gy_cl, dw, co_off=0, co_n=64, ci_off=0, ci_n=64, x_coff=0, gy_coff=0, accumulate=False,
                  scale=1.0, stream=None):
    """dW block (+)= scale * sum_v gy[v] (x) x[v+tap].  x_cl / gy_cl: cl tensors of the same dtype (bf16 -> kind::f16
    kernel, fp32 -> kind::tf32 kernel), either >= 64 channels (a 64-channel slice at x_coff / gy_coff) or a narrow
    tensor (all of its channels, the rest read as zero); dw: fp32 (Cout, Cin, [kt,] 3, 3)."""
    N, T, H, W, xp = x_cl.shape
    gp = gy_cl.shape[-1]
    if x_cl.dtype != gy_cl.dtype:
        raise HpvgError("conv_wgrad_cl: operands must have the same dtype")
    if (xp < 64 and x_coff) or (gp < 64 and gy_coff):
        raise HpvgError("conv_wgrad_cl: a narrow operand cannot be sliced")
    kt = dw.shape[2] if len(dw.shape) == 5 else 1
    isz = _isz(x_cl)
    fn = lib.hpvg_conv_wgrad_cl_tf32 if x_cl.dtype == F32 else lib.hpvg_conv_wgrad_cl
    if _flops is not None:
        _flops["wgrad"] += 2 * (27 if kt == 3 else 9) * min(co_n, dw.shape[0]) * min(ci_n, dw.shape[1]) * N * T * H * W
        _flops["wgrad_calls"] += 1
    timed = _prof is not None and _prof["mode"] == "all"
    if timed:
        e0, e1 = rt.Event(), rt.Event()
        e0.record(stream)
    check(fn(ctypes.c_void_p(x_cl.ptr + isz * x_coff), xp, ctypes.c_void_p(gy_cl.ptr + isz * gy_coff),
             gp, N, T, H, W, _p(dw), dw.shape[1], kt, co_off, co_n, ci_off, ci_n,
             1 if accumulate else 0, float(scale), _s(stream)), "conv_wgrad_cl")
    if timed:
        e1.record(stream)
        _prof["items"].append((("wgrad", None, N * T * H * W), e0, e1))
    return dw


def slice_act_cl(x_cl, h0=0, w0=0, sh=1, sw=1, out_hw=None, relu=False, out=None, stream=None):
    """Strided window of a bf16 channels-last tensor (N,T,H,W,C), optionally with ReLU:
    out[..., ho, wo, :] = act(x[..., h0 + ho*sh, w0 + wo*sw, :])."""
    N, T, H, W, C = x_cl.shape
    if x_cl.dtype != BF16:
        raise HpvgError("slice_act_cl works on bf16 channels-last tensors")
    Ho, Wo = out_hw if out_hw is not None else ((H - h0 + sh - 1) // sh, (W - w0 + sw - 1) // sw)
    if out is None:
        out = Tensor((N, T, Ho, Wo, C), BF16)
    check(lib.hpvg_slice_act_cl(_p(x_cl), N * T, H, W, C, Ho, Wo, h0, w0, sh, sw, 1 if relu else 0, _p(out), _s(stream)),
          "slice_act_cl")
    return out


def reflect_pad_cl(x_cl, pad_t=1, pad_hw=1, out=None, stream=None):
    """nn.Pad(mode='REFLECT') of a channels-last tensor (N,T,H,W,C) (reference networks_3d.py:65-68)."""
    N, T, H, W, C = x_cl.shape
    if out is None:
        out = Tensor((N, T + 2 * pad_t, H + 2 * pad_hw, W + 2 * pad_hw, C), x_cl.dtype)
    check(lib.hpvg_reflect_pad_cl(_p(x_cl), N, T, H, W, C * _isz(x_cl), int(pad_t), int(pad_hw), _p(out), _s(stream)),
          "reflect_pad_cl")
    return out


def lrelu_bwd_cl(ga, a, out=None, stream=None):
    out = out or Tensor(ga.shape, ga.dtype)
    fn = lib.hpvg_lrelu_bwd_cl_f32 if ga.dtype == F32 else lib.hpvg_lrelu_bwd_cl
    check(fn(_p(ga), _p(a), ga.size, _p(out), _s(stream)), "lrelu_bwd_cl")
    return out


def bn_bwd_cl(ga, y, saved, act=ACT_LRELU, out=None, dgamma=None, dbeta=None, accumulate=False, stream=None):
    voxels = int(np.prod(ga.shape[:-1]))
    out = out or Tensor(ga.shape, ga.dtype)
    fn = lib.hpvg_bn_bwd_cl_f32 if ga.dtype == F32 else lib.hpvg_bn_bwd_cl
    check(fn(_p(ga), _p(y), voxels, _p(saved), act, _p(out), _p(dgamma), _p(dbeta),
                             1 if accumulate else 0, _s(stream)), "bn_bwd_cl")
    return out


def colsum_cl(g, out, accumulate=False, stream=None):
    fn = lib.hpvg_colsum_cl_f32 if g.dtype == F32 else lib.hpvg_colsum_cl
    check(fn(_p(g), int(np.prod(g.shape[:-1])), _p(out), 1 if accumulate else 0, _s(stream)), "colsum_cl")
    return out


def mse_grad(out_t, target, coef, g=None, accumulate=False, stream=None):
    g = g or Tensor(out_t.shape, F32)
    check(lib.hpvg_mse_grad(_p(out_t), _p(target), out_t.size, float(coef), 1 if accumulate else 0, _p(g), _s(stream)),
          "mse_grad")
    return g


def tanh_bwd(g, out_t, gpre=None, stream=None):
    gpre = gpre or Tensor(g.shape, F32)
    check(lib.hpvg_tanh_bwd(_p(g), _p(out_t), g.size, _p(gpre), _s(stream)), "tanh_bwd")
    return gpre


def axpby(a, x, b, y, stream=None):
    check(lib.hpvg_axpby(float(a), _p(x), float(b), _p(y), x.size, _s(stream)), "axpby")
    return y


def fill(t, value, stream=None):
    check(lib.hpvg_fill(_p(t), float(value), t.size, _s(stream)), "fill")
    return t


def gather_tap(dw, tap, out, stream=None):
    """out[co][ci] = dw[co][ci][tap] for a (Cout, Cin, kt, 3, 3) tensor."""
    n = dw.shape[0] * dw.shape[1]
    check(lib.hpvg_gather_strided(_p(dw), n, dw.size // n, int(tap), _p(out), _s(stream)), "gather_strided")
    return out


def channel_sum(g, out, accumulate=False, stream=None):
    N, C = g.shape[0], g.shape[1]
    check(lib.hpvg_channel_sum(_p(g), N, C, g.size // (N * C), 1 if accumulate else 0, _p(out), _s(stream)),
          "channel_sum")
    return out


def kl_grad(mu, logvar, coef, gmu=None, glv=None, stream=None):
    gmu = gmu or Tensor(mu.shape, F32)
    glv = glv or Tensor(mu.shape, F32)
    check(lib.hpvg_kl_grad(_p(mu), _p(logvar), mu.size, float(coef), _p(gmu), _p(glv), _s(stream)), "kl_grad")
    return gmu, glv


def sn_grad(G, w, u, v, sigma, gw, accumulate=False, stream=None):
    cout = w.shape[0]
    check(lib.hpvg_sn_grad(_p(G), _p(w), _p(u), _p(v), _p(sigma), cout, w.size // cout, 1 if accumulate else 0, _p(gw),
                           _s(stream)), "sn_grad")
    return gw


def lerp(a, b, alpha, out=None, stream=None):
    out = out or Tensor(a.shape, F32)
    check(lib.hpvg_lerp(_p(a), _p(b), float(alpha), a.size, _p(out), _s(stream)), "lerp")
    return out


def gp_grad(g, lam, Gout=None, gp=None, stream=None):
    """losses.py:47-52: returns (G = d GP / d g, gp scalar tensor)."""
    N, C = g.shape[0], g.shape[1]
    Gout = Gout or Tensor(g.shape, F32)
    gp = gp or Tensor((1,), F32)
    check(lib.hpvg_gp_grad(_p(g), N, C, g.size // (N * C), float(lam), _p(Gout), _p(gp), _s(stream)), "gp_grad")
    return Gout, gp
