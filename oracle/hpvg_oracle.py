"""CPU oracle for the HP-VAE-GAN hot path (TEST INFRASTRUCTURE ONLY — never imported by the product package).

A restatement, in numpy + torch-CPU fp32, of the arithmetic the reference `SakiRinn/mindspore-hp-vae-gan` performs
on the generator / discriminator hot path.  Every function cites the reference file:line it follows.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this module.

PARITY PINNING:
  * pinned by the reference tree itself: trilinear known-answer (src/tools/trilinear.py:222-233), pyramid geometry
    value (src/utils/images.py:126), smoke-block output shapes (src/modules/networks_3d.py:554-593).
  * everything that lives inside the un-vendored `mindspore` package (conv, BatchNorm, LeakyReLU slope, Adam,
    ClipByNorm, L2Normalize, autodiff) is restated from MindSpore's published semantics and is "parity unpinned":
    MindSpore cannot be installed offline in this environment.  torch-CPU (oneDNN) is used here as the numerical
    engine for convolutions and autodiff; it is an independent implementation from the CUDA kernels under test.

MindSpore semantics assumed (marked ‡ in SURVEY.md): LeakyReLU alpha 0.2; BatchNorm eps 1e-5, moving-stat update
`moving = 0.9*moving + 0.1*batch` with the biased batch variance; Normal(sigma, mean) argument order;
L2Normalize eps 1e-12 (x / sqrt(max(sum x^2, eps))); Adam with eps outside the bias correction;
ClipByNorm x*c/max(||x||, c).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.2   # mindspore.nn.LeakyReLU default alpha (networks_3d.py:20)
BN_EPS = 1e-5       # mindspore.nn.BatchNorm3d default eps (networks_3d.py:52)
BN_MOMENTUM = 0.9   # mindspore.nn.BatchNorm3d default momentum: moving = m*moving + (1-m)*batch


# ----------------------------------------------------------------------------------------------------------------
# Options and pyramid geometry                                                      (src/utils/images.py:64-93)
# ----------------------------------------------------------------------------------------------------------------
def default_opt(**kw):
    """Defaults of train_video.py:232-294 plus the dataset-derived fields (src/datasets/video.py:28-35)."""
    o = SimpleNamespace(
        nc_im=3, nfc=64, latent_dim=128, vae_levels=3, enc_blocks=2, ker_size=3, num_layer=5, padd_size=1,
        scale_factor=0.75, noise_amp=0.1, min_size=32, max_size=256, img_size=256, sampling_rates=[4, 3, 2, 1],
        stop_scale_time=-1, lr_g=5e-4, lr_d=5e-4, beta1=0.5, lambda_grad=0.1, rec_weight=10.0, kl_weight=1.0,
        disc_loss_weight=1.0, lr_scale=0.2, train_depth=1, grad_clip=5.0, train_all=False, batch_size=1,
        org_fps=24.0, ar=0.75, const_amp=False)
    for k, v in kw.items():
        setattr(o, k, v)
    o.noise_amp_init = o.noise_amp
    o.scale_factor_init = o.scale_factor
    adjust_scales2image(o.img_size, o)
    if o.stop_scale_time == -1:
        o.stop_scale_time = o.stop_scale
    o.fps_lcm = int(np.lcm.reduce(o.sampling_rates))
    return o


def adjust_scales2image(size, opt):
    """images.py:64-71 (verbatim arithmetic)."""
    opt.num_scales = math.ceil((math.log(math.pow(opt.min_size / size, 1), opt.scale_factor_init))) + 1
    scale2stop = math.ceil(math.log(min([opt.max_size, size]) / size, opt.scale_factor_init))
    opt.stop_scale = opt.num_scales - scale2stop
    opt.scale1 = min(opt.max_size / size, 1)
    opt.scale_factor = math.pow(opt.min_size / size, 1 / opt.stop_scale)
    scale2stop = math.ceil(math.log(min([opt.max_size, size]) / size, opt.scale_factor_init))
    opt.stop_scale = opt.num_scales - scale2stop


def get_scales_by_index(index, scale_factor, stop_scale, img_size):
    """images.py:74-77."""
    scale = math.pow(scale_factor, stop_scale - index) + 1e-6
    return math.ceil(scale * img_size)


def get_fps_by_index(index, stop_scale_time, sampling_rates, org_fps):
    """images.py:80-84."""
    fps_index = int((index / stop_scale_time) * (len(sampling_rates) - 1))
    return org_fps / sampling_rates[fps_index], fps_index


def get_fps_td_by_index(index, stop_scale_time, sampling_rates, org_fps, fps_lcm):
    """images.py:87-93."""
    fps, fps_index = get_fps_by_index(index, stop_scale_time, sampling_rates, org_fps)
    every = sampling_rates[fps_index]
    return fps, fps_lcm // every + 1, fps_index


def scale_shape(opt, index):
    """(T, H, W) of pyramid level `index`: images.py:96-103 (3D) — [td, int(s*ar), s]."""
    s = get_scales_by_index(index, opt.scale_factor, opt.stop_scale, opt.img_size)
    _, td, _ = get_fps_td_by_index(index, opt.stop_scale_time, opt.sampling_rates, opt.org_fps, opt.fps_lcm)
    # synthetic benchmark clips may pin the time depth of a level (BASELINE.json config 3: 16 frames at the finest scale)
    td = getattr(opt, "td_override", {}).get(index, td)
    return (td, int(s * opt.ar), s)


def scale_shape_2d(opt, index):
    """images.py:110-114 (2D) — [int(s*ar), s]."""
    s = get_scales_by_index(index, opt.scale_factor, opt.stop_scale, opt.img_size)
    return (int(s * opt.ar), s)


# ----------------------------------------------------------------------------------------------------------------
# Data path between the decoder and the network            (src/datasets/generate_frames.py:42-46, video.py:45-86)
# ----------------------------------------------------------------------------------------------------------------
def _cv_sat_short(v):
    return max(-32768, min(32767, int(np.rint(np.float32(v) * np.float32(2048)))))   # saturate_cast<short>(f * 2048)


def _cv_x_coeffs(n_in, n_out):
    """OpenCV resize.cpp, INTER_LINEAR: half-pixel centres; at the borders the index is clamped AND the fraction
    zeroed (x axis only)."""
    scale = n_in / n_out
    i0 = np.empty(n_out, np.int64)
    a = np.empty((n_out, 2), np.int64)
    for d in range(n_out):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = np.float32(fx - np.float32(sx))
        if sx < 0:
            fx, sx = np.float32(0), 0
        if sx >= n_in - 1:
            fx, sx = np.float32(0), n_in - 1
        i0[d] = sx
        a[d] = _cv_sat_short(np.float32(1.0) - fx), _cv_sat_short(fx)
    return i0, a


def _cv_y_coeffs(n_in, n_out):
    """Rows are clamped into the image, the fraction is NOT touched (so a clamped border row is blended with itself)."""
    scale = n_in / n_out
    r0 = np.empty(n_out, np.int64)
    r1 = np.empty(n_out, np.int64)
    b = np.empty((n_out, 2), np.int64)
    for d in range(n_out):
        fy = np.float32((d + 0.5) * scale - 0.5)
        sy = int(np.floor(fy))
        fy = np.float32(fy - np.float32(sy))
        r0[d] = min(max(sy, 0), n_in - 1)
        r1[d] = min(max(sy + 1, 0), n_in - 1)
        b[d] = _cv_sat_short(np.float32(1.0) - fy), _cv_sat_short(fy)
    return r0, r1, b


def cv_resize_linear_u8(src, size):
    """cv2.resize(src, (W, H), interpolation=cv2.INTER_LINEAR) for uint8 HxWxC, restated: 11-bit fixed-point
    coefficients, int horizontal pass, vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2; an exact 2x
    decimation is the 2x2 box mean (cv::resize switches INTER_LINEAR to INTER_AREA there).  Pinned against cv2
    itself in tests/test_cpu_oracle.py (generate_frames.py:44-46)."""
    src = np.asarray(src, dtype=np.uint8)
    H, W = int(size[0]), int(size[1])
    Hs, Ws = src.shape[:2]
    if (Hs, Ws) == (H, W):
        return src.copy()
    s = src.astype(np.int64)
    if Hs == 2 * H and Ws == 2 * W:
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    xi, xa = _cv_x_coeffs(Ws, W)
    r0, r1, yb = _cv_y_coeffs(Hs, H)
    x1 = np.minimum(xi + 1, Ws - 1)
    hr = s[:, xi] * xa[:, 0][None, :, None] + s[:, x1] * xa[:, 1][None, :, None]
    b0, b1 = yb[:, 0][:, None, None], yb[:, 1][:, None, None]
    out = (((b0 * (hr[r0] >> 4)) >> 16) + ((b1 * (hr[r1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def frames_to_clip_np(frames, size, start=0, every=1, n_frames=None, hflip=False, bgr=False):
    """SingleVideoDataset.__getitem__ for one window (video.py:45-86) on already decoded uint8 frames (F, Hs, Ws, 3):
    BGR->RGB + resize (generate_frames.py:42-46), frames[start : ... : every] (video.py:52), /255 (video.py:56),
    horizontal flip (video.py:78-79), Normalize(mean .5, std .5)‡ (video.py:81-82), (C, T, H, W) (video.py:84).
    Returns float32 (1, 3, T, H, W)."""
    frames = np.asarray(frames, dtype=np.uint8)
    T = int(n_frames) if n_frames is not None else (frames.shape[0] - start + every - 1) // every
    sel = frames[start:start + (T - 1) * every + 1:every]
    if bgr:
        sel = sel[..., ::-1]
    res = np.stack([cv_resize_linear_u8(f, size) for f in sel])                  # (T, H, W, 3)
    x = res.transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255)           # (T, C, H, W) in [0, 1]
    if hflip:
        x = np.flip(x, -1)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x.transpose(1, 0, 2, 3)[None])


# ----------------------------------------------------------------------------------------------------------------
# Linear resize (UpsampleTrilinear3D / ResizeBilinear, align_corners)                 (src/tools/trilinear.py:171-254)
# ----------------------------------------------------------------------------------------------------------------
def linear_taps(n_in, n_out, align_corners=True):
    """Per-axis source indices and fp32 weights, ATen/MindSpore rule (validated against the KAT of
    trilinear.py:222-233).  Returns (i0 int32[n_out], i1 int32[n_out], l0 f32[n_out], l1 f32[n_out]).
    All arithmetic in IEEE fp32, in this order: scale = (in-1)/(out-1); r = scale*o; i0 = (int)r;
    i1 = i0 + (i0 < in-1); l1 = r - i0; l0 = 1 - l1.   (align_corners=False: r = max(scale*(o+0.5)-0.5, 0),
    scale = in/out.)"""
    o = np.arange(n_out, dtype=np.float32)
    if align_corners:
        scale = np.float32(0.0) if n_out <= 1 else np.float32(np.float32(n_in - 1) / np.float32(n_out - 1))
        r = (scale * o).astype(np.float32)
    else:
        scale = np.float32(np.float32(n_in) / np.float32(n_out))
        r = (scale * (o + np.float32(0.5)) - np.float32(0.5)).astype(np.float32)
        r = np.maximum(r, np.float32(0.0))
    i0 = r.astype(np.int32)
    i0 = np.minimum(i0, n_in - 1)
    i1 = i0 + (i0 < n_in - 1).astype(np.int32)
    l1 = (r - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1.0) - l1).astype(np.float32)
    return i0, i1, l0, l1


def resize_linear_np(x, size, align_corners=True):
    """x: np.float32 (N, C, *spatial) with 2 or 3 spatial dims; size: output spatial dims.
    Nested lerp, innermost axis first (the ATen/MindSpore evaluation order), fp32 throughout."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    nsp = x.ndim - 2
    assert nsp == len(size)
    y = x
    for ax in range(nsp - 1, -1, -1):
        n_in = y.shape[2 + ax]
        i0, i1, l0, l1 = linear_taps(n_in, int(size[ax]), align_corners)
        a = np.take(y, i0, axis=2 + ax)
        b = np.take(y, i1, axis=2 + ax)
        shp = [1] * y.ndim
        shp[2 + ax] = -1
        y = (l0.reshape(shp) * a + l1.reshape(shp) * b).astype(np.float32)
    return y


def resize_linear_bwd_np(gy, in_size, align_corners=True):
    """Adjoint of resize_linear_np: gy (N,C,*out) -> gx (N,C,*in_size).  float64 accumulation then fp32."""
    g = np.asarray(gy, dtype=np.float64)
    nsp = g.ndim - 2
    for ax in range(nsp):
        n_out = g.shape[2 + ax]
        n_in = int(in_size[ax])
        i0, i1, l0, l1 = linear_taps(n_in, n_out, align_corners)
        shp = list(g.shape)
        shp[2 + ax] = n_in
        out = np.zeros(shp, dtype=np.float64)
        gm = np.moveaxis(g, 2 + ax, 0)
        om = np.moveaxis(out, 2 + ax, 0)
        for o in range(n_out):
            om[i0[o]] += float(l0[o]) * gm[o]
            om[i1[o]] += float(l1[o]) * gm[o]
        g = out
    return g.astype(np.float32)


def resize_linear(x, size, align_corners=True):
    """torch (differentiable) version of resize_linear_np; same evaluation order."""
    nsp = x.dim() - 2
    y = x
    for ax in range(nsp - 1, -1, -1):
        n_in = y.shape[2 + ax]
        i0, i1, l0, l1 = linear_taps(n_in, int(size[ax]), align_corners)
        a = y.index_select(2 + ax, torch.from_numpy(i0.astype(np.int64)))
        b = y.index_select(2 + ax, torch.from_numpy(i1.astype(np.int64)))
        shp = [1] * y.dim()
        shp[2 + ax] = -1
        y = torch.from_numpy(l0).reshape(shp) * a + torch.from_numpy(l1).reshape(shp) * b
    return y


def upscale(video, index, opt):
    """images.py:96-107: resize to the (T,H,W) of pyramid level `index`, align_corners=True."""
    assert index > 0
    return resize_linear(video, scale_shape(opt, index), True)


def upscale_2d(image, index, opt):
    """images.py:110-119."""
    assert index > 0
    return resize_linear(image, scale_shape_2d(opt, index), True)


# ----------------------------------------------------------------------------------------------------------------
# Parameters (names = the reference's checkpoint contract, src/tools/pt2ms.py:129-188)
# ----------------------------------------------------------------------------------------------------------------
def _l2normalize_np(x, eps=1e-12):
    return (x / np.sqrt(max(float(np.sum(x.astype(np.float64) ** 2)), eps))).astype(np.float32)


def _conv_params(rng, prefix, cin, cout, k, nd, p, sn=False):
    shape = (cout, cin) + (k,) * nd
    p[prefix + "weight"] = (rng.standard_normal(shape) * 0.02).astype(np.float32)   # Normal(0.02, 0.0)
    p[prefix + "bias"] = np.zeros((cout,), np.float32)
    if sn:  # spectral_norm.py:137-140
        p[prefix + "weight_u"] = _l2normalize_np(rng.standard_normal((cout, 1)).astype(np.float32))
        p[prefix + "weight_v"] = _l2normalize_np(rng.standard_normal((cin * k ** nd, 1)).astype(np.float32))


def _bn_params(rng, prefix, c, p, nd):
    pre = prefix + ("bn2d." if nd == 3 else "")   # 3D BN wraps a 2D BN (pt2ms.py:173)
    p[pre + "gamma"] = (1.0 + rng.standard_normal((c,)) * 0.02).astype(np.float32)     # Normal(0.02, 1.0)
    p[pre + "beta"] = np.zeros((c,), np.float32)
    p[pre + "moving_mean"] = np.zeros((c,), np.float32)
    p[pre + "moving_variance"] = np.ones((c,), np.float32)


def _block_params(rng, prefix, cin, opt, p, nd):
    """One decoder / body block: ConvBlock(cin->N) + num_layer x ConvBlock(N->N) + Conv(N->nc_im)
    (networks_3d.py:377-381, 395-401)."""
    N, k = opt.nfc, opt.ker_size
    _conv_params(rng, f"{prefix}0.0.", cin, N, k, nd, p)
    _bn_params(rng, f"{prefix}0.1.", N, p, nd)
    for j in range(1, opt.num_layer + 1):
        _conv_params(rng, f"{prefix}{j}.0.", N, N, k, nd, p)
        _bn_params(rng, f"{prefix}{j}.1.", N, p, nd)
    _conv_params(rng, f"{prefix}{opt.num_layer + 1}.", N, opt.nc_im, k, nd, p)


def init_generator_params(opt, n_body, seed=0, nd=3):
    """Random-init parameters of GeneratorHPVAEGAN with `n_body` refinement stages (networks_3d.py:354-404).
    init_next_stage deep-copies the previous stage (networks_3d.py:404) — mirrored here."""
    rng = np.random.default_rng(seed)
    p = {}
    N, k = opt.nfc, opt.ker_size
    _conv_params(rng, "encode._features.0.0.", opt.nc_im, N, k, nd, p, sn=True)
    for i in range(1, opt.enc_blocks + 1):
        _conv_params(rng, f"encode._features.{i}.0.", N, N, k, nd, p, sn=True)
    _conv_params(rng, "encode._mu.0.", N, opt.latent_dim, k, nd, p)
    _conv_params(rng, "encode._logvar.0.", N, opt.latent_dim, k, nd, p)
    _block_params(rng, "decoder.", opt.latent_dim, opt, p, nd)
    for s in range(n_body):
        if s == 0:
            _block_params(rng, "body.0.", opt.nc_im, opt, p, nd)
        else:
            for key in [q for q in p if q.startswith(f"body.{s - 1}.")]:
                p[key.replace(f"body.{s - 1}.", f"body.{s}.", 1)] = p[key].copy()
    return p


def randomize_bn_stats(p, seed=1, opt=None, n_calib=1, nd=3):
    """Trained-checkpoint stand-in: moving statistics := the batch statistics of one random-mode forward (so
    activations stay O(1) instead of vanishing/saturating), then perturbed by a few percent so that eval-mode
    parity tests exercise a non-trivial BatchNorm fold; beta / hidden biases get small random values."""
    global BN_MOMENTUM
    rng = np.random.default_rng(seed)
    for k in list(p):
        if k.endswith("beta") or (k.endswith("bias") and p[k].shape[0] > 3):
            p[k] = (rng.standard_normal(p[k].shape) * 0.05).astype(np.float32)
    if opt is None:
        opt = default_opt()
    pt = to_torch(p)
    nb = n_body(pt)
    shp = scale_shape if nd == 3 else scale_shape_2d
    z = torch.from_numpy(rng.standard_normal((n_calib, opt.latent_dim) + shp(opt, 0)).astype(np.float32))
    noises = {s: torch.from_numpy(rng.standard_normal((n_calib, opt.nc_im) + shp(opt, s)).astype(np.float32))
              for s in range(1, nb + 1)}
    saved, BN_MOMENTUM = BN_MOMENTUM, 0.0
    try:
        with torch.no_grad():
            generator_forward(None, [1.0] + [0.1] * nb, pt, opt, noise_init=z, is_random=True, training=True,
                              noises=noises, nd=nd)
    finally:
        BN_MOMENTUM = saved
    for k in list(p):
        if k.endswith("moving_mean"):
            p[k] = (pt[k].numpy() + rng.standard_normal(p[k].shape) * 0.02).astype(np.float32)
        elif k.endswith("moving_variance"):
            p[k] = (pt[k].numpy() * (1.0 + 0.1 * (rng.random(p[k].shape) - 0.5))).astype(np.float32)
    return p


def init_discriminator_params(opt, seed=0, nd=3):
    """WDiscriminator3D (networks_3d.py:170-187): SN head nc_im->N, num_layer SN body blocks, plain tail N->1."""
    rng = np.random.default_rng(seed + 1000)
    p = {}
    N, k = opt.nfc, opt.ker_size
    _conv_params(rng, "head.0.", opt.nc_im, N, k, nd, p, sn=True)
    for j in range(opt.num_layer):
        _conv_params(rng, f"body.{j}.0.", N, N, k, nd, p, sn=True)
    _conv_params(rng, "tail.", N, 1, k, nd, p)
    return p


def to_torch(p, requires_grad=()):
    out = {}
    for k, v in p.items():
        t = torch.from_numpy(np.array(v, copy=True))
        if any(k.startswith(pref) for pref in requires_grad) and not (
                k.endswith("weight_u") or k.endswith("weight_v") or "moving_" in k):
            t.requires_grad_(True)
        out[k] = t
    return out


# ----------------------------------------------------------------------------------------------------------------
# Optional bf16 emulation.  The CUDA path stores weights and inter-layer activations in bfloat16.  LeakyReLU makes
# the GRADIENT a discontinuous function of the forward activations, so comparing gradients against a pure-fp32
# forward measures mask flips of near-zero pre-activations (~0.15 % of elements -> ~4 % rel-L2 per layer), not the
# backward kernels.  Inside `with bf16_emulation():` the oracle rounds at exactly the points where the CUDA path
# stores bf16 (straight-through gradient), which makes the gradient comparison a same-inputs comparison.
# ----------------------------------------------------------------------------------------------------------------
_QUANT = [False]


class _STQuant(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


def _q(x):
    return _STQuant.apply(x) if _QUANT[0] else x


class bf16_emulation:
    def __enter__(self):
        self._old = _QUANT[0]
        _QUANT[0] = True

    def __exit__(self, *a):
        _QUANT[0] = self._old


# ----------------------------------------------------------------------------------------------------------------
# Layers
# ----------------------------------------------------------------------------------------------------------------
def _conv(x, w, b, pad):
    w = _q(w)
    return F.conv3d(x, w, b, padding=pad) if x.dim() == 5 else F.conv2d(x, w, b, padding=pad)


def lrelu(x):
    return F.leaky_relu(x, LRELU_SLOPE)


def reflect_conv(x, w, b=None, act=False):
    """The `bn=False` branch of ConvBlock3DSN / ConvBlock2DSN (reference networks_3d.py:64-73, networks_2d.py:63-72):
    nn.Pad(mode='REFLECT') by one voxel on every spatial (and temporal) side, then a pad_mode='valid' convolution
    (bias-free in 3-D, with bias in 2-D) [+ LeakyReLU(0.2)]."""
    nd = w.dim() - 2
    y = F.conv3d(F.pad(x, (1,) * 6, mode="reflect"), w, b) if nd == 3 else F.conv2d(F.pad(x, (1,) * 4, mode="reflect"), w, b)
    return lrelu(y) if act else y


def sn_power_iteration(w, u, v, eps=1e-12):
    """spectral_norm.py:142-151: one power iteration; returns (sigma, u_new, v_new).  u, v are constants for
    autodiff (they are re-assigned onto non-trainable Parameters, spectral_norm.py:147-148)."""
    w2 = w.reshape(w.shape[0], -1)
    with torch.no_grad():
        v_new = w2.t() @ u
        v_new = v_new / torch.sqrt(torch.clamp((v_new ** 2).sum(), min=eps))
        u_new = w2 @ v_new
        u_new = u_new / torch.sqrt(torch.clamp((u_new ** 2).sum(), min=eps))
    sigma = (u_new.t() @ w2 @ v_new).reshape(())
    return sigma, u_new, v_new


def sn_conv(x, p, prefix, pad, update=True):
    """SpectualNormConv3d.construct (spectral_norm.py:142-155).  Mutates p[...weight_u/v] when update (Q5)."""
    w = p[prefix + "weight"]
    sigma, u_new, v_new = sn_power_iteration(w, p[prefix + "weight_u"], p[prefix + "weight_v"])
    if update:
        p[prefix + "weight_u"], p[prefix + "weight_v"] = u_new, v_new
    if _QUANT[0]:   # the CUDA path convolves with bf16(W) and applies 1/sigma and the bias in the fp32 epilogue
        bshape = [1, -1] + [1] * (x.dim() - 2)
        return _conv(x, w, None, pad) / sigma + p[prefix + "bias"].reshape(bshape)
    return _conv(x, w / sigma, p[prefix + "bias"], pad)


def batchnorm(x, p, prefix, training, update=True):
    """mindspore.nn.BatchNorm3d/2d (networks_3d.py:52).  training: batch stats (biased var) + moving update."""
    nd = x.dim() - 2
    pre = prefix + ("bn2d." if nd == 3 else "")
    shp = [1, -1] + [1] * nd
    g, b = p[pre + "gamma"].reshape(shp), p[pre + "beta"].reshape(shp)
    if training:
        dims = [0] + list(range(2, x.dim()))
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        if update:
            with torch.no_grad():
                p[pre + "moving_mean"] = BN_MOMENTUM * p[pre + "moving_mean"] + (1 - BN_MOMENTUM) * mean
                p[pre + "moving_variance"] = BN_MOMENTUM * p[pre + "moving_variance"] + (1 - BN_MOMENTUM) * var
    else:
        mean, var = p[pre + "moving_mean"], p[pre + "moving_variance"]
    return (x - mean.reshape(shp)) / torch.sqrt(var.reshape(shp) + BN_EPS) * g + b


def conv_block(x, p, prefix, pad, training, bn=True, act=True, taps=None):
    """ConvBlock3D / ConvBlock2D (networks_3d.py:45-54): conv(+bias) -> BN -> LeakyReLU(0.2)."""
    y = _conv(x, p[prefix + "0.weight"], p[prefix + "0.bias"], pad)
    if taps is not None:
        taps[prefix + "conv"] = y
    if bn:
        if training and _QUANT[0]:
            # training mode stores the raw conv output in bf16 before the statistics pass — MINUS the estimated position
            # of the LeakyReLU kink (kink-centred storage, csrc/elementwise.cu bn_center_multi_kernel): the offset comes
            # from the moving statistics BEFORE this forward's update and the current gamma / beta, and is itself a
            # bf16 value
            nd = y.dim() - 2
            pre = prefix + "1." + ("bn2d." if nd == 3 else "")
            g, b = p[pre + "gamma"].detach(), p[pre + "beta"].detach()
            cen = p[pre + "moving_mean"].detach().clone()
            ok = g.abs() > 1e-3
            cen[ok] = cen[ok] - (b[ok] * torch.sqrt(p[pre + "moving_variance"].detach()[ok] + BN_EPS) / g[ok])
            cen = cen.to(torch.bfloat16).to(torch.float32).reshape([1, -1] + [1] * nd)
            y = _q(y - cen) + cen
        y = batchnorm(y, p, prefix + "1.", training)
    if act:
        y = lrelu(y)
    y = _q(y)
    if taps is not None:
        taps[prefix + "out"] = y
    return y


def block_forward(x, p, prefix, opt, training, taps=None):
    """decoder / body[s] SequentialCell (networks_3d.py:377-381, 395-401)."""
    y = x
    for j in range(opt.num_layer + 1):
        y = conv_block(y, p, f"{prefix}{j}.", opt.padd_size, training, taps=taps)
    j = opt.num_layer + 1
    y = _conv(y, p[f"{prefix}{j}.weight"], p[f"{prefix}{j}.bias"], opt.ker_size // 2)
    if taps is not None:
        taps[f"{prefix}{j}.conv"] = y
    return y


def encode(x, p, opt, update_sn=True):
    """Encode3DVAE.construct (networks_3d.py:107-112): FeatureExtractor (3 SN blocks) -> mu, logvar convs."""
    pad = opt.ker_size // 2
    f = _q(x)
    for i in range(opt.enc_blocks + 1):
        f = _q(lrelu(sn_conv(f, p, f"encode._features.{i}.0.", pad, update_sn)))
    mu = _q(_conv(f, p["encode._mu.0.weight"], p["encode._mu.0.bias"], pad))
    logvar = _q(_conv(f, p["encode._logvar.0.weight"], p["encode._logvar.0.bias"], pad))
    return mu, logvar


def discriminator(x, p, opt, update_sn=True):
    """WDiscriminator3D.construct (networks_3d.py:189-193)."""
    pad = opt.ker_size // 2
    h = _q(lrelu(sn_conv(_q(x), p, "head.0.", pad, update_sn)))
    for j in range(opt.num_layer):
        h = _q(lrelu(sn_conv(h, p, f"body.{j}.0.", pad, update_sn)))
    return _conv(h, p["tail.weight"], p["tail.bias"], 1)


def n_body(p):
    s = 0
    while f"body.{s}.0.0.weight" in p:
        s += 1
    return s


def refinement_layers(start_idx, x_prev_out, noise_amp, p, opt, is_random, training, noises=None, nd=3, taps=None):
    """GeneratorHPVAEGAN.refinement_layers (networks_3d.py:434-451 / networks_2d.py:266-282).
    noises: dict idx+1 -> tensor (the N(0,1) draw of images.py:30-37) injected for reproducibility."""
    for idx in range(start_idx, n_body(p)):
        if opt.vae_levels == idx + 1 and not opt.train_all:
            x_prev_out = x_prev_out.detach()                                      # networks_3d.py:437-438
        up = upscale(x_prev_out, idx + 1, opt) if nd == 3 else upscale_2d(x_prev_out, idx + 1, opt)
        add_noise = is_random and (opt.vae_levels <= idx + 1 if nd == 3 else True)  # 2D: every scale (networks_2d.py:274)
        if add_noise:
            x_in = up + noises[idx + 1] * noise_amp[idx + 1]
        else:
            x_in = up
        x_in = _q(x_in)
        if taps is not None:
            taps[f"body.{idx}.in"] = x_in
        x_prev = block_forward(x_in, p, f"body.{idx}.", opt, training, taps=taps)
        x_prev_out = torch.tanh(x_prev + up)
        if taps is not None:
            taps[f"body.{idx}.out"] = x_prev_out
    return x_prev_out


def generator_forward(video, noise_amp, p, opt, noise_init=None, sample_init=None, is_random=False, training=False,
                      is_training_flag=False, eps=None, z_pred=None, noises=None, nd=3, update_sn=True, taps=None):
    """GeneratorHPVAEGAN.construct (networks_3d.py:406-432).  `eps` / `z_pred` / `noises` are the host-side
    N(0,1) draws (networks_3d.py:28-34, images.py:30-37) passed in explicitly."""
    mu = logvar = None
    if noise_init is None:
        mu, logvar = encode(video, p, opt, update_sn)
        if is_training_flag:
            z_vae = eps * torch.exp(logvar * 0.5) + mu
        else:
            z_vae = z_pred                                                       # Q2: pure N(0,1)
    else:
        z_vae = noise_init
    vae_out = torch.tanh(block_forward(_q(z_vae), p, "decoder.", opt, training, taps=taps))
    if sample_init is None:
        x = refinement_layers(0, vae_out, noise_amp, p, opt, is_random, training, noises, nd, taps)
    else:
        x = refinement_layers(sample_init[0], sample_init[1], noise_amp, p, opt, is_random, training, noises, nd, taps)
    if noise_init is None:
        return x, vae_out, mu, logvar
    return x, vae_out


# ----------------------------------------------------------------------------------------------------------------
# Losses                                                                               (src/modules/losses.py)
# ----------------------------------------------------------------------------------------------------------------
def kl_criterion(mu, logvar):
    """losses.py:5-7."""
    return (-0.5 * (1 + logvar - mu ** 2 - torch.exp(logvar))).mean()


def mse(a, b):
    return ((a - b) ** 2).mean()


def gradient_penalty(real, fake, alpha, pd, opt):
    """DWithLoss.calc_gradient_penalty (losses.py:47-52)."""
    x_hat = (alpha * real + (1 - alpha) * fake).detach().requires_grad_(True)
    out = discriminator(x_hat, pd, opt)
    grad, = torch.autograd.grad(out.sum(), x_hat, create_graph=True)
    norm = torch.sqrt((grad ** 2).sum(dim=1))
    return ((norm - 1) ** 2).mean() * opt.lambda_grad


def d_loss(real, fake, alpha, pd, opt):
    """DWithLoss.construct (losses.py:27-45) given the (stop-gradiented) fake."""
    err_real = -discriminator(real, pd, opt).mean()
    err_fake = discriminator(fake.detach(), pd, opt).mean()
    gp = gradient_penalty(real, fake.detach(), alpha, pd, opt)
    return err_real + err_fake + gp


# ----------------------------------------------------------------------------------------------------------------
# Optimiser                                                                        (src/modules/optimizers.py)
# ----------------------------------------------------------------------------------------------------------------
def clip_by_norm(g, clip):
    """mindspore.nn.ClipByNorm (optimizers.py:29): g*c / max(||g||2, c)."""
    n = math.sqrt(float((g.astype(np.float64) ** 2).sum()))
    return (g * np.float32(clip) / np.float32(max(n, clip))).astype(np.float32)


def adam_step(w, g, m, v, step, lr, beta1=0.5, beta2=0.999, eps=1e-8):
    """mindspore.nn.Adam (ops.Adam): m += (g-m)(1-b1); v += (g^2-v)(1-b2);
    w -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps).  step is 1-based."""
    w, g, m, v = (np.asarray(a, np.float32) for a in (w, g, m, v))
    m = (m + (g - m) * np.float32(1 - beta1)).astype(np.float32)
    v = (v + (g * g - v) * np.float32(1 - beta2)).astype(np.float32)
    lr_t = np.float32(lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step))
    w = (w - lr_t * m / (np.sqrt(v) + np.float32(eps))).astype(np.float32)
    return w, m, v


def g_loss(real, real_zero, noise_init, noise_amps, pg, pd, opt, is_vae, z_pred=None, eps=None, noises=None,
           is_training_flag=False, nd=3):
    """GWithLoss.construct (losses.py:70-103) in set_train() mode (BatchNorm batch statistics)."""
    x, vae_out, mu, logvar = generator_forward(real_zero, noise_amps, pg, opt, is_random=False, training=True,
                                               is_training_flag=is_training_flag, eps=eps, z_pred=z_pred, nd=nd)
    if is_vae:
        rec = mse(x, real) + mse(vae_out, real_zero)
        return opt.rec_weight * rec + opt.kl_weight * kl_criterion(mu, logvar)
    total = opt.rec_weight * mse(x, real)
    fake, _ = generator_forward(None, noise_amps, pg, opt, noise_init=noise_init, is_random=True, training=True,
                                noises=noises, nd=nd)
    fake = fake.detach()                                                          # Q1 (losses.py:94)
    return total + (-discriminator(fake, pd, opt).mean() * opt.disc_loss_weight)


# ----------------------------------------------------------------------------------------------------------------
# sinFID                                                                    (src/sinFID/c3d.py, inception.py, fid_score.py)
# ----------------------------------------------------------------------------------------------------------------
def c3d_block0(x, w, b, normalize_input=False):
    """C3D block 0 (c3d.py:62-66): conv1 = Conv3d(3, 64, 3x3x3, pad 1, bias), no activation inside the block;
    c3d.py:129-130: inputs in (0, 1) are scaled to (-1, 1) first when normalize_input."""
    x = torch.as_tensor(x)
    if normalize_input:
        x = 2 * x - 1
    return F.conv3d(x, torch.as_tensor(w), torch.as_tensor(b), padding=1)


INCEPTION_BN_EPS = 1e-3     # BasicConv2d's BatchNorm2d(eps=0.001) in InceptionV3


def inception_block0(x, params, normalize_input=False):
    """InceptionV3 block 0 (inception.py:66-72): Conv2d_1a_3x3 (3->32, k3, stride 2, no pad), Conv2d_2a_3x3 (32->32, k3,
    no pad), Conv2d_2b_3x3 (32->64, k3, pad 1); each conv(no bias) + BatchNorm(eval, eps 1e-3) + ReLU."""
    x = torch.as_tensor(x)
    if normalize_input:
        x = 2 * x - 1
    for name, stride, pad in (("Conv2d_1a", 2, 0), ("Conv2d_2a", 1, 0), ("Conv2d_2b", 1, 1)):
        x = F.conv2d(x, torch.as_tensor(params[name + ".conv.weight"]), None, stride=stride, padding=pad)
        g, b = torch.as_tensor(params[name + ".bn.gamma"]), torch.as_tensor(params[name + ".bn.beta"])
        m, v = torch.as_tensor(params[name + ".bn.moving_mean"]), torch.as_tensor(params[name + ".bn.moving_variance"])
        x = (x - m.reshape(1, -1, 1, 1)) / torch.sqrt(v.reshape(1, -1, 1, 1) + INCEPTION_BN_EPS) * g.reshape(1, -1, 1, 1) \
            + b.reshape(1, -1, 1, 1)
        x = F.relu(x)
    return x


def activation_statistics(feat):
    """fid_score.py:92-93,160-178 for ONE sample: positions are the observations — pred.transpose(0,2,3,1).reshape(-1, C),
    mu = mean, sigma = np.cov(rowvar=False).  feat: (1, C, [T,] H, W)."""
    f = np.asarray(feat, np.float64)
    act = np.moveaxis(f[0], 0, -1).reshape(-1, f.shape[1])
    return act.mean(axis=0), np.cov(act, rowvar=False)


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """fid_score.py:105-159 (Sutherland's stable form)."""
    from scipy import linalg
    diff = mu1 - mu2
    try:
        covmean, _ = linalg.sqrtm(sigma1.dot(sigma2), disp=False)
    except TypeError:       # scipy >= 1.16 dropped `disp` (and the error-estimate return value)
        covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


def svfid(real, fakes, feature_fn):
    """calculate_SVFID / calculate_SIFID (fid_score.py:181-242): per fake sample the Fréchet distance between its
    position statistics and the real sample's, averaged in float32.  real: (1,3,...); fakes: (n,3,...)."""
    m1, s1 = activation_statistics(feature_fn(real))
    vals = []
    for i in range(len(fakes)):
        m2, s2 = activation_statistics(feature_fn(fakes[i:i + 1]))
        vals.append(frechet_distance(m1, s1, m2, s2))
    return float(np.asarray(vals, np.float32).mean()), vals
