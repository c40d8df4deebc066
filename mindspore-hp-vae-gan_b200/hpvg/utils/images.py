"""Pyramid geometry, up-scaling and host noise generators — mirror of the reference's `src/utils/images.py`
(same function names and argument order), with the resize running on the GPU through libhpvg."""
import math
from types import SimpleNamespace

import numpy as np

from .. import ops
from ..runtime import from_numpy

__all__ = ["interpolate", "interpolate_3D", "adjust_scales2image", "generate_noise_size", "generate_noise_ref",
           "get_scales_by_index", "get_fps_td_by_index", "get_fps_by_index", "upscale", "upscale_2d", "scale_shape",
           "scale_shape_2d", "default_opt"]


def generate_noise_size(size=None, type="normal"):
    """images.py:17-27 — host numpy draw (global RNG, Q7), uploaded."""
    if type == "normal":
        return from_numpy(np.random.normal(size=size).astype("float32"))
    if type == "benoulli":
        return from_numpy(np.random.binomial(1, 0.5, size=size).astype("float32"))
    return from_numpy(np.random.uniform(0, 1, size=size).astype("float32"))


def generate_noise_ref(ref_shape, type="normal"):
    """images.py:30-37."""
    return generate_noise_size(tuple(ref_shape), type)


def interpolate_3D(input, size):
    """images.py:54-61: UpsampleTrilinear3D(size, align_corners=True)."""
    if len(input.shape) != 5:
        raise ValueError("interpolate_3D expects a 5-D tensor")
    return ops.resize3d(input, tuple(int(v) for v in size), align_corners=True)


def interpolate(input, size=None):
    """images.py:40-51: ResizeBilinear(size, align_corners=True) on (N,C,H,W) — the T == 1 case of the 3-D kernel."""
    if len(input.shape) == 4:
        n, c, h, w = input.shape
        y = ops.resize3d(input.view((n, c, 1, h, w)), (1, int(size[0]), int(size[1])), align_corners=True)
        return y.view((n, c, int(size[0]), int(size[1])))
    n, c, t, h, w = input.shape
    return ops.resize3d(input, (t, int(size[0]), int(size[1])), align_corners=True)


def adjust_scales2image(size, opt):
    """images.py:64-71."""
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
    """(T, H, W) of pyramid level `index` (images.py:99-101)."""
    s = get_scales_by_index(index, opt.scale_factor, opt.stop_scale, opt.img_size)
    _, td, _ = get_fps_td_by_index(index, opt.stop_scale_time, opt.sampling_rates, opt.org_fps, opt.fps_lcm)
    # synthetic benchmark clips may pin the time depth of a level (BASELINE.json config 3: "16-frame clips at the
    # finest scale"); the reference always derives it from the sampling rates
    td = getattr(opt, "td_override", {}).get(index, td)
    return (td, int(s * opt.ar), s)


def scale_shape_2d(opt, index):
    s = get_scales_by_index(index, opt.scale_factor, opt.stop_scale, opt.img_size)
    return (int(s * opt.ar), s)


def upscale(video, index, scale_factor, stop_scale, img_size, stop_scale_time, sampling_rates, org_fps, fps_lcm, ar):
    """images.py:96-107."""
    if index <= 0:
        raise ValueError("upscale: index must be positive")
    next_shape = get_scales_by_index(index, scale_factor, stop_scale, img_size)
    _, next_td, _ = get_fps_td_by_index(index, stop_scale_time, sampling_rates, org_fps, fps_lcm)
    return interpolate_3D(video, size=[next_td, int(next_shape * ar), next_shape])


def upscale_2d(image, index, scale_factor, stop_scale, img_size, ar):
    """images.py:110-119."""
    if index <= 0:
        raise ValueError("upscale_2d: index must be positive")
    next_shape = get_scales_by_index(index, scale_factor, stop_scale, img_size)
    return interpolate(image, size=[int(next_shape * ar), next_shape])


def default_opt(**kw):
    """The argparse defaults of train_video.py:232-294 plus the dataset-derived fields (datasets/video.py:28-35)."""
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
