"""Device-side twin of the reference's `src/datasets/video.py:SingleVideoDataset` / `image.py:SingleImageDataset`
(SURVEY.md §8f-4).  The reference re-decodes and re-resizes the whole video on the host for EVERY item
(video.py:52 -> generate_frames.py:23-52); here the decoded uint8 frames are uploaded once and every item is one
kernel (`hpvg_frames_to_clip`): cv2-exact fixed-point bilinear resize, frame window, /255, flip, Normalize, CTHW.

Decoding itself (cv2.VideoCapture) stays with the caller: pass the decoded frames, RGB or decoder-order BGR.
"""
import random

import numpy as np

from . import ops
from .runtime import Tensor, F32, from_numpy
from .utils import images as uimg

__all__ = ["SingleVideoDataset", "SingleImageDataset"]


class SingleVideoDataset:
    """video.py:13-94.  `frames`: uint8 numpy (F, H, W, 3) decoded frames (after `--start-frame` / `--max-frames`
    trimming, generate_frames.py:23-27).  Sets the same dataset-derived options the reference sets on `opt`
    (org_fps is the caller's: it comes from the container header)."""

    def __init__(self, opt, frames, bgr=False, stream=None):
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        if frames.ndim != 4 or frames.shape[-1] != 3:
            raise ValueError("frames must be (F, H, W, 3) uint8")
        self.opt = opt
        self.bgr = bool(bgr)
        self.stream = stream
        self.org_frame_size = [float(frames.shape[1]), float(frames.shape[2])]
        opt.ar = self.org_frame_size[0] / self.org_frame_size[1]            # H2W (video.py:33)
        opt.fps_lcm = int(np.lcm.reduce(opt.sampling_rates))               # video.py:35
        self.n_frames = int(frames.shape[0])
        self.d_frames = from_numpy(frames, stream=stream)                  # resident: decoded once, uploaded once

    def __len__(self):
        return (self.n_frames - self.opt.fps_lcm) * getattr(self.opt, "data_rep", 1)          # video.py:42-43

    def scaled_size(self, scale_idx):
        """video.py:88-92: [int(base * ar), base]."""
        base = uimg.get_scales_by_index(scale_idx, self.opt.scale_factor, self.opt.stop_scale, self.opt.img_size)
        return [int(base * self.opt.ar), base]

    def _clip(self, scale_idx, idx, every, hflip):
        H, W = self.scaled_size(scale_idx)
        T = self.opt.fps_lcm // every + 1                                  # frames[idx : idx + lcm + 1 : every]
        return ops.frames_to_clip(self.d_frames, (H, W), start=idx, every=every, n_frames=T, hflip=hflip,
                                  bgr=self.bgr, stream=self.stream)

    def __getitem__(self, idx):
        """video.py:45-73: (frames at opt.scale_idx, frames at scale 0 — zeros at scale 0), both (1, 3, T, H, W)."""
        opt = self.opt
        idx = idx % (self.n_frames - opt.fps_lcm)
        hflip = random.random() < 0.5 if getattr(opt, "hflip", False) else False
        every = opt.sampling_rates[opt.fps_index]
        frames = self._clip(opt.scale_idx, idx, every, hflip)
        if opt.scale_idx > 0:
            return frames, self._clip(0, idx, opt.sampling_rates[0], hflip)
        return frames, Tensor(frames.shape, F32).zero_(self.stream)


class SingleImageDataset:
    """image.py:13-60, same idea for one image: (1, 3, H, W) clips through the T == 1 case of the kernel."""

    def __init__(self, opt, image, bgr=False, stream=None):
        image = np.ascontiguousarray(image, dtype=np.uint8)
        if image.ndim != 3 or image.shape[-1] != 3:
            raise ValueError("image must be (H, W, 3) uint8")
        self.opt, self.bgr, self.stream = opt, bool(bgr), stream
        opt.ar = image.shape[0] / image.shape[1]
        self.d_image = from_numpy(image[None], stream=stream)

    def __len__(self):
        return getattr(self.opt, "data_rep", 1)

    def scaled_size(self, scale_idx):
        base = uimg.get_scales_by_index(scale_idx, self.opt.scale_factor, self.opt.stop_scale, self.opt.img_size)
        return [int(base * self.opt.ar), base]

    def _image(self, scale_idx, hflip):
        H, W = self.scaled_size(scale_idx)
        clip = ops.frames_to_clip(self.d_image, (H, W), n_frames=1, hflip=hflip, bgr=self.bgr, stream=self.stream)
        return clip.view((1, 3, H, W))

    def __getitem__(self, idx):
        opt = self.opt
        hflip = random.random() < 0.5 if getattr(opt, "hflip", False) else False
        img = self._image(opt.scale_idx, hflip)
        if opt.scale_idx > 0:
            return img, self._image(0, hflip)
        return img, Tensor(img.shape, F32).zero_(self.stream)
