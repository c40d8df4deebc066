"""GPU parity of the device-side data path: uint8 frames -> normalised fp32 clip, through the C ABI.
Bit-exact bar: the resize is integer arithmetic, the normalisation two correctly rounded fp32 divisions."""
import os
import types

import numpy as np
import pytest

from oracle import hpvg_oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("src,dst", [((45, 60), (24, 33)), ((45, 60), (57, 76)), ((48, 64), (24, 32)),
                                     ((30, 41), (30, 41)), ((186, 248), (192, 257)), ((9, 7), (3, 20))])
@pytest.mark.parametrize("hflip,bgr", [(False, False), (True, True)])
def test_frames_to_clip_matches_oracle_bit_exact(hpvg_gpu, src, dst, hflip, bgr):
    hp = hpvg_gpu
    fr = np.random.default_rng(src[1] * 100 + dst[0]).integers(0, 256, (13,) + src + (3,), dtype=np.uint8)
    got = hp.ops.frames_to_clip(hp.from_numpy(fr), dst, start=1, every=3, n_frames=4, hflip=hflip, bgr=bgr).numpy()
    want = orc.frames_to_clip_np(fr, dst, 1, 3, 4, hflip, bgr)
    assert got.shape == want.shape == (1, 3, 4) + dst
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_frames_to_clip_matches_reference_library_fixtures(hpvg_gpu):
    """tests/golden/frames_small.npz: outputs of cv2 + numpy run statement by statement as the reference does."""
    hp = hpvg_gpu
    g = np.load(os.path.join(GOLDEN, "frames_small.npz"))
    fr = hp.from_numpy(g["frames_bgr"])
    assert np.array_equal(hp.ops.frames_to_clip(fr, (24, 33), 0, 4, 4, False, True).numpy(), g["clip_s0"])
    assert np.array_equal(hp.ops.frames_to_clip(fr, (57, 76), 1, 3, 4, True, True).numpy(), g["clip_up_flip"])
    img = hp.from_numpy(g["image_bgr"][None])
    assert np.array_equal(hp.ops.frames_to_clip(img, (24, 33), n_frames=1, bgr=True).numpy(), g["image_s0"])
    assert np.array_equal(hp.ops.frames_to_clip(img, (48, 64), n_frames=1, bgr=True).numpy(), g["image_half"])


def test_frames_to_clip_rejects_window_past_the_end(hpvg_gpu):
    hp = hpvg_gpu
    fr = hp.from_numpy(np.zeros((5, 8, 8, 3), np.uint8))
    with pytest.raises(hp.HpvgError):
        hp.ops.frames_to_clip(fr, (4, 4), start=2, every=2, n_frames=3)


def test_single_video_dataset_items(hpvg_gpu):
    """src/datasets/video.py:45-73: (clip at opt.scale_idx, clip at scale 0), frame window by the scale's sampling
    rate, zeros for the second item at scale 0; sizes from get_scales_by_index * ar."""
    hp = hpvg_gpu
    from hpvg.datasets import SingleVideoDataset
    from hpvg.utils import images as uimg
    opt = uimg.default_opt(img_size=64)
    fr = np.random.default_rng(3).integers(0, 256, (13, 48, 64, 3), dtype=np.uint8)
    ds = SingleVideoDataset(opt, fr, bgr=True)
    assert opt.ar == 0.75 and opt.fps_lcm == 12 and len(ds) == 1
    for scale_idx in (0, 2, opt.stop_scale):
        opt.scale_idx = scale_idx
        _, td, opt.fps_index = uimg.get_fps_td_by_index(scale_idx, opt.stop_scale_time, opt.sampling_rates,
                                                        opt.org_fps, opt.fps_lcm)
        a, b = ds[0]
        size = tuple(ds.scaled_size(scale_idx))
        every = opt.sampling_rates[opt.fps_index]
        assert a.shape == (1, 3, td) + size
        assert np.array_equal(a.numpy(), orc.frames_to_clip_np(fr, size, 0, every, td, False, True))
        if scale_idx == 0:
            assert not b.numpy().any()
        else:
            assert np.array_equal(b.numpy(), orc.frames_to_clip_np(fr, tuple(ds.scaled_size(0)), 0,
                                                                  opt.sampling_rates[0], 4, False, True))


def test_single_image_dataset_items(hpvg_gpu):
    hp = hpvg_gpu
    from hpvg.datasets import SingleImageDataset
    from hpvg.utils import images as uimg
    g = np.load(os.path.join(GOLDEN, "frames_small.npz"))
    opt = uimg.default_opt(img_size=64)
    ds = SingleImageDataset(opt, g["image_bgr"], bgr=True)
    opt.scale_idx = 1
    a, b = ds[0]
    Ha, Wa = ds.scaled_size(1)
    H0, W0 = ds.scaled_size(0)
    assert a.shape == (1, 3, Ha, Wa) and b.shape == (1, 3, H0, W0)
    assert np.array_equal(a.numpy(), orc.frames_to_clip_np(g["image_bgr"][None], (Ha, Wa), 0, 1, 1, False, True)[:, :, 0])
