#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (torch-CPU fp32).  The reference itself cannot run here
(MindSpore is not installable offline), so these fixtures freeze the ORACLE; the oracle in turn is pinned to the
reference's own golden vectors in tests/test_cpu_oracle.py.  Run from the repo root: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hpvg_oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sample_small():
    img_size, n_body, seed = 64, 4, 11
    opt = orc.default_opt(img_size=img_size)
    p = orc.to_torch(orc.randomize_bn_stats(orc.init_generator_params(opt, n_body, seed=seed), opt=opt))
    rng = np.random.default_rng(5)
    z = rng.standard_normal((2, opt.latent_dim) + orc.scale_shape(opt, 0)).astype(np.float32)
    amps = np.array([1.0, 0.8, 0.6, 0.4, 0.3], np.float32)
    noises = {}
    for s in range(1, n_body + 1):
        if opt.vae_levels <= s:
            noises[s] = rng.standard_normal((2, 3) + orc.scale_shape(opt, s)).astype(np.float32)
    with torch.no_grad():
        x, vae = orc.generator_forward(None, list(amps), p, opt, noise_init=torch.from_numpy(z), is_random=True,
                                       noises={k: torch.from_numpy(v) for k, v in noises.items()})
    out = dict(img_size=img_size, n_body=n_body, seed=seed, z=z, amps=amps, x=x.numpy(), vae=vae.numpy())
    for k, v in noises.items():
        out["noise_%d" % k] = v
    np.savez_compressed(os.path.join(HERE, "sample_small.npz"), **out)
    print("sample_small:", x.shape, float(x.abs().mean()))


def ops_small():
    """Per-operator fixtures (small): resize fwd/bwd + tap tables, BatchNorm(train)+LeakyReLU, one spectral-norm power
    iteration, KL / MSE, ClipByNorm + Adam — the oracle's outputs on fixed seeded inputs."""
    rng = np.random.default_rng(123)
    out = {}
    # resize (images.py:54-61 / trilinear.py:171-254), align_corners=True, scale 0 -> 1 of the default pyramid
    x = rng.standard_normal((1, 3, 4, 24, 33)).astype(np.float32)
    gy = rng.standard_normal((1, 3, 4, 30, 41)).astype(np.float32)
    out["rs_x"], out["rs_gy"] = x, gy
    out["rs_y"] = orc.resize_linear_np(x, (4, 30, 41), True)
    out["rs_gx"] = orc.resize_linear_bwd_np(gy, (4, 24, 33), True)
    i0, i1, l0, l1 = orc.linear_taps(33, 41, True)
    out["rs_i0"], out["rs_i1"], out["rs_l0"], out["rs_l1"] = i0, i1, l0, l1
    # BatchNorm3d (train) + LeakyReLU(0.2) (networks_3d.py:52,20)
    y = (rng.standard_normal((2, 64, 2, 9, 10)) * 1.5 + 0.2).astype(np.float32)
    y = torch.from_numpy(y).to(torch.bfloat16).to(torch.float32).numpy()
    gamma = (1 + 0.1 * rng.standard_normal(64)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(64)).astype(np.float32)
    p = {"1.bn2d.gamma": torch.from_numpy(gamma), "1.bn2d.beta": torch.from_numpy(beta),
         "1.bn2d.moving_mean": torch.zeros(64), "1.bn2d.moving_variance": torch.ones(64)}
    out["bn_y"], out["bn_gamma"], out["bn_beta"] = y, gamma, beta
    out["bn_out"] = orc.lrelu(orc.batchnorm(torch.from_numpy(y), p, "1.", True)).numpy()
    out["bn_mm"], out["bn_mv"] = p["1.bn2d.moving_mean"].numpy(), p["1.bn2d.moving_variance"].numpy()
    # spectral norm (spectral_norm.py:142-151)
    w = (rng.standard_normal((64, 3, 3, 3, 3)) * 0.02).astype(np.float32)
    u = orc._l2normalize_np(rng.standard_normal((64, 1)).astype(np.float32))
    v = orc._l2normalize_np(rng.standard_normal((81, 1)).astype(np.float32))
    sigma, un, vn = orc.sn_power_iteration(torch.from_numpy(w), torch.from_numpy(u), torch.from_numpy(v))
    out["sn_w"], out["sn_u"], out["sn_v"] = w, u, v
    out["sn_sigma"], out["sn_u1"], out["sn_v1"] = np.float32(sigma), un.numpy(), vn.numpy()
    # losses (losses.py:5-7, nn.MSELoss)
    mu = rng.standard_normal((1, 128, 4, 6, 7)).astype(np.float32)
    lv = (0.3 * rng.standard_normal((1, 128, 4, 6, 7))).astype(np.float32)
    out["kl_mu"], out["kl_lv"] = mu, lv
    out["kl"] = np.float32(orc.kl_criterion(torch.from_numpy(mu), torch.from_numpy(lv)))
    out["mse"] = np.float32(orc.mse(torch.from_numpy(mu), torch.from_numpy(lv)))
    # ClipByNorm(5) + Adam (optimizers.py:33-43), two steps
    wt = rng.standard_normal((64, 27)).astype(np.float32)
    g1 = (rng.standard_normal((64, 27)) * 0.5).astype(np.float32)      # ||g|| > 5: clipped
    g2 = (rng.standard_normal((64, 27)) * 0.01).astype(np.float32)     # ||g|| < 5: untouched
    m = np.zeros_like(wt)
    vv = np.zeros_like(wt)
    out["ad_w0"], out["ad_g1"], out["ad_g2"] = wt, g1, g2
    w1, m, vv = orc.adam_step(wt, orc.clip_by_norm(g1, 5.0), m, vv, 1, 5e-4)
    w2, m, vv = orc.adam_step(w1, orc.clip_by_norm(g2, 5.0), m, vv, 2, 5e-4)
    out["ad_w1"], out["ad_w2"] = w1, w2
    np.savez_compressed(os.path.join(HERE, "ops_small.npz"), **out)
    print("ops_small:", sorted(out))


def frames_small():
    """Data-path fixtures (generate_frames.py:42-46, video.py:45-86, image.py:36-75) made by running the reference's
    OWN library calls here — cv2.cvtColor / cv2.resize(INTER_LINEAR) / numpy — on (a) seeded synthetic frames and
    (b) a crop of the reference's real image data/imgs/air_balloons.jpg.  The oracle restatement and the CUDA kernel
    must both reproduce these bits."""
    import cv2

    def reference_clip(frames_bgr, size, start, every, T, hflip):
        mem = []
        for image in frames_bgr:                                           # generate_frames.py:42-47
            rgb = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
            mem.append(cv2.resize(rgb, (size[1], size[0]), interpolation=cv2.INTER_LINEAR))
        frames = np.stack(mem)[start:start + (T - 1) * every + 1:every]    # video.py:52
        frames = frames.transpose(0, 3, 1, 2).astype(np.float32) / 255     # video.py:53-56
        if hflip:
            frames = np.flip(frames, -1)                                   # video.py:78-79
        frames = (frames - np.float32(0.5)) / np.float32(0.5)              # Normalize(mean .5, std .5), video.py:81-82
        return np.ascontiguousarray(frames.transpose(1, 0, 2, 3)[None])    # video.py:84

    rng = np.random.default_rng(2024)
    out = {}
    fr = rng.integers(0, 256, (13, 45, 60, 3), dtype=np.uint8)             # 13 decoded BGR frames, ar 0.75
    out["frames_bgr"] = fr
    out["clip_s0"] = reference_clip(fr, (24, 33), 0, 4, 4, False)          # scale 0: every 4th frame, 4 frames
    out["clip_up_flip"] = reference_clip(fr, (57, 76), 1, 3, 4, True)      # up-sampling + flip + offset window
    path = "/root/reference/data/imgs/air_balloons.jpg"
    if os.path.exists(path):
        img = cv2.imread(path)[40:136, 60:188]                             # 96 x 128 BGR crop of the real image
        out["image_bgr"] = img
        out["image_s0"] = reference_clip(img[None], (24, 33), 0, 1, 1, False)
        out["image_half"] = reference_clip(img[None], (48, 64), 0, 1, 1, False)   # exact 2x decimation (box mean)
    else:                                                                  # keep the previous real-image entries
        old = np.load(os.path.join(HERE, "frames_small.npz"))
        for k in ("image_bgr", "image_s0", "image_half"):
            out[k] = old[k]
    np.savez_compressed(os.path.join(HERE, "frames_small.npz"), **out)
    print("frames_small:", sorted(out))


if __name__ == "__main__":
    sample_small()
    ops_small()
    frames_small()
