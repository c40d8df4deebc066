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


if __name__ == "__main__":
    sample_small()
