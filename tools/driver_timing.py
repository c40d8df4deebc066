#!/usr/bin/env python
"""Per-iteration time of the reference-shaped progressive trainer (hpvg.driver.train_pyramid, train_video.py:413-419)
with every scale's iteration replayed as a CUDA graph, against the same iterations launched from Python.
usage: driver_timing.py [stop_scale=6] [niter=40]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
import hpvg
from hpvg import driver, networks_3d as n3
from hpvg.utils import images as uimg

stop = int(sys.argv[1]) if len(sys.argv) > 1 else 6
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hpvg.init(0)
for graph in (True, False):
    opt = uimg.default_opt()
    rng = np.random.default_rng(0)
    reals = {s: np.tanh(rng.standard_normal((1, 3) + uimg.scale_shape(opt, s))).astype(np.float32)
             for s in range(stop + 1)}
    G = n3.GeneratorHPVAEGAN(opt, seed=0)
    stamps = {}

    def on_iter(s, it, losses):
        hpvg.device_sync()
        stamps.setdefault(s, []).append(time.perf_counter())

    driver.train_pyramid(opt, G, n3.WDiscriminator3D, lambda s: reals[s], niter, stop_scale=stop, on_iter=on_iter,
                         graph=graph)
    out = []
    for s in sorted(stamps):
        t = np.diff(stamps[s][niter // 2:])          # second half: after the eager warm-up / capture iterations
        out.append("s%d %.2f" % (s, 1e3 * float(np.median(t))))
    print(("graph " if graph else "eager ") + "ms per iteration (median, incl. one device sync per iteration): " + "  ".join(out), flush=True)
