#!/usr/bin/env python
"""Small-batch sampling: per-batch time of the generator's random-mode forward (full 10-scale pyramid, 13x192x257)
through the per-layer Python path, the one-call C entry (hpvg_generator_sample) and the same entry replayed as a CUDA
graph.  The reference's eval loop generates one sample at a time (eval_video.py:62-76).
usage: sample_latency.py [batches=1,2,4] [iters=100]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
import hpvg
from hpvg import networks_3d as n3, sampling
from hpvg.utils import images as uimg

batches = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,2,4").split(",")]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hpvg.init(0)
opt = uimg.default_opt()
G = n3.GeneratorHPVAEGAN(opt, seed=0)
for _ in range(opt.stop_scale):
    G.init_next_stage()
G.set_train(False)
amps = [1.0] + [0.1] * opt.stop_scale
st = hpvg.Stream()
rng = np.random.default_rng(0)
for B in batches:
    z = hpvg.from_numpy(rng.standard_normal(sampling.z_init_size(opt, B)).astype(np.float32))
    direct = sampling.FusedSampler(G, amps, B, stream=st)
    graphed = sampling.FusedSampler(G, amps, B, stream=st, graph=True)
    out = hpvg.Tensor((B,) + direct.out_shape, hpvg.F32)
    paths = {"per-layer calls from Python": lambda i: G(z, amps, noise_init=z, isRandom=True, stream=st),
             "one C call (hpvg_generator_sample)": lambda i: direct(z, sample_base=i * B, out=out, stream=st),
             "the same, replayed as a CUDA graph": lambda i: graphed(z, sample_base=i * B, stream=st)}
    for name, fn in paths.items():
        for i in range(5):
            fn(i)
        st.sync()
        e0, e1 = hpvg.Event(), hpvg.Event()
        e0.record(st)
        for i in range(iters):
            fn(i)
        e1.record(st)
        e1.sync()
        ms = e0.elapsed_ms(e1) / iters
        print("batch %d  %-38s %7.3f ms per batch  %7.1f clips/s" % (B, name, ms, B / ms * 1e3), flush=True)
    graphed.close()
