#!/usr/bin/env python
"""Per-kernel-variant time breakdown of one sampling step (development tool)."""
import os, sys, time, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
import hpvg
from hpvg import networks_3d as n3, ops, sampling
from hpvg.utils import images as uimg
hpvg.init(0)
st = hpvg.Stream()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
opt = uimg.default_opt()
net = n3.GeneratorHPVAEGAN(opt)
for _ in range(opt.stop_scale): net.init_next_stage()
amps = [1.0] + [0.1] * opt.stop_scale
z = hpvg.from_numpy(np.random.default_rng(0).standard_normal(sampling.z_init_size(opt, B)).astype(np.float32))
def step():
    net.sample_counter = 0
    return net(z, amps, noise_init=z, isRandom=True, stream=st)
for _ in range(3): step()
st.sync()
t0 = time.perf_counter(); e0, e1 = hpvg.Event(), hpvg.Event(); e0.record(st)
for _ in range(5): step()
t_issue = time.perf_counter() - t0
e1.record(st); e1.sync()
print("batch %d: %.2f ms/step on device (events), host issue time %.2f ms/step" % (B, e0.elapsed_ms(e1) / 5, 1000 * t_issue / 5))
ops.start_profile(None)
step(); st.sync()
items = ops.stop_profile()
agg = collections.defaultdict(float)
for (mode, vox), ms in items: agg[mode] += ms
tot = sum(agg.values())
print("conv kernels in one step: total %.2f ms: " % tot + ", ".join("mode %d: %.2f ms" % (m, v) for m, v in sorted(agg.items())))
per = collections.defaultdict(lambda: [0, 0.0])
for (mode, vox), ms in items:
    per[(mode, vox)][0] += 1; per[(mode, vox)][1] += ms
print("mode  voxels/launch  launches   ms_total   ns/voxel   TFLOP/s(64->64)")
for (mode, vox), (n, ms) in sorted(per.items()):
    print("%4d %12d %8d %10.3f %10.3f %10.1f" % (mode, vox, n, ms, ms * 1e6 / (vox * n), (2*27*64*64*vox*n/ms/1e9) if mode == 0 else 0))
print("pool stats", hpvg.runtime._POOL_STATS)
