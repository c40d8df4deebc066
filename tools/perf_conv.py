#!/usr/bin/env python
"""Time the tcgen05 conv kernel variants with CUDA events on their own stream (development tool)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
import hpvg  # noqa: E402
from hpvg import ops  # noqa: E402

hpvg.init(0)
st = hpvg.Stream()
rng = np.random.default_rng(0)


def bench(mode_name, cin, cout, shape, iters=10, stats=False):
    """Precision follows HPVG_PRECISION (bf16: one kind::f16 launch; tf32: the two 32-channel kind::tf32 launches)."""
    N, T, H, W = shape
    dt = ops.cl_dtype()
    cpad = ops.narrow_pitch() if cin <= 8 else cin
    x_cl = hpvg.Tensor((N, T, H, W, cpad), dt).zero_()
    w = hpvg.from_numpy((rng.standard_normal((cout, cin, 3, 3, 3)) * 0.05).astype(np.float32))
    aff = ops.affine_from_bias(hpvg.from_numpy(np.zeros(max(cout, 1), np.float32)))
    wimgs = ops.build_wimgs(w, cin, cout)
    mode_name += "" if dt == hpvg.BF16 else "[tf32]"
    if cout <= 4:
        out = hpvg.Tensor((N, cout, T, H, W), hpvg.F32)
        run = lambda: ops.conv3d_cl_any(x_cl, w, aff, ops.ACT_TANH, cin, cout, out=out, wimgs=wimgs, stream=st)
    else:
        out = hpvg.Tensor((N, T, H, W, 64), dt)
        stt = hpvg.Tensor((2, 64), hpvg.F64).zero_() if stats else None
        run = lambda: ops.conv3d_cl_any(x_cl, w, aff, ops.ACT_LRELU, cin, cout, out=out, wimgs=wimgs, stats=stt, stream=st)
    for _ in range(3):
        run()
    st.sync()
    e0, e1 = hpvg.Event(), hpvg.Event()
    e0.record(st)
    for _ in range(iters):
        run()
    e1.record(st)
    e1.sync()
    ms = e0.elapsed_ms(e1) / iters
    vox = N * T * H * W
    flops = 2.0 * 27 * cin * cout * vox
    print("%-16s %-22s %8.3f ms  %8.1f TFLOP/s (algorithmic)  %6.2f ns/voxel" %
          (mode_name + ("+bnstats" if stats else ""), shape, ms, flops / ms / 1e9, ms * 1e6 / vox), flush=True)


if len(sys.argv) > 1:   # e.g. `perf_conv.py 64 64 8,13,192,257 [stats]` : one case only (ncu captures)
    cin, cout = int(sys.argv[1]), int(sys.argv[2])
    shape = tuple(int(v) for v in sys.argv[3].split(","))
    bench("%d->%d" % (cin, cout), cin, cout, shape, iters=3, stats=len(sys.argv) > 4)
    sys.exit(0)
for shape in [(1, 13, 192, 257), (1, 16, 192, 257), (4, 13, 192, 257), (8, 13, 192, 257), (1, 7, 153, 204),
              (16, 4, 24, 33), (64, 4, 24, 33)]:
    bench("64->64", 64, 64, shape)
for shape in [(1, 13, 192, 257), (8, 13, 192, 257)]:
    bench("64->64", 64, 64, shape, stats=True)
for shape in [(1, 13, 192, 257), (8, 13, 192, 257)]:
    bench("64->3", 64, 3, shape)
    bench("3->64", 3, 64, shape)
