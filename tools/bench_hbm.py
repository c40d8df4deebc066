#!/usr/bin/env python
"""Achieved HBM bandwidth of the bandwidth-bound kernels of the hot path on BATCHED synthetic inputs.

At batch 1 the resize / Adam / loss kernels move <= 10 MB and are launch-latency bound (SURVEY.md §8d caveat), so
their fraction of the HBM roofline is measured here on inputs of >= 0.5 GB, with CUDA events on the launching stream,
after warm-up.  "bytes" is the ALGORITHMIC traffic of one launch, the peak is
MEASURED_PEAKS.json:hbm_gbs (6545.6 GB/s copy bandwidth on this pool's B200) or the profiling recipe's fallback.

  python tools/bench_hbm.py [--json out.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
import hpvg  # noqa: E402
from hpvg import ops  # noqa: E402
from hpvg.runtime import BF16, F32, F64, Tensor  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


ITERS, WARM = [10], [3]


def timed(st, fn, iters=None, warm=None):
    iters, warm = iters or ITERS[0], WARM[0] if warm is None else warm
    for _ in range(warm):
        fn()
    st.sync()
    e0, e1 = hpvg.Event(), hpvg.Event()
    e0.record(st)
    for _ in range(iters):
        fn()
    e1.record(st)
    e1.sync()
    return e0.elapsed_ms(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--clips", type=int, default=48, help="3-channel clips per launch for the resize kernels")
    ap.add_argument("--wide", type=int, default=6, help="64-channel finest-scale clips per launch for the BN kernels")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3)
    args = ap.parse_args()
    hpvg.init(0)
    rows = measure(hpvg.Stream(), args)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


def measure(st, args=None, log=sys.stdout):
    """Times every bandwidth-bound kernel on `st`; returns one dict per kernel.  bench.py calls this at N == 1 so the
    HBM fractions sit in the same JSON line as the headline numbers (`hbm_kernels`)."""
    if args is None:
        args = argparse.Namespace(clips=48, wide=6, iters=10, warm=3)
    ITERS[0], WARM[0] = max(args.iters, 1), max(args.warm, 0)
    peak, kind = peak_gbs()
    rows = []

    def report(name, nbytes, ms, note=""):
        gbs = nbytes / ms / 1e6
        rows.append({"kernel": name, "bytes_per_launch": int(nbytes), "ms": ms, "achieved_gbs": gbs, "peak_gbs": peak,
                     "frac": gbs / peak, "peak_source": kind, "note": note})
        if log is not None:
            print("%-34s %9.1f MB  %8.3f ms  %8.1f GB/s  %5.1f%% of %s peak %s" %
                  (name, nbytes / 1e6, ms, gbs, 100 * gbs / peak, kind, note), file=log, flush=True)

    # ---------------------------------------------------------------- trilinear resize s8 -> s9 (images.py:54-61)
    N = args.clips
    si, so = (7, 153, 204), (13, 192, 257)
    vi, vo = int(np.prod(si)), int(np.prod(so))
    x = Tensor((N, 3) + si, F32).zero_(st)
    y = Tensor((N, 3) + so, F32)
    gx = Tensor((N, 3) + si, F32)
    xin = Tensor((N,) + so + (8,), BF16)
    report("resize3d_fwd s8->s9 x%d" % N, 4 * 3 * N * (vi + vo), timed(st, lambda: ops.resize3d(x, so, out=y, stream=st)))
    report("resize3d_bwd s9->s8 x%d" % N, 4 * 3 * N * (vi + vo),
           timed(st, lambda: ops.resize3d_bwd(y, si, out=gx, stream=st)))
    report("upsample_noise_pack (philox) x%d" % N, N * (4 * 3 * (vi + vo) + 16 * vo),
           timed(st, lambda: ops.upsample_noise_pack(x, so, amp=0.1, seed=1234, up=y, xin=xin, stream=st)),
           "resize + in-kernel N(0,1) + bf16 block input")
    del x, y, gx, xin

    # ---------------------------------------------------------------- frames -> clip (generate_frames.py:42-46, video.py:45-86)
    F_ = 64
    fr = Tensor((F_, 540, 960, 3), "uint8").zero_(st)
    clip = Tensor((1, 3, F_, 1080, 1920), F32)
    report("frames_to_clip 540p->1080p x%d" % F_, F_ * (540 * 960 * 3 + 1080 * 1920 * 12),
           timed(st, lambda: ops.frames_to_clip(fr, (1080, 1920), n_frames=F_, bgr=True, out=clip, stream=st)),
           "cv2-exact u8 bilinear + /255 + normalise, fp32 CTHW out")
    fr2 = Tensor((F_, 1080, 1920, 3), "uint8").zero_(st)
    report("frames_to_clip 1080p same size x%d" % F_, F_ * 1080 * 1920 * 15,
           timed(st, lambda: ops.frames_to_clip(fr2, (1080, 1920), n_frames=F_, bgr=True, out=clip, stream=st)))
    del fr, fr2, clip

    # ---------------------------------------------------------------- BatchNorm (train) on 64-channel bf16 cl
    B = args.wide
    vox = B * vo
    ycl = Tensor((B,) + so + (64,), BF16).zero_(st)
    acl = Tensor((B,) + so + (64,), BF16)
    gcl = Tensor((B,) + so + (64,), BF16).zero_(st)
    gamma = hpvg.from_numpy(np.ones(64, np.float32))
    beta = hpvg.from_numpy(np.zeros(64, np.float32))
    mm, mv = hpvg.from_numpy(np.zeros(64, np.float32)), hpvg.from_numpy(np.ones(64, np.float32))
    stats = Tensor((2, 64), F64).zero_(st)
    stats.copy_from_host(np.stack([np.zeros(64), np.full(64, float(vox))]), st)
    saved = Tensor((4, 64), F32)
    report("bn_train_apply_cl (1-pass BN+LReLU)", 4 * 64 * vox,
           timed(st, lambda: ops.bn_train_fused_cl(ycl, stats, gamma, beta, mm, mv, out=acl, saved=saved, stream=st)),
           "2 B in + 2 B out per element")
    s2 = Tensor((2, 64), F64)
    report("bn_stats_cl (standalone)", 2 * 64 * vox,
           timed(st, lambda: hpvg._lib.check(hpvg.lib.hpvg_bn_stats_cl(ycl.ptr, vox, s2.ptr, s2.ptr + 512, st.handle))))
    gy = Tensor((B,) + so + (64,), BF16)
    ms = timed(st, lambda: ops.bn_bwd_cl(gcl, ycl, saved, out=gy, stream=st))
    report("bn_bwd_cl (reduce + apply)", (4 + 6) * 64 * vox, ms, "pass 1 reads ga,y; pass 2 reads ga,y writes gy")
    report("lrelu_bwd_cl", 6 * 64 * vox, timed(st, lambda: ops.lrelu_bwd_cl(gcl, ycl, out=gy, stream=st)))
    f32 = Tensor((B, 64) + so, F32)
    report("unpack_cl bf16->f32 ncdhw", 6 * 64 * vox, timed(st, lambda: ops.unpack_cl(ycl, out=f32, stream=st)))
    report("pack_cl f32 ncdhw->bf16", 6 * 64 * vox, timed(st, lambda: ops.pack_cl(f32, out=ycl, stream=st)))
    del ycl, acl, gcl, gy, f32

    # ---------------------------------------------------------------- clip + Adam (optimizers.py:33-43)
    nt, per = 16, 4 * 1024 * 1024
    ps = [Tensor((per,), F32).zero_(st) for _ in range(nt)]
    gs = [Tensor((per,), F32).zero_(st) for _ in range(nt)]
    m1 = [Tensor((per,), F32).zero_(st) for _ in range(nt)]
    v1 = [Tensor((per,), F32).zero_(st) for _ in range(nt)]
    step = [0]

    def adam(clip):
        step[0] += 1
        ops.adam_clip_multi(ps, gs, m1, v1, [5e-4] * nt, step[0], clip=clip, stream=st)

    report("adam (no clip) 16 x 4M params", 28 * nt * per, timed(st, lambda: adam(0.0)), "28 B / parameter")
    report("clip-by-norm + adam 16 x 4M params", 32 * nt * per, timed(st, lambda: adam(5.0)), "+4 B / parameter norm pass")
    del ps, gs, m1, v1

    # ---------------------------------------------------------------- losses (losses.py:5-7, nn.MSELoss)
    n = 160 * 1024 * 1024
    a, b = Tensor((n,), F32).zero_(st), Tensor((n,), F32).zero_(st)
    out = Tensor((1,), F32)
    report("mse (160M elements)", 8 * n, timed(st, lambda: ops.mse(a, b, out=out, stream=st)))
    report("kl_criterion (160M elements)", 8 * n, timed(st, lambda: ops.kl_criterion(a, b, out=out, stream=st)))
    report("mse_grad", 12 * n, timed(st, lambda: ops.mse_grad(a, b, 0.5, g=a, stream=st)))
    return rows


if __name__ == "__main__":
    main()
