#!/usr/bin/env python
"""Config 5 on N GPUs of one box: shard `--samples` clips over the ranks, reduce each clip to its
sinFID moments on the device, ncclAllGather them over NVLink, Fréchet distance on rank 0 — and CHECK that the gathered
rows equal what rank 0 gets generating every sample alone (shard invariance through the real NCCL path).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/fid_multi_gpu.py --samples 32
(torch.distributed.run is only the launcher: the collective is issued by hpvg.dist through ctypes.)"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
import hpvg  # noqa: E402
from hpvg import dist, fid, networks_3d as n3, sampling  # noqa: E402
from hpvg.utils import images as uimg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=32)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--img-size", type=int, default=128)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
hpvg.init(int(os.environ.get("LOCAL_RANK", "0")))
st = hpvg.Stream()
comm = dist.from_env()
opt = uimg.default_opt(img_size=args.img_size)
net = n3.GeneratorHPVAEGAN(opt, seed=0)
for _ in range(opt.stop_scale):
    net.init_next_stage()
amps = [1.0] + [0.1] * opt.stop_scale
feats = fid.RandomFeatures3D(3, seed=5)
real = np.tanh(np.random.default_rng(1).standard_normal((1, 3) + uimg.scale_shape(opt, opt.stop_scale))).astype(np.float32)
t0 = time.perf_counter()
rows, count, _ = sampling.generate_moments(net, amps, args.samples, feats, comm, batch=args.batch, seed=3, stream=st)
dt = time.perf_counter() - t0
real_rows = fid.sample_moments(feats(hpvg.from_numpy(real), stream=st), stream=st).numpy(st)
value, _ = fid.svfid_from_moments(real_rows[0], rows, count)
ok = None
if rank == 0:
    alone, _, _ = sampling.generate_moments(net, amps, args.samples, feats, None, batch=args.batch, seed=3, stream=st)
    ok = bool(np.array_equal(alone, rows))
    print(json.dumps({"world": world, "samples": args.samples, "svfid": value, "gather_bytes": int(rows.nbytes),
                      "rows_equal_single_process": ok, "seconds": dt}), flush=True)
comm.barrier()
if hasattr(comm, "close"):
    comm.close()
sys.exit(0 if ok in (None, True) else 1)
