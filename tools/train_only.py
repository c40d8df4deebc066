#!/usr/bin/env python
"""Runs only the GAN-phase train iterations of bench.py (development tool: ncu launch lists / timing)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
sys.path.insert(0, ROOT)
import bench
import hpvg
from hpvg.utils import images as uimg
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hpvg.init(0)
st = hpvg.Stream()
opt = uimg.default_opt()
graph = not (len(sys.argv) > 3 and sys.argv[3] == 'eager')
frames = int(sys.argv[4]) if len(sys.argv) > 4 else None
peaks, kind = bench.load_peaks()
print(json.dumps(bench.train_iter_bench(hpvg, opt, steps, warm, st, peaks, kind, graph=graph, frames=frames)))
