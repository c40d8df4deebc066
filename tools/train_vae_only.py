#!/usr/bin/env python
"""Runs only the VAE-phase train iterations of bench.py (BASELINE config 2; development tool for ncu launch lists).
usage: train_vae_only.py [steps] [warmup] [eager|graph] [2d]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
sys.path.insert(0, ROOT)
import bench
import hpvg
from hpvg.utils import images as uimg
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 3
graph = not (len(sys.argv) > 3 and sys.argv[3] == "eager")
two_d = len(sys.argv) > 4 and sys.argv[4] == "2d"
hpvg.init(0)
st = hpvg.Stream()
peaks, kind = bench.load_peaks()
if two_d:
    r = bench.train_vae_bench(hpvg, uimg.default_opt(**bench.IMAGE_OPT), steps, warm, st, peaks, kind, bench.IMAGE_SCALE, graph, 2)
else:
    r = bench.train_vae_bench(hpvg, uimg.default_opt(), steps, warm, st, peaks, kind, 2, graph, 3)
print(json.dumps(r))
