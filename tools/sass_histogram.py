#!/usr/bin/env python
"""SASS evidence without a GPU: per kernel of the built objects, the count of the Blackwell-specific opcodes
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
UBLKCP = bulk copy, SYNCS = mbarrier, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents of the programmatic
dependent launch) plus the classic ones a recompiled sm_90 kernel would show instead (HMMA, LDGSTS).
usage: tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "UTCATOMSWS",
         "ACQBULK", "PREEXIT",
         "HMMA", "IMMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR", "MUFU", "FFMA", "DFMA"]

for obj in sorted(glob.glob(os.path.join(ROOT, "mindspore-hp-vae-gan_b200", "build", "*.o"))):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    print("==== %s" % os.path.basename(obj))
    name, counts, total = None, None, 0

    def flush():
        if name is None:
            return
        keep = {k: v for k, v in counts.items() if v}
        short = re.sub(r"^_ZN4hpvg\d+_GLOBAL__N__[0-9a-f_]+", "", name)
        print("  %-74s %6d instr  %s" % (short[:74], total, "  ".join("%s=%d" % (k, keep[k]) for k in WATCH if k in keep)))

    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            flush()
            name, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name is not None:
            op = m.group(1)
            total += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[w] += 1
                    break
    flush()
