#!/usr/bin/env python
"""Digest gpurun_out/prof_<tag>/ (made by tools/make_profiles.sh on the GPU box) into the tracked profiles/ directory:
launch-list shares, a one-line-per-kernel table of the `ncu --set full` captures, the HBM table and conv64_traffic.json."""
import csv, json, os, subprocess, sys, collections, re, shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
src = os.path.join(ROOT, "gpurun_out", "prof_" + tag)
dst = os.environ.get("HPVG_PROF_DST") or os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)


def launch_summary(path, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        ms = v / 1e6 if row[ui] in ("ns", "nsecond") else v / 1e3 if row[ui] in ("us", "usecond") else v
        name = re.sub(r"\(.*", "", row[ki])[:80]
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    out = ["%s" % title, "(ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)",
           "total %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values()))]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%6.2f%% %6d launches %10.3f ms  %s" % (100 * ms / tot, n, ms, k))
    return "\n".join(out) + "\n"


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def full_rows(rep):
    """One row per profiled launch with the FIXED column set KEYS (captures made with different options expose different
    raw columns; a metric a capture does not have stays empty)."""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    if len(r) < 3:
        return [], []
    h, units = r[0], r[1]
    idx = {k: h.index(k) for k in ["Kernel Name"] + KEYS if k in h}
    header = ["Kernel Name"] + [k + (" [%s]" % units[idx[k]] if k in idx and units[idx[k]] else "") for k in KEYS]
    rows = [[row[idx["Kernel Name"]]] + [row[idx[k]] if k in idx else "" for k in KEYS] for row in r[2:]]
    return header, rows


def main():
    for name, title in (("launches_sample.csv", "bench.py --steps 2 --warmup 3 --no-extras (sampling, 16 clips/step)"),
                        ("launches_train.csv", "tools/train_only.py 1 1 eager 16 (GAN-phase train iterations at the 16-frame finest scale)"),
                        ("launches_train_vae.csv", "tools/train_vae_only.py 1 0 eager (VAE-phase train iterations at scale 2, BASELINE config 2)")):
        p = os.path.join(src, name)
        if os.path.exists(p):
            open(os.path.join(dst, "%s_%s_summary.txt" % (tag, name[:-4])), "w").write(launch_summary(p, title))
            # keep the raw list too (small)
            shutil.copy(p, os.path.join(dst, "%s_%s" % (tag, name)))
    table = []
    hdr = None
    for rep in sorted(f for f in os.listdir(src) if f.endswith(".ncu-rep")):
        h, rows = full_rows(os.path.join(src, rep))
        if not rows:
            continue
        if rep.startswith("hbm_kernels"):     # section-based capture: its own columns -> its own file
            seen, sub = set(), []
            for row in rows:          # one representative launch per kernel
                key = re.sub(r"\(.*", "", row[0])
                if key in seen:
                    continue
                seen.add(key)
                sub.append(row)
            with open(os.path.join(dst, "%s_ncu_hbm_kernels.csv" % tag), "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(h)
                w.writerows(sub)
            continue
        if hdr is None:
            hdr = h
        table.append([rep] + rows[0])
        if rep == "conv64.ncu-rep":
            d = dict(zip(h, rows[0]))
            rd = float([v for k, v in d.items() if k.startswith("dram__bytes_read.sum")][0])
            wr = float([v for k, v in d.items() if k.startswith("dram__bytes_write.sum")][0])
            unit = [k for k in d if k.startswith("dram__bytes_read.sum")][0]
            mul = 1e6 if "Mbyte" in unit else 1e9 if "Gbyte" in unit else 1e3 if "Kbyte" in unit else 1.0
            vox = 8 * 13 * 192 * 257
            json.dump({"dram_bytes_per_voxel": (rd + wr) * mul / vox,
                       "source": "ncu --set full --clock-control none, conv3d_umma_kernel<64->64>, 8 x 13x192x257 voxels "
                                 "per launch: dram__bytes_read.sum %.1f + dram__bytes_write.sum %.1f (%s)" % (rd, wr, unit),
                       "algorithmic_bytes_per_voxel": 256}, open(os.path.join(dst, "conv64_traffic.json"), "w"))
    if table:
        with open(os.path.join(dst, "%s_ncu_full_summary.csv" % tag), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["capture"] + hdr)
            w.writerows(table)
    for name in ("hbm.json", "hbm.log", "perf_conv.log"):
        p = os.path.join(src, name)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(dst, "%s_%s" % (tag, name)))
    print("profiles/ updated from", src)


if __name__ == "__main__":
    main()
