#!/usr/bin/env python
"""bench.py — benchmarks of the B200-native HP-VAE-GAN hot path (contract: task statement; configs: BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                      the CPU restatement of the reference on the host cores

Headline (default, `--workload sample`, BASELINE config 4): sampled clips/s — one clip = one random-mode generator
forward through the full 10-scale pyramid (13 x 192 x 257 at the finest scale), weak-scaled over N GPUs, no collective
on the data path.  ONE JSON line on rank 0:
  value      device-resident throughput (noise already in HBM), CUDA events, max over ranks
  e2e        the same through the public API `hpvg.sampling.SamplePipeline` (what `sampling.generate` runs): per-sample
             random numbers of z drawn ON THE HOST (numpy, worker threads) into pinned memory, H2D, Box-Muller +
             generation on the device, D2H of every clip
  roofline   the dominant kernel (tcgen05 conv 64->64), per-launch CUDA events live in this run
  extras (objects in the same line): `tf32` (the fp32-accurate precision mode), `train` (config 3), `train_vae`
  (config 2), `train_image` (config 1, 2-D), `fid` (config 5: moments + NCCL all-gather, every N), `hbm_kernels`,
  `cpu_baseline`.  Each train object carries its own roofline (FLOP counted from the launches of one iteration),
  cpu_baseline (the oracle's iteration on the host cores) and e2e."""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "mindspore-hp-vae-gan_b200"))
sys.path.insert(0, ROOT)

METRIC = "sampled clips/s"
UNIT = "clips/s"
FLOP_PER_VOXEL_64 = 2 * 27 * 64 * 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="clips per step per GPU")
    ap.add_argument("--img-size", type=int, default=256)
    ap.add_argument("--workload", default="sample", choices=["sample", "train", "train_vae", "train_image", "fid"],
                    help="which measurement becomes the top-level metric/value of the line (default: sampled clips/s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=2, help="clips in the bounded CPU-baseline sample")
    ap.add_argument("--train-steps", type=int, default=10, help="train iterations timed for the `train` object")
    ap.add_argument("--train-frames", type=int, default=16,
                    help="frames at the finest scale for the train workload (BASELINE.json config 3: 16; the "
                         "reference's own schedule gives 13)")
    ap.add_argument("--no-extras", action="store_true", help="only the headline measurement")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-hbm", action="store_true")
    ap.add_argument("--no-fid", action="store_true")
    ap.add_argument("--no-tf32", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="train iterations launched eagerly instead of as a CUDA graph")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def committed_traffic():
    """dram__bytes_read+write per voxel of the 64->64 conv from the COMMITTED `ncu --set full` capture (not this run)."""
    p = os.path.join(ROOT, "profiles", "conv64_traffic.json")
    try:
        d = json.load(open(p))
        return float(d["dram_bytes_per_voxel"]), d.get("source", "profiles/conv64_traffic.json")
    except Exception:
        return None, None


def sample_workload_name(stop_scale, shape, batch):
    return ("eval_video.py random-noise sampling, full %d-scale pyramid, finest scale %dx%dx%d, %d clips/step/GPU, "
            "random-init weights" % ((stop_scale + 1,) + tuple(shape) + (batch,)))


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ================================================================================================= CPU oracle arms
# Everything in this section imports ONLY the oracle and torch-CPU (never hpvg): it is the reference arm
# (`--impl reference`) and the `cpu_baseline` legs.
def _cpu_threads(threads=None):
    import torch
    # all host cores: torchrun exports OMP_NUM_THREADS=1 to its workers, which would throttle the CPU arm
    torch.set_num_threads(threads or os.cpu_count() or 1)
    return torch.get_num_threads()


def cpu_sample_clips(opt_kw, n_clips, threads=None):
    """The CPU oracle (torch-CPU fp32, all host cores) generating n_clips full-pyramid samples -> (clips/s, threads)."""
    import torch
    from oracle import hpvg_oracle as orc
    cores = _cpu_threads(threads)
    opt = orc.default_opt(**opt_kw)
    p = orc.to_torch(orc.init_generator_params(opt, opt.stop_scale, seed=0))
    rng = np.random.default_rng(0)
    amps = [1.0] + [0.1] * opt.stop_scale

    def one():
        z = torch.from_numpy(rng.standard_normal((1, opt.latent_dim) + orc.scale_shape(opt, 0)).astype(np.float32))
        noises = {s: torch.from_numpy(rng.standard_normal((1, 3) + orc.scale_shape(opt, s)).astype(np.float32))
                  for s in range(opt.vae_levels, opt.stop_scale + 1)}
        with torch.no_grad():
            orc.generator_forward(None, amps, p, opt, noise_init=z, is_random=True, noises=noises)

    t0 = time.perf_counter()
    for _ in range(n_clips):
        one()
    return n_clips / (time.perf_counter() - t0), cores


def _cpu_adam(tensors, lr, clip):
    """optimizers.py:41-43 on the host: per-tensor ClipByNorm (clip > 0) + Adam, first step."""
    from oracle import hpvg_oracle as orc
    for t in tensors:
        if t.grad is None:
            continue
        g = t.grad.numpy()
        if clip:
            g = orc.clip_by_norm(g, clip)
        w, _, _ = orc.adam_step(t.detach().numpy(), g, np.zeros_like(g), np.zeros_like(g), 1, lr)
        t.data.copy_(__import__("torch").from_numpy(w))
        t.grad = None


def cpu_train_gan_iter(opt_kw, frames, reps=1, threads=None):
    """One GAN-phase train iteration at the finest scale on the CPU oracle (train_video.py:170-177): D step (fake clip
    through the whole pyramid, 3 D passes, WGAN-GP double backward, Adam) + G step (reconstruction forward in
    BatchNorm-train mode, backward of the last stage, random forward + D, ClippedAdam) -> (iter/s, threads)."""
    import torch
    from oracle import hpvg_oracle as orc
    cores = _cpu_threads(threads)
    opt = orc.default_opt(**opt_kw)
    if frames:
        opt.td_override = {opt.stop_scale: int(frames)}
    S = opt.stop_scale
    tg = orc.to_torch(orc.init_generator_params(opt, S, seed=0), requires_grad=("body.%d." % (S - 1),))
    td = orc.to_torch(orc.init_discriminator_params(opt, seed=1), requires_grad=("head.", "body.", "tail."))
    rng = np.random.default_rng(0)
    top, s0 = orc.scale_shape(opt, S), orc.scale_shape(opt, 0)
    real = torch.from_numpy(np.tanh(rng.standard_normal((1, 3) + top)).astype(np.float32))
    rz = torch.from_numpy(np.tanh(rng.standard_normal((1, 3) + s0)).astype(np.float32))
    amps = [1.0] + [0.1] * S

    def noises():
        return {s: torch.from_numpy(rng.standard_normal((1, 3) + orc.scale_shape(opt, s)).astype(np.float32))
                for s in range(opt.vae_levels, S + 1)}

    t0 = time.perf_counter()
    for _ in range(reps):
        nz = torch.from_numpy(rng.standard_normal((1, opt.latent_dim) + s0).astype(np.float32))
        with torch.no_grad():
            fake, _ = orc.generator_forward(None, amps, tg, opt, noise_init=nz, is_random=True, training=True,
                                            noises=noises())
        orc.d_loss(real, fake, float(rng.random()), td, opt).backward()
        _cpu_adam([t for t in td.values() if t.requires_grad], opt.lr_d, 0.0)
        orc.g_loss(real, rz, nz, amps, tg, {k: v.detach() for k, v in td.items()}, opt, False, z_pred=nz,
                   noises=noises()).backward()
        _cpu_adam([t for t in tg.values() if t.requires_grad], opt.lr_g, opt.grad_clip)
    return reps / (time.perf_counter() - t0), cores, top


def cpu_train_vae_iter(opt_kw, scale_idx, nd=3, reps=3, threads=None):
    """One VAE-phase train iteration on the CPU oracle (train_video.py:170-172 / train_image.py:151-153): G step with
    encode + decoder + body[-1] trainable, MSE + MSE + KL, per-tensor clip + Adam -> (iter/s, threads, shape)."""
    import torch
    from oracle import hpvg_oracle as orc
    cores = _cpu_threads(threads)
    opt = orc.default_opt(**opt_kw)
    shp = (lambda i: orc.scale_shape(opt, i)) if nd == 3 else (lambda i: orc.scale_shape_2d(opt, i))
    req = ("encode.", "decoder.", "body.%d." % (scale_idx - 1)) if scale_idx else ("encode.", "decoder.")
    tg = orc.to_torch(orc.init_generator_params(opt, scale_idx, seed=0, nd=nd), requires_grad=req)
    rng = np.random.default_rng(0)
    real = torch.from_numpy(np.tanh(rng.standard_normal((1, 3) + shp(scale_idx))).astype(np.float32))
    rz = torch.from_numpy(np.tanh(rng.standard_normal((1, 3) + shp(0))).astype(np.float32))
    amps = [1.0] + [0.1] * scale_idx
    t0 = time.perf_counter()
    for _ in range(reps):
        z = torch.from_numpy(rng.standard_normal((1, opt.latent_dim) + shp(0)).astype(np.float32))
        orc.g_loss(real, rz, None, amps, tg, None, opt, True, z_pred=z, nd=nd).backward()
        _cpu_adam([t for t in tg.values() if t.requires_grad], opt.lr_g, opt.grad_clip)
    return reps / (time.perf_counter() - t0), cores, shp(scale_idx)


IMAGE_OPT = {"img_size": 64}        # BASELINE config 1: "small pyramid" of train_image.py
IMAGE_SCALE = 2


def run_reference(args):
    """The reference arm: the CPU restatement of the reference (oracle/hpvg_oracle.py: numpy + torch-CPU/oneDNN fp32, all
    host cores) on the same workload.  MindSpore itself cannot be installed offline.  No hpvg import: the
    CUDA library is never mapped into this process."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import hpvg_oracle as orc
    opt_kw = {"img_size": args.img_size}
    opt = orc.default_opt(**opt_kw)
    W = max(0, args.warmup)
    if args.workload == "sample" or args.workload == "fid":
        metric, unit = METRIC, UNIT
        wl = sample_workload_name(opt.stop_scale, orc.scale_shape(opt, opt.stop_scale), 1) + \
            " (each step = 1 clip on the host CPU)"
        step = lambda: cpu_sample_clips(opt_kw, 1)      # noqa: E731
        sample = "%d steps x 1 clip"
    elif args.workload == "train":
        metric, unit = "video train iter/s", "iter/s"
        wl = "train_video.py GAN-phase iteration at the finest scale, %d frames (BASELINE config 3)" % args.train_frames
        step = lambda: cpu_train_gan_iter(opt_kw, args.train_frames, 1)[:2]      # noqa: E731
        sample = "%d steps x 1 iteration"
    elif args.workload == "train_vae":
        metric, unit = "video train iter/s", "iter/s"
        wl = "train_video.py VAE-phase iteration at scale 2 (BASELINE config 2)"
        step = lambda: cpu_train_vae_iter(opt_kw, 2, 3, 1)[:2]      # noqa: E731
        sample = "%d steps x 1 iteration"
    else:
        metric, unit = "image train iter/s", "iter/s"
        wl = "train_image.py VAE-phase iteration, img_size 64, scale %d (BASELINE config 1)" % IMAGE_SCALE
        step = lambda: cpu_train_vae_iter(IMAGE_OPT, IMAGE_SCALE, 2, 1)[:2]      # noqa: E731
        sample = "%d steps x 1 iteration"
    for _ in range(W):
        step()
    cores = os.cpu_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, cores = step()
    total = time.perf_counter() - t0
    value = args.steps / total
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": W, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": (sample % args.steps) + ", CPU restatement of the reference (torch-CPU/oneDNN fp32); "
                                   "MindSpore itself is not installable offline"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ================================================================================================= GPU arm: train
def _conv_roofline(prof, ms_total, peaks, peaks_kind, tf32=False):
    """prof: [(kind, mode, voxels, ms)] from ops.start_profile("all") -> roofline of the dominant kernel (the 64->64
    tcgen05 conv: forward and data-gradient launches) plus the weight-gradient kernel's line."""
    from hpvg import ops
    mode64 = ops.CONV_T32_64 if tf32 else ops.CONV_64_64
    per_launch = FLOP_PER_VOXEL_64 // (2 if tf32 else 1)        # a tf32 launch contracts 32 of the 64 input channels
    conv = [(v, ms) for k, m, v, ms in prof if k == "conv" and m == mode64]
    wg = [(v, ms) for k, m, v, ms in prof if k == "wgrad"]
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))) / (2.0 if tf32 else 1.0)
    out = {}
    for name, items, f in (("conv", conv, per_launch), ("wgrad", wg, FLOP_PER_VOXEL_64)):
        kms = sum(ms for _, ms in items)
        if not items or kms <= 0:
            continue
        ach = sum(v for v, _ in items) * f / (kms / 1e3) / 1e12
        out[name] = {"achieved": ach, "frac": ach / peak, "launches_timed": len(items),
                     "avg_launch_ms": kms / len(items), "share_of_step": kms / ms_total if ms_total else None}
    r = {"bound": "tensor", "kernel": "conv3d_umma_kernel<%s> fwd + dgrad (tcgen05 cta_group::2 implicit GEMM)" %
         ("32->64 kind::tf32" if tf32 else "64->64 kind::f16"), "peak": peak, "unit": "TFLOP/s",
         "peak_source": "%s bf16_tflops_sustained%s" % (peaks_kind, " / 2 (tf32 runs at half the bf16 rate)" if tf32 else "")}
    r.update(out.get("conv", {"achieved": 0.0, "frac": 0.0}))
    if "wgrad" in out:
        w = out["wgrad"]
        w["kernel"] = "conv3d_wgrad_kernel (wide 64x64 launches counted at 2*27*64*64 FLOP/voxel; narrow ones included)"
        r["wgrad"] = w
    return r


def _time_iters(hpvg, st, one_iter, steps, warmup):
    for _ in range(warmup):
        one_iter()
    st.sync()
    e0, e1 = hpvg.Event(), hpvg.Event()
    e0.record(st)
    last = None
    for _ in range(steps):
        last = one_iter()
    e1.record(st)
    e1.sync()
    return e0.elapsed_ms(e1) / steps, last


def train_iter_bench(hpvg, opt, steps, warmup, st, peaks, peaks_kind, graph=True, frames=None):
    """BASELINE config 3: one GAN-phase train iteration at the finest scale (train_video.py:170-177): D step (3 D
    forwards, 2 backwards, WGAN-GP double backward, Adam) + G step (reconstruction forward of the whole pyramid in
    BatchNorm-train mode, backward of the last stage, random forward + D forward for the loss value, ClippedAdam).
    Inputs come from pinned host buffers every iteration (the data loader + host noise of the reference); losses are
    read back: the timed number IS the end-to-end one."""
    from hpvg import networks_3d as n3, ops, sampling, train as T
    from hpvg.utils import images as uimg
    if frames:
        opt.td_override = {opt.stop_scale: int(frames)}
    G = n3.GeneratorHPVAEGAN(opt, seed=0)
    for _ in range(opt.stop_scale):
        G.init_next_stage()
    D = n3.WDiscriminator3D(opt, rng=np.random.default_rng(1))
    amps = [1.0] + [0.1] * opt.stop_scale
    top = uimg.scale_shape(opt, opt.stop_scale)
    rng = np.random.default_rng(0)
    shapes = {"real": (1, 3) + top, "real_zero": (1, 3) + uimg.scale_shape(opt, 0), "noise": sampling.z_init_size(opt, 1)}
    host = {k: hpvg.PinnedBuffer(int(np.prod(v)) * 4) for k, v in shapes.items()}
    for k, v in shapes.items():
        a = rng.standard_normal(v).astype(np.float32)
        host[k].as_array(v)[...] = np.tanh(a) if k != "noise" else a
    dev = {k: hpvg.Tensor(v, hpvg.F32) for k, v in shapes.items()}
    block = G.body[-1]

    def make_steps(device_side):
        optG = T.ClippedAdam(opt, [{"params": T.trainable_params(block), "lr": opt.lr_g}], opt.lr_g, beta1=opt.beta1,
                             beta2=0.999, device_step=device_side)
        optD = T.Adam(T.trainable_params(D), opt.lr_d, beta1=opt.beta1, beta2=0.999, device_step=device_side)
        g = T.TrainOneStepCell(T.GWithLoss(opt, D, G, device_rng=device_side), optG, cells_to_invalidate=[block])
        d = T.TrainOneStepCell(T.DWithLoss(opt, D, G, device_rng=device_side), optD, cells_to_invalidate=[D])
        g.set_train()
        d.set_train()
        return g, d

    nb = len(G.body)
    # second pinned noise buffer: the host draws iteration i+1's noise_init (numpy, like images.py:17-21) while the GPU
    # runs iteration i; an event per buffer guards its reuse
    noise_bufs = [host["noise"], hpvg.PinnedBuffer(host["noise"].nbytes)]
    noise_evs = [None, None]
    counter = [0]

    def upload():
        k = counter[0] & 1
        counter[0] += 1
        if noise_evs[k] is not None:
            noise_evs[k].sync()
        noise_bufs[k].as_array(shapes["noise"])[...] = np.random.standard_normal(shapes["noise"]).astype(np.float32)
        hpvg.lib.hpvg_h2d(dev["real"].ptr, host["real"].ptr, dev["real"].nbytes, st.handle)
        hpvg.lib.hpvg_h2d(dev["real_zero"].ptr, host["real_zero"].ptr, dev["real_zero"].nbytes, st.handle)
        hpvg.lib.hpvg_h2d(dev["noise"].ptr, noise_bufs[k].ptr, dev["noise"].nbytes, st.handle)
        ev = hpvg.Event()
        ev.record(st)
        noise_evs[k] = ev

    # (1) one EAGER iteration with every conv / weight-gradient launch bracketed by CUDA events and its algorithmic FLOP
    #     counted: the roofline leg (a graph replay cannot be instrumented per kernel)
    g_step, d_step = make_steps(graph)

    def eager_iter():
        upload()
        dl = d_step(dev["real"], dev["noise"], amps, stream=st)
        gl = g_step(dev["real"], dev["real_zero"], dev["noise"], amps, isVAE=False, trainable_body=(nb - 1,), stream=st)
        return dl, gl

    for _ in range(2):
        eager_iter()
    st.sync()
    ops.start_flop_count()
    ops.start_profile("all")
    e0, e1 = hpvg.Event(), hpvg.Event()
    e0.record(st)
    eager_iter()
    e1.record(st)
    e1.sync()
    prof = ops.stop_profile()
    fl = ops.stop_flop_count()
    eager_ms = e0.elapsed_ms(e1)
    flop_iter = fl["conv"] + fl["wgrad"]

    # (2) the timed iterations: one CUDA graph per iteration (or eager with --no-graph)
    if graph:
        it = T.GraphedIteration(st, g_step, d_step, dev["real"], dev["real_zero"], dev["noise"], amps,
                                dict(isVAE=False, trainable_body=(nb - 1,)))
        upload()
        it.warmup(max(warmup, 2))
        it.capture(concurrent=not os.environ.get("HPVG_NO_OVERLAP"))
        per_iter = it.kernels_per_launch

        def one_iter():
            upload()
            return it()
    else:
        one_iter, per_iter = eager_iter, None
    l0 = hpvg.lib.hpvg_launch_count()
    ms, (dl, gl) = _time_iters(hpvg, st, one_iter, steps, warmup)
    if per_iter is None:
        per_iter = (hpvg.lib.hpvg_launch_count() - l0) // (steps + warmup)
    roof = _conv_roofline(prof, eager_ms, peaks, peaks_kind)
    roof["iteration_tflops"] = flop_iter / (ms / 1e3) / 1e12
    roof["iteration_frac"] = roof["iteration_tflops"] / roof["peak"]
    roof["flop_per_iteration"] = flop_iter
    roof["flop_source"] = ("counted from the %d conv / dgrad and %d wgrad launches of one iteration: 2*27*Cin*Cout per "
                           "output voxel, real channel counts" % (fl["conv_calls"], fl["wgrad_calls"]))
    roof["note"] = "per-kernel lines from one event-instrumented EAGER iteration (%.1f ms); iteration_* from the timed " \
                   "graph replays" % eager_ms
    h2d = int(sum(t.nbytes for t in dev.values()))
    return {"metric": "video train iter/s", "value": 1000.0 / ms, "unit": "iter/s", "ms_per_iter": ms, "steps": steps,
            "warmup": warmup, "gpu_launches_per_iter": int(per_iter), "cuda_graph": bool(graph), "dtype": "bf16",
            "e2e": {"value": 1000.0 / ms, "unit": "iter/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 64,
                    "note": "identical to value: the host noise draw, the H2D of clip / noise and the loss read-back "
                            "are inside every timed iteration"},
            "roofline": roof,
            "last_losses": {"D": float(dl), "G": float(gl)},
            "config": {"workload": "train_video.py GAN-phase iteration (D step + G step, train_depth 1) at the finest "
                                   "scale %dx%dx%d (BASELINE config 3) of the full %d-scale pyramid, batch 1, synthetic "
                                   "clip, random-init weights" % (top + (opt.stop_scale + 1,))}}


def train_vae_bench(hpvg, opt, steps, warmup, st, peaks, peaks_kind, scale_idx=2, graph=True, nd=3):
    """BASELINE config 2 (nd=3) / config 1 (nd=2): VAE-phase train iteration at a coarse scale (train_video.py:170-172,
    train_image.py:151-153): one G step — reconstruction forward through encoder, decoder and the refinement stages in
    BatchNorm-train mode, MSE + MSE + KL, backward into encode / decoder / body[-1] (through the resize backward),
    per-tensor clip + Adam with the per-group learning rates of train_video.py:76-105."""
    from hpvg import driver, networks_2d as n2, networks_3d as n3, ops, sampling, train as T
    from hpvg.utils import images as uimg
    G = (n3 if nd == 3 else n2).GeneratorHPVAEGAN(opt, seed=0)
    for _ in range(scale_idx):
        G.init_next_stage()
    amps = [1.0] + [0.1] * scale_idx
    rng = np.random.default_rng(0)
    shp = (lambda i: uimg.scale_shape(opt, i)) if nd == 3 else (lambda i: uimg.scale_shape_2d(opt, i))
    shapes = {"real": (1, 3) + shp(scale_idx), "real_zero": (1, 3) + shp(0), "noise": (1, opt.latent_dim) + shp(0)}
    host = {k: hpvg.PinnedBuffer(int(np.prod(v)) * 4) for k, v in shapes.items()}
    for k, v in shapes.items():
        a = rng.standard_normal(v).astype(np.float32)
        host[k].as_array(v)[...] = np.tanh(a) if k != "noise" else a
    dev = {k: hpvg.Tensor(v, hpvg.F32) for k, v in shapes.items()}
    groups, body_idx, codec = driver.generator_param_groups(opt, G, scale_idx)
    optG = T.ClippedAdam(opt, groups, opt.lr_g, beta1=opt.beta1, beta2=0.999, device_step=graph)
    cells = [G.body[i] for i in body_idx] + [G.encode, G.decoder]
    g_step = T.TrainOneStepCell(T.GWithLoss(opt, None, G, device_rng=graph), optG, cells_to_invalidate=cells)
    G.set_train(True)
    kw = dict(isVAE=True, trainable_body=body_idx, train_codec=codec)

    def upload():
        for k in dev:
            hpvg.lib.hpvg_h2d(dev[k].ptr, host[k].ptr, dev[k].nbytes, st.handle)

    def eager_iter():
        upload()
        return g_step(dev["real"], dev["real_zero"], dev["noise"], amps, stream=st, **kw)

    for _ in range(2):
        eager_iter()
    st.sync()
    ops.start_flop_count()
    eager_iter()
    st.sync()
    fl = ops.stop_flop_count()
    flop_iter = fl["conv"] + fl["wgrad"]
    if graph:
        it = T.GraphedIteration(st, g_step, None, dev["real"], dev["real_zero"], dev["noise"], amps, kw)
        upload()
        it.warmup(max(warmup, 2))
        it.capture()
        per_iter = it.kernels_per_launch

        def one_iter():
            upload()
            return it()[1]
    else:
        per_iter, one_iter = None, eager_iter
    l0 = hpvg.lib.hpvg_launch_count()
    ms, gl = _time_iters(hpvg, st, one_iter, steps, warmup)
    if per_iter is None:
        per_iter = (hpvg.lib.hpvg_launch_count() - l0) // (steps + warmup)
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    tfl = flop_iter / (ms / 1e3) / 1e12
    dims = "x".join(str(v) for v in shp(scale_idx))
    name = ("train_video.py VAE-phase iteration (BASELINE config 2) at scale %d = %s of the 13-frame pyramid"
            if nd == 3 else "train_image.py VAE-phase iteration (BASELINE config 1: 2-D, small pyramid img_size %d) at "
                            "scale %%d = %%s" % opt.img_size) % (scale_idx, dims)
    return {"metric": "video train iter/s" if nd == 3 else "image train iter/s", "value": 1000.0 / ms, "unit": "iter/s",
            "ms_per_iter": ms, "steps": steps, "warmup": warmup, "gpu_launches_per_iter": int(per_iter),
            "cuda_graph": bool(graph), "dtype": "bf16", "last_loss": float(gl),
            "e2e": {"value": 1000.0 / ms, "unit": "iter/s", "h2d_bytes_per_step": int(sum(t.nbytes for t in dev.values())),
                    "d2h_bytes_per_step": 8, "note": "identical to value: H2D of the clips / noise and the loss "
                                                     "read-back are inside every timed iteration"},
            "roofline": {"bound": "latency", "iteration_tflops": tfl, "peak": peak, "unit": "TFLOP/s",
                         "iteration_frac": tfl / peak, "flop_per_iteration": flop_iter,
                         "note": "%d launches of <= %d voxels each: launch / dependency latency bound, not a roofline "
                                 "regime (DESIGN.md §6)" % (per_iter, int(np.prod(shp(scale_idx))))},
            "config": {"workload": name + ": G step with encode + decoder + body[-1] trainable, batch 1, synthetic "
                                          "input, random-init weights"}}


# ================================================================================================= GPU arm: sampling
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import hpvg
    from hpvg import fid as hfid, networks_3d as n3, ops, sampling
    from hpvg import dist as hdist
    from hpvg.utils import images as uimg
    hpvg.init(local_rank)
    st = hpvg.Stream()
    opt = uimg.default_opt(img_size=args.img_size)
    net = n3.GeneratorHPVAEGAN(opt, seed=0)
    for _ in range(opt.stop_scale):
        net.init_next_stage()
    amps = [1.0] + [0.1] * opt.stop_scale
    B = args.batch
    zshape = sampling.z_init_size(opt, B)
    top = uimg.scale_shape(opt, opt.stop_scale)
    out_shape = (B, opt.nc_im) + top
    rng = np.random.default_rng(1234 + rank)
    z_dev = hpvg.from_numpy(rng.standard_normal(zshape).astype(np.float32))
    W = max(args.warmup, 3)
    peaks, peaks_kind = load_peaks()
    extras = world == 1 and not args.no_extras

    def step_device():
        # inputs resident in HBM; the refinement noise (device Philox) is keyed by the running sample counter, so every
        # step generates NEW clips
        x, _ = net(z_dev, amps, noise_init=z_dev, isRandom=True, stream=st)
        return x

    def barrier():
        hpvg.device_sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        import torch
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_device(fn, steps, profile_mode=None):
        barrier()
        e0, e1 = hpvg.Event(), hpvg.Event()
        l0 = hpvg.lib.hpvg_launch_count()
        if profile_mode is not None:
            ops.start_profile(profile_mode)
        e0.record(st)
        for _ in range(steps):
            fn()
        e1.record(st)
        e1.sync()
        prof = ops.stop_profile() if profile_mode is not None else None
        ms = e0.elapsed_ms(e1)
        launches = hpvg.lib.hpvg_launch_count() - l0
        barrier()
        return max_over_ranks(ms), launches, prof

    def timed_pipeline(noise, steps, fused=False):
        """`steps` chunks of B clips through sampling.SamplePipeline (the public API): wall-clock of pipe.run() between
        two device-wide barriers — host draws, H2D, generation, D2H of every clip, the last clip on the host."""
        threads = max(1, min(8, (os.cpu_count() or 2) // max(local_world, 1)))
        pipe = sampling.SamplePipeline(net, amps, B, seed=7 + rank, stream=st, threads=threads, noise=noise, fused=fused)
        base = rank * 1000000
        warm = [[base + c * B + i for i in range(B)] for c in range(W)]
        chunks = [[base + (W + c) * B + i for i in range(B)] for c in range(steps)]
        sink_sum = [0.0]

        def sink(chunk, clips):     # the "device->host read of the step's result": touch one value per clip on the host
            sink_sum[0] += float(clips[:, 0, 0, 0, 0].sum())

        pipe.run(warm, sink)
        barrier()
        h0, d0 = pipe.h2d_bytes, pipe.d2h_bytes
        t0 = time.perf_counter()
        n = pipe.run(chunks, sink)
        hpvg.device_sync()
        ms = (time.perf_counter() - t0) * 1e3
        barrier()
        info = {"h2d": (pipe.h2d_bytes - h0) // steps, "d2h": (pipe.d2h_bytes - d0) // steps, "threads": threads, "n": n}
        pipe.close()
        return max_over_ranks(ms), info

    for _ in range(W):
        step_device()
    st.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, launches, prof = timed_device(step_device, args.steps, profile_mode=ops.CONV_64_64)
    ms_e2e, e2e_info = timed_pipeline("host", args.steps)
    ms_e2e_dev, e2e_dev_info = timed_pipeline("device", args.steps)
    ms_e2e_fused, e2e_fused_info = timed_pipeline("host", args.steps, fused=True)
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    clips = world * B * args.steps
    value = clips / (ms_dev / 1000.0)
    flops = sum(v * FLOP_PER_VOXEL_64 for v, _ in prof)
    kms = sum(ms for _, ms in prof)
    achieved = flops / (kms / 1000.0) / 1e12 if kms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    bpv, bpv_src = committed_traffic()
    roofline = {"bound": "tensor", "kernel": "conv3d_umma_kernel<64->64> (tcgen05 cta_group::2 implicit GEMM)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % peaks_kind,
                "launches_timed": len(prof), "avg_launch_ms": kms / max(len(prof), 1),
                "share_of_step": kms / ms_dev,
                "traffic": None if bpv is None else bpv * (sum(v for v, _ in prof) / max(len(prof), 1)),
                "traffic_unit": "bytes per launch (mean over timed launches)",
                "traffic_source": "COMMITTED ncu --set full capture (%s), dram bytes per voxel x this run's voxels per "
                                  "launch; not a live counter (algorithmic: 256 B/voxel)" % bpv_src}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": sample_workload_name(opt.stop_scale, top, B),
                   "parallelism": "replicated weights, samples sharded by index",
                   "l2_policy": "per-layer activations (%.0f MB at the finest scale) exceed the 126 MB L2" %
                                (B * np.prod(top) * 128 / 1e6)},
        "e2e": {"value": clips / (ms_e2e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(e2e_info["h2d"]),
                "d2h_bytes_per_step": int(e2e_info["d2h"]), "ms_per_step": ms_e2e / args.steps,
                "api": "hpvg.sampling.SamplePipeline.run (the loop inside sampling.generate)",
                "host_draw_threads": e2e_info["threads"],
                "note": "per-sample random numbers of z drawn on the host (numpy uniforms, worker threads) into pinned "
                        "memory, H2D, Box-Muller + generation on the device, D2H of every clip; wall clock between "
                        "device-wide barriers, the last clip on the host",
                "device_noise": {"value": clips / (ms_e2e_dev / 1000.0), "unit": UNIT, "h2d_bytes_per_step": 0,
                                 "d2h_bytes_per_step": int(e2e_dev_info["d2h"]),
                                 "note": "same call with noise='device': z from the device Philox generator keyed by "
                                         "(seed, sample index) — no host draw, clips still copied to the host"},
                "fused_entry": {"value": clips / (ms_e2e_fused / 1000.0), "unit": UNIT,
                                "h2d_bytes_per_step": int(e2e_fused_info["h2d"]),
                                "d2h_bytes_per_step": int(e2e_fused_info["d2h"]),
                                "note": "the same pipeline (host draw included) with the generation issued as ONE C call "
                                        "per chunk, hpvg_generator_sample (include/hpvg.h), instead of one call per "
                                        "layer from Python; bit-identical clips"}},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": roofline,
    }

    # ------------------------------------------------------------------ config 5: FID moments + NCCL all-gather (every N)
    if not args.no_fid and not args.no_extras:
        try:
            comm = hdist.from_env()
            feats = hfid.C3DBlock0()
            n_fid = world * B * 2

            def fid_pass(seed):
                return sampling.generate_moments(net, amps, n_fid, feats, comm, batch=B, seed=seed, stream=st,
                                                 threads=max(1, min(8, (os.cpu_count() or 2) // max(local_world, 1))))

            fid_pass(1)
            barrier()
            t0 = time.perf_counter()
            rows, count, _ = fid_pass(2)
            hpvg.device_sync()
            ms_fid = max_over_ranks((time.perf_counter() - t0) * 1e3)
            real = hpvg.from_numpy(np.tanh(np.random.default_rng(5).standard_normal((1, 3) + top)).astype(np.float32))
            real_row = hfid.sample_moments(feats(real, stream=st), stream=st).numpy(st)[0]
            svfid, _ = hfid.svfid_from_moments(real_row, rows, count)
            line["fid"] = {"metric": "sampled clips/s incl. sinFID moments", "value": n_fid / (ms_fid / 1e3),
                           "unit": UNIT, "n_gpus": world, "clips": n_fid, "ms": ms_fid,
                           "gather": "%s, %d ranks x %d rows x %d fp32" % (type(comm).__name__, world, n_fid // world,
                                                                          hfid.MOMENT_FLOATS),
                           "gather_bytes_per_rank": int(n_fid // world * hfid.MOMENT_FLOATS * 4),
                           "svfid_vs_synthetic_clip": svfid,
                           "config": {"workload": "eval_video.py + sinFID statistics (BASELINE config 5): per-sample "
                                                  "C3D-block-0 feature moments on the device, ONE all-gather, Frechet "
                                                  "distance on the host; random-init generator and feature weights"}}
            if hasattr(comm, "close"):
                comm.close()
        except Exception as e:   # evidence, never a reason to lose the headline line
            sys.stderr.write("bench: fid measurement failed: %r\n" % (e,))

    # ------------------------------------------------------------------ tf32 precision mode (N == 1)
    if extras and not args.no_tf32:
        try:
            hpvg.set_precision("tf32")
            net.ws = n3.Workspace()
            for _ in range(2):
                step_device()
            ms32, l32, prof32 = timed_device(step_device, max(2, args.steps // 3), profile_mode="all")
            steps32 = max(2, args.steps // 3)
            r32 = _conv_roofline(prof32, ms32, peaks, peaks_kind, tf32=True)
            line["tf32"] = {"metric": METRIC, "value": B * steps32 / (ms32 / 1e3), "unit": UNIT, "dtype": "tf32",
                            "steps": steps32, "ms_per_step": ms32 / steps32, "gpu_launches": int(l32), "roofline": r32,
                            "note": "same workload in the tf32 precision mode (fp32 channels-last activations, tcgen05 "
                                    "kind::tf32; per-layer rel-L2 <= 1e-3 vs the fp32 oracle): device-resident value"}
        except Exception as e:
            sys.stderr.write("bench: tf32 measurement failed: %r\n" % (e,))
        finally:
            hpvg.set_precision("bf16")
            net.ws = n3.Workspace()

    # ------------------------------------------------------------------ train workloads (N == 1)
    if extras and not args.no_train:
        del net
        hpvg.empty_cache()
        mk = lambda **kw: uimg.default_opt(img_size=args.img_size, **kw)      # noqa: E731
        try:
            line["train"] = train_iter_bench(hpvg, mk(), args.train_steps, 3, st, peaks, peaks_kind,
                                             graph=not args.no_graph, frames=args.train_frames)
        except hpvg.HpvgError as e:     # graph capture refused on this driver/box: same iteration, launched eagerly
            if args.no_graph:
                raise
            sys.stderr.write("bench: CUDA-graph train iteration failed (%s); falling back to eager launches\n" % e)
            hpvg.device_sync()
            line["train"] = train_iter_bench(hpvg, mk(), args.train_steps, 3, st, peaks, peaks_kind, graph=False,
                                             frames=args.train_frames)
        for key, fn in (("train_vae", lambda g: train_vae_bench(hpvg, mk(), 50, 5, st, peaks, peaks_kind, 2, g, 3)),
                        ("train_image", lambda g: train_vae_bench(hpvg, uimg.default_opt(**IMAGE_OPT), 50, 5, st, peaks,
                                                                  peaks_kind, IMAGE_SCALE, g, 2))):
            try:
                line[key] = fn(not args.no_graph)
            except Exception as e:
                sys.stderr.write("bench: %s with a CUDA graph failed (%r); eager\n" % (key, e))
                try:
                    hpvg.device_sync()
                    line[key] = fn(False)
                except Exception as e2:
                    sys.stderr.write("bench: %s failed: %r\n" % (key, e2))

    if extras and not args.no_hbm:
        # achieved HBM bandwidth of the bandwidth-bound kernels on batched inputs (tools/bench_hbm.py: >= 0.5 GB per
        # launch, CUDA events on our stream) against the measured copy peak: north_star's resize / BN / Adam target
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_hbm
            hpvg.empty_cache()
            line["hbm_kernels"] = [{"kernel": r["kernel"], "gbs": round(r["achieved_gbs"], 1), "frac": round(r["frac"], 3)}
                                   for r in bench_hbm.measure(st, log=None)]
            line["hbm_peak"] = {"gbs": bench_hbm.peak_gbs()[0], "source": bench_hbm.peak_gbs()[1] + " copy bandwidth"}
        except Exception as e:
            sys.stderr.write("bench: hbm_kernels table failed: %r\n" % (e,))

    # ------------------------------------------------------------------ CPU baselines (rank 0, N == 1, bounded samples)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores = cpu_sample_clips({"img_size": args.img_size}, args.cpu_clips)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d clips, CPU restatement of the reference (torch-CPU/oneDNN fp32), same "
                                          "pyramid" % args.cpu_clips}
        if "train_vae" in line:
            v, cores, _ = cpu_train_vae_iter({"img_size": args.img_size}, 2, 3, 3)
            line["train_vae"]["cpu_baseline"] = {"value": v, "unit": "iter/s", "cores": cores, "kind": "port",
                                                 "sample": "3 iterations of the same G step on the oracle"}
        if "train_image" in line:
            v, cores, _ = cpu_train_vae_iter(IMAGE_OPT, IMAGE_SCALE, 2, 3)
            line["train_image"]["cpu_baseline"] = {"value": v, "unit": "iter/s", "cores": cores, "kind": "port",
                                                   "sample": "3 iterations of the same G step on the oracle (nd=2)"}
        if "train" in line:
            v, cores, _ = cpu_train_gan_iter({"img_size": args.img_size}, args.train_frames, 1)
            line["train"]["cpu_baseline"] = {"value": v, "unit": "iter/s", "cores": cores, "kind": "port",
                                             "sample": "1 iteration (D step + G step incl. Adam) of the oracle at the "
                                                       "same %d-frame finest scale" % args.train_frames}

    if args.workload != "sample" and args.workload in line:
        o = line[args.workload]
        line.update(metric=o["metric"], value=o["value"], unit=o["unit"], config=o["config"],
                    ms_per_step=o.get("ms_per_iter", o.get("ms")), steps=o.get("steps", args.steps),
                    scaling="weak" if args.workload == "fid" else "replicas only")
        for k in ("e2e", "roofline", "cpu_baseline", "dtype"):
            if k in o:
                line[k] = o[k]
    # compact digest of the extra objects LAST on the line (a log tail keeps the end of a long line)
    digest = {}
    for k in ("train", "train_vae", "train_image", "tf32", "fid"):
        o = line.get(k)
        if not o:
            continue
        e = {"value": round(o["value"], 2), "unit": o["unit"]}
        r = o.get("roofline", {})
        for kk in ("frac", "iteration_frac"):
            if kk in r:
                e["roofline_" + kk] = round(r[kk], 3)
        if "cpu_baseline" in o:
            e["cpu"] = round(o["cpu_baseline"]["value"], 4)
        if "gpu_launches_per_iter" in o:
            e["launches"] = o["gpu_launches_per_iter"]
        digest[k] = e
    if "hbm_kernels" in line:
        digest["hbm_frac"] = {r["kernel"].split(" (")[0][:28]: r["frac"] for r in line["hbm_kernels"]}
    if digest:
        line["digest"] = digest
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
