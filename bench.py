#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native HP-VAE-GAN hot path.

Metric (BASELINE.json): sampled clips/s — `eval_video.py` semantics (one clip = one random-mode generator forward
through the full 10-scale pyramid, 13 frames x 192 x 257 at the finest scale), weak-scaled over N GPUs: every rank
generates `--batch` clips per step from its own noise, no collective on the data path.
`--workload train` instead times GAN-phase train iterations on 1 GPU when the training path is built.

  python bench.py --gpus 1 --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                      times the CPU oracle (the reference cannot be installed)

Prints ONE JSON line on rank 0 (contract in the task statement): value = device-resident throughput, e2e = through
the public API with host buffers (H2D of the noise, D2H of the clips inside the timed region), roofline for the
dominant kernel (tcgen05 conv 64->64) measured live with CUDA events, cpu_baseline on rank 0 at N == 1."""
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=2, help="clips in the bounded CPU-baseline sample")
    ap.add_argument("--workload", default="sample", choices=["sample", "train"],
                    help="sample: sampled clips/s (scales over GPUs); train: GAN-phase train iter/s on 1 GPU")
    ap.add_argument("--train-steps", type=int, default=10, help="train iterations timed for the extra `train` object")
    ap.add_argument("--no-train", action="store_true", help="skip the extra train-iter/s measurement at N == 1")
    ap.add_argument("--train-frames", type=int, default=16,
                    help="frames of the synthetic clip at the finest scale for the train workload (BASELINE.json config 3: "
                         "16; the reference's own schedule gives 13)")
    ap.add_argument("--no-hbm", action="store_true", help="skip the extra `hbm_kernels` table at N == 1")
    ap.add_argument("--no-graph", action="store_true", help="train iterations launched eagerly instead of as a CUDA graph")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def traffic_per_voxel():
    p = os.path.join(ROOT, "profiles", "conv64_traffic.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_voxel"])
    except Exception:
        return float("nan")


def workload_name(opt, batch):
    from hpvg.utils import images as uimg
    t, h, w = uimg.scale_shape(opt, opt.stop_scale)
    return ("eval_video.py random-noise sampling, full %d-scale pyramid, finest scale %dx%dx%d, %d clips/step/GPU, "
            "random-init weights" % (opt.stop_scale + 1, t, h, w, batch))


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


# ------------------------------------------------------------------------------------------------- CPU oracle arm
def cpu_sample_clips(opt_kw, n_clips, threads=None):
    """Time the CPU oracle (torch-CPU fp32, all host cores) generating n_clips full-pyramid samples."""
    import torch
    from oracle import hpvg_oracle as orc
    # all host cores: torchrun exports OMP_NUM_THREADS=1 to its workers, which would throttle the CPU arm
    torch.set_num_threads(threads or os.cpu_count() or 1)
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
    dt = time.perf_counter() - t0
    return n_clips / dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    opt_kw = {"img_size": args.img_size}
    from oracle import hpvg_oracle as orc
    opt = orc.default_opt(**opt_kw)
    for _ in range(min(args.warmup, 1)):
        cpu_sample_clips(opt_kw, 1)
    vals = []
    t0 = time.perf_counter()
    cores = os.cpu_count()
    for _ in range(args.steps):
        v, cores = cpu_sample_clips(opt_kw, 1)
        vals.append(v)
    total = time.perf_counter() - t0
    value = args.steps / total
    from hpvg.utils import images as uimg
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(uimg.default_opt(**opt_kw), 1) + " (each step = 1 clip on the host CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d steps x 1 clip, CPU restatement of the reference (torch-CPU/oneDNN fp32); "
                                   "MindSpore itself is not installable offline" % args.steps},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0



# ------------------------------------------------------------------------------------------------- train iter/s
def train_iter_bench(hpvg, opt, steps, warmup, st, graph=True, frames=None):
    """One GAN-phase train iteration at the finest scale (train_video.py:170-177): D step (3 D forwards, 2 backwards,
    WGAN-GP double backward, Adam) + G step (reconstruction forward of the whole pyramid in BatchNorm-train mode,
    backward of the last stage, random forward + D forward for the loss value, ClippedAdam).  Inputs come from
    pinned host buffers every iteration (the data loader + host noise of the reference); losses are read back."""
    from hpvg import networks_3d as n3, train as T, sampling
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
    optG = T.ClippedAdam(opt, [{"params": T.trainable_params(block), "lr": opt.lr_g}], opt.lr_g, beta1=opt.beta1,
                         beta2=0.999, device_step=graph)
    optD = T.Adam(T.trainable_params(D), opt.lr_d, beta1=opt.beta1, beta2=0.999, device_step=graph)
    g_step = T.TrainOneStepCell(T.GWithLoss(opt, D, G, device_rng=graph), optG, cells_to_invalidate=[block])
    d_step = T.TrainOneStepCell(T.DWithLoss(opt, D, G, device_rng=graph), optD, cells_to_invalidate=[D])
    g_step.set_train()
    d_step.set_train()
    nb = len(G.body)
    # second pinned noise buffer: the host draws iteration i+1's noise_init (numpy, like images.py:17-21) while the GPU
    # runs iteration i; an event per buffer guards its reuse
    host2 = hpvg.PinnedBuffer(host["noise"].nbytes)
    noise_bufs = [host["noise"], host2]
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
        def one_iter():
            upload()
            dl = d_step(dev["real"], dev["noise"], amps, stream=st)
            gl = g_step(dev["real"], dev["real_zero"], dev["noise"], amps, isVAE=False, trainable_body=(nb - 1,),
                        stream=st)
            return dl, gl
        per_iter = None

    for _ in range(warmup):
        one_iter()
    st.sync()
    l0 = hpvg.lib.hpvg_launch_count()
    e0, e1 = hpvg.Event(), hpvg.Event()
    e0.record(st)
    for _ in range(steps):
        dl, gl = one_iter()
    e1.record(st)
    e1.sync()
    ms = e0.elapsed_ms(e1) / steps
    if per_iter is None:
        per_iter = (hpvg.lib.hpvg_launch_count() - l0) // steps
    return {"metric": "video train iter/s", "value": 1000.0 / ms, "unit": "iter/s", "ms_per_iter": ms, "steps": steps,
            "warmup": warmup, "gpu_launches_per_iter": int(per_iter), "cuda_graph": bool(graph),
            "h2d_bytes_per_iter": int(sum(t.nbytes for t in dev.values())), "d2h_bytes_per_iter": 64,
            "last_losses": {"D": float(dl), "G": float(gl)},
            "config": {"workload": "train_video.py GAN-phase iteration (D step + G step, train_depth 1) at the finest "
                                   "scale %dx%dx%d (BASELINE.json config 3) of the full %d-scale pyramid, batch 1, synthetic clip, random-init "
                                   "weights; host noise draw, host->device copies of the clip/noise and loss "
                                   "read-backs included; the iteration is replayed as one CUDA graph"
                                   % (top + (opt.stop_scale + 1,))}}

def train_vae_bench(hpvg, opt, steps, warmup, st, scale_idx=2, graph=True):
    """BASELINE.json config 2: VAE-phase train iteration at the coarsest scales (train_video.py:170-172 with
    scale_idx < vae_levels): one G step — reconstruction forward through encoder, decoder and the refinement stages in
    BatchNorm-train mode, MSE + MSE + KL losses, backward into encode / decoder / body[-1] (through the resize
    backward), per-tensor clip + Adam with the per-group learning rates of train_video.py:76-105."""
    from hpvg import driver, networks_3d as n3, train as T, sampling
    from hpvg.utils import images as uimg
    G = n3.GeneratorHPVAEGAN(opt, seed=0)
    for _ in range(scale_idx):
        G.init_next_stage()
    amps = [1.0] + [0.1] * scale_idx
    rng = np.random.default_rng(0)
    shapes = {"real": (1, 3) + uimg.scale_shape(opt, scale_idx), "real_zero": (1, 3) + uimg.scale_shape(opt, 0),
              "noise": sampling.z_init_size(opt, 1)}
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
        per_iter = None

        def one_iter():
            upload()
            return g_step(dev["real"], dev["real_zero"], dev["noise"], amps, stream=st, **kw)

    for _ in range(warmup):
        one_iter()
    st.sync()
    l0 = hpvg.lib.hpvg_launch_count()
    e0, e1 = hpvg.Event(), hpvg.Event()
    e0.record(st)
    for _ in range(steps):
        gl = one_iter()
    e1.record(st)
    e1.sync()
    ms = e0.elapsed_ms(e1) / steps
    if per_iter is None:
        per_iter = (hpvg.lib.hpvg_launch_count() - l0) // steps
    t, h, w = uimg.scale_shape(opt, scale_idx)
    return {"metric": "video train iter/s", "value": 1000.0 / ms, "unit": "iter/s", "ms_per_iter": ms, "steps": steps,
            "warmup": warmup, "gpu_launches_per_iter": int(per_iter), "cuda_graph": bool(graph),
            "last_loss": float(gl),
            "config": {"workload": "train_video.py VAE-phase iteration (BASELINE.json config 2) at scale %d = %dx%dx%d of "
                                   "the 13-frame pyramid: G step with encode + decoder + body[-1] trainable, batch 1, "
                                   "synthetic clip, random-init weights; H2D of the clips and loss read-back included"
                                   % (scale_idx, t, h, w)}}


# ------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import hpvg
    from hpvg import networks_3d as n3, ops, sampling
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
    out_shape = (B, opt.nc_im) + uimg.scale_shape(opt, opt.stop_scale)
    rng = np.random.default_rng(1234 + rank)
    z_host = hpvg.PinnedBuffer(int(np.prod(zshape)) * 4)
    z_host.as_array(zshape)[...] = rng.standard_normal(zshape).astype(np.float32)
    out_host = hpvg.PinnedBuffer(int(np.prod(out_shape)) * 4)
    z_dev = hpvg.Tensor(zshape, hpvg.F32)
    hpvg.lib.hpvg_h2d(z_dev.ptr, z_host.ptr, z_dev.nbytes, st.handle)
    st.sync()

    def step_device():
        net.sample_counter = 0
        x, _ = net(z_dev, amps, noise_init=z_dev, isRandom=True, stream=st)
        return x

    # e2e: the user-facing call with HOST buffers.  Copies run on a second stream so that the D2H of clip batch i and the
    # H2D of noise batch i+1 overlap the generation of batch i+1 (double-buffered noise / clip tensors, events for
    # ordering); every byte still moves inside the timed region.
    cp = hpvg.Stream()       # device -> host copies
    cp_in = hpvg.Stream()    # host -> device copies (own stream: the next batch's noise must not queue behind a D2H)
    z_devs = [z_dev, hpvg.Tensor(zshape, hpvg.F32)]
    out_hosts = [out_host, hpvg.PinnedBuffer(int(np.prod(out_shape)) * 4)]
    e2e = {"i": 0, "h2d_done": [None, None], "d2h_done": [None, None], "gen_done": [None, None]}

    def step_e2e():
        k = e2e["i"] & 1
        e2e["i"] += 1
        # noise k: host -> device on the copy stream (its previous consumer, step i-2, finished: gen_done[k])
        if e2e["gen_done"][k] is not None:
            cp_in.wait_event(e2e["gen_done"][k])
        hpvg.lib.hpvg_h2d(z_devs[k].ptr, z_host.ptr, z_devs[k].nbytes, cp_in.handle)
        ev = hpvg.Event(); ev.record(cp_in); e2e["h2d_done"][k] = ev
        # generation on the main stream: needs noise k, and clip buffer k free (its D2H of step i-2 done)
        st.wait_event(ev)
        if e2e["d2h_done"][k] is not None:
            st.wait_event(e2e["d2h_done"][k])
        net.sample_counter = 0
        net.out_slot = k
        x, _ = net(z_devs[k], amps, noise_init=z_devs[k], isRandom=True, stream=st)
        ev = hpvg.Event(); ev.record(st); e2e["gen_done"][k] = ev
        # clip k: device -> host on the copy stream
        cp.wait_event(ev)
        hpvg.lib.hpvg_d2h(out_hosts[k].ptr, x.ptr, x.nbytes, cp.handle)
        ev = hpvg.Event(); ev.record(cp); e2e["d2h_done"][k] = ev
        return x

    def drain_e2e():
        cp.sync()

    def barrier():
        st.sync()
        cp.sync()
        cp_in.sync()
        if dist is not None:
            dist.barrier()

    def timed(fn, steps, profile=False, drain=None):
        barrier()
        e0, e1 = hpvg.Event(), hpvg.Event()
        l0 = hpvg.lib.hpvg_launch_count()
        if profile:
            ops.start_profile(ops.CONV_64_64)
        e0.record(st)
        for _ in range(steps):
            fn()
        if drain is not None:       # the last clips must have reached the host before the clock stops
            st.wait_event(e2e["d2h_done"][(e2e["i"] - 1) & 1])
        e1.record(st)
        e1.sync()
        if drain is not None:
            drain()
        prof = ops.stop_profile() if profile else None
        ms = e0.elapsed_ms(e1)
        launches = hpvg.lib.hpvg_launch_count() - l0
        barrier()
        if dist is not None:
            import torch
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, prof

    for _ in range(max(args.warmup, 3)):
        step_e2e()
    st.sync()
    cp.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, launches, prof = timed(step_device, args.steps, profile=True)
    ms_e2e, _, _ = timed(step_e2e, args.steps, drain=drain_e2e)
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    clips = world * B * args.steps
    value = clips / (ms_dev / 1000.0)
    e2e_value = clips / (ms_e2e / 1000.0)
    peaks, peaks_kind = load_peaks()
    flops = sum(v * FLOP_PER_VOXEL_64 for v, _ in prof)
    kms = sum(ms for _, ms in prof)
    achieved = flops / (kms / 1000.0) / 1e12 if kms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    roofline = {"bound": "tensor", "kernel": "conv3d_umma_kernel<64->64> (tcgen05 cta_group::2 implicit GEMM)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % peaks_kind,
                "launches_timed": len(prof), "avg_launch_ms": kms / max(len(prof), 1),
                "share_of_step": kms / ms_dev,
                # dram__bytes_read+write per voxel from the committed `ncu --set full` capture of this kernel
                # (profiles/conv64_traffic.json; algorithmic = 2 x 128 B/voxel), scaled to this run's mean launch
                "traffic": traffic_per_voxel() * (sum(v for v, _ in prof) / max(len(prof), 1)),
                "traffic_unit": "bytes per launch (mean over timed launches)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(opt, B), "parallelism": "replicated weights, samples sharded by index",
                   "l2_policy": "per-layer activations (%.0f MB at the finest scale) exceed the 126 MB L2" %
                                (B * np.prod(uimg.scale_shape(opt, opt.stop_scale)) * 128 / 1e6)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(z_dev.nbytes),
                "d2h_bytes_per_step": int(np.prod(out_shape)) * 4, "ms_per_step": ms_e2e / args.steps,
                "note": "pinned host noise in, host clips out, every step; copies on a second stream overlap the next "
                        "step's generation (double-buffered), the clock stops after the last clip reached the host"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": roofline,
    }
    if world == 1 and not args.no_train:
        del net
        try:
            line["train"] = train_iter_bench(hpvg, uimg.default_opt(img_size=args.img_size), args.train_steps, 3, st,
                                             graph=not args.no_graph, frames=args.train_frames)
        except hpvg.HpvgError as e:     # graph capture refused on this driver/box: same iteration, launched eagerly
            if args.no_graph:
                raise
            sys.stderr.write("bench: CUDA-graph train iteration failed (%s); falling back to eager launches\n" % e)
            hpvg.device_sync()
            line["train"] = train_iter_bench(hpvg, uimg.default_opt(img_size=args.img_size), args.train_steps, 3, st,
                                             graph=False, frames=args.train_frames)
        try:
            line["train_vae"] = train_vae_bench(hpvg, uimg.default_opt(img_size=args.img_size), 50, 5, st,
                                                graph=not args.no_graph)
        except hpvg.HpvgError as e:
            sys.stderr.write("bench: VAE-phase train measurement failed: %s\n" % e)
        if args.workload == "train":
            tr = line["train"]
            line.update(metric=tr["metric"], value=tr["value"], unit=tr["unit"], ms_per_step=tr["ms_per_iter"],
                        steps=tr["steps"], warmup=tr["warmup"], scaling="replicas only", config=tr["config"],
                        e2e={"value": tr["value"], "unit": tr["unit"], "h2d_bytes_per_step": tr["h2d_bytes_per_iter"],
                             "d2h_bytes_per_step": 64})
    if world == 1 and not args.no_hbm:
        # achieved HBM bandwidth of the bandwidth-bound kernels on batched inputs (tools/bench_hbm.py: >= 0.5 GB per
        # launch, CUDA events on our stream) against the measured copy peak: north_star's resize / BN / Adam target
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_hbm
            hpvg.empty_cache()
            line["hbm_kernels"] = [{"kernel": r["kernel"], "achieved_gbs": round(r["achieved_gbs"], 1),
                                    "frac": round(r["frac"], 3), "mb_per_launch": round(r["bytes_per_launch"] / 1e6, 1)}
                                   for r in bench_hbm.measure(st, log=None)]
            line["hbm_peak"] = {"gbs": bench_hbm.peak_gbs()[0], "source": bench_hbm.peak_gbs()[1] + " copy bandwidth"}
        except Exception as e:   # the table is evidence, never a reason to lose the headline line
            sys.stderr.write("bench: hbm_kernels table failed: %r\n" % (e,))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores = cpu_sample_clips({"img_size": args.img_size}, args.cpu_clips)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d clips, CPU restatement of the reference (torch-CPU/oneDNN fp32), "
                                          "same pyramid" % args.cpu_clips}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
