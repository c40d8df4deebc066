"""Progressive per-scale training — the step semantics of the reference's `train_video.py:22-227` / `train_image.py`
 on the libhpvg training step.  Only the compute-relevant logic is mirrored: the
discriminator warm start from the previous scale (:53-62), the generator parameter groups and their learning-rate
schedule (:76-105), the first-iteration noise-amplitude calibration (:152-166), the VAE / GAN phase switch (:170-177)
and the per-scale state that gets saved (:224-227).  Data loading, logging, progress bars and image dumps are out of
scope; `real_at(scale_idx)` supplies the (already resized, [-1,1]) clip of a scale as a float32 numpy array."""
import os

import numpy as np

from . import checkpoint, ops
from . import train as T
from .networks_3d import as5d
from .runtime import F32, Tensor, from_numpy


def generator_param_groups(opt, netG, scale_idx):
    """train_video.py:76-105 -> ([{'params': [(name, Tensor)], 'lr': float}], trainable body indices, train_codec)."""
    groups, body_idx, codec = [], [], False
    nb = len(netG.body)

    def body_groups(indices):
        k = len(indices)
        for j, s in enumerate(indices):
            groups.append({"params": T.trainable_params(netG.body[s], "body.%d." % s),
                           "lr": opt.lr_g * (opt.lr_scale ** (k - 1 - j))})
            body_idx.append(s)

    def codec_groups():
        lr = opt.lr_g * (opt.lr_scale ** scale_idx)
        groups.append({"params": T.trainable_params(netG.encode, "encode."), "lr": lr})
        groups.append({"params": T.trainable_params(netG.decoder, "decoder."), "lr": lr})

    if not opt.train_all:
        if opt.vae_levels < scale_idx + 1:
            depth = min(opt.train_depth, nb - opt.vae_levels + 1)
            body_groups(list(range(nb))[-depth:])
        else:
            codec_groups()
            codec = True
            body_groups(list(range(nb))[-opt.train_depth:] if nb else [])
    else:
        if nb < opt.train_depth:
            codec_groups()
            codec = True
            body_groups(list(range(nb)))
        else:
            body_groups(list(range(nb))[-opt.train_depth:])
    return groups, tuple(body_idx), codec


def calibrate_noise_amp(opt, netG, real, real_zero, noise_amps, scale_idx, stream=None):
    """train_video.py:152-166: append this scale's amplitude = noise_amp_init * RMSE(real, G(real_zero)) / batch."""
    if opt.const_amp:
        noise_amps.append(1)
        return 1.0
    if scale_idx == 0:
        noise_amps.append(1)
        return 1.0
    noise_amps.append(0)
    rec = netG(real_zero, noise_amps, isRandom=False, stream=stream)[0]
    mse = float(ops.mse(as5d(real), as5d(rec), stream=stream).numpy(stream)[0])
    amp = opt.noise_amp_init * float(np.sqrt(mse)) / opt.batch_size      # nn.RMSELoss
    noise_amps[-1] = amp
    return amp


def train_scale(opt, netG, D_cls, scale_idx, real, real_zero, noise_amps, niter, save_dir=None, prev_D_state=None,
                z_init_size=None, stream=None, on_iter=None, graph=False):
    """One call of the reference's `train(opt, netG)` for `scale_idx` (netG already has `scale_idx` body stages).
    real / real_zero: float32 numpy clips of this scale / of scale 0.  Returns (D or None, losses list, D state).

    graph=False: every iteration is launched kernel by kernel; all N(0,1) draws come from numpy on the host exactly
    where the reference makes them (Q7).
    graph=True: the first two iterations run the same way, then the iteration (D step + G step + both Adam updates, or
    the VAE-phase G step) is captured ONCE as a CUDA graph and replayed — what bench.py measures: at the coarse scales an
    iteration is launch-latency bound and replays several times faster than it can be launched from Python.  In the GAN
    phase the random numbers of `noise_init` are still drawn by numpy on the host every iteration (on a worker thread,
    one iteration ahead) and uploaded before each replay; the refinement noise inside the forwards comes from the device
    Philox generator keyed by a device-resident draw counter (the host cannot inject draws into a replayed graph), and
    the Adam step counter lives on the device."""
    vae_phase = opt.vae_levels >= scale_idx + 1
    D = None
    d_step = None
    if not vae_phase:
        D = D_cls(opt, rng=np.random.default_rng(1000 + scale_idx))
        if prev_D_state is not None and opt.vae_levels < scale_idx:        # warm start (:59-62)
            checkpoint.load_param_into_net(D, prev_D_state)
        optD = T.Adam(T.trainable_params(D), opt.lr_d, beta1=opt.beta1, beta2=0.999, device_step=graph)
        d_step = T.TrainOneStepCell(T.DWithLoss(opt, D, netG, device_rng=graph), optD, cells_to_invalidate=[D])
    groups, body_idx, codec = generator_param_groups(opt, netG, scale_idx)
    optG = T.ClippedAdam(opt, groups, opt.lr_g, beta1=opt.beta1, beta2=0.999, device_step=graph)
    to_invalidate = [netG.body[s] for s in body_idx] + ([netG.encode, netG.decoder] if codec else [])
    g_step = T.TrainOneStepCell(T.GWithLoss(opt, D, netG, device_rng=graph), optG, cells_to_invalidate=to_invalidate)
    netG.set_train(True)
    if D is not None:
        D.set_train(True)
    t_real, t_zero = from_numpy(np.asarray(real, np.float32)), from_numpy(np.asarray(real_zero, np.float32))
    if z_init_size is None:
        z_init_size = (1, opt.latent_dim) + tuple(netG.stage_shape(0)[(0 if netG.KT == 3 else 1):])
    losses = []
    if graph:
        losses = _train_scale_graphed(opt, netG, scale_idx, g_step, d_step, t_real, t_zero, noise_amps, niter,
                                      z_init_size, body_idx, codec, vae_phase, stream, on_iter)
        niter = 0
    for it in range(niter):
        noise_init = from_numpy(np.random.normal(size=z_init_size).astype(np.float32))      # images.py:17-21 (Q7)
        if it == 0:
            calibrate_noise_amp(opt, netG, t_real, t_zero, noise_amps, scale_idx, stream)
        if vae_phase:
            gl = g_step(t_real, t_zero, noise_init, noise_amps, isVAE=True, trainable_body=body_idx,
                        train_codec=codec, stream=stream)
            losses.append((None, gl))
        else:
            dl = d_step(t_real, noise_init, noise_amps, stream=stream)
            gl = g_step(t_real, t_zero, noise_init, noise_amps, isVAE=False, trainable_body=body_idx,
                        train_codec=codec, stream=stream)
            losses.append((dl, gl))
        if on_iter is not None:
            on_iter(scale_idx, it, losses[-1])
    d_state = checkpoint.state_dict(D) if D is not None else None
    if save_dir is not None:                                                        # train_video.py:224-227
        checkpoint.save_json({"noise_amps": [float(a) for a in noise_amps], "scale_idx": scale_idx},
                             os.path.join(save_dir, "intermediate.json"))
        checkpoint.save_checkpoint(netG, os.path.join(save_dir, "netG_%d" % scale_idx))
        if D is not None:
            checkpoint.save_checkpoint(D, os.path.join(save_dir, "netD_%d" % scale_idx))
    return D, losses, d_state


def _train_scale_graphed(opt, netG, scale_idx, g_step, d_step, t_real, t_zero, noise_amps, niter, z_init_size,
                         body_idx, codec, vae_phase, stream, on_iter):
    """The iteration loop of train_scale as CUDA-graph replays (see train_scale)."""
    from .runtime import PinnedBuffer, Stream
    from ._lib import check, lib
    st = stream or Stream()
    t_real, t_zero = as5d(t_real), as5d(t_zero)
    z5 = tuple(z_init_size) if len(z_init_size) == 5 else (z_init_size[0], z_init_size[1], 1) + tuple(z_init_size[2:])
    noise_dev = Tensor(z5, F32).zero_(st)
    pinned = [PinnedBuffer(noise_dev.nbytes), PinnedBuffer(noise_dev.nbytes)]
    uploaded = [None, None]
    # noise_init is consumed by the GAN phase only (the fake clip of the D step, the random clip of the G step); the VAE
    # phase never reads it (losses.py:77-91), so nothing is drawn there.  GAN phase: the host draws the random numbers of
    # iteration i+1 on a worker thread while the GPU replays iteration i — uniforms from a numpy Generator seeded from the
    # global numpy state (the reference draws from that state, images.py:17-21), turned into N(0,1) on the device after
    # the upload (ops.box_muller_: a quarter of the host cost of drawing normals)
    from concurrent.futures import ThreadPoolExecutor
    gen = np.random.default_rng(int(np.random.randint(0, 2 ** 31 - 1)))
    pool = ThreadPoolExecutor(max_workers=1) if not vae_phase else None
    drawn = {}

    def draw(it):
        k = it & 1
        if uploaded[k] is not None:
            uploaded[k].sync()          # the copy that last read this pinned buffer has completed
        gen.random(dtype=np.float32, out=pinned[k].as_array(z5))
        return k

    def upload(it):
        if vae_phase:
            return
        k = drawn.pop(it).result() if it in drawn else draw(it)
        check(lib.hpvg_h2d(noise_dev.ptr, pinned[k].ptr, noise_dev.nbytes, st.handle), "h2d")
        ev = T.Event()
        ev.record(st)
        uploaded[k] = ev
        ops.box_muller_(noise_dev, stream=st)
        if it + 1 < niter:
            drawn[it + 1] = pool.submit(draw, it + 1)

    kw = dict(isVAE=vae_phase, trainable_body=body_idx, train_codec=codec)
    losses, graphed = [], None
    for it in range(niter):
        upload(it)
        if it == 0:
            calibrate_noise_amp(opt, netG, t_real, t_zero, noise_amps, scale_idx, st)
            graphed = T.GraphedIteration(st, g_step, d_step, t_real, t_zero, noise_dev, noise_amps, kw)
        if it < 2:
            dl, gl = graphed._body(True)        # eager: first-use allocations, caches — and two real iterations
        else:
            if graphed.graph is None:
                st.sync()
                graphed.capture()
            dl, gl = graphed()
        losses.append((dl, gl))
        if on_iter is not None:
            on_iter(scale_idx, it, losses[-1])
    st.sync()
    if pool is not None:
        pool.shutdown(wait=True)
    if graphed is not None:
        graphed.destroy()
    return losses


def train_pyramid(opt, netG, D_cls, real_at, niter, start_scale=0, stop_scale=None, noise_amps=None, save_dir=None,
                  stream=None, on_iter=None, graph=False):
    """The `while opt.scale_idx < opt.stop_scale + 1` loop of train_video.py:413-419."""
    noise_amps = [] if noise_amps is None else noise_amps
    stop_scale = opt.stop_scale if stop_scale is None else stop_scale
    d_state, history = None, []
    real_zero = real_at(0)
    for scale_idx in range(start_scale, stop_scale + 1):
        if scale_idx > 0 and len(netG.body) < scale_idx:
            netG.init_next_stage()
        real = real_at(scale_idx)
        _, losses, d_state = train_scale(opt, netG, D_cls, scale_idx, real, real_zero if scale_idx > 0 else real,
                                         noise_amps, niter, save_dir, d_state, stream=stream, on_iter=on_iter,
                                         graph=graph)
        history.append(losses)
    return noise_amps, history
