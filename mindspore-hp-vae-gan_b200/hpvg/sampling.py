"""Sample generation — the loop of the reference's `eval_video.py:53-82` (`eval`), sharded by sample index.

One "sampled clip" = one random-mode generator forward from `Z_init_size` noise through the whole pyramid
(SURVEY §8 a16).  Samples are independent in eval mode (BatchNorm uses moving statistics), so sample i goes to rank
`i mod world` and each rank batches its local samples; no collective is needed during generation."""
import os

import numpy as np

from . import ops
from .runtime import BF16, F32, Tensor, from_numpy
from .utils import images as uimg


def z_init_size(opt, batch=1):
    """eval_video.py:36-39: [1, latent_dim, td, int(s0*ar), s0]."""
    td, h, w = uimg.scale_shape(opt, 0)
    return (batch, opt.latent_dim, td, h, w)


def local_chunks(num_samples, batch, rank, world):
    """Samples are dealt to ranks in contiguous chunks of `batch`: chunk c = [c*batch, (c+1)*batch) -> rank c mod world.
    (Any partition gives identical clips because all noise is keyed by the GLOBAL sample index.)"""
    chunks = []
    for c, start in enumerate(range(0, num_samples, batch)):
        if c % world == rank:
            chunks.append(list(range(start, min(start + batch, num_samples))))
    return chunks


def host_noise_for_sample(seed, index, shape, out=None):
    """Counter-based host draw (numpy Philox-free PCG64 keyed by [seed, sample index], float32 ziggurat): z of sample
    `index` does not depend on the number of ranks, the batch size or the drawing thread.  The reference draws z on the
    host with numpy too (eval_video.py:67 -> utils.generate_noise_ref), from the global generator."""
    g = np.random.default_rng([int(seed), int(index)])
    if out is None:
        return g.standard_normal(shape, dtype=np.float32)
    g.standard_normal(dtype=np.float32, out=out)
    return out


def host_uniform_for_sample(seed, index, shape, out=None):
    """The host half of a sample's z: uniforms in [0, 1) from the same counter-based generator (PCG64 keyed by
    [seed, sample index]).  The device turns consecutive pairs into N(0, 1) in place after the upload
    (`ops.box_muller_`): the random numbers are still drawn on the host, per sample, like the reference's
    (eval_video.py:67), at a quarter of the host cost of drawing normals (numpy: 3.6 vs 14 ns per value)."""
    g = np.random.default_rng([int(seed), int(index)])
    if out is None:
        return g.random(shape, dtype=np.float32)
    g.random(dtype=np.float32, out=out)
    return out


def fold_seed(netG, seed):
    """The device Philox key of the refinement noise follows `seed` (opt.manualSeed) as well: different seeds must not
    reuse the same refinement noise per sample index."""
    base = getattr(netG, "_noise_seed_base", None)
    if base is None:
        base = netG._noise_seed_base = int(netG.noise_seed)
    x = (int(seed) + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF          # splitmix64 finaliser
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    netG.noise_seed = (base ^ x ^ (x >> 31)) & 0x7FFFFFFFFFFFFFFF


class FusedSampler:
    """The generator's random-mode forward (eval_video.py:67-76 -> networks_3d.py:406-451) as ONE C call per batch:
    `hpvg_generator_sample` (include/hpvg.h) enqueues exactly the launches `GeneratorHPVAEGAN.construct(noise_init=z,
    isRandom=True)` makes, with the same arguments — the clips are bit-identical to that path — from a description of
    the prepared network (packed filter banks, folded-BatchNorm epilogue vectors) built here once.  bf16 precision
    mode, eval-mode BatchNorm.  This is the entry a maintainer binds when the per-layer cells are not needed
    ("one call per batch")."""

    def __init__(self, netG, noise_amps, batch, stream=None, graph=False):
        """graph=True: the call is captured once per batch size into a CUDA graph (z and the clip live in buffers of the
        sampler, the sample index in a device counter) and replayed — for small batches, where the ~80 launches of a
        forward cost more than their kernels (the reference's eval loop generates one sample at a time,
        eval_video.py:62-76)."""
        import ctypes
        from ._lib import HPVG_BLOCK_LAYERS, HPVG_MAX_LEVELS, HpvgGenerator, HpvgError, lib
        if ops.cl_dtype() != BF16:
            raise HpvgError("FusedSampler: bf16 precision mode only")
        if netG.training:
            raise HpvgError("FusedSampler: the generator must be in eval mode (set_train(False))")
        opt = netG.opt
        n_stages = len(netG.body)
        if n_stages + 1 > HPVG_MAX_LEVELS:
            raise HpvgError("FusedSampler: at most %d pyramid levels" % HPVG_MAX_LEVELS)
        self.net, self.batch = netG, int(batch)
        g = HpvgGenerator()
        g.n_stages, g.nc_im, g.latent_dim = n_stages, int(opt.nc_im), int(opt.latent_dim)
        for l in range(n_stages + 1):
            g.T[l], g.H[l], g.W[l] = (int(v) for v in netG.stage_shape(l))
            add = l >= 1 and netG.noise_at(l, True)
            g.noise_amp[l] = float(noise_amps[l]) if add else 0.0
            g.noise_seed[l] = ((netG.noise_seed + 0x632BE59BD9B4E019 * l) & 0xFFFFFFFFFFFFFFFF) if add else 0
        self._keep = []                               # tensors the description points at

        def fill(dst, block):
            layers = block.layers
            if len(layers) > HPVG_BLOCK_LAYERS:
                raise HpvgError("FusedSampler: at most %d layers per block" % HPVG_BLOCK_LAYERS)
            dst.n_layers = len(layers)
            for j, layer in enumerate(layers):
                layer._prepare(False, stream)         # filter bank(s) + (scale, shift) with eval-mode BatchNorm folded in
                if layer.sn:
                    raise HpvgError("FusedSampler: spectrally normalised layers are not part of the generator")
                dst.cin[j] = int(layer.cin)
                for h, img in enumerate(layer._wimgs):
                    dst.wimg[j][h] = img.ptr
                dst.scale[j] = layer._aff.ptr
                dst.shift[j] = layer._aff.ptr + 256
                self._keep += list(layer._wimgs) + [layer._aff]

        fill(g.decoder, netG.decoder)
        for sidx in range(n_stages):
            fill(g.body[sidx], netG.body[sidx])
        self.desc = g
        nbytes = int(lib.hpvg_generator_sample_workspace(ctypes.byref(g), self.batch))
        if nbytes <= 0:
            raise HpvgError("FusedSampler: bad generator description")
        self.ws = Tensor(((nbytes + 3) // 4,), F32)
        self.ws_bytes = nbytes
        self.out_shape = (int(opt.nc_im),) + tuple(int(v) for v in netG.stage_shape(n_stages))
        self.graph = bool(graph)
        self._graphs = {}            # n -> (Graph, z buffer, clip buffer)
        if self.graph:
            from .runtime import U64
            self._offset = Tensor((1,), U64).zero_()   # device counter: first sample index of the batch being generated
            self._offset_val = 0

    def __call__(self, z, sample_base=None, out=None, vae_out=None, stream=None):
        """z: fp32 (n <= batch, latent_dim, T0, H0, W0) device tensor -> fp32 (n, nc_im, T, H, W)."""
        import ctypes
        from ._lib import check, lib
        n = int(z.shape[0])
        if n > self.batch:
            raise ValueError("FusedSampler built for batches of %d" % self.batch)
        if sample_base is None:
            sample_base = self.net.sample_counter
        if self.graph:
            return self._replay(z, n, int(sample_base), out, vae_out, stream)
        if out is None:
            out = Tensor((n,) + self.out_shape, F32)
        check(lib.hpvg_generator_sample(ctypes.byref(self.desc), z.ptr, n, int(sample_base), None, out.ptr,
                                        None if vae_out is None else vae_out.ptr, self.ws.ptr, self.ws_bytes,
                                        None if stream is None else stream.handle), "generator_sample")
        self.net.sample_counter = int(sample_base) + n
        return out

    def _replay(self, z, n, sample_base, out, vae_out, stream):
        import ctypes
        from ._lib import HpvgError, check, lib
        from .runtime import Graph
        if stream is None:
            raise HpvgError("FusedSampler(graph=True) needs an explicit stream")
        if vae_out is not None:
            raise HpvgError("FusedSampler(graph=True) does not return vae_out")
        ent = self._graphs.get(n)
        if ent is None or ent[0].stream is not stream:
            zbuf = Tensor(tuple(z.shape), F32)
            obuf = Tensor((n,) + self.out_shape, F32)
            g = Graph(stream)
            with g:
                check(lib.hpvg_generator_sample(ctypes.byref(self.desc), zbuf.ptr, n, 0, self._offset.ptr, obuf.ptr, None,
                                                self.ws.ptr, self.ws_bytes, stream.handle), "generator_sample")
            ent = self._graphs[n] = (g, zbuf, obuf)
        g, zbuf, obuf = ent
        check(lib.hpvg_d2d(zbuf.ptr, z.ptr, z.nbytes, stream.handle), "d2d")
        # advance the device counter to this batch's first sample index (a kernel argument: nothing on the host has to
        # stay unchanged until the stream gets there; uint64 wrap-around makes a step backwards an addition too)
        check(lib.hpvg_counter_add(self._offset.ptr, (sample_base - self._offset_val) & 0xFFFFFFFFFFFFFFFF,
                                   stream.handle), "counter_add")
        self._offset_val = sample_base
        g.launch()
        if out is not None:
            check(lib.hpvg_d2d(out.ptr, obuf.ptr, obuf.nbytes, stream.handle), "d2d")
        self.net.sample_counter = sample_base + n
        return out if out is not None else obuf

    def close(self):
        for g, _, _ in self._graphs.values():
            g.destroy()
        self._graphs = {}


class SamplePipeline:
    """eval_video.py:53-82 as a stream: per chunk of `batch` samples
         host draw of z's random numbers (uniforms; worker threads, straight into pinned memory)  ->  H2D on a copy
         stream  ->  in-place Box-Muller + generator forward on `stream`  ->  D2H of the clips on a second copy stream into pinned memory  ->  sink.
    `depth` z buffers and two clip buffers are in flight, so the draw / copies of chunk i+1 overlap the generation of
    chunk i.  noise="device": z is drawn by the device Philox generator keyed by (seed, sample index) instead — no host
    draw, no H2D (the clips differ from the host-noise ones; stated wherever it is used)."""

    def __init__(self, netG, noise_amps, batch, seed=0, stream=None, threads=None, noise="host", depth=3, fused=False):
        from concurrent.futures import ThreadPoolExecutor
        from .runtime import Event, PinnedBuffer, Stream
        if noise not in ("host", "device"):
            raise ValueError("noise must be 'host' or 'device'")
        self.net, self.amps, self.batch, self.seed, self.noise = netG, list(noise_amps), int(batch), int(seed), noise
        self.opt = netG.opt
        self.st = stream or Stream()
        self.cp_in, self.cp_out = Stream(), Stream()
        self.zshape = z_init_size(self.opt, self.batch)
        self.depth = int(depth)
        zbytes = int(np.prod(self.zshape)) * 4
        self.z_dev = [Tensor(self.zshape, F32) for _ in range(self.depth)]
        self.z_host = [PinnedBuffer(zbytes) for _ in range(self.depth)] if noise == "host" else None
        self.gen_done = [None] * self.depth          # device z buffer k may be overwritten once its consumer has run
        self.h2d_done = [None] * self.depth          # pinned z buffer k may be redrawn once its upload has completed
        self.out_host, self.d2h_done = [None, None], [None, None]
        self._Event, self._Pinned = Event, PinnedBuffer
        self.pool = ThreadPoolExecutor(max_workers=threads or max(1, min(8, (os.cpu_count() or 2) // 2))) \
            if noise == "host" else None
        self.h2d_bytes = self.d2h_bytes = 0
        fold_seed(netG, seed)
        # fused=True: one C call per chunk (FusedSampler) instead of the per-layer launches from Python; same clips
        self.fused = FusedSampler(netG, self.amps, self.batch, stream=self.st) if fused else None
        self.fused_out = [None, None]

    def _draw_async(self, k, chunk):
        """Fill pinned z buffer k with the draws of `chunk` on the worker threads -> list of futures."""
        arr = self.z_host[k].as_array(self.zshape)
        shp = self.zshape[1:]
        return [self.pool.submit(host_uniform_for_sample, self.seed, i, shp, arr[j]) for j, i in enumerate(chunk)]

    def run(self, chunks, sink=None):
        """chunks: list of lists of global sample indices (each <= batch long).  sink(chunk, clips_view) is called with a
        numpy VIEW of the pinned clip buffer (valid until the next-but-one chunk); returns the number of clips."""
        Event = self._Event
        st, n_done = self.st, 0
        pending = []                                  # (chunk, out slot, d2h event)
        futs = {}
        if self.noise == "host":
            for k in range(min(self.depth - 1, len(chunks))):      # prefetch
                if self.h2d_done[k % self.depth] is not None:
                    self.h2d_done[k % self.depth].sync()
                futs[k] = self._draw_async(k % self.depth, chunks[k])
        for c, chunk in enumerate(chunks):
            k, o = c % self.depth, c & 1
            n = len(chunk)
            zd = self.z_dev[k] if n == self.batch else self.z_dev[k].view((n,) + self.zshape[1:], F32, 0)
            if self.noise == "host":
                for f in futs.pop(c):
                    f.result()
                nb = n * int(np.prod(self.zshape[1:])) * 4
                if self.gen_done[k] is not None:
                    self.cp_in.wait_event(self.gen_done[k])        # device-side: the previous reader of z_dev[k] is done
                check_h2d(zd.ptr, self.z_host[k].ptr, nb, self.cp_in)
                self.h2d_bytes += nb
                ev = Event(); ev.record(self.cp_in)
                self.h2d_done[k] = ev
                st.wait_event(ev)
                ops.box_muller_(zd, stream=st)        # host uniforms -> N(0,1), in place
                nxt = c + self.depth - 1
                if nxt < len(chunks):                 # next draw goes into the pinned buffer uploaded longest ago
                    kk = nxt % self.depth
                    if self.h2d_done[kk] is not None:
                        self.h2d_done[kk].sync()
                    futs[nxt] = self._draw_async(kk, chunks[nxt])
            else:
                per = int(np.prod(self.zshape[1:]))
                for j, i in enumerate(chunk):
                    ops.randn((per,), self.seed ^ 0x5A17, offset=i, out=zd.view((per,), F32, j * per * 4), stream=st)
            if self.d2h_done[o] is not None:
                st.wait_event(self.d2h_done[o])       # clip tensor / pinned buffer o is free again
            self.net.sample_counter = chunk[0]        # device Philox noise keyed by (seed, GLOBAL sample index, element)
            self.net.out_slot = o
            if self.fused is not None:
                if self.fused_out[o] is None:
                    self.fused_out[o] = Tensor((self.batch,) + self.fused.out_shape, F32)
                xo = self.fused_out[o] if n == self.batch else self.fused_out[o].view((n,) + self.fused.out_shape, F32, 0)
                x = self.fused(zd, sample_base=chunk[0], out=xo, stream=st)
            else:
                x, _ = self.net(zd, self.amps, noise_init=zd, isRandom=True, stream=st)
            ev = Event(); ev.record(st); self.gen_done[k] = ev
            if self.out_host[o] is None or self.out_host[o].nbytes < x.nbytes:
                self.out_host[o] = self._Pinned(int(np.prod((self.batch,) + tuple(x.shape[1:]))) * 4)
            self.cp_out.wait_event(ev)
            check_d2h(self.out_host[o].ptr, x.ptr, x.nbytes, self.cp_out)
            self.d2h_bytes += x.nbytes
            ev2 = Event(); ev2.record(self.cp_out); self.d2h_done[o] = ev2
            pending.append((chunk, o, ev2, tuple(x.shape)))
            while len(pending) > 1:                   # hand the previous chunk to the sink while this one runs
                n_done += self._deliver(pending.pop(0), sink)
        while pending:
            n_done += self._deliver(pending.pop(0), sink)
        return n_done

    def _deliver(self, item, sink):
        chunk, o, ev, shape = item
        ev.sync()
        if sink is not None:
            sink(chunk, self.out_host[o].as_array(shape))
        return len(chunk)

    def close(self):
        if self.pool is not None:
            self.pool.shutdown(wait=True)
            self.pool = None
        self.net.out_slot = 0


def check_h2d(dst, src, nbytes, stream):
    from ._lib import check, lib
    check(lib.hpvg_h2d(dst, src, nbytes, stream.handle), "h2d")


def check_d2h(dst, src, nbytes, stream):
    from ._lib import check, lib
    check(lib.hpvg_d2h(dst, src, nbytes, stream.handle), "d2h")


def generate(netG, noise_amps, num_samples, rank=0, world=1, batch=8, seed=0, stream=None, keep=True, noise="host",
             threads=None, fused=False):
    """Returns (indices, clips) for this rank: clips is a float32 array (n_local, 3, T, H, W) when keep=True.
    The loop runs as a SamplePipeline (host draw / copies overlapped with the generation); fused=True issues every
    chunk's forward as one C call (hpvg_generator_sample) — the same clips."""
    chunks = local_chunks(num_samples, batch, rank, world)
    idxs = [i for c in chunks for i in c]
    outs = []
    pipe = SamplePipeline(netG, noise_amps, batch, seed=seed, stream=stream, threads=threads, noise=noise, fused=fused)
    try:
        pipe.run(chunks, (lambda chunk, clips: outs.append(np.array(clips))) if keep else None)
        pipe.st.sync()
    finally:
        pipe.close()
    return idxs, (np.concatenate(outs) if outs and keep else None)


def generate_moments(netG, noise_amps, num_samples, features, comm=None, batch=8, seed=0, stream=None, keep=False,
                     threads=None):
    """eval_video.py:53-82 + the statistics half of calculate_SVFID (fid_score.py:219-242), sharded:
    every rank generates its samples, reduces each clip ON THE DEVICE to 4160 moment floats, and ONE all-gather
    (NCCL over NVLink on GPUs) collects them.  Returns (rows (num_samples, 4160) in global sample order — identical on
    every rank —, voxels per clip, local clips or None)."""
    from . import fid
    from .dist import SingleProcess, shard_rows, unshard_rows
    comm = comm or SingleProcess()
    opt = netG.opt
    shp = z_init_size(opt, 1)[1:]
    fold_seed(netG, seed)
    idx = shard_rows(num_samples, batch, comm.rank, comm.world)
    rows = Tensor((len(idx), fid.MOMENT_FLOATS), F32).zero_(stream)
    clips = []
    count = None
    from concurrent.futures import ThreadPoolExecutor
    starts = [c0 for c0 in range(0, len(idx), batch) if any(i >= 0 for i in idx[c0:c0 + batch])]

    def draw(c0):     # the host draw of a chunk (numpy releases the GIL): runs ahead of the GPU on worker threads
        ch = [i for i in idx[c0:c0 + batch] if i >= 0]
        return ch, np.stack([host_uniform_for_sample(seed, i, shp) for i in ch])

    pool = ThreadPoolExecutor(max_workers=threads or max(1, min(8, (os.cpu_count() or 2) // 2)))
    pending = [pool.submit(draw, c0) for c0 in starts]
    for c0, fut in zip(starts, pending):
        chunk, z = fut.result()
        tz = ops.box_muller_(from_numpy(z, stream=stream), stream=stream)      # same z as SamplePipeline's
        netG.sample_counter = chunk[0]
        x, _ = netG(tz, noise_amps, noise_init=tz, isRandom=True, stream=stream)
        f = features(x, stream=stream)
        count = int(np.prod(f.shape[1:4]))
        fid.sample_moments(f, out=rows.view((len(chunk), fid.MOMENT_FLOATS), F32, c0 * fid.MOMENT_FLOATS * 4),
                           stream=stream)
        if keep:
            clips.append(x.numpy(stream))
    pool.shutdown(wait=True)
    gathered = comm.all_gather_rows(rows, stream=stream)
    all_rows = unshard_rows(gathered, num_samples, batch, comm.world)
    return all_rows, count, (np.concatenate(clips) if clips else None)


def evaluate_svfid(netG, noise_amps, real_clip, num_samples, features=None, comm=None, batch=8, seed=0, stream=None):
    """Mean per-sample Fréchet distance between generated clips and the real clip (calculate_SVFID semantics)."""
    from . import fid
    features = features or fid.RandomFeatures3D(netG.opt.nc_im, kt=netG.KT)
    rows, count, _ = generate_moments(netG, noise_amps, num_samples, features, comm, batch, seed, stream)
    real_rows = fid.sample_moments(features(real_clip, stream=stream), stream=stream).numpy(stream)
    value, per_sample = fid.svfid_from_moments(real_rows[0], rows, count)
    return value, per_sample
