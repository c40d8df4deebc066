"""Sample generation — the loop of the reference's `eval_video.py:53-82` (`eval`), sharded by sample index.

One "sampled clip" = one random-mode generator forward from `Z_init_size` noise through the whole pyramid
(SURVEY §8 a16).  Samples are independent in eval mode (BatchNorm uses moving statistics), so sample i goes to rank
`i mod world` and each rank batches its local samples; no collective is needed during generation."""
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


def host_noise_for_sample(seed, index, shape):
    """Counter-based host draw: z of sample `index` does not depend on the number of ranks."""
    return np.random.default_rng([int(seed), int(index)]).standard_normal(shape).astype(np.float32)


def generate(netG, noise_amps, num_samples, rank=0, world=1, batch=8, seed=0, stream=None, keep=True):
    """Returns (indices, clips) for this rank: clips is a float32 array (n_local, 3, T, H, W) when keep=True."""
    opt = netG.opt
    shp = z_init_size(opt, 1)[1:]
    idxs, outs = [], []
    for chunk in local_chunks(num_samples, batch, rank, world):
        z = np.stack([host_noise_for_sample(seed, i, shp) for i in chunk])
        tz = from_numpy(z, stream=stream)
        netG.sample_counter = chunk[0]       # device Philox noise is keyed by (seed, global sample index, element)
        x, _ = netG(tz, noise_amps, noise_init=tz, isRandom=True, stream=stream)
        idxs += chunk
        if keep:
            outs.append(x.numpy(stream))
    return idxs, (np.concatenate(outs) if outs and keep else None)


def generate_moments(netG, noise_amps, num_samples, features, comm=None, batch=8, seed=0, stream=None, keep=False):
    """eval_video.py:53-82 + the statistics half of calculate_SVFID (fid_score.py:219-242), sharded:
    every rank generates its samples, reduces each clip ON THE DEVICE to 4160 moment floats, and ONE all-gather
    (NCCL over NVLink on GPUs) collects them.  Returns (rows (num_samples, 4160) in global sample order — identical on
    every rank —, voxels per clip, local clips or None)."""
    from . import fid
    from .dist import SingleProcess, shard_rows, unshard_rows
    comm = comm or SingleProcess()
    opt = netG.opt
    shp = z_init_size(opt, 1)[1:]
    idx = shard_rows(num_samples, batch, comm.rank, comm.world)
    rows = Tensor((len(idx), fid.MOMENT_FLOATS), F32).zero_(stream)
    clips = []
    count = None
    for c0 in range(0, len(idx), batch):
        chunk = [i for i in idx[c0:c0 + batch] if i >= 0]
        if not chunk:
            continue
        z = np.stack([host_noise_for_sample(seed, i, shp) for i in chunk])
        tz = from_numpy(z, stream=stream)
        netG.sample_counter = chunk[0]
        x, _ = netG(tz, noise_amps, noise_init=tz, isRandom=True, stream=stream)
        f = features(x, stream=stream)
        count = int(np.prod(f.shape[1:4]))
        fid.sample_moments(f, out=rows.view((len(chunk), fid.MOMENT_FLOATS), F32, c0 * fid.MOMENT_FLOATS * 4),
                           stream=stream)
        if keep:
            clips.append(x.numpy(stream))
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
