"""Single-video / single-image FID statistics.

The reference computes, PER SAMPLE, the mean and covariance over spatial positions of a 64-channel block-0 feature
map and the Fréchet distance to the real clip's statistics, then averages over samples
(`src/sinFID/fid_score.py:105-159, 160-178, 208-242`).  Its feature networks cannot be used: `C3D` is non-functional
as shipped (`c3d.py:59-96`) and both need `mindspore_hub` downloads.  Here:

  * the feature map is any 64-channel bf16 channels-last tensor; `RandomFeatures3D` is a fixed, seeded 3->64 conv +
    LeakyReLU on the tcgen05 conv kernel (a random-projection stand-in, stated as such in DESIGN.md);
  * per-sample moments are reduced ON THE DEVICE: mu = column sums (`hpvg_colsum_cl`), second moments = the centre tap
    of the tcgen05 weight-gradient kernel applied to (f, f) — sum_v f[v] f[v]^T is exactly that 64x64xV GEMM;
  * ranks exchange only the 4160 floats per sample (`hpvg/dist.py`), rank 0 evaluates the Fréchet distance on the host
    with the reference's formula."""
import numpy as np

from . import ops
from .networks_3d import ConvLayer, as5d
from .runtime import BF16, F32, Tensor

FEATURE_DIM = 64
MOMENT_FLOATS = FEATURE_DIM + FEATURE_DIM * FEATURE_DIM     # 4160 floats per sample


class RandomFeatures3D:
    """Seeded 3 -> 64 conv3x3x3 + LeakyReLU feature extractor (stand-in for the reference's block-0 features)."""

    def __init__(self, nc_im=3, seed=1234, kt=3):
        self.layer = ConvLayer(nc_im, FEATURE_DIM, act="lrelu", rng=np.random.default_rng(seed), kt=kt)
        # N(0, 0.02) filters give tiny activations; rescale so the moments are O(1)
        w = self.layer.p["weight"].numpy() * 10.0
        self.layer.p["weight"].copy_from_host(w)
        self.layer.invalidate()

    def weights(self):
        return self.layer.p["weight"].numpy(), self.layer.p["bias"].numpy()

    def __call__(self, x, stream=None):
        """x: fp32 (N,3,T,H,W) or (N,3,H,W) -> bf16 channels-last (N,T,H,W,64)."""
        return self.layer.forward_cl(ops.pack_cl(as5d(x), c_pitch=8, stream=stream, dtype=BF16), stream=stream)


def sample_moments(feat_cl, out=None, stream=None):
    """feat_cl: bf16 (N,T,H,W,64).  Returns fp32 (N, 4160): per sample [sum_v f (64) | sum_v f f^T (64x64)] — RAW sums;
    `moments_to_stats` turns them into (mu, unbiased covariance) like np.mean / np.cov(rowvar=False)."""
    N, T, H, W, C = feat_cl.shape
    assert C == FEATURE_DIM
    if out is None:
        out = Tensor((N, MOMENT_FLOATS), F32)
    per = T * H * W * C * 2
    scratch = Tensor((FEATURE_DIM, FEATURE_DIM, 3, 3, 3), F32)
    for n in range(N):
        f = feat_cl.view((1, T, H, W, C), BF16, n * per)
        row = n * MOMENT_FLOATS * 4
        ops.colsum_cl(f, out.view((FEATURE_DIM,), F32, row), accumulate=False, stream=stream)
        ops.conv_wgrad_cl(f, f, scratch, accumulate=False, stream=stream)
        # centre tap (1,1,1) of dW[co][ci][3][3][3] = sum_v f[v][co] * f[v][ci]
        ops.gather_tap(scratch, 13, out.view((FEATURE_DIM, FEATURE_DIM), F32, row + FEATURE_DIM * 4), stream=stream)
    return out


def moments_to_stats(row, count):
    """(sum f, sum f f^T, count) -> (mu, sigma) with np.cov's unbiased normalisation (fid_score.py:176-177)."""
    row = np.asarray(row, np.float64)
    s1, s2 = row[:FEATURE_DIM], row[FEATURE_DIM:].reshape(FEATURE_DIM, FEATURE_DIM)
    mu = s1 / count
    sigma = (s2 - count * np.outer(mu, mu)) / (count - 1)
    return mu, 0.5 * (sigma + sigma.T)


def _sqrtm_psd_product(s1, s2):
    """Matrix square root of s1 @ s2 (both symmetric PSD) through scipy when present, else an eigen route."""
    try:
        from scipy import linalg
        r = linalg.sqrtm(s1.dot(s2), disp=False)
        return r[0] if isinstance(r, tuple) else r
    except Exception:
        w, v = np.linalg.eigh(s1)
        root = (v * np.sqrt(np.clip(w, 0, None))) @ v.T
        m = root @ s2 @ root
        w2, _ = np.linalg.eigh(0.5 * (m + m.T))
        return np.diag(np.sqrt(np.clip(w2, 0, None)))     # same trace as sqrtm(s1 s2)


def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """fid_score.py:105-159: ||mu1-mu2||^2 + Tr(C1 + C2 - 2 sqrt(C1 C2)), with the singular-product fallback."""
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    if mu1.shape != mu2.shape or sigma1.shape != sigma2.shape:
        raise ValueError("mean / covariance shapes differ")
    diff = mu1 - mu2
    covmean = _sqrtm_psd_product(sigma1, sigma2)
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = _sqrtm_psd_product(sigma1 + offset, sigma2 + offset)
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


def svfid_from_moments(real_row, fake_rows, count):
    """calculate_SVFID (fid_score.py:219-242): mean over samples of the per-sample Fréchet distance to the real clip."""
    m1, s1 = moments_to_stats(real_row, count)
    vals = []
    for row in np.asarray(fake_rows):
        m2, s2 = moments_to_stats(row, count)
        vals.append(calculate_frechet_distance(m1, s1, m2, s2))
    return float(np.asarray(vals, np.float32).mean()), vals
