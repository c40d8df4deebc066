"""Single-video / single-image FID: feature networks, per-sample
statistics and the Fréchet distance of `src/sinFID/`.

The reference computes, PER SAMPLE, the mean and covariance over spatial positions of a 64-channel block-0 feature
map and the Fréchet distance to the real clip's statistics, then averages over samples
(`src/sinFID/fid_score.py:105-159, 160-178, 208-242`).  Its feature networks are `mindspore_hub` downloads
(`c3d.py:59-60`, `inception.py:62-64`) and `C3D` is non-functional as shipped (it loads the Inception model and then
reads `.conv1` from it).  What CAN be built offline is built here:

  * `C3DBlock0` — block 0 of C3D (`c3d.py:62-66`: `conv1` = Conv3d(3, 64, 3x3x3, pad 1) with bias, no activation inside
    the block) on the tcgen05 head-conv kernel;
  * `InceptionBlock0` — block 0 of InceptionV3 (`inception.py:66-72`: Conv2d_1a 3->32 k3 s2, Conv2d_2a 32->32 k3,
    Conv2d_2b 32->64 k3 pad 1, each conv(no bias) + BatchNorm(eps 1e-3) + ReLU) on the same conv kernels (stride-1
    zero-padded convolutions followed by `hpvg_slice_act_cl`: window / stride-2 pick + ReLU);
  both load their weights from an `.npz` (`load_npz`: MindSpore-hub / torchvision key names accepted); without a file
  they are seeded random-init networks of the SAME architecture — stated wherever a number is reported;
  * `RandomFeatures3D` — the round-1 random-projection stand-in (3->64 conv + LeakyReLU), kept for its tests;
  * per-sample moments are reduced ON THE DEVICE: mu = column sums (`hpvg_colsum_cl`), second moments = the centre tap
    of the tcgen05 weight-gradient kernel applied to (f, f) — sum_v f[v] f[v]^T is exactly that 64x64xV GEMM;
  * ranks exchange only the 4160 floats per sample (`hpvg/dist.py`), rank 0 evaluates the Fréchet distance on the host
    with the reference's formula."""
import numpy as np

from . import ops
from .networks_3d import ConvLayer, as5d
from .runtime import BF16, F32, HpvgError, Tensor, from_numpy

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


def _first(d, names):
    for n in names:
        if n in d:
            return np.asarray(d[n], np.float32)
    raise HpvgError("feature weights: none of %s in the file (has: %s ...)" % (names, sorted(d)[:6]))


def _two_x_minus_one(x, stream=None):
    """(0, 1) -> (-1, 1) (c3d.py:129-130, inception.py:141-142): lerp(x, 1, 2) = 2*x + (1 - 2)*1."""
    ones = ops.fill(Tensor(x.shape, F32), 1.0, stream)
    return ops.lerp(x, ones, 2.0, stream=stream)


class C3DBlock0:
    """Block 0 of the reference's C3D feature network (`src/sinFID/c3d.py:62-66`): conv1 = Conv3d(3, 64, kernel 3x3x3,
    padding 1, bias) — 64 feature channels at every voxel of the clip (BLOCK_INDEX_BY_DIM[64] = 0, c3d.py:15).
    `normalize_input=True` mirrors c3d.py:129-130 for inputs in (0, 1); generated clips are already in (-1, 1), which is
    what the network sees after that normalisation, so the default feeds them unchanged."""

    def __init__(self, nc_im=3, weights=None, seed=1234, normalize_input=False):
        rng = np.random.default_rng(seed)
        self.layer = ConvLayer(nc_im, FEATURE_DIM, act=None, rng=rng, kt=3)
        self.normalize_input = bool(normalize_input)
        self.pretrained = False
        if weights is None:
            # random-init stand-in of the same architecture (He-normal, fan-in 81): O(1) features for inputs in (-1, 1)
            w = rng.standard_normal((FEATURE_DIM, nc_im, 3, 3, 3)).astype(np.float32) * np.float32(np.sqrt(2.0 / 81.0))
            b = (rng.standard_normal(FEATURE_DIM) * 0.1).astype(np.float32)
            self.set_weights(w, b)
        else:
            self.load_npz(weights)

    def set_weights(self, w, b):
        if tuple(w.shape) != (FEATURE_DIM, self.layer.cin, 3, 3, 3) or tuple(b.shape) != (FEATURE_DIM,):
            raise HpvgError("C3D conv1 weights must be (64, %d, 3, 3, 3) / (64,)" % self.layer.cin)
        self.layer.p["weight"].copy_from_host(np.ascontiguousarray(w, np.float32))
        self.layer.p["bias"].copy_from_host(np.ascontiguousarray(b, np.float32))
        self.layer.invalidate()

    def load_npz(self, path):
        """conv1 weights from an .npz: MindSpore C3D names (`conv1.weight`, `conv1.bias`) or a bare (`weight`, `bias`)."""
        d = dict(np.load(path)) if not isinstance(path, dict) else path
        self.set_weights(_first(d, ["conv1.weight", "c3d.conv1.weight", "weight"]),
                         _first(d, ["conv1.bias", "c3d.conv1.bias", "bias"]))
        self.pretrained = True

    def weights(self):
        return self.layer.p["weight"].numpy(), self.layer.p["bias"].numpy()

    def __call__(self, x, stream=None):
        """x: fp32 (N,3,T,H,W) -> bf16 channels-last (N,T,H,W,64)."""
        x = as5d(x)
        if self.normalize_input:
            x = _two_x_minus_one(x, stream)
        return self.layer.forward_cl(ops.pack_cl(x, c_pitch=8, stream=stream, dtype=BF16), stream=stream)


class InceptionBlock0:
    """Block 0 of the reference's InceptionV3 feature network (`src/sinFID/inception.py:66-72`): Conv2d_1a_3x3
    (3 -> 32, stride 2, no padding), Conv2d_2a_3x3 (32 -> 32, no padding), Conv2d_2b_3x3 (32 -> 64, padding 1); each is
    conv(no bias) + BatchNorm(eps 1e-3, eval) + ReLU.  Output: 64 channels at ((H-3)//2+1-2) x ((W-3)//2+1-2) positions
    (BLOCK_INDEX_BY_DIM[64] = 0, inception.py:15).  The convolutions run on the stride-1 zero-padded tcgen05 kernels with
    the channel counts zero-padded to the kernels' 64; a "valid" output is the interior of the padded one and a stride-2
    output its odd positions, both picked (with the ReLU) by `hpvg_slice_act_cl`."""
    BN_EPS = 1e-3
    SPEC = (("Conv2d_1a", 3, 32), ("Conv2d_2a", 32, 32), ("Conv2d_2b", 32, 64))

    def __init__(self, weights=None, seed=4321, normalize_input=False):
        rng = np.random.default_rng(seed)
        self.normalize_input = bool(normalize_input)
        self.pretrained = False
        self.layers = [ConvLayer(3, 64, act=None, rng=rng, kt=1), ConvLayer(64, 64, act=None, rng=rng, kt=1),
                       ConvLayer(64, 64, act=None, rng=rng, kt=1)]
        self._aff = [None] * 3
        if weights is None:
            params = {}
            for name, cin, cout in self.SPEC:
                params[name + ".conv.weight"] = (rng.standard_normal((cout, cin, 3, 3)) *
                                                 np.sqrt(2.0 / (9 * cin))).astype(np.float32)
                params[name + ".bn.gamma"] = (1.0 + 0.1 * rng.standard_normal(cout)).astype(np.float32)
                params[name + ".bn.beta"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
                params[name + ".bn.moving_mean"] = (0.1 * rng.standard_normal(cout)).astype(np.float32)
                params[name + ".bn.moving_variance"] = (1.0 + 0.1 * rng.random(cout)).astype(np.float32)
            self.set_params(params)
        else:
            self.load_npz(weights)

    def set_params(self, params):
        """params: `<layer>.conv.weight`, `<layer>.bn.{gamma,beta,moving_mean,moving_variance}` for the three layers."""
        self.params = {k: np.asarray(v, np.float32) for k, v in params.items()}
        for li, (name, cin, cout) in enumerate(self.SPEC):
            w = self.params[name + ".conv.weight"]
            if tuple(w.shape) != (cout, cin, 3, 3):
                raise HpvgError("%s.conv.weight must be %s" % (name, (cout, cin, 3, 3)))
            layer = self.layers[li]
            full = np.zeros((64, layer.cin, 3, 3), np.float32)          # zero rows / columns beyond the real channels
            full[:cout, :cin] = w
            layer.p["weight"].copy_from_host(full)
            layer.invalidate()
            g, b = self.params[name + ".bn.gamma"], self.params[name + ".bn.beta"]
            m, v = self.params[name + ".bn.moving_mean"], self.params[name + ".bn.moving_variance"]
            scale = np.zeros(64, np.float32)
            shift = np.zeros(64, np.float32)
            scale[:cout] = g / np.sqrt(v + np.float32(self.BN_EPS))
            shift[:cout] = b - m * scale[:cout]
            self._aff[li] = from_numpy(np.stack([scale, shift]))

    def load_npz(self, path):
        """MindSpore-hub names (`Conv2d_1a.conv.weight`, `.bn.gamma/beta/moving_mean/moving_variance`) or torchvision
        names (`Conv2d_1a_3x3.conv.weight`, `.bn.weight/bias/running_mean/running_var`)."""
        d = dict(np.load(path)) if not isinstance(path, dict) else path
        params = {}
        for name, _, _ in self.SPEC:
            alts = [name, name + "_3x3", "inception." + name]
            params[name + ".conv.weight"] = _first(d, [a + ".conv.weight" for a in alts])
            for ours, theirs in (("gamma", ("gamma", "weight")), ("beta", ("beta", "bias")),
                                 ("moving_mean", ("moving_mean", "running_mean")),
                                 ("moving_variance", ("moving_variance", "running_var"))):
                params[name + ".bn." + ours] = _first(d, [a + ".bn." + t for a in alts for t in theirs])
        self.set_params(params)
        self.pretrained = True

    @staticmethod
    def out_hw(H, W):
        return ((H - 3) // 2 + 1 - 2, (W - 3) // 2 + 1 - 2)

    def __call__(self, x, stream=None):
        """x: fp32 (N,3,H,W) (or (N,3,1,H,W)) -> bf16 channels-last (N,1,Ho,Wo,64)."""
        x = as5d(x)
        if self.normalize_input:
            x = _two_x_minus_one(x, stream)
        N, _, T, H, W = x.shape
        h1, w1 = (H - 3) // 2 + 1, (W - 3) // 2 + 1
        l0, l1, l2 = self.layers
        y = self._conv(l0, ops.pack_cl(x, c_pitch=8, stream=stream, dtype=BF16), 0, stream)
        y = ops.slice_act_cl(y, 1, 1, 2, 2, (h1, w1), relu=True, stream=stream)        # stride 2, no padding
        y = self._conv(l1, y, 1, stream)
        y = ops.slice_act_cl(y, 1, 1, 1, 1, (h1 - 2, w1 - 2), relu=True, stream=stream)  # no padding
        y = self._conv(l2, y, 2, stream)
        return ops.slice_act_cl(y, relu=True, stream=stream)

    def _conv(self, layer, x_cl, li, stream):
        layer._prepare_wimgs(stream, BF16)
        return ops.conv3d_cl_any(x_cl, layer.p["weight"], self._aff[li], ops.ACT_NONE, layer.cin, 64,
                                 wimgs=layer._wimgs, stream=stream)


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
        try:
            r = linalg.sqrtm(s1.dot(s2), disp=False)
        except TypeError:       # scipy >= 1.16 dropped `disp`
            r = linalg.sqrtm(s1.dot(s2))
        return r[0] if isinstance(r, tuple) else r
    except ImportError:
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
