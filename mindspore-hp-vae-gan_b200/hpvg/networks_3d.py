"""Host-side mirror of the reference's `src/modules/networks_3d.py` cells over libhpvg kernels.

Same class names, constructor arguments, `construct` signatures and parameter names as the reference
(ConvBlock3D :45, ConvBlock3DSN :57, FeatureExtractor :76, Encode3DVAE :89, WDiscriminator3D :170,
GeneratorHPVAEGAN :354) so the parity tests read like the reference's own smoke blocks.  Instead of MindSpore
library ops every layer is one launch of the fused tcgen05 convolution (conv + bias + folded BatchNorm / 1/sigma +
LeakyReLU / tanh + residual), activations stay channels-last bf16 between layers.

Forward only here; the training step (autodiff restated by hand) lives in `train.py`."""
import numpy as np

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_TANH
from .runtime import BF16, F32, HpvgError, Tensor, from_numpy
from .utils import images as uimg


class Workspace:
    """Reusable device buffers keyed by (tag, shape, dtype): no allocation inside the steady-state loop."""

    def __init__(self):
        self._bufs = {}

    def get(self, tag, shape, dtype):
        key = (tag, tuple(int(s) for s in shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = Tensor(shape, dtype)
            self._bufs[key] = t
        return t

    def nbytes(self):
        return sum(t.nbytes for t in self._bufs.values())


class Cell:
    """Minimal stand-in for mindspore.nn.Cell: named parameters, train/eval flag, __call__ -> construct."""

    def __init__(self):
        self.training = False
        self._cells = {}

    def _add(self, name, cell):
        self._cells[name] = cell
        return cell

    def cells(self):
        return self._cells

    def set_train(self, mode=True):
        self.training = bool(mode)
        for c in self._cells.values():
            c.set_train(mode)
        return self

    def parameters_dict(self, prefix=""):
        out = {}
        for name, c in self._cells.items():
            out.update(c.parameters_dict(prefix + name + "."))
        return out

    def load_parameters(self, params, strict=True):
        """params: name -> numpy array (the reference's checkpoint contract, src/tools/pt2ms.py:129-188)."""
        mine = self.parameters_dict()
        for k, t in mine.items():
            if k in params:
                arr = np.asarray(params[k], np.float32)
                if arr.size != t.size:
                    raise HpvgError("parameter %s: size %d != %d" % (k, arr.size, t.size))
                t.copy_from_host(arr.reshape(t.shape))
            elif strict:
                raise HpvgError("missing parameter %s" % k)
        self.invalidate()

    def invalidate(self):
        for c in self._cells.values():
            c.invalidate()

    def __call__(self, *a, **k):
        return self.construct(*a, **k)


# Kink-centred bf16 storage of the pre-BatchNorm activation in training mode (csrc/elementwise.cu: bn_center_multi_kernel).
# HPVG_CENTRED_BN=0 stores the plain conv output (round-1 behaviour; kept for the A/B numbers of DESIGN.md §5.1).
import os as _os
CENTRED_BN_STORAGE = [_os.environ.get("HPVG_CENTRED_BN", "1") != "0"]


def conv_layers_of(cell):
    """Every ConvLayer under `cell`, depth first."""
    if isinstance(cell, ConvLayer):
        return [cell]
    out = []
    for c in cell.cells().values():
        out += conv_layers_of(c)
    return out


def prepack(cells, stream=None, dgrad=True):
    """Rebuild, in ONE launch, everything the coming forward / backward derives from the weights of `cells`' layers: the
    packed forward filter banks, the data-gradient banks (transposed / mirrored) and the (1, bias) epilogue vectors.
    Equivalent to what ConvLayer._prepare / train._dgrad_wimgs do lazily, layer by layer, one launch per bank — at the
    coarse scales those ~45 small launches sit on the critical path of a launch-latency-bound iteration.  bf16 kernel
    variants only (the tf32 mode keeps the lazy path)."""
    if ops.cl_dtype() != BF16:
        return 0
    entries, centres = [], []
    seen = set()
    for cell in cells:
        for l in conv_layers_of(cell):
            if id(l) in seen or l._wimgs is not None:
                continue
            seen.add(id(l))
            w = l.p["weight"]
            imgs = []
            for kw in ops.plan_wimgs(l.cin, l.cout, BF16):
                t = ops.wimg_tensor(kw["mode"])
                imgs.append(t)
                entries.append(dict(w=w, out=t, **kw))
            l._wimgs, l._wimgs_prec = imgs, BF16
            if dgrad:
                dimgs = []
                for kw in ops.plan_wimgs(l.cout, l.cin, BF16):      # roles of Cin / Cout swapped
                    t = ops.wimg_tensor(kw["mode"])
                    dimgs.append(t)
                    entries.append(dict(w=w, out=t, transpose_flip=True, **kw))
                l._dgrad_cache = (imgs, dimgs)
            if l.centred(BF16) and l.training:
                if l._aff_center is None:
                    centres.append(l.center_entry())
            elif not l.sn and (l.cout == 64 or l.cout <= 4) and l._aff_bias is None:
                aff = Tensor((2, 64), F32)
                entries.append(dict(bias=l.p["bias"], cout=l.cout, out=aff))
                l._aff_bias = aff
    ops.pack_weights_multi(entries, stream=stream)
    ops.bn_center_multi(centres, stream=stream)
    return len(entries) + len(centres)


class ConvLayer(Cell):
    """conv3x3x3(+bias) [-> BatchNorm3d] [-> act], or the spectrally normalised conv.  One fused kernel in eval mode."""

    def __init__(self, cin, cout, bn=False, sn=False, act=None, bn_prefix="1.bn2d.", conv_prefix="0.", kt=3, rng=None):
        super().__init__()
        self.cin, self.cout, self.bn, self.sn, self.kt = cin, cout, bn, sn, kt
        self.act = {None: ACT_NONE, "lrelu": ACT_LRELU, "tanh": ACT_TANH}[act]
        self.conv_prefix, self.bn_prefix = conv_prefix, bn_prefix
        rng = rng or np.random.default_rng(0)
        wshape = (cout, cin, 3, 3, 3) if kt == 3 else (cout, cin, 3, 3)
        self.p = {}
        # weight_init=Normal(0.02, 0.0) -> N(mean 0, sigma 0.02)  (networks_3d.py:49)
        self.p["weight"] = from_numpy((rng.standard_normal(wshape) * 0.02).astype(np.float32))
        self.p["bias"] = from_numpy(np.zeros(cout, np.float32))
        if sn:   # spectral_norm.py:137-140
            k = cin * 9 * kt
            u = rng.standard_normal((cout, 1)).astype(np.float32)
            v = rng.standard_normal((k, 1)).astype(np.float32)
            self.p["weight_u"] = from_numpy(u / max(np.linalg.norm(u), 1e-6))
            self.p["weight_v"] = from_numpy(v / max(np.linalg.norm(v), 1e-6))
        if bn:   # gamma_init=Normal(0.02, 1.0)  (networks_3d.py:52)
            self.p["gamma"] = from_numpy((1.0 + rng.standard_normal(cout) * 0.02).astype(np.float32))
            self.p["beta"] = from_numpy(np.zeros(cout, np.float32))
            self.p["moving_mean"] = from_numpy(np.zeros(cout, np.float32))
            self.p["moving_variance"] = from_numpy(np.ones(cout, np.float32))
        self._wimgs = None
        self._wimgs_prec = None
        self._aff = None        # epilogue vectors used by the current forward
        self._aff_bias = None   # cached (1, bias): valid until the parameters change
        self._aff_center = None  # cached (1, bias - center) of the kink-centred training forward (bf16 mode)
        self._center = None      # (64,) the offset itself (persistent buffer, rewritten when _aff_center is rebuilt)
        self._center_used = None  # the offset of the CURRENT forward's stored y (None: not centred)
        self._aff_eval = None   # cached folded eval-mode BatchNorm: valid until parameters / moving stats change
        self._aff_sn = None     # (1/sigma, bias) written by the power-iteration kernel
        self._sigma = None      # (sigma, 1/sigma)
        self._sn_fresh = False  # the power iteration of the coming forward was already run (batched per network)

    # ---- parameters
    def parameters_dict(self, prefix=""):
        out = {}
        for k, t in self.p.items():
            pre = self.bn_prefix if k in ("gamma", "beta", "moving_mean", "moving_variance") else self.conv_prefix
            out[prefix + pre + k] = t
        return out

    def invalidate(self):
        self._wimgs = None
        self._aff = None
        self._aff_bias = None
        self._aff_center = None
        self._aff_eval = None

    def copy_from(self, other):
        for k, t in self.p.items():
            t.copy_(other.p[k])
        self.invalidate()

    # ---- forward
    def _prepare_wimgs(self, stream=None, dtype=None):
        """(Re)build the packed filter bank(s) — no side effects on the spectral-norm state.  The banks are specific
        to the activation dtype (bf16 -> kind::f16 images, fp32 -> kind::tf32 images), so a switch rebuilds them."""
        dtype = dtype or ops.cl_dtype()
        if self._wimgs is None or self._wimgs_prec != dtype:
            self._wimgs = ops.build_wimgs(self.p["weight"], self.cin, self.cout, dtype=dtype, stream=stream)
            self._wimgs_prec = dtype

    def sn_entry(self, sigma=None, aff=None, u_copy=None, v_copy=None):
        """Table entry for ops.sn_power_iter_multi.  `sigma` / `aff` default to the layer's own persistent buffers;
        the training tape passes per-pass tensors so the backward sees the values THIS forward used (Q5)."""
        if sigma is None:
            if self._sigma is None:
                self._sigma = Tensor((2,), F32)
            sigma = self._sigma
        if aff is None:
            if self._aff_sn is None:
                self._aff_sn = Tensor((2, 64), F32)
            aff = self._aff_sn
        self._cur_sigma, self._cur_aff, self._cur_u, self._cur_v = sigma, aff, u_copy, v_copy
        return {"w": self.p["weight"], "u": self.p["weight_u"], "v": self.p["weight_v"], "sigma": sigma,
                "bias": self.p["bias"], "aff": aff, "u_copy": u_copy, "v_copy": v_copy}

    def _prepare(self, training, stream, dtype=None):
        """Filter bank + the epilogue vectors (bias / folded BN / 1/sigma) for this forward."""
        self._prepare_wimgs(stream, dtype)
        if self.sn:
            # Q5: u/v advance on EVERY forward, train or eval (spectral_norm.py:146-148)
            if not self._sn_fresh:
                ops.sn_power_iter_multi([self.sn_entry()], stream=stream)
            self._sn_fresh = False
            self._aff = self._cur_aff
            self._sigma_used = self._cur_sigma
        elif self.bn and not training:
            if self._aff_eval is None:
                self._aff_eval = ops.bn_fold_eval(self.p["gamma"], self.p["beta"], self.p["moving_mean"],
                                                  self.p["moving_variance"], self.p["bias"], stream=stream)
            self._aff = self._aff_eval
        elif self.centred(dtype):
            # training-mode BatchNorm, bf16 storage of y: store y minus the estimated LeakyReLU kink (bn_center_multi)
            if self._aff_center is None:
                ops.bn_center_multi([self.center_entry()], stream=stream)
            self._aff = self._aff_center
            self._center_used = self._center
        else:
            if self._aff_bias is None:
                self._aff_bias = ops.affine_from_bias(self.p["bias"], None, stream=stream)
            self._aff = self._aff_bias
        if not (self.bn and training and self.centred(dtype)):
            self._center_used = None

    def centred(self, dtype=None):
        """Does the training-mode forward store its pre-BatchNorm output kink-centred?  (bf16 activations only: the
        tf32 mode stores y unrounded, and the eval path folds BatchNorm into the conv epilogue and stores no y.)"""
        return self.bn and self.cout == 64 and (dtype or ops.cl_dtype()) == BF16 and CENTRED_BN_STORAGE[0]

    def center_entry(self):
        """Allocate (once) the offset / epilogue-vector buffers and return this layer's row for ops.bn_center_multi."""
        if self._center is None:
            self._center = Tensor((64,), F32)
        if self._aff_center is None:
            self._aff_center = Tensor((2, 64), F32)
        return (self.p["gamma"], self.p["beta"], self.p["moving_mean"], self.p["moving_variance"], self.p["bias"],
                self._center, self._aff_center)

    def forward_cl(self, x_cl, residual=None, out=None, ws=None, tag="", stream=None, saved=None, stats=None,
                   raw=None):
        """x_cl: bf16 channels-last.  Returns bf16 cl (Cout >= 64) or fp32 ncdhw (Cout <= 4).
        Training-mode BatchNorm: conv(+bias) with the batch statistics accumulated in its epilogue, then ONE
        normalise+activation pass.  `stats`: pre-zeroed fp64 (2,64) scratch (allocated here when absent);
        `saved`: dict that receives raw=y and bn=(scale, shift, mean, invstd) for the backward."""
        training = self.training
        self._prepare(training, stream, x_cl.dtype)
        if self.bn and training:
            N, T, H, W, _ = x_cl.shape
            if raw is None:
                dt = x_cl.dtype
                raw = ws.get(tag + ".raw", (N, T, H, W, self.cout), dt) if ws else Tensor((N, T, H, W, self.cout), dt)
            if stats is None:
                stats = Tensor((2, 64), "float64").zero_(stream)
            y = ops.conv3d_cl_any(x_cl, self.p["weight"], self._aff, ACT_NONE, self.cin, self.cout, out=raw,
                                  wimgs=self._wimgs, stats=stats, stream=stream)
            sv = Tensor((4, 64), F32) if saved is not None else None
            x = ops.bn_train_fused_cl(y, stats, self.p["gamma"], self.p["beta"], self.p["moving_mean"],
                                      self.p["moving_variance"], self.act, out=out, saved=sv, stream=stream,
                                      center=self._center_used)
            self._aff_eval = None   # moving stats changed: a later eval-mode fold must be rebuilt
            if saved is not None:
                saved.update(raw=y, bn=sv)
            return x
        return ops.conv3d_cl_any(x_cl, self.p["weight"], self._aff, self.act, self.cin, self.cout,
                                 residual=residual, out=out, wimgs=self._wimgs, stream=stream)


class ReflectConvLayer(ConvLayer):
    """The `bn=False` branch of ConvBlock3DSN / ConvBlock2DSN (networks_3d.py:64-71): nn.Pad(REFLECT) by one voxel on
    every spatial (and temporal) side, then a plain bias-free convolution with pad_mode='valid' [+ activation].  Computed
    as reflect pad -> the zero-padded conv kernel on the padded tensor -> its interior (the interior of a 'same'
    convolution never touches the zero padding, so it IS the valid convolution).  Unreachable from the reference's
    drivers (only FeatureExtractor(return_linear=True) builds it): forward only.  In the SequentialCell the conv is
    cell 1 (after the Pad), hence the parameter name `1.weight`; the 3-D cell has no bias, the 2-D one has."""

    def __init__(self, cin, cout, act=None, kt=3, rng=None):
        super().__init__(cin, cout, act=act, conv_prefix="1.", kt=kt, rng=rng)
        self.has_bias = kt == 1      # networks_3d.py:70 has_bias=False ; networks_2d.py:68 has_bias=True

    def parameters_dict(self, prefix=""):
        out = {prefix + self.conv_prefix + "weight": self.p["weight"]}
        if self.has_bias:
            out[prefix + self.conv_prefix + "bias"] = self.p["bias"]
        return out

    def forward_cl(self, x_cl, residual=None, out=None, ws=None, tag="", stream=None, saved=None, stats=None, raw=None):
        if saved is not None:
            raise HpvgError("ReflectConvLayer has no backward (the branch is unreachable from the reference's trainers)")
        if x_cl.dtype != BF16 or self.cout != 64:
            raise HpvgError("ReflectConvLayer: bf16 precision mode, 64 output channels")
        N, T, H, W, _ = x_cl.shape
        pt = 1 if self.kt == 3 else 0
        xp = ops.reflect_pad_cl(x_cl, pad_t=pt, pad_hw=1, stream=stream)
        self._prepare(False, stream, x_cl.dtype)          # filter bank + (1, 0) epilogue vectors (the bias stays zero)
        yp = ops.conv3d_cl_any(xp, self.p["weight"], self._aff, self.act, self.cin, self.cout, wimgs=self._wimgs,
                               stream=stream)
        if out is None:
            out = Tensor((N, T, H, W, self.cout), BF16)
        plane = (H + 2) * (W + 2) * self.cout * 2
        for n in range(N):                                 # frames pt .. pt+T-1 of sample n, rows / columns 1 .. H / W
            src = yp.view((1, T, H + 2, W + 2, self.cout), BF16, (n * (T + 2 * pt) + pt) * plane)
            dst = out.view((1, T, H, W, self.cout), BF16, n * T * H * W * self.cout * 2)
            ops.slice_act_cl(src, h0=1, w0=1, out_hw=(H, W), out=dst, stream=stream)
        return out


def sn_prepare_batch(layers, stream=None, entries=None):
    """Run the power iteration of every spectrally normalised layer in `layers` in one launch and mark them fresh."""
    sn = [l for l in layers if l.sn]
    if not sn:
        return
    ops.sn_power_iter_multi(entries if entries is not None else [l.sn_entry() for l in sn], stream=stream)
    for l in sn:
        l._sn_fresh = True


class BnStatsSlab:
    """fp64 (2,64) accumulators for the fused BatchNorm statistics of one network pass: one slab, one memset."""

    def __init__(self, n_slots=96):
        self.n_slots = n_slots
        self.t = None
        self.i = 0

    def reset(self, stream=None):
        if self.t is None:
            self.t = Tensor((self.n_slots, 2, 64), "float64")
        self.t.zero_(stream)
        self.i = 0

    def take(self):
        if self.t is None or self.i >= self.n_slots:
            raise HpvgError("BnStatsSlab exhausted: reset() it at the start of the pass / raise n_slots")
        v = self.t.view((2, 64), "float64", self.i * 1024)
        self.i += 1
        return v


def ConvBlock3D(in_channel, out_channel, ker_size=3, padding=1, stride=1, bn=True, act="lrelu", rng=None, kt=3,
                bn_prefix="1.bn2d."):
    """networks_3d.py:45-54.  Only the 3x3x3 / stride 1 / pad 1 configuration exists on the hot path.
    (kt=1, bn_prefix="1." gives the 2-D twin ConvBlock2D, networks_2d.py:44-53.)"""
    _check_geometry(ker_size, padding, stride)
    return ConvLayer(in_channel, out_channel, bn=bn, act=act, rng=rng, kt=kt, bn_prefix=bn_prefix)


def ConvBlock3DSN(in_channel, out_channel, ker_size=3, padding=1, stride=1, bn=True, act="lrelu", rng=None, kt=3):
    """networks_3d.py:57-73: `bn=True` selects the spectrally normalised conv (there is no BatchNorm in it)."""
    _check_geometry(ker_size, padding, stride)
    if not bn:
        return ReflectConvLayer(in_channel, out_channel, act=act, rng=rng, kt=kt)
    return ConvLayer(in_channel, out_channel, sn=True, act=act, rng=rng, kt=kt)


def _check_geometry(ker_size, padding, stride):
    if ker_size != 3 or padding != 1 or int(stride) != 1:
        raise HpvgError("hpvg kernels implement ker_size=3, padding=1, stride=1 (the reference's only hot-path "
                        "configuration, train_video.py:247-250); got k=%s p=%s s=%s" % (ker_size, padding, stride))


class Sequential(Cell):
    def __init__(self, layers):
        super().__init__()
        self.layers = list(layers)
        for i, l in enumerate(self.layers):
            self._add(str(i), l)

    def append(self, layer):
        self._add(str(len(self.layers)), layer)
        self.layers.append(layer)

    def __len__(self):
        return len(self.layers)

    def __getitem__(self, i):
        return self.layers[i]


class FeatureExtractor(Sequential):
    """networks_3d.py:76-86: num_blocks+1 SN blocks."""

    def __init__(self, in_channel, out_channel, ker_size, padding, stride, num_blocks=2, return_linear=False, rng=None,
                 kt=3):
        layers = [ConvBlock3DSN(in_channel, out_channel, ker_size, padding, stride, rng=rng, kt=kt)]
        for _ in range(num_blocks - 1):
            layers.append(ConvBlock3DSN(out_channel, out_channel, ker_size, padding, stride, rng=rng, kt=kt))
        # return_linear (networks_3d.py:83-84): the last block is the bias-free reflect-padded conv without activation
        layers.append(ConvBlock3DSN(out_channel, out_channel, ker_size, padding, stride, rng=rng, kt=kt,
                                    bn=not return_linear, act=None if return_linear else "lrelu"))
        super().__init__(layers)

    def construct_cl(self, x_cl, stream=None):
        for layer in self.layers:
            x_cl = layer.forward_cl(x_cl, stream=stream)
        return x_cl


class _Wrap(Sequential):
    """A SequentialCell holding exactly one conv layer (gives the `.0.` level of the reference's names)."""


def as5d(t):
    """(N,C,H,W) -> (N,C,1,H,W) view: 2-D data is the T == 1 case of every kernel."""
    if t is None or len(t.shape) == 5:
        return t
    n, c, h, w = t.shape
    return t.view((n, c, 1, h, w))


def as4d(t):
    if t is None or len(t.shape) == 4:
        return t
    n, c, _, h, w = t.shape
    return t.view((n, c, h, w))


class Encode3DVAE(Cell):
    """networks_3d.py:89-112."""
    KT = 3

    def __init__(self, opt, out_dim=None, num_blocks=2, rng=None):
        super().__init__()
        kt = self.KT
        output_dim = opt.nfc if out_dim is None else int(out_dim)
        self._features = self._add("_features", FeatureExtractor(opt.nc_im, opt.nfc, opt.ker_size, opt.ker_size // 2,
                                                                 1, num_blocks=num_blocks, rng=rng, kt=kt))
        # the SN blocks are SequentialCells themselves: encode._features.{i}.0.weight
        for l in self._features.layers:
            l.conv_prefix = "0."
        self._mu = self._add("_mu", ConvBlock3D(opt.nfc, output_dim, opt.ker_size, opt.ker_size // 2, 1, bn=False,
                                                act=None, rng=rng, kt=kt))
        self._logvar = self._add("_logvar", ConvBlock3D(opt.nfc, output_dim, opt.ker_size, opt.ker_size // 2, 1,
                                                        bn=False, act=None, rng=rng, kt=kt))

    def construct_cl(self, x_cl, stream=None):
        f = x_cl
        sn_prepare_batch(self._features.layers, stream)
        for l in self._features.layers:
            f = l.forward_cl(f, stream=stream)
        return self._mu.forward_cl(f, stream=stream), self._logvar.forward_cl(f, stream=stream), f

    def construct(self, x, stream=None):
        nd4 = len(x.shape) == 4
        mu, logvar, _ = self.construct_cl(ops.pack_cl(as5d(x), c_pitch=ops.narrow_pitch(), stream=stream), stream=stream)
        mu, logvar = ops.unpack_cl(mu, stream=stream), ops.unpack_cl(logvar, stream=stream)
        return (as4d(mu), as4d(logvar)) if nd4 else (mu, logvar)


class WDiscriminator3D(Cell):
    """networks_3d.py:170-193: SN head, num_layer SN body blocks, plain tail conv N->1."""
    KT = 3

    def __init__(self, opt, rng=None):
        super().__init__()
        N, kt = int(opt.nfc), self.KT
        self.head = self._add("head", ConvBlock3DSN(opt.nc_im, N, opt.ker_size, opt.ker_size // 2, stride=1, rng=rng,
                                                    kt=kt))
        self.body = self._add("body", Sequential([ConvBlock3DSN(N, N, opt.ker_size, opt.ker_size // 2, stride=1,
                                                                rng=rng, kt=kt) for _ in range(opt.num_layer)]))
        self.tail = self._add("tail", ConvLayer(N, 1, conv_prefix="", rng=rng, kt=kt))

    def construct(self, x, stream=None):
        nd4 = len(x.shape) == 4
        sn_prepare_batch([self.head] + self.body.layers, stream)
        h = self.head.forward_cl(ops.pack_cl(as5d(x), c_pitch=ops.narrow_pitch(), stream=stream), stream=stream)
        for l in self.body.layers:
            h = l.forward_cl(h, stream=stream)
        out = self.tail.forward_cl(h, stream=stream)
        return as4d(out) if nd4 else out


def _make_block(cin, opt, rng, kt=3, bn_prefix="1.bn2d."):
    """decoder / body stage: ConvBlock3D(cin->N) + num_layer x ConvBlock3D(N->N) + Conv3d(N->nc_im)
    (networks_3d.py:377-381, 395-401; networks_2d.py:206-213, 226-234 with kt=1)."""
    N = int(opt.nfc)
    layers = [ConvBlock3D(cin, N, opt.ker_size, opt.padd_size, stride=1, rng=rng, kt=kt, bn_prefix=bn_prefix)]
    for _ in range(opt.num_layer):
        layers.append(ConvBlock3D(N, N, opt.ker_size, opt.padd_size, stride=1, rng=rng, kt=kt, bn_prefix=bn_prefix))
    layers.append(ConvLayer(N, opt.nc_im, conv_prefix="", rng=rng, kt=kt))   # body.{s}.6.weight
    return Sequential(layers)


class GeneratorHPVAEGAN(Cell):
    """networks_3d.py:354-451."""
    KT = 3                      # temporal filter taps (1 for the 2-D twin)
    BN_PREFIX = "1.bn2d."       # the 3-D BatchNorm wraps a 2-D one (pt2ms.py:173)
    ENCODER = Encode3DVAE

    def stage_shape(self, index):
        """(T, H, W) of pyramid level `index` (images.py:96-107)."""
        return uimg.scale_shape(self.opt, index)

    def noise_at(self, index, is_random):
        """Refinement noise is added in random mode from the first GAN level on (networks_3d.py:443-446)."""
        return bool(is_random) and self.opt.vae_levels <= index

    def __init__(self, opt, is_training=False, seed=0):
        super().__init__()
        self.opt = opt
        self.is_training = is_training           # Q2: the reference's drivers leave this False
        self.vae_levels = opt.vae_levels
        self.train_all = opt.train_all
        self.N = int(opt.nfc)
        self._rng = np.random.default_rng(seed)
        self.encode = self._add("encode", self.ENCODER(opt, out_dim=opt.latent_dim, num_blocks=opt.enc_blocks,
                                                       rng=self._rng))
        self.decoder = self._add("decoder", _make_block(opt.latent_dim, opt, self._rng, self.KT, self.BN_PREFIX))
        self.body = self._add("body", Sequential([]))
        self.ws = Workspace()
        self.bn_slab = BnStatsSlab()
        self.noise_seed = 0x9E3779B97F4A7C15   # device Philox key for internally drawn noise
        self.sample_counter = 0
        self.out_slot = 0    # which of the alternating buffers receives the final clip (lets a caller overlap the
                             # device->host copy of clip i with the generation of clip i+1)

    def init_next_stage(self):
        """networks_3d.py:393-404: first stage is freshly initialised, later stages deep-copy the previous one."""
        stage = _make_block(self.opt.nc_im, self.opt, self._rng, self.KT, self.BN_PREFIX)
        if len(self.body) > 0:
            for new, old in zip(stage.layers, self.body[-1].layers):
                new.copy_from(old)
        self.body.append(stage)
        stage.set_train(self.training)

    # ---------------------------------------------------------------------------------------------------------
    def _run_block(self, block, x_cl, residual, tag, stream, out=None):
        ws = self.ws
        N, T, H, W, _ = x_cl.shape
        h = x_cl
        for j, layer in enumerate(block.layers[:-1]):
            buf = ws.get("%s.act%d" % (tag, j & 1), (N, T, H, W, self.N), x_cl.dtype)
            stats = self.bn_slab.take() if (layer.bn and layer.training) else None
            h = layer.forward_cl(h, out=buf, ws=ws, tag=tag, stream=stream, stats=stats)
        tail = block.layers[-1]
        tail.act = ACT_TANH    # tanh(block(x) [+ up]) is fused into the tail conv's epilogue (networks_3d.py:423,450)
        return tail.forward_cl(h, residual=residual, out=out, stream=stream)

    def construct(self, video, noise_amp, noise_init=None, sample_init=None, isRandom=False, noises=None, eps=None,
                  z_pred=None, stream=None):
        """Same contract as the reference (networks_3d.py:406-432).  Extra keyword-only hooks for testing:
        `noises` {scale index: fp32 ncdhw Tensor} replaces the internally drawn refinement noise, `eps` / `z_pred`
        replace the internally drawn reparameterisation noise."""
        if sample_init is not None and len(self.body) <= sample_init[0]:
            raise HpvgError("sample_init scale %d exceeds the %d stages" % (sample_init[0], len(self.body)))
        nd4 = any(t is not None and len(t.shape) == 4 for t in (video, noise_init))
        video, noise_init, eps, z_pred = as5d(video), as5d(noise_init), as5d(eps), as5d(z_pred)
        if sample_init is not None:
            sample_init = (sample_init[0], as5d(sample_init[1]))
        if noises is not None:
            noises = {k: as5d(v) for k, v in noises.items()}
        mu = logvar = None
        if self.training:
            self.bn_slab.reset(stream)
        if noise_init is None:
            mu, logvar = self.encode.construct(video, stream=stream)
            if self.is_training:
                if eps is None:
                    eps = from_numpy(np.random.normal(size=mu.shape).astype(np.float32))   # networks_3d.py:28-30
                z_vae = ops.reparam(mu, logvar, eps, stream=stream)
            else:
                z_vae = z_pred if z_pred is not None else from_numpy(
                    np.random.normal(size=mu.shape).astype(np.float32))                     # networks_3d.py:32-34
        else:
            z_vae = noise_init
        N = z_vae.shape[0]
        z_cl = ops.pack_cl(z_vae, out=self.ws.get("z", (N,) + tuple(z_vae.shape[2:]) + (z_vae.shape[1],),
                                                  ops.cl_dtype()), stream=stream)
        vae_out = self._run_block(self.decoder, z_cl, None, "dec", stream,
                                  out=self.ws.get("vae_out", (N, self.opt.nc_im) + tuple(z_vae.shape[2:]), F32))
        if sample_init is None:
            x = self.refinement_layers(0, vae_out, noise_amp, isRandom, noises=noises, stream=stream)
        else:
            x = self.refinement_layers(sample_init[0], sample_init[1], noise_amp, isRandom, noises=noises,
                                       stream=stream)
        self.sample_counter += N
        if nd4:
            x, vae_out, mu, logvar = as4d(x), as4d(vae_out), as4d(mu), as4d(logvar)
        if noise_init is None:
            return x, vae_out, mu, logvar
        return x, vae_out

    def refinement_layers(self, start_idx, x_prev_out, noise_amp, isRandom=False, noises=None, stream=None):
        """networks_3d.py:434-451."""
        opt = self.opt
        for idx in range(start_idx, len(self.body)):
            block = self.body[idx]
            size = self.stage_shape(idx + 1)
            N = x_prev_out.shape[0]
            up = self.ws.get("up%d" % idx, (N, opt.nc_im) + size, F32)
            xin = self.ws.get("xin%d" % idx, (N,) + size + (ops.narrow_pitch(),), ops.cl_dtype())
            add_noise = self.noise_at(idx + 1, isRandom)
            noise_t, seed, amp = None, 0, 0.0
            if add_noise:
                amp = float(noise_amp[idx + 1])
                if noises is not None and (idx + 1) in noises:
                    noise_t = noises[idx + 1]
                else:
                    seed = (self.noise_seed + 0x632BE59BD9B4E019 * (idx + 1)) & 0xFFFFFFFFFFFFFFFF
            ops.upsample_noise_pack(x_prev_out, size, noise=noise_t, amp=amp, seed=seed,
                                    sample_base=self.sample_counter, up=up, xin=xin, stream=stream)
            last = idx == len(self.body) - 1
            out = self.ws.get("out%d%s" % (idx, ".%d" % self.out_slot if last else ""), (N, opt.nc_im) + size, F32)
            x_prev_out = self._run_block(block, xin, up, "s%d" % idx, stream, out=out)
        return x_prev_out
