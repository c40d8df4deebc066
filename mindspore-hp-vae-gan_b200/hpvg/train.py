"""Training step of HP-VAE-GAN on libhpvg kernels — mirror of the reference's `src/modules/losses.py`
(DWithLoss :17, GWithLoss :59), `src/modules/optimizers.py` (ClippedAdam :33) and the way `train_video.py:65-113`
wires them through `nn.TrainOneStepCell`.

MindSpore's autodiff is restated by hand: every backward pass below is an explicit composition of the data-gradient
convolution (the forward tcgen05 kernel with a transposed / mirrored filter bank), the tcgen05 weight-gradient kernel,
and fused elementwise kernels (BatchNorm / LeakyReLU / tanh backward, spectral-norm chain rule, WGAN-GP pieces).
Reference quirks are mirrored on purpose: the adversarial generator term carries no gradient
(Q1), z is pure noise unless is_training (Q2), the GP alpha is drawn once (Q3), frozen blocks still run BatchNorm in
batch-statistics mode and update their moving stats (Q4), spectral-norm u/v advance on every forward (Q5)."""
import numpy as np

from . import ops
from .networks_3d import BnStatsSlab, ConvLayer, Workspace, as5d
from .ops import ACT_LRELU, ACT_LRELU_MASK, ACT_NONE, ACT_TANH
from .runtime import F32, U64, Event, Graph, HpvgError, Tensor, device_sync, from_numpy
from .utils import images as uimg

import os as _os
# The LeakyReLU backward is fused into the producing data-gradient conv (HPVG_ACT_LRELU_MASK: the epilogue multiplies by
# LeakyReLU'(stored activation)) instead of running as a separate pass.  Round 1 measured the fused iteration 0.6 ms
# SLOWER (the conv was paced by its epilogue and the extra 128 B/voxel mask load cost more than the separate pass at
# 97 % of the copy bandwidth); with the TMA-store epilogue and the cheaper MMA-issue loop the A/B flipped: 47.2 against
# 46.4 iter/s for the GAN-phase iteration at 16x192x257 (same box, alternating runs).  HPVG_MASK_FUSION=0 restores the
# separate pass.
_MASK_FUSION = _os.environ.get("HPVG_MASK_FUSION", "1") != "0"
NON_TRAINABLE = ("weight_u", "weight_v", "moving_mean", "moving_variance")


def trainable_params(cell, prefix=""):
    """[(name, Tensor)] — mindspore Cell.trainable_params(): everything except u/v and the BN moving statistics."""
    return [(k, t) for k, t in cell.parameters_dict(prefix).items() if not k.endswith(NON_TRAINABLE)]


class LossTerms:
    """Loss value assembled on the device side without stalling the stream: every term is reduced into its own slot of
    a small device vector; `finish()` enqueues ONE async copy into pinned host memory and returns a LazyLoss whose
    float() waits for that copy only when the caller actually looks at the number (the reference's loop reads losses
    just for logging, train_video.py:188-194).  A ring of host slots keeps earlier iterations' values readable."""
    SLOTS = 8
    RING = 16

    def __init__(self):
        from .runtime import PinnedBuffer
        self.dev = Tensor((self.SLOTS,), F32)
        self.host = PinnedBuffer(self.SLOTS * 4 * self.RING)
        self.coefs = []
        self._pending = [None] * self.RING
        self._next = 0

    def slot(self, coef):
        i = len(self.coefs)
        if i >= self.SLOTS:
            raise HpvgError("LossTerms: too many terms")
        self.coefs.append(float(coef))
        return self.dev.view((1,), F32, 4 * i)

    def begin(self):
        self.coefs = []

    def finish(self, stream=None):
        from ._lib import check, lib
        from .runtime import Event, _s
        k = self._next
        self._next = (k + 1) % self.RING
        if self._pending[k] is not None:
            float(self._pending[k])          # materialise the value that lived in this host slot
        check(lib.hpvg_d2h(self.host.ptr + k * self.SLOTS * 4, self.dev.ptr, self.dev.nbytes, _s(stream)), "d2h")
        ev = Event()
        ev.record(stream)
        lazy = LazyLoss(self.host, k * self.SLOTS, list(self.coefs), ev)
        self._pending[k] = lazy
        return lazy


class LazyLoss:
    def __init__(self, host, first, coefs, event):
        self._host, self._first, self._coefs, self._event, self._value = host, first, coefs, event, None

    def __float__(self):
        if self._value is None:
            self._event.sync()
            vals = self._host.as_array((LossTerms.SLOTS * LossTerms.RING,))[self._first:self._first + len(self._coefs)]
            self._value = float(sum(c * float(v) for c, v in zip(self._coefs, vals)))
            self._host = self._event = None
        return self._value

    def __repr__(self):
        return "LazyLoss(%r)" % float(self)


class GradBook:
    """name-less gradient store keyed by the parameter Tensor's device pointer."""

    def __init__(self):
        self._g = {}

    def of(self, param):
        g = self._g.get(param.ptr)
        if g is None:
            g = Tensor(param.shape, F32).zero_()
            device_sync()       # first use only: the memset ran on the legacy stream, consumers use their own stream
            self._g[param.ptr] = g
        return g

    def has(self, param):
        return param.ptr in self._g

    def zero(self, stream=None):
        for g in self._g.values():
            g.zero_(stream)


# ================================================================================================ layer level
_ZERO64 = {}


def _unit_affine(stream=None):
    """(scale=1, shift=0) epilogue vectors."""
    t = _ZERO64.get("unit")
    if t is None:
        t = from_numpy(np.concatenate([np.ones(64, np.float32), np.zeros(64, np.float32)]).reshape(2, 64))
        _ZERO64["unit"] = t
    return t


def layer_forward_train(layer, x_cl, ws, key, stream=None, slab=None, defer=None):
    """Forward of one ConvLayer keeping what its backward needs.  Returns (output, ctx).
    defer: list that receives (saved, moving_mean, moving_var) INSTEAD of updating the moving statistics in the BatchNorm
    kernel — used when several forwards of the network run concurrently (see GraphedIteration)."""
    N, T, H, W, _ = x_cl.shape
    layer._prepare(True, stream, x_cl.dtype)
    ctx = {"x": x_cl, "layer": layer}
    if layer.bn:
        # conv(+bias) with the batch statistics accumulated in its epilogue, then ONE normalise+LeakyReLU pass
        y = ws.get(key + ".y", (N, T, H, W, layer.cout), x_cl.dtype)
        stats = slab.take() if slab is not None else Tensor((2, 64), "float64").zero_(stream)
        ops.conv3d_cl_any(x_cl, layer.p["weight"], layer._aff, ACT_NONE, layer.cin, layer.cout, out=y,
                          wimgs=layer._wimgs, stats=stats, stream=stream)
        a = ws.get(key + ".a", (N, T, H, W, layer.cout), x_cl.dtype)
        saved = ws.get(key + ".saved", (4, 64), F32)
        cen = layer._center_used       # y is stored minus this offset (kink-centred bf16 storage), or None
        if defer is None:
            ops.bn_train_fused_cl(y, stats, layer.p["gamma"], layer.p["beta"], layer.p["moving_mean"],
                                  layer.p["moving_variance"], layer.act, out=a, saved=saved, stream=stream, center=cen)
        else:
            ops.bn_train_fused_cl(y, stats, layer.p["gamma"], layer.p["beta"], None, None, layer.act, out=a, saved=saved,
                                  stream=stream, center=cen)
            defer.append((saved, layer.p["moving_mean"], layer.p["moving_variance"], cen))
        layer._aff_eval = None
        ctx.update(y=y, a=a, saved=saved)
        return a, ctx
    if layer.sn:
        # the chain rule needs the (sigma, u, v) used by THIS forward (Q5): either the per-pass snapshot tensors the
        # batched power iteration wrote (sn_tape_entries), or explicit copies
        if layer._cur_u is not None:
            ctx.update(sigma=layer._cur_sigma, u=layer._cur_u, v=layer._cur_v, aff=layer._cur_aff)
        else:
            sig = Tensor((2,), F32).copy_(layer._cur_sigma, stream)
            u = Tensor(layer.p["weight_u"].shape, F32).copy_(layer.p["weight_u"], stream)
            v = Tensor(layer.p["weight_v"].shape, F32).copy_(layer.p["weight_v"], stream)
            aff = Tensor(layer._aff.shape, F32).copy_(layer._aff, stream)
            ctx.update(sigma=sig, u=u, v=v, aff=aff)
            layer._aff = aff
    if layer.cout <= 4:
        raise HpvgError("tail convs are handled by the block-level code")
    a = ws.get(key + ".a", (N, T, H, W, layer.cout), x_cl.dtype)
    ops.conv3d_cl_any(x_cl, layer.p["weight"], layer._aff, layer.act, layer.cin, layer.cout, out=a,
                      wimgs=layer._wimgs, stream=stream)
    ctx.update(a=a)
    return a, ctx


def sn_tape_prepare(layers, ws, tag, stream=None):
    """Batched power iteration (one launch) for a training pass: (sigma, 1/sigma), the conv epilogue vectors and a
    snapshot of the updated u, v go to per-pass workspace tensors that the backward of this pass reads."""
    from .networks_3d import sn_prepare_batch
    sn = [l for l in layers if l.sn]
    entries = []
    for i, l in enumerate(sn):
        k = "%s.sn%d" % (tag, i)
        entries.append(l.sn_entry(sigma=ws.get(k + ".sigma", (2,), F32), aff=ws.get(k + ".aff", (2, 64), F32),
                                  u_copy=ws.get(k + ".u", l.p["weight_u"].shape, F32),
                                  v_copy=ws.get(k + ".v", l.p["weight_v"].shape, F32)))
    sn_prepare_batch(sn, stream, entries=entries)


def _dgrad_wimgs(layer, stream=None):
    """Filter banks of the data-gradient convolution of `layer` (roles of Cin/Cout swapped, taps mirrored).  Cached on
    the layer until its parameters change (ConvLayer.invalidate): the D step back-propagates through every D layer
    three times per iteration with the same weights."""
    cached = getattr(layer, "_dgrad_cache", None)
    if cached is not None and layer._wimgs is not None and cached[0] is layer._wimgs:
        return cached[1]
    # forward channels cin -> cout; the data gradient maps cout -> cin with the transposed / mirrored bank
    imgs = ops.build_wimgs(layer.p["weight"], layer.cout, layer.cin, transpose_flip=True, stream=stream)
    # keyed by the identity of the forward filter-bank list: invalidate() drops that list, so a stale entry can never
    # be returned after a weight update
    if layer._wimgs is not None:
        layer._dgrad_cache = (layer._wimgs, imgs)
    return imgs


# ------------------------------------------------------------------------------------------------ backward branches
# Inside one layer's backward the weight gradient (wgrad kernel [+ the spectral-norm chain rule]) and the data gradient
# (dgrad conv -> the layer below) only share their INPUTS.  At the coarse scales every kernel is a few microseconds and
# the iteration is one long dependency chain (BASELINE config 2: 329 launches), so the weight-gradient work is issued on
# side streams — captured as parallel branches of the iteration's CUDA graph — and the chain that remains is
# BN-backward -> dgrad per layer.  A layer always maps to the same side stream, so repeated accumulations into one
# weight gradient stay ordered; the library's reduction scratch is per stream (csrc/api.cu).  The branch set is joined
# before the optimiser reads the gradients (GWithLoss.grad).
class WgradBranches:
    def __init__(self, main, sides):
        self.main, self.sides, self._slot, self._keep, self._used = main, list(sides), {}, [], set()

    def stream_for(self, layer):
        """Side stream of `layer`, made to wait for everything enqueued on the main stream so far."""
        i = self._slot.setdefault(id(layer), len(self._slot) % len(self.sides))
        side = self.sides[i]
        ev = Event()
        ev.record(self.main)
        side.wait_event(ev)
        self._keep.append(ev)
        self._used.add(i)
        return side

    def join(self):
        for i in sorted(self._used):
            ev = Event()
            ev.record(self.sides[i])
            self.main.wait_event(ev)
            self._keep.append(ev)
        self._used = set()


_BRANCHES = [None]


def _wgrad_stream(layer, stream):
    br = _BRANCHES[0]
    if br is None or stream is not br.main:
        return stream
    return br.stream_for(layer)


def conv_backward(layer, ctx, gy_cl, grads, ws, key, need_dx=True, want_dw=True, inv_sigma_aff=None, stream=None,
                  dw_target=None, mask_input=False, dx_out=None):
    """Backward of the convolution of `layer` given gy (bf16 cl, Cout channels [zero padded to 64 for tails]).
    Accumulates dW (into dw_target or grads) and db; returns dx: bf16 cl (Cin >= 64) or fp32 ncdhw (Cin <= 4).
    mask_input: the conv's input x is the stored output of a LeakyReLU (no BatchNorm in between): the data-gradient
    kernel multiplies by LeakyReLU'(x) in its epilogue, so the result is already the gradient wrt that layer's
    PRE-activation (saves a separate elementwise pass over a finest-scale activation tensor)."""
    x_cl = ctx["x"]
    N, T, H, W, xp = x_cl.shape
    cin, cout = layer.cin, layer.cout
    w = layer.p["weight"]
    if want_dw:
        dw = dw_target if dw_target is not None else grads.of(w)
        xw = ctx.get("x_wide", x_cl)             # head convs: the 64-channel zero-padded copy of the 8-channel input
        wst = ctx.get("wg_stream") or _wgrad_stream(layer, stream)
        ctx["wg_stream"] = wst
        for ob in range(max(1, cout // 64)):
            for ib in range(max(1, cin // 64)):
                ops.conv_wgrad_cl(xw, gy_cl, dw, co_off=ob * 64, co_n=min(cout, 64), ci_off=ib * 64,
                                  ci_n=min(cin, 64), x_coff=ib * 64, gy_coff=ob * 64, accumulate=True, stream=wst)
        ctx["wg_last"] = wst          # where this layer's weight gradient was produced (the sigma chain rule follows it)
    ctx["wg_stream"] = None           # a preset is valid for one call only (contexts persist across iterations)
    if not need_dx:
        return None
    imgs = _dgrad_wimgs(layer, stream)
    aff = inv_sigma_aff if inv_sigma_aff is not None else _unit_affine(stream)
    ss = (aff.view((64,), F32, 0), _unit_affine(stream).view((64,), F32, 256))     # (1 or 1/sigma, 0)
    dt = gy_cl.dtype
    if cin <= 8:       # head layer: 64 -> (<= 4) tail-type kernel, fp32 ncdhw result
        out = ws.get(key + ".dx3", (N, cin, T, H, W), F32)
        return ops.conv3d_cl_any(gy_cl, w, None, ACT_NONE, 64, cin, out=out, wimgs=imgs, stream=stream, scale_shift=ss)
    dx = dx_out if dx_out is not None else ws.get(key + ".dx", (N, T, H, W, cin), dt)
    fuse_mask = mask_input and _MASK_FUSION and dt != F32
    if mask_input and not fuse_mask:     # default: data gradient, then the (bandwidth-bound) lrelu backward pass
        raw = ws.get(key + ".dxraw", (N, T, H, W, cin), dt)
        conv_backward(layer, ctx, gy_cl, grads, ws, key, True, False, inv_sigma_aff, stream, None, False, raw)
        return ops.lrelu_bwd_cl(raw, x_cl, out=dx, stream=stream)
    act = ACT_LRELU_MASK if fuse_mask else ACT_NONE
    mask = x_cl if fuse_mask else None
    # tail layer (cout <= 4): gy is a narrow tensor -> head-type kernel; otherwise 64-channel blocks
    return ops.conv3d_cl_any(gy_cl, w, None, act, cout if cout > 4 else gy_cl.shape[-1], cin, out=dx, wimgs=imgs,
                             mask=mask, stream=stream, scale_shift=ss)


def layer_backward(layer, ctx, ga_cl, grads, ws, key, need_dx=True, trainable=True, stream=None, ga_masked=False,
                   mask_input=False):
    """Backward of conv -> [BN] -> LeakyReLU given ga (grad wrt the layer output, bf16 cl).
    ga_masked: ga is already the gradient wrt this layer's pre-activation (the upper layer's data-gradient conv applied
    LeakyReLU' in its epilogue); mask_input: ask OUR data-gradient conv to do the same for the layer below."""
    if layer.bn:
        dg = grads.of(layer.p["gamma"]) if trainable else None
        db = grads.of(layer.p["beta"]) if trainable else None
        gy = ops.bn_bwd_cl(ga_cl, ctx["y"], ctx["saved"], layer.act, out=ws.get(key + ".gy", ga_cl.shape, ga_cl.dtype),
                           dgamma=dg, dbeta=db, accumulate=True, stream=stream)
        ctx["wg_stream"] = None
        # The bias of a convolution in front of a training-mode BatchNorm has an IDENTICALLY ZERO gradient: the
        # BatchNorm backward projects the mean out of gy, so sum_v gy[v] == 0 in exact arithmetic.  The reference's
        # autodiff sums rounding noise there (and Adam turns that noise into +-lr steps of a parameter that cannot
        # change the output); here the gradient stays the exact 0 that GradBook.zero() wrote — one pass over gy per
        # layer less (3.9 % of the finest-scale train iteration in round 1's launch list).
        return conv_backward(layer, ctx, gy, grads, ws, key, need_dx, trainable, stream=stream)
    gz = ga_cl
    if layer.act == ACT_LRELU and not ga_masked:
        gz = ops.lrelu_bwd_cl(ga_cl, ctx["a"], out=ws.get(key + ".gz", ga_cl.shape, ga_cl.dtype), stream=stream)
    ctx["wg_stream"] = None
    if trainable:
        if layer.cout == 64:
            wst = ctx["wg_stream"] = _wgrad_stream(layer, stream)
            ops.colsum_cl(gz, grads.of(layer.p["bias"]), accumulate=True, stream=wst)
        else:   # 128-channel outputs (mu / logvar): column sums per 64-channel half via an fp32 view
            _colsum_wide(gz, grads.of(layer.p["bias"]), stream)
    if layer.sn:
        # zeroed on the stream that will accumulate into it (the side branch forked above, when there is one)
        ghat = ws.get(key + ".ghat", layer.p["weight"].shape, F32).zero_(ctx.get("wg_stream") or stream)
        dx = conv_backward(layer, ctx, gz, grads, ws, key, need_dx, trainable, inv_sigma_aff=ctx["aff"],
                           stream=stream, dw_target=ghat, mask_input=mask_input)
        if trainable:     # on the stream that produced ghat (a side branch when branches are active)
            ops.sn_grad(ghat, layer.p["weight"], ctx["u"], ctx["v"], ctx["sigma"], grads.of(layer.p["weight"]),
                        accumulate=True, stream=ctx.get("wg_last") or stream)
        return dx
    return conv_backward(layer, ctx, gz, grads, ws, key, need_dx, trainable, stream=stream, mask_input=mask_input)


def _colsum_wide(g_cl, out, stream=None):
    """bias gradient of a 128-channel bf16 cl tensor: unpack to fp32 ncdhw and reduce per channel."""
    f = ops.unpack_cl(g_cl, stream=stream)
    ops.channel_sum(f, out, accumulate=True, stream=stream)


# ================================================================================================ block level
def block_forward_train(block, x_cl, residual, ws, tag, stream=None, x_wide=None, out=None, slab=None, defer=None):
    """decoder / body stage in training mode: returns (tanh(block(x) [+ residual]) fp32 ncdhw, ctxs)."""
    ctxs = []
    h = x_cl
    for j, layer in enumerate(block.layers[:-1]):
        h, c = layer_forward_train(layer, h, ws, "%s.%d" % (tag, j), stream, slab=slab, defer=defer)
        if j == 0 and x_wide is not None:
            c["x_wide"] = x_wide
        ctxs.append(c)
    tail = block.layers[-1]
    tail.act = ACT_TANH
    tail._prepare(True, stream)
    o = tail.forward_cl(h, residual=residual, out=out, stream=stream)
    ctxs.append({"x": h, "layer": tail, "out": o})
    return o, ctxs


def block_backward(block, ctxs, g_out, grads, ws, tag, need_dx, trainable=True, stream=None):
    """g_out: grad wrt the block output tanh(pre [+ up]) (fp32 ncdhw).  Returns (g_pre, dx) where g_pre is the grad wrt
    the pre-activation (== grad wrt the residual `up`) and dx the grad wrt the block input (None unless need_dx)."""
    tail_ctx = ctxs[-1]
    tail = tail_ctx["layer"]
    o = tail_ctx["out"]
    N, C, T, H, W = o.shape
    g_pre = ops.tanh_bwd(g_out, o, gpre=ws.get(tag + ".gpre", o.shape, F32), stream=stream)
    if trainable:
        ops.channel_sum(g_pre, grads.of(tail.p["bias"]), accumulate=True, stream=stream)
    gy = ops.pack_cl(g_pre, out=ws.get(tag + ".gtail", (N, T, H, W, ops.narrow_pitch()), ops.cl_dtype()), stream=stream)
    ga = conv_backward(tail, tail_ctx, gy, grads, ws, tag + ".t", True, trainable, stream=stream)
    dx = None
    for j in range(len(block.layers) - 2, -1, -1):
        last = j == 0
        ga = layer_backward(block.layers[j], ctxs[j], ga, grads, ws, "%s.%d" % (tag, j), need_dx or not last,
                            trainable, stream)
        if last:
            dx = ga
    return g_pre, dx


# ================================================================================================ generator
class GeneratorTrainer:
    """Forward/backward of GeneratorHPVAEGAN in training mode (BatchNorm batch statistics everywhere, Q4)."""

    def __init__(self, netG, device_rng=False, salt=0):
        self.net = netG
        self.ws = Workspace()
        self.slab = BnStatsSlab()
        self.salt = int(salt)        # distinct Philox keys for the D-step and G-step generator passes
        self.host_draws = 0          # forward counter (host mirror of `draws`): fresh noise on every forward
        # device_rng: every N(0,1) draw the reference makes on the host inside a forward (z when no noise_init is
        # given, networks_3d.py:28-34; refinement noise, images.py:30-37) comes from the device Philox generator keyed
        # by a DEVICE-resident draw counter, so the whole step can be replayed as a CUDA graph with fresh noise.
        self.device_rng = device_rng
        self.draws = Tensor((1,), U64).zero_() if device_rng else None
        if device_rng:
            device_sync()

    def _draw(self, shape, stream):
        """N(0,1) of `shape`: host numpy global RNG like the reference (Q7), or the device generator (device_rng)."""
        if self.device_rng:
            return ops.randn(shape, (self.net.noise_seed ^ 0x5DEECE66D) + self.salt, d_offset=self.draws,
                             out=self.ws.get("zdraw", shape, F32), stream=stream)
        return from_numpy(np.random.normal(size=shape).astype(np.float32))

    def _stage_input(self, x_prev, idx, noise_amp, is_random, noises, stream, wide):
        net, opt = self.net, self.net.opt
        size = net.stage_shape(idx + 1)
        N = x_prev.shape[0]
        up = self.ws.get("up%d" % idx, (N, opt.nc_im) + size, F32)
        xin = self.ws.get("xin%d" % idx, (N,) + size + (ops.narrow_pitch(),), ops.cl_dtype())
        add_noise = net.noise_at(idx + 1, is_random)
        noise_t, seed, amp = None, 0, 0.0
        if add_noise:
            amp = float(noise_amp[idx + 1])
            if noises is not None and (idx + 1) in noises:
                noise_t = noises[idx + 1]
            else:
                seed = (net.noise_seed + 0x632BE59BD9B4E019 * (idx + 1) +
                        0x94D049BB133111EB * self.salt) & 0xFFFFFFFFFFFFFFFF
        base = net.sample_counter + (0 if self.device_rng else self.host_draws)
        ops.upsample_noise_pack(x_prev, size, noise=noise_t, amp=amp, seed=seed, sample_base=base,
                                up=up, xin=xin, stream=stream, d_sample_offset=self.draws)
        x_wide = None
        if wide:   # the head conv's weight gradient reads the 8-channel block input itself (TMA zero-fills 8..63)
            x_wide = xin
        return up, xin, x_wide

    def prepare_shared(self, stream=None):
        """Pack the filter banks / bias vectors of every decoder and body layer on `stream`.  Forwards that run
        concurrently on other streams afterwards only READ these caches."""
        for block in [self.net.decoder] + list(self.net.body.layers):
            for layer in block.layers:
                layer._prepare(True, stream)

    def forward(self, video, noise_amp, noise_init=None, is_random=False, noises=None, eps=None, z_pred=None,
                save_from=None, save_decoder=False, save_encoder=False, stream=None, defer_bn=False, enc_stream=None):
        """networks_3d.py:406-451 in set_train() mode.  Contexts are kept for the encoder / decoder when asked and
        for body stages with index >= save_from.  Returns dict(x, vae_out, mu, logvar, ctx...).
        defer_bn: do not touch the BatchNorm moving statistics; out["bn_deferred"] lists the pending updates."""
        net, opt, ws = self.net, self.net.opt, self.ws
        out = {"body_ctx": {}, "ups": {}}
        defer = [] if defer_bn else None
        out["bn_deferred"] = defer
        mu = logvar = None
        slab = self.slab
        slab.reset(stream)
        n_draw = 1 if noise_init is None else noise_init.shape[0]
        self.host_draws += n_draw
        if self.device_rng:
            ops.counter_add(self.draws, n_draw, stream)
        if noise_init is None:
            # enc_stream: with is_training=False (the reference's drivers, Q2) z is pure noise, so the encoder pass is
            # independent of the decoder / refinement pass and may run next to it on its own stream; the caller joins
            es = stream
            if enc_stream is not None and enc_stream is not stream:
                es = enc_stream
                fork = Event()
                fork.record(stream)
                es.wait_event(fork)
                out["enc_events"] = [fork]
            out["enc_stream"] = es
            enc = net.encode
            sn_tape_prepare(enc._features.layers, ws, "enc", es)
            x_cl = ops.pack_cl(video, c_pitch=ops.narrow_pitch(), stream=es)
            ectx = []
            h = x_cl
            xw = None
            if save_encoder:
                xw = x_cl
            for i, l in enumerate(enc._features.layers):
                h, c = layer_forward_train(l, h, ws, "enc.%d" % i, es)
                if i == 0 and xw is not None:
                    c["x_wide"] = xw
                ectx.append(c)
            mu_cl, cm = layer_forward_train(enc._mu, h, ws, "enc.mu", es)
            lv_cl, cl = layer_forward_train(enc._logvar, h, ws, "enc.lv", es)
            mu, logvar = ops.unpack_cl(mu_cl, stream=es), ops.unpack_cl(lv_cl, stream=es)
            out.update(enc_ctx=ectx, mu_ctx=cm, lv_ctx=cl)
            if es is not stream and net.is_training:      # z = eps*exp(.5 logvar)+mu needs the encoder's result
                ev = Event()
                ev.record(es)
                stream.wait_event(ev)
                out["enc_events"].append(ev)
            if net.is_training:
                if eps is None:
                    eps = self._draw(mu.shape, stream)
                z = ops.reparam(mu, logvar, eps, stream=stream)
                out["eps"] = eps
            else:
                z = z_pred if z_pred is not None else self._draw(mu.shape, stream)
        else:
            z = noise_init
        N = z.shape[0]
        z_cl = ops.pack_cl(z, out=ws.get("z", (N,) + tuple(z.shape[2:]) + (z.shape[1],), ops.cl_dtype()), stream=stream)
        vae_out, dctx = block_forward_train(net.decoder, z_cl, None, ws, "dec", stream, slab=slab, defer=defer,
                                            out=ws.get("vae_out", (N, opt.nc_im) + tuple(z.shape[2:]), F32))
        out.update(vae_out=vae_out, dec_ctx=dctx, mu=mu, logvar=logvar)
        x = vae_out
        for idx in range(len(net.body)):
            keep = save_from is not None and idx >= save_from
            up, xin, xw = self._stage_input(x, idx, noise_amp, is_random, noises, stream, wide=keep)
            x, bctx = block_forward_train(net.body[idx], xin, up, ws, "s%d" % idx, stream, x_wide=xw, slab=slab,
                                          defer=defer, out=ws.get("out%d" % idx, up.shape, F32))
            if keep:
                out["body_ctx"][idx] = bctx
            out["ups"][idx] = up
        out["x"] = x
        return out


class GWithLoss:
    """losses.py:59-107.  `grad()` returns (loss value as float, GradBook)."""

    def __init__(self, opt, netD, netG, device_rng=False):
        self._netD, self._netG, self.opt = netD, netG, opt
        self.rec_weight, self.kl_weight, self.disc_loss_weight = opt.rec_weight, opt.kl_weight, opt.disc_loss_weight
        self.trainer = GeneratorTrainer(netG, device_rng, salt=1)
        self.grads = GradBook()
        self.terms = LossTerms()
        self.wgrad_sides = None      # side streams for the weight-gradient branches (set by GraphedIteration.capture)
        self.enc_side = None         # side stream for the VAE-phase encoder pass (forward, KL term, backward)

    def recon_forward_args(self, isVAE, trainable_body, train_codec=False):
        """Keyword arguments of the reconstruction forward that grad() runs (for callers that run it themselves)."""
        _, enc_bw, save_from = self._plan(isVAE, trainable_body, train_codec)
        return dict(is_random=False, save_from=save_from, save_encoder=enc_bw)

    def _plan(self, isVAE, trainable_body, train_codec):
        """Which parts of the tape the backward walks.  GAN phase (losses.py:93-103): the reconstruction loss reaches every
        stage of body[-train_depth:] (train_video.py:78-86) — the chain runs down to the lowest trainable stage and stops
        only at the reference's stop_gradient (networks_3d.py:437-438); with --train-all and fewer stages than
        train_depth (train_video.py:95-103) encode / decoder are optimised too, and the chain continues into the decoder
        (and, with is_training=True, through the reparameterised z into the encoder)."""
        nb = len(self._netG.body)
        codec_bw = bool(isVAE or train_codec)
        enc_bw = bool(isVAE or (train_codec and self._netG.is_training))
        save_from = 0 if codec_bw else (min(trainable_body) if trainable_body else nb)
        return codec_bw, enc_bw, save_from

    def grad(self, real, real_zero, noise_init, noise_amps, isVAE=False, trainable_body=(), train_codec=False,
             noises=None, z_pred=None, eps=None, stream=None, finish=True, recon_fw=None, random_x=None,
             wait_recon=None, wait_random=None):
        """recon_fw / random_x: results of the reconstruction / random-mode generator forwards when the caller already
        ran them (possibly on other streams: wait_recon / wait_random are the events to wait for)."""
        """trainable_body: indices of body stages whose parameters are optimised (train_video.py:76-105);
        train_codec: whether encode/decoder are optimised (VAE phase)."""
        net, opt, tr = self._netG, self.opt, self.trainer
        real, real_zero, noise_init, z_pred, eps = (as5d(t) for t in (real, real_zero, noise_init, z_pred, eps))
        if noises is not None:
            noises = {k: as5d(v) for k, v in noises.items()}
        g = self.grads
        g.zero(stream)
        nb = len(net.body)
        if self.wgrad_sides and stream is not None:
            _BRANCHES[0] = WgradBranches(stream, self.wgrad_sides)
        try:
            return self._grad(real, real_zero, noise_init, noise_amps, isVAE, trainable_body, train_codec, noises,
                              z_pred, eps, stream, finish, recon_fw, random_x, wait_recon, wait_random)
        finally:
            _BRANCHES[0] = None

    def _grad(self, real, real_zero, noise_init, noise_amps, isVAE, trainable_body, train_codec, noises, z_pred, eps,
              stream, finish, recon_fw, random_x, wait_recon, wait_random):
        net, opt, tr = self._netG, self.opt, self.trainer
        g = self.grads
        nb = len(net.body)
        codec_bw, enc_bw, save_from = self._plan(isVAE, trainable_body, train_codec)
        if recon_fw is None:
            fw = tr.forward(real_zero, noise_amps, is_random=False, z_pred=z_pred, eps=eps, save_from=save_from,
                            save_encoder=enc_bw, stream=stream,
                            enc_stream=self.enc_side if (isVAE and stream is not None) else None)
        else:
            fw = recon_fw
            if wait_recon is not None:
                stream.wait_event(wait_recon)
        x, vae_out = fw["x"], fw["vae_out"]
        ws = tr.ws
        n = x.size
        terms = self.terms
        terms.begin()
        ops.mse(x, real, out=terms.slot(self.rec_weight), stream=stream)
        g_x = ops.mse_grad(x, real, self.rec_weight * 2.0 / n, g=ws.get("g_x", x.shape, F32), stream=stream)
        if isVAE:
            ops.mse(vae_out, real_zero, out=terms.slot(self.rec_weight), stream=stream)
            ops.kl_criterion(fw["mu"], fw["logvar"], out=terms.slot(self.kl_weight), stream=fw.get("enc_stream", stream))
        # ---- backward through the refinement stages (networks_3d.py:434-451)
        g_cur = g_x
        lowest = save_from
        for idx in range(nb - 1, lowest - 1, -1):
            stop_here = (opt.vae_levels == idx + 1 and not opt.train_all)      # stop_gradient on the stage input
            need_prev = (not stop_here) and (codec_bw or idx > lowest)
            g_pre, dx = block_backward(net.body[idx], fw["body_ctx"][idx], g_cur, g, ws, "s%d" % idx,
                                       need_dx=need_prev, trainable=idx in trainable_body, stream=stream)
            if not need_prev:
                g_cur = None
                break
            # grad wrt up = g_pre (residual) + dx (through the head conv); then through the resize
            ops.axpby(1.0, g_pre, 1.0, dx, stream=stream)
            prev_shape = (vae_out if idx == 0 else fw["ups"][idx - 1]).shape
            g_cur = ops.resize3d_bwd(dx, prev_shape[2:], out=ws.get("g_prev%d" % idx, prev_shape, F32), stream=stream)
        g_z_cl = None
        if isVAE:
            # ---- decoder: grad = chain from the stages + rec_weight*MSE(vae_out, real_zero)
            if nb == 0 or g_cur is None:
                g_v = ops.mse_grad(vae_out, real_zero, self.rec_weight * 2.0 / vae_out.size,
                                   g=ws.get("g_v", vae_out.shape, F32), stream=stream)
                if nb == 0:   # generated IS vae_out: both MSE terms act on it
                    ops.axpby(1.0, g_x, 1.0, g_v, stream=stream)
            else:
                g_v = ops.mse_grad(vae_out, real_zero, self.rec_weight * 2.0 / vae_out.size, g=g_cur, accumulate=True,
                                   stream=stream)
            _, g_z_cl = block_backward(net.decoder, fw["dec_ctx"], g_v, g, ws, "dec", need_dx=net.is_training,
                                       trainable=train_codec, stream=stream)
        elif codec_bw and (g_cur is not None or nb == 0):
            # ---- GAN phase with encode / decoder in the optimiser (--train-all, train_video.py:95-103): only the
            # reconstruction term rec_weight*MSE(x, real) reaches the decoder, through the whole refinement chain
            _, g_z_cl = block_backward(net.decoder, fw["dec_ctx"], g_cur if nb else g_x, g, ws, "dec",
                                       need_dx=net.is_training, trainable=True, stream=stream)
        if enc_bw and (isVAE or g_z_cl is not None):
            # ---- encoder: the KL term (VAE phase) always reaches it; the reconstruction terms only through the
            # reparameterised z, i.e. only with is_training=True (Q2: the reference's drivers leave it False, then z is
            # pure noise)
            mu, lv = fw["mu"], fw["logvar"]
            es = fw.get("enc_stream", stream) or stream      # the encoder's own branch when it ran on a side stream
            if es is not stream and net.is_training:          # the reparameterisation gradient comes from the decoder
                ev = Event()
                ev.record(stream)
                es.wait_event(ev)
                fw["enc_events"].append(ev)
            gmu, glv = ops.kl_grad(mu, lv, (self.kl_weight / mu.size) if isVAE else 0.0, stream=es)
            if net.is_training:
                g_z = ops.unpack_cl(g_z_cl, stream=es)
                ops.reparam_bwd(g_z, fw["eps"], lv, gmu, glv, stream=es)
            gmu_cl, glv_cl = ops.pack_cl(gmu, stream=es), ops.pack_cl(glv, stream=es)
            enc = net.encode
            d1 = layer_backward(enc._mu, fw["mu_ctx"], gmu_cl, g, ws, "enc.mu", True, train_codec, es)
            d2 = layer_backward(enc._logvar, fw["lv_ctx"], glv_cl, g, ws, "enc.lv", True, train_codec, es)
            ga = _add_cl(d1, d2, ws, es)
            for i in range(len(enc._features.layers) - 1, -1, -1):
                ga = layer_backward(enc._features.layers[i], fw["enc_ctx"][i], ga, g, ws, "enc.%d" % i, i > 0,
                                    train_codec, es)
            if es is not stream:      # join: the loss terms and the optimiser read what the encoder branch wrote
                ev = Event()
                ev.record(es)
                stream.wait_event(ev)
                fw["enc_events"].append(ev)
                self._enc_keep = fw["enc_events"]
        if not isVAE:
            # ---- adversarial term: value only, no gradient reaches G (Q1, losses.py:93-98)
            if random_x is None:
                random_x = tr.forward(None, noise_amps, noise_init=noise_init, is_random=True, noises=noises,
                                      stream=stream)["x"]
            elif wait_random is not None:
                stream.wait_event(wait_random)
            d_out = self._netD(random_x, stream=stream)
            ops.mean(d_out, out=terms.slot(-self.disc_loss_weight), stream=stream)
        if _BRANCHES[0] is not None:
            self._branch_keep = _BRANCHES[0]      # events stay alive as long as a graph captured from this call
            _BRANCHES[0].join()                   # the optimiser reads the gradients next
        return (terms.finish(stream) if finish else terms), g


def _add_cl(a_cl, b_cl, ws, stream=None):
    """a + b for bf16 cl tensors (through fp32)."""
    fa, fb = ops.unpack_cl(a_cl, stream=stream), ops.unpack_cl(b_cl, stream=stream)
    ops.axpby(1.0, fa, 1.0, fb, stream=stream)
    return ops.pack_cl(fb, stream=stream)


# ================================================================================================ discriminator
class DWithLoss:
    """losses.py:17-56: -mean D(real) + mean D(sg(G(z))) + lambda * mean((||grad_xhat sum D(xhat)||_2 - 1)^2)."""

    def __init__(self, opt, netD, netG, alpha=None, device_rng=False):
        self._netD, self._netG, self.opt = netD, netG, opt
        self.lambda_grad = opt.lambda_grad
        self.alpha = float(np.random.uniform()) if alpha is None else float(alpha)    # Q3: drawn once (losses.py:25)
        self.trainer = GeneratorTrainer(netG, device_rng, salt=2)
        self.ws = Workspace()
        self.grads = GradBook()
        self.terms = LossTerms()

    # ---- one D forward keeping the tape
    def _forward(self, x, tag, stream):
        D, ws = self._netD, self.ws
        N = x.shape[0]
        x8 = ops.pack_cl(x, out=ws.get(tag + ".x8", (N,) + tuple(x.shape[2:]) + (ops.narrow_pitch(),), ops.cl_dtype()),
                         stream=stream)
        xw = x8       # narrow operand of the head conv's weight gradient
        ctxs = []
        sn_tape_prepare(self._layers(), ws, tag, stream)
        h, c = layer_forward_train(D.head, x8, ws, tag + ".h", stream)
        c["x_wide"] = xw
        ctxs.append(c)
        for j, l in enumerate(D.body.layers):
            h, c = layer_forward_train(l, h, ws, "%s.b%d" % (tag, j), stream)
            ctxs.append(c)
        D.tail._prepare(True, stream)
        out = D.tail.forward_cl(h, out=ws.get(tag + ".out", (N, 1) + tuple(x.shape[2:]), F32), stream=stream)
        ctxs.append({"x": h, "layer": D.tail, "out": out})
        return out, ctxs

    def _layers(self):
        D = self._netD
        return [D.head] + list(D.body.layers)

    def _backward_first_order(self, ctxs, coef, tag, stream):
        """d(coef * sum D(x))/dW for one pass."""
        D, ws, g = self._netD, self.ws, self.grads
        out = ctxs[-1]["out"]
        N, _, T, H, W = out.shape
        go = ops.fill(ws.get(tag + ".go", out.shape, F32), coef, stream)
        ops.channel_sum(go, g.of(D.tail.p["bias"]), accumulate=True, stream=stream)
        gy = ops.pack_cl(go, out=ws.get(tag + ".gyt", (N, T, H, W, ops.narrow_pitch()), ops.cl_dtype()), stream=stream)
        # every D layer is SN-conv + LeakyReLU (no BatchNorm): each data-gradient conv applies the LeakyReLU' of the
        # layer below in its epilogue, so no separate lrelu-backward pass runs
        ga = conv_backward(D.tail, ctxs[-1], gy, g, ws, tag + ".t", True, True, stream=stream, mask_input=True)
        layers = self._layers()
        for j in range(len(layers) - 1, -1, -1):
            ga = layer_backward(layers[j], ctxs[j], ga, g, ws, "%s.l%d" % (tag, j), j > 0, True, stream,
                                ga_masked=True, mask_input=j > 0)

    def _gradient_penalty(self, ctxs, tag, stream):
        """calc_gradient_penalty (losses.py:47-52) and its gradient wrt D's weights (second order)."""
        D, ws, g = self._netD, self.ws, self.grads
        layers = self._layers()
        out = ctxs[-1]["out"]
        N, _, T, H, W = out.shape
        # (1) input gradient of sum D(xhat): deltas[j] = grad wrt the pre-activation of layer j
        ones = ops.fill(ws.get(tag + ".ones", out.shape, F32), 1.0, stream)
        d_out = ops.pack_cl(ones, out=ws.get(tag + ".dout", (N, T, H, W, ops.narrow_pitch()), ops.cl_dtype()),
                            stream=stream)
        # LeakyReLU' of the layer below is applied in each data-gradient conv's epilogue: its output IS delta[j-1]
        nl = len(layers)
        dshape = ctxs[-1]["x"].shape
        deltas = [ws.get("%s.delta%d" % (tag, j), dshape, ops.cl_dtype()) for j in range(nl)]
        conv_backward(D.tail, ctxs[-1], d_out, g, ws, tag + ".gt", True, False, stream=stream, mask_input=True,
                      dx_out=deltas[nl - 1])
        ga = None
        for j in range(nl - 1, -1, -1):
            ga = conv_backward(layers[j], ctxs[j], deltas[j], g, ws, "%s.g%d" % (tag, j), True, False,
                               inv_sigma_aff=ctxs[j]["aff"], stream=stream, mask_input=j > 0,
                               dx_out=deltas[j - 1] if j > 0 else None)
        grad_x = ga                                                   # fp32 ncdhw (N, 3, T, H, W)
        Gx, gp = ops.gp_grad(grad_x, self.lambda_grad, Gout=ws.get(tag + ".G", grad_x.shape, F32),
                             gp=self.terms.slot(1.0), stream=stream)
        # (2) d GP / d W: push G forward through the SAME linear maps, masked by the LeakyReLU pattern of xhat
        xi8 = ops.pack_cl(Gx, out=ws.get(tag + ".xi8", (N, T, H, W, ops.narrow_pitch()), ops.cl_dtype()), stream=stream)
        xi_wide = xi8
        unit = _unit_affine(stream)
        zero_shift = unit.view((64,), F32, 256)
        xi = xi8
        for j, layer in enumerate(layers):
            ghat = ws.get("%s.ghat%d" % (tag, j), layer.p["weight"].shape, F32).zero_(stream)
            xw = xi_wide if j == 0 else xi
            ops.conv_wgrad_cl(xw, deltas[j], ghat, co_n=64, ci_n=min(layer.cin, 64), accumulate=True, stream=stream)
            ops.sn_grad(ghat, layer.p["weight"], ctxs[j]["u"], ctxs[j]["v"], ctxs[j]["sigma"],
                        g.of(layer.p["weight"]), accumulate=True, stream=stream)
            # eta = conv(xi; W/sigma) (no bias, no activation); xi_next = eta * LeakyReLU'(a_j)
            layer._prepare_wimgs(stream)
            so = _scale_only(ctxs[j]["aff"], zero_shift, ws, tag, j, stream)
            if _MASK_FUSION and xi.dtype != F32:   # the mask of this layer's own LeakyReLU applied in the conv epilogue
                xi = ops.conv3d_cl_any(xi, layer.p["weight"], so, ACT_LRELU_MASK, layer.cin, 64,
                                       out=ws.get("%s.xi%d" % (tag, j), deltas[j].shape, xi.dtype), wimgs=layer._wimgs,
                                       mask=ctxs[j]["a"], stream=stream)
            else:
                eta = ops.conv3d_cl_any(xi, layer.p["weight"], so, ACT_NONE, layer.cin, 64,
                                        out=ws.get("%s.eta%d" % (tag, j), deltas[j].shape, xi.dtype), wimgs=layer._wimgs,
                                        stream=stream)
                xi = ops.lrelu_bwd_cl(eta, ctxs[j]["a"], out=ws.get("%s.xi%d" % (tag, j), eta.shape, eta.dtype),
                                      stream=stream)
        ops.conv_wgrad_cl(xi, d_out, g.of(D.tail.p["weight"]), co_n=1, ci_n=64, accumulate=True, stream=stream)
        return gp

    def grad(self, real, noise_init, noise_amps, noises=None, fake=None, stream=None, finish=True, wait_fake=None):
        real, noise_init, fake = as5d(real), as5d(noise_init), as5d(fake)
        if noises is not None:
            noises = {k: as5d(v) for k, v in noises.items()}
        g = self.grads
        g.zero(stream)
        if fake is None:
            fake = self.trainer.forward(None, noise_amps, noise_init=noise_init, is_random=True, noises=noises,
                                        stream=stream)["x"]                      # stop_gradient (losses.py:29-30)
        V = real.size // real.shape[1]
        terms = self.terms
        terms.begin()
        out_r, ctx_r = self._forward(real, "R", stream)
        self._backward_first_order(ctx_r, -1.0 / V, "R", stream)
        if wait_fake is not None:       # the fake clip was generated on another stream
            stream.wait_event(wait_fake)
        out_f, ctx_f = self._forward(fake, "F", stream)
        self._backward_first_order(ctx_f, 1.0 / V, "F", stream)
        xhat = ops.lerp(real, fake, self.alpha, out=self.ws.get("xhat", real.shape, F32), stream=stream)
        out_x, ctx_x = self._forward(xhat, "X", stream)
        self._gradient_penalty(ctx_x, "X", stream)
        ops.mean(out_r, out=terms.slot(-1.0), stream=stream)
        ops.mean(out_f, out=terms.slot(1.0), stream=stream)
        return (terms.finish(stream) if finish else terms), g


def _scale_only(aff, zero_shift, ws, tag, j, stream):
    """(scale = 1/sigma, shift = 0) epilogue vectors built from a layer's (1/sigma, bias) pair."""
    t = ws.get("%s.so%d" % (tag, j), (2, 64), F32)
    t.view((64,), F32, 0).copy_(aff.view((64,), F32, 0), stream)
    t.view((64,), F32, 256).copy_(zero_shift, stream)
    return t


# ================================================================================================ optimisers
class Adam:
    """mindspore.nn.Adam over a flat list or a list of {"params": [...], "lr": x} groups (train_video.py:65,76-108)."""

    def __init__(self, params, learning_rate=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, clip=0.0, device_step=False):
        self.beta1, self.beta2, self.eps, self.clip = beta1, beta2, eps, clip
        # device_step: the 1-based step counter (bias correction) lives on the device so the update can be replayed
        # inside a CUDA graph
        self.d_step = Tensor((1,), U64).zero_() if device_step else None
        self.items = []      # (param Tensor, lr)
        if params and isinstance(params[0], dict):
            for grp in params:
                lr = grp.get("lr", learning_rate)
                for p in grp["params"]:
                    self.items.append((p[1] if isinstance(p, tuple) else p, lr))
        else:
            for p in params:
                self.items.append((p[1] if isinstance(p, tuple) else p, learning_rate))
        self.m = [Tensor(p.shape, F32).zero_() for p, _ in self.items]
        self.v = [Tensor(p.shape, F32).zero_() for p, _ in self.items]
        device_sync()           # the memsets above ran on the legacy stream
        self.step = 0

    def apply(self, grads, stream=None):
        self.step += 1
        if self.d_step is not None:
            ops.counter_add(self.d_step, 1, stream)
        ps = [p for p, _ in self.items]
        gs = [grads.of(p) for p in ps]
        ops.adam_clip_multi(ps, gs, self.m, self.v, [lr for _, lr in self.items], self.step, self.beta1, self.beta2,
                            self.eps, self.clip, stream=stream, d_step=self.d_step)


class ClippedAdam(Adam):
    """optimizers.py:33-43: per-tensor ClipByNorm(opt.grad_clip) then Adam."""

    def __init__(self, opt, params, learning_rate=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, device_step=False):
        super().__init__(params, learning_rate, beta1, beta2, eps, clip=float(opt.grad_clip), device_step=device_step)


class TrainOneStepCell:
    """nn.TrainOneStepCell(loss_cell, optimizer): forward, backward, optimiser step; returns the loss."""

    def __init__(self, network, optimizer, cells_to_invalidate=()):
        self.network, self.optimizer = network, optimizer
        self.cells = list(cells_to_invalidate)

    def set_train(self, mode=True):
        for c in (self.network._netG, self.network._netD):
            c.set_train(mode)
        return self

    def __call__(self, *args, **kw):
        loss, grads = self.network.grad(*args, **kw)
        self.optimizer.apply(grads, stream=kw.get("stream"))
        for c in self.cells:
            c.invalidate()     # packed filter banks / folded BN vectors are stale after the update
        return loss


class GraphedIteration:
    """One whole training iteration of train_video.py:170-177 (GAN phase: D step then G step; VAE phase: G step) captured
    ONCE as a CUDA graph and replayed: ~750 kernel launches per iteration stop costing host time, the GPU runs the
    iteration back to back.  Requirements (all provided by this module): the loss cells are built with
    device_rng=True and the optimisers with device_step=True, so every per-iteration varying quantity (noise draws,
    Adam bias correction) is read from device memory; inputs live in fixed device tensors the caller refreshes
    (hpvg_h2d on the same stream) before each launch; losses are read back after the launch.

    Usage:   it = GraphedIteration(stream, g_step, d_step, inputs..., g_kwargs)
             it.warmup(2); it.capture(); loop: [refresh inputs]; d_loss, g_loss = it()"""

    def __init__(self, stream, g_step, d_step, real, real_zero, noise_init, noise_amps, g_kwargs):
        self.stream, self.g_step, self.d_step = stream, g_step, d_step
        self.real, self.real_zero, self.noise_init, self.amps = real, real_zero, noise_init, list(noise_amps)
        self.g_kwargs = dict(g_kwargs)
        self.graph = None
        self.kernels_per_launch = 0
        self.side = None
        if d_step is not None:
            # third generator pass (the G step's random clip) gets its own workspace / draw counter so that it can run
            # next to the reconstruction pass
            self.trainer3 = GeneratorTrainer(g_step.network._netG, device_rng=True, salt=3)
        for cell in (g_step, d_step):
            if cell is None:
                continue
            if cell.optimizer.d_step is None or not cell.network.trainer.device_rng:
                raise HpvgError("GraphedIteration needs device_step=True optimisers and device_rng=True loss cells")

    def _body(self, finish, concurrent=False):
        st = self.stream
        # make the iteration CLOSED: packed filter banks / epilogue vectors derived from trainable weights must be
        # rebuilt inside the iteration, never carried over from the previous one through a Python-side cache (a
        # replayed graph would keep reading the buffer that was current at capture time)
        cells = []
        for cell in (self.d_step, self.g_step):
            if cell is not None:
                for c in cell.cells:
                    c.invalidate()
                    cells.append(c)
        from .networks_3d import prepack
        prepack(cells, st)     # every filter bank / epilogue vector of the trainable layers in one launch
        if self.d_step is None:      # VAE phase: one forward, nothing to overlap
            return None, self.g_step(self.real, self.real_zero, self.noise_init, self.amps, stream=st, finish=finish,
                                     **self.g_kwargs)
        # GAN phase (train_video.py:175-177).  The three generator forwards of an iteration — the D step's fake clip, the
        # G step's reconstruction and the G step's random clip — are independent of each other and of the D passes on
        # the real clip, and their nine coarse scales are far too small to fill 148 SMs.  They are issued on side
        # streams (captured as parallel branches of the graph) so that their small kernels fill the gaps of the main
        # stream's finest-scale kernels.  Semantics are unchanged: all three read the same generator weights (G's
        # update comes last), D passes and both Adam steps stay in order on the main stream, the BatchNorm moving
        # statistics are updated afterwards, forward by forward, in the reference's order (Q4).
        d_loss, g_loss = self.d_step.network, self.g_step.network
        s1, s2, s3 = self.side if concurrent else (st, st, st)
        g_loss.trainer.prepare_shared(st)
        fork = Event()
        fork.record(st)
        evs = []
        for s in (s1, s2, s3):
            if s is not st:
                s.wait_event(fork)
        fake = d_loss.trainer.forward(None, self.amps, noise_init=self.noise_init, is_random=True, stream=s1,
                                      defer_bn=True)
        rkw = g_loss.recon_forward_args(self.g_kwargs.get("isVAE", False), self.g_kwargs.get("trainable_body", ()),
                                        self.g_kwargs.get("train_codec", False))
        recon = g_loss.trainer.forward(self.real_zero, self.amps, stream=s2, defer_bn=True, **rkw)
        rand = self.trainer3.forward(None, self.amps, noise_init=self.noise_init, is_random=True, stream=s3,
                                     defer_bn=True)
        for s in (s1, s2, s3):
            ev = Event()
            ev.record(s)
            evs.append(ev)
        dl = self.d_step(self.real, self.noise_init, self.amps, stream=st, finish=finish, fake=fake["x"],
                         wait_fake=evs[0] if concurrent else None)
        gl = self.g_step(self.real, self.real_zero, self.noise_init, self.amps, stream=st, finish=finish,
                         recon_fw=recon, random_x=rand["x"], wait_recon=evs[1] if concurrent else None,
                         wait_random=evs[2] if concurrent else None, **self.g_kwargs)
        for fw in (fake, recon, rand):     # one launch per forward: a layer appears once per launch, order = reference order
            ops.bn_moving_update_multi(fw["bn_deferred"], stream=st)
        self._events = [fork] + evs      # keep the events alive as long as the captured graph
        return dl, gl

    def warmup(self, n=2):
        """Eager iterations: first-use allocations, cudaFuncSetAttribute, cached epilogue vectors."""
        out = None
        for _ in range(n):
            out = self._body(True)
        self.stream.sync()
        return out

    def capture(self, concurrent=True, wgrad_branches=2):
        """wgrad_branches: number of side streams the G step's weight-gradient kernels are spread over (0 = all on the
        main stream)."""
        from ._lib import lib
        from .runtime import Stream
        n0 = lib.hpvg_launch_count()
        if concurrent and self.d_step is not None and self.side is None:
            self.side = (Stream(), Stream(), Stream())
        if self.d_step is None and concurrent and self.g_step.network.enc_side is None:
            self.g_step.network.enc_side = Stream()
        if wgrad_branches and self.g_step.network.wgrad_sides is None:
            self.g_step.network.wgrad_sides = tuple(Stream() for _ in range(int(wgrad_branches)))
        self.graph = Graph(self.stream)
        with self.graph:
            self._terms = self._body(False, concurrent=concurrent and self.d_step is not None)
        self.kernels_per_launch = int(lib.hpvg_launch_count() - n0)
        return self

    def __call__(self):
        self.graph.launch()
        return tuple(t.finish(self.stream) if t is not None else None for t in self._terms)

    def destroy(self):
        if self.graph is not None:
            self.graph.destroy()
            self.graph = None
