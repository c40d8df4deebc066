"""Parity of the hand-restated training step (GWithLoss / DWithLoss / ClippedAdam, src/modules/losses.py and
optimizers.py) against torch-CPU autograd on the oracle, same weights / noise / inputs.
Tolerance: north_star rel-L2 <= 1e-2 per layer for bf16 gradients; deep chains accumulate, stated per assertion."""
import numpy as np
import pytest
import torch

from oracle import hpvg_oracle as orc
from util import rel_l2

pytestmark = pytest.mark.gpu


def _setup(hp, n_body, seed=3, bias_noise=True):
    from hpvg import networks_3d as n3
    from hpvg.utils import images as uimg
    opt, oopt = uimg.default_opt(), orc.default_opt()
    pg = orc.init_generator_params(oopt, n_body, seed=seed)
    pd = orc.init_discriminator_params(oopt, seed=seed)
    rng = np.random.default_rng(seed)
    if bias_noise:
        for p in (pg, pd):
            for k in p:
                if k.endswith("bias") or k.endswith("beta"):
                    p[k] = (rng.standard_normal(p[k].shape) * 0.05).astype(np.float32)
    G = n3.GeneratorHPVAEGAN(opt)
    for _ in range(n_body):
        G.init_next_stage()
    G.load_parameters(pg)
    D = n3.WDiscriminator3D(opt)
    D.load_parameters(pd)
    return G, D, opt, oopt, pg, pd, rng


def _grads_by_name(cell, book, names):
    pd = cell.parameters_dict()
    return {k: book.of(pd[k]).numpy() for k in names}


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def _report(got, ref, tol, what, min_cos=None, only=None):
    """rel-L2 (and optionally cosine) of every gradient tensor; prints the table, asserts at the end."""
    worst, rows, bad = 0.0, [], []
    for k in ref:
        if only is not None and not only(k):
            continue
        r = ref[k]
        if np.linalg.norm(r) < 1e-5:
            # conv biases in front of a BatchNorm have an analytically ZERO gradient (BN removes the mean); the
            # reference value is fp32 rounding noise, ours is bf16 rounding noise: absolute criterion
            if np.linalg.norm(got[k]) >= 5e-3:
                bad.append("%s should be ~0 (|g|=%.2e)" % (k, np.linalg.norm(got[k])))
            continue
        e, c = rel_l2(got[k], r), _cos(got[k], r)
        rows.append("%-36s rel-L2 %.3e  cos %.5f  |ref| %.3e" % (k, e, c, np.linalg.norm(r)))
        worst = max(worst, e)
        if e >= tol or (min_cos is not None and c < min_cos):
            bad.append("%s rel-L2 %.3e cos %.5f" % (k, e, c))
    print("---- %s" % what)
    print("\n".join(rows))
    assert not bad, "%s: gradients outside tolerance %.1e: %s" % (what, tol, bad)
    return worst


# Conditioning note (measured with the oracle alone, tools/ + DESIGN.md §parity): in this randomly initialised
# network a 1e-3 relative perturbation of the FORWARD activations changes the BatchNorm-path weight gradients by
# 6-11 % rel-L2 (LeakyReLU mask flips coupled through batch statistics).  A bf16 forward differs from the oracle's
# by ~3e-3 per layer, so end-to-end gradients can only agree to that conditioning; the kernels themselves are held to
# rel-L2 <= 1e-2 by the teacher-forced per-layer test below and by tests/test_gpu_backward.py.
E2E_BN_TOL, E2E_BN_COS = 0.20, 0.98


@pytest.mark.parametrize("shape", [(1, 4, 30, 41), (1, 13, 192, 257)])
def test_per_layer_backward_teacher_forced(hpvg_gpu, shape):
    """Every conv+BN+LeakyReLU layer of a refinement stage, fed with the ORACLE's input activation and upstream
    gradient: dW, dgamma, dbeta and dx within rel-L2 1e-2 (north_star, bf16) — at scale 1 and at the finest scale
    (13 x 192 x 257, BASELINE.json's size: BatchNorm statistics and their backward over 641 k voxels)."""
    hp = hpvg_gpu
    from hpvg import networks_3d as n3, ops, train as T
    from util import bf16_round
    G, D, opt, oopt, pg, pd, rng = _setup(hp, 1)
    G.set_train(True)
    x3 = bf16_round(rng.standard_normal((1, 3) + shape[1:]) * 0.5)
    up = rng.standard_normal((1, 3) + shape[1:]).astype(np.float32) * 0.3
    gout = rng.standard_normal((1, 3) + shape[1:]).astype(np.float32)
    tg = orc.to_torch(pg, requires_grad=("body.",))
    taps = {}
    xt = torch.from_numpy(x3).requires_grad_(True)
    with orc.bf16_emulation():
        pre = orc.block_forward(xt, tg, "body.0.", oopt, True, taps=taps)
        out = torch.tanh(pre + torch.from_numpy(up))
    for v in taps.values():
        if v.requires_grad:
            v.retain_grad()
    out.backward(torch.from_numpy(gout))
    block = G.body[0]
    pdict = G.parameters_dict()
    ws = n3.Workspace()
    for j in range(opt.num_layer + 1):
        layer = block.layers[j]
        xin = x3 if j == 0 else taps["body.0.%d.out" % (j - 1)].detach().numpy()
        x_cl = ops.pack_cl(hp.from_numpy(xin), c_pitch=8 if j == 0 else 64)
        a, ctx = T.layer_forward_train(layer, x_cl, ws, "tf%d" % j)
        if j == 0:     # head conv weight gradient: zero-padded 64-channel copy (small case) or the narrow tensor itself
            ctx["x_wide"] = ops.pack_cl(hp.from_numpy(xin), c_pitch=64, zero_to=64) if shape[1] == 4 else x_cl
        ga = taps["body.0.%d.out" % j].grad.numpy()
        book = T.GradBook()
        dx = T.layer_backward(layer, ctx, ops.pack_cl(hp.from_numpy(ga)), book, ws, "tf%d" % j, need_dx=True)
        pre_n = "body.0.%d." % j
        for nm in ("0.weight", "1.bn2d.gamma", "1.bn2d.beta"):
            e = rel_l2(book.of(pdict[pre_n + nm]).numpy(), tg[pre_n + nm].grad.numpy())
            assert e < 1e-2, "layer %d %s rel-L2 %.3e" % (j, nm, e)
        ref_dx = xt.grad.numpy() if j == 0 else taps["body.0.%d.out" % (j - 1)].grad.numpy()
        got_dx = dx.numpy() if j == 0 else ops.unpack_cl(dx).numpy()
        e = rel_l2(got_dx, ref_dx)
        assert e < 1e-2, "layer %d dx rel-L2 %.3e" % (j, e)
    # the tail conv 64 -> 3 (networks_3d.py:399): weight gradient from the narrow (8-channel) output gradient, data
    # gradient through the transposed bank — fed with the oracle's gradient wrt the tail's output
    jt = opt.num_layer + 1
    tail = block.layers[jt]
    x_t = ops.pack_cl(hp.from_numpy(taps["body.0.%d.out" % (jt - 1)].detach().numpy()))
    g_pre = taps["body.0.%d.conv" % jt].grad.numpy()
    book = T.GradBook()
    dx = T.conv_backward(tail, {"x": x_t, "layer": tail}, ops.pack_cl(hp.from_numpy(g_pre), c_pitch=8), book, ws, "tft",
                         need_dx=True, want_dw=True)
    e = rel_l2(book.of(pdict["body.0.%d.weight" % jt]).numpy(), tg["body.0.%d.weight" % jt].grad.numpy())
    assert e < 1e-2, "tail weight rel-L2 %.3e" % e
    e = rel_l2(ops.unpack_cl(dx).numpy(), taps["body.0.%d.out" % (jt - 1)].grad.numpy())
    assert e < 1e-2, "tail dx rel-L2 %.3e" % e


def test_vae_phase_g_step(hpvg_gpu):
    """scale_idx = 1 (VAE phase, one refinement stage): loss and gradients of encode / decoder / body.0."""
    hp = hpvg_gpu
    from hpvg import train as T
    G, D, opt, oopt, pg, pd, rng = _setup(hp, 1)
    s0, s1 = orc.scale_shape(oopt, 0), orc.scale_shape(oopt, 1)
    real = np.tanh(rng.standard_normal((1, 3) + s1)).astype(np.float32)
    real_zero = np.tanh(rng.standard_normal((1, 3) + s0)).astype(np.float32)
    z = rng.standard_normal((1, 128) + s0).astype(np.float32)
    amps = [1.0, 0.0]
    tg = orc.to_torch(pg, requires_grad=("encode.", "decoder.", "body."))
    with orc.bf16_emulation():      # same quantisation points as the CUDA path (see oracle: bf16_emulation)
        loss_ref = orc.g_loss(torch.from_numpy(real), torch.from_numpy(real_zero), None, amps, tg, None, oopt, True,
                              z_pred=torch.from_numpy(z))
    loss_ref.backward()
    names = [k for k, t in tg.items() if t.requires_grad]
    ref = {k: tg[k].grad.numpy() for k in names}
    G.set_train(True)
    gl = T.GWithLoss(opt, D, G)
    loss, book = gl.grad(hp.from_numpy(real), hp.from_numpy(real_zero), None, amps, isVAE=True, trainable_body=(0,),
                         train_codec=True, z_pred=hp.from_numpy(z))
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    got = _grads_by_name(G, book, names)
    # encoder: only the (smooth) KL term reaches it -> tight; decoder/body: conditioning-limited (see note above)
    _report(got, ref, 1e-2, "VAE phase / encoder", only=lambda k: k.startswith("encode."))
    _report(got, ref, E2E_BN_TOL, "VAE phase / decoder + body (BatchNorm path)", min_cos=E2E_BN_COS,
            only=lambda k: not k.startswith("encode."))
    # moving statistics were updated exactly like the oracle's (Q4)
    mm = G.parameters_dict()["decoder.0.1.bn2d.moving_mean"].numpy()
    assert np.allclose(mm, tg["decoder.0.1.bn2d.moving_mean"].numpy(), atol=2e-3)


def test_gan_phase_g_and_d_steps(hpvg_gpu):
    """scale_idx = 3 (first GAN scale): G step trains body[-1] only; D step incl. the WGAN-GP double backward."""
    hp = hpvg_gpu
    from hpvg import train as T
    G, D, opt, oopt, pg, pd, rng = _setup(hp, 3, seed=5)
    s0, s3 = orc.scale_shape(oopt, 0), orc.scale_shape(oopt, 3)
    real = np.tanh(rng.standard_normal((1, 3) + s3)).astype(np.float32)
    real_zero = np.tanh(rng.standard_normal((1, 3) + s0)).astype(np.float32)
    z_pred = rng.standard_normal((1, 128) + s0).astype(np.float32)
    noise_init = rng.standard_normal((1, 128) + s0).astype(np.float32)
    nz = {3: rng.standard_normal((1, 3) + s3).astype(np.float32)}
    amps = [1.0, 0.0, 0.0, 0.3]
    alpha = 0.37
    G.set_train(True)
    D.set_train(True)
    # ------------------------------------------------ D step (runs first, train_video.py:175)
    tg = orc.to_torch(pg)
    td = orc.to_torch(pd, requires_grad=("head.", "body.", "tail."))
    with torch.no_grad(), orc.bf16_emulation():
        fake_ref, _ = orc.generator_forward(None, amps, tg, oopt, noise_init=torch.from_numpy(noise_init),
                                            is_random=True, training=True,
                                            noises={k: torch.from_numpy(v) for k, v in nz.items()})
    with orc.bf16_emulation():
        dloss_ref = orc.d_loss(torch.from_numpy(real), fake_ref, alpha, td, oopt)
    dloss_ref.backward()
    dnames = [k for k, t in td.items() if t.requires_grad]
    dref = {k: td[k].grad.numpy() for k in dnames}
    dl = T.DWithLoss(opt, D, G, alpha=alpha)
    dloss, dbook = dl.grad(hp.from_numpy(real), hp.from_numpy(noise_init), amps,
                           noises={k: hp.from_numpy(v) for k, v in nz.items()})
    assert abs(float(dloss) - float(dloss_ref)) < 2e-2 * max(abs(float(dloss_ref)), 1e-2), (float(dloss), float(dloss_ref))
    _report(_grads_by_name(D, dbook, dnames), dref, E2E_BN_TOL, "D step (incl. WGAN-GP double backward)",
            min_cos=E2E_BN_COS)
    # spectral-norm state advanced three times (real, fake, xhat) exactly like the oracle's (Q5)
    assert np.allclose(D.parameters_dict()["body.2.0.weight_u"].numpy(), td["body.2.0.weight_u"].numpy(), atol=1e-4)
    # ------------------------------------------------ G step
    tg2 = orc.to_torch({k: v.detach().numpy() for k, v in tg.items()}, requires_grad=("body.2.",))
    with orc.bf16_emulation():
        gloss_ref = orc.g_loss(torch.from_numpy(real), torch.from_numpy(real_zero), torch.from_numpy(noise_init), amps,
                               tg2, {k: v.detach() for k, v in td.items()}, oopt, False,
                               z_pred=torch.from_numpy(z_pred), noises={k: torch.from_numpy(v) for k, v in nz.items()})
    gloss_ref.backward()
    gnames = [k for k, t in tg2.items() if t.requires_grad]
    gref = {k: tg2[k].grad.numpy() for k in gnames}
    gl = T.GWithLoss(opt, D, G)
    gloss, gbook = gl.grad(hp.from_numpy(real), hp.from_numpy(real_zero), hp.from_numpy(noise_init), amps, isVAE=False,
                           trainable_body=(2,), z_pred=hp.from_numpy(z_pred),
                           noises={k: hp.from_numpy(v) for k, v in nz.items()})
    assert abs(float(gloss) - float(gloss_ref)) < 2e-2 * abs(float(gloss_ref)), (float(gloss), float(gloss_ref))
    _report(_grads_by_name(G, gbook, gnames), gref, E2E_BN_TOL, "G step (GAN phase)", min_cos=E2E_BN_COS)


@pytest.mark.parametrize("scale,shape", [(7, (7, 121, 162)), (9, (13, 192, 257))])
def test_d_step_with_gradient_penalty_at_large_scales(hpvg_gpu, scale, shape):
    """DWithLoss (losses.py:27-56) at realistic sizes — scale 7 of the default pyramid (137 k voxels, ragged tiles in H
    and W) and the finest scale (13 x 192 x 257 = 641 k voxels, BASELINE.json's size) — on given real / fake clips:
    loss, first-order terms and the WGAN-GP double backward of all seven discriminator layers against torch-CPU
    autograd (create_graph=True) on the oracle."""
    hp = hpvg_gpu
    from hpvg import train as T
    G, D, opt, oopt, pg, pd, rng = _setup(hp, 0, seed=9)
    s7 = orc.scale_shape(oopt, scale)
    assert s7 == shape
    real = np.tanh(rng.standard_normal((1, 3) + s7)).astype(np.float32)
    fake = np.tanh(rng.standard_normal((1, 3) + s7)).astype(np.float32)
    alpha = 0.61
    D.set_train(True)
    td = orc.to_torch(pd, requires_grad=("head.", "body.", "tail."))
    with orc.bf16_emulation():
        dloss_ref = orc.d_loss(torch.from_numpy(real), torch.from_numpy(fake), alpha, td, oopt)
    dloss_ref.backward()
    dnames = [k for k, t in td.items() if t.requires_grad]
    dref = {k: td[k].grad.numpy() for k in dnames}
    dl = T.DWithLoss(opt, D, G, alpha=alpha)
    dloss, dbook = dl.grad(hp.from_numpy(real), None, None, fake=hp.from_numpy(fake))
    assert abs(float(dloss) - float(dloss_ref.detach())) < 2e-2 * max(abs(float(dloss_ref.detach())), 1e-2)
    # no BatchNorm in D => none of the conditioning caveat below: the north_star per-layer tolerance applies end to end
    _report(_grads_by_name(D, dbook, dnames), dref, 1e-2, "D step at scale %d (incl. WGAN-GP double backward)" % scale,
            min_cos=0.9999)


def test_train_one_step_updates_parameters_like_clipped_adam(hpvg_gpu):
    hp = hpvg_gpu
    from hpvg import train as T
    G, D, opt, oopt, pg, pd, rng = _setup(hp, 3, seed=7)
    s0, s3 = orc.scale_shape(oopt, 0), orc.scale_shape(oopt, 3)
    real = np.tanh(rng.standard_normal((1, 3) + s3)).astype(np.float32)
    real_zero = np.tanh(rng.standard_normal((1, 3) + s0)).astype(np.float32)
    z_pred = rng.standard_normal((1, 128) + s0).astype(np.float32)
    noise_init = rng.standard_normal((1, 128) + s0).astype(np.float32)
    amps = [1.0, 0.0, 0.0, 0.3]
    nz = {3: rng.standard_normal((1, 3) + s3).astype(np.float32)}
    block = G.body[-1]
    groups = [{"params": T.trainable_params(block), "lr": opt.lr_g}]
    optim = T.ClippedAdam(opt, groups, opt.lr_g, beta1=opt.beta1, beta2=0.999)
    gl = T.GWithLoss(opt, D, G)
    step = T.TrainOneStepCell(gl, optim, cells_to_invalidate=[block])
    step.set_train()
    before = {k: t.numpy() for k, t in T.trainable_params(block)}
    loss = step(hp.from_numpy(real), hp.from_numpy(real_zero), hp.from_numpy(noise_init), amps, isVAE=False,
                trainable_body=(2,), z_pred=hp.from_numpy(z_pred), noises={k: hp.from_numpy(v) for k, v in nz.items()})
    assert np.isfinite(float(loss))
    grads = {k: gl.grads.of(t).numpy() for k, t in T.trainable_params(block)}
    for k, t in T.trainable_params(block):
        want, _, _ = orc.adam_step(before[k], orc.clip_by_norm(grads[k], opt.grad_clip), np.zeros_like(before[k]),
                                   np.zeros_like(before[k]), 1, opt.lr_g, 0.5, 0.999)
        assert rel_l2(t.numpy(), want) < 1e-5, k
        assert not np.array_equal(t.numpy(), before[k]) or np.linalg.norm(grads[k]) == 0


def test_device_randn_statistics_and_counter(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    ctr = hp.Tensor((1,), hp.U64).zero_()
    a = ops.randn((1, 128, 4, 24, 33), seed=77, d_offset=ctr).numpy()
    assert abs(a.mean()) < 5e-3 and abs(a.std() - 1.0) < 5e-3
    assert np.array_equal(a, ops.randn(a.shape, seed=77, d_offset=ctr).numpy())      # same key -> same draw
    ops.counter_add(ctr, 1)
    b = ops.randn(a.shape, seed=77, d_offset=ctr).numpy()
    assert not np.array_equal(a, b) and abs(float((a * b).mean())) < 5e-3            # fresh, uncorrelated draw
    assert np.array_equal(b, ops.randn(a.shape, seed=77, offset=1).numpy())          # device counter == host offset


@pytest.mark.parametrize("n_body", [3, 9])
def test_cuda_graph_iteration_matches_eager_iteration(hpvg_gpu, n_body):
    """The whole GAN-phase iteration (D step + G step + both Adam updates) replayed as ONE CUDA graph must evolve the
    weights exactly like the eager launch sequence: same kernels, device-resident noise counters and Adam step.
    n_body = 3: the first GAN scale; n_body = 9: the full pyramid at the finest scale 13 x 192 x 257 — the size at which
    the graph's three parallel generator branches really overlap with the main stream's kernels."""
    hp = hpvg_gpu
    from hpvg import train as T

    def run(graphed, iters=3):
        G, D, opt, oopt, pg, pd, rng = _setup(hp, n_body, seed=5)
        G.noise_seed = 0x1234567
        s0, s3 = orc.scale_shape(oopt, 0), orc.scale_shape(oopt, n_body)
        st = hp.Stream()
        real = hp.from_numpy(np.tanh(rng.standard_normal((1, 3) + s3)).astype(np.float32))
        real_zero = hp.from_numpy(np.tanh(rng.standard_normal((1, 3) + s0)).astype(np.float32))
        noise = hp.from_numpy(rng.standard_normal((1, 128) + s0).astype(np.float32))
        amps = [1.0, 0.0, 0.0] + [0.3] * (n_body - 2)
        block = G.body[-1]
        optG = T.ClippedAdam(opt, [{"params": T.trainable_params(block), "lr": opt.lr_g}], opt.lr_g, beta1=0.5,
                             beta2=0.999, device_step=True)
        optD = T.Adam(T.trainable_params(D), opt.lr_d, beta1=0.5, beta2=0.999, device_step=True)
        g_step = T.TrainOneStepCell(T.GWithLoss(opt, D, G, device_rng=True), optG, cells_to_invalidate=[block])
        d_step = T.TrainOneStepCell(T.DWithLoss(opt, D, G, alpha=0.37, device_rng=True), optD, cells_to_invalidate=[D])
        g_step.set_train()
        d_step.set_train()
        it = T.GraphedIteration(st, g_step, d_step, real, real_zero, noise, amps,
                                dict(isVAE=False, trainable_body=(n_body - 1,)))
        losses = []
        it.warmup(1)
        if graphed:
            it.capture()
            assert it.kernels_per_launch > 100
            for _ in range(iters):
                losses.append(it())
        else:
            for _ in range(iters):
                losses.append(it._body(True))
        st.sync()
        out = ({k: t.numpy() for k, t in D.parameters_dict().items()},
               {k: t.numpy() for k, t in block.parameters_dict("b.").items()},
               [(float(a), float(b)) for a, b in losses])
        if graphed:
            it.destroy()
        return out

    d_e, g_e, l_e = run(False)
    d_g, g_g, l_g = run(True)
    for (de, ge), (dg, gg) in zip(l_e, l_g):
        assert abs(de - dg) < 2e-3 * max(1.0, abs(de)) and abs(ge - gg) < 2e-3 * max(1.0, abs(ge)), (l_e, l_g)
    # fp64 atomics in the BatchNorm statistics make the summation order non-deterministic at the 1e-16 level; after
    # three Adam steps the weights agree to float32 round-off
    for ref, got in ((d_e, d_g), (g_e, g_g)):
        for k in ref:
            if k.startswith("b.") and k.endswith(".0.bias") and not k.startswith("b.6"):
                # conv bias in front of a BatchNorm: its gradient is analytically ZERO, so Adam normalises pure
                # rounding noise into +-lr steps — not reproducible run to run even eagerly (measured); bounded only
                assert np.abs(got[k] - ref[k]).max() <= 2 * 3 * 5e-4 + 1e-6, k
                continue
            assert rel_l2(got[k], ref[k]) < 2e-3, k
    assert l_e[0] != l_e[1]        # noise really changes from iteration to iteration


def test_vae_phase_with_reparameterised_z_is_training_true(hpvg_gpu):
    """The non-default `GeneratorHPVAEGAN(opt, is_training=True)` branch (networks_3d.py:414-417): z = eps*exp(.5 lv)+mu,
    so the reconstruction losses reach the encoder through the decoder's input gradient and the reparam backward."""
    hp = hpvg_gpu
    from hpvg import networks_3d as n3, train as T
    G0, D, opt, oopt, pg, pd, rng = _setup(hp, 1)
    G = n3.GeneratorHPVAEGAN(opt, is_training=True)
    G.init_next_stage()
    G.load_parameters(pg)
    s0, s1 = orc.scale_shape(oopt, 0), orc.scale_shape(oopt, 1)
    real = np.tanh(rng.standard_normal((1, 3) + s1)).astype(np.float32)
    real_zero = np.tanh(rng.standard_normal((1, 3) + s0)).astype(np.float32)
    eps = rng.standard_normal((1, 128) + s0).astype(np.float32)
    amps = [1.0, 0.0]
    tg = orc.to_torch(pg, requires_grad=("encode.", "decoder.", "body."))
    with orc.bf16_emulation():
        loss_ref = orc.g_loss(torch.from_numpy(real), torch.from_numpy(real_zero), None, amps, tg, None, oopt, True,
                              eps=torch.from_numpy(eps), is_training_flag=True)
    loss_ref.backward()
    names = [k for k, t in tg.items() if t.requires_grad and k.startswith("encode.")]
    ref = {k: tg[k].grad.numpy() for k in names}
    G.set_train(True)
    gl = T.GWithLoss(opt, D, G)
    loss, book = gl.grad(hp.from_numpy(real), hp.from_numpy(real_zero), None, amps, isVAE=True, trainable_body=(0,),
                         train_codec=True, eps=hp.from_numpy(eps))
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    got = _grads_by_name(G, book, names)
    # the reconstruction gradient dominates the KL one by orders of magnitude here, so this exercises the new path
    # the chain is the longest in the model (body BN x6 + decoder BN x6 before reaching the encoder): conditioning-
    # limited like the other BatchNorm-path gradients (see the note above), measured rel-L2 0.15-0.21, cos 0.98-0.99
    _report(got, ref, 0.30, "VAE phase, is_training=True / encoder through the reparameterisation", min_cos=0.97)


def test_gan_phase_train_depth_2_reaches_both_stages(hpvg_gpu):
    """--train-depth 2 (train_video.py:78-86): the optimiser holds body[-2:], and the reconstruction loss must reach BOTH
    stages — the backward chain runs from the top stage through the resize adjoint into the stage below and stops at
    the lowest trainable one (round-1 advisor finding: only the top stage received a gradient)."""
    hp = hpvg_gpu
    from hpvg import train as T
    nb = 5
    G, D, opt, oopt, pg, pd, rng = _setup(hp, nb, seed=11)
    s0, st = orc.scale_shape(oopt, 0), orc.scale_shape(oopt, nb)
    real = np.tanh(rng.standard_normal((1, 3) + st)).astype(np.float32)
    real_zero = np.tanh(rng.standard_normal((1, 3) + s0)).astype(np.float32)
    z_pred = rng.standard_normal((1, 128) + s0).astype(np.float32)
    noise_init = rng.standard_normal((1, 128) + s0).astype(np.float32)
    nz = {s: rng.standard_normal((1, 3) + orc.scale_shape(oopt, s)).astype(np.float32) for s in range(3, nb + 1)}
    amps = [1.0, 0.0, 0.0, 0.3, 0.25, 0.2]
    G.set_train(True)
    D.set_train(True)
    tg = orc.to_torch(pg, requires_grad=("body.3.", "body.4."))
    td = orc.to_torch(pd)
    with orc.bf16_emulation():
        gloss_ref = orc.g_loss(torch.from_numpy(real), torch.from_numpy(real_zero), torch.from_numpy(noise_init), amps,
                               tg, td, oopt, False, z_pred=torch.from_numpy(z_pred),
                               noises={k: torch.from_numpy(v) for k, v in nz.items()})
    gloss_ref.backward()
    names = [k for k, t in tg.items() if t.requires_grad]
    ref = {k: tg[k].grad.numpy() for k in names}
    assert np.linalg.norm(ref["body.3.2.0.weight"]) > 0
    gl = T.GWithLoss(opt, D, G)
    loss, book = gl.grad(hp.from_numpy(real), hp.from_numpy(real_zero), hp.from_numpy(noise_init), amps, isVAE=False,
                         trainable_body=(3, 4), z_pred=hp.from_numpy(z_pred),
                         noises={k: hp.from_numpy(v) for k, v in nz.items()})
    assert abs(float(loss) - float(gloss_ref)) < 2e-2 * abs(float(gloss_ref))
    got = _grads_by_name(G, book, names)
    assert np.linalg.norm(got["body.3.2.0.weight"]) > 0, "the lower trainable stage received no gradient"
    _report(got, ref, E2E_BN_TOL, "G step, GAN phase, train_depth 2", min_cos=E2E_BN_COS)


@pytest.mark.parametrize("mode", ["bf16", "tf32"])
def test_gan_phase_train_all_reaches_the_decoder(hpvg_gpu, mode):
    """--train-all with fewer stages than train_depth (train_video.py:95-103): encode / decoder / every stage are in the
    optimiser in the GAN phase too, there is no stop_gradient (networks_3d.py:437), and the reconstruction term flows
    through the whole chain into the decoder; the encoder sees nothing (z is pure noise, Q2; no KL term in this phase)."""
    hp = hpvg_gpu
    import contextlib
    from hpvg import networks_3d as n3, train as T
    from hpvg.utils import images as uimg
    nb = 3
    hp.set_precision(mode)
    emu = orc.bf16_emulation if mode == "bf16" else contextlib.nullcontext     # tf32: the plain fp32 oracle
    opt, oopt = uimg.default_opt(train_all=True, train_depth=5), orc.default_opt(train_all=True, train_depth=5)
    pg = orc.init_generator_params(oopt, nb, seed=13)
    pd = orc.init_discriminator_params(oopt, seed=13)
    rng = np.random.default_rng(13)
    G = n3.GeneratorHPVAEGAN(opt)
    for _ in range(nb):
        G.init_next_stage()
    G.load_parameters(pg)
    D = n3.WDiscriminator3D(opt)
    D.load_parameters(pd)
    s0, st = orc.scale_shape(oopt, 0), orc.scale_shape(oopt, nb)
    real = np.tanh(rng.standard_normal((1, 3) + st)).astype(np.float32)
    real_zero = np.tanh(rng.standard_normal((1, 3) + s0)).astype(np.float32)
    z_pred = rng.standard_normal((1, 128) + s0).astype(np.float32)
    noise_init = rng.standard_normal((1, 128) + s0).astype(np.float32)
    nz = {3: rng.standard_normal((1, 3) + st).astype(np.float32)}
    amps = [1.0, 0.0, 0.0, 0.3]
    G.set_train(True)
    D.set_train(True)
    tg = orc.to_torch(pg, requires_grad=("encode.", "decoder.", "body."))
    td = orc.to_torch(pd)
    with emu():
        gloss_ref = orc.g_loss(torch.from_numpy(real), torch.from_numpy(real_zero), torch.from_numpy(noise_init), amps,
                               tg, td, oopt, False, z_pred=torch.from_numpy(z_pred),
                               noises={k: torch.from_numpy(v) for k, v in nz.items()})
    gloss_ref.backward()
    names = [k for k, t in tg.items() if t.requires_grad and t.grad is not None and not k.startswith("encode.")]
    ref = {k: tg[k].grad.numpy() for k in names}
    assert np.linalg.norm(ref["decoder.0.0.weight"]) > 0
    try:
        gl = T.GWithLoss(opt, D, G)
        loss, book = gl.grad(hp.from_numpy(real), hp.from_numpy(real_zero), hp.from_numpy(noise_init), amps, isVAE=False,
                             trainable_body=(0, 1, 2), train_codec=True, z_pred=hp.from_numpy(z_pred),
                             noises={k: hp.from_numpy(v) for k, v in nz.items()})
        assert abs(float(loss) - float(gloss_ref)) < 2e-2 * abs(float(gloss_ref))
        got = _grads_by_name(G, book, names)
    finally:
        hp.set_precision("bf16")
    assert np.linalg.norm(got["decoder.0.0.weight"]) > 0, "the decoder received no gradient"
    # the longest BatchNorm + LeakyReLU chain in the model (3 stages + decoder = 24 layers below the loss): the
    # conditioning of the end-to-end gradient (mask flips) grows with every layer crossed — measured
    # 0.33 rel-L2 / cos 0.94 at the bottom of the chain in bf16; the tf32 mode against the plain fp32 oracle is the
    # tighter statement
    tol, cos = (0.5, 0.90) if mode == "bf16" else (0.25, 0.97)
    _report(got, ref, tol, "G step, GAN phase, --train-all (decoder + 3 stages), %s" % mode, min_cos=cos)


def test_cuda_graph_vae_iteration_matches_eager_iteration(hpvg_gpu):
    """The VAE-phase iteration (d_step None: encoder + decoder + refinement stage, resize backward, KL / MSE terms,
    ClippedAdam with per-group learning rates) captured as a CUDA graph evolves the weights like the eager sequence.
    The capture allocates temporaries and the layers' packed filter banks; between replays the test churns the caching
    allocator — memory baked into the graph must stay reserved for it (runtime: _GRAPH_OWNED)."""
    hp = hpvg_gpu
    from hpvg import driver, train as T

    def run(graphed, iters=4):
        scale_idx = 2
        G, D, opt, oopt, pg, pd, rng = _setup(hp, scale_idx, seed=21)
        G.noise_seed = 0x7654321
        st = hp.Stream()
        real = hp.from_numpy(np.tanh(rng.standard_normal((1, 3) + orc.scale_shape(oopt, scale_idx))).astype(np.float32))
        real_zero = hp.from_numpy(np.tanh(rng.standard_normal((1, 3) + orc.scale_shape(oopt, 0))).astype(np.float32))
        noise = hp.from_numpy(rng.standard_normal((1, 128) + orc.scale_shape(oopt, 0)).astype(np.float32))
        amps = [1.0, 0.0, 0.0]
        groups, body_idx, codec = driver.generator_param_groups(opt, G, scale_idx)
        optG = T.ClippedAdam(opt, groups, opt.lr_g, beta1=0.5, beta2=0.999, device_step=True)
        cells = [G.body[i] for i in body_idx] + [G.encode, G.decoder]
        g_step = T.TrainOneStepCell(T.GWithLoss(opt, None, G, device_rng=True), optG, cells_to_invalidate=cells)
        G.set_train(True)
        it = T.GraphedIteration(st, g_step, None, real, real_zero, noise, amps,
                                dict(isVAE=True, trainable_body=body_idx, train_codec=codec))
        it.warmup(1)
        losses = []
        if graphed:
            it.capture()
            assert it.kernels_per_launch > 100
            for _ in range(iters):
                junk = [hp.Tensor((n,), hp.F32).zero_(st) for n in (64, 4096, 27 * 64 * 64 * 2, 16384)]   # allocator churn
                del junk
                for c in cells:
                    c.invalidate()
                losses.append(float(it()[1]))
        else:
            for _ in range(iters):
                losses.append(float(it._body(True)[1]))
        st.sync()
        out = {k: t.numpy() for k, t in G.parameters_dict().items()}
        if graphed:
            it.destroy()
        return out, losses

    p_e, l_e = run(False)
    p_e2, _ = run(False)
    p_g, l_g = run(True)
    for a, b in zip(l_e, l_g):
        assert abs(a - b) < 2e-3 * max(1.0, abs(a)), (l_e, l_g)
    assert l_e[0] != l_e[1]
    for k in p_e:
        if k.endswith(".0.bias") and (k.startswith("body.") or k.startswith("decoder.")) and ".6." not in k:
            # conv bias in front of a BatchNorm: analytically zero gradient, Adam amplifies rounding noise (see above)
            assert np.abs(p_g[k] - p_e[k]).max() <= 2 * 4 * 5e-4 + 1e-6, k
            continue
        # Adam divides by sqrt(v): where a gradient element is tiny its +-lr step follows rounding noise (fp64 atomics in
        # the BatchNorm statistics are order-dependent at 1e-16), so two EAGER runs already differ slightly; the graph
        # must be no further from an eager run than eager runs are from each other (plus float32 round-off)
        noise = rel_l2(p_e2[k], p_e[k])
        assert rel_l2(p_g[k], p_e[k]) < 1e-3 + 2 * noise, (k, rel_l2(p_g[k], p_e[k]), noise)
