"""Run-to-run reproducibility.  Every cross-block reduction on the training path sums its block partials in a FIXED
order (csrc/det_reduce.cuh: the last block to finish adds them in block order; Adam's gradient norms and the NCDHW
channel sums through a second fixed-order pass; the weight gradient through wgrad_reduce_kernel): no result depends on
the order in which blocks happen to be scheduled, so repeated runs agree BIT FOR BIT — including whole train iterations."""
import numpy as np
import pytest

from oracle import hpvg_oracle as orc

pytestmark = pytest.mark.gpu


def _same(a, b):
    return np.array_equal(np.asarray(a).view(np.uint8), np.asarray(b).view(np.uint8))


@pytest.mark.parametrize("mode", ["bf16", "tf32"])
def test_reductions_are_bitwise_reproducible(hpvg_gpu, mode):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    hp.set_precision(mode)
    try:
        rng = np.random.default_rng(0)
        shape = (2, 64, 5, 48, 65)
        x = hp.from_numpy((rng.standard_normal(shape) * 3 + 0.7).astype(np.float32))
        ga = hp.from_numpy(rng.standard_normal(shape).astype(np.float32))
        w = hp.from_numpy((rng.standard_normal((64, 64, 3, 3, 3)) * 0.05).astype(np.float32))
        aff = ops.affine_from_bias(hp.from_numpy((rng.standard_normal(64) * 0.1).astype(np.float32)))
        x_cl, ga_cl = ops.pack_cl(x), ops.pack_cl(ga)
        n = int(np.prod(x_cl.shape[:-1]))
        gamma, beta = hp.from_numpy(np.ones(64, np.float32)), hp.from_numpy(np.zeros(64, np.float32))
        runs = []
        for _ in range(3):
            stats = hp.Tensor((2, 64), hp.F64).zero_()
            y = ops.conv3d_cl_any(x_cl, w, aff, ops.ACT_NONE, 64, 64, stats=stats)       # fused BatchNorm statistics
            saved = hp.Tensor((4, 64), hp.F32)
            a = ops.bn_train_fused_cl(y, stats, gamma, beta, None, None, ops.ACT_LRELU, saved=saved)
            dg, db = hp.Tensor((64,), hp.F32), hp.Tensor((64,), hp.F32)
            gy = ops.bn_bwd_cl(ga_cl, y, saved, ops.ACT_LRELU, dgamma=dg, dbeta=db)
            cs = hp.Tensor((64,), hp.F32)
            ops.colsum_cl(gy, cs)
            dw = hp.Tensor((64, 64, 3, 3, 3), hp.F32)
            ops.conv_wgrad_cl(x_cl, gy, dw)
            ch = hp.Tensor((64,), hp.F32)
            ops.channel_sum(ga, ch)
            Gout, gp = ops.gp_grad(hp.from_numpy(rng.standard_normal((1, 3, 5, 48, 65)).astype(np.float32) * 0 + 0.3), 0.1)
            runs.append([stats.numpy(), a.numpy(), saved.numpy(), gy.numpy(), dg.numpy(), db.numpy(), cs.numpy(),
                         dw.numpy(), ch.numpy(), ops.mse(x, ga).numpy(), ops.kl_criterion(ga, ga).numpy(),
                         ops.mean(x).numpy(), gp.numpy()])
        for r in runs[1:]:
            for i, (p, q) in enumerate(zip(runs[0], r)):
                assert _same(p, q), "result %d differs between two runs" % i
        assert np.isfinite(runs[0][0]).all() and abs(runs[0][0][0].sum()) > 0
    finally:
        hp.set_precision("bf16")


def test_clipped_adam_is_bitwise_reproducible(hpvg_gpu):
    hp, ops = hpvg_gpu, hpvg_gpu.ops
    rng = np.random.default_rng(1)
    sizes = [64 * 64 * 27, 64, 3 * 64 * 27, 128 * 64 * 27 + 3]
    init = [(rng.standard_normal(n).astype(np.float32), (rng.standard_normal(n) * 4).astype(np.float32)) for n in sizes]
    outs = []
    for _ in range(3):
        ps = [hp.from_numpy(p) for p, _ in init]
        gs = [hp.from_numpy(g) for _, g in init]
        ms = [hp.Tensor((n,), hp.F32).zero_() for n in sizes]
        vs = [hp.Tensor((n,), hp.F32).zero_() for n in sizes]
        for step in (1, 2):
            ops.adam_clip_multi(ps, gs, ms, vs, [5e-4] * len(sizes), step, clip=5.0)
        outs.append([p.numpy() for p in ps])
    for o in outs[1:]:
        assert all(_same(a, b) for a, b in zip(outs[0], o))
    # and it is ClipByNorm + Adam (optimizers.py:41-43)
    w, m, v = init[0][0], np.zeros(sizes[0], np.float32), np.zeros(sizes[0], np.float32)
    for step in (1, 2):
        w, m, v = orc.adam_step(w, orc.clip_by_norm(init[0][1], 5.0), m, v, step, 5e-4, 0.5, 0.999)
    assert np.allclose(outs[0][0], w, rtol=1e-5, atol=1e-7)


def test_train_iterations_are_bitwise_reproducible(hpvg_gpu):
    """Two independent runs of two GAN-phase iterations (D step incl. the WGAN-GP double backward + G step, device
    Philox noise, spectral-norm updates, BatchNorm statistics, clip + Adam) from the same initial state end with
    IDENTICAL bits in every generator and discriminator parameter."""
    hp = hpvg_gpu
    from hpvg import networks_3d as n3, train as T
    from hpvg.utils import images as uimg

    def run():
        opt, oopt = uimg.default_opt(), orc.default_opt()
        nb = 4
        G = n3.GeneratorHPVAEGAN(opt)
        for _ in range(nb):
            G.init_next_stage()
        G.load_parameters(orc.init_generator_params(oopt, nb, seed=2))
        D = n3.WDiscriminator3D(opt)
        D.load_parameters(orc.init_discriminator_params(oopt, seed=2))
        G.noise_seed = 0xABCDEF
        rng = np.random.default_rng(2)
        st = hp.Stream()
        real = hp.from_numpy(np.tanh(rng.standard_normal((1, 3) + orc.scale_shape(oopt, nb))).astype(np.float32))
        real_zero = hp.from_numpy(np.tanh(rng.standard_normal((1, 3) + orc.scale_shape(oopt, 0))).astype(np.float32))
        noise = hp.from_numpy(rng.standard_normal((1, 128) + orc.scale_shape(oopt, 0)).astype(np.float32))
        amps = [1.0, 0.0, 0.0, 0.3, 0.25]
        block = G.body[-1]
        optG = T.ClippedAdam(opt, [{"params": T.trainable_params(block), "lr": opt.lr_g}], opt.lr_g, beta1=0.5,
                             beta2=0.999, device_step=True)
        optD = T.Adam(T.trainable_params(D), opt.lr_d, beta1=0.5, beta2=0.999, device_step=True)
        g_step = T.TrainOneStepCell(T.GWithLoss(opt, D, G, device_rng=True), optG, cells_to_invalidate=[block])
        d_step = T.TrainOneStepCell(T.DWithLoss(opt, D, G, alpha=0.41, device_rng=True), optD, cells_to_invalidate=[D])
        g_step.set_train()
        d_step.set_train()
        it = T.GraphedIteration(st, g_step, d_step, real, real_zero, noise, amps, dict(isVAE=False, trainable_body=(nb - 1,)))
        losses = [it._body(True) for _ in range(2)]
        st.sync()
        out = {"D." + k: t.numpy() for k, t in D.parameters_dict().items()}
        out.update({"G." + k: t.numpy() for k, t in G.parameters_dict().items()})
        return out, [(float(a), float(b)) for a, b in losses]

    p1, l1 = run()
    p2, l2 = run()
    assert l1 == l2
    diff = [k for k in p1 if not _same(p1[k], p2[k])]
    assert not diff, "parameters differ between two runs: %s" % diff[:5]
    # ... and the same bits with programmatic dependent launch switched off (csrc/launch.cuh): a kernel that touched
    # global memory before its griddepcontrol.wait, or overwrote a buffer its still-running predecessor reads, would
    # show up here as a difference
    before = hp.lib.hpvg_set_pdl(0)
    try:
        p3, l3 = run()
    finally:
        hp.lib.hpvg_set_pdl(before)
    assert before == 1, "programmatic dependent launch is the default"
    assert l1 == l3
    diff = [k for k in p1 if not _same(p1[k], p3[k])]
    assert not diff, "parameters differ between PDL on / off: %s" % diff[:5]
