"""2-D (image) path — SURVEY.md §8 a10 / config 1: the reference's `networks_2d.py` graph on the same kernels (T == 1).

Parity against the oracle run with nd=2 (F.conv2d, BatchNorm2d naming, bilinear align_corners resize, refinement
noise at every scale), small pyramid of config 1 (`--img-size 64` -> stop_scale 4, widths [33,39,46,54,65])."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hpvg_oracle as orc
from util import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-2
TOL_E2E = 5e-2


def _build(hp, n_body, seed=13, img_size=64):
    from hpvg import networks_2d as n2
    from hpvg.utils import images as uimg
    opt, oopt = uimg.default_opt(img_size=img_size), orc.default_opt(img_size=img_size)
    params = orc.randomize_bn_stats(orc.init_generator_params(oopt, n_body, seed=seed, nd=2), opt=oopt, nd=2)
    net = n2.GeneratorHPVAEGAN(opt)
    for _ in range(n_body):
        net.init_next_stage()
    net.load_parameters(params)
    return net, opt, oopt, params


def test_small_pyramid_geometry_config1(hpvg_gpu):
    from hpvg.utils import images as uimg
    opt = uimg.default_opt(img_size=64)
    assert opt.stop_scale == 4
    assert [uimg.scale_shape_2d(opt, i)[1] for i in range(5)] == [33, 39, 46, 54, 65]
    assert [uimg.scale_shape_2d(opt, i) for i in range(5)] == [orc.scale_shape_2d(orc.default_opt(img_size=64), i)
                                                               for i in range(5)]


def test_parameter_names_2d_follow_reference_checkpoint_contract(hpvg_gpu):
    """src/tools/pt2ms.py:30-89: 2-D BatchNorm lives at `...1.gamma` (no bn2d level), conv weights are 4-D."""
    net, opt, oopt, params = _build(hpvg_gpu, 2)
    mine = net.parameters_dict()
    assert set(mine) == set(params)
    assert mine["decoder.0.1.gamma"].shape == (64,)
    assert mine["body.1.6.weight"].shape == (3, 64, 3, 3)
    assert mine["encode._features.0.0.weight_v"].shape == (27, 1)


@pytest.mark.parametrize("case", [((24, 33), (29, 39)), ((40, 54), (48, 65)), ((7, 5), (20, 31))])
def test_bilinear_resize_align_corners(hpvg_gpu, case):
    """ops.ResizeBilinear(size, align_corners=True) (images.py:40-51) == T=1 case of the linear-resize kernel."""
    hp = hpvg_gpu
    from hpvg.utils import images as uimg
    (hi, wi), (ho, wo) = case
    x = np.random.default_rng(1).standard_normal((2, 3, hi, wi)).astype(np.float32)
    y = uimg.interpolate(hp.from_numpy(x), size=[ho, wo])
    assert y.shape == (2, 3, ho, wo)
    ref = F.interpolate(torch.from_numpy(x), size=(ho, wo), mode="bilinear", align_corners=True).numpy()
    assert np.abs(y.numpy() - ref).max() < 2e-6
    ref2 = orc.resize_linear(torch.from_numpy(x), (ho, wo)).numpy()
    assert np.abs(y.numpy() - ref2).max() < 2e-6


def test_image_sample_matches_oracle_full_small_pyramid(hpvg_gpu):
    """eval_image.py:53-60 semantics: random-mode forward from Z_init noise, noise injected at EVERY scale."""
    hp = hpvg_gpu
    net, opt, oopt, params = _build(hp, 4)
    rng = np.random.default_rng(2)
    z = rng.standard_normal((2, 128) + orc.scale_shape_2d(oopt, 0)).astype(np.float32)
    amps = [1.0, 0.5, 0.4, 0.3, 0.2]
    nz = {s: rng.standard_normal((2, 3) + orc.scale_shape_2d(oopt, s)).astype(np.float32) for s in range(1, 5)}
    tz = hp.from_numpy(z)
    x, vae = net(tz, amps, noise_init=tz, isRandom=True, noises={k: hp.from_numpy(v) for k, v in nz.items()})
    assert x.shape == (2, 3) + orc.scale_shape_2d(oopt, 4) and vae.shape == (2, 3) + orc.scale_shape_2d(oopt, 0)
    with torch.no_grad():
        rx, rv = orc.generator_forward(None, amps, orc.to_torch(params), oopt, noise_init=torch.from_numpy(z),
                                       is_random=True, noises={k: torch.from_numpy(v) for k, v in nz.items()}, nd=2)
    assert rel_l2(vae.numpy(), rv.numpy()) < TOL
    assert rel_l2(x.numpy(), rx.numpy()) < TOL_E2E
    # the noise really is used at scale 1 (< vae_levels): without it the output changes
    x2, _ = net(tz, amps, noise_init=tz, isRandom=False)
    assert rel_l2(x2.numpy(), rx.numpy()) > 1e-2


def test_discriminator_and_encoder_2d(hpvg_gpu):
    hp = hpvg_gpu
    from hpvg import networks_2d as n2
    from hpvg.utils import images as uimg
    opt, oopt = uimg.default_opt(img_size=64), orc.default_opt(img_size=64)
    pd = orc.init_discriminator_params(oopt, seed=4, nd=2)
    D = n2.WDiscriminator2D(opt)
    D.load_parameters(pd)
    x = np.tanh(np.random.default_rng(5).standard_normal((1, 3, 40, 54))).astype(np.float32)
    td = orc.to_torch(pd)
    out = D(hp.from_numpy(x))
    with torch.no_grad():
        ref = orc.discriminator(torch.from_numpy(x), td, oopt)
    assert out.shape == (1, 1, 40, 54)
    assert rel_l2(out.numpy(), ref.numpy()) < TOL
    assert np.allclose(D.parameters_dict()["head.0.weight_u"].numpy(), td["head.0.weight_u"].numpy(), atol=1e-5)
    net, _, _, params = _build(hp, 0)
    tp = orc.to_torch(params)
    img = np.tanh(np.random.default_rng(6).standard_normal((1, 3) + orc.scale_shape_2d(oopt, 0))).astype(np.float32)
    mu, lv = net.encode(hp.from_numpy(img))
    with torch.no_grad():
        rmu, rlv = orc.encode(torch.from_numpy(img), tp, oopt)
    assert mu.shape == rmu.shape and rel_l2(mu.numpy(), rmu.numpy()) < TOL and rel_l2(lv.numpy(), rlv.numpy()) < TOL


def test_train_image_vae_step_loss_and_update(hpvg_gpu):
    """train_image.py VAE-phase iteration at scale 1: loss parity with the oracle and a finite ClippedAdam update."""
    hp = hpvg_gpu
    from hpvg import networks_2d as n2, train as T
    from hpvg.utils import images as uimg
    opt, oopt = uimg.default_opt(img_size=64), orc.default_opt(img_size=64)
    rng = np.random.default_rng(7)
    pg = orc.init_generator_params(oopt, 1, seed=7, nd=2)
    pd = orc.init_discriminator_params(oopt, seed=7, nd=2)
    G = n2.GeneratorHPVAEGAN(opt)
    G.init_next_stage()
    G.load_parameters(pg)
    D = n2.WDiscriminator2D(opt)
    D.load_parameters(pd)
    s0, s1 = orc.scale_shape_2d(oopt, 0), orc.scale_shape_2d(oopt, 1)
    real = np.tanh(rng.standard_normal((1, 3) + s1)).astype(np.float32)
    real_zero = np.tanh(rng.standard_normal((1, 3) + s0)).astype(np.float32)
    z = rng.standard_normal((1, 128) + s0).astype(np.float32)
    amps = [1.0, 0.0]
    tg = orc.to_torch(pg, requires_grad=("encode.", "decoder.", "body."))
    with orc.bf16_emulation():
        loss_ref = orc.g_loss(torch.from_numpy(real), torch.from_numpy(real_zero), None, amps, tg, None, oopt, True,
                              z_pred=torch.from_numpy(z), nd=2)
    loss_ref.backward()
    G.set_train(True)
    gl = T.GWithLoss(opt, D, G)
    loss, book = gl.grad(hp.from_numpy(real), hp.from_numpy(real_zero), None, amps, isVAE=True, trainable_body=(0,),
                         train_codec=True, z_pred=hp.from_numpy(z))
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    mine = G.parameters_dict()
    for k in ("encode._mu.0.weight", "encode._features.2.0.weight", "decoder.6.weight"):
        g, r = book.of(mine[k]).numpy(), tg[k].grad.numpy()
        assert g.shape == r.shape
        cos = float((g.ravel() @ r.ravel()) / (np.linalg.norm(g) * np.linalg.norm(r)))
        assert cos > 0.98, (k, cos)
    params = T.trainable_params(G)
    optim = T.ClippedAdam(opt, [{"params": params, "lr": opt.lr_g}], opt.lr_g, beta1=opt.beta1, beta2=0.999)
    before = mine["decoder.6.weight"].numpy().copy()
    optim.apply(book)
    after = mine["decoder.6.weight"].numpy()
    assert np.isfinite(after).all() and np.abs(after - before).max() > 0


def test_discriminator_2d_at_the_full_image_size(hpvg_gpu):
    """WDiscriminator2D on a 192 x 257 image (the finest scale of the default 256-pixel pyramid): with T == 1 the 198
    strips of the conv kernel are single planes, so every CTA pair's work range crosses strip boundaries."""
    hp = hpvg_gpu
    from hpvg import networks_2d as n2
    from hpvg.utils import images as uimg
    opt, oopt = uimg.default_opt(), orc.default_opt()
    pd = orc.init_discriminator_params(oopt, seed=14, nd=2)
    D = n2.WDiscriminator2D(opt)
    D.load_parameters(pd)
    x = np.tanh(np.random.default_rng(15).standard_normal((2, 3, 192, 257))).astype(np.float32)
    with torch.no_grad():
        ref = orc.discriminator(torch.from_numpy(x), orc.to_torch(pd), oopt)
    out = D(hp.from_numpy(x))
    assert out.shape == (2, 1, 192, 257)
    assert rel_l2(out.numpy(), ref.numpy()) < TOL
