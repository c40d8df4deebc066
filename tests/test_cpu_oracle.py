"""CPU suite: pins the oracle against every golden vector the reference tree holds for this path,
checks the oracle's internal consistency, the host-side geometry mirror, and that libhpvg.so loads and exports every
symbol declared in include/hpvg.h.  No GPU calls."""
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import hpvg_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


# ------------------------------------------------------------------------------------------- reference golden vectors
def test_trilinear_known_answer_from_reference_docstring():
    """src/tools/trilinear.py:222-233: 2x2 -> (2,4,4), align_corners=False."""
    x = np.arange(1, 5, dtype=np.float32).reshape(1, 1, 1, 2, 2)
    want_plane = np.array([[1.0, 1.25, 1.75, 2.0], [1.5, 1.75, 2.25, 2.5], [2.5, 2.75, 3.25, 3.5], [3.0, 3.25, 3.75, 4.0]],
                          np.float32)
    y = orc.resize_linear_np(x, (2, 4, 4), align_corners=False)
    assert y.shape == (1, 1, 2, 4, 4)
    assert np.array_equal(y[0, 0, 0], want_plane) and np.array_equal(y[0, 0, 1], want_plane)


def test_trilinear_shape_example_from_reference_docstring():
    """trilinear.py:217-220: (2,3,4,512,256) -> output_size [4,64,48]."""
    y = orc.resize_linear_np(np.zeros((2, 3, 4, 32, 16), np.float32), (4, 64, 48), True)
    assert y.shape == (2, 3, 4, 64, 48)


def test_geometry_value_from_reference_main_block():
    """src/utils/images.py:126 prints get_scales_by_index(3, 0.7937005259840998, 9, 256) == 65."""
    assert orc.get_scales_by_index(3, 0.7937005259840998, 9, 256) == 65


def test_default_pyramid():
    """SURVEY §8c golden (3): widths / time depths at the argparse defaults (images.py:64-93)."""
    opt = orc.default_opt()
    assert opt.stop_scale == 9
    assert abs(opt.scale_factor - 0.7937005259840998) < 1e-15
    shapes = [orc.scale_shape(opt, i) for i in range(10)]
    assert [s[2] for s in shapes] == [33, 41, 51, 65, 81, 102, 129, 162, 204, 257]
    assert [s[0] for s in shapes] == [4, 4, 4, 5, 5, 5, 7, 7, 7, 13]
    assert shapes[9] == (13, 192, 257) and shapes[0] == (4, 24, 33)


def test_smoke_block_shapes_networks_3d():
    """networks_3d.py:554-593: ones (8,3,4,2,2), one init_next_stage, sf .75, stop_scale 9, ar 1 ->
    outputs (8,3,4,26,26) and (8,3,4,2,2)."""
    opt = orc.default_opt()
    opt.scale_factor, opt.stop_scale, opt.img_size, opt.ar = 0.75, 9, 256, 1.0
    opt.stop_scale_time = 9
    p = orc.to_torch(orc.init_generator_params(opt, 1, seed=0))
    z = torch.ones(8, opt.latent_dim, 4, 2, 2)
    x, vae = orc.generator_forward(None, [1.0, 1.0], p, opt, noise_init=z, is_random=False)
    assert tuple(vae.shape) == (8, 3, 4, 2, 2)
    assert tuple(x.shape) == (8, 3, 4, 26, 26)


# ------------------------------------------------------------------------------------------- oracle self-consistency
@pytest.mark.parametrize("case", [((4, 24, 33), (4, 30, 41)), ((5, 9, 7), (7, 12, 10)), ((1, 8, 8), (1, 11, 13))])
def test_resize_matches_aten(case):
    (ti, hi, wi), size = case
    x = np.random.default_rng(0).standard_normal((2, 3, ti, hi, wi)).astype(np.float32)
    ref = F.interpolate(torch.from_numpy(x), size=size, mode="trilinear", align_corners=True).numpy()
    got = orc.resize_linear_np(x, size, True)
    assert np.max(np.abs(got - ref)) < 2e-6
    got_t = orc.resize_linear(torch.from_numpy(x), size, True).numpy()
    assert np.max(np.abs(got_t - got)) < 1e-6


def test_resize_taps_partition_of_unity_and_monotone():
    for n_in, n_out in [(33, 41), (4, 5), (7, 13), (192, 192), (5, 1), (1, 7)]:
        i0, i1, l0, l1 = orc.linear_taps(n_in, n_out, True)
        assert np.all(l0 + l1 == np.float32(1.0)) or np.max(np.abs(l0 + l1 - 1)) < 1e-7
        assert np.all(i0 >= 0) and np.all(i1 <= n_in - 1) and np.all(np.diff(i0) >= 0)
        assert i0[0] == 0 and (n_out == 1 or i1[-1] == n_in - 1)


def test_resize_backward_is_adjoint():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((1, 2, 4, 6, 5)).astype(np.float32)
    gy = rng.standard_normal((1, 2, 5, 8, 7)).astype(np.float32)
    y = orc.resize_linear_np(x, (5, 8, 7), True)
    gx = orc.resize_linear_bwd_np(gy, (4, 6, 5), True)
    assert abs(float((y.astype(np.float64) * gy).sum()) - float((x.astype(np.float64) * gx).sum())) < 1e-4
    xt = torch.from_numpy(x).requires_grad_(True)
    orc.resize_linear(xt, (5, 8, 7), True).backward(torch.from_numpy(gy))
    assert np.max(np.abs(xt.grad.numpy() - gx)) < 1e-5


def test_sn_sigma_matches_svd_after_iterations():
    rng = np.random.default_rng(2)
    w = torch.from_numpy(rng.standard_normal((64, 3, 3, 3, 3)).astype(np.float32))
    u = torch.from_numpy(rng.standard_normal((64, 1)).astype(np.float32))
    v = torch.from_numpy(rng.standard_normal((81, 1)).astype(np.float32))
    for _ in range(200):
        sigma, u, v = orc.sn_power_iteration(w, u, v)
    top = torch.linalg.svdvals(w.reshape(64, -1))[0]
    assert abs(float(sigma) - float(top)) / float(top) < 1e-3


def test_adam_and_clip_reference_formulae():
    g = np.full(100, 3.0, np.float32)          # ||g|| = 30 > 5 -> scaled to norm 5
    c = orc.clip_by_norm(g, 5.0)
    assert abs(np.linalg.norm(c) - 5.0) < 1e-5
    small = np.full(4, 0.1, np.float32)
    assert np.allclose(orc.clip_by_norm(small, 5.0), small)
    w, m, v = orc.adam_step(np.ones(3, np.float32), np.full(3, 0.5, np.float32), np.zeros(3, np.float32),
                            np.zeros(3, np.float32), 1, 1e-3, 0.5, 0.999)
    # step 1: m = .25, v = 2.5e-4, lr_t = 1e-3*sqrt(1-.999)/(1-.5)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / 0.5
    assert np.allclose(w, 1 - lr_t * 0.25 / (np.sqrt(2.5e-4) + 1e-8), rtol=1e-6)


def test_generator_eval_is_batch_independent():
    """eval-mode BN => samples are independent: batching must not change results."""
    opt = orc.default_opt(img_size=64)
    p = orc.to_torch(orc.randomize_bn_stats(orc.init_generator_params(opt, 2, seed=3), opt=opt))
    rng = np.random.default_rng(0)
    t0 = orc.scale_shape(opt, 0)
    z = torch.from_numpy(rng.standard_normal((2, 128) + t0).astype(np.float32))
    amps = [1.0, 0.5, 0.25]
    with torch.no_grad():
        both, _ = orc.generator_forward(None, amps, p, opt, noise_init=z)
        one, _ = orc.generator_forward(None, amps, p, opt, noise_init=z[1:2])
    assert torch.allclose(both[1:2], one, atol=1e-4)


# ------------------------------------------------------------------------------------------- committed golden fixtures
def test_oracle_matches_committed_golden_fixtures():
    """tests/golden/*.npz were generated by tests/golden/make_golden.py from this oracle + torch-CPU; they freeze
    the oracle's behaviour so that a silent change of the checker is caught."""
    path = os.path.join(GOLDEN, "sample_small.npz")
    g = np.load(path)
    opt = orc.default_opt(img_size=int(g["img_size"]))
    p = orc.to_torch(orc.randomize_bn_stats(orc.init_generator_params(opt, int(g["n_body"]), seed=int(g["seed"])), opt=opt))
    z = torch.from_numpy(g["z"])
    noises = {int(k[6:]): torch.from_numpy(g[k]) for k in g.files if k.startswith("noise_")}
    with torch.no_grad():
        x, vae = orc.generator_forward(None, list(g["amps"]), p, opt, noise_init=z, is_random=True, noises=noises)
    assert np.max(np.abs(x.numpy() - g["x"])) < 1e-4
    assert np.max(np.abs(vae.numpy() - g["vae"])) < 1e-4


# ------------------------------------------------------------------------------------------- host mirror + C ABI
def test_host_geometry_mirror_matches_oracle():
    from hpvg.utils import images as u
    for kw in ({}, {"img_size": 64}, {"img_size": 128, "ar": 0.5625}):
        a, b = orc.default_opt(**kw), u.default_opt(**kw)
        assert a.stop_scale == b.stop_scale and a.scale_factor == b.scale_factor
        for i in range(a.stop_scale + 1):
            assert orc.scale_shape(a, i) == u.scale_shape(b, i)
            assert orc.scale_shape_2d(a, i) == u.scale_shape_2d(b, i)


def test_library_loads_and_exports_every_declared_symbol():
    import hpvg
    hdr = open(os.path.join(ROOT, "include", "hpvg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hpvg_[a-z0-9_]+|Hpvg[A-Za-z0-9]+)\s*\(", hdr))
    assert len(declared) >= 45
    for name in sorted(declared):
        assert hasattr(hpvg.lib, name), "libhpvg.so does not export %s" % name
    assert hpvg.lib.hpvg_version() >= 100


def test_product_path_fails_loudly_without_gpu():
    import hpvg
    if hpvg.lib.hpvg_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(hpvg.HpvgError):
        hpvg.init(0)
    with pytest.raises(hpvg.HpvgError):
        hpvg.Tensor((4,), hpvg.F32)


@pytest.mark.parametrize("n_in,n_out", [(33, 41), (24, 30), (4, 5), (5, 7), (7, 13), (204, 257), (153, 192), (1, 4),
                                        (6, 1), (257, 257)])
@pytest.mark.parametrize("align", [True, False])
def test_resize_tables_bit_exact_host(n_in, n_out, align):
    """The bit-exact contract (north_star): indices and fp32 weights identical to the reference rule."""
    import hpvg
    i0, i1, l0, l1 = hpvg.ops.linear_taps(n_in, n_out, align)
    r0, r1, m0, m1 = orc.linear_taps(n_in, n_out, align)
    assert np.array_equal(i0, r0) and np.array_equal(i1, r1)
    assert np.array_equal(l0.view(np.uint32), m0.view(np.uint32))
    assert np.array_equal(l1.view(np.uint32), m1.view(np.uint32))


def test_c_abi_argument_errors():
    import ctypes
    import hpvg
    i0 = (ctypes.c_int32 * 4)()
    f0 = (ctypes.c_float * 4)()
    assert hpvg.lib.hpvg_linear_taps(0, 4, 1, i0, i0, f0, f0) == -2          # HPVG_E_ARG
    assert b"positive" in hpvg.lib.hpvg_last_error()
    assert hpvg.lib.hpvg_conv_wimg_bytes(0) == 2 * 27 * 32 * 128
    assert hpvg.lib.hpvg_conv_wimg_bytes(99) == 0


def test_oracle_reproduces_operator_fixtures():
    """tests/golden/ops_small.npz (made by tests/golden/make_golden.py) pins the per-operator oracle functions against
    drift: resize fwd/bwd + tap tables, BatchNorm(train), spectral-norm iteration, KL / MSE, ClipByNorm + Adam."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ops_small.npz"))
    assert np.array_equal(orc.resize_linear_np(g["rs_x"], (4, 30, 41), True), g["rs_y"])
    assert np.array_equal(orc.resize_linear_bwd_np(g["rs_gy"], (4, 24, 33), True), g["rs_gx"])
    i0, i1, l0, l1 = orc.linear_taps(33, 41, True)
    assert np.array_equal(i0, g["rs_i0"]) and np.array_equal(i1, g["rs_i1"])
    assert np.array_equal(l0.view(np.uint32), g["rs_l0"].view(np.uint32))
    assert np.array_equal(l1.view(np.uint32), g["rs_l1"].view(np.uint32))
    p = {"1.bn2d.gamma": torch.from_numpy(g["bn_gamma"]), "1.bn2d.beta": torch.from_numpy(g["bn_beta"]),
         "1.bn2d.moving_mean": torch.zeros(64), "1.bn2d.moving_variance": torch.ones(64)}
    out = orc.lrelu(orc.batchnorm(torch.from_numpy(g["bn_y"]), p, "1.", True)).numpy()
    assert np.allclose(out, g["bn_out"], atol=1e-6) and np.allclose(p["1.bn2d.moving_mean"].numpy(), g["bn_mm"], atol=1e-7)
    sigma, un, vn = orc.sn_power_iteration(torch.from_numpy(g["sn_w"]), torch.from_numpy(g["sn_u"]),
                                           torch.from_numpy(g["sn_v"]))
    assert abs(float(sigma) - float(g["sn_sigma"])) < 1e-7 and np.allclose(un.numpy(), g["sn_u1"], atol=1e-7)
    assert abs(float(orc.kl_criterion(torch.from_numpy(g["kl_mu"]), torch.from_numpy(g["kl_lv"]))) - float(g["kl"])) < 1e-6
    m = np.zeros_like(g["ad_w0"])
    v = np.zeros_like(g["ad_w0"])
    w1, m, v = orc.adam_step(g["ad_w0"], orc.clip_by_norm(g["ad_g1"], 5.0), m, v, 1, 5e-4)
    w2, m, v = orc.adam_step(w1, orc.clip_by_norm(g["ad_g2"], 5.0), m, v, 2, 5e-4)
    assert np.array_equal(w1, g["ad_w1"]) and np.array_equal(w2, g["ad_w2"])


# ------------------------------------------------------------------------------------------- data path
@pytest.mark.parametrize("src,dst", [((186, 248), (24, 33)), ((186, 248), (192, 257)), ((50, 60), (121, 162)),
                                     ((100, 120), (50, 60)), ((64, 64), (64, 64)), ((9, 7), (3, 20)),
                                     ((720, 1280), (153, 204))])
def test_cv_resize_restatement_matches_cv2(src, dst):
    """generate_frames.py:44-46 / image.py:72: the oracle's fixed-point bilinear is cv2.resize(INTER_LINEAR) on
    uint8, bit for bit — checked against the library the reference itself calls (cv2 is in this image)."""
    cv2 = pytest.importorskip("cv2")
    img = np.random.default_rng(src[0] * 1000 + dst[1]).integers(0, 256, src + (3,), dtype=np.uint8)
    assert np.array_equal(orc.cv_resize_linear_u8(img, dst), cv2.resize(img, (dst[1], dst[0]),
                                                                        interpolation=cv2.INTER_LINEAR))


def test_oracle_reproduces_data_path_fixtures():
    """tests/golden/frames_small.npz was produced by the reference's own library calls (cv2.cvtColor / cv2.resize /
    numpy, statement by statement as in generate_frames.py:42-47 and video.py:52-84) on seeded frames and on a crop
    of the reference's real image: the restatement must reproduce those bits."""
    g = np.load(os.path.join(GOLDEN, "frames_small.npz"))
    fr = g["frames_bgr"]
    assert np.array_equal(orc.frames_to_clip_np(fr, (24, 33), 0, 4, 4, False, True), g["clip_s0"])
    assert np.array_equal(orc.frames_to_clip_np(fr, (57, 76), 1, 3, 4, True, True), g["clip_up_flip"])
    img = g["image_bgr"][None]
    assert np.array_equal(orc.frames_to_clip_np(img, (24, 33), 0, 1, 1, False, True), g["image_s0"])
    assert np.array_equal(orc.frames_to_clip_np(img, (48, 64), 0, 1, 1, False, True), g["image_half"])
    assert g["clip_s0"].shape == (1, 3, 4, 24, 33) and float(np.abs(g["clip_s0"]).max()) <= 1.0


def test_frames_to_clip_argument_errors():
    import ctypes
    import hpvg
    buf = (ctypes.c_uint8 * 16)()
    out = (ctypes.c_float * 16)()
    # window runs past the decoded frames: 3 frames, start 1, every 2, T 2 -> needs frame 3
    assert hpvg.lib.hpvg_frames_to_clip(buf, 3, 2, 2, 0, 1, 2, 2, 2, 2, 0, out, None) == -2
    assert b"past the decoded frames" in hpvg.lib.hpvg_last_error()
    assert hpvg.lib.hpvg_frames_to_clip(buf, 3, 2, 2, 0, 0, 0, 2, 2, 2, 0, out, None) == -2      # every == 0


def test_sinfid_restatement_shapes_and_identities():
    """src/sinFID: block-0 shapes (c3d.py:62-66 keeps the clip's size at 64 channels; inception.py:66-72 halves and crops),
    position statistics (fid_score.py:160-178) and the Fréchet distance (fid_score.py:105-159)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 3, 4, 17, 21)).astype(np.float32)
    w, b = rng.standard_normal((64, 3, 3, 3, 3)).astype(np.float32) * 0.1, rng.standard_normal(64).astype(np.float32)
    f = orc.c3d_block0(x, w, b)
    assert tuple(f.shape) == (1, 64, 4, 17, 21)
    assert np.allclose(orc.c3d_block0((x + 1) / 2, w, b, normalize_input=True).numpy(), f.numpy(), atol=1e-4)
    params = {}
    for name, cin, cout in (("Conv2d_1a", 3, 32), ("Conv2d_2a", 32, 32), ("Conv2d_2b", 32, 64)):
        params[name + ".conv.weight"] = rng.standard_normal((cout, cin, 3, 3)).astype(np.float32) * 0.1
        params[name + ".bn.gamma"] = np.ones(cout, np.float32)
        params[name + ".bn.beta"] = np.zeros(cout, np.float32)
        params[name + ".bn.moving_mean"] = np.zeros(cout, np.float32)
        params[name + ".bn.moving_variance"] = np.ones(cout, np.float32)
    g = orc.inception_block0(rng.standard_normal((2, 3, 65, 81)).astype(np.float32), params)
    assert tuple(g.shape) == (2, 64, (65 - 3) // 2 + 1 - 2, (81 - 3) // 2 + 1 - 2) and float(g.min()) >= 0.0
    mu, sigma = orc.activation_statistics(f.numpy())
    assert mu.shape == (64,) and sigma.shape == (64, 64)
    act = f.numpy()[0].reshape(64, -1).T
    assert np.allclose(mu, act.mean(0)) and np.allclose(sigma, np.cov(act, rowvar=False))
    assert abs(orc.frechet_distance(mu, sigma, mu, sigma)) < 1e-3
    # 1-D Gaussians: d^2 = (m1-m2)^2 + (s1-s2)^2
    d = orc.frechet_distance(np.array([1.0]), np.array([[4.0]]), np.array([3.0]), np.array([[9.0]]))
    assert abs(d - (4.0 + 1.0)) < 1e-9
    v, per = orc.svfid(x, np.concatenate([x, x * 0.5]), lambda c: orc.c3d_block0(c, w, b).numpy())
    assert abs(per[0]) < 1e-3 and per[1] > 0 and abs(v - np.mean(per)) < 1e-6


def test_host_noise_is_counter_based_and_seed_folding_changes_the_device_key():
    from hpvg import sampling
    a = sampling.host_noise_for_sample(3, 17, (4, 5))
    buf = np.empty((4, 5), np.float32)
    assert sampling.host_noise_for_sample(3, 17, (4, 5), out=buf) is buf and np.array_equal(a, buf)
    assert a.dtype == np.float32 and not np.array_equal(a, sampling.host_noise_for_sample(4, 17, (4, 5)))

    class Net:
        noise_seed = 0x9E3779B97F4A7C15
    n = Net()
    sampling.fold_seed(n, 1)
    k1 = n.noise_seed
    sampling.fold_seed(n, 2)
    k2 = n.noise_seed
    sampling.fold_seed(n, 1)
    assert k1 != k2 and n.noise_seed == k1 and 0 <= k1 < 2 ** 63


def test_generator_description_layout_matches_the_c_struct():
    """include/hpvg.h HpvgGenerator / HpvgBlock against their ctypes mirrors (hpvg/_lib.py): the workspace size the C side
    computes from a description filled through ctypes must be the documented sum — which it only is when every field
    sits where the C struct has it (no compute, no GPU)."""
    import ctypes
    import hpvg
    from hpvg._lib import HpvgBlock, HpvgGenerator
    assert ctypes.sizeof(HpvgBlock) == 4 + 4 * 8 + 4 + 8 * 16 + 8 * 8 + 8 * 8      # int, int[8], pad, ptr[8][2], ptr[8], ptr[8]
    g = HpvgGenerator()
    g.n_stages, g.nc_im, g.latent_dim = 2, 3, 128
    shapes = [(4, 24, 33), (4, 30, 41), (4, 38, 51)]
    for l, (t, h, w) in enumerate(shapes):
        g.T[l], g.H[l], g.W[l] = t, h, w
    g.noise_amp[2] = 0.5
    g.noise_seed[2] = 0xFFFFFFFFFFFFFFFF
    g.body[1].n_layers = 5          # a field far into the struct
    N = 3
    al = lambda b: (b + 255) // 256 * 256                                       # noqa: E731
    v0, vmax = 4 * 24 * 33, 4 * 38 * 51
    want = (al(N * v0 * 128 * 2) + al(N * v0 * 64 * 4) + 2 * al(N * vmax * 64 * 2) + al(N * vmax * 3 * 4) +
            al(N * vmax * 16) + 2 * al(N * vmax * 3 * 4))
    assert hpvg.lib.hpvg_generator_sample_workspace(ctypes.byref(g), N) == want
    assert hpvg.lib.hpvg_generator_sample_workspace(ctypes.byref(g), 0) == 0
    g.n_stages = 16
    assert hpvg.lib.hpvg_generator_sample_workspace(ctypes.byref(g), N) == 0     # more levels than HPVG_MAX_LEVELS
    # argument errors are reported before any launch
    g.n_stages = 2
    assert hpvg.lib.hpvg_generator_sample(ctypes.byref(g), None, N, 0, None, None, None, None, 0, None) == -2
    blk = HpvgBlock()
    blk.cin[0] = 128
    assert hpvg.lib.hpvg_block_fwd_eval_workspace(ctypes.byref(blk), 2, 4, 24, 33) == 2 * al(2 * v0 * 64 * 2) + al(2 * v0 * 64 * 4)
